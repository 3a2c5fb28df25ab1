"""small workload that touches every kernel of gnb_small.cu plus the rewritten helper kernels (for compute-sanitizer)"""
import sys
import numpy as np
sys.path.insert(0, ".")
from gaunegf_b200 import synthetic as sy
from gaunegf_b200._native import Context
ctx = Context(0)
for N, nc in ((20, 3), (64, 5), (90, 11)):
    F, S = sy.hermitian_pair(N, seed=N)
    ctx.set_system(F, S)
    ctx.sigma_clear()
    for i in sy.end_contacts(N, nc):
        ctx.sigma_add_const_block(i, -0.1j * np.eye(nc))
    E = np.linspace(-1, 1, 7)
    for reg in (1, 0):
        ctx.lib.gnb_dev_set_option(b"small_reg", reg)
        T = ctx.transmission(E, 0, -1)
        d = ctx.dos(E)[0]
        G = ctx.green(E[:2] + 0.1j)
        P = ctx.gr_int(E + 0.2j, np.ones(7) / 7)
    ctx.lib.gnb_dev_set_option(b"small_reg", 1)
    print(N, T[:2], d[:1])
# block engine kernels rewritten this round (assemble, panel_save, panel_prep) at a padded size
N, nc = 200, 12
F, S = sy.hermitian_pair(N, seed=1)
ctx.set_system(F, S)
ctx.sigma_clear()
for i in sy.end_contacts(N, nc):
    ctx.sigma_add_const_block(i, -0.1j * np.eye(nc))
E = np.linspace(-1, 1, 5)
print(ctx.transmission(E, 0, -1)[:2], ctx.gr_int(E + 0.2j, np.ones(5) / 5)[0, :1])
