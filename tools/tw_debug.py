"""dev: T(E) real-structure path vs the oracle for several sizes, warp tournament on/off"""
import sys
import numpy as np
sys.path.insert(0, ".")
from gaunegf_b200 import synthetic as sy
from gaunegf_b200._native import Context
from oracle import negf_oracle as O
ctx = Context(0)
for N, nc in ((128, 8), (160, 8), (192, 16), (256, 16), (320, 16), (384, 16), (512, 32), (1024, 64)):
    F, S = sy.hermitian_pair(N, seed=1)
    s1, s2 = sy.block_sigma_vectors(N, nc, 0.1)
    E = np.linspace(-0.5, 0.5, 5)
    Tref = O.calculate_transmission(F, S, O.SigmaCalculator(s1, s2), E)
    out = []
    for tw, f32 in ((0, 1), (3, 1), (1, 1), (2, 1), (1, 0)):
        ctx.lib.gnb_dev_set_option(b"tourn_warp", tw)
        ctx.lib.gnb_dev_set_option(b"tourn_fp32", f32)
        ctx.set_system(F, S); ctx.sigma_clear()
        ctx.sigma_add_const_block(np.arange(nc), np.diag(s1[:nc]))
        ctx.sigma_add_const_block(np.arange(N - nc, N), np.diag(s2[N - nc:]))
        T = ctx.transmission(E)
        out.append(float(np.max(np.abs(T - Tref)) / np.max(np.abs(Tref))))
    print(N, nc, " ".join("%.2e" % v for v in out), flush=True)
