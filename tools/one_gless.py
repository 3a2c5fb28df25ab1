"""one GrLessInt workload (for ncu launch lists): python tools/one_gless.py N M nc"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from gaunegf_b200 import synthetic as sy
from gaunegf_b200._native import Context
N, M, nc = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
ctx = Context(0)
F, S = sy.hermitian_pair(N, seed=3)
ctx.set_system(F, S); ctx.sigma_clear()
rng = np.random.default_rng(1)
for inds in (np.arange(nc), np.arange(N - nc, N)):
    b = rng.standard_normal((nc, nc)) * 0.02
    ctx.sigma_add_const_block(inds, (b + b.T) / 2 - 0.1j * np.eye(nc))
E = np.linspace(-0.25, 0.25, M)
w = np.full(M, 0.5 / M)
if len(sys.argv) > 4:
    ctx.gless_int(E, w, -1)
    t = time.perf_counter(); ctx.gless_int(E, w, -1); print("ms", (time.perf_counter() - t) * 1e3)
    t = time.perf_counter(); ctx.transmission(E, 0, -1); print("T ms (cold)", (time.perf_counter() - t) * 1e3)
    t = time.perf_counter(); ctx.transmission(E, 0, -1); print("T ms", (time.perf_counter() - t) * 1e3)
else:
    print(ctx.gless_int(E, w, -1)[0, :2])
