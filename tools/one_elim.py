"""one elimination workload (for ncu launch lists): python tools/one_elim.py N M T|G"""
import sys
import numpy as np
sys.path.insert(0, ".")
from gaunegf_b200 import synthetic as sy
from gaunegf_b200._native import Context
N, M, mode = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
nc = 64 if N >= 1024 else max(N // 16, 2)
ctx = Context(0)
F, S = sy.hermitian_pair(N, seed=1)
s1, s2 = sy.block_sigma_vectors(N, nc, 0.1)
ctx.set_system(F, S); ctx.sigma_clear()
ctx.sigma_add_const_block(np.arange(nc), np.diag(s1[:nc]))
ctx.sigma_add_const_block(np.arange(N - nc, N), np.diag(s2[N - nc:]))
if mode == "T":
    print(ctx.transmission(np.linspace(-0.5, 0.5, M))[:2])
else:
    z, w = sy.contour_points(2 * (M // 2), -30.0, 0.0)
    print(ctx.gr_int(z, w)[0, :2])
