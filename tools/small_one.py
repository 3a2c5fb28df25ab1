"""one T(E) call on the small path (for ncu): python tools/small_one.py N nc M"""
import sys
import numpy as np
sys.path.insert(0, ".")
from gaunegf_b200 import synthetic as sy
from gaunegf_b200._native import Context
N, nc, M = (int(x) for x in sys.argv[1:4])
ctx = Context(0)
F, S = sy.hermitian_pair(N, seed=N)
ctx.set_system(F, S)
ctx.sigma_clear()
for i in sy.end_contacts(N, nc):
    ctx.sigma_add_const_block(i, -0.1j * np.eye(nc))
E = np.linspace(-3, 3, M)
for _ in range(2):
    T = ctx.transmission(E, 0, -1)
print(T[:3])
