import sys, time, os
sys.path.insert(0, ".")
import numpy as np
from gaunegf_b200 import synthetic as sy
from gaunegf_b200._native import Context
ctx = Context(0)
print("cpus", os.cpu_count(), len(os.sched_getaffinity(0)))
for N in (1024, 2048):
    F, S = sy.hermitian_pair(N, seed=1)
    Fc, Sc = F.astype(complex), S.astype(complex)
    for name, (a, b) in (("real", (F, S)), ("complex", (Fc, Sc))):
        ctx.set_system(a, b)
        t = time.perf_counter()
        for _ in range(10): ctx.set_system(a, b)
        dt = (time.perf_counter() - t) / 10
        a2 = a.copy()
        t = time.perf_counter()
        for k in range(10):
            a2[0, 0] += 1e-9
            ctx.set_system(a2, b)
        dt2 = (time.perf_counter() - t) / 10
        print(N, name, "same %.2f ms  F changed %.2f ms" % (dt * 1e3, dt2 * 1e3))
