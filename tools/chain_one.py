"""one surfG1D fixed-point solve (for ncu launch lists): python tools/chain_one.py n_lead M eta"""
import sys
import time
import numpy as np
sys.path.insert(0, ".")
from gaunegf_b200 import synthetic as sy
from gaunegf_b200.surfG1D import surfG
nl, M, eta = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
F, S, li, taus = sy.lead_device_lead(nl, 4 * nl, seed=2, s_off=0.0)
g = surfG(F, S, [list(i) for i in li], [list(t) for t in taus], eta=eta)
E = np.linspace(-1, 1, M)
if len(sys.argv) < 5:
    g.g(E[:4], 0)          # warm-up (skipped with a 4th argument, for ncu launch lists)
t = time.perf_counter()
g0 = g.g(E, 0)
dt = time.perf_counter() - t
its = np.array([g.last_iters[(complex(e), 0)][0] for e in E])
print(f"n_lead={nl} M={M} eta={eta}: {dt:.3f} s, iterations min/med/max {its.min()}/{int(np.median(its))}/{its.max()}, "
      f"{dt / its.max() * 1e6:.1f} us per lock-step iteration")
