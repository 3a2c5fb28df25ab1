"""cfg4 at full size (dev probe): surfG1D Sigma(E), N = 512 device + 128-orbital lead cells, 256 energies."""
import sys, time, io, contextlib
import numpy as np
sys.path.insert(0, ".")
from gaunegf_b200 import synthetic as sy, transport as tr
from gaunegf_b200.surfG1D import surfG
eta = float(sys.argv[1]) if len(sys.argv) > 1 else 1e-4
M = int(sys.argv[2]) if len(sys.argv) > 2 else 256
F, S, inds, taus = sy.lead_device_lead(128, 512, seed=2, s_off=0.0)
g = surfG(F, S, [list(i) for i in inds], [list(t) for t in taus], eta=eta)
E = np.linspace(-1, 1, M)
for rep in range(2):
    t = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        T = tr.cohTransE(E, F, S, g)
    dt = time.perf_counter() - t
    print(f"eta={eta} M={M}: {dt:.3f} s -> {M/dt:.1f} E/s; T[:3]={np.array(T[:3])}", flush=True)
t = time.perf_counter(); g0 = g.g(E, 0); dt = time.perf_counter() - t
its = np.array([g.last_iters[(complex(e), 0)][0] for e in E])
print(f"g.g(E, 0) alone: {dt:.3f} s; iterations min/median/max {its.min()}/{int(np.median(its))}/{its.max()}")
