"""96 < N <= 192 on the energy-grid calls: thread-block-cluster kernels (one launch: assembly + inverse) vs the lock-step
block engine, by batch size (dev tool; writes gpurun_out/small_probe_r2.json)."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from gaunegf_b200 import synthetic as sy          # noqa: E402
from gaunegf_b200._native import Context          # noqa: E402

ctx = Context(0)
ctx.lib.gnb_dev_set_option(b"small_cluster_maxm", 1 << 20)       # the probe decides, not the dispatch


def timed(f, reps=7):
    f()
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        f()
        ts.append(time.perf_counter() - t)
    return float(np.median(ts))


out = []
for N in (112, 128, 160, 192):
    nc = N // 8
    F, S = sy.hermitian_pair(N, seed=N)
    ctx.set_system(F, S)
    ctx.sigma_clear()
    for i in sy.end_contacts(N, nc):
        ctx.sigma_add_const_block(i, -0.1j * np.eye(nc))
    for M in (2, 12, 36, 108, 324, 1024, 4096):
        E = np.linspace(-3, 3, M)
        z = E + 0.05j
        w = np.full(M, 1.0 / M, dtype=complex)
        row = {"N": N, "M": M}
        for cl, tag in ((1, "cluster"), (0, "block")):
            ctx.lib.gnb_dev_set_option(b"small_cluster", cl)
            row[f"GrInt_{tag}_ms"] = 1e3 * timed(lambda: ctx.gr_int(z, w))
            row[f"DOS_{tag}_ms"] = 1e3 * timed(lambda: ctx.dos(E))
        ctx.lib.gnb_dev_set_option(b"small_cluster", 1)
        out.append(row)
        print(json.dumps(row), flush=True)
json.dump(out, open("gpurun_out/small_probe_r2.json", "w"), indent=1)
