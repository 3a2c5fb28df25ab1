"""A/B sweep of developer options on the bench workload (dev tool, one gpurun call for many variants):
    python tools/sweep.py "name1:opt=a,opt=b" "name2:..."        (the empty option list is the default build)
Each variant runs in a fresh process (the switches are process-wide): T(E), N = 1024, 1250 energies, 64-orbital contacts;
prints energy points / s (best of 4 synchronous calls) and appends to gpurun_out/sweep.json.
GNB_SWEEP_MODE=G runs GrInt (296 points) instead; GNB_SWEEP_CPLX=1 uses a complex F."""
import json
import os
import subprocess
import sys
import time

CHILD = r'''
import os, sys, time
import numpy as np
sys.path.insert(0, ".")
from gaunegf_b200 import synthetic as sy
from gaunegf_b200._native import Context
N, nc = int(os.environ.get("GNB_SWEEP_N", "1024")), int(os.environ.get("GNB_SWEEP_NC", "64"))
M = int(os.environ.get("GNB_SWEEP_M", "1250"))
ctx = Context(0)
F, S = sy.hermitian_pair(N, seed=1, complex_F=os.environ.get("GNB_SWEEP_CPLX") == "1")
s1, s2 = sy.block_sigma_vectors(N, nc, 0.1)
ctx.set_system(F, S); ctx.sigma_clear()
ctx.sigma_add_const_block(np.arange(nc), np.diag(s1[:nc]))
ctx.sigma_add_const_block(np.arange(N - nc, N), np.diag(s2[N - nc:]))
E = np.linspace(-0.5, 0.5, M)
if os.environ.get("GNB_SWEEP_MODE") == "G":
    z, w = sy.contour_points(296, -30.0, 0.0)
    run, npts = (lambda: ctx.gr_int(z, w)), 296
else:
    run, npts = (lambda: ctx.transmission(E)), M
ref = run()
best = 1e30
for _ in range(4):
    t = time.perf_counter(); out = run(); best = min(best, time.perf_counter() - t)
chk = float(np.sum(np.abs(out)))
print("RESULT", npts / best, best * 1e3, chk, ctx.last_elim_flops)
'''

results = {}
for spec in sys.argv[1:]:
    name, _, opts = spec.partition(":")
    env = dict(os.environ)
    extra = []
    for kv in filter(None, opts.split(",")):
        if kv.startswith("WS="):
            env["GNB_WS_GIB"] = kv[3:]
        else:
            extra.append(kv)
    env["GNB_DEV_OPTS"] = ",".join(extra)
    t = time.time()
    res = subprocess.run([sys.executable, "-c", CHILD], capture_output=True, text=True, env=env, timeout=600)
    line = [l for l in res.stdout.splitlines() if l.startswith("RESULT")]
    if not line:
        print(f"{name:28s} FAILED rc={res.returncode} {res.stderr[-300:]}", flush=True)
        results[name] = None
        continue
    _, eps, ms, chk, fl = line[0].split()
    results[name] = {"opts": opts, "points_per_s": float(eps), "ms": float(ms), "checksum": float(chk), "elim_flops": float(fl)}
    print(f"{name:28s} {float(eps):10.0f} pts/s  {float(ms):8.2f} ms  checksum {float(chk):.12e}  ({time.time() - t:.0f}s)", flush=True)
path = os.path.join("gpurun_out", "sweep.json")
old = json.load(open(path)) if os.path.exists(path) else {}
old.update(results)
json.dump(old, open(path, "w"), indent=1)
