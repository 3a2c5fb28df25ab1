"""Per-launch event trace of one elimination (dev tool): python tools/trace_elim.py N M T|G [streams]"""
import sys, os, ctypes as C, collections
import numpy as np
sys.path.insert(0, ".")
from gaunegf_b200 import synthetic as sy
from gaunegf_b200._native import Context
N, M, mode = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
streams = int(sys.argv[4]) if len(sys.argv) > 4 else 4
nc = 64 if N >= 1024 else max(N // 16, 2)
ctx = Context(0)
ctx.lib.gnb_dev_set_option(b"rec_streams", streams)
for opt in ("rk_m3", "tourn_fp32"):
    if opt.upper() in os.environ:
        ctx.lib.gnb_dev_set_option(opt.encode(), int(os.environ[opt.upper()]))
F, S = sy.hermitian_pair(N, seed=1)
s1, s2 = sy.block_sigma_vectors(N, nc, 0.1)
ctx.set_system(F, S); ctx.sigma_clear()
ctx.sigma_add_const_block(np.arange(nc), np.diag(s1[:nc]))
ctx.sigma_add_const_block(np.arange(N - nc, N), np.diag(s2[N - nc:]))
E = np.linspace(-0.5, 0.5, M)
z, w = sy.contour_points(2 * (M // 2), -30.0, 0.0)
run = (lambda: ctx.transmission(E)) if mode == "T" else (lambda: ctx.gr_int(z, w))
run()
ctx.lib.gnb_dev_trace_start()
run()
path = f"gpurun_out/trace_{mode}{N}_s{streams}.txt"
ctx.lib.gnb_dev_trace_dump(path.encode())
rows = [l.split() for l in open(path)]
t0 = min(float(r[3]) for r in rows); t1 = max(float(r[4]) for r in rows)
print(f"span {t1 - t0:.3f} ms, {len(rows)} scopes")
by = collections.defaultdict(float)
for r in rows: by[r[0]] += float(r[4]) - float(r[3])
for k, v in sorted(by.items(), key=lambda kv: -kv[1]): print(f"  {k:10s} sum of scope durations {v:8.3f} ms")
# union of gemm intervals (tensor pipe busy estimate)
iv = sorted((float(r[3]), float(r[4])) for r in rows if r[0].startswith("gemm"))
busy, cur0, cur1 = 0.0, None, None
for a, b in iv:
    if cur1 is None or a > cur1:
        if cur1 is not None: busy += cur1 - cur0
        cur0, cur1 = a, b
    else: cur1 = max(cur1, b)
if cur1 is not None: busy += cur1 - cur0
print(f"  union of gemm intervals {busy:.3f} ms = {100 * busy / (t1 - t0):.1f}% of span")
