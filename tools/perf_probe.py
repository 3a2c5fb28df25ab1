"""Quick device-side throughput probe of the engine (dev tool)."""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
from gaunegf_b200 import synthetic as sy
from gaunegf_b200._native import Context
import torch

def ev(f, n=2):
    f(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(n):
        t = time.perf_counter(); f(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t)
    return best

ctx = Context(0)
import ctypes, os
for opt in ("two_level", "gemm_pipe", "engine_rec", "rec_streams", "tourn_fp32", "rk_m3", "rk_m3_mink", "rk_kskip"):
    if opt.upper() in os.environ:
        ctx.lib.gnb_dev_set_option(opt.encode(), int(os.environ[opt.upper()]))
ctx.set_timing(True)
out = {}
for N, nc, M in ((256, 16, 1184), (512, 32, 592), (1024, 64, 296), (2048, 64, 74)):
    F, S = sy.hermitian_pair(N, seed=1)
    s1, s2 = sy.block_sigma_vectors(N, nc, 0.1)
    ctx.set_system(F, S); ctx.sigma_clear()
    ctx.sigma_add_const_block(np.arange(nc), np.diag(s1[:nc]))
    ctx.sigma_add_const_block(np.arange(N - nc, N), np.diag(s2[N - nc:]))
    E = np.linspace(-0.5, 0.5, M)
    z, w = sy.contour_points(2 * (M // 2), -30.0, 0.0)
    t = ev(lambda: ctx.transmission(E)); te = ctx.last_elim_ms
    out[f"T_N{N}"] = dict(M=M, s=t, eps=M / t, elim_ms=te, tflops=(8 / 3 * N**3 + 8 * N * N * nc) * M / (te * 1e-3) / 1e12)
    t = ev(lambda: ctx.gr_int(z, w)); te = ctx.last_elim_ms
    out[f"G_N{N}"] = dict(M=len(z), s=t, eps=len(z) / t, elim_ms=te, tflops=8 * N**3 * len(z) / (te * 1e-3) / 1e12)
    print(json.dumps({k: v for k, v in out.items() if k.endswith(f"N{N}")}), flush=True)
json.dump(out, open("gpurun_out/perf_probe.json", "w"), indent=1)
