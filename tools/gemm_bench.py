"""Rank-K update kernel alone (dev tool): TFLOP/s vs K and CTA shape."""
import ctypes as C, json, sys
sys.path.insert(0, ".")
from gaunegf_b200._native import Context
ctx = Context(0)
fn = ctx.lib.gnb_dev_gemm_bench
fn.restype = C.c_int
fn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
out = {}
for n, M in ((1024, 296), (512, 1184)):
    for k in (32, 64):
        for bm in (64, 32, 0):
            ms = C.c_double()
            rc = fn(ctx.h, M, n, k, bm, 5, C.byref(ms))
            assert rc == 0, ctx.lib.gnb_last_error(ctx.h)
            tf = 8.0 * n * n * k * M / (ms.value * 1e-3) / 1e12
            out[f"n{n}_k{k}_bm{bm}"] = round(tf, 2)
            print(f"n={n} M={M} k={k} bm={bm}: {ms.value:.3f} ms  {tf:.2f} TFLOP/s", flush=True)
json.dump(out, open("gpurun_out/gemm_bench.json", "w"), indent=1)
