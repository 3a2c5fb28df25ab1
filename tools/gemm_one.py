"""one configuration of the rank-K kernel (for ncu): python tools/gemm_one.py n M k bm"""
import ctypes as C, sys
sys.path.insert(0, ".")
from gaunegf_b200._native import Context
n, M, k, bm = (int(x) for x in sys.argv[1:5])
ctx = Context(0)
fn = ctx.lib.gnb_dev_gemm_bench
fn.restype = C.c_int
fn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
ms = C.c_double()
assert fn(ctx.h, M, n, k, bm, 3, C.byref(ms)) == 0
print(n, M, k, bm, ms.value, 8.0 * n * n * k * M / (ms.value * 1e-3) / 1e12)
