import ctypes as C, sys, os
sys.path.insert(0, ".")
from gaunegf_b200 import _native
for v in range(4):
    lib = _native.load_library(os.path.join("gaunegf_b200/_lib", f"libgnb_v{v}.so"))
    h = C.c_void_p(); assert lib.gnb_create(C.byref(h), 0) == 0
    fn = lib.gnb_dev_gemm_bench; fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
    for k in (32, 64):
        ms = C.c_double(); assert fn(h, 296, 1024, k, 0, 5, C.byref(ms)) == 0
        print(f"variant {v} k={k}: {ms.value:.3f} ms {8.0*1024*1024*k*296/(ms.value*1e-3)/1e12:.2f} TF/s", flush=True)
    lib.gnb_destroy(h)
