"""gaunegf_b200 — B200-native energy-grid Green's-function path behind GauNEGF's Python API.

    from gaunegf_b200 import transport, density, integrate
    from gaunegf_b200.surfG1D import surfG
    from gaunegf_b200.surfGBethe import surfGB, surfGBAt
    from gaunegf_b200.surfGTester import surfGTest

numpy in, numpy out; all arithmetic of the path runs in libgaunegf_b200.so (sm_100a CUDA, C ABI in
include/gaunegf_b200.h).  There is no CPU fallback: calls raise without the library or a GPU.
"""
__version__ = "0.1.0"
