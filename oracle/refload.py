"""Load the UNMODIFIED reference package (wliverno/GauNEGF) under the numpy-backed jax shim.

TEST INFRASTRUCTURE ONLY — used by tests/golden/make_golden.py (to generate committed golden
vectors in the build container) and by the `-m "not gpu"` tests that pin oracle/negf_oracle.py
against the reference when /root/reference is present.  /root/reference does not exist on the
GPU box; nothing on the product path or in the `-m gpu` tests calls this.
"""
import importlib
import os
import sys

SHIM_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refshim")


def reference_available(path="/root/reference"):
    return os.path.isdir(os.path.join(path, "gauNEGF"))


def load_reference(path="/root/reference"):
    """Returns a dict of reference modules. Real jax/matplotlib/gauopen win if installed."""
    if not reference_available(path):
        raise RuntimeError(f"reference not found at {path}")
    for mod in ("jax", "matplotlib", "gauopen"):
        try:
            if mod not in sys.modules:
                importlib.import_module(mod)
        except Exception:
            if SHIM_DIR not in sys.path:
                sys.path.insert(0, SHIM_DIR)
            importlib.import_module(mod)
    if path not in sys.path:
        sys.path.insert(0, path)
    cfg = importlib.import_module("gauNEGF.config")
    cfg.LOG_PERFORMANCE = False  # integrate.py:28-32 would create a log file in CWD
    cfg.LOG_LEVEL = "CRITICAL"
    names = ["config", "utils", "integrate", "transport", "density", "surfG1D", "surfGBethe",
             "surfGTester", "matTools"]
    return {n: importlib.import_module(f"gauNEGF.{n}") for n in names}
