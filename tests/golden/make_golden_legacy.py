"""Golden vectors for the rest of the public API surface (legacy wrappers and option branches not covered by
make_golden.py), produced by the UNMODIFIED reference under the numpy-backed jax shim.

Run in the build container only:   python tests/golden/make_golden_legacy.py   ->  tests/golden/legacy_api.npz
"""
import contextlib
import io
import os
import sys
import tempfile

import numpy as np
import scipy.io as sio

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.refload import load_reference  # noqa: E402
from gaunegf_b200 import synthetic as sy    # noqa: E402
sys.path.insert(0, os.path.join(ROOT, "tests"))
from legacy_cases import run_cases           # noqa: E402

R = load_reference()
out = run_cases(R["transport"], R["density"], R["surfGTester"].surfGTest, sio, tempfile)
path = os.path.join(HERE, "legacy_api.npz")
np.savez_compressed(path, **out)
print(f"legacy_api: {os.path.getsize(path) / 1024:.1f} KiB")
for k, v in out.items():
    print(k, np.shape(v), np.ravel(v)[:3])
