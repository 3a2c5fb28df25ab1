"""Golden vectors for SURVEY.md §8(f) N2 — grid fits and Fermi-level searches (gauNEGF/density.py:836-1515) —
produced by the UNMODIFIED reference under the numpy-backed jax shim (oracle/refshim).

Run in the build container only:   python tests/golden/make_golden_n2.py     ->  tests/golden/n2_fermi.npz
Inputs are rebuilt from seeds by tests (gaunegf_b200.synthetic), only the reference's answers are stored.
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.refload import load_reference  # noqa: E402
from gaunegf_b200 import synthetic as sy    # noqa: E402

R = load_reference()
de, sgt, s1d = R["density"], R["surfGTester"], R["surfG1D"]

N, NC, NE, EMIN = 24, 4, 10, -20.0          # the system every case below uses (tests rebuild it)


def system():
    F, S = sy.hermitian_pair(N, seed=11)
    return F, S, sgt.surfGTest(F, S, sy.end_contacts(N, NC), -0.1j, -0.1j)


def quiet(f, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return f(*a, **k)


def scal(t):
    """tuple of scalars / None / matrices -> float vector (None -> nan, matrices dropped)"""
    return np.array([np.nan if x is None else float(np.real(x)) for x in t if np.ndim(x) == 0 or x is None])


out = {}
F, S, g = system()
out["calcEmin"] = np.array([quiet(de.calcEmin, F, S, g)])
out["integralFit"] = scal(quiet(de.integralFit, F, S, g, 0.0))
out["integralFitNEGF"] = np.array([quiet(de.integralFitNEGF, F, S, g, 0.0, 0.2)])
for name in ("calcFermiBisect", "calcFermiSecant", "calcFermiMuller", "calcFermiPolyFit"):
    for tag, (Ef0, npts) in {"a": (0.0, 24), "b": (0.4, 12)}.items():
        F, S, g = system()
        r = quiet(getattr(de, name), g, NE, EMIN, Ef0, npts)
        out[f"{name}_{tag}"] = scal(r)
        out[f"{name}_{tag}_P"] = np.asarray(r[2])
F, S, g = system()
out["calcFermi"] = scal(quiet(de.calcFermi, g, NE, EMIN, 5.0, 0.0, 24, 16))
F, S, g = system()
out["getFermiContact"] = np.array([quiet(de.getFermiContact, g, NE)])

# 1-D chain contact (getFermi1DContact): lead cell of 6 orbitals, 3 electrons per cell
Fc, Sc, li, taus = sy.lead_device_lead(6, 12, seed=4, s_off=0.03)
gs = s1d.surfG(Fc, Sc, [list(i) for i in li], [list(t) for t in taus], eta=1e-4)
out["getFermi1DContact"] = scal(quiet(de.getFermi1DContact, gs, 3, 0, 1e-3, -1e6, 0.0, 30))

# N4: energy-independent analytic density + bisectFermi (density.py:276-382)
V4, Vc4, D4, Gam4 = sy.analytic_density_case(30, seed=5)
out["n4_density"] = de.density(V4, Vc4, D4, Gam4, -50.0, 0.1)
out["n4_bisectFermi"] = np.array([quiet(de.bisectFermi, V4, Vc4, D4, Gam4, 12.0)])

path = os.path.join(HERE, "n2_fermi.npz")
np.savez_compressed(path, **out)
print(f"n2_fermi: {os.path.getsize(path) / 1024:.1f} KiB")
for k, v in out.items():
    if not k.endswith("_P"):
        print(k, v)
