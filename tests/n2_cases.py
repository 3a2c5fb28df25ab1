"""The §8(f) N2 cases of tests/golden/make_golden_n2.py, run on a given density module (shared by the CPU
host-logic test, which substitutes the numpy oracle for the GPU drivers, and the GPU parity test)."""
import contextlib
import io

import numpy as np

from gaunegf_b200 import synthetic as sy

N, NC, NE, EMIN = 24, 4, 10, -20.0


def quiet(f, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return f(*a, **k)


def scal(t):
    return np.array([np.nan if x is None else float(np.real(x)) for x in t if np.ndim(x) == 0 or x is None])


def run_cases(de, surfGTest, surfG):
    def system():
        F, S = sy.hermitian_pair(N, seed=11)
        return F, S, surfGTest(F, S, sy.end_contacts(N, NC), -0.1j, -0.1j)

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
    if surfG is not None:
        Fc, Sc, li, taus = sy.lead_device_lead(6, 12, seed=4, s_off=0.03)
        gs = surfG(Fc, Sc, [list(i) for i in li], [list(t) for t in taus], eta=1e-4)
        out["getFermi1DContact"] = scal(quiet(de.getFermi1DContact, gs, 3, 0, 1e-3, -1e6, 0.0, 30))
    return out


def compare(out, gold, tol):
    """searches amplify rounding differences of the integrals by 1/DOS; counts and brackets are discrete"""
    for k, v in out.items():
        ref = gold[k]
        assert v.shape == ref.shape, k
        assert np.array_equal(np.isnan(v), np.isnan(ref)), k
        m = ~np.isnan(ref) if ref.dtype.kind == "f" else np.ones(ref.shape, bool)
        err = np.max(np.abs(v[m] - ref[m])) if m.any() else 0.0
        assert err <= tol * max(1.0, np.max(np.abs(ref[m]))), (k, err, v, ref)
