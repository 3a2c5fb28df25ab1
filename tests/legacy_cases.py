"""Cases of tests/golden/make_golden_legacy.py, run on given transport / density modules (the reference's when the
goldens are made, gaunegf_b200's in the GPU test)."""
import contextlib
import io
import os

import numpy as np

from gaunegf_b200 import synthetic as sy


def quiet(f, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return f(*a, **k)


def run_cases(tr, de, surfGTest, sio, tempfile):
    out = {}
    N, nc = 20, 3
    F, S = sy.hermitian_pair(N, seed=31)
    inds = sy.end_contacts(N, nc)
    g = surfGTest(F, S, inds, -0.1j, -0.15j)
    s1, s2 = sy.block_sigma_vectors(N, nc, 0.1)
    # spin-resolved current / transmission on a 2N x 2N collinear system with N x N self-energies (kron expansion)
    F2, S2 = sy.hermitian_pair(2 * N, seed=32, complex_F=True)
    out["currentSpin_u"] = np.asarray(quiet(tr.currentSpin, F2, S2, s1, s2, 0.0, 0.3, 0.0, "u", 0.02))
    out["current_T300"] = np.asarray(quiet(tr.current, F, S, s1, s2, 0.1, 0.2, 300.0, "r", 0.02))
    out["currentE"] = np.asarray(quiet(tr.currentE, F, S, g, 0.0, 0.3, 0.0, "r", 0.02))
    E = np.linspace(-0.6, 0.7, 7)
    Tt, T4 = quiet(tr.cohTransSpinE, E, np.kron(np.eye(2), F), np.kron(np.eye(2), S), g, "u")
    out["cohTransSpinE_tot"], out["cohTransSpinE_4"] = np.asarray(Tt), np.asarray(T4)
    # currentF: inputs through a .mat file (transport.py:847-875)
    m1, m2 = np.diag(s1).astype(complex), np.diag(s2).astype(complex)
    with tempfile.TemporaryDirectory() as d:
        fn = os.path.join(d, "junction.mat")
        sio.savemat(fn, {"F": F, "S": S, "sig1": m1, "sig2": m2, "fermi": 0.05, "qV": 0.25, "spin": "r"})
        out["currentF"] = np.asarray(quiet(tr.currentF, fn, 0.02, 0.0))
    # density option branches
    out["densityGridTrap"] = np.asarray(quiet(de.densityGridTrap, F, S, g, -0.2, 0.3, 0, 12, 0.0))
    out["densityGridN_T300"] = np.asarray(quiet(de.densityGridN, F, S, g, -0.2, 0.3, -1, 16, 300.0, False))
    out["densityRealN_T300"] = np.asarray(quiet(de.densityRealN, F, S, g, -12.0, 0.1, 24, 300.0, False))
    out["densityComplexN_legendre"] = np.asarray(quiet(de.densityComplexN, F, S, g, -12.0, 0.1, 24, 0.0, False, "legendre"))
    out["densityComplexN_mid"] = np.asarray(quiet(de.densityComplexN, F, S, g, -12.0, 0.1, 24, 0.0, False, "midpoint"))
    out["densityComplexN_T300"] = np.asarray(quiet(de.densityComplexN, F, S, g, -12.0, 0.1, 24, 300.0, False))
    out["densityGrid_adaptive"] = np.asarray(quiet(de.densityGrid, F, S, g, -0.2, 0.3, None, 1e-4, 0.0))
    out["densityReal_adaptive"] = np.asarray(quiet(de.densityReal, F, S, g, -12.0, 0.1, 1e-3, 0.0, 200))
    return out
