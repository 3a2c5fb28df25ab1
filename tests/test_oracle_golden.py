"""Pin the numpy oracle (oracle/negf_oracle.py) to the reference's behaviour: every golden vector
under tests/golden/ was produced by the UNMODIFIED reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from conftest import relerr
from gaunegf_b200 import synthetic as sy
from oracle import negf_oracle as O

TOL = 1e-10


def test_cfg1_chain_transmission_dos_current(golden):
    G = golden("cfg1_chain")
    F, S, s1, s2 = sy.chain(64)
    calc = O.SigmaCalculator(s1, s2, energy_dependent=False)
    T = O.calculate_transmission(F, S, calc, G["E"])
    assert relerr(T, G["T"]) < 1e-10
    tot, per = O.calculate_dos(F, S, calc, G["Ed"])
    assert relerr(tot, G["dos_tot"]) < TOL and relerr(per, G["dos_site"]) < TOL
    for (qV, Tk), ref in zip(G["cur_args"], G["cur"]):
        assert abs(O.calculate_current(F, S, calc, 0.0, qV, T=Tk, dE=0.01) - ref) <= 1e-10 * abs(ref)
    F2, S2, _, _ = sy.chain(40)
    calc2 = O.SigmaCalculator(G["sm1"], G["sm2"], energy_dependent=False)
    assert relerr(O.calculate_transmission(F2.astype(complex), S2, calc2, G["E2"]), G["Tm"]) < TOL


def test_cfg1_spin(golden):
    G = golden("cfg1_spin")
    Fs, Ss = sy.hermitian_pair(24, seed=5, complex_F=True)
    su1, su2 = sy.block_sigma_vectors(12, 3, 0.1)
    calc = O.SigmaCalculator(su1, su2, energy_dependent=False)
    for spin, kT, k4 in (("u", "Tu", "Tu4"), ("g", "Tg", "Tg4")):
        T, T4 = O.calculate_transmission(Fs, Ss, calc, G["Es"], spin=spin)
        assert relerr(T, G[kT]) < TOL and relerr(T4, G[k4]) < TOL


def test_cfg2_density48(golden):
    G = golden("cfg2_density48")
    N = 48
    F, S = sy.hermitian_pair(N, seed=0)
    inds = sy.end_contacts(N, 6)
    g = O.surfGTest(F, S, inds, -0.1j, -0.1j)
    assert relerr(O.densityComplex(F, S, g, -30.0, 0.0, 1e-4, 0.0), G["Pc"]) < TOL
    assert relerr(O.densityComplexN(F, S, g, -30.0, 0.0, 54, 300.0, "ant"), G["PcN"]) < TOL
    assert relerr(O.densityComplexN(F, S, g, -30.0, 0.0, 40, 0.0, "legendre"), G["PcNl"]) < TOL
    assert relerr(O.densityRealN(F, S, g, -8.0, 0.0, 64, 0.0), G["PrN"]) < TOL
    assert relerr(O.densityReal(F, S, g, -8.0, 0.0, 1e-2, 0.0), G["Pr"]) < TOL
    assert relerr(O.densityGridN(F, S, g, -0.25, 0.25, -1, 60, 0.0), G["PgN"]) < TOL
    assert relerr(O.densityGridN(F, S, g, 0.25, -0.25, None, 60, 300.0), G["PgN0"]) < TOL
    assert relerr(O.densityGrid(F, S, g, -0.25, 0.25, 0, 1e-4, 0.0), G["Pg"]) < TOL
    assert relerr(O.GrInt(F, S, g, G["z"], G["w"]), G["GI"]) < TOL
    assert relerr(O.GrLessInt(F, S, g, G["z"].real, G["w"].real, 0), G["GL"]) < TOL
    assert relerr(O.GrLessInt(F, S, g, G["z"].real, G["w"].real, None), G["GLn"]) < TOL


def test_cfg2_density256(golden):
    G = golden("cfg2_density256")
    N = 256
    F, S = sy.hermitian_pair(N, seed=0)
    g = O.surfGTest(F, S, sy.end_contacts(N, 16), -0.1j, -0.1j)
    trace = []
    P = O.densityComplex(F, S, g, -30.0, 0.0, 1e-4, 0.0, trace=trace)
    assert trace[-1][0] == 162                      # SURVEY probe: converges at 162 points
    assert abs(np.trace(P @ S) - G["trPS"]) < 1e-10 * abs(G["trPS"])
    assert relerr(P[G["ii"], G["jj"]], G["samp"]) < TOL


def test_cfg3_transmission(golden):
    G = golden("cfg3_trans")
    N = 256
    F, S = sy.hermitian_pair(N, seed=1)
    s1, s2 = sy.block_sigma_vectors(N, 16, 0.1)
    calc = O.SigmaCalculator(s1, s2, energy_dependent=False)
    assert relerr(O.calculate_transmission(F, S, calc, G["E3b"]), G["T3b"]) < TOL
    assert abs(O.calculate_current(F, S, calc, 0.0, 0.4, T=0.0, dE=0.01) - G["I3b"]) < 1e-10 * abs(G["I3b"])


@pytest.mark.parametrize("tag,eta", [("a", 0.05), ("b", 1e-4)])
def test_cfg4_surfg1d(golden, tag, eta):
    G = golden("cfg4_surfg1d")
    F, S, inds, taus = sy.lead_device_lead(16, 32, seed=2, s_off=0.05)
    g = O.surfG1D(F, S, inds, taus, eta=eta)
    E4 = G["E4"]
    g0 = np.array([g.g(e, 0) for e in E4])
    its = np.array([g.last_iters[(complex(e), 0)][0] for e in E4])
    conv = its < g.MAX_ITER
    # converged energies: damped contraction, roundoff does not amplify -> 1e-10
    assert relerr(g0[conv], G["g0_" + tag][conv]) < TOL
    sig0 = np.array([g.sigma(e, 0)[np.ix_(inds[0], inds[0])] for e in E4])
    sig1 = np.array([g.sigma(e, 1)[np.ix_(inds[1], inds[1])] for e in E4])
    conv1 = np.array([g.last_iters[(complex(e), 1)][0] for e in E4]) < g.MAX_ITER
    assert relerr(sig0[conv], G["sig0_" + tag][conv]) < TOL
    assert relerr(sig1[conv1], G["sig1_" + tag][conv1]) < TOL
    if tag == "a":
        assert conv.all() and conv1.all()
        calc = O.SigmaCalculator(g, energy_dependent=True)
        assert relerr(O.calculate_transmission(F, S, calc, E4), G["T4"]) < TOL
        assert relerr(O.GrInt(F, S, g, G["zc"], np.array([1.0, 0.5j, -0.25])), G["GI4"]) < TOL
        assert relerr(O.GrLessInt(F, S, g, E4[:4], np.ones(4) * 0.1, -1), G["GL4"]) < TOL
        assert relerr(O.calculate_dos(F, S, calc, E4)[0], G["dos4"]) < TOL


def bethe_from_golden(G, F, S):
    gl = [O.surfGBAt(G["H"][i], G["Slist"][i], G["Vlist"][i], float(G["eta"])) for i in range(2)]
    lens, flat = G["nInd_len"], list(G["nInd_flat"])
    nil, p = [], 0
    for c in lens:
        cl = []
        for n in c:
            cl.append(flat[p:p + n])
            p += n
        nil.append(cl)
    return O.surfGB(F, S, gl, G["indsLists"], nil)


def test_cfg5_bethe(golden):
    G = golden("cfg5_bethe")
    Nb = int(G["Nb"])
    F, S = sy.hermitian_pair(Nb, seed=3)
    gB = bethe_from_golden(G, F, S)
    at = gB.gList[0]
    for k, E in enumerate(G["E5"]):
        assert relerr(at.sigmaK(E), G["sigK"][k]) < TOL
        assert relerr(at.sigma(E), G["sigS"][k]) < TOL
        assert relerr(gB.sigma(E, 0), G["sigB0"][k]) < TOL
        assert relerr(gB.sigmaTot(E), G["sigBt"][k]) < TOL
        assert abs(at.DOS(E) - G["dosB"][k]) < 1e-9 * abs(G["dosB"][k])
    mu = float(G["fermi"])
    assert relerr(O.densityGridN(F, S, gB, mu - 0.25, mu + 0.25, -1, 6, 0.0), G["PgB"]) < TOL


def test_oracle_vs_live_reference_when_present():
    """in the build container the UNMODIFIED reference (under the jax->numpy shim) is run side by side with the oracle
    on inputs that are NOT in the golden files; skipped where /root/reference does not exist (the GPU box)"""
    import contextlib
    import io
    from oracle.refload import reference_available, load_reference
    if not reference_available():
        pytest.skip("/root/reference not present")
    R = load_reference()
    tr, it, sgt = R["transport"], R["integrate"], R["surfGTester"]
    N = 30
    F, S = sy.hermitian_pair(N, seed=77)
    s1, s2 = sy.block_sigma_vectors(N, 4, 0.15)
    E = np.linspace(-1.3, 1.1, 23)
    with contextlib.redirect_stdout(io.StringIO()):
        Tref = np.array(tr.cohTrans(E, F, S, s1, s2))
        dref, pref = tr.DOS(E, F, S, s1, s2)
    calc = O.SigmaCalculator(s1, s2, energy_dependent=False)
    assert relerr(O.calculate_transmission(F, S, calc, E), Tref) < TOL
    tot, per = O.calculate_dos(F, S, calc, E)
    assert relerr(tot, np.array(dref)) < TOL and relerr(per, np.array(pref)) < TOL
    inds = sy.end_contacts(N, 4)
    g, og = sgt.surfGTest(F, S, inds, -0.1j, -0.2j), O.surfGTest(F, S, inds, -0.1j, -0.2j)
    z, w = sy.contour_points(10, -6.0, 0.2)
    assert relerr(O.GrInt(F, S, og, z, w), np.asarray(it.GrInt(F, S, g, z, w))) < TOL
    Er, wr = np.linspace(-0.4, 0.4, 9), np.full(9, 0.1)
    for ind in (None, 0, -1):
        assert relerr(O.GrLessInt(F, S, og, Er, wr, ind), np.asarray(it.GrLessInt(F, S, g, Er, wr, ind))) < TOL


# ---- BASELINE-size goldens (tests/golden/make_golden_full.py): a bounded part of each on the CPU -------------------
def test_cfg4_full_size_oracle(golden):
    """n_c = 128 fixed point: the oracle's iteration counts and g equal the reference's (3 of the 32 golden problems;
    the GPU test covers all of them)"""
    from full_cases import cfg4_system
    G = golden("cfg4_full")
    F, S, li, taus = cfg4_system()
    og = O.surfG1D(F, S, li, taus, eta=0.02)
    ii, jj = G["ii"], G["jj"]
    for k, c in ((2, 0), (7, 1), (12, 1)):
        gm = og.g(G["E"][k], c)
        assert og.last_iters[(complex(G["E"][k]), c)][0] == G["iters_a"][k, c]
        assert relerr(gm[ii, jj], G["g_samp_a"][k, c]) < TOL
        assert abs(np.linalg.norm(gm) - G["g_fro_a"][k, c]) < TOL * G["g_fro_a"][k, c]


def test_cfg5_full_size_oracle(golden):
    """N = 2048 GrLessInt(ind = -1) with Bethe contacts: oracle == reference (4 energies, about 10 s of LAPACK)"""
    from full_cases import bethe_atoms, check_sampled, nind_lists
    G = golden("cfg5_full")
    N = int(G["N"])
    F, S = sy.hermitian_pair(N, seed=3)
    gB = O.surfGB(F, S, bethe_atoms(G, O.surfGBAt, eta=float(G["eta"])), G["indsLists"], nind_lists(G))
    check_sampled(O.GrLessInt(F, S, gB, G["Eg"], G["wg"], -1), G, "GL_last", TOL, relerr)


def test_bethe_orthonormal_and_spin_oracle(golden):
    """Xi Sigma Xi and the spin kron (surfGBethe.py:529-539): oracle == reference"""
    from full_cases import bethe_atoms, nind_lists, spin_system
    G = golden("bethe_xi")
    Nb = int(G["Nb"])
    F, S = sy.hermitian_pair(Nb, seed=3)
    gO = O.surfGB(F, S, bethe_atoms(G, O.surfGBAt, "o_"), G["o_indsLists"], nind_lists(G, "o_"), Xi=G["o_Xi"],
                  orthonormal=True)
    E = G["E_o"]
    assert relerr(gO.sigma(E[0], 0), G["o_sig0"]) < TOL and relerr(gO.sigmaTot(E[1]), G["o_sigT"]) < TOL
    calc = O.SigmaCalculator(gO, energy_dependent=True)
    assert relerr(O.calculate_transmission(F, S, calc, E), G["o_T"]) < TOL
    assert relerr(O.GrLessInt(F, S, gO, E, np.array([0.2, 0.5, 0.3]), -1), G["o_GL"]) < TOL
    assert relerr(O.GrInt(F, S, gO, E + 0.3j, np.array([0.2, 0.5j, 0.3])), G["o_GI"]) < TOL
    ii, jj = G["ii"], G["jj"]
    for sp in ("u", "g"):
        F2, S2 = spin_system(F, S, sp)
        gS = O.surfGB(F2, S2, bethe_atoms(G, O.surfGBAt, sp + "_"), G[sp + "_indsLists"], nind_lists(G, sp + "_"), spin=sp)
        Es = G["E_" + sp]
        assert relerr(gS.sigmaTot(Es[0])[ii, jj], G[sp + "_sigT_samp"]) < TOL
        P = O.GrInt(F2, S2, gS, Es + 0.2j, np.array([1.0, -0.5j]))
        assert relerr(P[ii, jj], G[sp + "_GI_samp"]) < TOL and relerr(np.diag(P), G[sp + "_GI_diag"]) < TOL
        P = O.GrLessInt(F2, S2, gS, Es, np.array([0.6, 0.4]), 0)
        assert relerr(P[ii, jj], G[sp + "_GL_samp"]) < TOL and relerr(np.diag(P), G[sp + "_GL_diag"]) < TOL
