"""GPU parity of the C-ABI (libgaunegf_b200.so, called through ctypes) against the numpy oracle.
Tolerance: 1e-10 relative (to the largest element) in complex128, as BASELINE.json states."""
import numpy as np
import pytest

from conftest import relerr
from gaunegf_b200 import synthetic as sy
from oracle import negf_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(scope="module")
def ctx():
    from gaunegf_b200._native import Context
    c = Context(0)
    yield c
    c.close()


def const_system(ctx, N, nc, seed=0, complex_F=False):
    F, S = sy.hermitian_pair(N, seed=seed, complex_F=complex_F)
    inds = sy.end_contacts(N, nc)
    rng = np.random.default_rng(seed + 100)
    blks = []
    for _ in inds:
        b = rng.standard_normal((nc, nc)) * 0.02
        blks.append((b + b.T) / 2 - 0.1j * np.eye(nc))
    ctx.set_system(F, S)
    ctx.sigma_clear()
    sig = []
    for i, b in zip(inds, blks):
        ctx.sigma_add_const_block(i, b)
        s = np.zeros((N, N), dtype=complex)
        s[np.ix_(i, i)] = b
        sig.append(s)
    return F, S, inds, sig


@pytest.mark.parametrize("n", [1, 5, 31, 32, 33, 64, 100, 257, 300, 520])
def test_inverse_batch(ctx, n):
    rng = np.random.default_rng(n)
    A = rng.standard_normal((7, n, n)) + 1j * rng.standard_normal((7, n, n))
    Ai = ctx.inverse_batch(A)
    ref = np.linalg.inv(A)
    for k in range(7):
        assert relerr(Ai[k], ref[k]) < 1e-11 * max(1.0, np.linalg.cond(A[k]) / 100)


def test_inverse_singular_raises(ctx):
    A = np.zeros((2, 40, 40), dtype=complex)
    with pytest.raises(np.linalg.LinAlgError):
        ctx.inverse_batch(A)


@pytest.mark.parametrize("N,nc", [(48, 6), (64, 1), (130, 16), (300, 40)])
def test_green_dos_transmission(ctx, N, nc):
    F, S, inds, sig = const_system(ctx, N, nc, seed=N, complex_F=(N == 130))
    E = np.concatenate([np.linspace(-1, 1, 9), [0.3 + 0.7j, -2 + 0.01j]])
    st = sig[0] + sig[1]
    G = ctx.green(E)
    Gref = np.array([O.gr_matrix(st, e, F, S) for e in E])
    for k in range(len(E)):
        assert relerr(G[k], Gref[k]) < TOL
    tot, per = ctx.dos(E)
    assert relerr(per, np.array([-np.imag(np.diag(g)) / np.pi for g in Gref])) < TOL
    assert relerr(tot, np.array([-np.imag(np.trace(g)) / np.pi for g in Gref])) < TOL
    Er = E[:9].real
    T = ctx.transmission(Er, 0, -1)
    g1 = 1j * (sig[0] - sig[0].conj().T)
    g2 = 1j * (sig[1] - sig[1].conj().T)
    Tref = np.array([O.transmission_restricted(e, F, S, st, g1, g2) for e in Er])
    assert relerr(T, Tref) < 1e-10
    Td = ctx.transmission_dense(Er, st, g1, g2)
    assert relerr(Td, Tref) < 1e-10


@pytest.mark.parametrize("N,nc", [(48, 6), (200, 24)])
def test_integrals(ctx, N, nc):
    F, S, inds, sig = const_system(ctx, N, nc, seed=3)
    g = type("G", (), {})()
    g.sigmaTot = lambda E: sig[0] + sig[1]
    g.sigma = lambda E, i: sig[i]
    z, w = sy.contour_points(18, -12.0, 0.0)
    assert relerr(ctx.gr_int(z, w), O.GrInt(F, S, g, z, w)) < TOL
    Er = np.linspace(-0.4, 0.4, 11)
    wr = np.linspace(0.1, 0.3, 11)
    assert relerr(ctx.gless_int(Er, wr, 1), O.GrLessInt(F, S, g, Er, wr, 1)) < TOL
    assert relerr(ctx.gless_int(Er, wr, -1), O.GrLessInt(F, S, g, Er, wr, None)) < TOL
    st = sig[0] + sig[1]
    gam = 1j * (st - st.conj().T)
    assert relerr(ctx.gless_int_dense(Er, wr, st, gam), O.GrLessInt(F, S, g, Er, wr, None)) < TOL
    assert relerr(ctx.gr_int_dense(z, w, np.broadcast_to(st, (len(z), N, N)).copy()), O.GrInt(F, S, g, z, w)) < TOL


@pytest.mark.parametrize("N,nc", [(48, 6), (200, 24)])
def test_segmented_gr_int_equals_separate_calls(ctx, N, nc):
    """gnb_gr_int_seg (several nested quadrature levels in one batch): every segment equals its own gnb_gr_int call
    bit for bit, with ragged and empty segments, across chunk boundaries, and for a dense Sigma"""
    F, S, inds, sig = const_system(ctx, N, nc, seed=5)
    z, w = sy.contour_points(18, -12.0, 0.0)
    ends = [2, 6, 6, 18]
    seg = ctx.gr_int_seg(z, w, ends)
    assert seg.shape == (4, N, N)
    lo = 0
    for s_, hi in enumerate(ends):
        ref = ctx.gr_int(z[lo:hi], w[lo:hi])
        assert np.array_equal(seg[s_], ref)
        lo = hi
    st = sig[0] + sig[1]
    segd = ctx.gr_int_seg(z, w, ends, sig=st)
    assert relerr(segd[3], seg[3]) < TOL and relerr(segd[0], seg[0]) < TOL
    ctx.set_workspace_limit(64 << 20)                      # force several chunks: segments straddle them
    try:
        z2, w2 = sy.contour_points(162, -12.0, 0.0)
        ends2 = [36, 144, 162]
        seg2 = ctx.gr_int_seg(z2, w2, ends2)
        lo = 0
        for s_, hi in enumerate(ends2):
            assert relerr(seg2[s_], O.GrInt(F, S, type("G", (), {"sigmaTot": lambda self, E: st})(), z2[lo:hi], w2[lo:hi])) < TOL
            lo = hi
    finally:
        ctx.set_workspace_limit(48 << 30)
    with pytest.raises(Exception):
        ctx.gr_int_seg(z, w, [4, 2, 18])


@pytest.mark.parametrize("N", [130, 256])
def test_block_engine_singular_raises(ctx, N):
    """exactly singular A (zero rows off the contacts) on the block engine: numpy.linalg.LinAlgError like the reference's
    solve(), on the real-structure T(E) path (FP32 pivot order + FP64 pivot-block inverse detects the zero pivot), the
    complex Gauss-Jordan path and GrLessInt"""
    nc = 16
    Z = np.zeros((N, N))
    ctx.set_system(Z, Z)
    ctx.sigma_clear()
    ctx.sigma_add_const_block(np.arange(nc), -0.1j * np.eye(nc))
    ctx.sigma_add_const_block(np.arange(N - nc, N), -0.1j * np.eye(nc))
    E = np.linspace(-0.3, 0.3, 5)
    with pytest.raises(np.linalg.LinAlgError):
        ctx.transmission(E, 0, -1)
    with pytest.raises(np.linalg.LinAlgError):
        ctx.green(E + 0.1j)
    with pytest.raises(np.linalg.LinAlgError):
        ctx.gless_int(E, np.ones(5), -1)
    # and the context is usable afterwards
    F, S, inds, sig = const_system(ctx, N, nc, seed=3)
    st = sig[0] + sig[1]
    assert relerr(ctx.green(E[:1] + 0.1j)[0], O.gr_matrix(st, E[0] + 0.1j, F, S)) < TOL


def test_set_system_cached_uploads_only_what_changed(ctx):
    """gnb_set_system_cached: F and S are compared with the context's pinned shadows; an in-place change of either is
    seen, only the changed matrix is re-sent, and the results follow the new values"""
    N, nc = 70, 6
    F, S, inds, sig = const_system(ctx, N, nc, seed=11)
    st = sig[0] + sig[1]
    E = np.array([0.13, -0.4 + 0.2j])
    ctx.set_system(F, S)
    assert ctx.last_system_upload == 0                     # const_system has just installed the same arrays
    F2 = np.array(F, dtype=complex)
    F2[5, 7] += 0.01; F2[7, 5] += 0.01
    ctx.set_system(F2, S)
    assert ctx.last_system_upload == 1
    G = ctx.green(E)
    assert relerr(G, np.array([O.gr_matrix(st, e, F2, S) for e in E])) < TOL
    F2[0, 0] += 1e-13                                      # in-place change of one entry of the caller's array
    ctx.set_system(F2, S)
    assert ctx.last_system_upload == 1
    S2 = np.array(S, dtype=complex)
    S2[3, 3] *= 1.001
    ctx.set_system(F2, S2)
    assert ctx.last_system_upload == 2
    assert relerr(ctx.green(E), np.array([O.gr_matrix(st, e, F2, S2) for e in E])) < TOL
    Fc = F2 + 0.003j * (np.triu(np.ones((N, N)), 1) - np.tril(np.ones((N, N)), -1))      # complex Hermitian F
    ctx.set_system(Fc, S2)
    Er = np.linspace(-0.3, 0.3, 5)
    g1 = 1j * (sig[0] - sig[0].conj().T); g2 = 1j * (sig[1] - sig[1].conj().T)
    assert relerr(ctx.transmission(Er, 0, -1), np.array([O.transmission_restricted(e, Fc, S2, st, g1, g2) for e in Er])) < TOL
    ctx.set_system(F2, S2)                                 # back to real input: the real-structure shortcut applies again
    assert relerr(ctx.transmission(Er, 0, -1), np.array([O.transmission_restricted(e, F2, S2, st, g1, g2) for e in Er])) < TOL
    # read-only comparison (what rank 0 runs for all ranks) and the comparison-free update that follows the agreed flags
    assert ctx.system_differs(F2, S2) == 0 and ctx.system_differs(F2, S2, full=False) == 0
    F3 = F2.copy(); F3[N - 1, N - 2] += 1e-9
    S3 = S2.copy(); S3[0, 0] *= 1.0005
    assert ctx.system_differs(F3, S2) == 1 and ctx.system_differs(F2, S3) == 2 and ctx.system_differs(F3, S3) == 3
    assert ctx.system_differs(F2, S3, full=False) == 2                       # element 0 is in the sample
    ctx.set_system_known(F3, S3, 2)                                           # only S is taken over ...
    assert ctx.last_system_upload == 2 and ctx.system_differs(F3, S3) == 1    # ... F is still the old one
    ctx.set_system_known(F3, S3, 1)
    assert ctx.system_differs(F3, S3) == 0
    assert relerr(ctx.green(E), np.array([O.gr_matrix(st, e, F3, S3) for e in E])) < TOL
    Fs, Ss = sy.hermitian_pair(40, seed=2)
    assert ctx.system_differs(Fs, Ss) == 3                                    # another size: nothing comparable
    ctx.set_system_known(Fs, Ss, 0)                                           # falls back to the comparing call
    assert ctx.last_system_upload == 3 and ctx.N == 40


def test_chunking_matches_single_pass(ctx):
    F, S, inds, sig = const_system(ctx, 96, 8, seed=9)
    z, w = sy.contour_points(54, -12.0, 0.0)
    full = ctx.gr_int(z, w)
    ctx.set_workspace_limit(64 << 20)
    try:
        small = ctx.gr_int(z, w)
        T1 = ctx.transmission(np.linspace(-1, 1, 300))
    finally:
        ctx.set_workspace_limit(16 << 30)
    assert relerr(small, full) < 1e-13
    T2 = ctx.transmission(np.linspace(-1, 1, 300))
    assert np.array_equal(T1, T2)


def test_cfg1_golden_transmission(ctx, golden):
    G = golden("cfg1_chain")
    F, S, s1, s2 = sy.chain(64)
    ctx.set_system(F, S)
    ctx.sigma_clear()
    ctx.sigma_add_const_block([0], [[s1[0]]])
    ctx.sigma_add_const_block([63], [[s2[63]]])
    T = ctx.transmission(G["E"])
    assert relerr(T, G["T"]) < 1e-10
    tot, per = ctx.dos(G["Ed"])
    assert relerr(tot, G["dos_tot"]) < TOL and relerr(per, G["dos_site"]) < TOL


@pytest.mark.parametrize("tag,eta", [("a", 0.05), ("b", 1e-4)])
def test_chain1d_sigma(ctx, golden, tag, eta):
    G = golden("cfg4_surfg1d")
    F, S, inds, taus = sy.lead_device_lead(16, 32, seed=2, s_off=0.05)
    og = O.surfG1D(F, S, inds, taus, eta=eta)
    ctx.set_system(F, S)
    ctx.sigma_clear()
    for i in range(2):
        ctx.sigma_add_chain1d(inds[i], og.aList[i], og.aSList[i], og.bList[i], og.bSList[i], og.tauList[i],
                              og.stauList[i], eta, 1e-5, 0.1, 2000)
    E4 = G["E4"]
    g0, iters, diffs = ctx.sigma_eval(0, 1, E4, (16, 16))
    ref_it = []
    for e in E4:
        og.g(e, 0)
        ref_it.append(og.last_iters[(complex(e), 0)][0])
    ref_it = np.array(ref_it)
    assert np.array_equal(iters, ref_it), (iters, ref_it)
    conv = ref_it < 2000
    assert relerr(g0[conv], G["g0_" + tag][conv]) < TOL
    s0, _, _ = ctx.sigma_eval(0, 0, E4, (16, 16))
    assert relerr(s0[conv], G["sig0_" + tag][conv]) < TOL
    if tag == "a":
        T = ctx.transmission(E4)
        assert relerr(T, G["T4"]) < TOL
        assert relerr(ctx.gr_int(G["zc"], np.array([1.0, 0.5j, -0.25])), G["GI4"]) < TOL
        assert relerr(ctx.gless_int(E4[:4], np.ones(4) * 0.1, 1), G["GL4"]) < TOL
        assert relerr(ctx.dos(E4)[0], G["dos4"]) < TOL


def test_bethe_sigma(ctx, golden):
    G = golden("cfg5_bethe")
    Nb = int(G["Nb"])
    F, S = sy.hermitian_pair(Nb, seed=3)
    ctx.set_system(F, S)
    ctx.sigma_clear()
    lens, flat = G["nInd_len"], list(G["nInd_flat"])
    p = 0
    for c in range(2):
        nbl = []
        for n in lens[c]:
            nbl.append(flat[p:p + n])
            p += n
        ctx.sigma_add_bethe(G["indsLists"][c], nbl, G["H"][c], G["Slist"][c], G["Vlist"][c], float(G["eta"]), 1e-5, 0.5, 1000)
    E5 = G["E5"]
    sK, _, _ = ctx.sigma_eval(0, 1, E5, (12, 9, 9))
    assert relerr(sK, G["sigK"]) < TOL
    sS, _, _ = ctx.sigma_eval(0, 2, E5, (9, 9, 9))
    assert relerr(sS, G["sigS"]) < TOL
    blk, _, _ = ctx.sigma_eval(0, 0, E5, (27, 27))
    ii = np.asarray(G["indsLists"][0]).reshape(-1)
    ref = G["sigB0"][:, ii][:, :, ii]
    assert relerr(blk, ref) < TOL
    Gd = ctx.green(E5)
    for k, e in enumerate(E5):
        assert relerr(Gd[k], O.gr_matrix(G["sigBt"][k], e, F, S)) < TOL
    mu = float(G["fermi"])
    from scipy.special import roots_legendre
    x, w = roots_legendre(6)
    mid = 0.25
    Eg = mid * (x + 1) + mu - 0.25
    out = ctx.gless_int(Eg, mid * w * 1.0, 1) / (2 * np.pi)
    assert relerr(out, G["PgB"]) < TOL


# ---------------------------------------------------------------------------------------------------
# Recursive multi-level engine (gnb_rec.cu): A/B against the two-level engine and its own switches
# ---------------------------------------------------------------------------------------------------
def _set(ctx, **opts):
    for k, v in opts.items():
        assert ctx.lib.gnb_dev_set_option(k.encode(), int(v)) == 0


DEFAULTS = dict(engine_rec=1, rk_m3=1, rk_m3_mink=64, rk_kskip=1, contacts_last=1, tourn_fp32=1, rec_streams=2,
                rk_real=1, mixed_layout=1, rk_strip=1, rk_wsolve_mma=1, rk_fin_mma=1, rk_augreal=1, rk_wskip=1,
                rk_wsolve_fused=1, rk_lookahead=0, rk_la_ctas=1, rk_cs=4, tourn_warp=249, gless_mixed=1, rk_tcap_k=0,
                rk_lowprio=0, tournq_cplx_min_m=400)
# round-2 switches: (option, value A, value B)
ROUND2_SWITCHES = [("rk_augreal", 1, 0), ("rk_wskip", 1, 0), ("rk_wsolve_fused", 1, 0), ("rk_lookahead", 0, 1),
                   ("rk_cs", 4, 3), ("tourn_warp", 249, 1), ("tourn_warp", 249, 9), ("tourn_warp", 249, 25),
                   ("gless_mixed", 1, 0), ("rk_tcap_k", 0, 256), ("tournq_cplx_min_m", 400, 1)]


@pytest.mark.parametrize("N,nc", [(96, 8), (100, 7), (256, 16), (416, 33), (600, 40)])
def test_recursive_engine_matches_two_level_engine_and_numpy(ctx, N, nc):
    """padded sizes (N, naug not multiples of 32), both elimination modes, against numpy and the other engine"""
    F, S, inds, sig = const_system(ctx, N, nc, seed=N + 1)
    st = sig[0] + sig[1]
    E = np.array([0.13 + 0j, -0.8 + 0.5j, 0.41 + 1e-6j])
    Er = np.linspace(-0.7, 0.9, 5)
    z, w = sy.contour_points(6, -5.0, 0.0)
    out = {}
    try:
        for eng in (1, 0):
            _set(ctx, engine_rec=eng)
            out[eng] = (ctx.green(E), ctx.transmission(Er, 0, -1), ctx.gr_int(z, w), ctx.gless_int(Er, np.ones(5) * 0.2, -1))
    finally:
        _set(ctx, **DEFAULTS)
    Gref = np.array([O.gr_matrix(st, e, F, S) for e in E])
    for k in range(len(E)):
        assert relerr(out[1][0][k], Gref[k]) < TOL
    for a, b in zip(out[1], out[0]):
        assert relerr(a, b) < TOL
    g1 = 1j * (sig[0] - sig[0].conj().T)
    g2 = 1j * (sig[1] - sig[1].conj().T)
    Tref = np.array([O.transmission_restricted(e, F, S, st, g1, g2) for e in Er])
    assert relerr(out[1][1], Tref) < 1e-10


@pytest.mark.parametrize("opt", ["rk_m3", "rk_kskip", "contacts_last", "tourn_fp32", "rec_streams", "rk_real",
                                 "mixed_layout", "rk_strip", "rk_wsolve_mma", "rk_fin_mma"])
def test_recursive_engine_switches_do_not_change_results(ctx, opt):
    """3M arithmetic, block-upper K skipping, contacts-last ordering, FP32 nominating rounds and sub-batch streams
    are performance switches: results must stay within the parity tolerance of the plain path."""
    N, nc = 320, 32
    F, S, inds, sig = const_system(ctx, N, nc, seed=7)
    E = np.concatenate([np.linspace(-1, 1, 70), [0.2 + 0.3j]])
    Er = np.linspace(-1, 1, 70)
    res = {}
    try:
        for v in ((1, 0) if opt != "rec_streams" else (1, 3)):
            _set(ctx, **{opt: v})
            res[v] = (ctx.green(E[-2:]), ctx.transmission(Er, 0, -1), ctx.dos(E)[0])
    finally:
        _set(ctx, **DEFAULTS)
    a, b = list(res.values())
    for x, y in zip(a, b):
        assert relerr(x, y) < TOL


@pytest.mark.parametrize("opt,va,vb", ROUND2_SWITCHES)
def test_round2_switches_do_not_change_results(ctx, opt, va, vb):
    """real arithmetic on the augmented columns, the skipped W write, the fused leaf pair, the look-ahead, cache hints, the
    tournament kernels (k_tournw / k_tournq with 2 or 4 rows per lane, FP64 or FP32-order final round), the mixed layout
    of GrLessInt and short-lived rank-K CTAs are performance switches: T(E), the GrLessInt integral and a complex-F
    T(E) (complex panels) must stay within the parity tolerance, and match the oracle"""
    N, nc = 416, 33
    F, S, inds, sig = const_system(ctx, N, nc, seed=13)
    Er = np.linspace(-0.9, 0.9, 130)
    wr = np.linspace(0.1, 0.3, 130)
    res = {}
    try:
        for v in (va, vb):
            _set(ctx, **{opt: v})
            ctx.set_system(F, S)
            out = [ctx.transmission(Er, 0, -1), ctx.gless_int(Er, wr, -1), ctx.gless_int(Er, wr, 0)]
            Fc, _ = sy.hermitian_pair(N, seed=13, complex_F=True)
            ctx.set_system(Fc, S)
            out.append(ctx.transmission(Er[:40], 0, -1))
            res[v] = out
    finally:
        _set(ctx, **DEFAULTS)
    for x, y in zip(res[va], res[vb]):
        assert relerr(x, y) < TOL
    st = sig[0] + sig[1]
    g1 = 1j * (sig[0] - sig[0].conj().T)
    g2 = 1j * (sig[1] - sig[1].conj().T)
    Tref = np.array([O.transmission_restricted(e, F, S, st, g1, g2) for e in Er[::13]])
    assert relerr(res[vb][0][::13], Tref) < TOL


def test_full_size_properties_n1024(ctx):
    """BASELINE size (N = 1024, 64-orbital contacts): size-independent properties instead of a CPU comparison of
    every point -- G A = I, T from the contact-column solve == T from the full inverse, reciprocity T12 == T21,
    0 <= T <= min(n1, n2); plus one energy against numpy."""
    N, nc = 1024, 64
    F, S = sy.hermitian_pair(N, seed=1)
    s1, s2 = sy.block_sigma_vectors(N, nc, 0.1)
    ctx.set_system(F, S)
    ctx.sigma_clear()
    ctx.sigma_add_const_block(np.arange(nc), np.diag(s1[:nc]))
    ctx.sigma_add_const_block(np.arange(N - nc, N), np.diag(s2[N - nc:]))
    E = np.linspace(-0.5, 0.5, 40)
    T12 = ctx.transmission(E, 0, -1)
    T21 = ctx.transmission(E, -1, 0)
    assert np.all(T12 > -1e-12) and np.all(T12 < nc + 1e-9)
    assert relerr(T12, T21) < 1e-10
    G = ctx.green(E[:3])
    sig = np.diag(s1 + s2)
    g1, g2 = np.diag(-2 * s1.imag), np.diag(-2 * s2.imag)
    for k in range(3):
        A = E[k] * S - F - sig
        assert np.abs(G[k] @ A - np.eye(N)).max() < 1e-9
        Tfull = np.trace(g1 @ G[k] @ g2 @ G[k].conj().T).real
        assert abs(Tfull - T12[k]) < 1e-9 * max(1.0, abs(Tfull))
    assert relerr(G[0], np.linalg.inv(E[0] * S - F - sig)) < TOL


def test_real_structure_shortcut_matches_complex_path(ctx):
    """real F, S and real energies: with the contact orbitals ordered last every column left of them stays exactly
    real, and the rank-K / forward-W kernels skip the imaginary DMMAs there (rk_real).  Same results as the full
    complex arithmetic, and as a system whose F carries an imaginary part that disables the shortcut."""
    N, nc = 352, 40
    F, S, inds, sig = const_system(ctx, N, nc, seed=11)
    Er = np.linspace(-0.9, 0.9, 33)
    st = sig[0] + sig[1]
    g1 = 1j * (sig[0] - sig[0].conj().T)
    g2 = 1j * (sig[1] - sig[1].conj().T)
    try:
        T1 = ctx.transmission(Er, 0, -1)
        _set(ctx, rk_real=0)
        T0 = ctx.transmission(Er, 0, -1)
    finally:
        _set(ctx, **DEFAULTS)
    assert relerr(T1, T0) < TOL
    Tref = np.array([O.transmission_restricted(e, F, S, st, g1, g2) for e in Er[::4]])
    assert relerr(T1[::4], Tref) < 1e-10
    # complex energies must not take the shortcut (A is complex everywhere) -- compared through GrLessInt-free T path
    Fc, Sc, indsc, sigc = const_system(ctx, N, nc, seed=11, complex_F=True)
    Tc = ctx.transmission(Er[:5], 0, -1)
    stc = sigc[0] + sigc[1]
    Tcref = np.array([O.transmission_restricted(e, Fc, Sc, stc, 1j * (sigc[0] - sigc[0].conj().T),
                                                1j * (sigc[1] - sigc[1].conj().T)) for e in Er[:5]])
    assert relerr(Tc, Tcref) < 1e-10


@pytest.mark.parametrize("n,nE", [(24, 37), (104, 7)])
def test_chain_contacts_joint_batch_equals_separate(ctx, n, nE):
    """two 1-D chain contacts with the same block size iterate as ONE lock-step batch (developer switch chain_joint)
    whose converged problems retire (chain_compact): per-problem arithmetic is unchanged, so transmission, iteration
    counts and GrLessInt must agree exactly"""
    F, S, li, taus = sy.lead_device_lead(n, 2 * n, seed=5, s_off=0.04)
    E = np.linspace(-0.8, 0.9, nE)
    res = {}
    try:
        for joint, compact in ((1, 1), (0, 0), (1, 0)):
            _set(ctx, chain_joint=joint, chain_compact=compact)
            ctx.set_system(F, S)
            ctx.sigma_clear()
            for k in range(2):
                i, t = np.asarray(li[k]), np.asarray(taus[k])
                ctx.sigma_add_chain1d(i, F[np.ix_(i, i)], S[np.ix_(i, i)], F[np.ix_(t, i)], S[np.ix_(t, i)],
                                      F[np.ix_(t, i)], S[np.ix_(t, i)], 0.03, 1e-5, 0.1)
            T = ctx.transmission(E, 0, -1)
            its = [ctx.sigma_eval(k, 1, E.astype(complex), (n, n))[1] for k in range(2)]
            P = ctx.gless_int(E, np.full(E.size, 0.05), -1)
            res[joint + compact] = (T, its[0], its[1], P)
    finally:
        _set(ctx, chain_joint=1, chain_compact=1)
    for a, b, c_ in zip(res[2], res[1], res[0]):          # joint + compacted / joint / separate
        assert np.array_equal(a, b) and np.array_equal(a, c_)
    res[1] = res[2]
    og = O.surfG1D(F, S, li, taus, eta=0.03)
    Tref = O.calculate_transmission(F, S, O.SigmaCalculator(og, energy_dependent=True), E)
    assert relerr(res[1][0], Tref) < TOL
