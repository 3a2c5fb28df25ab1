"""GPU parity of the one-CTA-per-energy shared-memory path (gnb_small.cu, N <= 119) against numpy/the oracle and
against the lock-step elimination engine (developer switch small_fused=0).  Tolerance 1e-10 relative (BASELINE.json)."""
import numpy as np
import pytest

from conftest import relerr
from gaunegf_b200 import synthetic as sy
from oracle import negf_oracle as O
from test_gpu_engine import const_system, ctx  # noqa: F401  (module-scoped context fixture)

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _small(ctx, on, reg=1):
    """on: one-CTA-per-energy path; reg: register-resident (N <= 96) or shared-memory-resident kernel"""
    ctx.lib.gnb_dev_set_option(b"small_fused", int(on))
    ctx.lib.gnb_dev_set_option(b"small_reg", int(reg))


@pytest.mark.parametrize("reg", [1, 0])
@pytest.mark.parametrize("n", [1, 2, 3, 31, 32, 33, 64, 65, 96, 97, 118, 119, 120, 127, 128, 129, 160, 191, 192, 193])
def test_small_inverse_batch(ctx, n, reg):
    """utils.inv (utils.py:52-54): reg=1: n <= 96 in the registers of one CTA, n <= 128 in those of a 2-CTA cluster,
    n <= 192 in those of a 4-CTA cluster (small batches), above on the block engine; reg=0: n <= 119 in shared memory"""
    rng = np.random.default_rng(n)
    A = rng.standard_normal((9, n, n)) + 1j * rng.standard_normal((9, n, n))
    A[3] = np.triu(A[3]) + np.eye(n) * 1e-3          # forces row exchanges to matter little / much
    A[4][[0, n - 1]] = A[4][[n - 1, 0]]
    A[5] = np.roll(np.eye(n), 1, axis=0) * (1 + 2j) + 1e-3 * A[5]      # every pivot off the diagonal
    _small(ctx, 1, reg)
    try:
        Ai = ctx.inverse_batch(A)
    finally:
        _small(ctx, 1)
    ref = np.linalg.inv(A)
    for k in range(9):
        assert relerr(Ai[k], ref[k]) < 1e-11 * max(1.0, np.linalg.cond(A[k]) / 100)


def test_cluster_inverse_matches_block_engine(ctx):
    """96 < n <= 128: 2-CTA-cluster kernel (DSMEM publishing, cluster barrier) vs the block engine"""
    rng = np.random.default_rng(5)
    for n in (100, 128):
        A = rng.standard_normal((300, n, n)) + 1j * rng.standard_normal((300, n, n))
        res = {}
        try:
            for cl in (1, 0):
                ctx.lib.gnb_dev_set_option(b"small_cluster", cl)
                res[cl] = ctx.inverse_batch(A)
        finally:
            ctx.lib.gnb_dev_set_option(b"small_cluster", 1)
        assert relerr(res[1], res[0]) < 1e-10
        assert relerr(res[1][7] @ A[7], np.eye(n)) < 1e-10


@pytest.mark.parametrize("N,nc", [(100, 12), (128, 16), (129, 9), (160, 20), (192, 24)])
def test_cluster_energy_grid_modes(ctx, N, nc):
    """96 < N <= 192 on the energy-grid calls (BASELINE north star: small orbital counts stay on chip): Green's function,
    DOS and the contour integral of a SMALL batch run on the 2- / 4-CTA cluster kernels (assembly + inverse in one
    launch); same results as the block engine (small_cluster=0) and the oracle; large batches stay on the block engine"""
    F, S, inds, sig = const_system(ctx, N, nc, seed=N, complex_F=(N == 129))
    st = sig[0] + sig[1]
    E = np.concatenate([np.linspace(-1.2, 1.1, 9), [0.3 + 0.7j, -2 + 0.01j]])
    z, w = sy.contour_points(36, -8.0, 0.0)
    out = {}
    try:
        for cl in (1, 0):
            ctx.lib.gnb_dev_set_option(b"small_cluster", cl)
            n0 = ctx.launches
            res = [ctx.green(E[-4:]), *ctx.dos(E), ctx.gr_int(z, w), ctx.gr_int_dense(z, w, st), ctx.dos_dense(E, st)[0],
                   ctx.gr_int_seg(z, w, [4, 12, 36])]
            out[cl] = (res, ctx.launches - n0)
    finally:
        ctx.lib.gnb_dev_set_option(b"small_cluster", 1)
    for a, b in zip(out[1][0], out[0][0]):
        assert relerr(a, b) < TOL
    assert out[1][1] < out[0][1] / 4                        # one launch per call instead of the block engine's chain
    Gref = np.array([O.gr_matrix(st, e, F, S) for e in E[-4:]])
    assert relerr(out[1][0][0], Gref) < TOL
    assert relerr(out[1][0][3], O.GrInt(F, S, _ConstG(st), z, w)) < TOL
    # a batch above the cluster limit takes the block engine and agrees
    z2, w2 = sy.contour_points(162, -8.0, 0.0)
    assert relerr(ctx.gr_int(z2, w2), O.GrInt(F, S, _ConstG(st), z2, w2)) < TOL


def test_small_singular_raises(ctx):
    A = np.zeros((3, 17, 17), dtype=complex)
    with pytest.raises(np.linalg.LinAlgError):
        ctx.inverse_batch(A)


@pytest.mark.parametrize("N,nc", [(7, 2), (33, 5), (64, 1), (64, 25), (100, 31), (119, 40)])
def test_small_matches_block_engine_and_oracle(ctx, N, nc):
    F, S, inds, sig = const_system(ctx, N, nc, seed=N + nc, complex_F=(N == 33))
    st = sig[0] + sig[1]
    E = np.concatenate([np.linspace(-1.2, 1.1, 150), [0.3 + 0.7j, -2 + 0.01j]])
    Er = E[:150].real
    z, w = sy.contour_points(10, -8.0, 0.0)
    out = {}
    try:
        for on, reg in ((1, 1), (1, 0), (0, 0)):
            _small(ctx, on, reg)
            out[on + reg] = (ctx.green(E[-4:]), ctx.transmission(Er, 0, -1), ctx.transmission(Er, 1, 0),
                             ctx.transmission(Er, 0, 0), *ctx.dos(E), ctx.gr_int(z, w), ctx.dos_dense(E, st)[0],
                             ctx.gr_int_dense(z, w, st))
    finally:
        _small(ctx, 1)
    for a, b, c in zip(out[2], out[1], out[0]):      # registers / shared memory / block engine
        assert relerr(a, c) < TOL and relerr(b, c) < TOL
    out[1] = out[2]
    Gref = np.array([O.gr_matrix(st, e, F, S) for e in E[-4:]])
    assert relerr(out[1][0], Gref) < TOL
    g1 = 1j * (sig[0] - sig[0].conj().T)
    g2 = 1j * (sig[1] - sig[1].conj().T)
    Tref = np.array([O.transmission_restricted(e, F, S, st, g1, g2) for e in Er])
    assert relerr(out[1][1], Tref) < 1e-10
    assert relerr(out[1][2], Tref) < 1e-10     # Tr[G1 G G2 G+] symmetric
    assert relerr(out[1][6], O.GrInt(F, S, _ConstG(st), z, w)) < TOL


class _ConstG:
    def __init__(self, st):
        self.st = st

    def sigmaTot(self, E):
        return self.st


def test_small_overlapping_contacts_and_chunks(ctx):
    """contacts that share orbitals are subtracted one after the other; chunked calls equal one pass"""
    N = 40
    F, S = sy.hermitian_pair(N, seed=3)
    i1, i2 = np.arange(0, 12), np.arange(8, 20)
    rng = np.random.default_rng(5)
    b1 = rng.standard_normal((12, 12)) * 0.05 - 0.1j * np.eye(12)
    b2 = rng.standard_normal((12, 12)) * 0.05 - 0.2j * np.eye(12)
    ctx.set_system(F, S)
    ctx.sigma_clear()
    ctx.sigma_add_const_block(i1, b1)
    ctx.sigma_add_const_block(i2, b2)
    s1 = np.zeros((N, N), complex); s1[np.ix_(i1, i1)] = b1
    s2 = np.zeros((N, N), complex); s2[np.ix_(i2, i2)] = b2
    E = np.linspace(-1, 1, 3000)
    T = ctx.transmission(E, 0, 1)
    g1, g2 = 1j * (s1 - s1.conj().T), 1j * (s2 - s2.conj().T)
    Tref = np.array([O.transmission_restricted(e, F, S, s1 + s2, g1, g2) for e in E])
    assert relerr(T, Tref) < 1e-10
    _small(ctx, 1, 0)
    try:
        assert relerr(ctx.transmission(E, 0, 1), Tref) < 1e-10
    finally:
        _small(ctx, 1)
    ctx.set_workspace_limit(64 << 20)          # -> 3 chunks of the energy list
    try:
        T2 = ctx.transmission(E, 0, 1)
        d2 = ctx.dos(E)[0]
    finally:
        ctx.set_workspace_limit(48 << 30)
    assert np.array_equal(T, T2)
    assert np.array_equal(d2, ctx.dos(E)[0])


def test_small_cfg1_golden_both_paths(ctx, golden):
    """BASELINE cfg 1 (64-orbital chain, 1000 energies): shared-memory kernel and block engine vs the reference's output"""
    G = golden("cfg1_chain")
    F, S, s1, s2 = sy.chain(64)
    ctx.set_system(F, S)
    ctx.sigma_clear()
    ctx.sigma_add_const_block([0], [[s1[0]]])
    ctx.sigma_add_const_block([63], [[s2[63]]])
    try:
        for on in (1, 0):
            _small(ctx, on)
            n0 = ctx.launches
            T = ctx.transmission(G["E"])
            assert relerr(T, G["T"]) < 1e-10
            if on:
                assert ctx.launches - n0 == 1          # the whole call is ONE kernel launch
    finally:
        _small(ctx, 1)
