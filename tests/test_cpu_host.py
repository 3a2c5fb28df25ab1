"""CPU-only checks (`-m "not gpu"`): the C-ABI library loads and exports every symbol the header
declares, the product path fails loudly without a GPU (no CPU fallback), host-side logic
(quadrature nodes, sigma planning, checkpoint batching, energy sharding over 2 gloo ranks)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from gaunegf_b200 import synthetic as sy


def test_library_exports_every_header_symbol():
    from gaunegf_b200 import _native
    from gaunegf_b200.build import build
    build()
    hdr = open(os.path.join(ROOT, "include", "gaunegf_b200.h")).read()
    declared = set(re.findall(r"\b(gnb_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = _native.load_library()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/gaunegf_b200.h but not exported"
    assert declared == set(_native.SIGNATURES), declared ^ set(_native.SIGNATURES)
    assert b"sm_100a" in lib.gnb_version()
    # developer switches / probes live in their own header and are not part of the drop-in ABI; nothing else is exported
    dev = open(os.path.join(ROOT, "include", "gaunegf_b200_dev.h")).read()
    dev_declared = set(re.findall(r"\b(gnb_dev_[a-z0-9_]+)\s*\(", dev))
    assert dev_declared == set(_native.DEV_SIGNATURES), dev_declared ^ set(_native.DEV_SIGNATURES)
    out = subprocess.run(["nm", "-D", "--defined-only", _native.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (gnb_[a-z0-9_]+)$", out, flags=re.M))
    assert exported == declared | dev_declared, exported ^ (declared | dev_declared)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gaunegf_b200 import transport as tr
    F, S, s1, s2 = sy.chain(8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        tr.cohTrans([0.0], F, S, s1, s2)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "gaunegf_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("the oracle", "").lower() or f == "synthetic.py" or \
                    not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f


def test_ant_points_nested_and_ratio():
    from gaunegf_b200.density import getANTPoints, fermi
    x2, w2 = getANTPoints(6)
    x1, _ = getANTPoints(2)
    assert len(x2) == 6 and np.isin(np.round(x1, 14), np.round(x2, 14)).all()
    x3, w3 = getANTPoints(18)
    old = np.isin(np.round(x3, 14), np.round(x2, 14))
    assert abs(np.sum(w3[old]) / np.sum(w2) - 1 / 3) < 1e-12
    xs, ws = getANTPoints(162)
    assert abs(np.sum(ws * np.exp(-xs ** 2)) - 1.4936482656248540) < 1e-9
    assert list(fermi(np.array([1 + 1j, -1 + 1j, 0, 1j]), 0, 0)) == [0, 1, 1, 0]   # lexicographic complex <=


def test_adaptive_driver_logic():
    from gaunegf_b200.density import integratePointsAdaptiveANT
    calls = []

    def cp(x, w):
        calls.append(len(x))
        return np.array([[np.sum(w * np.cos(x))]])
    import io, contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        val = integratePointsAdaptiveANT(cp, tol=1e-10)
    assert abs(val[0, 0] - 2 * np.sin(1.0)) < 1e-9
    assert calls[:4] == [2, 4, 12, 36]            # only the NEW nodes of each level are evaluated


def test_sigma_planning():
    from gaunegf_b200.sigma_plan import ArrayPlan, ObjectPlan, DESC, DENSE_CONST, DENSE_CALL
    from gaunegf_b200.surfGTester import surfGTest
    from gaunegf_b200.surfG1D import surfG
    N = 12
    s1 = np.zeros(N, complex); s1[0] = -0.1j
    s2 = np.zeros(N, complex); s2[-2:] = -0.1j
    p = ArrayPlan([s1, s2], N)
    assert p.kind == DESC and list(p.inds[0]) == [0] and list(p.inds[1]) == [10, 11]
    assert ArrayPlan([np.full(N, -0.1j), s2], N).kind == DENSE_CONST
    with pytest.raises(ValueError):
        ArrayPlan([np.zeros(5), np.zeros(5)], N)
    F, S = sy.hermitian_pair(N, 0)
    assert ObjectPlan(surfGTest(F, S, [[0, 1], [10, 11]], -0.1j), N).kind == DENSE_CONST
    Fx, Sx, inds, taus = sy.lead_device_lead(4, 8, 2)
    assert ObjectPlan(surfG(Fx, Sx, inds, taus, eta=0.01), Fx.shape[0]).kind == DESC

    class Mock:
        def sigmaTot(self, E): return np.zeros((N, N), complex)
        def sigma(self, E, i): return np.zeros((N, N), complex)
    assert ObjectPlan(Mock(), N).kind == DENSE_CALL


def test_checkpoint_block_cadence():
    from gaunegf_b200.transport import _blocks
    rem = np.arange(23)
    blocks = _blocks(rem, "x.npz", 5)
    assert [len(b) for b in blocks] == [1, 5, 5, 5, 5, 2]       # writes after idx 0, 5, 10, ... like the reference
    assert np.array_equal(np.concatenate(blocks), rem)
    assert [len(b) for b in _blocks(rem, None, 5)] == [23]
    assert _blocks(np.array([], dtype=int), "x", 5) == []


def test_surfg_setF_and_bethe_extended_matrices():
    from gaunegf_b200.surfG1D import surfG
    from gaunegf_b200.surfGBethe import surfGBAt
    Fx, Sx, inds, taus = sy.lead_device_lead(4, 8, 2)
    g = surfG(Fx, Sx, inds, taus, eta=0.01)
    F2 = Fx + 0.01 * np.eye(len(Fx))
    g.setF(F2)
    assert np.allclose(g.F[np.ix_(inds[0], inds[0])], F2[np.ix_(taus[0], taus[0])])
    assert np.allclose(g.aList[0], g.F[np.ix_(inds[0], inds[0])])
    G = np.load(os.path.join(ROOT, "tests", "golden", "cfg5_bethe.npz"))
    at = surfGBAt(G["H"][0], G["Slist"][0], G["Vlist"][0], 1e-4)
    assert at.F.shape == (117, 117) and np.allclose(at.F, at.F.conj().T) and np.allclose(at.S, at.S.T)


GLOO_SCRIPT = r"""
import os, sys
import numpy as np
sys.path.insert(0, os.environ["GNB_ROOT"])
import torch.distributed as dist
dist.init_process_group("gloo")
from gaunegf_b200 import parallel
rank, world = parallel.dist_info()
assert world == 2
N, M = 6, 11
E = np.linspace(-1, 1, M) + 0.1j
w = np.linspace(0.5, 1.5, M) * (1 + 0.2j)
A0 = np.arange(N * N).reshape(N, N) * 0.01
def f(e): return np.linalg.inv(e * np.eye(N) - A0)
seen = []
def partial(El, wl, out):
    seen.append(len(El))
    return sum((wk * f(ek) for ek, wk in zip(El, wl)), np.zeros((N, N), complex))
tot = parallel.sharded_matrix_sum(N, E, w, partial)
ref = sum(wk * f(ek) for ek, wk in zip(E, w))
assert np.max(np.abs(tot - ref)) < 1e-12 * np.max(np.abs(ref)), "matrix sum"
assert seen == [len(parallel.shard_indices(M, rank, world))]
T = parallel.sharded_per_energy(E.real, lambda El: El ** 2)
assert np.allclose(T, E.real ** 2)
D = parallel.sharded_per_energy(E.real, lambda El: np.outer(El, np.arange(3.0)), width=3)
assert np.allclose(D, np.outer(E.real, np.arange(3.0)))
# fewer energies than ranks: one rank gets an empty slice
one = parallel.sharded_matrix_sum(N, E[:1], w[:1], partial)
assert np.max(np.abs(one - w[0] * f(E[0]))) < 1e-12
# several weighted sums in one batch (the nested levels of the adaptive quadrature): every segment is sharded on its own,
# ragged and empty segments included, one all-reduce for all of them
ends = [3, 3, 4, 11]
def partial_seg(El, wl, local_ends, out):
    res, lo = [], 0
    for hi in local_ends:
        res.append(sum((wk * f(ek) for ek, wk in zip(El[lo:hi], wl[lo:hi])), np.zeros((N, N), complex)))
        lo = int(hi)
    return np.array(res)
sums = parallel.sharded_matrix_sums(N, E, w, ends, partial_seg)
lo = 0
for s_, hi in enumerate(ends):
    ref_s = sum((wk * f(ek) for ek, wk in zip(E[lo:hi], w[lo:hi])), np.zeros((N, N), complex))
    assert np.max(np.abs(sums[s_] - ref_s)) < 1e-12 * max(1.0, np.max(np.abs(ref_s))), ("segment", s_)
    lo = hi
# parallel.set_system: rank 0's full comparison (or any rank's sanity sample) decides for every rank
class FakeCtx:
    device = 0
    def __init__(self, full_says, sample_says):
        self.full_says, self.sample_says, self.uploads, self.asked = full_says, sample_says, 0, []
        self.last_system_upload, self.system_uploads_skipped = 3, 0
    def system_differs(self, F, S, full=True):
        self.asked.append(full)
        return self.full_says if full else self.sample_says
    def set_system_known(self, F, S, changed):
        self.uploads.append(changed)
for full_says, sample_says, expect in ((0, 0, 0), (1, 0, 1), (0, 2, 2), (2, 1, 3), (3, 3, 3)):
    c = FakeCtx(full_says, sample_says)
    c.uploads = []
    parallel.set_system(c, None, None)
    assert c.asked == [rank == 0], "rank 0 compares in full, the others sample"
    assert c.uploads == [expect], (rank, full_says, sample_says, c.uploads)     # every rank acts on the agreed bits
dist.destroy_process_group()
print("rank", rank, "ok")
"""


GLOO_CKPT_SCRIPT = r"""
import os, sys
import numpy as np
sys.path.insert(0, os.environ["GNB_ROOT"])
import torch.distributed as dist
dist.init_process_group("gloo")
from gaunegf_b200 import parallel, transport as tr
rank, world = parallel.dist_info()
calls = []
def fake_batch(F, S, calc, energies, spin):          # stands in for the GPU batch: T(E) = E^2 on this rank's shard
    calls.append(len(energies))
    return parallel.sharded_per_energy(np.asarray(energies), lambda El: El ** 2)
tr._transmission_batch = fake_batch
ck = os.environ["GNB_CKPT"]
E = np.linspace(0.0, 2.0, 23)
T = tr.calculate_transmission(None, None, None, E, checkpoint_file=ck, checkpoint_interval=5)
assert np.allclose(T, E ** 2)
dist.barrier()
d = np.load(ck)                                      # complete, readable file; only rank 0 wrote it
assert np.allclose(d["transmission"], E ** 2) and not os.path.exists(ck + ".tmp.npz")
dist.barrier()
# resume from a partial checkpoint: every rank must agree on what is left (rank 0 reads and broadcasts)
if rank == 0:
    part = E ** 2
    part[9:] = -1
    np.savez(ck, transmission=part, energy_list=E)
dist.barrier()
calls.clear()
T2 = tr.calculate_transmission(None, None, None, E, checkpoint_file=ck, checkpoint_interval=5)
assert np.allclose(T2, E ** 2) and sum(calls) == 23 - 9, calls
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_checkpoint_two_gloo_ranks(tmp_path):
    """under torchrun only rank 0 writes the .npz (atomically) and resume state is broadcast"""
    script = tmp_path / "gloo_ckpt.py"
    script.write_text(GLOO_CKPT_SCRIPT)
    env = dict(os.environ, GNB_ROOT=ROOT, OMP_NUM_THREADS="1", GNB_CKPT=str(tmp_path / "ck.npz"))
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29733", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.count("ok") == 2


def test_bethe_objects_get_a_device_plan():
    """surfGB-shaped objects (ours or the reference's: attributes gList / indsLists / nIndLists) are described to the
    device, including orthonormal (Xi) and spin-expanded ones; the install sequence is transform first, then contacts"""
    import types
    from gaunegf_b200.sigma_plan import ObjectPlan, DESC, DENSE_CALL
    at = types.SimpleNamespace(H=np.eye(9), Slist=[np.zeros((9, 9))] * 12, Vlist=[np.eye(9)] * 12, eta=1e-4)
    inds = [[np.arange(9), np.arange(9, 18)], [np.arange(18, 27)]]
    nil = [[[0, 1], [2]], [[3]]]
    log = []

    class Ctx:
        N = 60
        def sigma_clear(self): log.append("clear")
        def sigma_set_transform(self, n, Xi, mode): log.append(("xform", n, Xi is not None, mode))
        def sigma_add_bethe(self, i, nb, *a): log.append(("bethe", len(i), nb))
    ref = types.SimpleNamespace(gList=[at, at], indsLists=inds, nIndLists=nil, Xi=np.eye(30), Sdict={"sss": 0}, spin="g")
    plan = ObjectPlan(ref, 60)
    assert plan.kind == DESC
    plan.install(Ctx())
    assert log == ["clear", ("xform", 30, True, 2), ("bethe", 18, [[0, 1], [2]]), ("bethe", 9, [[3]])]
    log.clear()
    ref.Sdict, ref.spin = {"sss": 0.1}, "r"
    Ctx.N = 30
    ObjectPlan(ref, 30).install(Ctx())
    assert log[0] == "clear" and log[1][0] == "bethe"          # plain case: no transform
    assert ObjectPlan(types.SimpleNamespace(sigma=None, sigmaTot=None), 30).kind == DENSE_CALL


def test_energy_sharding_two_gloo_ranks(tmp_path):
    script = tmp_path / "gloo_shard.py"
    script.write_text(GLOO_SCRIPT)
    env = dict(os.environ, GNB_ROOT=ROOT, OMP_NUM_THREADS="1")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29731", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.count("ok") == 2


def test_n2_fermi_search_host_logic(golden, monkeypatch):
    """the grid fits and Fermi searches are host root finding over the GPU drivers.  With the
    drivers replaced by the numpy oracle the host logic must reproduce the reference's answers
    (tests/golden/n2_fermi.npz, made by the unmodified reference)."""
    import gaunegf_b200.density as D
    from gaunegf_b200.surfGTester import surfGTest
    from oracle import negf_oracle as O
    from n2_cases import run_cases, compare
    monkeypatch.setattr(D, "GrInt", O.GrInt)
    monkeypatch.setattr(D, "GrLessInt", O.GrLessInt)
    # the speculative multi-level batches of the adaptive quadrature: each level is its own oracle integral
    monkeypatch.setattr(D, "GrIntLevels", lambda F, S, g, levels: [O.GrInt(F, S, g, E, w) for E, w in levels])
    monkeypatch.setattr(D, "_compute_dos_at_energy", O.compute_dos_at_energy)
    monkeypatch.setattr(D, "inv", np.linalg.inv)           # utils.inv (GPU) of the eigenvalue estimates
    compare(run_cases(D, surfGTest, None), golden("n2_fermi"), 1e-9)


def test_adaptive_ant_speculation_is_transparent():
    """integratePointsAdaptiveANT with a `levels` batch evaluator (several nested levels per GPU batch) returns the same
    value, visits the same levels and prints the same text as the level-by-level reference loop (density.py:211-273)."""
    import contextlib, io
    import gaunegf_b200.density as D
    rng = np.random.default_rng(3)
    A = rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4))

    def f(x, w):                                   # a smooth matrix-valued integrand: converges at N = 54 or 162
        return sum(wk * np.linalg.inv(np.eye(4) * (2.5 + xk) + 0.1 * A) for xk, wk in zip(x, w))

    calls = []

    def g(x, w):
        calls.append(("point", len(x)))
        return f(x, w)

    def levels(pairs):
        calls.append(("levels", [len(x) for x, _ in pairs]))
        return [f(x, w) for x, w in pairs]

    out0, out1 = io.StringIO(), io.StringIO()
    with contextlib.redirect_stdout(out0):
        ref = D.integratePointsAdaptiveANT(g, tol=1e-9)
    seq = list(calls)
    calls.clear()
    g.levels = levels
    with contextlib.redirect_stdout(out1):
        spec = D.integratePointsAdaptiveANT(g, tol=1e-9)
    assert np.array_equal(ref, spec) and out0.getvalue() == out1.getvalue()
    assert all(kind == "levels" for kind, _ in calls) and len(calls) < len(seq)
    assert calls[0] == ("levels", [2, 4, 12])
    # maxN cuts the speculation, too
    calls.clear()
    with contextlib.redirect_stdout(io.StringIO()):
        D.integratePointsAdaptiveANT(g, tol=0.0, maxN=54)
    assert [c[1] for c in calls] == [[2, 4, 12], [36]]


def test_n4_analytic_density(golden):
    """closed-form density for constant self-energies + its bisection (host numpy)"""
    import contextlib, io
    from gaunegf_b200 import density as D
    G = golden("n2_fermi")
    V, Vc, Dv, Gam = sy.analytic_density_case(30, seed=5)
    P = D.density(V, Vc, Dv, Gam, -50.0, 0.1)
    assert np.max(np.abs(P - G["n4_density"])) < 1e-10 * np.max(np.abs(G["n4_density"]))
    with contextlib.redirect_stdout(io.StringIO()):
        mu = D.bisectFermi(V, Vc, Dv, Gam, 12.0)
    assert abs(mu - G["n4_bisectFermi"][0]) < 1e-9
