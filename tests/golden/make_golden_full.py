"""BASELINE-size golden vectors for cfg 4 (surfG1D, 128-orbital lead cells, N = 768) and cfg 5 (Bethe-lattice
contacts, N = 2048, GrLessInt / densityGridN), produced by running the UNMODIFIED reference
(/root/reference/gauNEGF) under the numpy-backed jax shim (oracle/refshim).

Run in the build container only:   python tests/golden/make_golden_full.py [cfg4] [cfg5] [bethe_xi]
Outputs (small: sampled entries + traces + norms, never full N x N matrices):
  cfg4_full.npz   iteration counts of the reference's fixed point for 16 energies x 2 contacts (read from the final
                  state the reference's own lax.while_loop call returns), sampled g / Sigma entries, T(E)
  cfg5_full.npz   GrLessInt (ind = None and -1) and densityGridN at N = 2048 with contacts built by the reference's own
                  surfGB constructor from a mock Gaussian `bar`
  bethe_xi.npz    orthonormal-lattice (Xi Sigma Xi, surfGBethe.py:530-533) and spin-expanded ('u', 'g': :536-539)
                  surfGB objects on a small system
"""
import contextlib
import io
import os
import sys
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.refload import load_reference  # noqa: E402
from gaunegf_b200 import synthetic as sy    # noqa: E402

R = load_reference()
tr, de, it, s1d, sgb = R["transport"], R["density"], R["integrate"], R["surfG1D"], R["surfGBethe"]


def quiet(f, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return f(*a, **k)


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in arrs.items()})
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB", flush=True)


def samples(N, n, seed):
    rng = np.random.default_rng(seed)
    return rng.integers(0, N, n), rng.integers(0, N, n)


def mock_bar(n_dev_funcs, Lz, dev_z=(0.0, 0.0)):
    """Two 3-atom Au(111) triangles (nn 2.88 A) Lz apart with 9 basis functions per atom (the order the reference
    sorts by, surfGBethe.py:131-132) around device atoms that carry the remaining basis functions."""
    ang = 1 / sgb.bohr_to_ang
    a = 2.88
    tri = np.array([[0, 0, 0], [a, 0, 0], [a / 2, a * np.sqrt(3) / 2, 0]])
    cen = np.array([a / 2, a / (2 * np.sqrt(3)), 0.0])
    ndev_atoms = (n_dev_funcs + 3) // 4
    zs = np.linspace(dev_z[0], dev_z[1], ndev_atoms)
    dev = np.array([cen + [0.0, 0.0, z] for z in zs])
    coords = np.vstack([tri, dev, tri + [0, 0, Lz]])
    typ9 = [0, 1001, 1002, 1003, 2001, 2002, 2003, 2004, 2005]
    ibfatm, ibftyp, left = [], [], n_dev_funcs
    for at in range(1, len(coords) + 1):
        if 3 < at <= 3 + ndev_atoms:
            k = min(4, left)
            left -= k
            ibfatm += [at] * k
            ibftyp += [0, 1001, 1002, 1003][:k]
        else:
            ibfatm += [at] * 9
            ibftyp += typ9
    bar = types.SimpleNamespace(ibfatm=np.array(ibfatm), ibftyp=np.array(ibftyp), c=(coords * ang).ravel())
    n_at = len(coords)
    return bar, [[1, 2, 3], [n_at - 2, n_at - 1, n_at]]


def bethe_parts(gB):
    parts = dict(H=[np.array(x.H) for x in gB.gList], Slist=[np.array(x.Slist) for x in gB.gList],
                 Vlist=[np.array(x.Vlist) for x in gB.gList], fermi=float(gB.gList[0].fermi),
                 indsLists=np.array(gB.indsLists))
    parts["nInd_len"] = np.array([[len(x) for x in c] for c in gB.nIndLists])
    parts["nInd_flat"] = np.array([v for c in gB.nIndLists for x in c for v in x], dtype=int)
    return parts


def build_surfGB(F, S, contacts, bar, latdir, latFile, **kw):
    cwd = os.getcwd()
    os.chdir(latdir)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            return sgb.surfGB(F, S, contacts, bar, latFile=latFile, **kw)
    finally:
        os.chdir(cwd)


# ---- cfg 4 at BASELINE size ------------------------------------------------------------------------
def cfg4():
    import jax.lax as lax                       # the shim's module: the reference calls lax.while_loop (surfG1D.py:287)
    final = []
    orig = lax.while_loop

    def recording(cond, body, init):            # the reference source stays unmodified: only its final state is read
        out = orig(cond, body, init)
        final.append((int(out[0]), float(out[1])))
        return out

    F, S, li, taus = sy.lead_device_lead(128, 512, seed=2, s_off=0.0)
    E_all = np.linspace(-1, 1, 256)
    idx = np.arange(3, 256, 16)                 # 16 of the 256 energies of the BASELINE grid
    E = E_all[idx]
    out = {"idx": idx, "E": E}
    ii, jj = samples(128, 96, 7)
    for eta, tag in ((1e-4, "b"), (0.02, "a")):
        g = s1d.surfG(F, S, [list(i) for i in li], [list(t) for t in taus], eta=eta)
        iters = np.zeros((len(E), 2), dtype=int)
        diffs = np.zeros((len(E), 2))
        gs = np.zeros((len(E), 2, 96), dtype=complex)
        gfro = np.zeros((len(E), 2))
        gtr = np.zeros((len(E), 2), dtype=complex)
        ss = np.zeros((len(E), 2, 96), dtype=complex)
        sfro = np.zeros((len(E), 2))
        lax.while_loop = recording
        try:
            t0 = time.time()
            for k, e in enumerate(E):
                for c in (0, 1):
                    final.clear()
                    gm = np.array(g.g(e, c, 1e-5, 0.1))
                    iters[k, c], diffs[k, c] = final[-1]
                    gs[k, c], gfro[k, c], gtr[k, c] = gm[ii, jj], np.linalg.norm(gm), np.trace(gm)
                    sm = np.array(g.sigma(e, c))[np.ix_(li[c], li[c])]
                    ss[k, c], sfro[k, c] = sm[ii, jj], np.linalg.norm(sm)
                print(f"cfg4 eta={eta} E[{k}]={e:+.4f} iters={iters[k]} diff={diffs[k]} ({time.time() - t0:.0f}s)", flush=True)
        finally:
            lax.while_loop = orig
        out.update({f"iters_{tag}": iters, f"diffs_{tag}": diffs, f"g_samp_{tag}": gs, f"g_fro_{tag}": gfro,
                    f"g_tr_{tag}": gtr, f"sig_samp_{tag}": ss, f"sig_fro_{tag}": sfro})
        n_t = 8 if tag == "b" else 4            # T(E): the reference re-evaluates both fixed points 4x per energy
        out[f"T_{tag}"] = quiet(tr.cohTransE, E[:n_t], F, S, g)
        print(f"cfg4 eta={eta} T done ({time.time() - t0:.0f}s)", flush=True)
    save("cfg4_full", ii=ii, jj=jj, **out)


# ---- cfg 5 at BASELINE size ------------------------------------------------------------------------
def cfg5():
    N = 2048
    bar, contacts = mock_bar(N - 54, Lz=24.0, dev_z=(9.0, 15.0))
    assert len(bar.ibfatm) == N
    F, S = sy.hermitian_pair(N, seed=3)
    t0 = time.time()
    gB = build_surfGB(F, S, contacts, bar, "/root/reference", "Au", eta=1e-4)
    print(f"cfg5 surfGB built ({time.time() - t0:.0f}s)", flush=True)
    mu = float(gB.gList[0].fermi)
    ii, jj = samples(N, 512, 11)
    Eg = mu + np.array([-0.2, -0.05, 0.1, 0.23])
    wg = np.array([0.4, 0.3, 0.2, 0.1])
    out = dict(bethe_parts(gB), eta=1e-4, N=N, ii=ii, jj=jj, Eg=Eg, wg=wg)
    for tag, ind in (("last", -1), ("all", None)):
        P = np.array(it.GrLessInt(F, S, gB, Eg, wg, ind))
        out.update({f"GL_{tag}_samp": P[ii, jj], f"GL_{tag}_diag": np.diag(P), f"GL_{tag}_fro": np.linalg.norm(P),
                    f"GL_{tag}_trS": np.trace(P @ S)})
        print(f"cfg5 GrLessInt ind={ind} ({time.time() - t0:.0f}s)", flush=True)
    P = np.array(quiet(de.densityGridN, F, S, gB, mu - 0.25, mu + 0.25, -1, 8, 0.0, False))
    out.update(PgN_samp=P[ii, jj], PgN_diag=np.diag(P), PgN_fro=np.linalg.norm(P), PgN_trS=np.trace(P @ S))
    zc = np.array([mu - 3.0 + 2.0j, mu + 0.5j])
    wc = np.array([0.7 - 0.2j, 0.3j])
    P = np.array(it.GrInt(F, S, gB, zc, wc))
    out.update(zc=zc, wc=wc, GI_samp=P[ii, jj], GI_diag=np.diag(P), GI_fro=np.linalg.norm(P))
    print(f"cfg5 done ({time.time() - t0:.0f}s)", flush=True)
    save("cfg5_full", **out)


# ---- orthonormal-lattice and spin-expanded surfGB ------------------------------------------------------
def spin_system(F, S, sp, seed=17):
    """2N x 2N spin-expanded F, S ('u': block order, 'g': spinor order) with the spin degeneracy broken"""
    F2 = np.kron(np.eye(2), F) if sp == "u" else np.kron(F, np.eye(2))
    S2 = np.kron(np.eye(2), S) if sp == "u" else np.kron(S, np.eye(2))
    D = np.random.default_rng(seed).standard_normal(F2.shape) * 0.02
    return F2 + (D + D.T) / 2, S2


def bethe_xi():
    # Au2.bethe is the reference's own orthonormal parameter file (all overlap integrals zero): Sdict['sss'] == 0
    # selects the Xi Sigma Xi branch (surfGBethe.py:530-533)
    bar, contacts = mock_bar(8, Lz=9.0, dev_z=(3.8, 5.2))
    Nb = len(bar.ibfatm)
    F, S = sy.hermitian_pair(Nb, seed=3)
    out = {"Nb": Nb}
    gO = build_surfGB(F, S, contacts, bar, "/root/reference", "Au2", eta=1e-4)
    mu = float(gO.gList[0].fermi)
    E = np.array([mu - 2.0, mu, mu + 1.5])
    out.update({"o_" + k: v for k, v in bethe_parts(gO).items()})
    out["o_Xi"] = np.array(gO.Xi)
    out["E_o"] = E
    out["o_sig0"] = np.array(gO.sigma(E[0], 0))
    out["o_sigT"] = np.array(gO.sigmaTot(E[1]))
    out["o_T"] = quiet(tr.cohTransE, E, F, S, gO)
    out["o_dos"] = quiet(tr.DOSE, E, F, S, gO)[0]
    out["o_GL"] = np.array(it.GrLessInt(F, S, gO, E, np.array([0.2, 0.5, 0.3]), -1))
    out["o_GI"] = np.array(it.GrInt(F, S, gO, E + 0.3j, np.array([0.2, 0.5j, 0.3])))
    # spin-expanded contacts: F, S are 2N x 2N, Sigma is kron-expanded (surfGBethe.py:536-539)
    ii, jj = samples(2 * Nb, 256, 5)
    out["ii"], out["jj"] = ii, jj
    for sp in ("u", "g"):
        F2, S2 = spin_system(F, S, sp)
        gS = build_surfGB(F2, S2, contacts, bar, "/root/reference", "Au", spin=sp, eta=1e-4)
        mus = float(gS.gList[0].fermi)
        Es = np.array([mus - 1.0, mus + 0.4])
        out.update({f"{sp}_" + k: v for k, v in bethe_parts(gS).items()})
        out[f"E_{sp}"] = Es
        sT = np.array(gS.sigmaTot(Es[0]))
        out[f"{sp}_sigT_samp"], out[f"{sp}_sigT_diag"], out[f"{sp}_sigT_fro"] = sT[ii, jj], np.diag(sT), np.linalg.norm(sT)
        for nm, P in (("GI", np.array(it.GrInt(F2, S2, gS, Es + 0.2j, np.array([1.0, -0.5j])))),
                      ("GL", np.array(it.GrLessInt(F2, S2, gS, Es, np.array([0.6, 0.4]), 0)))):
            out[f"{sp}_{nm}_samp"], out[f"{sp}_{nm}_diag"], out[f"{sp}_{nm}_fro"] = P[ii, jj], np.diag(P), np.linalg.norm(P)
        out[f"{sp}_dos"] = quiet(tr.DOSE, Es, F2, S2, gS)[0]
    save("bethe_xi", **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["cfg4", "cfg5", "bethe_xi"]
    for w in which:
        {"cfg4": cfg4, "cfg5": cfg5, "bethe_xi": bethe_xi}[w]()
    print("done")
