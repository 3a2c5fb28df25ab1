"""Generate the committed golden vectors by running the UNMODIFIED reference (wliverno/GauNEGF,
/root/reference) under the numpy-backed jax shim (oracle/refshim) on seeded synthetic inputs.

Run in the build container only:   python tests/golden/make_golden.py
Outputs: tests/golden/*.npz (small: full matrices only for N <= 48, summaries above that).
The reference has no stored golden values of its own, so these fixtures are what
pins oracle/negf_oracle.py and the CUDA path to the reference's actual behaviour.
"""
import contextlib
import io
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.refload import load_reference  # noqa: E402
from gaunegf_b200 import synthetic as sy    # noqa: E402

R = load_reference()
tr, de, it, s1d, sgb, sgt = (R["transport"], R["density"], R["integrate"], R["surfG1D"],
                             R["surfGBethe"], R["surfGTester"])


def quiet(f, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return f(*a, **k)


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in arrs.items()})
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def sample_idx(N, n=64, seed=123):
    rng = np.random.default_rng(seed)
    return rng.integers(0, N, n), rng.integers(0, N, n)


# ---- cfg 1: cohTrans / DOS / current on the 64-orbital chain --------------------------------
F, S, sig1, sig2 = sy.chain(64)
E = np.linspace(-3, 3, 1000)
T = quiet(tr.cohTrans, E, F, S, sig1, sig2)
Ed = np.linspace(-2.5, 2.5, 40)
dos_tot, dos_site = quiet(tr.DOS, Ed, F, S, sig1, sig2)
cur = [quiet(tr.current, F, S, sig1, sig2, 0.0, qV, T=Tk, dE=0.01)
       for qV, Tk in ((0.5, 0.0), (-0.5, 0.0), (0.3, 300.0))]
# block-form (matrix) self-energies on 8-site blocks of a 40-site chain with complex F
F2, S2, _, _ = sy.chain(40)
F2 = F2.astype(complex)
sm1 = np.zeros((40, 40), dtype=complex)
sm2 = np.zeros((40, 40), dtype=complex)
sm1[:8, :8] = -0.1j * np.eye(8) + 0.02 * np.ones((8, 8))
sm2[-8:, -8:] = -0.15j * np.eye(8)
E2 = np.linspace(-2, 2, 50)
Tm = quiet(tr.cohTrans, E2, F2, S2, sm1, sm2)
save("cfg1_chain", E=E, T=T, Ed=Ed, dos_tot=dos_tot, dos_site=dos_site,
     cur_args=np.array([(0.5, 0.0), (-0.5, 0.0), (0.3, 300.0)]), cur=cur, E2=E2, Tm=Tm,
     sm1=sm1, sm2=sm2)

# spin-resolved ('u' block-diagonal and 'g' spinor) on a small random system
Fs, Ss = sy.hermitian_pair(24, seed=5, complex_F=True)
su1, su2 = sy.block_sigma_vectors(12, 3, 0.1)
Es = np.linspace(-1, 1, 7)
Tu, Tu4 = quiet(tr.cohTransSpin, Es, Fs, Ss, su1, su2, 'u')
Tg, Tg4 = quiet(tr.cohTransSpin, Es, Fs, Ss, su1, su2, 'g')
save("cfg1_spin", Es=Es, Tu=Tu, Tu4=Tu4, Tg=Tg, Tg4=Tg4)

# ---- cfg 2: densities with constant sigma (surfGTest) ---------------------------------------
N = 48
F, S = sy.hermitian_pair(N, seed=0)
inds = sy.end_contacts(N, 6)
g = sgt.surfGTest(F, S, [list(inds[0]), list(inds[1])], -0.1j, -0.1j)
Pc = quiet(de.densityComplex, F, S, g, -30.0, 0.0, 1e-4, 0.0)
PcN = quiet(de.densityComplexN, F, S, g, -30.0, 0.0, 54, 300.0, False, 'ant')
PcNl = quiet(de.densityComplexN, F, S, g, -30.0, 0.0, 40, 0.0, False, 'legendre')
PrN = quiet(de.densityRealN, F, S, g, -8.0, 0.0, 64, 0.0, False)
Pr = quiet(de.densityReal, F, S, g, -8.0, 0.0, 1e-2, 0.0)
PgN = quiet(de.densityGridN, F, S, g, -0.25, 0.25, -1, 60, 0.0, False)
PgN0 = quiet(de.densityGridN, F, S, g, 0.25, -0.25, None, 60, 300.0, False)
Pg = quiet(de.densityGrid, F, S, g, -0.25, 0.25, 0, 1e-4, 0.0)
z = np.array([-1.0 + 0.5j, 0.3 + 0.01j, 2.0 + 0j, -0.2 + 2j])
w = np.array([0.3 + 0.1j, 1.0, -0.5j, 0.25])
GI = np.array(it.GrInt(F, S, g, z, w))
GL = np.array(it.GrLessInt(F, S, g, z.real, w.real, 0))
GLn = np.array(it.GrLessInt(F, S, g, z.real, w.real, None))
save("cfg2_density48", Pc=Pc, PcN=PcN, PcNl=PcNl, PrN=PrN, Pr=Pr, PgN=PgN, PgN0=PgN0, Pg=Pg,
     z=z, w=w, GI=GI, GL=GL, GLn=GLn)

N = 256
F, S = sy.hermitian_pair(N, seed=0)
inds = sy.end_contacts(N, 16)
g = sgt.surfGTest(F, S, [list(inds[0]), list(inds[1])], -0.1j, -0.1j)
buf = io.StringIO()
with contextlib.redirect_stdout(buf):
    P = de.densityComplex(F, S, g, -30.0, 0.0, 1e-4, 0.0)
ii, jj = sample_idx(N)
save("cfg2_density256", trPS=np.trace(P @ S), fro=np.linalg.norm(P), ii=ii, jj=jj, samp=P[ii, jj],
     log=np.array(buf.getvalue()))

# ---- cfg 3: N=1024 transmission (16 energies; 0.4 s each on CPU) ----------------------------
N = 1024
F, S = sy.hermitian_pair(N, seed=1)
s1, s2 = sy.block_sigma_vectors(N, 64, 0.1)
E3 = np.linspace(-0.5, 0.5, 16)
T3 = quiet(tr.cohTrans, E3, F, S, s1, s2)
N = 256
F, S = sy.hermitian_pair(N, seed=1)
s1, s2 = sy.block_sigma_vectors(N, 16, 0.1)
E3b = np.linspace(-0.5, 0.5, 64)
T3b = quiet(tr.cohTrans, E3b, F, S, s1, s2)
I3b = quiet(tr.current, F, S, s1, s2, 0.0, 0.4, T=0.0, dE=0.01)
save("cfg3_trans", E3=E3, T3=T3, E3b=E3b, T3b=T3b, I3b=I3b)

# ---- cfg 4: surfG1D Sigma(E) -----------------------------------------------------------------
F, S, inds, taus = sy.lead_device_lead(16, 32, seed=2, s_off=0.05)
E4 = np.linspace(-1, 1, 9)
out = {}
for eta, tag in ((0.05, "a"), (1e-4, "b")):
    g = s1d.surfG(F, S, [list(i) for i in inds], [list(t) for t in taus], eta=eta)
    out["sig0_" + tag] = np.array([np.array(g.sigma(e, 0))[np.ix_(inds[0], inds[0])] for e in E4])
    out["sig1_" + tag] = np.array([np.array(g.sigma(e, 1))[np.ix_(inds[1], inds[1])] for e in E4])
    out["g0_" + tag] = np.array([np.array(g.g(e, 0, 1e-5, 0.1)) for e in E4])
    if tag == "a":
        out["T4"] = quiet(tr.cohTransE, E4, F, S, g)
        zc = np.array([-0.5 + 0.3j, 0.2 + 1.0j, 0.7 + 0.05j])
        out["zc"] = zc
        out["GI4"] = np.array(it.GrInt(F, S, g, zc, np.array([1.0, 0.5j, -0.25])))
        out["GL4"] = np.array(it.GrLessInt(F, S, g, E4[:4], np.ones(4) * 0.1, -1))
        out["dos4"], _ = quiet(tr.DOSE, E4, F, S, g)
save("cfg4_surfg1d", E4=E4, **out)

# ---- cfg 5: Bethe lattice ---------------------------------------------------------------------
# Mock Gaussian `bar`: two 3-atom Au(111) triangles (nn 2.88 A) 9 A apart + 2 light device atoms.
ang = 1 / sgb.bohr_to_ang
a = 2.88
tri = np.array([[0, 0, 0], [a, 0, 0], [a / 2, a * np.sqrt(3) / 2, 0]])
Lz = 9.0
coords = np.vstack([tri, [[a / 2, a / (2 * np.sqrt(3)), Lz / 2 - 0.7], [a / 2, a / (2 * np.sqrt(3)), Lz / 2 + 0.7]],
                    tri + [0, 0, Lz]])
natom = len(coords)
typ9 = [0, 1001, 1002, 1003, 2001, 2002, 2003, 2004, 2005]
ibfatm, ibftyp = [], []
for at in range(1, natom + 1):
    if at in (4, 5):
        ibfatm += [at] * 4
        ibftyp += [0, 1001, 1002, 1003]
    else:
        ibfatm += [at] * 9
        ibftyp += typ9
bar = types.SimpleNamespace(ibfatm=np.array(ibfatm), ibftyp=np.array(ibftyp), c=(coords * ang).ravel())
Nb = len(ibfatm)
Fb, Sb = sy.hermitian_pair(Nb, seed=3)
cwd = os.getcwd()
os.chdir("/root/reference")
try:
    with contextlib.redirect_stdout(io.StringIO()):
        gB = sgb.surfGB(Fb, Sb, [[1, 2, 3], [6, 7, 8]], bar, latFile='Au', eta=1e-4)
finally:
    os.chdir(cwd)
at0 = gB.gList[0]
E5 = np.array([-3.0, float(at0.fermi), 2.5])
parts = dict(H=[np.array(x.H) for x in gB.gList], Slist=[np.array(x.Slist) for x in gB.gList],
             Vlist=[np.array(x.Vlist) for x in gB.gList], fermi=float(at0.fermi), eta=1e-4,
             indsLists=np.array(gB.indsLists), Nb=Nb)
nil = gB.nIndLists
parts["nInd_len"] = np.array([[len(x) for x in c] for c in nil])
parts["nInd_flat"] = np.array([v for c in nil for x in c for v in x], dtype=int)
sigK = np.array([np.array(at0.sigmaK(e, 1e-5, 0.5)) for e in E5])
sigS = np.array([np.array(at0.sigma(e, 1e-5, 0.5)) for e in E5])
sigB0 = np.array([np.array(gB.sigma(e, 0)) for e in E5])
sigBt = np.array([np.array(gB.sigmaTot(e)) for e in E5])
dosB = np.array([float(at0.DOS(e)) for e in E5])
PgB = quiet(de.densityGridN, Fb, Sb, gB, float(at0.fermi) - 0.25, float(at0.fermi) + 0.25, -1, 6, 0.0, False)
save("cfg5_bethe", E5=E5, sigK=sigK, sigS=sigS, sigB0=sigB0, sigBt=sigBt, dosB=dosB, PgB=PgB, **parts)
print("done")
