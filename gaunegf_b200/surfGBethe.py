"""Bethe-lattice contacts — the Sigma(E) part of gauNEGF/surfGBethe.py (surfGBAt.sigmaK/sigma/
sigmaTot/DOS: :958-1155; surfGB.sigma/sigmaTot: :479-575) on the B200.

The geometry detection / Slater-Koster construction of the reference's surfGB constructor
(surfGBethe.py:106-477) needs a Gaussian `bar` object and is one-time setup, out of scope for this
path: build the 9x9 onsite matrix H and the 12 hopping / overlap blocks with
the reference (or any Slater-Koster code) and hand them to surfGBAt / surfGB.from_parts.
"""
import numpy as np

from ._native import default_context
from .config import (ETA, TEMPERATURE, SURFACE_GREEN_CONVERGENCE, BETHE_MAX_ITER, BETHE_MIXING, ENERGY_MIN,
                     FERMI_CALCULATION_TOL)

dim = 9


class surfGBAt:
    """One Bethe-lattice atom: H (9x9), Slist / Vlist (12 x 9x9) — same constructor as the reference."""

    def __init__(self, H, Slist, Vlist, eta, T=TEMPERATURE):
        assert np.shape(H) == (dim, dim), f"Error with H dim, should be {dim}x{dim}"
        for S_, V_ in zip(Slist, Vlist):
            assert np.shape(S_) == (dim, dim), f"Error with S dim, should be {dim}x{dim}"
            assert np.shape(V_) == (dim, dim), f"Error with F dim, should be {dim}x{dim}"
        self.H = np.array(H)
        self.Slist = [np.array(s) for s in Slist]
        self.Vlist = [np.array(v) for v in Vlist]
        self.NN = len(Slist)
        assert self.NN == 12, "Error: surfGBAt only implemented for FCC using 12 NN"
        self.eta = eta
        self.T = T
        self.fermi = None
        self.last_iters = {}
        self.updateH()

    def updateH(self, fermi=None):
        """shift to a new Fermi level and rebuild the 13-site extended F / S (surfGBethe.py:914-955)"""
        if fermi is not None and self.fermi is not None and fermi != self.fermi:
            d = fermi - self.fermi
            self.H = self.H + d * np.eye(dim)
            for j, S_ in enumerate(self.Slist):
                self.Vlist[j] = self.Vlist[j] + d * S_
            self.fermi = fermi
        n = self.NN
        H0x = np.kron(np.eye(n + 1), self.H).astype(complex)
        S0x = np.eye(dim * (n + 1), dtype=complex)
        for i in range(n):
            sl = slice(i * dim, (i + 1) * dim)
            S0x[-dim:, sl] = self.Slist[i]
            S0x[sl, -dim:] = self.Slist[i].T
            H0x[-dim:, sl] = self.Vlist[i]
            H0x[sl, -dim:] = self.Vlist[i].conj().T
        self.F, self.S = H0x, S0x

    def _raw(self, E, which, conv, mix):
        ctx = default_context()
        ctx.set_system(np.eye(dim), np.eye(dim))
        ctx.sigma_clear()
        ctx.sigma_add_bethe(np.arange(dim), [[]], self.H, self.Slist, self.Vlist, self.eta, conv, mix, BETHE_MAX_ITER)
        Es = np.atleast_1d(np.asarray(E, dtype=complex))
        out, iters, diffs = ctx.sigma_eval(0, which, Es, (12 if which == 1 else 9, dim, dim))
        for e, it, d in zip(Es, iters, diffs):
            self.last_iters[(which, complex(e))] = (int(it), float(d))
        return out[0] if np.ndim(E) == 0 else out

    def sigmaK(self, E, conv=SURFACE_GREEN_CONVERGENCE, mix=BETHE_MIXING):
        """12 bulk direction self-energies (surfGBethe.py:958-1030)"""
        return self._raw(E, 1, conv, mix)

    def sigma(self, E, conv=SURFACE_GREEN_CONVERGENCE, mix=BETHE_MIXING):
        """9 surface direction self-energies (surfGBethe.py:1032-1108)"""
        return self._raw(E, 2, conv, mix)

    def setF(self, F, mu1, mu2):
        pass

    _gnb_batched = True        # sigmaTot accepts an array of energies (one GPU call for the whole batch)

    def sigmaTot(self, E, conv=SURFACE_GREEN_CONVERGENCE):
        """self-energy of the 13-site extended system (surfGBethe.py:1110-1136): every neighbour site carries the
        bulk total minus the direction pointing back at the centre; the centre itself carries none"""
        Es = np.atleast_1d(np.asarray(E, dtype=complex))
        sigK = np.asarray(self.sigmaK(Es, conv))                  # (M, 12, 9, 9)
        n = self.NN
        sig = np.zeros((len(Es), (n + 1) * dim, (n + 1) * dim), dtype=complex)
        tot = np.sum(sigK, axis=1)
        for k in range(n):
            sig[:, k * dim:(k + 1) * dim, k * dim:(k + 1) * dim] = tot - sigK[:, (k + 6) % 12]
        return sig[0] if np.ndim(E) == 0 else sig

    def calcFermi(self, ne, fGuess=5, tol=FERMI_CALCULATION_TOL):
        """Fermi level of the Bethe lattice holding `ne` electrons on the centre atom's 9 orbitals
        (surfGBethe.py:1159-1186): density.getFermiContact on the extended system F, S"""
        from .density import getFermiContact
        self.fermi = getFermiContact(self, ne, tol, ENERGY_MIN, 1000, T=self.T, nOrbs=dim)
        return self.fermi

    def DOS(self, E):
        """bulk DOS of the Bethe lattice (surfGBethe.py:1139-1155); the 9x9 inverse runs on the GPU"""
        sig = np.sum(np.asarray(self.sigma(E)), axis=0)
        ctx = default_context()
        Gr = ctx.inverse_batch((E - 1j * self.eta) * np.eye(dim) - self.H - sig)
        return -np.trace(Gr).imag / np.pi


_SPIN_MODE = {'r': 0, 'u': 1, 'ro': 1, 'g': 2}


def is_bethe_object(g):
    """a surfGB-shaped object: ours, or one built by the reference's constructor (surfGBethe.py:106-248)"""
    return (all(hasattr(g, a) for a in ("gList", "indsLists", "nIndLists"))
            and all(all(hasattr(a, x) for x in ("H", "Slist", "Vlist", "eta")) for a in g.gList))


def bethe_orthonormal(g):
    """the reference decides by its parameter file: Sdict['sss'] == 0 selects Xi Sigma Xi (surfGBethe.py:530-533)"""
    if hasattr(g, "orthonormal"):
        return bool(g.orthonormal)
    return hasattr(g, "Sdict") and g.Sdict.get("sss", 1.0) == 0


def install_bethe(ctx, g, conv=SURFACE_GREEN_CONVERGENCE, spin_mode=None):
    """device description of a surfGB-shaped object: one Bethe contact per contact of g; with an orthonormal
    parameter set and / or a spin-expanded system also the Sigma -> expand(Xi Sigma Xi) transform
    (surfGBethe.py:529-539).  spin_mode overrides the object's own (transport's spinor -> block reordering)."""
    orth = bethe_orthonormal(g)
    mode = _SPIN_MODE[getattr(g, "spin", 'r')] if spin_mode is None else spin_mode
    if orth or mode:
        n = ctx.N // 2 if mode else ctx.N
        ctx.sigma_set_transform(n, np.asarray(g.Xi) if orth else None, mode)
    for at, inds, nbs in zip(g.gList, g.indsLists, g.nIndLists):
        ctx.sigma_add_bethe(np.concatenate([np.asarray(i) for i in inds]), [list(x) for x in nbs], np.asarray(at.H),
                            np.asarray(at.Slist), np.asarray(at.Vlist), at.eta, conv, BETHE_MIXING, BETHE_MAX_ITER)


class surfGB:
    """Device with Bethe-lattice contacts (surfG protocol).  Build with from_parts()."""

    def __init__(self, F, S, contacts=None, bar=None, latFile='Au', spin='r', eta=ETA, T=TEMPERATURE):
        raise NotImplementedError(
            "surfGB(F, S, contacts, bar, ...) needs Gaussian's `bar` geometry and the Slater-Koster setup of the "
            "reference (surfGBethe.py:106-477), which is outside this path; use surfGB.from_parts(...) with the "
            "H / Slist / Vlist / index lists that setup produces.")

    @classmethod
    def from_parts(cls, F, S, gList, indsLists, nIndLists, Xi=None, orthonormal=False, spin='r', eta=ETA):
        self = object.__new__(cls)
        self.F, self.S, self.N = F, S, len(F) if spin == 'r' else len(F) // 2
        self.gList = gList
        self.indsLists = [[np.asarray(a) for a in c] for c in indsLists]
        self.nIndLists = [[list(a) for a in c] for c in nIndLists]
        self.Xi, self.orthonormal, self.spin, self.eta = Xi, orthonormal, spin, eta
        return self

    def _gnb_install(self, ctx, conv=SURFACE_GREEN_CONVERGENCE, spin_mode=None):
        install_bethe(ctx, self, conv, spin_mode)

    def sigma(self, E, i, conv=SURFACE_GREEN_CONVERGENCE):
        """N x N self-energy of contact i (surfGBethe.py:479-542)"""
        surf = np.asarray(self.gList[i].sigma(E, conv))
        sig = np.zeros((self.N, self.N), dtype=complex)
        for nInds, Finds in zip(self.nIndLists[i], self.indsLists[i]):
            sigAtom = np.sum(surf[:9], axis=0)
            for n in nInds:
                sigAtom = sigAtom - surf[n]
            sig[np.ix_(Finds, Finds)] = sigAtom
        if self.orthonormal:
            sig = self.Xi @ sig @ self.Xi
        if self.spin in ('u', 'ro'):
            sig = np.kron(np.eye(2), sig)
        elif self.spin == 'g':
            sig = np.kron(sig, np.eye(2))
        return sig

    def sigmaTot(self, E, conv=SURFACE_GREEN_CONVERGENCE):
        return sum(self.sigma(E, i, conv) for i in range(len(self.indsLists)))

    def getSigma(self, Elist=(None, None), conv=SURFACE_GREEN_CONVERGENCE):
        E0 = self.gList[0].fermi if Elist[0] is None else Elist[0]
        E1 = self.gList[-1].fermi if Elist[1] is None else Elist[1]
        return (self.sigma(E0, 0, conv), self.sigma(E1, -1, conv))

    def updateFermi(self, i, Ef):
        self.gList[i].updateH(Ef)

    def setF(self, F, muL, muR):
        self.F = F
        if self.gList[0].fermi != muL:
            self.updateFermi(0, muL)
        if self.gList[-1].fermi != muR:
            self.updateFermi(-1, muR)
