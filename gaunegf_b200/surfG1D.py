"""1-D chain contacts — drop-in for gauNEGF/surfG1D.py (class surfG, surfG protocol).

The surface Green's function is the reference's DAMPED FIXED POINT (surfG1D.py:260-288), not a
Sancho-Rubio decimation: g0 = inv(A), g <- relax*inv(A - B g B^H) + (1-relax)*g until the largest
element-wise relative change is <= conv (cap 2000).  It runs on the B200 batched over energies
(gnb_sigma.cu); g / sigma / sigmaTot accept a scalar energy like the reference or an array of
energies (returning a stacked result), and last_iters exposes the per-energy iteration counts.
"""
import numpy as np

from ._native import default_context
from .config import ETA, SURFACE_GREEN_CONVERGENCE, SURFACE_RELAXATION_FACTOR, SURFACE_GREEN_MAX_ITER
from .utils import fractional_matrix_power


class surfG:
    def __init__(self, Fock, Overlap, indsList, taus=None, staus=None, alphas=None, aOverlaps=None,
                 betas=None, bOverlaps=None, eta=ETA):
        self.F = np.array(Fock)
        self.S = np.array(Overlap)
        self.X = np.array(fractional_matrix_power(Overlap, -0.5))
        self.indsList = [np.array(inds) for inds in indsList]
        if taus is None:                                  # surfG1D.py:133-134
            taus = [self.indsList[-1], self.indsList[0]]
        taus = [np.array(tau) for tau in taus]
        if len(np.shape(taus[0])) == 1:                   # connection indices: couplings from F / S
            self.tauFromFock = True
            self.tauInds = taus
            self.tauList = [self.F[np.ix_(taus[0], self.indsList[0])], self.F[np.ix_(taus[1], self.indsList[-1])]]
            self.stauList = [self.S[np.ix_(taus[0], self.indsList[0])], self.S[np.ix_(taus[1], self.indsList[-1])]]
        else:
            if staus is None:
                raise ValueError("staus (coupling overlaps) are required when taus are matrices")
            self.tauFromFock = False
            self.tauList = [np.array(tau) for tau in taus]
            self.stauList = [np.array(stau) for stau in staus]
        if alphas is None:
            self.contactFromFock = True
            self.setContacts()
        else:
            self.contactFromFock = False
            self.setContacts(alphas, aOverlaps, betas, bOverlaps)
            self.fermiList = [None] * len(indsList)
        self.eta = eta
        self.num_contacts = len(indsList)
        self.last_iters = {}

    def setContacts(self, alphas=None, aOverlaps=None, betas=None, bOverlaps=None):
        if self.contactFromFock:
            self.aList = [np.array(self.F[np.ix_(inds, inds)]) for inds in self.indsList]
            self.aSList = [np.array(self.S[np.ix_(inds, inds)]) for inds in self.indsList]
            self.bList = [np.array(tau) for tau in self.tauList]          # beta == tau (surfG1D.py:217-218)
            self.bSList = [np.array(stau) for stau in self.stauList]
        else:
            self.aList = [np.array(a) for a in alphas]
            self.aSList = [np.array(a) for a in aOverlaps]
            self.bList = [np.array(b) for b in betas]
            self.bSList = [np.array(b) for b in bOverlaps]

    # ---- device description -------------------------------------------------------------
    def _add_contact(self, ctx, i, conv, relFactor):
        n = len(self.indsList[i])
        for name, m in (("tau", self.tauList[i]), ("stau", self.stauList[i])):
            if np.shape(m) != (n, n):
                raise ValueError(f"contact {i}: {name} has shape {np.shape(m)}; the reference adds t g t^H on the "
                                 f"{n} contact orbitals, which needs len(tau indices) == len(contact indices)")
        ctx.sigma_add_chain1d(self.indsList[i], self.aList[i], self.aSList[i], self.bList[i], self.bSList[i],
                              self.tauList[i], self.stauList[i], self.eta, conv, relFactor, SURFACE_GREEN_MAX_ITER)

    def _gnb_install(self, ctx, conv=SURFACE_GREEN_CONVERGENCE, relFactor=SURFACE_RELAXATION_FACTOR):
        for i in range(self.num_contacts):
            self._add_contact(ctx, i, conv, relFactor)

    def _eval(self, E, i, which, conv, relFactor):
        ctx = default_context()
        ctx.set_system(self.F, self.S)
        ctx.sigma_clear()
        i = i % self.num_contacts
        self._add_contact(ctx, i, conv, relFactor)
        Es = np.atleast_1d(np.asarray(E, dtype=complex))
        n = len(self.indsList[i])
        out, iters, diffs = ctx.sigma_eval(0, which, Es, (n, n))
        for e, it, d in zip(Es, iters, diffs):
            self.last_iters[(complex(e), i)] = (int(it), float(d))
        return out

    # ---- surfG protocol -------------------------------------------------------------------
    def g(self, E, i, conv=SURFACE_GREEN_CONVERGENCE, relFactor=SURFACE_RELAXATION_FACTOR):
        out = self._eval(E, i, 1, conv, relFactor)
        return out[0] if np.ndim(E) == 0 else out

    def sigma(self, E, i, conv=SURFACE_GREEN_CONVERGENCE):
        blocks = self._eval(E, i, 0, conv, SURFACE_RELAXATION_FACTOR)
        inds = self.indsList[i]
        N = self.F.shape[0]
        full = np.zeros((blocks.shape[0], N, N), dtype=complex)
        full[np.ix_(np.arange(blocks.shape[0]), inds, inds)] += blocks
        return full[0] if np.ndim(E) == 0 else full

    def sigmaTot(self, E, conv=SURFACE_GREEN_CONVERGENCE):
        total = None
        for i in range(self.num_contacts):
            s = self.sigma(E, i, conv)
            total = s if total is None else total + s
        return total

    def setF(self, F, mu1=None, mu2=None):
        """Update the Fock matrix (and, for fully specified contacts, shift them to new chemical
        potentials).  The fully-specified branch follows the reference's INTENT (shift alpha by dmu*I
        and beta by dmu*Sbeta, surfG1D.py:335-342); the reference code itself fails there because it
        calls .at[] on Python lists."""
        self.F = np.array(F)
        if self.tauFromFock:
            taus, indsList = self.tauInds, self.indsList
            self.F[np.ix_(indsList[0], indsList[0])] = self.F[np.ix_(taus[0], taus[0])].copy()
            self.F[np.ix_(indsList[-1], indsList[-1])] = self.F[np.ix_(taus[1], taus[1])].copy()
            self.tauList = [np.array(self.F[np.ix_(taus[0], indsList[0])]), np.array(self.F[np.ix_(taus[1], indsList[-1])])]
            self.stauList = [np.array(self.S[np.ix_(taus[0], indsList[0])]), np.array(self.S[np.ix_(taus[1], indsList[-1])])]
            if self.contactFromFock:
                self.setContacts()
        if not self.contactFromFock:
            if self.fermiList[0] is None:
                self.fermiList[0] = mu1
                self.fermiList[-1] = mu2
            else:
                for i, mu in zip([0, -1], [mu1, mu2]):
                    f = self.fermiList[i]
                    if f is not None and mu is not None and f != mu:
                        d = mu - f
                        self.aList[i] = self.aList[i] + d * np.eye(len(self.aList[i]))
                        self.bList[i] = self.bList[i] + d * self.bSList[i]
                        self.fermiList[i] = mu
