"""Turn the reference's self-energy inputs (arrays, or duck-typed surfG objects: SURVEY.md §1) into
the device-side description the C ABI evaluates: constant contact blocks, 1-D chain / Bethe fixed
points, or — for arbitrary Python objects with sigmaTot()/sigma() — per-energy dense matrices that
the caller evaluates on the host and the GPU inverts (API compatibility, not a compute fallback)."""
import numpy as np

from .config import (SURFACE_GREEN_CONVERGENCE, SURFACE_RELAXATION_FACTOR, SURFACE_GREEN_MAX_ITER,
                     BETHE_MAX_ITER, BETHE_MIXING)

DESC, DENSE_CONST, DENSE_CALL = "desc", "dense_const", "dense_call"


def _is_bethe(g):
    from .surfGBethe import is_bethe_object
    return is_bethe_object(g)


def support(mat, tol=0.0):
    """orbitals on which a self-energy matrix (or its adjoint) is non-zero"""
    m = np.abs(np.asarray(mat)) > tol
    return np.nonzero(m.any(axis=0) | m.any(axis=1))[0]


def gamma_of(sig):
    return 1j * (sig - sig.conj().T)


class ArrayPlan:
    """energy-independent sig1 / sig2 arrays (vectors -> diagonal): transport.SigmaCalculator inputs"""

    def __init__(self, sigs, N):
        self.N = N
        self._src = [np.asarray(s) for s in sigs]
        self._full = None
        self.inds, self.blocks = [], []
        for s in self._src:
            if s.ndim == 1:                      # vector -> diagonal: never materialise the N x N matrix
                if s.shape != (N,):
                    raise ValueError(f"self-energy of shape {(s.shape[0], s.shape[0])} does not match a {N}x{N} system")
                i = np.nonzero(s)[0]
                self.inds.append(i)
                self.blocks.append(np.diag(s[i]).astype(complex))
            else:
                if s.shape != (N, N):
                    raise ValueError(f"self-energy of shape {s.shape} does not match a {N}x{N} system")
                i = support(s)
                self.inds.append(i)
                self.blocks.append(s[np.ix_(i, i)].astype(complex))
        # compact (low-rank) path when every contact touches at most half of the orbitals
        self.kind = DESC if all(0 < len(i) <= max(1, N // 2) for i in self.inds) else DENSE_CONST

    @property
    def full(self):
        if self._full is None:
            self._full = [np.diag(s).astype(complex) if s.ndim == 1 else s.astype(complex) for s in self._src]
        return self._full

    def install(self, ctx):
        ctx.sigma_clear()
        if self.kind == DESC:
            for b, i in zip(self.blocks, self.inds):
                ctx.sigma_add_const_block(i, b)

    def sigma_total(self, E=None):
        return sum(self.full)

    def sigma(self, E, i):
        return self.full[i]

    def ncontacts(self):
        return len(self.full)


class ObjectPlan:
    """surfG-protocol objects: g.sigmaTot(E), g.sigma(E, i)"""

    def __init__(self, g, N):
        self.g, self.N = g, N
        self.kind = DENSE_CALL
        self.full = None
        self.spin = getattr(g, "spin", 'r') if hasattr(g, "gList") else 'r'     # spin expansion done by the object itself
        if hasattr(g, "_gnb_install"):                      # our own surfG / surfGB / surfGBAt
            self.kind = DESC
        elif _is_bethe(g):
            self.kind = DESC                                 # a surfGB built by the reference's own constructor
        elif all(hasattr(g, a) for a in ("aList", "aSList", "bList", "bSList", "tauList", "stauList", "indsList", "eta")) \
                and all(np.shape(t)[0] == np.shape(t)[1] == len(i) for t, i in zip(g.tauList, g.indsList)):
            self.kind = DESC                                 # a reference-style surfG1D object
        elif isinstance(getattr(g, "sig", None), list) and all(np.shape(s) == (N, N) for s in g.sig):
            self.kind = DENSE_CONST                          # surfGTest-style constant matrices
            self.full = [np.asarray(s, dtype=complex) for s in g.sig]

    def install(self, ctx, spin_mode=None):
        ctx.sigma_clear()
        if self.kind != DESC:
            return
        g = self.g
        if _is_bethe(g):
            from .surfGBethe import install_bethe
            install_bethe(ctx, g, spin_mode=spin_mode)
            return
        if hasattr(g, "_gnb_install"):
            g._gnb_install(ctx)
            return
        for i in range(len(g.indsList)):
            ctx.sigma_add_chain1d(np.asarray(g.indsList[i]), g.aList[i], g.aSList[i], g.bList[i], g.bSList[i],
                                  g.tauList[i], g.stauList[i], g.eta, SURFACE_GREEN_CONVERGENCE,
                                  SURFACE_RELAXATION_FACTOR, SURFACE_GREEN_MAX_ITER)

    def sigma_total(self, E=None):
        if self.full is not None:
            return sum(self.full)
        return np.asarray(self.g.sigmaTot(E), dtype=complex)

    def sigma(self, E, i):
        if self.full is not None:
            return self.full[i]
        return np.asarray(self.g.sigma(E, i), dtype=complex)

    def ncontacts(self):
        for attr in ("indsList", "indsLists", "sig", "gList"):
            if hasattr(self.g, attr):
                return len(getattr(self.g, attr))
        return 2

    def sigma_total_batch(self, Elist):
        if self.full is None and getattr(self.g, "_gnb_batched", False):      # provider evaluates the batch in one call
            return np.asarray(self.g.sigmaTot(np.asarray(Elist)), dtype=complex)
        return np.stack([self.sigma_total(E) for E in Elist])

    def sigma_batch(self, Elist, i):
        return np.stack([self.sigma(E, i) for E in Elist])
