"""utils.inv drop-in on the GPU + the setup-time matrix power (reference gauNEGF/utils.py)."""
import numpy as np

from ._native import default_context


def inv(A):
    """A^-1 as the reference computes it, solve(A, I) (utils.py:52-54), on the B200.
    Accepts one (n, n) matrix or a batch (M, n, n)."""
    A = np.asarray(A)
    return default_context().inverse_batch(A)


def fractional_matrix_power(S, power):
    """S^p through a Hermitian eigendecomposition (utils.py:12-48).  Setup-time only (contact
    constructors); SURVEY.md §8(f) N4 keeps it on the host."""
    w, v = np.linalg.eigh(np.asarray(S))
    w = np.maximum(w, 1e-16)
    return v @ np.diag(np.power(w, power)) @ v.conj().T


def eig(A):
    """eigenvalues / right eigenvectors of a general matrix (utils.py:56-58).  Setup-time host linear algebra
    (Fermi-level guesses of density.getFermiContact); nothing on the energy grid calls it."""
    return np.linalg.eig(np.asarray(A))


def eigh(A):
    """eigh as the reference calls it (utils.py:60-62): lower triangle of A, ascending eigenvalues."""
    return np.linalg.eigh(np.asarray(A))
