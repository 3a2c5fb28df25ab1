"""formSigma — the one matTools helper on the hot path (gauNEGF/matTools.py:39-74).  The Gaussian
matrix I/O of that module needs gauopen and is out of scope."""
import numpy as np


def formSigma(inds, V, nsto, S=0):
    """nsto x nsto self-energy: background -i*1e-9*S (identity if S is not given), with V on the
    listed orbitals (scalar -> diagonal entries, matrix -> the inds x inds block)."""
    if isinstance(S, int):
        S = np.eye(nsto)
    sigma = np.array(-1j * 1e-9 * S, dtype=complex)
    if isinstance(V, (int, complex, float)):
        for i in inds:
            sigma[i, i] = V
    else:
        sigma[np.ix_(inds, inds)] = V
    return sigma
