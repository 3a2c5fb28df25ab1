"""Seeded synthetic Hermitian F/S pairs and contact layouts (Gaussian/gauopen are unavailable).

These are the inputs of the five BASELINE.json configs, parameterised by size so
that parity tests can run them small and bench.py at full size.  Pure numpy; used by tests/,
bench.py, __graft_entry__.smoke() and tests/golden/make_golden.py.
"""
import numpy as np


def chain(N=64, t=-1.0, eps=0.0, gamma=0.1):
    """cfg 1: tight-binding chain, S = I, wide-band contacts -i*gamma on the end orbitals
    (same system as the reference's tests/test_transport_checkpointing.py:22-58 builder)."""
    F = np.zeros((N, N))
    i = np.arange(N - 1)
    F[i, i + 1] = t
    F[i + 1, i] = t
    F[np.arange(N), np.arange(N)] = eps
    S = np.eye(N)
    sig1 = np.zeros(N, dtype=complex)
    sig2 = np.zeros(N, dtype=complex)
    sig1[0] = -1j * gamma
    sig2[-1] = -1j * gamma
    return F, S, sig1, sig2


def hermitian_pair(N, seed=0, complex_F=False):
    """cfg 2/3/5 device: random symmetric F (scale 0.5) and SPD overlap S = I + 2B/sqrt(N)."""
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((N, N))
    F = (A + A.T) / 2 * 0.5
    B = rng.random((N, N)) * 0.2
    B = (B + B.T) / 2
    np.fill_diagonal(B, 0.0)
    S = np.eye(N) + 2 * B / np.sqrt(N)
    if complex_F:
        C = rng.standard_normal((N, N)) * 0.1
        F = F + 1j * (C - C.T) / 2
    return F, S


def end_contacts(N, nc):
    return [np.arange(nc), np.arange(N - nc, N)]


def block_sigma_vectors(N, nc, gamma=0.1):
    """cfg 3: constant -i*gamma on the first / last nc orbitals, as length-N vectors."""
    s1 = np.zeros(N, dtype=complex)
    s2 = np.zeros(N, dtype=complex)
    s1[:nc] = -1j * gamma
    s2[N - nc:] = -1j * gamma
    return s1, s2


def lead_device_lead(n_lead=128, n_dev=512, seed=2, s_off=0.0):
    """cfg 4: extended block-tridiagonal system [lead | device | lead] for surfG1D.

    Returns F, S (size n_dev + 2 n_lead), indsList (lead orbitals) and taus (the device-side
    orbitals each lead couples to) in the layout of the reference's
    tests/test_transport_checkpointing.py:174-178 example.
    """
    rng = np.random.default_rng(seed)
    nb = n_lead
    assert n_dev % nb == 0
    nblk = n_dev // nb + 2
    N = nblk * nb
    F = np.zeros((N, N))
    S = np.eye(N)
    H0 = rng.standard_normal((nb, nb))
    H0 = (H0 + H0.T) / 2 * 0.5
    V0 = rng.standard_normal((nb, nb)) * 0.2 / np.sqrt(nb) * 4
    for b in range(nblk):
        sl = slice(b * nb, (b + 1) * nb)
        Hb = H0 if b in (0, 1, nblk - 2, nblk - 1) else None
        if Hb is None:
            Hb = rng.standard_normal((nb, nb))
            Hb = (Hb + Hb.T) / 2 * 0.5
        F[sl, sl] = Hb
        if b + 1 < nblk:
            sr = slice((b + 1) * nb, (b + 2) * nb)
            F[sl, sr] = V0
            F[sr, sl] = V0.T
            if s_off:
                S[sl, sr] = s_off * np.eye(nb)
                S[sr, sl] = s_off * np.eye(nb)
    inds = [np.arange(0, nb), np.arange(N - nb, N)]
    taus = [np.arange(nb, 2 * nb), np.arange(N - 2 * nb, N - nb)]
    return F, S, inds, taus


def contour_points(n, Emin=-30.0, mu=0.0):
    """ANT-style nested Gauss-Chebyshev nodes on the upper semicircle through (Emin, mu): the
    energy/weight lists densityComplexN hands to GrInt at T = 0 (reference density.py:699-722)."""
    k = np.arange(1, n + 1, 2)
    th = k * np.pi / (2 * n)
    s, c = np.sin(th), np.cos(th)
    x = 1.0 + 0.21220659078919378103 * s * c * (3 + 2 * s * s) - k / n
    x = np.concatenate((x, -x))
    w = np.concatenate((s ** 4, s ** 4)) * 16.0 / (3 * n)
    center, r = (Emin + mu) / 2, (mu - Emin) / 2
    theta = np.pi / 2 * (x + 1)
    z = center + r * np.exp(1j * theta)
    dz = 1j * r * np.exp(1j * theta)
    occ = (z <= mu) * 1
    return z, (np.pi / 2) * w * occ * dz


def analytic_density_case(N, seed=5, nc=4):
    """Inputs of the energy-independent analytic density (density.py:276-329): eigenvectors V, Vc = inv(V)^H and
    eigenvalues D of X^H (F + Sigma) X with constant contact self-energies, and Gamma in the same basis."""
    F, S = hermitian_pair(N, seed=seed)
    X = np.linalg.inv(np.linalg.cholesky(S)).conj().T
    sig = np.zeros((N, N), dtype=complex)
    sig[:nc, :nc] = -0.1j * np.eye(nc)
    sig[-nc:, -nc:] = -0.2j * np.eye(nc)
    D, V = np.linalg.eig(X.conj().T @ (F + sig) @ X)
    return V, np.linalg.inv(V).conj().T, D, X.conj().T @ (1j * (sig - sig.conj().T)) @ X
