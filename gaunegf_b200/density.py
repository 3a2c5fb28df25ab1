"""Density-matrix quadrature drivers — drop-in for the energy-dependent part of gauNEGF/density.py
(:64-119, 211-273, 385-834).  Nodes and weights are built on the host exactly as the reference
builds them; every batch of energies is evaluated on the B200 through integrate.GrInt/GrLessInt.
"""
import numpy as np
from scipy.special import roots_legendre

from .config import (TEMPERATURE, ADAPTIVE_INTEGRATION_TOL, FERMI_CALCULATION_TOL, N_KT, MAX_CYCLES,
                     MAX_GRID_POINTS)
from .integrate import GrInt, GrLessInt
from ._native import default_context
from .sigma_plan import ObjectPlan, DESC, DENSE_CONST

# CONSTANTS (density.py:59-61)
har_to_eV = 27.211386   # eV/Hartree
kB = 8.617e-5           # eV/Kelvin


def fermi(E, mu, T):
    """Fermi-Dirac occupation (density.py:64-86); T = 0 is the step (E <= mu), with numpy's
    lexicographic ordering for complex E."""
    kT = kB * T
    if kT == 0:
        return (E <= mu) * 1
    return 1 / (np.exp((E - mu) / kT) + 1)


def getANTPoints(N):
    """ANT.Gaussian's nested, modified Gauss-Chebyshev nodes and weights on (-1, 1) (density.py:88-119).
    Always an even number of points; the N-point set contains the N/3-point set."""
    k = np.arange(1, N + 1, 2)
    theta = k * np.pi / (2 * N)
    xs, xcc = np.sin(theta), np.cos(theta)
    x = 1.0 + 0.21220659078919378103 * xs * xcc * (3 + 2 * xs * xs) - k / N
    w = xs ** 4 * 16.0 / (3 * N)
    return np.concatenate((x, -1 * x)), np.concatenate((w, w))


def integratePointsAdaptiveANT(computePoint, tol=ADAPTIVE_INTEGRATION_TOL, maxN=MAX_GRID_POINTS, debug=False):
    """Nested refinement N = 2, 6, 18, ... (density.py:211-273): each level rescales the previous
    integral by the weight ratio of the re-used nodes and adds ONLY the new nodes, which go to the
    GPU as one batch (2, 4, 12, 36, 108, 324 energies)."""
    prev_x = prev_sumW = P = new_P = None
    N = 2
    maxDP = 1e10
    while N <= maxN:
        x, w = getANTPoints(N)
        if prev_x is None:
            P = computePoint(x[0:2], w[0:2])
        else:
            old_mask = np.isin(np.round(x, 14), np.round(prev_x, 14))
            assert int(old_mask.sum()) == prev_x.size, "Old nodes mismatch"
            ratio = float(np.sum(w[old_mask]) / prev_sumW)
            new_mask = ~old_mask
            new_P = P * ratio
            new_P += computePoint(x[new_mask], w[new_mask])
            maxDP = np.max(np.abs(new_P - P))
            if debug:
                P_debug = computePoint(x, w)
                print(f"N={N}, nested-weight ratio ~ {ratio:.3f}, maxDP={maxDP:.3e}")
                print(f"Direct Calculation: N={N}, maxDP={np.max(np.abs(P_debug - P)):.3e}, "
                      f"maxDiff={np.max(np.abs(P_debug - new_P)):.3e}")
            P = new_P.copy()
            if maxDP < tol:
                print(f'Adaptive integration converged to {maxDP:.3e} in {N} points.')
                return new_P
        prev_x = x
        prev_sumW = float(np.sum(w))
        N *= 3
    N /= 3
    print(f'Adaptive integration reached full grid ({N} points), final error {maxDP:.3e}')
    return new_P


# ---- equilibrium, real axis (density.py:385-484) -----------------------------------------------
def densityRealN(F, S, g, Emin, mu, N=100, T=TEMPERATURE, showText=True):
    Emax = mu + N_KT * kB * T
    mid = (Emax - Emin) / 2
    x, w = roots_legendre(N)
    x = np.real(x)
    Elist = mid * (x + 1) + Emin
    weights = mid * w * fermi(Elist, mu, T)
    if showText:
        print(f'Integrating {N} points along real axis...')
    defInt = GrInt(F, S, g, Elist, weights)
    if showText:
        print('Integration done!')
    return (-1 + 0j) * np.imag(defInt) / (np.pi)


def densityReal(F, S, g, Emin, mu, tol=ADAPTIVE_INTEGRATION_TOL, T=TEMPERATURE, maxN=MAX_CYCLES, debug=False):
    P = np.zeros_like(F)
    N = 1
    maxDP = 1e9
    while N < maxN:
        P_ = P.copy()
        P = densityRealN(F, S, g, Emin, mu, N, T, showText=False)
        maxDP = np.max(np.abs(P - P_))
        if maxDP < tol:
            print(f'Adaptive integration converged to {maxDP:.3e} in {N} points.')
            return P
        N *= 2
    print(f'Warning: adaptive integration not converged after {maxN} points: maxDP={maxDP:.2E}')
    return P


# ---- non-equilibrium window (density.py:487-658) ------------------------------------------------
def _window(mu1, mu2, T):
    muLo, muHi = min(mu1, mu2), max(mu1, mu2)
    dInt = np.sign(mu2 - mu1)
    Emax = muHi + N_KT * kB * T
    Emin = muLo - N_KT * kB * T
    return muLo, muHi, dInt, Emin, (Emax - Emin) / 2


def densityGridN(F, S, g, mu1, mu2, ind=None, N=100, T=TEMPERATURE, showText=True):
    muLo, muHi, dInt, Emin, mid = _window(mu1, mu2, T)
    x, w = roots_legendre(N)
    x = np.real(x)
    energies = mid * (x + 1) + Emin
    dfermi = fermi(energies, muHi, T) - fermi(energies, muLo, T)
    weights = mid * w * dfermi * dInt
    if showText:
        print(f'Real integration over {N} points...')
    den = GrLessInt(F, S, g, energies, weights, ind)
    if showText:
        print('Integration done!')
    return den / (2 * np.pi)


def densityGridTrap(F, S, g, mu1, mu2, ind=None, N=100, T=TEMPERATURE):
    """Midpoint ("trapezoid") variant (density.py:547-603): the reference loops serially; the same
    N-1 midpoints and weights go to the GPU as one batch."""
    muLo, muHi, dInt, Emin, mid = _window(mu1, mu2, T)
    Egrid = np.linspace(Emin, Emin + 2 * mid, N)
    print(f'Real integration over {N} points...')
    E = (Egrid[1:] + Egrid[:-1]) / 2
    dE = Egrid[1:] - Egrid[:-1]
    weights = (fermi(E, muHi, T) - fermi(E, muLo, T)) * dE * dInt
    den = GrLessInt(F, S, g, E, weights, ind) if N > 1 else np.zeros(np.shape(F), dtype=complex)
    print('Integration done!')
    return den / (2 * np.pi)


def densityGrid(F, S, g, mu1, mu2, ind=None, tol=ADAPTIVE_INTEGRATION_TOL, T=TEMPERATURE, debug=False):
    muLo, muHi, dInt, Emin, mid = _window(mu1, mu2, T)

    def computePoint(x, w):
        E = mid * (x + 1) + Emin
        dFermi = fermi(E, muHi, T) - fermi(E, muLo, T)
        return GrLessInt(F, S, g, E, mid * w * dFermi * dInt, ind)

    den = integratePointsAdaptiveANT(computePoint, tol=tol, debug=debug)
    if debug:
        print('Integration done!')
    return den / (2 * np.pi)


# ---- equilibrium, complex contour (density.py:660-816) ------------------------------------------
def _semicircle(Emin, mu, T):
    broadening = 10 * kB * T
    Emax = mu - broadening
    return (Emin + Emax) / 2, (Emax - Emin) / 2, broadening


def densityComplexN(F, S, g, Emin, mu, N=100, T=TEMPERATURE, showText=True, method='ant'):
    center, r, broadening = _semicircle(Emin, mu, T)
    if method == 'legendre':
        x, w = roots_legendre(N)
    elif method == 'chebyshev':
        k = np.arange(1, N + 1)
        x = np.cos(k * np.pi / (N + 1))
        w = (np.pi / (N + 1)) * (np.sin(k * np.pi / (N + 1)) ** 2) / np.sqrt(1 - (x ** 2))
    elif method == 'ant':
        x, w = getANTPoints(N)
    else:   # midpoint rule, both end points included (density.py:714-716)
        x = np.linspace(-1, 1, N)
        w = 2 * np.ones(N) / N
    theta = np.pi / 2 * (x + 1)
    Elist = center + r * np.exp(1j * theta)
    dz = 1j * r * np.exp(1j * theta)
    weights = (np.pi / 2) * w * fermi(Elist, mu, T) * dz
    if showText:
        print(f'Complex Integration over {N} points...')
    lineInt = GrInt(F, S, g, Elist, weights)
    if T > 0:
        if showText:
            print('Integrating Fermi Broadening')
        Nbroad = int(N // 8)
        if method in ('legendre', 'chebyshev', 'ant'):
            x_fermi, w_fermi = roots_legendre(Nbroad)
        else:
            x_fermi = np.linspace(-1, 1, Nbroad)
            w_fermi = 2 * np.ones(Nbroad) / Nbroad
        Elist = broadening * (x_fermi) + mu
        weights = broadening * w_fermi * fermi(Elist, mu, T)
        lineInt += GrInt(F, S, g, Elist, weights)
    if showText:
        print('Integration done!')
    return (1 + 0j) * np.imag(lineInt) / np.pi


def densityComplex(F, S, g, Emin, mu, tol=ADAPTIVE_INTEGRATION_TOL, T=TEMPERATURE, debug=False):
    center, r, broadening = _semicircle(Emin, mu, T)

    def computePoint(x, w):
        theta = np.pi / 2 * (x + 1)
        z = center + r * np.exp(1j * theta)
        dz = 1j * r * np.exp(1j * theta)
        return GrInt(F, S, g, z, (np.pi / 2) * w * dz * fermi(z, mu, T))

    print('Complex Contour Integration:')
    lineInt = integratePointsAdaptiveANT(computePoint, tol=tol, debug=debug)
    if T > 0:
        print('Integrating Fermi Broadening:')

        def computePointBroadening(x, w):
            E = broadening * (x) + mu
            return GrInt(F, S, g, E, broadening * w * fermi(E, mu, T))

        lineInt += integratePointsAdaptiveANT(computePointBroadening, tol=tol, debug=debug)
    return (1 + 0j) * np.imag(lineInt) / np.pi


# ---- integration limit (density.py:49-54, 821-834) ------------------------------------------------
def _compute_dos_at_energy(E, F, S, sigma_total):
    """-Im Tr G / pi at one energy with a given total self-energy (density.py:49-54), on the GPU."""
    ctx = default_context()
    ctx.set_system(F, S)
    ctx.sigma_clear()
    tot, _ = ctx.dos_dense(np.array([E]), np.asarray(sigma_total, dtype=complex), per_site=False)
    return float(tot[0])


def calcEmin(F, S, g, tol=FERMI_CALCULATION_TOL, maxN=MAX_CYCLES):
    """Lower integration bound from the DOS tail (density.py:821-834).  The generalised eigenvalue
    estimate is setup-time host linear algebra; the DOS samples run on the GPU."""
    # eigh of inv(S) @ F, exactly as the reference does (it reads the lower triangle only)
    D = np.linalg.eigvalsh(np.linalg.solve(np.asarray(S), np.eye(len(S))) @ np.asarray(F))
    Emin = min(D.real.flatten()) - 5
    counter = 0
    dP = _compute_dos_at_energy(Emin, F, S, g.sigmaTot(Emin))
    while dP > tol and counter < maxN:
        Emin -= 1
        dP = _compute_dos_at_energy(Emin, F, S, g.sigmaTot(Emin))
        counter += 1
    if counter == maxN:
        print(f'Warning: Emin still not within tolerance (final value = {dP}) after {maxN} energy samples')
    print(f'Calculated Emin: {Emin} eV, DOS = {dP:.2E}')
    return Emin
