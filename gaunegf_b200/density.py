"""Density-matrix quadrature drivers — drop-in for the energy-dependent part of gauNEGF/density.py
(:64-119, 211-273, 385-834).  Nodes and weights are built on the host exactly as the reference
builds them; every batch of energies is evaluated on the B200 through integrate.GrInt/GrLessInt.
"""
import numpy as np
from scipy.special import roots_legendre

from .config import (TEMPERATURE, ADAPTIVE_INTEGRATION_TOL, FERMI_CALCULATION_TOL, FERMI_SEARCH_CYCLES, N_KT,
                     ENERGY_MIN, MAX_CYCLES, MAX_GRID_POINTS)
from .integrate import GrInt, GrLessInt, GrIntLevels
from ._native import default_context
from .utils import inv
from .sigma_plan import ObjectPlan, DESC, DENSE_CONST

# CONSTANTS (density.py:59-61)
har_to_eV = 27.211386   # eV/Hartree
kB = 8.617e-5           # eV/Kelvin
FERMI_DEBUG = False     # density.py:57


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


# Nested levels evaluated per GPU batch when the integrand offers `computePoint.levels` (speculation: the nodes of
# level 3N are known before level N's convergence test).  Groups of nested grid sizes: a group costs one launch
# chain; levels computed past the converged one are discarded.
_SPECULATIVE_GROUPS = ((2, 6, 18), (54, 162), (486,), (1458,), (4374,))


def integratePointsAdaptiveANT(computePoint, tol=ADAPTIVE_INTEGRATION_TOL, maxN=MAX_GRID_POINTS, debug=False):
    """Nested refinement N = 2, 6, 18, ... (density.py:211-273): each level rescales the previous
    integral by the weight ratio of the re-used nodes and adds ONLY the new nodes, which go to the
    GPU as one batch (2, 4, 12, 36, 108, 324 energies).

    If `computePoint` has an attribute `levels` (a function of [(x, w), ...] returning the list of integrals), the new
    nodes of SEVERAL consecutive levels are evaluated in one GPU batch (2 + 4 + 12, then 36 + 108, ...): the host
    arithmetic (P * ratio + new, the convergence test and the prints) is the reference's, level by level, on sums
    that are the same as those of separate calls."""
    many = None if debug else getattr(computePoint, "levels", None)
    ahead = {}                       # N -> integral over the NEW nodes of level N, computed speculatively

    def new_nodes_sum(N, x_new, w_new, later):
        if many is None:
            return computePoint(x_new, w_new)
        if N not in ahead:
            group = next((g for g in _SPECULATIVE_GROUPS if N in g), (N,))
            todo = [(N, x_new, w_new)] + [t for t in later(group) if t[0] > N]
            for (n, _, _), val in zip(todo, many([(xx, ww) for _, xx, ww in todo])):
                ahead[n] = val
        return ahead.pop(N)

    def later_levels(group):
        """(N, new x, new w) of the levels of `group` above the current one, within maxN"""
        out = []
        for n in group:
            if n > maxN or n < 6:
                continue
            xa, wa = getANTPoints(n)
            xp, _ = getANTPoints(n // 3)
            mask = ~np.isin(np.round(xa, 14), np.round(xp, 14))
            out.append((n, xa[mask], wa[mask]))
        return out

    prev_x = prev_sumW = P = new_P = None
    N = 2
    maxDP = 1e10
    while N <= maxN:
        x, w = getANTPoints(N)
        if prev_x is None:
            P = new_nodes_sum(N, x[0:2], w[0:2], later_levels)
        else:
            old_mask = np.isin(np.round(x, 14), np.round(prev_x, 14))
            assert int(old_mask.sum()) == prev_x.size, "Old nodes mismatch"
            ratio = float(np.sum(w[old_mask]) / prev_sumW)
            new_mask = ~old_mask
            new_P = P * ratio
            new_P += new_nodes_sum(N, x[new_mask], w[new_mask], later_levels)
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

    def contour(x, w):
        theta = np.pi / 2 * (x + 1)
        z = center + r * np.exp(1j * theta)
        dz = 1j * r * np.exp(1j * theta)
        return z, (np.pi / 2) * w * dz * fermi(z, mu, T)

    def computePoint(x, w):
        return GrInt(F, S, g, *contour(x, w))

    computePoint.levels = lambda pairs: GrIntLevels(F, S, g, [contour(x, w) for x, w in pairs])
    print('Complex Contour Integration:')
    lineInt = integratePointsAdaptiveANT(computePoint, tol=tol, debug=debug)
    if T > 0:
        print('Integrating Fermi Broadening:')

        def tail(x, w):
            E = broadening * (x) + mu
            return E, broadening * w * fermi(E, mu, T)

        def computePointBroadening(x, w):
            return GrInt(F, S, g, *tail(x, w))

        computePointBroadening.levels = lambda pairs: GrIntLevels(F, S, g, [tail(x, w) for x, w in pairs])
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
    D = np.linalg.eigvalsh(inv(np.asarray(S)) @ np.asarray(F))          # S^-1 on the GPU (utils.inv), eigvalsh on the host
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


# ---- grid fits and Fermi-level searches (density.py:836-1515) ------------------
# Host-side root finding over the GPU drivers above: every probe of the electron count is one
# contour integral (densityComplexN -> integrate.GrInt -> one batched launch sequence on the B200).
def _diag_change(new, old):
    return max(abs(np.diag(new - old)))


def _doubling_fit(evaluate, start, tol, cap, report):
    """Double the point count until the largest change of a diagonal density element is <= tol
    (the loop shared by integralFit / integralFitNEGF, density.py:878-913, 949-964).  Returns the
    last count that was still needed (the converged count halved), as the reference does."""
    n, change, prev = start, np.inf, None
    while change > tol and n < cap:
        n *= 2
        cur = evaluate(n)
        change = _diag_change(cur, 0 * cur if prev is None else prev)
        report(change, cur)
        prev = cur
    return n, change


def integralFit(F, S, g, mu, Eminf=ENERGY_MIN, tol=FERMI_CALCULATION_TOL, T=TEMPERATURE, maxN=MAX_CYCLES):
    """(Emin, N1, N2): contour start from the DOS tail, then the number of complex-contour and of real-axis
    points needed for `tol` on the density diagonal (density.py:836-914)."""
    Emin = calcEmin(F, S, g, tol, maxN)

    Ncomplex, dP = _doubling_fit(lambda n: np.real(densityComplexN(F, S, g, Emin, mu, n, T=T)), 4, tol, maxN,
                                 lambda d, rho: print(f"MaxDP = {d:.2E}, N = {sum(np.diag(rho).real):2f}"))
    if dP < tol:
        Ncomplex /= 2
    elif Ncomplex >= maxN and dP > tol:
        print(f'Warning: Ncomplex still not within tolerance (final value = {dP})')
    print(f'Final Ncomplex: {Ncomplex}')

    Nreal, dP = _doubling_fit(lambda n: np.real(densityRealN(F, S, g, Eminf, Emin, n, T=0)), 8, tol, maxN,
                              lambda d, rho: print(f"MaxDP = {d:.2E}"))
    if dP < tol:
        Nreal /= 2
    elif Nreal >= maxN and dP > tol:
        print(f'Warning: Nreal still not within tolerance (final value = {dP})')
    print(f'Final Nreal: {Nreal}')
    return Emin, Ncomplex, Nreal


def integralFitNEGF(F, S, g, fermi, qV, Eminf=ENERGY_MIN, tol=FERMI_CALCULATION_TOL, T=TEMPERATURE,
                    maxGrid=MAX_GRID_POINTS):
    """Number of real-axis points for the non-equilibrium window integrals (density.py:916-964)."""
    def both_windows(n):
        rho = np.real(densityGridN(F, S, g, fermi, fermi + (qV / 2), ind=0, N=n, T=T))
        return rho + np.real(densityGridN(F, S, g, fermi, fermi - (qV / 2), ind=-1, N=n, T=T))

    N, dP = _doubling_fit(both_windows, 8, tol, maxGrid, lambda d, rho: print(f"MaxDP = {d:.2E}"))
    if dP < tol:
        N /= 2
    elif N >= maxGrid and dP > tol:
        print(f'Warning: N still not within tolerance (final value = {dP})')
    print(f'Final Nnegf: {N}')
    return N


def _electrons(P, S, nOrbs=0):
    PS = P @ S
    return np.trace(PS if nOrbs == 0 else PS[-nOrbs:, -nOrbs:])


def _mid_gap(F, S, ne, per_cell=1, hermitian=False):
    """sorted real eigenvalues of inv(S) F and the middle of the gap above orbital per_cell*ne"""
    M = inv(np.asarray(S)) @ np.asarray(F)                              # S^-1 on the GPU (utils.inv)
    orbs = np.sort(np.real(np.linalg.eigvalsh(M) if hermitian else np.linalg.eigvals(M)))
    k = per_cell * int(ne)
    return orbs, (orbs[k - 1] + orbs[k]) / 2


def getFermiContact(g, ne, tol=FERMI_CALCULATION_TOL, Eminf=ENERGY_MIN, maxcycles=MAX_CYCLES, T=TEMPERATURE,
                    nOrbs=0):
    """Fermi energy of a contact treated as its own system (density.py:967-1003)."""
    S, F = g.S, g.F
    orbs, fermi_guess = _mid_gap(F, S, ne)
    Emin, N1, N2 = integralFit(F, S, g, fermi_guess, Eminf, tol, T, maxN=maxcycles)
    return calcFermi(g, ne, Emin, max(orbs), fermi_guess, N1, N2, Eminf, T, tol, maxcycles, nOrbs)[0]


def getFermi1DContact(gSys, ne, ind=0, tol=FERMI_CALCULATION_TOL, Eminf=ENERGY_MIN, T=TEMPERATURE,
                      maxcycles=MAX_CYCLES):
    """(fermi, Emin, N1, N2) of the periodic chain behind contact `ind` of a surfG1D system
    (density.py:1005-1053): integration parameters from a two-cell model, search on the one-cell contact."""
    from .surfG1D import surfG
    F, S = gSys.aList[ind], gSys.aSList[ind]
    tau, stau = gSys.bList[ind], gSys.bSList[ind]
    inds = np.arange(len(F))
    g = surfG(F, S, [inds], [tau], [stau], eta=1e-6)
    F2 = np.block([[F, tau], [tau.conj().T, F]])
    S2 = np.block([[S, stau], [stau.T, S]])
    g2 = surfG(F2, S2, [inds], [tau], [stau], eta=1e-6)
    orbs, fermi_guess = _mid_gap(F2, S2, ne, per_cell=2, hermitian=True)
    Emin, N1, N2 = integralFit(F2, S2, g2, fermi_guess, Eminf, tol, T, maxN=maxcycles)
    return calcFermi(g, ne, Emin, max(orbs), fermi_guess, N1, N2, Eminf, T, tol, maxcycles)


def calcFermi(g, ne, Emin, Emax, fermiGuess=0, N1=100, N2=50, Eminf=ENERGY_MIN, T=TEMPERATURE,
              tol=FERMI_CALCULATION_TOL, maxcycles=MAX_CYCLES, nOrbs=0):
    """Bisection on the electron count of the contact's own system between Emin and Emax
    (density.py:1056-1143).  With N1/N2 = None the reference passes keywords its adaptive drivers do not
    accept and raises; here those branches run the adaptive drivers."""
    dos_eminf = _compute_dos_at_energy(Eminf, g.F, g.S, g.sigmaTot(Eminf))
    print(f'Eminf DOS = {dos_eminf}')

    def below(Tlow):
        if N2 is None:
            return _quiet(densityReal, g.F, g.S, g, Eminf, Emin, tol, Tlow)
        return densityRealN(g.F, g.S, g, Eminf, Emin, N2, Tlow, showText=False)

    def contour(mu):
        if N1 is None:
            return _quiet(densityComplex, g.F, g.S, g, Emin, mu, tol, T)
        return densityComplexN(g.F, g.S, g, Emin, mu, N1, T, showText=False, method='legendre')

    fermi = fermiGuess
    nELow = _electrons(below(T), g.S, nOrbs)
    print(f'Electrons below lowest onsite energy: {nELow}')
    if nELow >= ne:
        raise Exception('Calculated Fermi energy is below lowest orbital energy!')
    Ncurr, counter, lBound, uBound = -1, 0, Emin, Emax
    print('Calculating Fermi energy using bisection:')
    while abs(ne - Ncurr) > tol and uBound - lBound > tol / 10 and counter < maxcycles:
        g.setF(g.F, fermi, fermi)
        Ncurr = _electrons(np.real(below(0) + contour(fermi)), g.S, nOrbs)
        dN = ne - Ncurr
        if dN > 0 and fermi > lBound:
            lBound = fermi
        elif dN < 0 and fermi < uBound:
            uBound = fermi
        if abs(ne - Ncurr) > tol:
            fermi = (uBound + lBound) / 2
        print("DN:", dN, "Fermi:", fermi, "Bounds:", lBound, uBound)
        counter += 1
    if abs(ne - Ncurr) > tol and counter > maxcycles:
        print(f'Warning: Fermi energy still not within tolerance! Ef = {fermi:.2f} eV, N = {Ncurr:.2f})')
    print(f'Finished after {counter} iterations, Ef = {fermi:.2f}')
    return fermi, Emin, N1, N2


def _quiet(f, *a, **k):
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        return f(*a, **k)


def _contour_density(g, Emin, N, tol, T):
    """P(mu) used by the device-level searches: fixed N-point contour, or the adaptive one for N = None"""
    assert_n = len(g.F)
    if N is None:
        return assert_n, (lambda mu: densityComplex(g.F, g.S, g, Emin, mu, tol, T))
    return assert_n, (lambda mu: densityComplexN(g.F, g.S, g, Emin, mu, N, T))


def _probe(g, pMu, mu):
    """move the contacts to mu, integrate, count electrons"""
    g.setF(g.F, mu, mu)
    P = pMu(mu)
    return P, np.trace(P @ g.S).real


def _track(bounds, E, n):
    """bracket bookkeeping of the Muller / PolyFit searches: n = N(E) - ne"""
    u, l = bounds
    if n > 0:
        u = E if u is None else min(u, E)
    elif n < 0:
        l = E if l is None else max(l, E)
    return u, l


def calcFermiBisect(g, ne, Emin, Ef, N, tol=ADAPTIVE_INTEGRATION_TOL, conv=FERMI_CALCULATION_TOL,
                    maxcycles=FERMI_SEARCH_CYCLES, T=TEMPERATURE, uBound=None, lBound=None):
    """Bracket-then-bisect search (density.py:1145-1201) -> (Ef, dE, P).  The bracket expansion sizes its steps
    from a DOS sample that the reference takes with F and S exchanged (density.py:1176, SURVEY.md appendix
    A.5); that call is reproduced as it is, because it decides which energies are visited."""
    nbf, pMu = _contour_density(g, Emin, N, tol, T)
    assert ne < nbf, "Number of electrons cannot exceed number of basis functions!"
    E = Ef + 0.0
    dE = tol
    counter = 0
    P, Ncurr = _probe(g, pMu, E)
    while None in [uBound, lBound] and counter < maxcycles:
        if Ncurr > ne:
            uBound = E + 0.0
            Ef = uBound
            E -= dE
        if Ncurr < ne:
            lBound = E + 0.0
            Ef = lBound
            E += dE
        if FERMI_DEBUG:
            print(f"DEBUG: Ef={Ef:.2f}, dN={ne-Ncurr:.2E}, dE={dE:.2E}")
        dos = _compute_dos_at_energy(E, g.S, g.F, g.sigmaTot(E))
        dE = max(2 * abs(Ncurr - ne) / dos, dE)
        counter += 1
        P, Ncurr = _probe(g, pMu, E)
    while abs(ne - Ncurr) > conv and counter < maxcycles and uBound != lBound:
        dN = ne - Ncurr
        if dN > 0 and Ef > lBound:
            lBound = Ef + 0.0
        elif dN < 0 and Ef < uBound:
            uBound = Ef + 0.0
        Ef = (uBound + lBound) / 2
        dE = uBound - lBound
        if FERMI_DEBUG:
            print(f"DEBUG: Ef={Ef:.2f}, dN={dN:.2E}, dE={dE:.2E}")
        counter += 1
        if abs(dN) > conv:
            g.setF(g.F, Ef, Ef)
            P = pMu(Ef)
            Ncurr = np.trace(P @ g.S)
    if counter == maxcycles:
        print(f'Warning: Max cycles reached, convergence = {abs(Ncurr-ne):.2E}')
    elif uBound == lBound:
        print(f'Warning: Bisection failed, convergence = {abs(Ncurr-ne):.2E}')
    return Ef, dE, P


def calcFermiSecant(g, ne, Emin, Ef, N, tol=ADAPTIVE_INTEGRATION_TOL, conv=FERMI_CALCULATION_TOL,
                    maxcycles=FERMI_SEARCH_CYCLES, T=TEMPERATURE):
    """Secant search on N(mu) - ne (density.py:1203-1238) -> (Ef, dE, P, |N - ne|)."""
    nbf, pMu = _contour_density(g, Emin, N, tol, T)
    assert ne < nbf, "Number of electrons cannot exceed number of basis functions!"
    P, nCurr = _probe(g, pMu, Ef)
    dE = conv
    counter = 0
    while abs(nCurr - ne) > conv and counter < maxcycles:
        Ef += dE
        P, nNext = _probe(g, pMu, Ef)
        if FERMI_DEBUG:
            print(f"DEBUG: Ef={Ef:.2f}, dN={nNext-ne:.2E}, dE={dE:.2E}")
        counter += 1
        if abs(nNext - nCurr) < 1e-10:
            print('Warning: change in ne low, reducing step size')
            dE *= 0.1
            continue
        dE = dE * ((ne - nCurr) / (nNext - nCurr)) - dE
        nCurr = nNext + 0.0
    Ef += dE
    if counter == maxcycles:
        print(f'Warning: Max cycles reached, convergence = {abs(nCurr-ne):.2E}')
    return Ef, dE, P, abs(nCurr - ne)


def calcFermiMuller(g, ne, Emin, Ef, N, tol=ADAPTIVE_INTEGRATION_TOL, conv=FERMI_CALCULATION_TOL,
                    maxcycles=FERMI_SEARCH_CYCLES, T=TEMPERATURE):
    """Muller's method (parabola through the last three probes) on N(mu) - ne (density.py:1240-1331)
    -> (Ef, dE, P, |N - ne|, uBound, lBound)."""
    nbf, pMu = _contour_density(g, Emin, N, tol, T)
    assert ne < nbf, "Number of electrons cannot exceed number of basis functions!"
    pts = [Ef, Ef - conv, Ef + conv]          # E2 (newest), E1, E0
    vals, bounds = [], (None, None)
    for E in pts:
        P, cnt = _probe(g, pMu, E)
        n = cnt - ne
        bounds = _track(bounds, E, n)
        if abs(n) < conv:
            return E, 0, P, abs(n), bounds[0], bounds[1]
        vals.append(n)
    (E2, E1, E0), (n2, n1, n0) = pts, vals
    counter = 3
    while counter < maxcycles:
        h0, h1 = E0 - E2, E1 - E2
        d0, d1 = n0 - n2, n1 - n2
        det = h0 * h1 * (h0 - h1)
        a = (d0 * h1 - h0 * d1) / det
        b = (h0 * h0 * d1 - h1 * h1 * d0) / det
        disc = np.sqrt(b * b - 4 * a * n2) if b * b > 4 * a * n2 else 0
        if b < 0:
            disc = -disc
        dE = -2 * n2 / (b + disc)
        Enext = E2 + dE
        if abs(Enext - E1) < abs(Enext - E0):     # keep the two old points that are nearest to the new one
            E0, E1, n0, n1 = E1, E0, n1, n0
        if abs(Enext - E2) < abs(Enext - E1):
            E1, n1 = E2, n2
        E2 = Enext
        P, cnt = _probe(g, pMu, E2)
        n2 = cnt - ne
        bounds = _track(bounds, E2, n2)
        if abs(n2) < conv:
            break
        if FERMI_DEBUG:
            print(f"DEBUG: Ef={E2:.2f}, dN={n2:.2E}, dE={dE:.2E}")
        counter += 1
    if counter == maxcycles:
        print(f'Warning: Max cycles reached, convergence = {abs(n2):.2E}')
    return E2, dE, P, abs(n2), bounds[0], bounds[1]


def calcFermiPolyFit(g, ne, Emin, Ef, N, tol=ADAPTIVE_INTEGRATION_TOL, conv=FERMI_CALCULATION_TOL,
                     maxcycles=FERMI_SEARCH_CYCLES, T=TEMPERATURE, order=3):
    """Root of a Huber-robust polynomial fit through all probes so far, smoothed by a shape-preserving
    PCHIP interpolant (density.py:1333-1515) -> (Ef, dE, P, |N - ne|, uBound, lBound)."""
    from scipy.interpolate import PchipInterpolator
    from scipy.optimize import least_squares
    nbf, pMu = _contour_density(g, Emin, N, tol, T)
    assert ne < nbf, "Number of electrons cannot exceed number of basis functions!"
    bounds = (None, None)
    E = Ef
    P, cnt = _probe(g, pMu, E)
    n = cnt - ne
    if abs(n) < conv:
        return E, 0, P, abs(n), None, None
    Es, ns = [E], [n]
    step, n_first, counter = conv * 10, n, 1
    while counter < maxcycles:                    # second probe far enough above the first to see N rise
        E = Ef + step
        P, cnt = _probe(g, pMu, E)
        n = cnt - ne
        bounds = _track(bounds, E, n)
        if abs(n) < conv:
            return E, step, P, abs(n), bounds[0], bounds[1]
        if n - n_first > 0:
            break
        step *= 10
        counter += 1
        if FERMI_DEBUG:
            print(f'Warning: Tried Ef = {E:2f} eV (too close to {Ef:2f} to get accurate dN {n-n_first:.2E})')
    Es.append(E)
    ns.append(n)
    dE = step
    while counter < maxcycles:
        deg = min(len(ns) - 1, order)
        Esort, nsort = zip(*sorted(zip(Es, ns)))
        smooth = PchipInterpolator(Esort, nsort)(Es)
        fit = least_squares(lambda cf: np.polyval(cf, Es) - smooth, np.polyfit(Es, ns, deg), loss='huber',
                            f_scale=ADAPTIVE_INTEGRATION_TOL)
        roots = np.roots(fit.x)
        E_next = roots[np.argmin(np.abs(roots - Es[-1]))].real
        wrong_way = (ns[-1] > 0 and E_next > Es[-1]) or (ns[-1] < 0 and E_next < Es[-1])
        if wrong_way:                             # N(mu) must rise with mu: drop the probe, step 10 dE the right way
            E_next = Es[-1] - abs(dE) * 10 if ns[-1] > 0 else Es[-1] + abs(dE) * 10
            Es.pop()
            ns.pop()
            counter -= 1
            if FERMI_DEBUG:
                print('Warning: monotonicity exception corrected!')
        E = E_next
        P, cnt = _probe(g, pMu, E)
        n = cnt - ne
        bounds = _track(bounds, E, n)
        Es.append(E)
        ns.append(n)
        dE = E - Es[-2]
        if abs(n) < conv:
            break
        if FERMI_DEBUG:
            print(f"Iter {counter}: E = {E:.6f}, n-ne = {n:.3e}, dE = {dE:.3e}, order = {deg}")
        counter += 1
    if counter >= maxcycles:
        print(f'Warning: Max cycles reached, convergence = {abs(n):.2E}')
    return E, dE, P, abs(n), bounds[0], bounds[1]


# ---- energy-INDEPENDENT analytic density (density.py:276-382) -----------------------
# No energy grid: with a constant self-energy the integral of G Gamma G^+ has a closed form in the eigenbasis of
# the (non-Hermitian) effective Hamiltonian.  Setup-time host linear algebra, one N x N eigenproblem per call site.
def density(V, Vc, D, Gam, Emin, mu):
    """Closed-form integral_{Emin}^{mu} G Gamma G^+ dE / 2 pi for an energy-independent Gamma (Eq. 27 of
    PRB 65, 165401; density.py:276-329).  V, D: eigenvectors / eigenvalues of the effective Hamiltonian in the
    orthogonalised basis, Vc = inv(V)^H."""
    D = np.asarray(D)

    def log_difference(limit):                    # log(1 - limit/D_i) - conj(log(1 - limit/D_j))
        lg = np.emath.log(1 - (limit / D))
        return lg[:, None] - lg.conj()[None, :]

    weights = (log_difference(mu) - log_difference(Emin)) / (2 * np.pi * (D[:, None] - D.conj()[None, :]))
    return V @ (weights * (Vc.conj().T @ Gam @ Vc)) @ V.conj().T


def bisectFermi(V, Vc, D, Gam, Nexp, conv=FERMI_CALCULATION_TOL, Eminf=ENERGY_MIN):
    """Chemical potential at which the analytic density holds Nexp electrons, by bisection between the lowest
    and highest eigenvalue (density.py:331-382)."""
    lo, hi = min(D.real), max(D.real)
    dN, Niter = Nexp, 0
    while abs(dN) > conv and Niter < 1000:
        fermi_level = (lo + hi) / 2
        dN = np.trace(density(V, Vc, D, Gam, Eminf, fermi_level)).real - Nexp
        if dN > 0:
            hi = fermi_level
        else:
            lo = fermi_level
        Niter += 1
    if Niter >= 1000:
        print('Warning: Bisection search timed out after 1000 iterations!')
    print(f'Bisection fermi search converged to {dN:.2E} in {Niter} iterations.')
    return fermi_level


def integratePoints(computePointFunc, numPoints, parallel=False, numWorkers=None, chunkSize=None, debug=False):
    """sum_i computePointFunc(i), i = 0 .. numPoints-1 (density.py:121-208).  The reference's process-pool branch
    exists to spread CPU solves over cores; here every point function already runs its linear algebra on the GPU
    (batch the points through GrInt / GrLessInt instead), so the points are simply accumulated in order and
    `parallel`, `numWorkers`, `chunkSize` are accepted for signature compatibility only."""
    if debug:
        print(f'Number of points to integrate: {numPoints}')
    result = np.zeros_like(computePointFunc(0))
    for i in range(int(numPoints)):
        result += computePointFunc(i)
    return result
