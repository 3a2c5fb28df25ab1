"""CPU oracle: a plain numpy restatement of GauNEGF's energy-grid Green's-function hot path.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this module; the product path (gaunegf_b200/) never does and
fails loudly when its CUDA library is missing.

Every function cites the reference file:line (under /root/reference/gauNEGF/) whose arithmetic
it restates.  Parity is PINNED: tests/test_oracle_golden.py checks this module against
(a) golden vectors produced by running the unmodified reference under the jax→numpy shim
(tests/golden/make_golden.py, make_golden_n2.py; fixtures committed under tests/golden/) and (b) the live reference
when /root/reference is present.  Arithmetic is numpy + LAPACK (zgesv), float64/complex128 — the
same algorithm family the reference's JAX-CPU path dispatches to (the reference ships no golden
values of its own: SURVEY.md §4).
"""
import numpy as np
from scipy.special import roots_legendre

# --- config.py:8-23 (values are part of parity) -------------------------------------------
ETA = 1e-6
SURFACE_GREEN_CONVERGENCE = 1e-5
SURFACE_RELAXATION_FACTOR = 0.1
ADAPTIVE_INTEGRATION_TOL = 1e-4
N_KT = 10
MAX_CYCLES = 1000
MAX_GRID_POINTS = 1000
ENERGY_STEP = 0.001
kB = 8.617e-5          # density.py:61, transport.py:36
eoverh = 3.874e-5      # transport.py:35


# --- utils.py:52-54 ------------------------------------------------------------------------
def inv(A):
    A = np.asarray(A)
    return np.linalg.solve(A, np.eye(A.shape[0]))


# --- integrate.py:67-82 --------------------------------------------------------------------
def gr_matrix(sigTot, E, F, S):
    return inv(E * S - F - sigTot)


def gless_matrix(sig, sigTot, E, F, S):
    Gr = inv(E * S - F - sigTot)
    gamma = 1j * (sig - sig.conj().T)
    return Gr @ gamma @ Gr.conj().T


# --- integrate.py:84-208 (vmap then sum(axis=0) == ordered accumulation) -------------------
def GrInt(F, S, g, Elist, weights):
    Elist = np.asarray(Elist)
    weights = np.asarray(weights)
    assert Elist.size == weights.size
    assert F.shape == S.shape and F.shape[0] == F.shape[1]
    acc = np.zeros(F.shape, dtype=complex)
    for E, w in zip(Elist, weights):
        acc = acc + w * gr_matrix(np.asarray(g.sigmaTot(E)), E, F, S)
    return acc


def GrLessInt(F, S, g, Elist, weights, ind=None):
    Elist = np.asarray(Elist)
    weights = np.asarray(weights)
    assert Elist.size == weights.size
    acc = np.zeros(F.shape, dtype=complex)
    for E, w in zip(Elist, weights):
        sigTot = np.asarray(g.sigmaTot(E))
        sig = sigTot if ind is None else np.asarray(g.sigma(E, ind))
        acc = acc + w * gless_matrix(sig, sigTot, E, F, S)
    return acc


# --- transport.py:40-146 -------------------------------------------------------------------
class SigmaCalculator:
    def __init__(self, sig1, sig2=None, energy_dependent=None):
        self.sig1, self.sig2 = sig1, sig2
        if energy_dependent is None:
            energy_dependent = hasattr(sig1, "sigma") and hasattr(sig1, "sigmaTot")
        self.energy_dependent = energy_dependent
        if energy_dependent and sig2 is not None:
            raise ValueError("For energy-dependent calculations, provide only surfG object as sig1")
        if not energy_dependent and sig2 is None:
            raise ValueError("For energy-independent calculations, provide both sig1 and sig2")

    @staticmethod
    def _expand(sig, spin, matrix_size):
        if spin in ("u", "ro", "g") and matrix_size is not None and matrix_size == 2 * sig.shape[0]:
            return np.kron(np.eye(2), sig) if spin in ("u", "ro") else np.kron(sig, np.eye(2))
        return sig

    def get_sigma_total(self, E, spin=None, matrix_size=None):
        if self.energy_dependent:
            tot = np.asarray(self.sig1.sigmaTot(E))
        else:
            a, b = np.asarray(self.sig1), np.asarray(self.sig2)
            tot = np.diag(a + b) if a.ndim == 1 else a + b
        return self._expand(tot, spin, matrix_size)

    def get_sigma(self, E, contact_index, spin=None, matrix_size=None):
        if self.energy_dependent:
            sig = np.asarray(self.sig1.sigma(E, contact_index))
        else:
            if contact_index == 0:
                raw = self.sig1
            elif contact_index in (-1, 1):
                raw = self.sig2
            else:
                raise ValueError(f"Invalid contact_index {contact_index}")
            raw = np.asarray(raw)
            sig = np.diag(raw) if raw.ndim == 1 else raw
        return self._expand(sig, spin, matrix_size)

    def get_gamma(self, E, contact_index, spin=None, matrix_size=None):
        sig = self.get_sigma(E, contact_index, spin, matrix_size)
        return 1j * (sig - sig.conj().T)


# --- transport.py:150-190 ------------------------------------------------------------------
def transmission_restricted(E, F, S, sigma_total, gamma1, gamma2):
    Gr = inv(E * S - F - sigma_total)
    return float(np.real(np.trace(gamma1 @ Gr @ gamma2 @ Gr.conj().T)))


def transmission_spin_block(E, F, S, sigma_total, gamma1, gamma2):
    Gr = inv(E * S - F - sigma_total)
    Ga = Gr.conj().T
    N = F.shape[0] // 2
    up, dn = slice(0, N), slice(N, 2 * N)
    blocks = [(up, up, up, up), (up, dn, up, dn), (dn, up, dn, up), (dn, dn, dn, dn)]
    # (row, col) of Gr/Ga block i; gamma1 uses the row spin, gamma2 the column spin (:170-173)
    Ts = []
    for (r, c, _, _) in blocks:
        t = gamma1[r, r] @ Gr[r, c] @ gamma2[c, c]
        Ts.append(float(np.real(np.trace(t @ Ga[r, c]))))
    return float(np.sum(Ts)), np.array(Ts)


def dos_kernel(E, F, S, sigma_total):
    Gr = inv(E * S - F - sigma_total)
    per_site = -np.imag(np.diag(Gr)) / np.pi
    return float(np.sum(per_site)), per_site


def compute_dos_at_energy(E, F, S, sigma_total):   # density.py:49-54
    return float(-np.imag(np.trace(inv(E * S - F - sigma_total))) / np.pi)


# --- transport.py:193-271, 376-483 (no checkpointing here: that is host logic, not arithmetic)
def transmission_single_energy(E, F, S, calc, spin=None):
    spin = spin or "r"
    n = F.shape[0]
    st = calc.get_sigma_total(E, spin, n)
    g1 = calc.get_gamma(E, 0, spin, n)
    g2 = calc.get_gamma(E, -1, spin, n)
    if spin == "r":
        return transmission_restricted(E, F, S, st, g1, g2)
    if spin in ("u", "ro"):
        return transmission_spin_block(E, F, S, st, g1, g2)
    if spin == "g":
        N = n // 2
        p = np.concatenate([np.arange(0, 2 * N, 2), np.arange(1, 2 * N, 2)])
        ix = np.ix_(p, p)
        return transmission_spin_block(E, F[ix], S[ix], st[ix], g1[ix], g2[ix])
    raise ValueError(f"Unknown spin configuration '{spin}'. Use 'r', 'u', 'ro', or 'g'")


def calculate_transmission(F, S, calc, energy_list, spin=None):
    spin = spin or "r"
    energy_list = np.asarray(energy_list)
    T = np.zeros(len(energy_list))
    Ts = np.zeros((len(energy_list), 4)) if spin != "r" else None
    for i, E in enumerate(energy_list):
        r = transmission_single_energy(E, F, S, calc, spin)
        if isinstance(r, tuple):
            T[i], Ts[i] = r[0], r[1]
        else:
            T[i] = r
    return T if Ts is None else (T, Ts)


def calculate_dos(F, S, calc, energy_list):        # transport.py:486-607, spin 'r'
    energy_list = np.asarray(energy_list)
    tot = np.zeros(len(energy_list))
    per = np.zeros((len(energy_list), F.shape[0]))
    for i, E in enumerate(energy_list):
        tot[i], per[i] = dos_kernel(E, F, S, calc.get_sigma_total(E, "r", F.shape[0]))
    return tot, per


def current_grid(fermi, qV, T=0.0, dE=ENERGY_STEP):  # transport.py:652-675
    dE = -abs(dE) if qV < 0 else abs(dE)
    muL, muR = fermi - qV / 2, fermi + qV / 2
    if T == 0:
        return np.arange(muL, muR, dE), muL, muR
    spread = np.sign(dE) * N_KT * kB * T
    return np.arange(muL - spread, muR + spread, dE), muL, muR


def calculate_current(F, S, calc, fermi, qV, T=0.0, spin="r", dE=ENERGY_STEP):  # :610-720
    from scipy.integrate import trapezoid
    if np.allclose(0, qV):
        return 0.0
    grid, muL, muR = current_grid(fermi, qV, T, dE)
    if len(grid) == 0:
        raise ValueError("No energies in integration window. Check fermi, qV, and dE.")
    trans = calculate_transmission(F, S, calc, grid, spin="r")
    if T == 0:
        I = eoverh * trapezoid(trans, grid)
    else:
        df = np.abs(1 / (np.exp((grid - muR) / (kB * T)) + 1) - 1 / (np.exp((grid - muL) / (kB * T)) + 1))
        I = eoverh * trapezoid(trans * df, grid)
    return 2 * I if spin == "r" else I


# --- density.py:64-119 ---------------------------------------------------------------------
def fermi(E, mu, T):
    kT = kB * T
    if kT == 0:
        return (E <= mu) * 1
    return 1 / (np.exp((E - mu) / kT) + 1)


def getANTPoints(N):
    k = np.arange(1, N + 1, 2)
    th = k * np.pi / (2 * N)
    s, c = np.sin(th), np.cos(th)
    x = 1.0 + 0.21220659078919378103 * s * c * (3 + 2 * s * s) - k / N
    w = s ** 4 * 16.0 / (3 * N)
    return np.concatenate((x, -x)), np.concatenate((w, w))


# --- density.py:211-273 --------------------------------------------------------------------
def integratePointsAdaptiveANT(computePoint, tol=ADAPTIVE_INTEGRATION_TOL, maxN=MAX_GRID_POINTS,
                               trace=None):
    prev_x = prev_sumW = P = new_P = None
    N = 2
    while N <= maxN:
        x, w = getANTPoints(N)
        if prev_x is None:
            P = computePoint(x[0:2], w[0:2])
        else:
            old = np.isin(np.round(x, 14), np.round(prev_x, 14))
            assert int(old.sum()) == prev_x.size, "Old nodes mismatch"
            ratio = float(np.sum(w[old]) / prev_sumW)
            new_P = P * ratio
            new_P = new_P + computePoint(x[~old], w[~old])
            maxDP = np.max(np.abs(new_P - P))
            P = new_P.copy()
            if trace is not None:
                trace.append((N, float(maxDP)))
            if maxDP < tol:
                return new_P
        prev_x, prev_sumW = x, float(np.sum(w))
        N *= 3
    return new_P


# --- density.py:385-484 --------------------------------------------------------------------
def densityRealN(F, S, g, Emin, mu, N=100, T=0.0):
    Emax = mu + N_KT * kB * T
    mid = (Emax - Emin) / 2
    x, w = roots_legendre(N)
    x = np.real(x)
    E = mid * (x + 1) + Emin
    wts = mid * w * fermi(E, mu, T)
    return (-1 + 0j) * np.imag(GrInt(F, S, g, E, wts)) / np.pi


def densityReal(F, S, g, Emin, mu, tol=ADAPTIVE_INTEGRATION_TOL, T=0.0, maxN=MAX_CYCLES):
    P = np.zeros_like(F)
    N = 1
    while N < maxN:
        P_prev = P.copy()
        P = densityRealN(F, S, g, Emin, mu, N, T)
        if np.max(np.abs(P - P_prev)) < tol:
            return P
        N *= 2
    return P


# --- density.py:487-658 --------------------------------------------------------------------
def _grid_window(mu1, mu2, T):
    muLo, muHi = min(mu1, mu2), max(mu1, mu2)
    dInt = np.sign(mu2 - mu1)
    Emax, Emin = muHi + N_KT * kB * T, muLo - N_KT * kB * T
    return muLo, muHi, dInt, Emin, (Emax - Emin) / 2


def densityGridN(F, S, g, mu1, mu2, ind=None, N=100, T=0.0):
    muLo, muHi, dInt, Emin, mid = _grid_window(mu1, mu2, T)
    x, w = roots_legendre(N)
    x = np.real(x)
    E = mid * (x + 1) + Emin
    wts = mid * w * (fermi(E, muHi, T) - fermi(E, muLo, T)) * dInt
    return GrLessInt(F, S, g, E, wts, ind) / (2 * np.pi)


def densityGrid(F, S, g, mu1, mu2, ind=None, tol=ADAPTIVE_INTEGRATION_TOL, T=0.0, trace=None):
    muLo, muHi, dInt, Emin, mid = _grid_window(mu1, mu2, T)

    def computePoint(x, w):
        E = mid * (x + 1) + Emin
        wts = mid * w * (fermi(E, muHi, T) - fermi(E, muLo, T)) * dInt
        return GrLessInt(F, S, g, E, wts, ind)

    return integratePointsAdaptiveANT(computePoint, tol=tol, trace=trace) / (2 * np.pi)


# --- density.py:660-816 --------------------------------------------------------------------
def _contour(Emin, mu, T):
    broad = 10 * kB * T
    Emax = mu - broad
    return (Emin + Emax) / 2, (Emax - Emin) / 2, broad


def densityComplexN(F, S, g, Emin, mu, N=100, T=0.0, method="ant"):
    center, r, broad = _contour(Emin, mu, T)
    if method == "legendre":
        x, w = roots_legendre(N)
    elif method == "chebyshev":
        k = np.arange(1, N + 1)
        x = np.cos(k * np.pi / (N + 1))
        w = (np.pi / (N + 1)) * (np.sin(k * np.pi / (N + 1)) ** 2) / np.sqrt(1 - x ** 2)
    elif method == "ant":
        x, w = getANTPoints(N)
    else:
        x = np.linspace(-1, 1, N)
        w = 2 * np.ones(N) / N
    th = np.pi / 2 * (x + 1)
    z = center + r * np.exp(1j * th)
    dz = 1j * r * np.exp(1j * th)
    line = GrInt(F, S, g, z, (np.pi / 2) * w * fermi(z, mu, T) * dz)
    if T > 0:
        Nb = int(N // 8)
        if method in ("legendre", "chebyshev", "ant"):
            xf, wf = roots_legendre(Nb)
        else:
            xf, wf = np.linspace(-1, 1, Nb), 2 * np.ones(Nb) / Nb
        Eb = broad * xf + mu
        line = line + GrInt(F, S, g, Eb, broad * wf * fermi(Eb, mu, T))
    return (1 + 0j) * np.imag(line) / np.pi


def densityComplex(F, S, g, Emin, mu, tol=ADAPTIVE_INTEGRATION_TOL, T=0.0, trace=None):
    center, r, broad = _contour(Emin, mu, T)

    def computePoint(x, w):
        th = np.pi / 2 * (x + 1)
        z = center + r * np.exp(1j * th)
        dz = 1j * r * np.exp(1j * th)
        return GrInt(F, S, g, z, (np.pi / 2) * w * dz * fermi(z, mu, T))

    line = integratePointsAdaptiveANT(computePoint, tol=tol, trace=trace)
    if T > 0:
        def computeBroad(x, w):
            E = broad * x + mu
            return GrInt(F, S, g, E, broad * w * fermi(E, mu, T))
        line = line + integratePointsAdaptiveANT(computeBroad, tol=tol)
    return (1 + 0j) * np.imag(line) / np.pi


# --- matTools.py:39-74, surfGTester.py:62-132 ----------------------------------------------
def formSigma(inds, V, nsto, S=0):
    if isinstance(S, int):
        S = np.eye(nsto)
    sigma = np.array(-1j * 1e-9 * S, dtype=complex)
    if isinstance(V, (int, complex, float)):
        for i in inds:
            sigma[i, i] = V
    else:
        sigma[np.ix_(inds, inds)] = V
    return sigma


class surfGTest:
    def __init__(self, Fock, Overlap, indsList, sig1, sig2=None):
        self.F, self.S, self.N, self.indsList = Fock, Overlap, len(Fock), indsList
        self.sig = [formSigma(indsList[0], sig1, self.N, Overlap),
                    formSigma(indsList[1], sig1 if sig2 is None else sig2, self.N, Overlap)]

    def sigma(self, E, i, conv=SURFACE_GREEN_CONVERGENCE):
        return self.sig[i]

    def sigmaTot(self, E, conv=SURFACE_GREEN_CONVERGENCE):
        tot = np.zeros((self.N, self.N), dtype=complex)
        for i in range(len(self.indsList)):
            tot += self.sig[i]
        return tot

    def setF(self, F, mu1=None, mu2=None):
        self.F = F


# --- surfG1D.py:83-399 ---------------------------------------------------------------------
class surfG1D:
    """1-D chain contacts: damped fixed point g <- 0.1*inv(A - B g B^H) + 0.9*g (NOT Sancho-Rubio)."""
    MAX_ITER = 2000

    def __init__(self, Fock, Overlap, indsList, taus=None, staus=None, alphas=None, aOverlaps=None,
                 betas=None, bOverlaps=None, eta=ETA):
        self.F, self.S = np.array(Fock), np.array(Overlap)
        self.indsList = [np.array(i) for i in indsList]
        if taus is None:
            taus = [self.indsList[-1], self.indsList[0]]
        taus = [np.array(t) for t in taus]
        if taus[0].ndim == 1:     # index form (:136-140)
            self.tauList = [self.F[np.ix_(taus[0], self.indsList[0])], self.F[np.ix_(taus[1], self.indsList[-1])]]
            self.stauList = [self.S[np.ix_(taus[0], self.indsList[0])], self.S[np.ix_(taus[1], self.indsList[-1])]]
        else:
            self.tauList = taus
            self.stauList = [np.array(s) for s in staus]
        if alphas is None:        # contactFromFock (:200-218): beta == tau
            self.aList = [self.F[np.ix_(i, i)] for i in self.indsList]
            self.aSList = [self.S[np.ix_(i, i)] for i in self.indsList]
            self.bList = [np.array(t) for t in self.tauList]
            self.bSList = [np.array(s) for s in self.stauList]
        else:
            self.aList = [np.array(a) for a in alphas]
            self.aSList = [np.array(a) for a in aOverlaps]
            self.bList = [np.array(b) for b in betas]
            self.bSList = [np.array(b) for b in bOverlaps]
        self.eta = eta
        self.num_contacts = len(indsList)
        self.last_iters = {}

    def g(self, E, i, conv=SURFACE_GREEN_CONVERGENCE, relFactor=SURFACE_RELAXATION_FACTOR):
        z = E + 1j * self.eta
        A = z * self.aSList[i] - self.aList[i]
        B = z * self.bSList[i] - self.bList[i]
        Bd = B.conj().T
        g = inv(A)
        count, diff = 0, np.inf
        while diff > conv and count < self.MAX_ITER:
            g_new = inv(A - B @ g @ Bd)
            diff = np.max(np.abs(g_new - g) / np.maximum(np.abs(g_new), 1e-12))
            g = g_new * relFactor + g * (1 - relFactor)
            count += 1
        self.last_iters[(complex(E), i)] = (count, float(diff))
        return g

    def sigma(self, E, i, conv=SURFACE_GREEN_CONVERGENCE):
        out = np.zeros(self.F.shape, dtype=complex)
        inds = self.indsList[i]
        t = E * self.stauList[i] - self.tauList[i]       # no i*eta here (:370)
        out[np.ix_(inds, inds)] += t @ self.g(E, i, conv) @ t.conj().T
        return out

    def sigmaTot(self, E, conv=SURFACE_GREEN_CONVERGENCE):
        out = np.zeros(self.F.shape, dtype=complex)
        for i in range(self.num_contacts):
            out = out + self.sigma(E, i, conv)
        return out


# --- surfGBethe.py:832-1155 ----------------------------------------------------------------
class surfGBAt:
    """Bethe-lattice atom: 12-direction bulk fixed point + 9-direction surface fixed point."""
    dim, NN, MAX_ITER = 9, 12, 1000

    def __init__(self, H, Slist, Vlist, eta, T=0.0):
        self.H = np.array(H)
        self.Slist = [np.array(s) for s in Slist]
        self.Vlist = [np.array(v) for v in Vlist]
        self.eta, self.T = eta, T
        self.last_iters = {}

    def sigmaK(self, E, conv=SURFACE_GREEN_CONVERGENCE, mix=0.5):
        z = E - 1j * self.eta                             # E MINUS i*eta (:995)
        sigK = np.array([-1j * np.eye(self.dim) for _ in range(self.NN)], dtype=complex)
        A = z * np.eye(self.dim) - self.H
        count, diff = 0, np.inf
        while diff > conv and count < self.MAX_ITER:
            old = sigK.copy()
            tot = np.sum(sigK, axis=0)                   # frozen for the whole sweep (:1007)
            for k in range(self.NN):
                gK = np.linalg.inv(A - tot + sigK[(k + 6) % 12])   # in-place updated sigK (:1011)
                B = z * self.Slist[k] - self.Vlist[k]
                sigK[k] = mix * (B @ gK @ B.conj().T) + (1 - mix) * old[k]
            diff = np.max(np.abs(sigK - old)) / np.max(np.abs(old))
            count += 1
        self.last_iters[("K", complex(E))] = (count, float(diff))
        return sigK

    def sigma(self, E, conv=SURFACE_GREEN_CONVERGENCE, mix=0.5):
        z = E - 1j * self.eta
        sig = self.sigmaK(E, conv, mix)[:9].copy()
        A = z * np.eye(self.dim) - self.H
        count, diff = 0, np.inf
        while diff > conv and count < self.MAX_ITER:
            old = sig.copy()
            g = np.linalg.inv(A - np.sum(sig, axis=0))
            for k in (0, 1, 2, 6, 7, 8):
                B = z * self.Slist[k] - self.Vlist[k]
                sig[k] = mix * (B @ g @ B.conj().T) + (1 - mix) * old[k]
            diff = np.max(np.abs(sig - old)) / np.max(np.abs(old))
            count += 1
        self.last_iters[("S", complex(E))] = (count, float(diff))
        return sig

    def sigmaTot(self, E, conv=SURFACE_GREEN_CONVERGENCE):
        d, n = self.dim, self.NN
        out = np.zeros(((n + 1) * d, (n + 1) * d), dtype=complex)
        sigK = self.sigmaK(E, conv)
        tot = np.sum(sigK, axis=0)
        for k in range(n):
            out[k * d:(k + 1) * d, k * d:(k + 1) * d] = tot - sigK[(k + 6) % 12]
        return out

    def DOS(self, E):
        Gr = np.linalg.inv((E - 1j * self.eta) * np.eye(self.dim) - self.H - np.sum(self.sigma(E), axis=0))
        return float(-np.trace(Gr).imag / np.pi)


class surfGB:
    """N x N scatter of per-atom Bethe self-energies (surfGBethe.py:479-575).  The geometry /
    Slater-Koster setup (ctor, :106-477) is out of scope: the parts it produces are passed in."""

    def __init__(self, F, S, gList, indsLists, nIndLists, Xi=None, orthonormal=False, spin="r"):
        self.F, self.S, self.N = F, S, (len(F) if spin == "r" else len(F) // 2)   # spin: F is 2N x 2N (:114-123)
        self.gList, self.indsLists, self.nIndLists = gList, indsLists, nIndLists
        self.Xi, self.orthonormal, self.spin = Xi, orthonormal, spin

    def sigma(self, E, i, conv=SURFACE_GREEN_CONVERGENCE):
        sig = np.zeros((self.N, self.N), dtype=complex)
        surf = self.gList[i].sigma(E, conv)
        for nInds, Finds in zip(self.nIndLists[i], self.indsLists[i]):
            atom = np.sum(surf[:9], axis=0)
            for n in nInds:
                atom = atom - surf[n]
            sig[np.ix_(Finds, Finds)] = atom
        if self.orthonormal:
            sig = self.Xi @ sig @ self.Xi
        if self.spin in ("u", "ro"):
            sig = np.kron(np.eye(2), sig)
        elif self.spin == "g":
            sig = np.kron(sig, np.eye(2))
        return sig

    def sigmaTot(self, E, conv=SURFACE_GREEN_CONVERGENCE):
        return sum(self.sigma(E, i, conv) for i in range(len(self.indsLists)))
