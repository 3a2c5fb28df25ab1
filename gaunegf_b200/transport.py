"""Coherent transport on the B200 — drop-in for gauNEGF/transport.py.

Same public names, argument order, defaults, return types, prints and .npz checkpoint schema as the
reference (transport.py:40-1107).  What changes is where the arithmetic runs: the reference loops
over energies in Python, one jitted `inv` + three dense N^3 products per point
(transport.py:150-157, 452-469); here every not-yet-computed energy of a call goes to the GPU as
one batch, and T(E) = Tr[Gamma1 G Gamma2 G^H] uses only the contact columns of G (Gamma is non-zero
on the contact orbitals only), i.e. a forward elimination + n_c-column back-substitution instead
of a full inverse.
"""
import os

import numpy as np
import scipy.io as io
from scipy.integrate import trapezoid

from . import parallel
from ._native import default_context
from .config import ENERGY_STEP, N_KT, TEMPERATURE
from .sigma_plan import DESC, DENSE_CONST, ArrayPlan, ObjectPlan, gamma_of

# CONSTANTS (transport.py:33-37)
har_to_eV = 27.211386   # eV/Hartree
eoverh = 3.874e-5       # A/eV
kB = 8.617e-5           # eV/Kelvin
V_to_au = 0.03675       # Volts to Hartree/elementary Charge

_SPINS = ('r', 'u', 'ro', 'g')


class SigmaCalculator:
    """Unified access to energy-independent (arrays) and energy-dependent (surfG objects)
    self-energies (transport.py:40-146)."""

    def __init__(self, sig1, sig2=None, energy_dependent=None):
        self.sig1 = sig1
        self.sig2 = sig2
        if energy_dependent is None:
            self.energy_dependent = hasattr(sig1, 'sigma') and hasattr(sig1, 'sigmaTot')
        else:
            self.energy_dependent = energy_dependent
        if self.energy_dependent and sig2 is not None:
            raise ValueError("For energy-dependent calculations, provide only surfG object as sig1")
        if not self.energy_dependent and sig2 is None:
            raise ValueError("For energy-independent calculations, provide both sig1 and sig2")

    # -- host-side views (used by callers that want the matrices; the batch drivers below do not
    #    materialise N x N matrices per energy unless the sigma object is an opaque Python callable)
    @staticmethod
    def _spin_expand(mat, spin, matrix_size):
        if spin in ('u', 'ro', 'g') and matrix_size is not None and matrix_size == 2 * mat.shape[0]:
            if spin in ('u', 'ro'):
                return np.kron(np.eye(2), mat)
            return np.kron(mat, np.eye(2))
        return mat

    def get_sigma_total(self, E, spin=None, matrix_size=None):
        if self.energy_dependent:
            total = np.asarray(self.sig1.sigmaTot(E))
        else:
            a, b = np.asarray(self.sig1), np.asarray(self.sig2)
            total = np.diag(a + b) if a.ndim == 1 else a + b
        return self._spin_expand(total, spin, matrix_size)

    def get_sigma(self, E, contact_index, spin=None, matrix_size=None):
        if self.energy_dependent:
            sigma = np.asarray(self.sig1.sigma(E, contact_index))
        else:
            if contact_index == 0:
                raw = self.sig1
            elif contact_index == -1 or contact_index == 1:
                raw = self.sig2
            else:
                raise ValueError(f"Invalid contact_index {contact_index}")
            raw = np.asarray(raw)
            sigma = np.diag(raw) if raw.ndim == 1 else raw
        return self._spin_expand(sigma, spin, matrix_size)

    def get_gamma(self, E, contact_index, spin=None, matrix_size=None):
        sigma = self.get_sigma(E, contact_index, spin, matrix_size)
        return 1j * (sigma - np.conj(sigma).T)

    # -- device plan
    def _plan(self, N):
        if self.energy_dependent:
            return ObjectPlan(self.sig1, N)
        return ArrayPlan([self.sig1, self.sig2], N)


def _described_spin_mode(plan, g, spin, n):
    """spin mode (1: kron(I2, .), 2: kron(., I2)) when a surfGB-shaped object's spin expansion can run on the
    device for a 2N x 2N system treated with `spin`; 0 otherwise (host evaluation of g.sigmaTot / g.sigma)"""
    from .surfGBethe import is_bethe_object
    if plan.kind != DESC or not is_bethe_object(g) or spin not in ('u', 'ro', 'g'):
        return 0
    own = getattr(g, "spin", 'r')
    want = 1 if spin in ('u', 'ro') else 2
    if own == 'r':                       # SigmaCalculator expands an N x N Sigma itself (transport.py:92-103)
        return want if getattr(g, "N", None) is not None and 2 * g.N == n else 0
    return want if (1 if own in ('u', 'ro') else 2) == want else 0


def _batched_sigma(calc, energies, spin, n, which):
    if which == 'tot':
        return np.stack([calc.get_sigma_total(E, spin, n) for E in energies]).astype(complex)
    return np.stack([calc.get_gamma(E, which, spin, n) for E in energies]).astype(complex)


def _transmission_batch(F, S, calc, energies, spin):
    """T(E) for a batch of energies on the GPU. Returns (M,) or ((M,), (M,4)) for spin != 'r'."""
    ctx = default_context()
    n = np.shape(F)[0]
    energies = np.asarray(energies, dtype=float)
    if spin == 'r':
        parallel.set_system(ctx, F, S)
        plan = calc._plan(n)
        plan.install(ctx)
        if plan.kind == DESC:
            return parallel.sharded_per_energy(energies, lambda E: ctx.transmission(E, 0, -1))
        if plan.kind == DENSE_CONST:
            st = plan.sigma_total()
            g1, g2 = gamma_of(plan.sigma(None, 0)), gamma_of(plan.sigma(None, -1))
            return parallel.sharded_per_energy(energies, lambda E: ctx.transmission_dense(E, st, g1, g2))

        def generic(E):
            out = np.empty(E.size)
            step = max(1, (1 << 30) // (48 * n * n))
            for k in range(0, E.size, step):
                Ek = E[k:k + step]
                st, g1, g2 = (_batched_sigma(calc, Ek, spin, n, which) for which in ('tot', 0, -1))
                ctx.set_system(F, S)       # the provider's sigma() may itself have used this context
                ctx.sigma_clear()
                out[k:k + step] = ctx.transmission_dense(Ek, st, g1, g2)
            return out
        return parallel.sharded_per_energy(energies, generic)

    if spin not in ('u', 'ro', 'g'):
        raise ValueError(f"Unknown spin configuration '{spin}'. Use 'r', 'u', 'ro', or 'g'")
    # spin-resolved (transport.py:159-181, 247-269): one 2N x 2N inverse, four spin-block traces
    Fm, Sm = np.asarray(F), np.asarray(S)
    perm = None
    if spin == 'g':           # spinor -> block order (transport.py:257-268)
        half = n // 2
        perm = np.concatenate([np.arange(0, 2 * half, 2), np.arange(1, 2 * half, 2)])
        ix = np.ix_(perm, perm)
        Fm, Sm = Fm[ix], Sm[ix]
    parallel.set_system(ctx, Fm, Sm)
    ctx.sigma_clear()
    if calc.energy_dependent:
        plan = calc._plan(n)
        mode = _described_spin_mode(plan, calc.sig1, spin, n)
        if mode:
            # Bethe contacts: Sigma blocks, Xi Sigma Xi and the spin expansion all on the device; after the
            # spinor -> block reordering above a 'g' expansion kron(Sigma, I2) reads kron(I2, Sigma)
            plan.install(ctx, spin_mode=1)
            T4 = parallel.sharded_per_energy(energies, ctx.transmission_spin_described, width=4)
            return T4.sum(axis=1), T4

    def spin_fn(E):
        const = not calc.energy_dependent
        Es = E[:1] if const else E
        st = _batched_sigma(calc, Es, spin, n, 'tot')
        g1 = _batched_sigma(calc, Es, spin, n, 0)
        g2 = _batched_sigma(calc, Es, spin, n, -1)
        if perm is not None:
            st, g1, g2 = (a[:, perm][:, :, perm] for a in (st, g1, g2))
        if const:
            st, g1, g2 = st[0], g1[0], g2[0]
        ctx.set_system(Fm, Sm)
        ctx.sigma_clear()
        return ctx.transmission_spin(E, st, g1, g2)

    T4 = parallel.sharded_per_energy(energies, spin_fn, width=4)
    return T4.sum(axis=1), T4


def transmission_single_energy(E, F_jax, S_jax, sigma_calc, spin=None):
    """T at one energy (transport.py:193-271): float for 'r', (total, [4 spin blocks]) otherwise."""
    spin = spin or 'r'
    res = _transmission_batch(F_jax, S_jax, sigma_calc, np.array([E]), spin)
    if isinstance(res, tuple):
        return float(res[0][0]), res[1][0].tolist()
    return float(res[0])


def _dos_batch(F, S, calc, energies, spin):
    ctx = default_context()
    n = np.shape(F)[0]
    energies = np.asarray(energies, dtype=float)
    if spin not in _SPINS:
        raise ValueError(f"Unknown spin configuration '{spin}'. Use 'r', 'u', 'ro', or 'g'")
    parallel.set_system(ctx, F, S)

    def _host_sigma_dos(E):        # Sigma(E) from the provider on the host, inverse and reduction on the GPU
        st = _batched_sigma(calc, E, spin, n, 'tot')
        ctx.set_system(F, S)       # the provider's sigmaTot() may itself have used this context
        ctx.sigma_clear()
        return np.column_stack(ctx.dos_dense(E, st)[::-1])

    if spin == 'r':
        plan = calc._plan(n)
        plan.install(ctx)
        if plan.kind == DESC:
            fn = lambda E: np.column_stack(ctx.dos(E)[::-1])                   # noqa: E731
        elif plan.kind == DENSE_CONST:
            st = plan.sigma_total()
            fn = lambda E: np.column_stack(ctx.dos_dense(E, st)[::-1])        # noqa: E731
        else:
            fn = _host_sigma_dos
    else:
        ctx.sigma_clear()
        mode = _described_spin_mode(calc._plan(n), calc.sig1, spin, n) if calc.energy_dependent else 0
        if mode:
            calc._plan(n).install(ctx, spin_mode=mode)
            fn = lambda E: np.column_stack(ctx.dos(E)[::-1])                   # noqa: E731
        elif calc.energy_dependent:
            fn = _host_sigma_dos
        else:
            st = calc.get_sigma_total(None, spin, n).astype(complex)
            fn = lambda E: np.column_stack(ctx.dos_dense(E, st)[::-1])        # noqa: E731
    res = parallel.sharded_per_energy(energies, fn, width=n + 1)             # columns: per-site..., total
    return res[:, n], res[:, :n]


def dos_single_energy(E, F_jax, S_jax, sigma_calc, spin=None):
    """DOS at one energy (transport.py:274-373)."""
    spin = spin or 'r'
    tot, per = _dos_batch(F_jax, S_jax, sigma_calc, np.array([E]), spin)
    per = per[0]
    if spin == 'r':
        return float(tot[0]), per
    half = len(per) // 2
    if spin in ('u', 'ro'):
        up, dn = per[:half], per[half:]
        return np.sum(up) + np.sum(dn), per, up, dn
    a, b = per[0::2], per[1::2]
    return np.sum(per), per, a, b


def _blocks(remaining, checkpoint_file, checkpoint_interval):
    """batches of not-yet-computed energies: everything at once without a checkpoint file, else the
    reference's write cadence (after local index 0, interval, 2*interval, ...: transport.py:464)"""
    if not checkpoint_file or len(remaining) == 0:
        return [remaining] if len(remaining) else []
    step = max(1, int(checkpoint_interval))
    cuts = [0, 1] + list(range(step + 1, len(remaining), step)) + [len(remaining)]
    cuts = sorted(set(c for c in cuts if c <= len(remaining)))
    return [remaining[a:b] for a, b in zip(cuts[:-1], cuts[1:]) if b > a]


def _load_checkpoint(checkpoint_file):
    """checkpoint contents as a dict of arrays, identical on every rank: rank 0 reads the file and broadcasts it
    (ranks that read a file another rank is replacing could disagree on what is left to do)"""
    def read():
        if checkpoint_file and os.path.exists(checkpoint_file):
            with np.load(checkpoint_file, allow_pickle=True) as data:
                return {k: np.array(data[k]) for k in data.files}
        return None
    return parallel.rank0_value(read)


def _save_checkpoint(checkpoint_file, **arrays):
    """rank 0 writes <file>.tmp and renames it over the checkpoint (never a half-written .npz)"""
    if parallel.dist_info()[0] != 0:
        return
    target = checkpoint_file if str(checkpoint_file).endswith(".npz") else str(checkpoint_file) + ".npz"   # np.savez's rule
    tmp = target + ".tmp.npz"
    np.savez(tmp, **arrays)
    os.replace(tmp, target)


def calculate_transmission(F, S, sigma_calculator, energy_list,
                           spin=None, checkpoint_file=None,
                           checkpoint_interval=10):
    """T(E) over an energy list with the reference's -1-sentinel .npz checkpointing
    (transport.py:376-483).  Not-yet-computed energies are evaluated on the GPU in batches."""
    energy_list = np.asarray(energy_list)
    n_energies = len(energy_list)
    if spin is None:
        spin = 'r'
    if spin not in _SPINS:
        raise ValueError(f"Unknown spin configuration '{spin}'. Use 'r', 'u', 'ro', or 'g'")
    open_shell = spin in ('u', 'ro', 'g')

    transmission = -1 * np.ones(n_energies)
    spin_trans = -1 * np.ones((n_energies, 4)) if open_shell else None
    data = _load_checkpoint(checkpoint_file)
    if data is not None:
        if 'energy_list' in data:
            saved = data['energy_list']
            if np.shape(saved) != np.shape(energy_list) or not np.allclose(saved, energy_list, rtol=1e-10):
                print("Warning: energy_list in checkpoint doesn't match. Starting fresh.")
                spin_trans = None if not open_shell else spin_trans
            else:
                if 'transmission' in data:
                    transmission = np.array(data['transmission'], dtype=float)
                if open_shell and 'spin_transmission' in data:
                    spin_trans = np.array(data['spin_transmission'], dtype=float)

    def save():
        if spin_trans is not None:
            _save_checkpoint(checkpoint_file, transmission=transmission, spin_transmission=spin_trans,
                             energy_list=energy_list)
        else:
            _save_checkpoint(checkpoint_file, transmission=transmission, energy_list=energy_list)

    remaining = np.where(transmission == -1)[0]
    for block in _blocks(remaining, checkpoint_file, checkpoint_interval):
        res = _transmission_batch(F, S, sigma_calculator, energy_list[block], spin)
        if isinstance(res, tuple):
            transmission[block] = res[0]
            spin_trans[block] = res[1]
        else:
            transmission[block] = res
        if checkpoint_file:
            save()
    if checkpoint_file:
        save()
    if spin_trans is not None:
        return transmission, spin_trans
    return transmission


def calculate_dos(F, S, sigma_calculator, energy_list,
                  spin=None, checkpoint_file=None,
                  checkpoint_interval=10):
    """DOS(E) with checkpointing (transport.py:486-607): (total (M,), per-site (M,N)[, spin (M,2)])."""
    energy_list = np.asarray(energy_list)
    n_energies = len(energy_list)
    n_sites = np.shape(F)[0]
    if spin is None:
        spin = 'r'
    if spin not in _SPINS:
        raise ValueError(f"Unknown spin configuration '{spin}'. Use 'r', 'u', 'ro', or 'g'")
    open_shell = spin in ('u', 'ro', 'g')
    dos_total = -1 * np.ones(n_energies)
    dos_per_site = -1 * np.ones((n_energies, n_sites))
    dos_spin = -1 * np.ones((n_energies, 2)) if open_shell else None
    data = _load_checkpoint(checkpoint_file)
    if data is not None:
        if 'energy_list' in data:
            saved = data['energy_list']
            if np.shape(saved) != np.shape(energy_list) or not np.allclose(saved, energy_list, rtol=1e-10):
                print("Warning: energy_list in checkpoint doesn't match. Starting fresh.")
            else:
                if 'dos_total' in data:
                    dos_total = np.array(data['dos_total'], dtype=float)
                if 'dos_per_site' in data:
                    dos_per_site = np.array(data['dos_per_site'], dtype=float)
                if open_shell and 'dos_spin' in data:
                    dos_spin = np.array(data['dos_spin'], dtype=float)

    def save():
        if dos_spin is not None:
            _save_checkpoint(checkpoint_file, dos_total=dos_total, dos_per_site=dos_per_site, dos_spin=dos_spin,
                             energy_list=energy_list)
        else:
            _save_checkpoint(checkpoint_file, dos_total=dos_total, dos_per_site=dos_per_site, energy_list=energy_list)

    remaining = np.where(dos_total == -1)[0]
    for block in _blocks(remaining, checkpoint_file, checkpoint_interval):
        tot, per = _dos_batch(F, S, sigma_calculator, energy_list[block], spin)
        dos_per_site[block] = per
        if open_shell:
            half = n_sites // 2
            if spin == 'g':
                up, dn = per[:, 0::2].sum(axis=1), per[:, 1::2].sum(axis=1)
                dos_total[block] = per.sum(axis=1)
            else:
                up, dn = per[:, :half].sum(axis=1), per[:, half:].sum(axis=1)
                dos_total[block] = up + dn
            dos_spin[block, 0], dos_spin[block, 1] = up, dn
        else:
            dos_total[block] = tot
        if checkpoint_file:
            save()
    if checkpoint_file:
        save()
    if dos_spin is not None:
        return dos_total, dos_per_site, dos_spin
    return dos_total, dos_per_site


def calculate_current(F, S, sigma_calculator, fermi, qV, T=TEMPERATURE, spin=None, dE=ENERGY_STEP,
                      **kwargs):
    """Landauer current at bias qV (transport.py:610-720): grid arange(muL, muR, dE) (widened by
    N_KT kT at T > 0), T(E) on the GPU, trapezoid on the host, x2 for spin 'r'."""
    if fermi is None or qV is None:
        raise ValueError("fermi and qV must be provided for current calculations")
    if spin is None:
        spin = 'r'
    if np.allclose(0, qV):
        return 0.0 if spin == 'r' else [0.0, 0.0, 0.0, 0.0]
    dE = -1 * abs(dE) if qV < 0 else abs(dE)
    muL = fermi - qV / 2
    muR = fermi + qV / 2
    if T == 0:
        grid = np.arange(muL, muR, dE)
    else:
        spread = np.sign(dE) * N_KT * kB * T
        grid = np.arange(muL - spread, muR + spread, dE)
    if len(grid) == 0:
        raise ValueError("No energies in integration window. Check fermi, qV, and dE.")
    result = calculate_transmission(F, S, sigma_calculator, grid, spin=spin, **kwargs)
    if isinstance(result, tuple):
        trans, spin_trans = np.asarray(result[0]), np.asarray(result[1])
    else:
        trans, spin_trans = np.asarray(result), None
    if T == 0:
        weight = 1.0
    else:
        weight = np.abs(1 / (np.exp((grid - muR) / (kB * T)) + 1) - 1 / (np.exp((grid - muL) / (kB * T)) + 1))
    if spin_trans is not None:
        current_spin = [eoverh * trapezoid(spin_trans[:, i] * weight, grid) for i in range(4)]
        return sum(current_spin), current_spin
    current_total = eoverh * trapezoid(trans * weight, grid)
    if spin == 'r':
        current_total *= 2
    return current_total


# ---- legacy API (transport.py:723-1107) --------------------------------------------------------
def current(F, S, sig1, sig2, fermi, qV, T=TEMPERATURE, spin="r", dE=ENERGY_STEP):
    sigma_calc = SigmaCalculator(sig1, sig2, energy_dependent=False)
    return calculate_current(F, S, sigma_calc, fermi=fermi, qV=qV, T=T, spin=spin, dE=dE)


def currentSpin(F, S, sig1, sig2, fermi, qV, T=TEMPERATURE, spin="r", dE=ENERGY_STEP):
    sigma_calc = SigmaCalculator(sig1, sig2, energy_dependent=False)
    result = calculate_current(F, S, sigma_calc, fermi=fermi, qV=qV, T=T, spin=spin, dE=dE)
    if isinstance(result, tuple):
        return result[1]
    return [0, 0, 0, 0]


def currentE(F, S, g, fermi, qV, T=TEMPERATURE, spin="r", dE=ENERGY_STEP):
    sigma_calc = SigmaCalculator(g, energy_dependent=True)
    return calculate_current(F, S, sigma_calc, fermi=fermi, qV=qV, T=T, spin=spin, dE=dE)


def currentF(fn, dE=ENERGY_STEP, T=TEMPERATURE):
    matfile = io.loadmat(fn)
    return current(matfile["F"], matfile["S"], matfile["sig1"], matfile["sig2"],
                   matfile["fermi"][0, 0], matfile["qV"][0, 0], T, matfile["spin"][0], dE=dE)


def _plain(seq):
    """float64 arrays as Python floats: str() gives the same text as for numpy scalars, several times faster"""
    return seq.tolist() if isinstance(seq, np.ndarray) and seq.dtype == np.float64 and seq.ndim == 1 else seq


def _report(Elist, label, values):
    """the reference prints one line per energy (transport.py:910-911, 1031-1032, 1104-1105); same text, one write"""
    lines = [f"Energy: {E} eV, {label}= {v}" for E, v in zip(_plain(Elist), _plain(values))]
    if lines:
        print("\n".join(lines))


def cohTrans(Elist, F, S, sig1, sig2):
    sigma_calc = SigmaCalculator(sig1, sig2, energy_dependent=False)
    transmissions = calculate_transmission(F, S, sigma_calc, Elist, spin='r')
    _report(Elist, "Transmission", transmissions)
    return transmissions.tolist()


def cohTransSpin(Elist, F, S, sig1, sig2, spin='u'):
    sigma_calc = SigmaCalculator(sig1, sig2, energy_dependent=False)
    result = calculate_transmission(F, S, sigma_calc, Elist, spin=spin)
    if isinstance(result, tuple):
        transmissions, spin_transmissions = result
        for i, E in enumerate(Elist):
            print("Energy:", E, "eV, Transmission=", transmissions[i], ", Tspin=", spin_transmissions[i])
        return (transmissions.tolist(), spin_transmissions)
    _report(Elist, "Transmission", result)
    return (result.tolist(), np.zeros((len(Elist), 4)))


def DOS(Elist, F, S, sig1, sig2):
    sigma_calc = SigmaCalculator(sig1, sig2, energy_dependent=False)
    dos_values, dos_per_site_list = calculate_dos(F, S, sigma_calc, Elist, spin='r')
    return dos_values.tolist(), dos_per_site_list


def cohTransE(Elist, F, S, g):
    sigma_calc = SigmaCalculator(g, energy_dependent=True)
    transmissions = calculate_transmission(F, S, sigma_calc, Elist, spin='r')
    _report(Elist, "Transmission", transmissions)
    return transmissions.tolist()


def cohTransSpinE(Elist, F, S, g, spin='u'):
    sigma_calc = SigmaCalculator(g, energy_dependent=True)
    result = calculate_transmission(F, S, sigma_calc, Elist, spin=spin)
    if isinstance(result, tuple):
        transmissions, spin_transmissions = result
        for i, E in enumerate(Elist):
            print("Energy:", E, "eV, Transmission=", transmissions[i], ", Tspin=", spin_transmissions[i])
        return transmissions, spin_transmissions
    _report(Elist, "Transmission", result)
    return result, np.zeros((len(Elist), 4))


def DOSE(Elist, F, S, g):
    sigma_calc = SigmaCalculator(g, energy_dependent=True)
    dos_values, dos_per_site_list = calculate_dos(F, S, sigma_calc, Elist, spin='r')
    _report(Elist, "DOS", dos_values)
    return dos_values.tolist(), dos_per_site_list


# ---- the reference's single-point kernels (transport.py:150-190), same names and signatures -------------------
# One energy per call, like the reference's jitted functions; the batched drivers above are the fast path.
def _one_energy_system(F, S):
    ctx = default_context()
    ctx.set_system(np.asarray(F), np.asarray(S))
    ctx.sigma_clear()
    return ctx


def _transmission_kernel_restricted(E, F, S, sigma_total, gamma1, gamma2):
    """Re Tr[Gamma1 G Gamma2 G^H] at one energy (transport.py:150-157)."""
    ctx = _one_energy_system(F, S)
    return float(ctx.transmission_dense(np.array([E]), np.asarray(sigma_total, dtype=complex),
                                        np.asarray(gamma1, dtype=complex), np.asarray(gamma2, dtype=complex))[0])


def _transmission_kernel_spin_block(E, F, S, sigma_total, gamma1, gamma2):
    """(total, [uu, ud, du, dd]) spin-block transmissions of a 2N x 2N system (transport.py:159-181)."""
    ctx = _one_energy_system(F, S)
    T4 = ctx.transmission_spin(np.array([E]), np.asarray(sigma_total, dtype=complex),
                               np.asarray(gamma1, dtype=complex), np.asarray(gamma2, dtype=complex))[0]
    return float(np.sum(T4)), np.asarray(T4)


def _dos_kernel(E, F, S, sigma_total):
    """(total, per-orbital) -Im diag(G)/pi at one energy (transport.py:183-190)."""
    ctx = _one_energy_system(F, S)
    tot, per = ctx.dos_dense(np.array([E]), np.asarray(sigma_total, dtype=complex))
    return float(tot[0]), per[0]
