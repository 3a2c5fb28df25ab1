"""Energy-batch integrators GrInt / GrLessInt — drop-in for gauNEGF/integrate.py:146-208.

The reference vmaps `solve(E S - F - Sigma, I)` over the energies on one device and sums M x N x N
temporaries (integrate.py:97-142).  Here the whole energy list goes to the B200 in one C-ABI call:
assembly, batched block elimination on the FP64 tensor pipe and the weighted reduction all happen on
the device and only the N x N result comes back.  With torch.distributed initialised (torchrun, one
process per GPU) the energies are sharded cyclically and the partial sums all-reduced once.
"""
import logging

import numpy as np

from . import parallel
from ._native import default_context
from .sigma_plan import DESC, DENSE_CONST, ObjectPlan, gamma_of

parallel_logger = logging.getLogger('gauNEGF.integrate')   # same logger name as the reference (:36)

# host-side staging bound for the generic (Python-callable sigma) path, bytes
_DENSE_STAGE_BYTES = 2 << 30


def _check(F, S, Elist, weights):
    Elist = np.asarray(Elist)
    weights = np.asarray(weights)
    assert Elist.size == weights.size, "Elist and weights must have the same length"
    assert np.shape(F) == np.shape(S), "F and S must have the same shape"
    assert np.shape(F)[0] == np.shape(F)[1], "F and S must be square matrices"
    return Elist.reshape(-1), weights.reshape(-1)


def _device_of(ctx):
    import torch
    return torch.device("cuda", ctx.device)


def _reinstall(ctx, F, S):
    ctx.set_system(F, S)
    ctx.sigma_clear()


def _gr_partial(ctx, plan, N, F, S):
    def run(E, w, out):
        ptr_ = out.data_ptr() if out is not None else None
        if plan.kind == DESC:
            return ctx.gr_int(E, w, out_device_ptr=ptr_)
        if plan.kind == DENSE_CONST:
            return ctx.gr_int_dense(E, w, plan.sigma_total(), out_device_ptr=ptr_)
        # generic surfG object: evaluate sigmaTot(E) on the host in bounded batches, invert on the GPU
        step = max(1, _DENSE_STAGE_BYTES // (16 * N * N))
        acc = np.zeros((N, N), dtype=complex)
        for k in range(0, E.size, step):
            st = plan.sigma_total_batch(E[k:k + step])
            _reinstall(ctx, F, S)          # g.sigmaTot may itself have used this context (e.g. a surfGBAt)
            acc += ctx.gr_int_dense(E[k:k + step], w[k:k + step], st)
        if out is not None:
            import torch
            out.copy_(torch.from_numpy(acc))
            return None
        return acc
    return run


def GrInt(F, S, g, Elist, weights):
    """sum_k weights[k] * G^R(Elist[k]),  G^R = (E S - F - g.sigmaTot(E))^-1   (integrate.py:146-173)"""
    Elist, weights = _check(F, S, Elist, weights)
    N = np.shape(F)[0]
    ctx = default_context()
    parallel.set_system(ctx, F, S)
    plan = ObjectPlan(g, N)
    plan.install(ctx)
    parallel_logger.info("Calculating G^R with GInt on B200: %dx%d, %d energies", N, N, Elist.size)
    run = _gr_partial(ctx, plan, N, F, S)
    _, world = parallel.dist_info()
    return parallel.sharded_matrix_sum(N, Elist, weights, run, device=_device_of(ctx) if world > 1 else None)


def GrIntLevels(F, S, g, levels):
    """[GrInt(F, S, g, E_l, w_l) for (E_l, w_l) in levels] evaluated as ONE batch on the GPU (gnb_gr_int_seg): the
    nested levels of the adaptive quadratures (density.py:234-270) are known before the convergence test of the current
    one, so several of them share one launch chain.  Each level's sum runs over that level's energies only and in
    their order, exactly as its own GrInt call would."""
    levels = [_check(F, S, E, w) for E, w in levels]
    N = np.shape(F)[0]
    ctx = default_context()
    parallel.set_system(ctx, F, S)
    plan = ObjectPlan(g, N)
    plan.install(ctx)
    if plan.kind not in (DESC, DENSE_CONST):            # host-evaluated Sigma objects: level by level
        return [GrInt(F, S, g, E, w) for E, w in levels]
    Eall = np.concatenate([E for E, _ in levels])
    wall = np.concatenate([w for _, w in levels])
    ends = np.cumsum([E.size for E, _ in levels])
    parallel_logger.info("Calculating G^R with GInt on B200: %dx%d, %d energies in %d levels", N, N, Eall.size, len(levels))

    def run(E, w, local_ends, out):
        ptr_ = out.data_ptr() if out is not None else None
        sig = None if plan.kind == DESC else plan.sigma_total()
        return ctx.gr_int_seg(E, w, local_ends, sig=sig, out_device_ptr=ptr_)

    _, world = parallel.dist_info()
    res = parallel.sharded_matrix_sums(N, Eall, wall, ends, run, device=_device_of(ctx) if world > 1 else None)
    return [res[i] for i in range(len(levels))]


def GrLessInt(F, S, g, Elist, weights, ind=None):
    """sum_k weights[k] * G^R Gamma G^A with Gamma = i(sigma - sigma^H); sigma = g.sigmaTot(E) if
    ind is None else g.sigma(E, ind)   (integrate.py:177-208)"""
    Elist, weights = _check(F, S, Elist, weights)
    N = np.shape(F)[0]
    ctx = default_context()
    parallel.set_system(ctx, F, S)
    plan = ObjectPlan(g, N)
    plan.install(ctx)
    nct = plan.ncontacts()
    parallel_logger.info("Calculating G< with GInt on B200: %dx%d, %d energies", N, N, Elist.size)

    def run(E, w, out):
        ptr_ = out.data_ptr() if out is not None else None
        if plan.kind == DESC:
            contact = -1 if ind is None else (ind + nct if ind < 0 else ind)
            return ctx.gless_int(E, w, contact, out_device_ptr=ptr_)
        if plan.kind == DENSE_CONST:
            sig = plan.sigma_total() if ind is None else plan.sigma(None, ind)
            return ctx.gless_int_dense(E, w, plan.sigma_total(), gamma_of(sig), out_device_ptr=ptr_)
        step = max(1, _DENSE_STAGE_BYTES // (32 * N * N))
        acc = np.zeros((N, N), dtype=complex)
        for k in range(0, E.size, step):
            Ek = E[k:k + step]
            st = plan.sigma_total_batch(Ek)
            sg = st if ind is None else plan.sigma_batch(Ek, ind)
            gam = 1j * (sg - sg.conj().transpose(0, 2, 1))
            _reinstall(ctx, F, S)
            acc += ctx.gless_int_dense(Ek, w[k:k + step], st, gam)
        if out is not None:
            import torch
            out.copy_(torch.from_numpy(acc))
            return None
        return acc

    _, world = parallel.dist_info()
    return parallel.sharded_matrix_sum(N, Elist, weights, run, device=_device_of(ctx) if world > 1 else None)


# ---- the reference's per-point matrix functions (integrate.py:67-82), same names and signatures ----------------
def _gr_matrix_ops(sigTot, E, F, S):
    """G^R(E) = (E S - F - Sigma)^-1 at one energy (integrate.py:67-71)."""
    ctx = default_context()
    ctx.set_system(np.asarray(F), np.asarray(S))
    ctx.sigma_clear()
    return ctx.green_dense(np.array([E], dtype=complex), np.asarray(sigTot, dtype=complex))[0]


def _gless_matrix_ops(sig, sigTot, E, F, S):
    """G^R i(sig - sig^H) G^A at one energy (integrate.py:74-82)."""
    ctx = default_context()
    ctx.set_system(np.asarray(F), np.asarray(S))
    ctx.sigma_clear()
    return ctx.gless_int_dense(np.array([E], dtype=complex), np.ones(1, dtype=complex), np.asarray(sigTot, dtype=complex),
                               gamma_of(np.asarray(sig, dtype=complex)))
