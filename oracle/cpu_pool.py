"""Best-effort CPU arm of the reference algorithm: an energy-parallel process pool with ONE BLAS thread per worker.

TEST / BENCH INFRASTRUCTURE ONLY (bench.py's cpu_baseline and `--impl reference` legs).  This is the reference
author's own recipe for using all host cores on an energy grid (reference tests/benchmark_sigma_parallelization.py:
27-30 pins the BLAS threads, :178-212 maps the energies over a ProcessPoolExecutor); the per-energy arithmetic is the
numpy/LAPACK restatement in oracle/negf_oracle.py (transport.py:150-157: full solve(A, I) + Gamma1 G Gamma2 G^H).

Workers are spawned (never forked: the bench process may hold a CUDA context), rebuild the seeded synthetic inputs
themselves and keep them for the life of the pool, so a timed map() moves only energies and T(E) values.
"""
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_state = {}


def _init(n_orb, n_contact, seed):
    for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[v] = "1"
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    try:
        from threadpoolctl import threadpool_limits
        _state["limit"] = threadpool_limits(limits=1)
    except Exception:
        pass
    from gaunegf_b200 import synthetic as sy
    from oracle import negf_oracle as O
    F, S = sy.hermitian_pair(n_orb, seed=seed)
    s1, s2 = sy.block_sigma_vectors(n_orb, n_contact, 0.1)
    _state.update(F=F, S=S, calc=O.SigmaCalculator(s1, s2, energy_dependent=False), O=O)


def _work(energies):
    O = _state["O"]
    return O.calculate_transmission(_state["F"], _state["S"], _state["calc"], np.asarray(energies))


class TransmissionPool:
    """T(E) of the bench workload on `workers` single-threaded processes"""

    def __init__(self, n_orb, n_contact, seed, workers=None):
        self.workers = workers or os.cpu_count() or 1
        self.pool = mp.get_context("spawn").Pool(self.workers, initializer=_init, initargs=(n_orb, n_contact, seed))
        self.pool.map(_work, [[0.0]] * self.workers)          # every worker imported, built its inputs, warmed LAPACK

    def run(self, energies):
        """(seconds, T) for one pass over `energies`, split evenly over the workers"""
        parts = [p for p in np.array_split(np.asarray(energies, dtype=float), self.workers) if len(p)]
        t = time.perf_counter()
        res = self.pool.map(_work, [list(p) for p in parts], chunksize=1)
        dt = time.perf_counter() - t
        return dt, np.concatenate(res)

    def close(self):
        self.pool.close()
        self.pool.join()
