"""Dev check of the elimination engines against numpy on seeded systems: python tools/check_engine.py [N ...]"""
import sys
import numpy as np
sys.path.insert(0, ".")
from gaunegf_b200 import synthetic as sy
from gaunegf_b200._native import Context

Ns = [int(a) for a in sys.argv[1:]] or [64, 96, 100, 128, 160, 256, 600, 1024]
ctx = Context(0)
for N in Ns:
    nc = max(N // 16, 2)
    F, S = sy.hermitian_pair(N, seed=1)
    s1, s2 = sy.block_sigma_vectors(N, nc, 0.1)
    ctx.set_system(F, S); ctx.sigma_clear()
    ctx.sigma_add_const_block(np.arange(nc), np.diag(s1[:nc]))
    ctx.sigma_add_const_block(np.arange(N - nc, N), np.diag(s2[N - nc:]))
    E = np.array([0.1 + 0j, -1 + 2j, 0.37 + 1e-6j])
    G = ctx.green(E)
    T = ctx.transmission(E.real)
    for k, e in enumerate(E):
        A = e * S - F - np.diag(s1 + s2)
        G0 = np.linalg.inv(A)
        err = np.abs(G[k] - G0).max() / np.abs(G0).max()
        A = e.real * S - F - np.diag(s1 + s2)
        G0 = np.linalg.inv(A)
        g1 = np.diag(-2 * s1.imag); g2 = np.diag(-2 * s2.imag)
        T0 = np.trace(g1 @ G0 @ g2 @ G0.conj().T).real
        print(f"N={N} E={e}: jordan relerr {err:.1e}  T {T[k]:.12g} vs {T0:.12g} rel {abs(T[k]-T0)/max(abs(T0),1e-300):.1e}", flush=True)
    if err > 1e-8:
        D = np.abs(G[k] - np.linalg.inv(e * S - F - np.diag(s1 + s2)))
        bad = D > 1e-8 * np.abs(G0).max()
        rows = np.unique(np.nonzero(bad)[0] // 32); cols = np.unique(np.nonzero(bad)[1] // 32)
        print("   bad row blocks", rows, "bad col blocks", cols)
