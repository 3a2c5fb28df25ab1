"""Accuracy of the T(E) path on ill-conditioned systems (energies at eigenvalues of the isolated device, weak contacts):
FP32 pivot order + FP64 inverse in the final tournament round (default) against the FP64 final round and numpy (dev tool)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from gaunegf_b200 import synthetic as sy
from gaunegf_b200._native import Context
from oracle import negf_oracle as O
N, nc = 320, 16
F, S = sy.hermitian_pair(N, seed=4)
ev = np.linalg.eigvalsh(np.linalg.solve(S, F))
ctx = Context(0)
for gam in (1e-1, 1e-3, 1e-5):
    s1 = np.zeros(N, complex); s2 = np.zeros(N, complex)
    s1[:nc] = -1j * gam; s2[N - nc:] = -1j * gam
    E = np.concatenate([ev[100:140] + 1e-9, ev[100:140] + 1e-6])
    st = np.diag(s1 + s2); g1 = np.diag(-2 * s1.imag); g2 = np.diag(-2 * s2.imag)
    Tref = np.array([O.transmission_restricted(e, F, S, st, g1, g2) for e in E])
    conds = np.array([np.linalg.cond(e * S - F - st) for e in E[:5]])
    out = {}
    for tw in (249, 25):
        ctx.lib.gnb_dev_set_option(b"tourn_warp", tw)
        ctx.set_system(F, S); ctx.sigma_clear()
        ctx.sigma_add_const_block(np.arange(nc), np.diag(s1[:nc]))
        ctx.sigma_add_const_block(np.arange(N - nc, N), np.diag(s2[N - nc:]))
        out[tw] = ctx.transmission(E, 0, -1)
    ctx.lib.gnb_dev_set_option(b"tourn_warp", 249)
    rel = lambda a: float(np.max(np.abs(a - Tref) / np.maximum(np.abs(Tref), 1e-300)))
    print(f"gamma={gam:g} cond~{conds.max():.1e}: max rel dev of T  FP32-order final {rel(out[249]):.2e}   FP64 final {rel(out[25]):.2e}   max T {Tref.max():.3g}")
