"""Full-size sanity of the remaining BASELINE configs (dev probe): cfg5-sized GrLessInt at N=2048, cfg2 densityComplex at N=256."""
import sys, time, io, contextlib
import numpy as np
sys.path.insert(0, ".")
from gaunegf_b200 import synthetic as sy, density as de
from gaunegf_b200._native import Context
from gaunegf_b200.surfGTester import surfGTest

def rel(a, b): return float(np.abs(a - b).max() / np.abs(b).max())

ctx = Context(0)
N, nc = 2048, 63
F, S = sy.hermitian_pair(N, seed=3)
i1, i2 = np.arange(nc), np.arange(N - nc, N)
rng = np.random.default_rng(5)
def blk():
    b = rng.standard_normal((nc, nc)) * 0.02
    return (b + b.T) / 2 - 0.1j * np.eye(nc)
b1, b2 = blk(), blk()
ctx.set_system(F, S); ctx.sigma_clear(); ctx.sigma_add_const_block(i1, b1); ctx.sigma_add_const_block(i2, b2)
M = 96
E = np.linspace(-0.25, 0.25, M); w = np.full(M, 0.5 / M)
for rep in range(2):
    t = time.perf_counter(); P = ctx.gless_int(E, w, -1); dt = time.perf_counter() - t
    print(f"N=2048 GrLessInt (both contacts, naug=126) M={M}: {dt*1e3:.1f} ms -> {M/dt:.1f} E/s, "
          f"{(8/3*N**3 + 16*N*N*2*nc)*M/dt/1e12:.1f} TF/s algorithmic", flush=True)
sig = np.zeros((N, N), complex); sig[np.ix_(i1, i1)] = b1; sig[np.ix_(i2, i2)] = b2
gam = 1j * (sig - sig.conj().T)
Pd = ctx.gless_int_dense(E[:8], w[:8], sig, gam)
P8 = ctx.gless_int(E[:8], w[:8], -1)
print("contact-column path vs dense-Gamma full-inverse path:", rel(P8, Pd))
z, wz = sy.contour_points(24, -30.0, 0.0)
t = time.perf_counter(); G = ctx.gr_int(z, wz); dt = time.perf_counter() - t
print(f"N=2048 GrInt M=24: {dt*1e3:.1f} ms -> {24/dt:.1f} E/s, {8*N**3*24/dt/1e12:.1f} TF/s algorithmic")
# cfg2
N = 256
F, S = sy.hermitian_pair(N, seed=0)
inds = sy.end_contacts(N, 16)
g = surfGTest(F, S, [list(inds[0]), list(inds[1])], -0.1j, -0.1j)
for rep in range(2):
    t = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()) as buf:
        P = de.densityComplex(F, S, g, -30.0, 0.0, 1e-4, 0.0)
    dt = time.perf_counter() - t
    print(f"cfg2 densityComplex N=256: {dt*1e3:.1f} ms, Tr(PS)={np.trace(P @ S).real:.6f}; {buf.getvalue().strip().splitlines()[-1]}")
