"""All five BASELINE.json configs at their stated sizes through the public API on ONE B200, with the numpy/LAPACK
port of the reference (oracle/) timed on the host cores on a bounded sample of the same energies.
Dev tool (not the driver's bench):  python tools/bench_configs.py  ->  gpurun_out/configs.json"""
import contextlib
import io
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from gaunegf_b200 import synthetic as sy, transport as tr, density as de, integrate as it      # noqa: E402
from gaunegf_b200.surfG1D import surfG                                                          # noqa: E402
from gaunegf_b200.surfGBethe import surfGB, surfGBAt                                            # noqa: E402
from gaunegf_b200.surfGTester import surfGTest                                                  # noqa: E402
from oracle import negf_oracle as O                                                             # noqa: E402

out = {"host_cores": os.cpu_count()}


def quiet(f, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return f(*a, **k)


def timed(f, reps=2):
    best = None
    for _ in range(reps):
        t = time.perf_counter()
        r = f()
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
    return r, best


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


def record(name, **kw):
    out[name] = kw
    print(name, json.dumps(kw), flush=True)


# cfg 1: cohTrans, 64-orbital chain, 1000 energies
F, S, s1, s2 = sy.chain(64)
E = np.linspace(-3, 3, 1000)
T, dt = timed(lambda: np.array(quiet(tr.cohTrans, E, F, S, s1, s2)), 3)
Tc, dtc = timed(lambda: O.calculate_transmission(F, S, O.SigmaCalculator(s1, s2), E), 1)
record("cfg1_cohTrans_chain64", energies=1000, gpu_s=dt, gpu_pts_per_s=1000 / dt, cpu_s=dtc, cpu_pts_per_s=1000 / dtc,
       cpu_sample="all 1000 energies", max_rel_diff=rel(T, Tc))

# cfg 2: densityComplex, N = 256 (adaptive contour, tol 1e-4)
N = 256
F, S = sy.hermitian_pair(N, seed=0)
inds = sy.end_contacts(N, 16)
g = surfGTest(F, S, [list(inds[0]), list(inds[1])], -0.1j, -0.1j)
P, dt = timed(lambda: quiet(de.densityComplex, F, S, g, -30.0, 0.0, 1e-4, 0.0))
og = O.surfGTest(F, S, inds, -0.1j, -0.1j)
Pc, dtc = timed(lambda: O.densityComplex(F, S, og, -30.0, 0.0, 1e-4, 0.0), 1)
record("cfg2_densityComplex_N256", gpu_s=dt, cpu_s=dtc, cpu_sample="the whole adaptive integral", speedup=dtc / dt,
       trace_PS=float(np.trace(P @ S).real), max_rel_diff=rel(P, Pc))
Pn, dt = timed(lambda: quiet(de.densityComplexN, F, S, g, -30.0, 0.0, 486, 0.0, False))
record("cfg2_densityComplexN486_N256", energies=486, gpu_s=dt, gpu_pts_per_s=486 / dt)

# cfg 3: T(E) at N = 1024 (the driver's bench.py measures this one with warm-up, events and clocks)
N = 1024
F, S = sy.hermitian_pair(N, seed=1)
s1, s2 = sy.block_sigma_vectors(N, 64, 0.1)
calc = tr.SigmaCalculator(s1, s2)
E = np.linspace(-0.5, 0.5, 1250)
T, dt = timed(lambda: tr.calculate_transmission(F, S, calc, E))
Tc, dtc = timed(lambda: O.calculate_transmission(F, S, O.SigmaCalculator(s1, s2), E[:8]), 1)
record("cfg3_transmission_N1024", energies=1250, gpu_s=dt, gpu_pts_per_s=1250 / dt, cpu_pts_per_s=8 / dtc,
       cpu_sample="first 8 energies", max_rel_diff=rel(T[:8], Tc))
I, dt = timed(lambda: tr.calculate_current(F, S, calc, 0.0, 0.5, 0.0, "r", 0.001), 1)
record("cfg3_current_qV0.5_N1024", energies=500, gpu_s=dt, current_A=float(I))

# cfg 4: surfG1D Sigma(E), N = 512 device + 128-orbital lead cells, 256 energies, eta = 1e-4
F, S, li, taus = sy.lead_device_lead(128, 512, seed=2, s_off=0.0)
g = surfG(F, S, [list(i) for i in li], [list(t) for t in taus], eta=1e-4)
E = np.linspace(-1, 1, 256)
T, dt = timed(lambda: np.array(quiet(tr.cohTransE, E, F, S, g)), 1)
g0, dtg = timed(lambda: g.g(E, 0), 1)
its = np.array([g.last_iters[(complex(e), 0)][0] for e in E])
og = O.surfG1D(F, S, li, taus, eta=1e-4)
Es = E[[3, 100, 200]]
Tc, dtc = timed(lambda: O.calculate_transmission(F, S, O.SigmaCalculator(og, energy_dependent=True), Es), 1)
itc = np.array([og.last_iters[(complex(e), 0)][0] for e in Es])
conv = itc < 2000
record("cfg4_surfG1D_N768_lead128", energies=256, gpu_s=dt, gpu_pts_per_s=256 / dt, sigma_only_s=dtg,
       iterations_min_med_max=[int(its.min()), int(np.median(its)), int(its.max())], cpu_pts_per_s=3 / dtc,
       cpu_sample="3 energies (indices 3, 100, 200)", iteration_counts_match=bool(np.array_equal(its[[3, 100, 200]], itc)),
       max_rel_diff_converged=(rel(T[[3, 100, 200]][conv], Tc[conv]) if conv.any() else None))

# cfg 5: Bethe-lattice contacts on N = 2048, densityGridN (contact parts from the reference's own constructor)
G = np.load(os.path.join("tests", "golden", "cfg5_bethe.npz"))
N = 2048
F, S = sy.hermitian_pair(N, seed=3)
gl = [surfGBAt(G["H"][i], G["Slist"][i], G["Vlist"][i], float(G["eta"])) for i in range(2)]
lens, flat = G["nInd_len"], list(G["nInd_flat"])
nil, p = [], 0
for c in lens:
    cl = []
    for n in c:
        cl.append(flat[p:p + n])
        p += n
    nil.append(cl)
indsLists = [[np.arange(9 * a, 9 * a + 9) for a in range(3)], [np.arange(N - 27 + 9 * a, N - 18 + 9 * a) for a in range(3)]]
gB = surfGB.from_parts(F, S, gl, indsLists, nil, eta=float(G["eta"]))
mu = float(G["fermi"])
NG = 96
P, dt = timed(lambda: quiet(de.densityGridN, F, S, gB, mu - 0.25, mu + 0.25, -1, NG, 0.0, False))
ogl = [O.surfGBAt(G["H"][i], G["Slist"][i], G["Vlist"][i], float(G["eta"])) for i in range(2)]
ogB = O.surfGB(F, S, ogl, indsLists, nil)
Pc, dtc = timed(lambda: O.densityGridN(F, S, ogB, mu - 0.25, mu + 0.25, -1, 2, 0.0), 1)
P2 = quiet(de.densityGridN, F, S, gB, mu - 0.25, mu + 0.25, -1, 2, 0.0, False)
record("cfg5_bethe_densityGridN_N2048", energies=NG, gpu_s=dt, gpu_pts_per_s=NG / dt, cpu_pts_per_s=2 / dtc,
       cpu_sample="densityGridN with N = 2 points", max_rel_diff_2pt=rel(P2, Pc))

json.dump(out, open(os.path.join("gpurun_out", "configs.json"), "w"), indent=1)
