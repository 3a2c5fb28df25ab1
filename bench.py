#!/usr/bin/env python
"""bench.py — Green's-function energy points / second on B200 (BASELINE.json metric).

Workload (BASELINE.json configs[2], the N=1024 config the metric is quoted on): coherent transmission
T(E) of a synthetic N=1024 molecular junction (seeded Hermitian F/S, 64-orbital constant contacts),
10 000-point energy grid sharded over 8 GPUs = 1250 energy points per GPU per step (weak scaling:
per-GPU work fixed).  A "step" is one pass of the hot path over one such batch.

  python bench.py [--gpus N --steps K --warmup W]            our arm (one process per GPU)
  python bench.py --impl reference [...]                     the reference algorithm on the host CPU

Output: ONE JSON line (rank 0).
  value      whole-job energy points/s with F, S and the contact blocks resident in HBM
  e2e        the same metric through the public API (transport.calculate_transmission) with host buffers; F changes
             every step, so F, S travel host -> device inside the timed region every step
  roofline   the rank-K update kernel family (the dominant kernel) against the FP64 tensor-pipe ceiling MEASURED IN
             THIS RUN (gnb_dev_fp64_peak), plus step_frac = executed flops of all elimination launches / step time / peak
  parity     max relative deviation of the GPU results of THIS run (at this world size) from the CPU port on a sample
  secondary* density-matrix integration (GrInt, one NCCL all-reduce per call; cfg5 densityGridN at N=2048), complex-F
             T(E), N=512
  cpu_baseline  the numpy/LAPACK port of the reference (oracle/) on this box's host cores, bounded sample, two variants:
             as shipped (one process, all BLAS threads) and best effort (process pool, 1 BLAS thread per worker)
"""
import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ORB = 1024
N_CONTACT = 64
SEED = 1
E_PER_GPU = 1250          # 10 000 energies over 8 GPUs
CPU_SAMPLE = 8            # energies per as-shipped CPU step (about 1 s on 16 cores)
POOL_PER_WORKER = 2       # energies per worker per best-effort CPU step
METRIC = "Green's-function energy points/sec at N=1024"
UNIT = "energy points/s"


def workload_desc(n_gpus):
    return {
        "workload": f"cfg3 transmission T(E): synthetic N={N_ORB} junction (seed {SEED}), {N_CONTACT}-orbital constant "
                    f"contacts, {E_PER_GPU} energy points per GPU per step ({E_PER_GPU * n_gpus} total), grid [-0.5, 0.5] eV",
        "N": N_ORB, "contact_orbitals": N_CONTACT, "energies_per_gpu_per_step": E_PER_GPU,
        "parallelism": f"energy-grid sharding x{n_gpus} (no data-path collective; all-gather of T)",
        "l2": "per-step working set (1250 x 10.5 MB matrices) exceeds the 126 MB L2; no flush needed",
    }


def make_inputs():
    from gaunegf_b200 import synthetic as sy
    F, S = sy.hermitian_pair(N_ORB, seed=SEED)
    s1, s2 = sy.block_sigma_vectors(N_ORB, N_CONTACT, 0.1)
    return F, S, s1, s2


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_info():
    """CPU model and BLAS build the cpu_baseline numbers were measured with"""
    model, blas = "unknown", "unknown"
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip()
                break
    except Exception:
        pass
    try:
        cfg = np.show_config(mode="dicts")
        b = cfg.get("Build Dependencies", {}).get("blas", {})
        blas = f"{b.get('name', '?')} {b.get('version', '')}".strip()
    except Exception:
        pass
    return {"cpu_model": model, "blas": blas}


@contextlib.contextmanager
def all_blas_threads():
    """torchrun exports OMP_NUM_THREADS=1, which would starve the CPU arm: set the BLAS thread count explicitly"""
    n = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_limits
        with threadpool_limits(limits=n):
            yield n
    except ImportError:
        yield n


def cpu_reference_step(F, S, s1, s2, energies):
    from oracle import negf_oracle as O
    calc = O.SigmaCalculator(s1, s2, energy_dependent=False)
    t = time.perf_counter()
    T = O.calculate_transmission(F, S, calc, energies)
    return time.perf_counter() - t, T


def cpu_arms(F, S, s1, s2, steps, warmup, rng):
    """both CPU variants of the reference algorithm on bounded samples of the bench workload.
    Returns (as_shipped dict, best_effort dict, last as-shipped energies, their T)."""
    from oracle.cpu_pool import TransmissionPool
    cores = os.cpu_count() or 1
    with all_blas_threads():
        for _ in range(max(1, warmup)):
            cpu_reference_step(F, S, s1, s2, rng.uniform(-0.5, 0.5, 2))
        tot, Es, Tc = 0.0, None, None
        for _ in range(steps):
            Es = np.sort(rng.uniform(-0.5, 0.5, CPU_SAMPLE))
            dt, Tc = cpu_reference_step(F, S, s1, s2, Es)
            tot += dt
    shipped = {"value": CPU_SAMPLE * steps / tot, "unit": UNIT, "cores": cores, "threads": cores, "processes": 1,
               "ms_per_step": 1e3 * tot / steps,
               "sample": f"{CPU_SAMPLE} energies of the N={N_ORB} workload per step, one Python process, BLAS threads = {cores}"}
    pool = TransmissionPool(N_ORB, N_CONTACT, SEED, workers=cores)
    try:
        n_pool = POOL_PER_WORKER * cores
        for _ in range(max(1, min(warmup, 2))):
            pool.run(rng.uniform(-0.5, 0.5, cores))
        ptot = 0.0
        for _ in range(steps):
            dt, _ = pool.run(np.sort(rng.uniform(-0.5, 0.5, n_pool)))
            ptot += dt
    finally:
        pool.close()
    best = {"value": n_pool * steps / ptot, "unit": UNIT, "cores": cores, "threads": 1, "processes": cores,
            "ms_per_step": 1e3 * ptot / steps,
            "sample": f"{n_pool} energies per step over a pool of {cores} processes, 1 BLAS thread each (the reference's "
                      "own recipe: tests/benchmark_sigma_parallelization.py:27-30, 178-212)"}
    return shipped, best, Es, Tc


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    F, S, s1, s2 = make_inputs()
    rng = np.random.default_rng(0)
    shipped, best, _, _ = cpu_arms(F, S, s1, s2, args.steps, args.warmup, rng)
    top = best if best["value"] >= shipped["value"] else shipped
    val = top["value"]
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": top["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "complex128 (f64)", "data": "synthetic", "config": workload_desc(args.gpus),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": top["cores"], "kind": "port",
                             "sample": top["sample"] + "; the reference's algorithm (full solve(A, I) + Gamma1 G Gamma2 G^H, "
                                       "transport.py:150-157) restated in numpy/LAPACK (oracle/negf_oracle.py). value = the "
                                       "faster of the two variants below",
                             "variant": "best_effort" if top is best else "as_shipped",
                             "as_shipped": shipped, "best_effort": best, **host_info()},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


def cfg5_system():
    """BASELINE cfg 5: N = 2048 device with two 3-atom Bethe-lattice (Au) contacts; contact parameters as produced by the
    reference's own constructor (tests/golden/cfg5_full.npz, committed fixture)."""
    from gaunegf_b200 import synthetic as sy
    from gaunegf_b200.surfGBethe import surfGB, surfGBAt
    G = np.load(os.path.join(ROOT, "tests", "golden", "cfg5_full.npz"))
    N = int(G["N"])
    F, S = sy.hermitian_pair(N, seed=3)
    gl = [surfGBAt(G["H"][i], G["Slist"][i], G["Vlist"][i], float(G["eta"])) for i in range(2)]
    lens, flat = G["nInd_len"], list(G["nInd_flat"])
    nil, p = [], 0
    for c in lens:
        cl = []
        for n in c:
            cl.append([int(v) for v in flat[p:p + n]])
            p += n
        nil.append(cl)
    return G, F, S, surfGB.from_parts(F, S, gl, G["indsLists"], nil, eta=float(G["eta"]))


def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from gaunegf_b200 import transport as tr, integrate as it, density as de, parallel, synthetic as sy
    from gaunegf_b200._native import default_context

    F, S, s1, s2 = make_inputs()
    ctx = default_context(local)
    nc = N_CONTACT
    M_total = E_PER_GPU * world
    E_all = np.linspace(-0.5, 0.5, M_total)
    E_loc = np.ascontiguousarray(E_all[rank::world])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(steps):
            fn(k)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        barrier()
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def install_t1024(Fm=F):
        ctx.set_system(Fm, S)
        ctx.sigma_clear()
        ctx.sigma_add_const_block(np.arange(nc), np.diag(s1[:nc]))
        ctx.sigma_add_const_block(np.arange(N_ORB - nc, N_ORB), np.diag(s2[N_ORB - nc:]))

    # ---- device-resident leg: F, S, contact blocks already in HBM ---------------------------
    install_t1024()
    T_last = [None]

    def step_dev(_k):
        T_last[0] = ctx.transmission(E_loc, 0, -1)

    for k in range(args.warmup):
        step_dev(k)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    peak = ctx.fp64_peak(300.0)            # FP64 tensor-pipe ceiling of this GPU, now (clocks sampled alongside)
    l0 = ctx.launches
    ms = timed(step_dev, args.steps)
    launches = ctx.launches - l0
    elim_flops_step = ctx.last_elim_flops
    value = M_total * args.steps / (ms * 1e-3)
    # roofline leg: the same steps again with a CUDA-event pair around every launch of the rank-K update kernel
    # (on the stream the kernel is launched on); kept out of `value` because the per-launch events cost time.
    ctx.set_timing(True)
    ctx.gemm_stats(reset=True)
    roof_steps = max(1, min(args.steps, 3))
    ms_roof = timed(step_dev, roof_steps)
    gemm_ms, gemm_flops, gemm_n = ctx.gemm_stats(reset=True)
    ctx.set_timing(False)
    clocks = sampler.stop() if rank == 0 else None

    # ---- end-to-end leg: public API, host buffers, H2D + D2H inside the timed region ---------
    # F changes every step (an SCF loop hands a new Fock matrix to every call) and is uploaded every step; S is the same
    # array every step: the library finds it unchanged (memcmp against its pinned shadow) and keeps the resident copy
    Fv = []
    for d in (0.0, 1e-13):
        Fx = F.astype(np.complex128)
        Fx[0, 0] += d
        Fv.append(torch.from_numpy(Fx).pin_memory().numpy())
    Sp = torch.from_numpy(S.astype(np.complex128)).pin_memory().numpy()
    calc = tr.SigmaCalculator(s1, s2, energy_dependent=False)
    T_e2e = [None]

    def step_e2e(k):
        T_e2e[0] = tr.calculate_transmission(Fv[k % 2], Sp, calc, E_all)

    step_e2e(1)
    e2e_steps = max(2, min(args.steps, 6))
    skipped0 = ctx.system_uploads_skipped
    ms_e2e = timed(step_e2e, e2e_steps)
    assert ctx.system_uploads_skipped == skipped0, "e2e leg must upload F, S every step"
    e2e_val = M_total * e2e_steps / (ms_e2e * 1e-3)
    sent = bin(ctx.last_system_upload).count("1")       # matrices that went over PCIe in the last step (F changed, S did not)
    h2d = sent * N_ORB * N_ORB * 16 + 2 * (nc * nc * 16 + nc * 4) + E_loc.size * 16
    d2h = E_loc.size * 8
    step_e2e(0)
    ms_e2e_res = timed(lambda k: step_e2e(0), 3)          # unchanged F, S: they stay resident, nothing is re-sent
    T_full = T_e2e[0].copy()                              # T(E_all) with F = Fv[0], gathered from every rank

    # ---- secondary: density-matrix contour integration (GrInt, full G + on-device reduction + all-reduce)
    Mg = 296
    z, w = sy.contour_points(Mg * world, -30.0, 0.0)

    class ConstG:      # constant contacts through the surfG protocol -> compact device description
        indsList = [np.arange(nc), np.arange(N_ORB - nc, N_ORB)]
        def _gnb_install(self, c):
            c.sigma_add_const_block(np.arange(nc), np.diag(s1[:nc]))
            c.sigma_add_const_block(np.arange(N_ORB - nc, N_ORB), np.diag(s2[N_ORB - nc:]))
        def sigmaTot(self, E):
            return np.diag(s1 + s2)
        def sigma(self, E, i):
            return np.diag(s1 if i == 0 else s2)
    gobj = ConstG()
    P = [None]
    brk = []

    def step_grint(_k):
        P[0] = it.GrInt(Fv[0], Sp, gobj, z, w)
        if world > 1:
            brk.append(dict(parallel.last_breakdown))

    step_grint(0)
    brk.clear()
    g_steps = 4
    ms_g = timed(step_grint, g_steps)
    grint_val = Mg * world * g_steps / (ms_g * 1e-3)
    # parity of the sharded integral at this world size: 8 contour points split over the ranks vs the CPU port
    z8, w8 = sy.contour_points(8, -30.0, 0.0)
    P8 = it.GrInt(Fv[0], Sp, gobj, z8, w8)

    # ---- secondary: complex F (spin-orbit-like Hermitian F): no real-structure shortcut anywhere, device-resident
    Fc, _ = sy.hermitian_pair(N_ORB, seed=SEED, complex_F=True)
    install_t1024(Fc)

    def step_cplx(_k):
        T_last[0] = ctx.transmission(E_loc, 0, -1)

    step_cplx(0)
    c_steps = max(1, min(args.steps, 3))
    ms_c = timed(step_cplx, c_steps)
    cplx_val = M_total * c_steps / (ms_c * 1e-3)
    cplx_flops_step = ctx.last_elim_flops

    # ---- secondary: the same T(E) path at N = 512 (BASELINE metric names N = 512 / 1024), device-resident
    N5, nc5, M5 = 512, 32, 2500
    F5, S5 = sy.hermitian_pair(N5, seed=1)
    s15, s25 = sy.block_sigma_vectors(N5, nc5, 0.1)
    E5 = np.ascontiguousarray(np.linspace(-0.5, 0.5, M5 * world)[rank::world])
    ctx.set_system(F5, S5)
    ctx.sigma_clear()
    ctx.sigma_add_const_block(np.arange(nc5), np.diag(s15[:nc5]))
    ctx.sigma_add_const_block(np.arange(N5 - nc5, N5), np.diag(s25[N5 - nc5:]))

    def step_n512(_k):
        T_last[0] = ctx.transmission(E5, 0, -1)

    step_n512(0)
    n512_steps = max(1, min(args.steps, 5))
    ms_5 = timed(step_n512, n512_steps)
    n512_val = M5 * world * n512_steps / (ms_5 * 1e-3)

    # ---- secondary: BASELINE cfg 5, densityGridN at N = 2048 with Bethe-lattice contacts, sharded + one all-reduce
    G5, Fb, Sb, gB = cfg5_system()
    mu = float(G5["fermi"])
    NG = 96 * world

    def quiet(f, *a):
        with contextlib.redirect_stdout(io.StringIO()):
            return f(*a)

    def step_cfg5(_k):
        P[0] = quiet(de.densityGridN, Fb, Sb, gB, mu - 0.25, mu + 0.25, -1, NG, 0.0, False)

    step_cfg5(0)
    ms_b = timed(step_cfg5, 2)
    cfg5_val = NG * 2 / (ms_b * 1e-3)
    Pg8 = quiet(de.densityGridN, Fb, Sb, gB, mu - 0.25, mu + 0.25, -1, 8, 0.0, False)    # the golden's 8-point grid
    cfg5_par = max(rel(Pg8[G5["ii"], G5["jj"]], G5["PgN_samp"]), rel(np.diag(Pg8), G5["PgN_diag"]))

    if rank == 0:
        achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
        traffic, traffic_src, traffic_alg = None, None, None
        for name in ("r02_rk_gemm_ncu.json", "r01c_rk_gemm_ncu.json"):
            try:
                prof = json.load(open(os.path.join(ROOT, "profiles", name)))
                traffic = prof.get("dram_bytes_per_launch")
                traffic_alg = prof.get("algorithmic_bytes_per_launch")
                traffic_src = (f"profiles/{name}: dram__bytes_read.sum + dram__bytes_write.sum of the LARGEST launch of the family "
                               "(top-level far update, K = 512, 625 matrices) from an ncu --set full capture; static, not "
                               "re-measured in this run; traffic_algorithmic = C tile read + written once + the packed operands once")
                break
            except Exception:
                pass
        W_T = 8 / 3 * N_ORB ** 3 + 8 * N_ORB ** 2 * nc
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "complex128 (f64)", "data": "synthetic", "config": workload_desc(world),
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "gaunegf_b200.transport.calculate_transmission (pinned numpy in, numpy out); F differs from the "
                           "previous call every step and is uploaded every step; S repeats (as in an SCF loop) and stays resident after "
                           "the library's comparison with its pinned shadow copy (h2d_bytes_per_step counts what was sent)",
                    "ms_per_step": ms_e2e / e2e_steps,
                    "resident_value": M_total * 3 / (ms_e2e_res * 1e-3),
                    "resident_note": "same call with F, S unchanged since the previous call: they stay in HBM (h2d = energies only)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor",
                         "kernel": "k_rk_gemm / k_rk_gemm_rp (rank-K update C -= P W of the complex128 elimination, K = 32..512, packed "
                                   "operands by cp.async.bulk, DMMA.8x8x4; _rp = real-packed operands of the real columns). "
                                   "achieved counts EXECUTED arithmetic in 4-multiplication-equivalent flops: 8 per complex "
                                   "MAC, 4 where the panel is real, 2 where panel and pivot rows are real (real F, S, E with "
                                   "the contact orbitals ordered last keep every column left of the contacts exactly real); "
                                   "complex tiles with K >= 64 use 3M arithmetic (3 DMMAs for an 8-flop MAC)",
                         "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": (achieved / peak) if achieved else None,
                         "traffic": traffic, "traffic_algorithmic": traffic_alg, "traffic_source": traffic_src,
                         "peak_source": "measured in this run: gnb_dev_fp64_peak (register-resident DMMA.8x8x4 issue loop on every "
                                        "SM, best of ~0.3 s of launches), clocks as in `clocks`",
                         "launches_timed": int(gemm_n),
                         "algorithmic_flops_per_launch_avg": gemm_flops / gemm_n if gemm_n else None,
                         "kernel_share_of_step": gemm_ms / ms_roof if ms_roof > 0 else None,
                         "roofline_leg_ms_per_step": ms_roof / roof_steps,
                         "step_frac": elim_flops_step / (ms / args.steps * 1e-3) / 1e12 / peak,
                         "step_executed_tflops": elim_flops_step / (ms / args.steps * 1e-3) / 1e12,
                         "step_frac_note": "executed flops of ALL rank-K, forward-W and back-substitution launches of one step "
                                           "(host-counted, gnb_last_elim_flops) / device step time / peak; tournament pivoting "
                                           "(< 1 % of the flops, mostly FP32) is not counted",
                         "step_algorithmic_tflops": W_T * E_loc.size * args.steps / (ms * 1e-3) / 1e12,
                         "step_algorithmic_note": "complex128 flop count of the contact-column algorithm (SURVEY 8d: 8/3 N^3 "
                                                  "+ 8 N^2 n2 per energy) - NOT a roofline fraction: real columns run real "
                                                  "arithmetic, so fewer flops are executed"},
            "secondary": {"what": f"integrate.GrInt contour integration N={N_ORB}: full G(E) per point + on-device weighted "
                                  f"reduction + one NCCL all-reduce per call, {Mg} points per GPU per call (public API, host buffers, "
                                  "F and S resident)",
                          "value": grint_val, "unit": UNIT, "ms_per_call": ms_g / g_steps,
                          "algorithmic_tflops_per_gpu": 8 * N_ORB ** 3 * grint_val / world / 1e12,
                          "frac_of_fp64_peak": 8 * N_ORB ** 3 * grint_val / world / 1e12 / peak,
                          "frac_note": "algorithmic 8 N^3 per point / measured peak; the update runs 3M arithmetic (6 executed flops "
                                       "per 8 algorithmic), so this can approach 1.33",
                          "rank0_breakdown_ms": ({k: float(np.mean([b[k] for b in brk])) for k in ("partial_ms", "allreduce_ms", "d2h_ms")}
                                                 if brk else None)},
            "secondary_complex_F": {"what": f"T(E) at N={N_ORB} with a complex Hermitian F (no real-structure shortcut), "
                                            f"{E_PER_GPU} energy points per GPU per step, device-resident",
                                    "value": cplx_val, "unit": UNIT, "ms_per_step": ms_c / c_steps,
                                    "step_frac": cplx_flops_step / (ms_c / c_steps * 1e-3) / 1e12 / peak,
                                    "step_algorithmic_tflops": W_T * cplx_val / world / 1e12},
            "secondary_n512": {"what": f"T(E) at N={N5}, {nc5}-orbital constant contacts, {M5} energy points per GPU per step, "
                                       f"device-resident (same path and kernels as the headline value)",
                               "value": n512_val, "unit": UNIT,
                               "step_algorithmic_tflops": (8 / 3 * N5 ** 3 + 8 * N5 ** 2 * nc5) * n512_val / world / 1e12},
            "secondary_cfg5": {"what": "BASELINE cfg 5: density.densityGridN (GrLessInt, ind=-1) at N=2048 with two 27-orbital "
                                       "Bethe-lattice contacts, 96 grid points per GPU per call, one NCCL all-reduce per call "
                                       "(public API, host buffers)",
                               "value": cfg5_val, "unit": UNIT, "ms_per_call": ms_b / 2,
                               "algorithmic_tflops_per_gpu": (8 / 3 * 2048 ** 3 + 16 * 2048 ** 2 * 54) * cfg5_val / world / 1e12,
                               "parity_vs_reference_golden": cfg5_par,
                               "parity_note": "8-point densityGridN computed at THIS world size vs the unmodified reference's "
                                              "result (tests/golden/cfg5_full.npz: 512 sampled entries + diagonal)"},
        }
        # ---- parity at this world size + CPU baseline -------------------------------------------------
        from oracle import negf_oracle as O
        rng = np.random.default_rng(0)
        idx = np.sort(rng.choice(M_total, CPU_SAMPLE, replace=False))      # spread over every rank's shard
        with all_blas_threads():
            _, Tc = cpu_reference_step(F, S, s1, s2, E_all[idx])
            P8c = O.GrInt(F, S, _ConstOracle(s1, s2), z8, w8)
        line["parity"] = {"T_max_rel_diff_vs_cpu": rel(T_full[idx], Tc),
                          "T_sample": f"{CPU_SAMPLE} random energies of the {M_total}-point grid (shards of all {world} ranks), "
                                      "GPU values from the e2e call of this run",
                          "GrInt_max_rel_diff_vs_cpu": rel(P8, P8c),
                          "GrInt_sample": f"8-point contour integral sharded over {world} rank(s) vs numpy/LAPACK",
                          "cfg5_max_rel_diff_vs_reference": cfg5_par, "tolerance": 1e-10}
        if world == 1 and not args.no_cpu:
            shipped, best, _, _ = cpu_arms(F, S, s1, s2, 2, 1, rng)
            top = best if best["value"] >= shipped["value"] else shipped
            line["cpu_baseline"] = {"value": top["value"], "unit": UNIT, "cores": top["cores"], "kind": "port",
                                    "sample": top["sample"] + " (numpy/LAPACK port of the reference algorithm, "
                                                              "oracle/negf_oracle.py); value = the faster variant",
                                    "variant": "best_effort" if top is best else "as_shipped",
                                    "as_shipped": shipped, "best_effort": best,
                                    "max_rel_diff_vs_gpu": line["parity"]["T_max_rel_diff_vs_cpu"], **host_info()}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


class _ConstOracle:
    """constant diagonal contacts for the oracle's GrInt (surfG protocol)"""

    def __init__(self, s1, s2):
        self.tot = np.diag(s1 + s2)

    def sigmaTot(self, E):
        return self.tot


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline timing legs (parity is still checked)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
