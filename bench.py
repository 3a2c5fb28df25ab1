#!/usr/bin/env python
"""bench.py — Green's-function energy points / second on B200 (BASELINE.json metric).

Workload (BASELINE.json configs[2], the N=1024 config the metric is quoted on): coherent transmission
T(E) of a synthetic N=1024 molecular junction (seeded Hermitian F/S, 64-orbital constant contacts),
10 000-point energy grid sharded over 8 GPUs = 1250 energy points per GPU per step (weak scaling:
per-GPU work fixed).  A "step" is one pass of the hot path over one such batch.

  python bench.py [--gpus N --steps K --warmup W]            our arm (one process per GPU)
  python bench.py --impl reference [...]                     the reference algorithm on the host CPU

Output: ONE JSON line (rank 0).  `value` = whole-job energy points/s with F, S and the contact blocks
resident in HBM; `e2e` = the same metric through the public API (transport.calculate_transmission)
with host buffers, H2D/D2H inside the timed region; `roofline` = the rank-K update kernel (the
dominant kernel) against the measured FP64 tensor-pipe peak; `cpu_baseline` = the numpy/LAPACK
port of the reference (oracle/) on this box's host cores, bounded sample.
"""
import argparse
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
E_PER_GPU = 1250          # 10 000 energies over 8 GPUs
CPU_SAMPLE = 8            # energies per CPU-reference step (about 0.5 s each on 16 cores)
METRIC = "Green's-function energy points/sec at N=1024"
UNIT = "energy points/s"


def workload_desc(n_gpus):
    return {
        "workload": f"cfg3 transmission T(E): synthetic N={N_ORB} junction (seed 1), {N_CONTACT}-orbital constant "
                    f"contacts, {E_PER_GPU} energy points per GPU per step ({E_PER_GPU * n_gpus} total), grid [-0.5, 0.5] eV",
        "N": N_ORB, "contact_orbitals": N_CONTACT, "energies_per_gpu_per_step": E_PER_GPU,
        "parallelism": f"energy-grid sharding x{n_gpus} (no data-path collective; all-gather of T)",
        "l2": "per-step working set (1250 x 16.8 MB matrices) exceeds the 126 MB L2; no flush needed",
    }


def make_inputs():
    from gaunegf_b200 import synthetic as sy
    F, S = sy.hermitian_pair(N_ORB, seed=1)
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


def fp64_peak():
    """FP64 tensor-pipe peak: MEASURED_PEAKS.json holds only bf16/HBM, so the denominator is the
    DMMA.8x8x4 issue-rate ceiling measured on this pool's B200 by tools/fp64_peak.cu
    (profiles/r01_fp64_peak_dmma_dfma.json)."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r01_fp64_peak_dmma_dfma.json")))
        return max(v for k, v in d.items() if k.startswith("dmma884")), "measured DMMA.8x8x4 ceiling (tools/fp64_peak.cu, this pool)"
    except Exception:
        return 37.0, "fallback: B200 nominal FP64 37 TFLOP/s"


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
    return {"cpu_model": model, "blas": blas, "threads": "BLAS default = all host cores, one Python process"}


def cpu_reference_step(F, S, s1, s2, energies):
    from oracle import negf_oracle as O
    calc = O.SigmaCalculator(s1, s2, energy_dependent=False)
    t = time.perf_counter()
    T = O.calculate_transmission(F, S, calc, energies)
    return time.perf_counter() - t, T


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    F, S, s1, s2 = make_inputs()
    rng = np.random.default_rng(0)
    for _ in range(args.warmup):
        cpu_reference_step(F, S, s1, s2, rng.uniform(-0.5, 0.5, 2))
    tot = 0.0
    for _ in range(args.steps):
        dt, _ = cpu_reference_step(F, S, s1, s2, np.sort(rng.uniform(-0.5, 0.5, CPU_SAMPLE)))
        tot += dt
    val = CPU_SAMPLE * args.steps / tot
    cores = os.cpu_count()
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "complex128 (f64)", "data": "synthetic", "config": workload_desc(args.gpus),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{CPU_SAMPLE} energies of the same N={N_ORB} workload per step: the reference's algorithm "
                                       "(full solve(A, I) + Gamma1 G Gamma2 G^H, transport.py:150-157) restated in numpy/LAPACK "
                                       "(oracle/negf_oracle.py), BLAS threads = all host cores", **host_info()},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from gaunegf_b200 import transport as tr
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
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        barrier()
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- device-resident leg: F, S, contact blocks already in HBM ---------------------------
    ctx.set_system(F, S)
    ctx.sigma_clear()
    ctx.sigma_add_const_block(np.arange(nc), np.diag(s1[:nc]))
    ctx.sigma_add_const_block(np.arange(N_ORB - nc, N_ORB), np.diag(s2[N_ORB - nc:]))
    T_last = [None]

    def step_dev():
        T_last[0] = ctx.transmission(E_loc, 0, -1)

    for _ in range(args.warmup):
        step_dev()
    l0 = ctx.launches
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(step_dev, args.steps)
    launches = ctx.launches - l0
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
    Fp = torch.from_numpy(F.astype(np.complex128)).pin_memory().numpy()
    Sp = torch.from_numpy(S.astype(np.complex128)).pin_memory().numpy()
    calc = tr.SigmaCalculator(s1, s2, energy_dependent=False)
    T_e2e = [None]

    def step_e2e():
        T_e2e[0] = tr.calculate_transmission(Fp, Sp, calc, E_all)

    step_e2e()
    e2e_steps = max(1, min(args.steps, 5))
    ms_e2e = timed(step_e2e, e2e_steps)
    e2e_val = M_total * e2e_steps / (ms_e2e * 1e-3)
    h2d = 2 * N_ORB * N_ORB * 16 + 2 * (nc * nc * 16 + nc * 4) + E_loc.size * 16
    d2h = E_loc.size * 8

    # ---- secondary: density-matrix contour integration (GrInt, full G + on-device reduction + all-reduce)
    from gaunegf_b200 import integrate as it, synthetic as sy
    from gaunegf_b200.sigma_plan import ArrayPlan
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

    def step_grint():
        P[0] = it.GrInt(Fp, Sp, gobj, z, w)

    step_grint()
    ms_g = timed(step_grint, 2)
    grint_val = Mg * world * 2 / (ms_g * 1e-3)

    # ---- secondary: the same T(E) path at N = 512 (BASELINE metric names N = 512 / 1024), device-resident
    N5, nc5, M5 = 512, 32, 2500
    F5, S5 = sy.hermitian_pair(N5, seed=1)
    s15, s25 = sy.block_sigma_vectors(N5, nc5, 0.1)
    E5 = np.ascontiguousarray(np.linspace(-0.5, 0.5, M5 * world)[rank::world])
    ctx.set_system(F5, S5)
    ctx.sigma_clear()
    ctx.sigma_add_const_block(np.arange(nc5), np.diag(s15[:nc5]))
    ctx.sigma_add_const_block(np.arange(N5 - nc5, N5), np.diag(s25[N5 - nc5:]))

    def step_n512():
        T_last[0] = ctx.transmission(E5, 0, -1)

    step_n512()
    n512_steps = max(1, min(args.steps, 5))
    ms_5 = timed(step_n512, n512_steps)
    n512_val = M5 * world * n512_steps / (ms_5 * 1e-3)
    ctx.set_system(F, S)                       # back to the N = 1024 system (the CPU leg below compares against it)
    ctx.sigma_clear()
    ctx.sigma_add_const_block(np.arange(nc), np.diag(s1[:nc]))
    ctx.sigma_add_const_block(np.arange(N_ORB - nc, N_ORB), np.diag(s2[N_ORB - nc:]))

    if rank == 0:
        peak, peak_src = fp64_peak()
        achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r01c_rk_gemm_ncu.json"))).get("dram_bytes_per_launch")
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "complex128 (f64)", "data": "synthetic", "config": workload_desc(world),
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "gaunegf_b200.transport.calculate_transmission (pinned numpy in, numpy out)",
                    "ms_per_step": ms_e2e / e2e_steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor",
                         "kernel": "k_rk_gemm / k_rk_gemm_rp (rank-K update C -= P W of the complex128 elimination, K = 32..512, packed "
                                   "operands by cp.async.bulk, DMMA.8x8x4; _rp = real-packed operands of the real columns). "
                                   "achieved counts EXECUTED arithmetic in 4-multiplication-equivalent flops: 8 per complex "
                                   "MAC, 4 where the panel is real, 2 where panel and pivot rows are real (real F, S, E with "
                                   "the contact orbitals ordered last keep every column left of the contacts exactly real); "
                                   "complex tiles with K >= 64 use 3M arithmetic (3 DMMAs for an 8-flop MAC), so the rate can "
                                   "exceed the DMMA ceiling",
                         "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": (achieved / peak) if achieved else None,
                         "traffic": traffic, "peak_source": peak_src, "launches_timed": int(gemm_n),
                         "algorithmic_flops_per_launch_avg": gemm_flops / gemm_n if gemm_n else None,
                         "kernel_share_of_step": gemm_ms / ms_roof if ms_roof > 0 else None,
                         "roofline_leg_ms_per_step": ms_roof / roof_steps,
                         "step_algorithmic_tflops": (8 / 3 * N_ORB ** 3 + 8 * N_ORB ** 2 * nc) * E_loc.size * args.steps
                                                    / (ms * 1e-3) / 1e12,
                         "step_algorithmic_note": "complex128 flop count of the contact-column algorithm (SURVEY 8d: 8/3 N^3 "
                                                  "+ 8 N^2 n2 per energy); the executed count is lower because real columns "
                                                  "are updated with real arithmetic"},
            "secondary": {"what": f"integrate.GrInt contour integration N={N_ORB}: full G(E) per point + on-device weighted "
                                  f"reduction + one NCCL all-reduce per call, {Mg} points per GPU per call (public API, host buffers)",
                          "value": grint_val, "unit": UNIT,
                          "algorithmic_tflops_per_gpu": 8 * N_ORB ** 3 * grint_val / world / 1e12},
            "secondary_n512": {"what": f"T(E) at N={N5}, {nc5}-orbital constant contacts, {M5} energy points per GPU per step, "
                                       f"device-resident (same path and kernels as the headline value)",
                               "value": n512_val, "unit": UNIT,
                               "step_algorithmic_tflops": (8 / 3 * N5 ** 3 + 8 * N5 ** 2 * nc5) * n512_val / 1e12},
        }
        if world == 1 and not args.no_cpu:
            rng = np.random.default_rng(0)
            Es = np.sort(rng.uniform(-0.5, 0.5, CPU_SAMPLE))
            cpu_reference_step(F, S, s1, s2, Es[:2])
            dt, Tc = cpu_reference_step(F, S, s1, s2, Es)
            Tg = ctx.transmission(Es, 0, -1)
            line["cpu_baseline"] = {"value": CPU_SAMPLE / dt, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"{CPU_SAMPLE} energies of the same workload (numpy/LAPACK port of the reference "
                                              f"algorithm, oracle/negf_oracle.py), BLAS threads = all host cores",
                                    "max_rel_diff_vs_gpu": float(np.max(np.abs(Tg - Tc)) / np.max(np.abs(Tc))), **host_info()}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
