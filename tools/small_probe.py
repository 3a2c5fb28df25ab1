"""Times the one-CTA-per-energy shared-memory path against the lock-step block engine (dev switch small_fused)."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from gaunegf_b200 import synthetic as sy          # noqa: E402
from gaunegf_b200._native import Context          # noqa: E402
import torch                                       # noqa: E402

ctx = Context(0)
out = []


def timed(f, reps=5):
    f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        f()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t)
    return float(np.median(ts))


CASES = ((64, 1, 1000), (64, 1, 20000), (32, 4, 20000), (96, 16, 20000), (119, 32, 20000), (64, 1, 12), (96, 16, 12), (119, 32, 12), (119, 32, 108))
if len(sys.argv) > 1:
    CASES = CASES[:int(sys.argv[1])]
for N, nc, M in CASES:
    if nc == 1:
        F, S, s1, s2 = sy.chain(N)
        blocks = [([0], [[s1[0]]]), ([N - 1], [[s2[N - 1]]])]
    else:
        F, S = sy.hermitian_pair(N, seed=N)
        inds = sy.end_contacts(N, nc)
        blocks = [(i, -0.1j * np.eye(nc)) for i in inds]
    ctx.set_system(F, S)
    ctx.sigma_clear()
    for i, b in blocks:
        ctx.sigma_add_const_block(i, b)
    E = np.linspace(-3, 3, M)
    z = E + 0.05j
    w = np.full(M, 1.0 / M, dtype=complex)
    row = {"N": N, "nc": nc, "M": M}
    for on, reg, tag in ((1, 1, "reg"), (1, 0, "smem"), (0, 0, "block")):
        if tag == "reg" and N > 96:
            continue
        ctx.lib.gnb_dev_set_option(b"small_fused", on)
        ctx.lib.gnb_dev_set_option(b"small_reg", reg)
        row[f"T_{tag}_pts_per_s"] = M / timed(lambda: ctx.transmission(E, 0, -1))
        row[f"DOS_{tag}_pts_per_s"] = M / timed(lambda: ctx.dos(E))
        row[f"GrInt_{tag}_pts_per_s"] = M / timed(lambda: ctx.gr_int(z, w))
    ctx.lib.gnb_dev_set_option(b"small_fused", 1)
    ctx.lib.gnb_dev_set_option(b"small_reg", 1)
    row["T_flops_per_pt_gj"] = 8.0 * N ** 3
    best = "reg" if N <= 96 else "smem"
    row["T_%s_tflops" % best] = row["T_%s_pts_per_s" % best] * 8.0 * N ** 3 / 1e12
    out.append(row)
    print(json.dumps(row))
json.dump(out, open("gpurun_out/small_probe.json", "w"), indent=1)
