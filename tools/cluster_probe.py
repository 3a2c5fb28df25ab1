"""2-CTA-cluster register inverse vs block engine at n = 128 / 112 (device-resident timing through inverse_batch)"""
import json
import sys
import time
import numpy as np
sys.path.insert(0, ".")
from gaunegf_b200._native import Context
import torch
ctx = Context(0)
rng = np.random.default_rng(0)
out = []
for n in (112, 128):
    for M in (16, 64, 148, 512, 2048):
        A = rng.standard_normal((M, n, n)) + 1j * rng.standard_normal((M, n, n))
        row = {"n": n, "M": M}
        for cl in (1, 0):
            ctx.lib.gnb_dev_set_option(b"small_cluster", cl)
            ctx.inverse_batch(A)
            ts = []
            for _ in range(5):
                torch.cuda.synchronize()
                t = time.perf_counter()
                ctx.inverse_batch(A)
                ts.append(time.perf_counter() - t)
            row["cluster_ms" if cl else "block_ms"] = 1e3 * float(np.median(ts))
        ctx.lib.gnb_dev_set_option(b"small_cluster", 1)
        out.append(row)
        print(json.dumps(row))
json.dump(out, open("gpurun_out/cluster_probe.json", "w"), indent=1)
