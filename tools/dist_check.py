"""Multi-GPU parity check (run under torchrun, one rank per GPU, NCCL): the energy-sharded drivers must return the
same results as the numpy oracle on EVERY rank.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/dist_check.py"""
import contextlib
import io
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from gaunegf_b200 import synthetic as sy, transport as tr, integrate as it, density as de    # noqa: E402
from gaunegf_b200.surfGTester import surfGTest                                                  # noqa: E402
from oracle import negf_oracle as O                                                             # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / np.max(np.abs(b)))


N, nc = 200, 12
F, S = sy.hermitian_pair(N, seed=9)
inds = sy.end_contacts(N, nc)
g, og = surfGTest(F, S, inds, -0.1j, -0.2j), O.surfGTest(F, S, inds, -0.1j, -0.2j)
s1, s2 = sy.block_sigma_vectors(N, nc, 0.1)
E = np.linspace(-1, 1, 101)                      # not a multiple of the world size: ragged shards
z, w = sy.contour_points(37, -8.0, 0.1)
errs = {
    "T": rel(tr.calculate_transmission(F, S, tr.SigmaCalculator(s1, s2), E),
             O.calculate_transmission(F, S, O.SigmaCalculator(s1, s2), E)),
    "dos": rel(tr.calculate_dos(F, S, tr.SigmaCalculator(s1, s2), E)[1], O.calculate_dos(F, S, O.SigmaCalculator(s1, s2), E)[1]),
    "GrInt": rel(it.GrInt(F, S, g, z, w), O.GrInt(F, S, og, z, w)),
    "GrLessInt": rel(it.GrLessInt(F, S, g, E[:33], np.full(33, 0.03), -1), O.GrLessInt(F, S, og, E[:33], np.full(33, 0.03), -1)),
    "GrInt_fewer_points_than_ranks": rel(it.GrInt(F, S, g, z[:3], w[:3]), O.GrInt(F, S, og, z[:3], w[:3])),
}
with contextlib.redirect_stdout(io.StringIO()):
    P = de.densityComplexN(F, S, g, -8.0, 0.0, 36, 0.0, False)
errs["densityComplexN"] = rel(P, O.densityComplexN(F, S, og, -8.0, 0.0, 36, 0.0))
# adaptive contour integration: several nested levels per batch, every level sharded on its own (sharded_matrix_sums)
with contextlib.redirect_stdout(io.StringIO()):
    Pa = de.densityComplex(F, S, g, -8.0, 0.0, 1e-6, 0.0)
    Pa_ref = O.densityComplex(F, S, og, -8.0, 0.0, 1e-6, 0.0)
errs["densityComplex_adaptive"] = rel(Pa, Pa_ref)
lv = it.GrIntLevels(F, S, g, [(z[:2], w[:2]), (z[2:6], w[2:6]), (z[6:6], w[6:6]), (z[6:], w[6:])])
errs["GrIntLevels"] = max(rel(lv[0], O.GrInt(F, S, og, z[:2], w[:2])), rel(lv[1], O.GrInt(F, S, og, z[2:6], w[2:6])),
                          float(np.max(np.abs(lv[2]))), rel(lv[3], O.GrInt(F, S, og, z[6:], w[6:])))
# the same F, S again: the ranks agree that nothing changed; a changed F is taken over by every rank
F2 = F.copy(); F2[3, 3] += 0.01
errs["GrInt_changed_F"] = rel(it.GrInt(F2, S, g, z, w), O.GrInt(F2, S, og, z, w))
worst = torch.tensor([max(errs.values())], device="cuda", dtype=torch.float64)
dist.all_reduce(worst, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"world={world}", {k: f"{v:.1e}" for k, v in errs.items()}, "max over ranks", f"{worst.item():.1e}")
assert worst.item() < 1e-10, errs
dist.destroy_process_group()
