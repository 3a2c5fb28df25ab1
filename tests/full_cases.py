"""Shared builders for the BASELINE-size golden cases (tests/golden/make_golden_full.py): the same seeded inputs
rebuilt on the GPU box / for the oracle, with the Bethe contact parts the reference's constructor produced."""
import numpy as np

from gaunegf_b200 import synthetic as sy


def nind_lists(G, prefix=""):
    lens, flat = G[prefix + "nInd_len"], list(G[prefix + "nInd_flat"])
    nil, p = [], 0
    for c in lens:
        cl = []
        for n in c:
            cl.append([int(v) for v in flat[p:p + n]])
            p += n
        nil.append(cl)
    return nil


def bethe_atoms(G, cls, prefix="", eta=1e-4):
    return [cls(G[prefix + "H"][i], G[prefix + "Slist"][i], G[prefix + "Vlist"][i], eta) for i in range(2)]


def spin_system(F, S, sp, seed=17):
    """same construction as make_golden_full.spin_system"""
    F2 = np.kron(np.eye(2), F) if sp == "u" else np.kron(F, np.eye(2))
    S2 = np.kron(np.eye(2), S) if sp == "u" else np.kron(S, np.eye(2))
    D = np.random.default_rng(seed).standard_normal(F2.shape) * 0.02
    return F2 + (D + D.T) / 2, S2


def check_sampled(P, G, key, tol, relerr):
    """sampled entries, diagonal and Frobenius norm of an N x N result against the golden summaries"""
    ii, jj = G["ii"], G["jj"]
    scale = np.max(np.abs(G[key + "_diag"]))
    assert np.max(np.abs(P[ii, jj] - G[key + "_samp"])) / max(scale, np.max(np.abs(G[key + "_samp"]))) < tol, key
    assert relerr(np.diag(P), G[key + "_diag"]) < tol, key
    assert abs(np.linalg.norm(P) - float(G[key + "_fro"])) < tol * float(G[key + "_fro"]), key


def cfg4_system():
    return sy.lead_device_lead(128, 512, seed=2, s_off=0.0)
