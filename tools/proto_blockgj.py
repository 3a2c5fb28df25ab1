"""numpy prototype of the batched block elimination engine (tournament pivoting, nb=32):
validates the algorithm the CUDA kernels in gaunegf_b200/csrc implement, step for step.
Dev tool only (not imported by the package or the tests)."""
import numpy as np

NB = 32
GROUP = 256


def cabs1(z):
    return np.abs(z.real) + np.abs(z.imag)


def gepp_select(rows_vals, want):
    """GEPP without physical swaps on a (n x w) block; returns order of chosen local rows,
    and the compact LU (want x w) in pivot order."""
    a = rows_vals.copy()
    n, w = a.shape
    alive = np.ones(n, bool)
    order, LU = [], np.zeros((min(want, n), w), complex)
    for j in range(min(want, n, w)):
        m = np.where(alive, cabs1(a[:, j]), -1.0)
        p = int(np.argmax(m))
        order.append(p)
        alive[p] = False
        LU[j] = a[p]
        piv = a[p, j]
        l = a[:, j] / piv
        l[~alive] = 0
        a[alive, j] = l[alive]
        a[:, j + 1:] -= np.outer(l, a[p, j + 1:])
    return order, LU


def tournament(A, r0, c0, w):
    n = A.shape[0]
    cand = np.arange(r0, n)
    while True:
        groups = [cand[i:i + GROUP] for i in range(0, len(cand), GROUP)]
        nxt, last = [], None
        for g in groups:
            order, LU = gepp_select(A[g, c0:c0 + w], w)
            nxt.extend(g[order])
            last = LU
        cand = np.array(nxt)
        if len(groups) == 1:
            return cand, last


def apply_perm(A, r0, chosen, perm):
    w = len(chosen)
    tgt = list(range(r0, r0 + w))
    src = list(chosen)
    vac = sorted(set(src) - set(tgt))
    dis = sorted(set(tgt) - set(src))
    new_rows = {t: s for t, s in zip(tgt, src)}
    new_rows.update({v: d for v, d in zip(vac, dis)})
    keys = list(new_rows)
    vals = [new_rows[k] for k in keys]
    A[keys] = A[vals]
    perm[keys] = perm[vals]


def lu_solve_compact(LU, B):
    w = LU.shape[0]
    X = B.astype(complex).copy()
    for i in range(w):
        X[i] -= LU[i, :i] @ X[:i]
    for i in range(w - 1, -1, -1):
        X[i] = (X[i] - LU[i, i + 1:w] @ X[i + 1:]) / LU[i, i]
    return X


def gj_inverse(A0):
    A = A0.astype(complex).copy()
    n = A.shape[0]
    perm = np.arange(n)
    for c0 in range(0, n, NB):
        w = min(NB, n - c0)
        chosen, LU = tournament(A, c0, c0, w)
        apply_perm(A, c0, chosen, perm)
        K = slice(c0, c0 + w)
        rowblk = A[K].copy()
        rowblk[:, K] = np.eye(w)
        W = lu_solve_compact(LU, rowblk)
        P = A[:, K].copy()
        P[K] = 0
        A[:, K] = 0
        A[K] = W
        A -= P @ W
    G = np.empty_like(A)
    G[:, perm] = A
    return G


def forward_solve(A0, B0):
    """block Gaussian elimination on [A|B] then unit-block-upper back substitution."""
    n = A0.shape[0]
    A = np.hstack([A0.astype(complex), B0.astype(complex)])
    perm = np.arange(n)
    for c0 in range(0, n, NB):
        w = min(NB, n - c0)
        chosen, LU = tournament(A, c0, c0, w)
        apply_perm(A, c0, chosen, perm)
        K = slice(c0, c0 + w)
        A[K, c0 + w:] = lu_solve_compact(LU, A[K, c0 + w:])
        A[c0 + w:, c0 + w:] -= A[c0 + w:, K] @ A[K, c0 + w:]
    X = A[:, n:].copy()
    nblk = (n + NB - 1) // NB
    for b in range(nblk - 1, 0, -1):
        c0 = b * NB
        K = slice(c0, min(c0 + NB, n))
        X[:c0] -= A[:c0, K] @ X[K]
    return X


if __name__ == "__main__":
    import sys
    sys.path.insert(0, ".")
    from gaunegf_b200 import synthetic as sy
    for N in (64, 100, 256, 600):
        F, S = sy.hermitian_pair(N, seed=0)
        s1, s2 = sy.block_sigma_vectors(N, max(N // 16, 2), 0.1)
        for E in (0.1 + 0j, -1 + 2j, 0.37 + 1e-6j):
            A = E * S - F - np.diag(s1 + s2)
            G0 = np.linalg.inv(A)
            G = gj_inverse(A)
            B = np.eye(N)[:, -8:]
            X = forward_solve(A, B)
            print(N, E, "cond %.1e" % np.linalg.cond(A),
                  "gj relerr %.1e" % (np.abs(G - G0).max() / np.abs(G0).max()),
                  "fwd relerr %.1e" % (np.abs(X - G0[:, -8:]).max() / np.abs(G0).max()))


# ---------------------------------------------------------------------------------------------
# Two-level variant (outer width 64 = two inner 32-blocks): near columns step by step, far columns
# with ONE rank-64 update per outer step.  Mirrors gnb_eliminate's two-level path kernel by kernel.
# ---------------------------------------------------------------------------------------------
def net_moves(r0, chosen):
    w = len(chosen)
    tgt = list(range(r0, r0 + w))
    src = list(chosen)
    vac = sorted(set(src) - set(tgt))
    dis = sorted(set(tgt) - set(src))
    return tgt + vac, src + dis


def move_rows(A, cols, dst, src):
    A[np.ix_(dst, cols)] = A[np.ix_(src, cols)]


def two_level_jordan(A0):
    A = A0.astype(complex).copy()
    n = A.shape[0]
    perm = np.arange(n)
    Pws = np.zeros((n, 64), complex)
    allc = np.arange(n)
    for c0 in range(0, n, 64):
        wa = min(32, n - c0)
        wb = min(32, n - c0 - 32) if n - c0 > 32 else 0
        near = np.arange(c0, c0 + wa + wb)
        far = np.setdiff1d(allc, near)
        Ka = np.arange(c0, c0 + wa)
        rows_not_a = np.setdiff1d(allc, Ka)
        chosen, LU = tournament(A, c0, c0, wa)
        inv_a = np.linalg.inv(A[np.ix_(chosen, Ka)])
        dst_a, src_a = net_moves(c0, chosen)
        perm[dst_a] = perm[src_a]
        move_rows(A, near, dst_a, src_a)
        rb = A[np.ix_(Ka, near)].copy(); rb[:, :wa] = np.eye(wa)
        A[np.ix_(Ka, near)] = inv_a @ rb
        Pws[:] = 0
        Pws[rows_not_a, :wa] = A[np.ix_(rows_not_a, Ka)]
        A[np.ix_(rows_not_a, Ka)] = 0
        A[np.ix_(rows_not_a, near)] -= Pws[rows_not_a, :wa] @ A[np.ix_(Ka, near)]
        if wb:
            Kb = np.arange(c0 + 32, c0 + 32 + wb)
            rows_not_b = np.setdiff1d(allc, Kb)
            chosen, LU = tournament(A, c0 + 32, c0 + 32, wb)
            inv_b = np.linalg.inv(A[np.ix_(chosen, Kb)])
            dst_b, src_b = net_moves(c0 + 32, chosen)
            perm[dst_b] = perm[src_b]
            move_rows(A, near, dst_b, src_b)
            Pws[dst_b, :32] = Pws[src_b, :32]
            rb = A[np.ix_(Kb, near)].copy(); rb[:, wa:wa + wb] = np.eye(wb)
            A[np.ix_(Kb, near)] = inv_b @ rb
            Pws[rows_not_b, 32:32 + wb] = A[np.ix_(rows_not_b, Kb)]
            Pws[Kb, 32:] = 0
            A[np.ix_(rows_not_b, Kb)] = 0
            A[np.ix_(rows_not_b, near)] -= Pws[rows_not_b, 32:32 + wb] @ A[np.ix_(Kb, near)]
        # far columns
        move_rows(A, far, dst_a, src_a)
        A[np.ix_(Ka, far)] = inv_a @ A[np.ix_(Ka, far)]
        if wb:
            move_rows(A, far, dst_b, src_b)
            R = A[np.ix_(Kb, far)] - Pws[Kb, :wa] @ A[np.ix_(Ka, far)]
            A[np.ix_(Kb, far)] = inv_b @ R
            upd = rows_not_b
            A[np.ix_(upd, far)] -= Pws[upd, :32 + wb] @ A[np.ix_(np.arange(c0, c0 + 32 + wb), far)]
        else:
            A[np.ix_(rows_not_a, far)] -= Pws[rows_not_a, :wa] @ A[np.ix_(Ka, far)]
    G = np.empty_like(A)
    G[:, perm] = A
    return G


def two_level_forward(A0, B0):
    n = A0.shape[0]
    A = np.hstack([A0.astype(complex), B0.astype(complex)])
    ncol = A.shape[1]
    for c0 in range(0, n, 64):
        wa = min(32, n - c0)
        wb = min(32, n - c0 - 32) if n - c0 > 32 else 0
        near = np.arange(c0, c0 + wa + wb)
        far = np.arange(c0 + wa + wb, ncol)
        Ka = np.arange(c0, c0 + wa)
        chosen, LU = tournament(A, c0, c0, wa)
        inv_a = np.linalg.inv(A[np.ix_(chosen, Ka)])
        dst_a, src_a = net_moves(c0, chosen)
        move_rows(A, near, dst_a, src_a)
        if wb:
            Kb = np.arange(c0 + 32, c0 + 32 + wb)
            A[np.ix_(Ka, Kb)] = inv_a @ A[np.ix_(Ka, Kb)]
            A[c0 + 32:n, c0 + 32:c0 + 32 + wb] -= A[c0 + 32:n, c0:c0 + 32] @ A[np.ix_(Ka, Kb)]
            chosen, LU = tournament(A, c0 + 32, c0 + 32, wb)
            inv_b = np.linalg.inv(A[np.ix_(chosen, Kb)])
            dst_b, src_b = net_moves(c0 + 32, chosen)
            move_rows(A, near, dst_b, src_b)
        move_rows(A, far, dst_a, src_a)
        A[np.ix_(Ka, far)] = inv_a @ A[np.ix_(Ka, far)]
        if wb:
            move_rows(A, far, dst_b, src_b)
            R = A[np.ix_(Kb, far)] - A[np.ix_(Kb, Ka)] @ A[np.ix_(Ka, far)]
            A[np.ix_(Kb, far)] = inv_b @ R
            lo = c0 + 32 + wb
            A[lo:n, lo:] -= A[lo:n, c0:lo] @ A[c0:lo, lo:]
        else:
            lo = c0 + wa
            A[lo:n, lo:] -= A[lo:n, c0:lo] @ A[c0:lo, lo:]
    X = A[:, n:].copy()
    nblk = (n + NB - 1) // NB
    for b in range(nblk - 1, 0, -1):
        c0 = b * NB
        K = slice(c0, min(c0 + NB, n))
        X[:c0] -= A[:c0, K] @ X[K]
    return X


if __name__ == "__main__":
    for N in (40, 64, 100, 200, 333):
        F, S = sy.hermitian_pair(N, seed=1)
        A = (0.2 + 0.01j) * S - F
        G0 = np.linalg.inv(A)
        G = two_level_jordan(A)
        X = two_level_forward(A, np.eye(N)[:, -8:])
        print("two-level", N, "jordan %.1e" % (np.abs(G - G0).max() / np.abs(G0).max()),
              "forward %.1e" % (np.abs(X - G0[:, -8:]).max() / np.abs(G0).max()))
