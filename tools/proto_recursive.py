"""numpy model of the recursive (multi-level) block elimination that gnb_elim.cu implements.

Every helper below corresponds to one CUDA kernel family and touches exactly the buffers that kernel
touches, so data-flow hazards (which buffer holds forward / final values at which time) are validated here:

  A    : the matrices, [N x (N + naug)], updated in place
  Pbuf : saved panel columns  (GEMM left operand; "Ppk" on the device, packed)
  Lbuf : JORDAN only - row blocks L_ab (b < a) extracted from Pbuf when block a's pivot rows are known
  Wbuf : normalised pivot rows, forward values (GEMM right operand; "Wpk" on the device)

Dev tool only (not imported by the package or the tests)."""
import sys
import numpy as np

sys.path.insert(0, ".")
from tools.proto_blockgj import tournament, net_moves  # noqa: E402

NB = 32
STATS = {}


def gemm(C, rows, cols, P, prow, pcols, W, wrows, wcols, tag):
    """C[rows, cols] -= P[prow, pcols] @ W[wrows, wcols]   (the rank-K update kernel)"""
    k = len(pcols)
    if len(rows) == 0 or len(cols) == 0 or k == 0:
        return
    C[np.ix_(rows, cols)] -= P[np.ix_(prow, pcols)] @ W[np.ix_(wrows, wcols)]
    key = (tag, k)
    STATS[key] = STATS.get(key, 0) + 8.0 * len(rows) * len(cols) * k


class Elim:
    def __init__(self, A0, naug_cols=None, jordan=False, leaf=32):
        n = A0.shape[0]
        assert n % NB == 0
        self.n = n
        self.jordan = jordan
        B = np.zeros((n, 0), complex) if naug_cols is None else naug_cols
        self.A = np.hstack([A0.astype(complex), B.astype(complex)])
        self.ncol = self.A.shape[1]
        self.P = np.zeros((n, n), complex)
        self.L = np.zeros((n, n), complex)
        self.W = np.zeros((n, self.ncol), complex)
        self.perm = np.arange(n)
        self.moves = {}     # block index -> (dst, src)
        self.inv = {}       # block index -> inverse of the pivot block
        self.top = 0        # first column of the outermost range whose panels are still live

    # ---- kernels ---------------------------------------------------------------------------
    def k_tournament(self, c0):
        chosen, _ = tournament(self.A, c0, c0, NB)
        self.inv[c0] = np.linalg.inv(self.A[np.ix_(chosen, np.arange(c0, c0 + NB))])
        self.moves[c0] = net_moves(c0, chosen)
        dst, src = self.moves[c0]
        self.perm[dst] = self.perm[src]

    def k_moves_A(self, blocks, cols):
        """apply the row moves of `blocks` (in order) to A[:, cols]"""
        for c0 in blocks:
            dst, src = self.moves[c0]
            self.A[np.ix_(dst, cols)] = self.A[np.ix_(src, cols)]

    def k_moves_P(self, c0, cols):
        """row moves of block c0 applied to the saved panels Pbuf[:, cols]"""
        if len(cols) == 0:
            return
        dst, src = self.moves[c0]
        self.P[np.ix_(dst, cols)] = self.P[np.ix_(src, cols)]

    def k_extract_L(self, c0, cols):
        """JORDAN: rows of block c0 in the older live panels become L_ab blocks (used by the forward W
        solve) and are cleared in Pbuf so that one GEMM over all rows applies  -(strictly upper) W."""
        K = np.arange(c0, c0 + NB)
        if len(cols):
            self.L[np.ix_(K, cols)] = self.P[np.ix_(K, cols)]
            self.P[np.ix_(K, cols)] = 0

    def k_panel(self, c0):
        """near columns of the base step: save P, finish the pivot columns."""
        n = self.n
        K = np.arange(c0, c0 + NB)
        if self.jordan:
            notK = np.setdiff1d(np.arange(n), K)
            self.P[np.ix_(notK, K)] = self.A[np.ix_(notK, K)]
            self.P[np.ix_(K, K)] = 0
            self.A[np.ix_(K, K)] = self.inv[c0]
            self.A[np.ix_(notK, K)] = -self.P[np.ix_(notK, K)] @ self.inv[c0]
            self.W[np.ix_(K, K)] = self.inv[c0]
        else:
            below = np.arange(c0 + NB, n)
            self.P[np.ix_(below, K)] = self.A[np.ix_(below, K)]

    def k_wsolve(self, c0, cols, prev):
        """W_b = inv_b (A[rows_b, cols] - sum_{a in prev} L_ba W_a); written to A and Wbuf.
        prev = earlier blocks of the same leaf group (fused pre-update, FMA code on the device)."""
        K = np.arange(c0, c0 + NB)
        R = self.A[np.ix_(K, cols)].copy()
        Lsrc = self.L if self.jordan else self.P
        for a in prev:
            Ka = np.arange(a, a + NB)
            R -= Lsrc[np.ix_(K, Ka)] @ self.W[np.ix_(Ka, cols)]
        Wv = self.inv[c0] @ R
        self.A[np.ix_(K, cols)] = Wv
        self.W[np.ix_(K, cols)] = Wv

    # ---- recursion -------------------------------------------------------------------------
    def trsm(self, c0, w, cols):
        """forward W for blocks [c0, c0+w) on `cols` (moves already applied)"""
        nb = w // NB
        if nb <= 2:
            blocks = list(range(c0, c0 + w, NB))
            for i, b in enumerate(blocks):
                self.k_wsolve(b, cols, blocks[:i])
            return
        h = (nb + 1) // 2 * NB
        self.trsm(c0, h, cols)
        rows2 = np.arange(c0 + h, c0 + w)
        Lsrc = self.L if self.jordan else self.P
        gemm(self.A, rows2, cols, Lsrc, rows2, np.arange(c0, c0 + h), self.W, np.arange(c0, c0 + h), cols, "trsm")
        self.trsm(c0 + h, w - h, cols)

    def apply_far(self, c0, w, cols):
        if len(cols) == 0:
            return
        blocks = list(range(c0, c0 + w, NB))
        self.k_moves_A(blocks, cols)
        self.trsm(c0, w, cols)
        kc = np.arange(c0, c0 + w)
        if self.jordan:
            rows = np.arange(self.n)          # rows of the block too: Pbuf is strictly-upper there
        else:
            rows = np.arange(c0 + w, self.n)
        gemm(self.A, rows, cols, self.P, rows, kc, self.W, kc, cols, "far")

    def factor(self, c0, w):
        if w == NB:
            self.k_tournament(c0)
            K = np.arange(c0, c0 + NB)
            self.k_moves_A([c0], K)
            live = np.arange(self.top, c0)
            self.k_moves_P(c0, live)
            if self.jordan:
                self.k_extract_L(c0, live)
            self.k_panel(c0)
            return
        h = (w // NB + 1) // 2 * NB
        self.factor(c0, h)
        hi = self.ncol if (not self.jordan and c0 + w == self.n) else c0 + w     # aug columns ride along
        self.apply_far(c0, h, np.arange(c0 + h, hi))
        self.factor(c0 + h, w - h)
        if self.jordan:
            self.apply_far(c0 + h, w - h, np.arange(c0, c0 + h))

    def run(self):
        n = self.n
        self.factor(0, n)
        if self.jordan:
            G = np.empty((n, n), complex)
            G[:, self.perm] = self.A[:, :n]
            return G
        aug = np.arange(n, self.ncol)
        self.apply_far(n - NB, NB, aug)        # the last block is nobody's left sibling
        self.backsub(0, n, aug)
        return self.A[:, n:].copy()

    def backsub(self, c0, w, cols):
        """X[c0:c0+w] <- unit-block-upper solve, W blocks above the diagonal come from Wbuf"""
        if w <= NB:
            return
        h = (w // NB + 1) // 2 * NB
        self.backsub(c0 + h, w - h, cols)
        r1 = np.arange(c0, c0 + h)
        k2 = np.arange(c0 + h, c0 + w)
        # X1 -= W[r1, k2] X2 : left operand = rows of Wbuf (device: packed like a panel), right operand = A rows
        gemm(self.A, r1, cols, self.W, r1, k2, self.A, k2, cols, "back")
        self.backsub(c0, h, cols)


if __name__ == "__main__":
    from gaunegf_b200 import synthetic as sy
    for N in (64, 96, 256, 416, 1024):
        F, S = sy.hermitian_pair(N, seed=1)
        A = (0.2 + 0.01j) * S - F
        G0 = np.linalg.inv(A)
        nc = 8
        STATS.clear()
        X = Elim(A, np.eye(N)[:, -nc:], jordan=False).run()
        fstats = dict(STATS)
        STATS.clear()
        G = Elim(A, None, jordan=True).run()
        jstats = dict(STATS)
        print(N, "forward %.1e" % (np.abs(X - G0[:, -nc:]).max() / np.abs(G0).max()),
              "jordan %.1e" % (np.abs(G - G0).max() / np.abs(G0).max()), flush=True)
        for name, st, tot in (("forward", fstats, 8 / 3 * N**3 + 8 * N * N * nc), ("jordan", jstats, 8.0 * N**3)):
            byk = {}
            for (tag, k), fl in st.items():
                byk[k] = byk.get(k, 0) + fl
            s = sum(byk.values())
            print("   %s gemm flops / algorithmic = %.3f ; by K: %s" % (
                name, s / tot, ", ".join("%d:%.1f%%" % (k, 100 * v / s) for k, v in sorted(byk.items()))))
