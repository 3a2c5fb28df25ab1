"""ctypes binding of libgaunegf_b200.so (C ABI: include/gaunegf_b200.h).

There is NO CPU fallback: if the library is missing or no CUDA device is present, creating a
context raises.  numpy in, numpy out; torch is only used (elsewhere) for device buffers and NCCL.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "libgaunegf_b200.so")

HOST, DEVICE = 0, 1
ERR_SINGULAR = 3

_lib = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_vp = C.c_void_p

SIGNATURES = {
    "gnb_create": (C.c_int, [C.POINTER(_vp), C.c_int]),
    "gnb_destroy": (C.c_int, [_vp]),
    "gnb_last_error": (C.c_char_p, [_vp]),
    "gnb_version": (C.c_char_p, []),
    "gnb_set_stream": (C.c_int, [_vp, _vp]),
    "gnb_set_workspace_limit": (C.c_int, [_vp, C.c_size_t]),
    "gnb_launch_count": (C.c_int64, [_vp]),
    "gnb_last_elim_ms": (C.c_double, [_vp]),
    "gnb_last_elim_flops": (C.c_double, [_vp]),
    "gnb_set_timing": (C.c_int, [_vp, C.c_int]),
    "gnb_gemm_stats": (C.c_int, [_vp, _dp, _dp, C.POINTER(C.c_int64), C.c_int]),
    "gnb_set_system": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_int]),
    "gnb_set_system_cached": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_int, _vp]),
    "gnb_system_differs": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_int, C.c_int, _vp]),
    "gnb_set_system_known": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_int, C.c_int, _vp]),
    "gnb_sigma_clear": (C.c_int, [_vp]),
    "gnb_sigma_set_dense0": (C.c_int, [_vp, _vp, C.c_int]),
    "gnb_sigma_add_const_block": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "gnb_sigma_add_chain1d": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_double, C.c_double,
                                        C.c_double, C.c_int]),
    "gnb_sigma_add_bethe": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, C.c_double, C.c_double,
                                      C.c_double, C.c_int]),
    "gnb_sigma_set_transform": (C.c_int, [_vp, C.c_int, _vp, C.c_int, C.c_int]),
    "gnb_sigma_eval": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    "gnb_green": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_int]),
    "gnb_transmission": (C.c_int, [_vp, C.c_int, _vp, C.c_int, C.c_int, _vp]),
    "gnb_dos": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp]),
    "gnb_gr_int": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, C.c_int]),
    "gnb_gr_int_seg": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_int, _vp, _vp, C.c_int]),
    "gnb_gless_int": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_int, _vp, C.c_int]),
    "gnb_green_dense": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_long, _vp, C.c_int]),
    "gnb_transmission_dense": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_long, _vp, C.c_long, _vp, C.c_long, _vp]),
    "gnb_transmission_spin": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_long, _vp, C.c_long, _vp, C.c_long, _vp]),
    "gnb_dos_dense": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_long, _vp, _vp]),
    "gnb_gr_int_dense": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, C.c_long, _vp, C.c_int]),
    "gnb_gr_int_seg_dense": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_int, _vp, _vp, C.c_long, _vp, C.c_int]),
    "gnb_gless_int_dense": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, C.c_long, _vp, C.c_long, _vp, C.c_int]),
    "gnb_inverse_batch": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp, C.c_int]),
}


# developer switches / probes (include/gaunegf_b200_dev.h): process-wide, not part of the drop-in ABI
DEV_SIGNATURES = {
    "gnb_dev_set_option": (C.c_int, [C.c_char_p, C.c_int]),
    "gnb_dev_trace_start": (C.c_int, []),
    "gnb_dev_trace_dump": (C.c_int, [C.c_char_p]),
    "gnb_dev_gemm_bench": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _dp]),
    "gnb_dev_fp64_peak": (C.c_int, [_vp, C.c_double, _dp]),
}


def load_library(path=None):
    """dlopen the in-tree library and bind every symbol the header declares."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(
            f"{p} not found: build it with `python -m gaunegf_b200.build` (nvcc, sm_100a). "
            "gaunegf_b200 has no CPU fallback.")
    lib = C.CDLL(p)
    for name, (res, args) in list(SIGNATURES.items()) + list(DEV_SIGNATURES.items()):
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    if path is None:
        _lib = lib
    return lib


def c128(a):
    """contiguous complex128 view/copy (row-major) of an array-like"""
    return np.ascontiguousarray(np.asarray(a), dtype=np.complex128)


def ptr(a):
    return a.ctypes.data_as(_vp) if a is not None else None


class GnbError(RuntimeError):
    pass


class Context:
    """One context per GPU / host thread; owns the device workspace."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = _vp()
        rc = self.lib.gnb_create(C.byref(h), int(device))
        if rc != 0 or not h:
            raise RuntimeError(
                "gaunegf_b200: could not create a CUDA context on device %d (rc=%d). A B200 (sm_100a) "
                "GPU is required; there is no CPU fallback." % (device, rc))
        self.h = h
        self.device = device
        self.N = 0
        self.last_system_upload = 3
        self.system_uploads_skipped = 0
        # workspace limit in GiB (default 48 GiB inside the library): GNB_WS_GIB=100
        if os.environ.get("GNB_WS_GIB"):
            self.lib.gnb_set_workspace_limit(self.h, C.c_size_t(int(float(os.environ["GNB_WS_GIB"]) * (1 << 30))))
        # developer A/B switches: GNB_DEV_OPTS="rk_m3=1,tourn_group=256"
        for kv in filter(None, os.environ.get("GNB_DEV_OPTS", "").split(",")):
            k, v = kv.split("=")
            if self.lib.gnb_dev_set_option(k.strip().encode(), int(v)) != 0:
                raise ValueError(f"GNB_DEV_OPTS: unknown developer option {k!r}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.gnb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc == 0:
            return
        msg = self.lib.gnb_last_error(self.h).decode()
        if rc == ERR_SINGULAR:
            raise np.linalg.LinAlgError(msg or "Singular matrix")
        if rc == 2:
            raise ValueError(msg)
        raise GnbError(f"libgaunegf_b200 rc={rc}: {msg}")

    # -- plumbing -------------------------------------------------------------------------
    def set_stream(self, cuda_stream):
        self.check(self.lib.gnb_set_stream(self.h, _vp(cuda_stream)))

    def set_workspace_limit(self, nbytes):
        self.check(self.lib.gnb_set_workspace_limit(self.h, int(nbytes)))

    def set_timing(self, on):
        self.check(self.lib.gnb_set_timing(self.h, int(bool(on))))

    def gemm_stats(self, reset=False):
        """(ms, algorithmic flops, launches) of the rank-K update kernel while timing was on"""
        ms, fl, n = C.c_double(), C.c_double(), C.c_int64()
        self.check(self.lib.gnb_gemm_stats(self.h, C.byref(ms), C.byref(fl), C.byref(n), int(reset)))
        return ms.value, fl.value, n.value

    @property
    def launches(self):
        return int(self.lib.gnb_launch_count(self.h))

    @property
    def last_elim_ms(self):
        return float(self.lib.gnb_last_elim_ms(self.h))

    @property
    def last_elim_flops(self):
        return float(self.lib.gnb_last_elim_flops(self.h))

    def fp64_peak(self, ms_target=200.0):
        """FP64 tensor-pipe (DMMA.8x8x4) ceiling of this GPU in TFLOP/s, measured now (developer probe)"""
        out = C.c_double()
        self.check(self.lib.gnb_dev_fp64_peak(self.h, float(ms_target), C.byref(out)))
        return out.value

    # -- system / sigma -------------------------------------------------------------------
    def set_system(self, F, S):
        """F, S of the following calls.  They stay resident in HBM: the library compares the arrays with pinned shadow
        copies (one multi-threaded pass) and uploads only the matrix that changed (gnb_set_system_cached) — the reference
        re-sends F and S on every integrator call (integrate.py:92-95), the adaptive drivers make up to six such calls
        per integral, and an SCF step changes F but not S."""
        Fa, Sa = np.asarray(F), np.asarray(S)
        assert Fa.shape == Sa.shape, "F and S must have the same shape"
        assert Fa.ndim == 2 and Fa.shape[0] == Fa.shape[1], "F and S must be square matrices"
        (Fc, fr), (Sc, sr) = self._as_input(Fa), self._as_input(Sa)
        self.N = Fc.shape[0]
        up = C.c_int(0)
        self.check(self.lib.gnb_set_system_cached(self.h, self.N, ptr(Fc), ptr(Sc), fr | (sr << 1), C.byref(up)))
        self.last_system_upload = up.value                 # bit 0: F was sent, bit 1: S was sent
        if up.value == 0:
            self.system_uploads_skipped += 1

    @staticmethod
    def _as_input(a):           # real float64 arrays go to the library as they are (no complex copy per call)
        if a.dtype == np.float64 and a.flags.c_contiguous:
            return a, 1
        return c128(a), 0

    def system_differs(self, F, S, full=True):
        """read-only: which of (F, S) differ from the resident pair?  bit 0 = F, bit 1 = S (3: nothing comparable resident).
        full=False: size + strided sample (parallel.set_system)"""
        Fa, Sa = np.asarray(F), np.asarray(S)
        if Fa.shape != Sa.shape or Fa.ndim != 2 or Fa.shape[0] != Fa.shape[1]:
            return 3
        (Fc, fr), (Sc, sr) = self._as_input(Fa), self._as_input(Sa)
        d = C.c_int(3)
        self.check(self.lib.gnb_system_differs(self.h, Fc.shape[0], ptr(Fc), ptr(Sc), fr | (sr << 1), int(bool(full)), C.byref(d)))
        return int(d.value)

    def set_system_known(self, F, S, changed):
        """set_system without the comparison: `changed` (bit 0 = F, bit 1 = S) names what to copy and upload"""
        Fa, Sa = np.asarray(F), np.asarray(S)
        assert Fa.shape == Sa.shape and Fa.ndim == 2 and Fa.shape[0] == Fa.shape[1], "F and S must be square matrices of one size"
        (Fc, fr), (Sc, sr) = self._as_input(Fa), self._as_input(Sa)
        self.N = Fc.shape[0]
        up = C.c_int(0)
        self.check(self.lib.gnb_set_system_known(self.h, self.N, ptr(Fc), ptr(Sc), fr | (sr << 1), int(changed) & 3, C.byref(up)))
        self.last_system_upload = up.value
        if up.value == 0:
            self.system_uploads_skipped += 1

    def set_system_device(self, N, F_ptr, S_ptr):
        self.N = int(N)
        self.check(self.lib.gnb_set_system(self.h, self.N, _vp(F_ptr), _vp(S_ptr), DEVICE))

    def sigma_set_transform(self, n, Xi=None, spin_mode=0):
        """Sigma_tot = expand(Xi Sigma Xi) (surfGBethe.py:529-539); contacts added afterwards index the n-space"""
        Xi = None if Xi is None else c128(Xi)
        if Xi is not None:
            assert Xi.shape == (n, n)
        self.check(self.lib.gnb_sigma_set_transform(self.h, int(n), ptr(Xi), int(spin_mode), HOST))

    def sigma_clear(self):
        self.check(self.lib.gnb_sigma_clear(self.h))

    def sigma_set_dense0(self, sig0):
        sig0 = c128(sig0)
        assert sig0.shape == (self.N, self.N)
        self.check(self.lib.gnb_sigma_set_dense0(self.h, ptr(sig0), HOST))

    def sigma_add_const_block(self, inds, blk):
        inds = np.ascontiguousarray(inds, dtype=np.int32)
        blk = c128(blk)
        assert blk.shape == (len(inds), len(inds))
        self.check(self.lib.gnb_sigma_add_const_block(self.h, len(inds), ptr(inds), ptr(blk)))

    def sigma_add_chain1d(self, inds, alpha, Salpha, beta, Sbeta, tau, stau, eta, conv, relax, max_iter=2000):
        inds = np.ascontiguousarray(inds, dtype=np.int32)
        mats = [c128(m) for m in (alpha, Salpha, beta, Sbeta, tau, stau)]
        n = len(inds)
        for m in mats:
            if m.shape != (n, n):
                raise ValueError("chain1d: every contact matrix must be (len(inds), len(inds)); got %s" % (m.shape,))
        self.check(self.lib.gnb_sigma_add_chain1d(self.h, n, ptr(inds), *[ptr(m) for m in mats],
                                                  float(eta), float(conv), float(relax), int(max_iter)))

    def sigma_add_bethe(self, atom_inds, nb_lists, H, Slist, Vlist, eta, conv, mix=0.5, max_iter=1000):
        inds = np.ascontiguousarray(np.asarray(atom_inds).reshape(-1), dtype=np.int32)
        natoms = len(nb_lists)
        assert inds.size == natoms * 9
        off = np.zeros(natoms + 1, dtype=np.int32)
        off[1:] = np.cumsum([len(x) for x in nb_lists])
        dirs = np.ascontiguousarray([d for x in nb_lists for d in x] or [0], dtype=np.int32)
        H, Sl, Vl = c128(H), c128(Slist), c128(Vlist)
        assert H.shape == (9, 9) and Sl.shape == (12, 9, 9) and Vl.shape == (12, 9, 9)
        self.check(self.lib.gnb_sigma_add_bethe(self.h, natoms, ptr(inds), ptr(off), ptr(dirs), ptr(H), ptr(Sl),
                                                ptr(Vl), float(eta), float(conv), float(mix), int(max_iter)))

    def sigma_eval(self, contact, which, E, shape):
        E = c128(np.atleast_1d(E))
        M = E.size
        out = np.empty((M,) + tuple(shape), dtype=np.complex128)
        iters = np.zeros(M, dtype=np.int32)
        diffs = np.zeros(M, dtype=np.float64)
        self.check(self.lib.gnb_sigma_eval(self.h, int(contact), int(which), M, ptr(E), ptr(out), ptr(iters), ptr(diffs)))
        return out, iters, diffs

    # -- reductions ------------------------------------------------------------------------
    def green(self, E):
        E = c128(np.atleast_1d(E))
        G = np.empty((E.size, self.N, self.N), dtype=np.complex128)
        self.check(self.lib.gnb_green(self.h, E.size, ptr(E), ptr(G), HOST))
        return G

    def transmission(self, E, ca=0, cb=-1):
        E = c128(np.atleast_1d(E))
        T = np.empty(E.size, dtype=np.float64)
        self.check(self.lib.gnb_transmission(self.h, E.size, ptr(E), int(ca), int(cb), ptr(T)))
        return T

    def dos(self, E, per_site=True):
        E = c128(np.atleast_1d(E))
        tot = np.empty(E.size, dtype=np.float64)
        per = np.empty((E.size, self.N), dtype=np.float64) if per_site else None
        self.check(self.lib.gnb_dos(self.h, E.size, ptr(E), ptr(tot), ptr(per)))
        return tot, per

    def gr_int(self, E, w, out_device_ptr=None):
        E, w = c128(np.atleast_1d(E)), c128(np.atleast_1d(w))
        assert E.size == w.size, "Elist and weights must have the same length"
        if out_device_ptr is not None:
            self.check(self.lib.gnb_gr_int(self.h, E.size, ptr(E), ptr(w), _vp(out_device_ptr), DEVICE))
            return None
        out = np.empty((self.N, self.N), dtype=np.complex128)
        self.check(self.lib.gnb_gr_int(self.h, E.size, ptr(E), ptr(w), ptr(out), HOST))
        return out

    def gr_int_seg(self, E, w, seg_end, sig=None, out_device_ptr=None):
        """len(seg_end) weighted sums over consecutive energy ranges in one batch -> (nseg, N, N); sig: dense Sigma_tot
        (one matrix or one per energy) instead of the described contacts."""
        E, w = c128(np.atleast_1d(E)), c128(np.atleast_1d(w))
        assert E.size == w.size, "Elist and weights must have the same length"
        ends = np.ascontiguousarray(seg_end, dtype=np.int32)
        if sig is None:
            call = lambda o, loc: self.lib.gnb_gr_int_seg(self.h, E.size, ptr(E), ptr(w), ends.size, ptr(ends), o, loc)
        else:
            s, ss = self._dense(sig, E.size)
            call = lambda o, loc: self.lib.gnb_gr_int_seg_dense(self.h, E.size, ptr(E), ptr(w), ends.size, ptr(ends), ptr(s), ss, o, loc)
        if out_device_ptr is not None:
            self.check(call(_vp(out_device_ptr), DEVICE))
            return None
        out = np.empty((ends.size, self.N, self.N), dtype=np.complex128)
        self.check(call(ptr(out), HOST))
        return out

    def gless_int(self, E, w, contact=-1, out_device_ptr=None):
        E, w = c128(np.atleast_1d(E)), c128(np.atleast_1d(w))
        assert E.size == w.size, "Elist and weights must have the same length"
        if out_device_ptr is not None:
            self.check(self.lib.gnb_gless_int(self.h, E.size, ptr(E), ptr(w), int(contact), _vp(out_device_ptr), DEVICE))
            return None
        out = np.empty((self.N, self.N), dtype=np.complex128)
        self.check(self.lib.gnb_gless_int(self.h, E.size, ptr(E), ptr(w), int(contact), ptr(out), HOST))
        return out

    # -- dense (caller-evaluated sigma) variants -----------------------------------------------
    def _dense(self, a, M):
        if a is None:
            return None, 0
        a = c128(a)
        if a.ndim == 2:
            assert a.shape == (self.N, self.N)
            return a, 0
        assert a.shape == (M, self.N, self.N)
        return a, self.N * self.N

    def green_dense(self, E, sig):
        E = c128(np.atleast_1d(E))
        s, ss = self._dense(sig, E.size)
        G = np.empty((E.size, self.N, self.N), dtype=np.complex128)
        self.check(self.lib.gnb_green_dense(self.h, E.size, ptr(E), ptr(s), ss, ptr(G), HOST))
        return G

    def transmission_dense(self, E, sig, gam1, gam2):
        E = c128(np.atleast_1d(E))
        s, ss = self._dense(sig, E.size)
        g1, s1 = self._dense(gam1, E.size)
        g2, s2 = self._dense(gam2, E.size)
        T = np.empty(E.size, dtype=np.float64)
        self.check(self.lib.gnb_transmission_dense(self.h, E.size, ptr(E), ptr(s), ss, ptr(g1), s1, ptr(g2), s2, ptr(T)))
        return T

    def transmission_spin_described(self, E):
        """spin-resolved T(E) from the described (spin-expanded) self-energies, contacts 0 and -1"""
        E = c128(np.atleast_1d(E))
        T4 = np.empty((E.size, 4), dtype=np.float64)
        self.check(self.lib.gnb_transmission_spin(self.h, E.size, ptr(E), None, 0, None, 0, None, 0, ptr(T4)))
        return T4

    def transmission_spin(self, E, sig, gam1, gam2):
        E = c128(np.atleast_1d(E))
        s, ss = self._dense(sig, E.size)
        g1, s1 = self._dense(gam1, E.size)
        g2, s2 = self._dense(gam2, E.size)
        T4 = np.empty((E.size, 4), dtype=np.float64)
        self.check(self.lib.gnb_transmission_spin(self.h, E.size, ptr(E), ptr(s), ss, ptr(g1), s1, ptr(g2), s2, ptr(T4)))
        return T4

    def dos_dense(self, E, sig, per_site=True):
        E = c128(np.atleast_1d(E))
        s, ss = self._dense(sig, E.size)
        tot = np.empty(E.size, dtype=np.float64)
        per = np.empty((E.size, self.N), dtype=np.float64) if per_site else None
        self.check(self.lib.gnb_dos_dense(self.h, E.size, ptr(E), ptr(s), ss, ptr(tot), ptr(per)))
        return tot, per

    def gr_int_dense(self, E, w, sig, out_device_ptr=None):
        E, w = c128(np.atleast_1d(E)), c128(np.atleast_1d(w))
        assert E.size == w.size, "Elist and weights must have the same length"
        s, ss = self._dense(sig, E.size)
        if out_device_ptr is not None:
            self.check(self.lib.gnb_gr_int_dense(self.h, E.size, ptr(E), ptr(w), ptr(s), ss, _vp(out_device_ptr), DEVICE))
            return None
        out = np.empty((self.N, self.N), dtype=np.complex128)
        self.check(self.lib.gnb_gr_int_dense(self.h, E.size, ptr(E), ptr(w), ptr(s), ss, ptr(out), HOST))
        return out

    def gless_int_dense(self, E, w, sig, gam, out_device_ptr=None):
        E, w = c128(np.atleast_1d(E)), c128(np.atleast_1d(w))
        assert E.size == w.size, "Elist and weights must have the same length"
        s, ss = self._dense(sig, E.size)
        g, gs = self._dense(gam, E.size)
        if out_device_ptr is not None:
            self.check(self.lib.gnb_gless_int_dense(self.h, E.size, ptr(E), ptr(w), ptr(s), ss, ptr(g), gs,
                                                    _vp(out_device_ptr), DEVICE))
            return None
        out = np.empty((self.N, self.N), dtype=np.complex128)
        self.check(self.lib.gnb_gless_int_dense(self.h, E.size, ptr(E), ptr(w), ptr(s), ss, ptr(g), gs, ptr(out), HOST))
        return out

    def inverse_batch(self, A):
        A = c128(A)
        single = A.ndim == 2
        if single:
            A = A[None]
        assert A.ndim == 3 and A.shape[1] == A.shape[2]
        out = np.empty_like(A)
        self.check(self.lib.gnb_inverse_batch(self.h, A.shape[1], A.shape[0], ptr(A), ptr(out), HOST))
        return out[0] if single else out


_default_ctx = {}


def default_context(device=None):
    """Process-wide context for the current GPU (LOCAL_RANK-aware under torchrun)."""
    if device is None:
        device = int(os.environ.get("GAUNEGF_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]
