"""TEST INFRASTRUCTURE ONLY (oracle). ctypes view of the plain-C restatement oracle
(oracle/gpis_oracle.c → libgpisoracle.so [fp32] / libgpisoracle64.so [fp64 shadow]).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def build(force=False):
    """Compile the C restatement (and, when /root/reference is present, oracle/_ref)."""
    need = force or not all(
        os.path.exists(os.path.join(_HERE, f)) for f in ("libgpisoracle.so", "libgpisoracle64.so", "libgpisoracle64r.so")
    )
    if need:
        subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"] + (["-B"] if force else []))
    subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])


class Oracle:
    def __init__(self, double=False, cov_float=False):
        """double: linear algebra (and, unless cov_float, the covariance entries) in fp64. cov_float (with double): the
        covariance entries are evaluated and rounded exactly as in the fp32 build, only the linear algebra is fp64 —
        the arbiter of the parity tests (gpis_oracle.c, COV_FLOAT)."""
        name = ("libgpisoracle64r.so" if cov_float else "libgpisoracle64.so") if double else "libgpisoracle.so"
        path = os.path.join(_HERE, name)
        if not os.path.exists(path):
            build()
        self.L = L = C.CDLL(path)
        self.rt = np.float64 if double else np.float32
        assert L.gpo_real_bytes() == np.dtype(self.rt).itemsize
        vp = C.c_void_p
        rp = np.ctypeslib.ndpointer(dtype=self.rt, flags="C_CONTIGUOUS")
        fp = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
        ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
        sig = {
            "gpo_matern_train": (C.c_int, [C.c_int, fp, fp, C.c_int, C.c_float, fp, fp, vp]),
            "gpo_gp_train": (vp, [C.c_int, fp, C.c_int, C.c_float, C.c_float]),
            "gpo_gp_free": (None, [vp]),
            "gpo_gp_n": (C.c_int, [vp]),
            "gpo_gp_ng": (C.c_int, [vp]),
            "gpo_gp_chol_fail": (C.c_int, [vp]),
            "gpo_gp_get": (None, [vp, vp, vp, vp]),
            "gpo_gp_test": (None, [vp, rp, C.c_int, rp]),
            "gpo_map_create": (vp, [C.c_int, C.c_int, fp, C.c_float, vp, C.c_float, C.c_float, C.c_float, vp]),
            "gpo_map_free": (None, [vp]),
            "gpo_map_test": (None, [vp, fp, C.c_int, rp, vp, vp]),
            "gpo_sort_replay": (None, [fp, C.c_int, ip]),
            "gpo_sort_replay_heapsorts": (C.c_int, []),
            "gpo_obs2d_train": (vp, [fp, fp, C.c_int, C.c_int]),
            "gpo_obs1d_train": (vp, [fp, fp, C.c_int]),
            "gpo_obs_free": (None, [vp]),
            "gpo_obs_ntiles": (C.c_int, [vp]),
            "gpo_obs_tile_counts": (None, [vp, ip]),
            "gpo_obs_bounds": (C.c_int, [vp, vp, vp]),
            "gpo_obs_test": (None, [vp, fp, C.c_int, rp, rp]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args

    # ---------------------------------------------------------------- covariance
    def matern_train(self, dim, x, gradflag, scale, sigx, siggrad):
        x = np.ascontiguousarray(x, np.float32)
        gf = np.ascontiguousarray(gradflag, np.float32)
        N = x.shape[0]
        n = N + dim * int((gf > 0.5).sum())
        K = np.zeros((n, n), self.rt)
        self.L.gpo_matern_train(dim, x, gf, N, scale, np.ascontiguousarray(sigx, np.float32),
                                np.ascontiguousarray(siggrad, np.float32), K.ctypes.data_as(C.c_void_p))
        return K

    # ---------------------------------------------------------------- leaf GP
    def gp_train(self, dim, samples, scale, noise):
        return OracleGP(self, dim, samples, scale, noise)

    def sort_replay(self, keys):
        k = np.ascontiguousarray(keys, np.float32).ravel()
        idx = np.zeros(k.size, np.int32)
        self.L.gpo_sort_replay(k, k.size, idx)
        return idx

    def make_map(self, dim, centres, cluster_half, gps, search_half, var_thre, noise, boxes=None):
        return OracleMap(self, dim, centres, cluster_half, gps, search_half, var_thre, noise, boxes)

    def obs2d(self, vu, zinv, ni, nj):
        return OracleObs(self, 2, np.ascontiguousarray(vu, np.float32).ravel(),
                         np.ascontiguousarray(zinv, np.float32).ravel(), ni, nj)

    def obs1d(self, theta, f):
        return OracleObs(self, 1, np.ascontiguousarray(theta, np.float32).ravel(),
                         np.ascontiguousarray(f, np.float32).ravel(), 0, 0)


class OracleGP:
    def __init__(self, o, dim, samples, scale, noise):
        self.o = o
        self.dim = dim
        s = np.ascontiguousarray(samples, np.float32)
        self.N = s.shape[0]
        self.h = o.L.gpo_gp_train(dim, s, self.N, scale, noise) if self.N > 0 else None
        self.n = o.L.gpo_gp_n(self.h) if self.h else 0
        self.ng = o.L.gpo_gp_ng(self.h) if self.h else 0

    @property
    def chol_fail(self):
        return self.o.L.gpo_gp_chol_fail(self.h)

    def factors(self):
        alpha = np.zeros(self.n, self.o.rt)
        L = np.zeros((self.n, self.n), self.o.rt)
        gf = np.zeros(self.N, np.float32)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        self.o.L.gpo_gp_get(self.h, p(alpha), p(L), p(gf))
        return alpha, L, gf

    def test(self, x, res=None):
        x = np.ascontiguousarray(x, self.o.rt)
        m = x.shape[0]
        if res is None:
            res = np.zeros((m, 2 * (1 + self.dim)), self.o.rt)
        self.o.L.gpo_gp_test(self.h, x, m, res)
        return res

    def __del__(self):
        if getattr(self, "h", None):
            self.o.L.gpo_gp_free(self.h)
            self.h = None


class OracleMap:
    def __init__(self, o, dim, centres, cluster_half, gps, search_half, var_thre, noise, boxes=None):
        self.o = o
        self.dim = dim
        self.gps = list(gps)  # keep alive
        c = np.ascontiguousarray(centres, np.float32)
        arr = (C.c_void_p * max(len(gps), 1))(*[(g.h if g is not None else None) for g in gps])
        b = np.ascontiguousarray(boxes, np.float32) if boxes is not None else None
        bp = b.ctypes.data_as(C.c_void_p) if b is not None else None
        self.h = o.L.gpo_map_create(dim, len(gps), c, cluster_half, arr, search_half, var_thre, noise, bp)

    def test(self, x, res=None, want_choice=False):
        x = np.ascontiguousarray(x, np.float32)
        m = x.shape[0]
        if res is None:
            res = np.zeros((m, 2 * (1 + self.dim)), self.o.rt)
        chosen = np.zeros((m, 4), np.int32)
        tie = np.zeros(m, np.int32)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        self.o.L.gpo_map_test(self.h, x, m, res, p(chosen), p(tie))
        return (res, chosen, tie) if want_choice else res

    def __del__(self):
        if getattr(self, "h", None):
            self.o.L.gpo_map_free(self.h)
            self.h = None


class OracleObs:
    def __init__(self, o, d, a, b, ni, nj):
        self.o = o
        self.d = d
        self.h = o.L.gpo_obs2d_train(a, b, ni, nj) if d == 2 else o.L.gpo_obs1d_train(a, b, a.size)

    def test(self, xt, val=None, var=None):
        xt = np.ascontiguousarray(xt, np.float32)
        m = xt.shape[0] if self.d == 2 else xt.size
        val = np.zeros(m, self.o.rt) if val is None else np.ascontiguousarray(val, self.o.rt).copy()
        var = np.zeros(m, self.o.rt) if var is None else np.ascontiguousarray(var, self.o.rt).copy()
        self.o.L.gpo_obs_test(self.h, xt, m, val, var)
        return val, var

    def tile_counts(self):
        n = self.o.L.gpo_obs_ntiles(self.h)
        c = np.zeros(max(n, 1), np.int32)
        self.o.L.gpo_obs_tile_counts(self.h, c)
        return c[:n]

    def bounds(self):
        bi = np.zeros(70000, np.float32)
        bj = np.zeros(70000, np.float32)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        k = self.o.L.gpo_obs_bounds(self.h, p(bi), p(bj))
        return bi[: k >> 16].copy(), bj[: k & 0xFFFF].copy()

    def __del__(self):
        if getattr(self, "h", None):
            self.o.L.gpo_obs_free(self.h)
            self.h = None
