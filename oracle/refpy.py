"""TEST INFRASTRUCTURE ONLY (oracle). ctypes view of oracle/_ref/libgpisref.so.

The library is the UNMODIFIED reference (/root/reference/cpp/src/*.cpp) built by
oracle/Makefile against oracle/eigen_shim, plus oracle/ref_harness.cpp. Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "libgpisref.so")

f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def available(variant=None):
    return os.path.exists(_path_of(variant))


def _path_of(variant):
    return _PATH if not variant else os.path.join(_HERE, "_ref", f"libgpisref_{variant}.so")


_lib = None
_variant = None


def use_variant(variant):
    """Switch to a variant build of the reference (oracle/Makefile: params_variants/<variant>/params.h override, e.g.
    "rtimes25"); None = the default build. Objects created before the switch keep working with their own library."""
    global _lib, _variant
    if variant != _variant:
        _variant = variant
        _lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_path_of(_variant))
        vp = C.c_void_p
        sig = {
            "ref_hardware_concurrency": (C.c_int, []),
            "ref_std_sort_indices": (None, [f32p, C.c_int, i32p]),
            "ref_std_sort_killer": (None, [C.c_int, C.c_int, f32p]),
            "ref_matern_train": (C.c_int, [C.c_int, f32p, f32p, C.c_int, C.c_float, f32p, f32p, f32p, C.c_int]),
            "ref_matern_test": (C.c_int, [C.c_int, f32p, f32p, C.c_int, f32p, C.c_int, C.c_float, f32p, C.c_int]),
            "ref_ou_train": (None, [C.c_int, f32p, C.c_int, C.c_float, C.c_float, f32p]),
            "ref_gp_train": (vp, [C.c_int, f32p, C.c_int, C.c_float, C.c_float]),
            "ref_gp_free": (None, [vp]),
            "ref_gp_n": (C.c_int, [vp]),
            "ref_gp_nsamples": (C.c_int, [vp]),
            "ref_gp_get": (None, [vp, vp, vp, vp]),
            "ref_gp_test": (None, [vp, f32p, C.c_int, f32p]),
            "ref_obs2d_create": (vp, []),
            "ref_obs2d_free": (None, [vp]),
            "ref_obs2d_train": (C.c_int, [vp, f32p, f32p, C.c_int, C.c_int]),
            "ref_obs2d_test": (None, [vp, f32p, C.c_int, f32p, f32p]),
            "ref_obs2d_tiles": (C.c_int, [vp, i32p, C.c_int]),
            "ref_obs2d_partition": (C.c_int, [vp, f32p, f32p, C.c_int]),
            "ref_obs1d_create": (vp, []),
            "ref_obs1d_free": (None, [vp]),
            "ref_obs1d_train": (C.c_int, [vp, f32p, f32p, C.c_int]),
            "ref_obs1d_test": (None, [vp, f32p, C.c_int, f32p, f32p]),
            "ref_obs1d_ranges": (C.c_int, [vp, f32p, C.c_int]),
            "ref3_create": (vp, []),
            "ref3_create_cam": (vp, [C.c_float] * 4 + [C.c_int] * 2),
            "ref3_destroy": (None, [vp]),
            "ref3_reset": (None, [vp]),
            "ref3_set_cam": (None, [vp] + [C.c_float] * 4 + [C.c_int] * 2),
            "ref3_update": (None, [vp, f32p, C.c_int, f32p]),
            "ref3_update_timed": (None, [vp, f32p, C.c_int, f32p, f64p, i32p]),
            "ref3_test": (C.c_int, [vp, f32p, C.c_int, f32p]),
            "ref3_get_all_points": (C.c_int, [vp, vp, C.c_int]),
            "ref3_all_samples": (C.c_int, [vp, vp, C.c_int]),
            "ref3_clusters": (C.c_int, [vp, vp, vp, vp, C.c_int]),
            "ref3_root": (None, [vp, f32p]),
            "ref3_cluster_boxes": (C.c_int, [vp, vp, C.c_int]),
            "ref2_cluster_boxes": (C.c_int, [vp, vp, C.c_int]),
            "ref3_train_set": (C.c_int, [vp, f32p, C.c_float, C.c_float, vp, C.c_int]),
            "ref3_cluster_gp": (C.c_int, [vp, f32p, vp, vp, vp, C.c_int]),
            "ref3_candidates": (C.c_int, [vp, f32p, C.c_float, f32p, f32p, C.c_int]),
            "ref3_insert_samples": (C.c_int, [vp, f32p, C.c_int]),
            "ref3_update_gps": (C.c_int, [vp, vp, vp]),
            "ref3_update_nogp": (None, [vp, f32p, C.c_int, f32p]),
            "ref3_activate": (C.c_int, [vp, f32p, f32p]),
            "ref3_tree_dump": (C.c_int, [vp, f32p, f32p, vp, C.c_int, f32p, C.c_int, f32p, i32p]),
            "ref3_tree_load": (C.c_int, [vp, vp, C.c_int, f32p, C.c_int, f32p]),
            "ref2_create": (vp, []),
            "ref2_destroy": (None, [vp]),
            "ref2_reset": (None, [vp]),
            "ref2_update": (None, [vp, f32p, f32p, C.c_int, f32p]),
            "ref2_update_timed": (None, [vp, f32p, f32p, C.c_int, f32p, f64p, i32p]),
            "ref2_test": (C.c_int, [vp, f32p, C.c_int, f32p]),
            "ref2_get_all_points": (C.c_int, [vp, vp, C.c_int]),
            "ref2_all_samples": (C.c_int, [vp, vp, C.c_int]),
            "ref2_clusters": (C.c_int, [vp, vp, vp, vp, C.c_int]),
            "ref2_root": (None, [vp, f32p]),
            "ref2_train_set": (C.c_int, [vp, f32p, C.c_float, C.c_float, vp, C.c_int]),
            "ref2_cluster_gp": (C.c_int, [vp, f32p, vp, vp, vp, C.c_int]),
            "ref2_candidates": (C.c_int, [vp, f32p, C.c_float, f32p, f32p, C.c_int]),
            "ref2_insert_samples": (C.c_int, [vp, f32p, C.c_int]),
            "ref2_update_gps": (C.c_int, [vp]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def std_sort_indices(keys):
    """std::sort of 0..n-1 by keys[a] < keys[b] with the libstdc++ the reference was built with."""
    k = f32(keys).ravel()
    idx = np.zeros(k.size, np.int32)
    lib().ref_std_sort_indices(k, k.size, idx)
    return idx


def std_sort_killer(n, ties=1):
    """Keys that drive this std::sort into its heapsort fallback (McIlroy's adversary); ties > 1 makes groups of equal keys."""
    k = np.zeros(n, np.float32)
    lib().ref_std_sort_killer(n, ties, k)
    return k


class RefGP:
    """OnGPIS trained on an explicit sample list (cpp/src/OnGPIS.cpp:34-149)."""

    def __init__(self, dim, samples, scale, noise):
        self.dim = dim
        s = f32(samples)
        self.h = lib().ref_gp_train(dim, s, s.shape[0], scale, noise)
        self.n = lib().ref_gp_n(self.h)
        self.N = s.shape[0]

    def factors(self):
        alpha = np.zeros(self.n, np.float32)
        L = np.zeros((self.n, self.n), np.float32)  # column-major on the C side
        gf = np.zeros(self.N, np.float32)
        lib().ref_gp_get(self.h, _ptr(alpha), _ptr(L), _ptr(gf))
        return alpha, L.T.copy(), gf  # L.T: row-major view of the column-major buffer

    def test(self, x, res=None):
        x = f32(x)
        m = x.shape[0]
        w = 2 * (1 + self.dim)
        if res is None:
            res = np.zeros((m, w), np.float32)
        lib().ref_gp_test(self.h, x, m, res)
        return res

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_gp_free(self.h)
            self.h = None


class _RefMapBase:
    dim = 0
    pfx = ""

    def _fn(self, name):
        return getattr(lib(), self.pfx + name)

    def reset(self):
        self._fn("reset")(self.h)

    def test(self, x, res=None):
        x = f32(x)
        m = x.shape[0]
        w = 2 * (1 + self.dim)
        if res is None:
            res = np.zeros((m, w), np.float32)
        ok = self._fn("test")(self.h, x, m, res)
        return res if ok else None

    def all_points(self):
        n = self._fn("get_all_points")(self.h, None, 0)
        out = np.zeros((n, self.dim), np.float32)
        if n:
            self._fn("get_all_points")(self.h, _ptr(out), n)
        return out

    def all_samples(self):
        w = 2 * self.dim + 3
        n = self._fn("all_samples")(self.h, None, 0)
        out = np.zeros((n, w), np.float32)
        if n:
            self._fn("all_samples")(self.h, _ptr(out), n)
        return out

    def clusters(self):
        n = self._fn("clusters")(self.h, None, None, None, 0)
        c = np.zeros((n, self.dim), np.float32)
        ns = np.zeros(n, np.int32)
        tr = np.zeros(n, np.int32)
        if n:
            self._fn("clusters")(self.h, _ptr(c), _ptr(ns), _ptr(tr), n)
        return c, ns, tr

    def cluster_boxes(self):
        """Effective (ancestor-intersected) float boxes of the non-empty clusters, DFS order: (n, 2*dim)."""
        n = self._fn("clusters")(self.h, None, None, None, 0)
        b = np.zeros((n, 2 * self.dim), np.float32)
        if n:
            self._fn("cluster_boxes")(self.h, _ptr(b), n)
        return b

    def root(self):
        r = np.zeros(4, np.float32)
        self._fn("root")(self.h, r)
        return r[: self.dim].copy(), float(r[self.dim])

    def train_set(self, centre, half, rtimes, cap=8192):
        w = 2 * self.dim + 3
        out = np.zeros((cap, w), np.float32)
        n = self._fn("train_set")(self.h, f32(centre), half, rtimes, _ptr(out), cap)
        assert n <= cap
        return out[:n].copy()

    def cluster_gp(self, centre, cap_n=4096):
        N = C.c_int(0)
        n = self._fn("cluster_gp")(self.h, f32(centre), None, None, C.byref(N), 0)
        if n == 0:
            return None
        alpha = np.zeros(n, np.float32)
        L = np.zeros((n, n), np.float32)
        self._fn("cluster_gp")(self.h, f32(centre), _ptr(alpha), _ptr(L), C.byref(N), n)
        return alpha, L.T.copy(), N.value

    def candidates(self, x, half, cap=512):
        c = np.zeros((cap, self.dim), np.float32)
        d = np.zeros(cap, np.float32)
        n = self._fn("candidates")(self.h, f32(x), half, c, d, cap)
        assert n <= cap
        return c[:n].copy(), d[:n].copy()

    def insert_samples(self, samples):
        s = f32(samples)
        return self._fn("insert_samples")(self.h, s, s.shape[0])


class RefMap3(_RefMapBase):
    """The reference GPisMap3 (cpp/include/GPisMap3.h:84-140)."""
    dim = 3
    pfx = "ref3_"

    def __init__(self, cam=None):
        if cam is None:
            self.h = lib().ref3_create()
        else:
            self.h = lib().ref3_create_cam(*cam)

    def set_cam(self, fx, fy, cx, cy, w, h):
        lib().ref3_set_cam(self.h, fx, fy, cx, cy, w, h)

    def update(self, depth_colmajor, pose12, timed=False, nogp=False):
        d = f32(depth_colmajor).ravel()
        p = f32(pose12)
        if nogp:      # everything but updateGPs (the samples do not depend on the leaf GPs)
            lib().ref3_update_nogp(self.h, d, d.size, p)
            return
        if timed:
            ph = np.zeros(5, np.float64)
            cnt = np.zeros(2, np.int32)
            lib().ref3_update_timed(self.h, d, d.size, p, ph, cnt)
            return ph, cnt
        lib().ref3_update(self.h, d, d.size, p)

    def tree_dump(self, lo, hi):
        """Pre-order snapshot of the octree restricted to the box [lo,hi]: (flags uint8, samples (ns, 9), root [cx,cy,cz,half])."""
        lo, hi = f32(lo), f32(hi)
        root = np.zeros(4, np.float32)
        ns = np.zeros(1, np.int32)
        nf = lib().ref3_tree_dump(self.h, lo, hi, None, 0, np.zeros(9, np.float32), 0, root, ns)
        flags = np.zeros(max(nf, 1), np.uint8)
        smp = np.zeros((max(int(ns[0]), 1), 9), np.float32)
        lib().ref3_tree_dump(self.h, lo, hi, flags.ctypes.data_as(C.c_void_p), nf, smp, int(ns[0]), root, ns)
        return flags[:nf], smp[:int(ns[0])], root

    def tree_load(self, flags, samples, root):
        flags = np.ascontiguousarray(flags, np.uint8)
        smp = f32(samples)
        n = lib().ref3_tree_load(self.h, flags.ctypes.data_as(C.c_void_p), flags.size, smp, smp.shape[0], f32(root))
        if n < 0:
            raise ValueError("tree snapshot is inconsistent")
        return n

    def activate(self, lo, hi):
        return lib().ref3_activate(self.h, f32(lo), f32(hi))

    def update_gps(self, lo=None, hi=None):
        lo_ = f32(lo) if lo is not None else None
        hi_ = f32(hi) if hi is not None else None
        return lib().ref3_update_gps(self.h, _ptr(lo_), _ptr(hi_))

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref3_destroy(self.h)
            self.h = None


class RefMap2(_RefMapBase):
    """The reference GPisMap (cpp/include/GPisMap.h:70-119)."""
    dim = 2
    pfx = "ref2_"

    def __init__(self):
        self.h = lib().ref2_create()

    def update(self, theta, ranges, pose6, timed=False):
        t = f32(theta).ravel()
        r = f32(ranges).ravel()
        p = f32(pose6)
        if timed:
            ph = np.zeros(5, np.float64)
            cnt = np.zeros(2, np.int32)
            lib().ref2_update_timed(self.h, t, r, t.size, p, ph, cnt)
            return ph, cnt
        lib().ref2_update(self.h, t, r, t.size, p)

    def update_gps(self):
        return lib().ref2_update_gps(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref2_destroy(self.h)
            self.h = None


class RefObs2D:
    """ObsGP2D (cpp/src/ObsGP.cpp:204-463)."""

    def __init__(self):
        self.h = lib().ref_obs2d_create()

    def train(self, vu, zinv, ni, nj):
        self._vu = f32(vu).ravel().copy()
        self._z = f32(zinv).ravel().copy()
        return lib().ref_obs2d_train(self.h, self._vu, self._z, ni, nj)

    def test(self, xt, val=None, var=None):
        xt = f32(xt)
        m = xt.shape[0]
        val = np.zeros(m, np.float32) if val is None else f32(val).copy()
        var = np.zeros(m, np.float32) if var is None else f32(var).copy()
        lib().ref_obs2d_test(self.h, xt, m, val, var)
        return val, var

    def tiles(self, cap=1 << 16):
        n = np.zeros(cap, np.int32)
        k = lib().ref_obs2d_tiles(self.h, n, cap)
        return n[:k].copy()

    def partition(self, cap=4096):
        a = np.zeros(cap, np.float32)
        b = np.zeros(cap, np.float32)
        k = lib().ref_obs2d_partition(self.h, a, b, cap)
        return a[: k >> 16].copy(), b[: k & 0xFFFF].copy()

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_obs2d_free(self.h)
            self.h = None


class RefObs1D:
    """ObsGP1D (cpp/src/ObsGP.cpp:76-187)."""

    def __init__(self):
        self.h = lib().ref_obs1d_create()

    def train(self, theta, f):
        self._t = f32(theta).ravel().copy()
        self._f = f32(f).ravel().copy()
        return lib().ref_obs1d_train(self.h, self._t, self._f, self._t.size)

    def test(self, xt, val=None, var=None):
        xt = f32(xt).ravel()
        m = xt.size
        val = np.zeros(m, np.float32) if val is None else f32(val).copy()
        var = np.zeros(m, np.float32) if var is None else f32(var).copy()
        lib().ref_obs1d_test(self.h, xt, m, val, var)
        return val, var

    def ranges(self, cap=4096):
        a = np.zeros(cap, np.float32)
        k = lib().ref_obs1d_ranges(self.h, a, cap)
        return a[:k].copy()

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_obs1d_free(self.h)
            self.h = None


def matern_train(dim, x, gradflag, scale, sigx, siggrad):
    """covFnc.cpp:111-124. x: (N, dim). Returns K (n, n)."""
    x = f32(x)
    N = x.shape[0]
    ng = int((np.asarray(gradflag) > 0.5).sum())
    n = N + dim * ng
    K = np.zeros((n, n), np.float32)
    lib().ref_matern_train(dim, x, f32(gradflag), N, scale, f32(sigx), f32(siggrad), K, n * n)
    return K.T.copy()


def matern_test(dim, x, gradflag, xt, scale):
    """covFnc.cpp:126-139. Returns K* (n, m(1+dim))."""
    x = f32(x)
    xt = f32(xt)
    N, m = x.shape[0], xt.shape[0]
    ng = int((np.asarray(gradflag) > 0.5).sum())
    n = N + dim * ng
    K = np.zeros((m * (1 + dim), n), np.float32)
    lib().ref_matern_test(dim, x, f32(gradflag), N, xt, m, scale, K, K.size)
    return K.T.copy()
