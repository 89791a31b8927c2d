"""ctypes binding of the drop-in C++ classes GPisMap / GPisMap3 (include/gpismap/*.h, built into
gpismap_b200/libgpismap_host.so on top of libgpis_b200.so). Same method names and argument meaning
as the reference classes (cpp/include/GPisMap3.h:118-127, cpp/include/GPisMap.h:98-106)."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgpismap_host.so")

_libs = {}


def lib(path=None):
    path = path or LIB_PATH
    if path not in _libs:
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is not built (python -m gpismap_b200.build); there is no CPU fallback")
        L = C.CDLL(path)
        vp, f, i = C.c_void_p, C.c_float, C.c_int
        for pfx, posew in (("gm3", 12), ("gm2", 6)):
            getattr(L, pfx + "_create").restype = vp
            getattr(L, pfx + "_create").argtypes = [i]
            getattr(L, pfx + "_destroy").argtypes = [vp]
            getattr(L, pfx + "_destroy").restype = None
            getattr(L, pfx + "_reset").argtypes = [vp]
            getattr(L, pfx + "_test").argtypes = [vp, vp, i, vp]
            getattr(L, pfx + "_get_all_points").argtypes = [vp, vp, i]
            getattr(L, pfx + "_all_samples").argtypes = [vp, vp, i]
            getattr(L, pfx + "_leaves").argtypes = [vp, vp, vp, i]
            getattr(L, pfx + "_insert_samples").argtypes = [vp, vp, i]
            getattr(L, pfx + "_train_active").argtypes = [vp]
            getattr(L, pfx + "_activate_all").argtypes = [vp]
            getattr(L, pfx + "_timing").argtypes = [vp, vp, vp, vp]
            getattr(L, pfx + "_ctx").argtypes = [vp]
            getattr(L, pfx + "_ctx").restype = vp
        L.gm3_create_cam.restype = vp
        L.gm3_create_cam.argtypes = [i, f, f, f, f, i, i]
        L.gm3_set_cam.argtypes = [vp, f, f, f, f, i, i]
        L.gm3_set_tuning.argtypes = [vp, f, f, f, f, f]
        L.gm3_update.argtypes = [vp, vp, i, vp]
        L.gm2_update.argtypes = [vp, vp, vp, i, vp]
        _libs[path] = L
    return _libs[path]


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class _MapBase:
    dim = 0
    pfx = ""

    def _fn(self, name):
        return getattr(self.L, self.pfx + "_" + name)

    def close(self):
        if getattr(self, "h", None):
            self._fn("destroy")(self.h)
            self.h = None

    __del__ = close

    def reset(self):
        self._fn("reset")(self.h)

    resetMap = reset

    def test(self, x, res=None):
        """x: (n, dim) float32 → res (n, 2(1+dim)), read-modify-write like the reference."""
        x = np.ascontiguousarray(x, np.float32)
        n = x.shape[0]
        w = 2 * (1 + self.dim)
        if res is None:
            res = np.zeros((n, w), np.float32)
        ok = self._fn("test")(self.h, _p(x), n, _p(res))
        return res if ok else None

    def getAllPoints(self):
        n = self._fn("get_all_points")(self.h, None, 0)
        out = np.zeros((n, self.dim), np.float32)
        if n:
            self._fn("get_all_points")(self.h, _p(out), n)
        return out

    def all_samples(self):
        w = 2 * self.dim + 3
        n = self._fn("all_samples")(self.h, None, 0)
        out = np.zeros((n, w), np.float32)
        if n:
            self._fn("all_samples")(self.h, _p(out), n)
        return out

    def leaves(self):
        n = self._fn("leaves")(self.h, None, None, 0)
        c = np.zeros((n, self.dim), np.float32)
        k = np.zeros(n, np.int32)
        if n:
            self._fn("leaves")(self.h, _p(c), _p(k), n)
        return c, k

    def insert_samples(self, s):
        s = np.ascontiguousarray(s, np.float32)
        return self._fn("insert_samples")(self.h, _p(s), s.shape[0])

    def train_active(self):
        return self._fn("train_active")(self.h)

    def activate_all(self):
        """Mark every non-empty leaf dirty; the next train_active() retrains the whole map on its current samples."""
        return self._fn("activate_all")(self.h)

    def timing(self):
        ph = np.zeros(5, np.float64)
        cnt = np.zeros(3, np.int32)
        ms = C.c_float(0)
        self._fn("timing")(self.h, _p(ph), _p(cnt), C.byref(ms))
        return ph, cnt, ms.value

    def ctx_handle(self):
        return self._fn("ctx")(self.h)


class GPisMap3(_MapBase):
    dim = 3
    pfx = "gm3"

    def __init__(self, device=0, cam=None, libpath=None, rtimes=0.0, tree_min_half=0.0, tree_max_half=0.0,
                 tree_init_root_half=0.0, tree_cluster_half=0.0):
        """The tree / training-ball constants (reference: compile-time macros, params.h:40-44) default to the
        reference's values; pass e.g. rtimes=2.5 for BASELINE configs[4]'s larger leaves."""
        self.L = lib(libpath)
        self.h = self.L.gm3_create(device) if cam is None else self.L.gm3_create_cam(device, *cam)
        if rtimes or tree_min_half or tree_max_half or tree_init_root_half or tree_cluster_half:
            if not self.L.gm3_set_tuning(self.h, rtimes, tree_min_half, tree_max_half, tree_init_root_half, tree_cluster_half):
                raise RuntimeError("setTuning refused")

    def resetCam(self, fx, fy, cx, cy, w, h):
        self.L.gm3_set_cam(self.h, fx, fy, cx, cy, w, h)

    def update(self, dataz_colmajor, pose12):
        d = np.ascontiguousarray(dataz_colmajor, np.float32).ravel()
        p = np.ascontiguousarray(pose12, np.float32)
        self.L.gm3_update(self.h, _p(d), d.size, _p(p))


class GPisMap(_MapBase):
    dim = 2
    pfx = "gm2"

    def __init__(self, device=0, libpath=None):
        self.L = lib(libpath)
        self.h = self.L.gm2_create(device)

    def update(self, theta, ranges, pose6):
        t = np.ascontiguousarray(theta, np.float32).ravel()
        r = np.ascontiguousarray(ranges, np.float32).ravel()
        p = np.ascontiguousarray(pose6, np.float32)
        self.L.gm2_update(self.h, _p(t), _p(r), t.size, _p(p))
