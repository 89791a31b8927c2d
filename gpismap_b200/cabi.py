"""ctypes binding of the C ABI in include/gpis_b200.h (libgpis_b200.so).

This is plumbing for tests and bench.py; the product boundary is the C ABI itself (and the C++
classes in include/gpismap/ that sit on it). There is no CPU fallback: if the library is missing
or no sm_100 device is present, construction raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgpis_b200.so")
if os.environ.get("GPIS_B200_LIB"):      # development builds (e.g. -DT2_TIMING) next to the production library
    LIB_PATH = os.environ["GPIS_B200_LIB"]

EXPORTS = [
    "gpis_config_default", "gpis_create", "gpis_destroy", "gpis_reset", "gpis_last_error", "gpis_device",
    "gpis_leaves_update", "gpis_leaves_mark", "gpis_leaves_set_boxes", "gpis_leaves_erase", "gpis_rebase", "gpis_get_rebase", "gpis_leaf_get",
    "gpis_query", "gpis_query_device", "gpis_query_debug", "gpis_leaf_index",
    "gpis_obs_train_2d", "gpis_obs_train_1d", "gpis_obs_test",
    "gpis_get_stats", "gpis_comm_unique_id", "gpis_comm_init", "gpis_replicate",
    "gpis_snapshot_save", "gpis_snapshot_load", "gpis_samples_set", "gpis_leaves_train_dirty", "gpis_frame_eval", "gpis_reeval",
    "gpis_set_train_mode", "gpis_train_kick", "gpis_train_wait",
]


class Config(C.Structure):
    _fields_ = [
        ("dim", C.c_int32), ("device", C.c_int32), ("map_scale", C.c_float), ("map_noise", C.c_float),
        ("cluster_half", C.c_float), ("search_half", C.c_float), ("var_thre", C.c_float),
        ("obs_scale", C.c_float), ("obs_noise", C.c_float), ("max_leaves", C.c_int32), ("reserved0", C.c_int32),
        ("arena_chunk_bytes", C.c_uint64),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("leaves", C.c_int64), ("leaves_trained", C.c_int64), ("arena_bytes_used", C.c_int64),
        ("arena_bytes_reserved", C.c_int64),
        ("last_train_leaves", C.c_int64), ("last_train_sum_N", C.c_int64), ("last_train_sum_n", C.c_int64),
        ("last_train_flops", C.c_double), ("last_train_bytes", C.c_double), ("last_train_ms", C.c_float),
        ("last_train_skipped", C.c_int32),
        ("last_query_n", C.c_int64), ("last_query_evals", C.c_int64),
        ("last_query_flops", C.c_double), ("last_query_bytes_gather", C.c_double),
        ("last_query_bytes_compulsory", C.c_double), ("last_query_ms", C.c_float), ("last_query_eval_ms", C.c_float),
        ("kernel_launches", C.c_int64), ("last_query_items", C.c_int64 * 4),
        ("last_replicate_bytes", C.c_int64), ("last_replicate_records", C.c_int64), ("last_replicate_ms", C.c_float),
        ("reserved1", C.c_int32),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libgpis_b200.so is not built (python -m gpismap_b200.build); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        vp, i32, i64, f = C.c_void_p, C.c_int32, C.c_int64, C.c_float
        L.gpis_config_default.argtypes = [C.POINTER(Config), C.c_int]
        L.gpis_create.argtypes = [C.POINTER(vp), C.POINTER(Config)]
        L.gpis_destroy.argtypes = [vp]
        L.gpis_destroy.restype = None
        L.gpis_reset.argtypes = [vp]
        L.gpis_last_error.argtypes = [vp]
        L.gpis_last_error.restype = C.c_char_p
        L.gpis_device.argtypes = [vp]
        L.gpis_leaves_update.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp]
        L.gpis_leaves_mark.argtypes = [vp, C.c_int, vp, vp]
        L.gpis_leaves_erase.argtypes = [vp, C.c_int, vp]
        L.gpis_leaves_set_boxes.argtypes = [vp, C.c_int, vp, vp]
        L.gpis_rebase.argtypes = [vp, vp, C.c_int]
        L.gpis_get_rebase.argtypes = [vp, vp, C.POINTER(C.c_int)]
        L.gpis_leaf_get.argtypes = [vp, vp, vp, vp, vp, vp, vp, C.c_int]
        L.gpis_query.argtypes = [vp, vp, i64, vp]
        L.gpis_query_device.argtypes = [vp, vp, i64, vp]
        L.gpis_query_debug.argtypes = [vp, vp, i64, vp, vp, vp]
        L.gpis_leaf_index.argtypes = [vp, vp]
        L.gpis_obs_train_2d.argtypes = [vp, vp, vp, C.c_int, C.c_int]
        L.gpis_obs_train_1d.argtypes = [vp, vp, vp, C.c_int]
        L.gpis_obs_test.argtypes = [vp, vp, C.c_int, C.c_int, vp, vp]
        L.gpis_get_stats.argtypes = [vp, C.POINTER(Stats)]
        L.gpis_set_eval_version.argtypes = [vp, C.c_int]
        L.gpis_comm_unique_id.argtypes = [vp]
        L.gpis_comm_init.argtypes = [vp, C.c_int, C.c_int, vp]
        L.gpis_replicate.argtypes = [vp, C.c_int]
        L.gpis_snapshot_save.argtypes = [vp, C.c_char_p]
        L.gpis_snapshot_load.argtypes = [vp, C.c_char_p]
        L.gpis_set_train_mode.argtypes = [vp, C.c_int]
        L.gpis_train_wait.argtypes = [vp]
        L.gpis_train_kick.argtypes = [vp]
        _lib = L
    return _lib


def default_config(dim, device=0):
    cfg = Config()
    rc = lib().gpis_config_default(C.byref(cfg), dim)
    assert rc == 0
    cfg.device = device
    return cfg


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def cells_of(centres, cluster_half):
    """Integer lattice cell of leaf centres: floor(c / (2*half)); centres are odd multiples of half."""
    c = np.asarray(centres, np.float64)
    return np.floor(c / (2.0 * float(cluster_half))).astype(np.int32)


class Ctx:
    """Thin owner of a gpis_ctx."""

    def __init__(self, dim=3, device=0, cfg=None, borrowed=None):
        self.cfg = cfg if cfg is not None else default_config(dim, device)
        self.dim = self.cfg.dim
        self.borrowed = borrowed is not None
        if self.borrowed:      # a gpis_ctx* owned by someone else (e.g. a GPisMap3): never destroyed here
            self.h = C.c_void_p(borrowed)
            return
        self.h = C.c_void_p()
        rc = lib().gpis_create(C.byref(self.h), C.byref(self.cfg))
        if rc != 0:
            msg = lib().gpis_last_error(self.h).decode() if self.h else ""
            raise RuntimeError(f"gpis_create failed ({rc}): {msg} — no CPU fallback exists")

    def _ck(self, rc):
        if rc != 0:
            raise RuntimeError(f"gpis error {rc}: {lib().gpis_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None):
            if not self.borrowed and lib is not None:      # `lib` is gone during interpreter shutdown
                lib().gpis_destroy(self.h)
            self.h = None

    __del__ = close

    def reset(self):
        self._ck(lib().gpis_reset(self.h))

    def rebase(self, root_min_cell, levels):
        r = np.ascontiguousarray(root_min_cell, np.int32)
        self._ck(lib().gpis_rebase(self.h, _p(r), levels))

    def get_rebase(self):
        r = np.zeros(3, np.int32)
        lv = C.c_int(0)
        self._ck(lib().gpis_get_rebase(self.h, _p(r), C.byref(lv)))
        return r, lv.value

    def leaves_update(self, cells, centres, offsets, samples):
        cells = np.ascontiguousarray(cells, np.int32)
        centres = np.ascontiguousarray(centres, np.float32)
        offsets = np.ascontiguousarray(offsets, np.int32)
        samples = np.ascontiguousarray(samples, np.float32)
        n = cells.shape[0]
        status = np.zeros(max(n, 1), np.int32)
        self._ck(lib().gpis_leaves_update(self.h, n, _p(cells), _p(centres), _p(offsets), _p(samples), _p(status)))
        return status[:n]

    def leaves_mark(self, cells, centres):
        cells = np.ascontiguousarray(cells, np.int32)
        centres = np.ascontiguousarray(centres, np.float32)
        self._ck(lib().gpis_leaves_mark(self.h, cells.shape[0], _p(cells), _p(centres)))

    def leaves_set_boxes(self, cells, boxes):
        cells = np.ascontiguousarray(cells, np.int32)
        boxes = np.ascontiguousarray(boxes, np.float32)
        self._ck(lib().gpis_leaves_set_boxes(self.h, cells.shape[0], _p(cells), _p(boxes)))

    def leaves_erase(self, cells):
        cells = np.ascontiguousarray(cells, np.int32)
        self._ck(lib().gpis_leaves_erase(self.h, cells.shape[0], _p(cells)))

    def leaf_index(self, cell):
        c = np.ascontiguousarray(cell, np.int32)
        return lib().gpis_leaf_index(self.h, _p(c))

    def leaf_get(self, cell, want_L=True):
        c = np.ascontiguousarray(cell, np.int32)
        N = C.c_int32(0)
        ng = C.c_int32(0)
        n = lib().gpis_leaf_get(self.h, _p(c), C.byref(N), C.byref(ng), None, None, None, 0)
        if n == 0:
            return None
        alpha = np.zeros(n, np.float32)
        L = np.zeros((n, n), np.float32) if want_L else None
        gf = np.zeros(N.value, np.float32)
        lib().gpis_leaf_get(self.h, _p(c), C.byref(N), C.byref(ng), _p(alpha), _p(L), _p(gf), n)
        return dict(N=N.value, ng=ng.value, n=n, alpha=alpha, L=L, gradflag=gf)

    def query(self, x, res=None, debug=False):
        x = np.ascontiguousarray(x, np.float32)
        n = x.shape[0]
        w = 2 * (1 + self.dim)
        if res is None:
            res = np.zeros((n, w), np.float32)
        assert res.dtype == np.float32 and res.flags.c_contiguous and res.shape == (n, w)
        if debug:
            chosen = np.zeros((n, 4), np.int32)
            tie = np.zeros(n, np.int32)
            self._ck(lib().gpis_query_debug(self.h, _p(x), n, _p(res), _p(chosen), _p(tie)))
            return res, chosen, tie
        self._ck(lib().gpis_query(self.h, _p(x), n, _p(res)))
        return res

    def query_device(self, x_ptr, n, res_ptr):
        self._ck(lib().gpis_query_device(self.h, C.c_void_p(x_ptr), n, C.c_void_p(res_ptr)))

    def obs_train_2d(self, vu, zinv, ni, nj):
        vu = np.ascontiguousarray(vu, np.float32)
        zinv = np.ascontiguousarray(zinv, np.float32)
        self._ck(lib().gpis_obs_train_2d(self.h, _p(vu), _p(zinv), ni, nj))

    def obs_train_1d(self, theta, f):
        theta = np.ascontiguousarray(theta, np.float32)
        f = np.ascontiguousarray(f, np.float32)
        self._ck(lib().gpis_obs_train_1d(self.h, _p(theta), _p(f), theta.size))

    def obs_test(self, xt, d, val=None, var=None):
        xt = np.ascontiguousarray(xt, np.float32)
        m = xt.size // d
        val = np.zeros(m, np.float32) if val is None else np.ascontiguousarray(val, np.float32).copy()
        var = np.zeros(m, np.float32) if var is None else np.ascontiguousarray(var, np.float32).copy()
        self._ck(lib().gpis_obs_test(self.h, _p(xt), d, m, _p(val), _p(var)))
        return val, var

    def set_train_mode(self, mode):
        self._ck(lib().gpis_set_train_mode(self.h, mode))

    def train_wait(self):
        self._ck(lib().gpis_train_wait(self.h))

    def stats(self):
        s = Stats()
        self._ck(lib().gpis_get_stats(self.h, C.byref(s)))
        return {k: (list(getattr(s, k)) if k == "last_query_items" else getattr(s, k)) for k, _ in Stats._fields_}

    @staticmethod
    def comm_unique_id():
        buf = (C.c_ubyte * 128)()
        rc = lib().gpis_comm_unique_id(buf)
        if rc != 0:
            raise RuntimeError(f"gpis_comm_unique_id failed ({rc}): NCCL not available")
        return bytes(buf)

    def comm_init(self, rank, world, id128):
        buf = (C.c_ubyte * 128).from_buffer_copy(id128)
        self._ck(lib().gpis_comm_init(self.h, rank, world, buf))

    def replicate(self, root=0):
        self._ck(lib().gpis_replicate(self.h, root))

    def snapshot_save(self, path):
        self._ck(lib().gpis_snapshot_save(self.h, os.fsencode(path)))

    def snapshot_load(self, path):
        self._ck(lib().gpis_snapshot_load(self.h, os.fsencode(path)))

    def set_eval_version(self, v):
        self._ck(lib().gpis_set_eval_version(self.h, v))
