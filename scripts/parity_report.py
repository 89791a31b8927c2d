"""Development aid (GPU): row-error statistics of the CUDA path against the reference fixtures for both training
kernels (gpis_set_train_version 1 / 2): tests/golden/{map2d,map3d,seq3d}.npz."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
from gpismap_b200 import cabi, hostapi
G = os.path.join(ROOT, "tests", "golden")


def rep(rows, ref, dim, label):
    w = 1 + dim
    ev = (ref[:, w] < 1.0) & (rows[:, w] < 1.0)
    ef, eg, evr = H._errs(rows[ev], ref[ev], dim)
    out = [label, int(ev.sum())]
    for name, e, tol in (("f", ef, 1e-4), ("grad", eg, 1e-4), ("var", evr, 1e-3)):
        out.append(f"{name}: med {np.median(e):.2e} p99 {np.percentile(e, 99):.2e} max {e.max():.2e} within {100 * (e < tol).mean():.2f}%")
    print(" | ".join(str(x) for x in out), flush=True)


def fixture_map(name, dim, tv):
    g = dict(np.load(os.path.join(G, name)))
    P = H.P3 if dim == 3 else H.P2
    ctx = cabi.Ctx(dim)
    pitch = 2.0 * np.float64(np.float32(P["half"]))
    root_min = np.round((g["root_c"].astype(np.float64) - float(g["root_half"])) / pitch).astype(np.int32)
    levels = int(round(np.log2(float(g["root_half"]) / np.float64(np.float32(P["half"])))))
    ctx.rebase(root_min, levels)
    cells = cabi.cells_of(g["centres"], P["half"])
    ctx.leaves_update(cells, g["centres"], g["offsets"], g["samples"])
    ctx.leaves_set_boxes(cells, g["boxes"])
    got = ctx.query(g["X"], g["init"].copy())
    rep(got, g["rows"], dim, f"{name} train v{tv}")
    ctx.close()


def seq3d(tv):
    g = dict(np.load(os.path.join(G, "seq3d.npz")))
    m = None
    for k in range(len(g["cam"])):
        cam = int(g["cam"][k])
        c = tuple(np.float32(H.BIGBIRD_CAMS[n][cam - 1]) for n in ("fx", "fy", "cx", "cy")) + (640, 480)
        if m is None:
            m = hostapi.GPisMap3(cam=c)
                    else:
            m.resetCam(*c)
        dz = np.zeros(640 * 480, np.float32)
        a, b = g["depth_off"][k], g["depth_off"][k + 1]
        dz[g["depth_idx"][a:b]] = g["depth_val"][a:b]
        m.update(dz, g["pose12"][k])
    rep(m.test(g["X"]), g["rows"], 3, f"seq3d train v{tv}")
    m.close()


for tv in (1,):
    fixture_map("map2d.npz", 2, tv)
    fixture_map("map3d.npz", 3, tv)
    seq3d(tv)
