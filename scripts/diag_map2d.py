"""Development aid (GPU): the rows of tests/golden/map2d.npz where the CUDA path is furthest from the reference."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
from gpismap_b200 import cabi
from oracle import oraclepy
g = dict(np.load(os.path.join(ROOT, "tests", "golden", "map2d.npz")))
P = H.P2
ctx = cabi.Ctx(2)
pitch = 2.0 * np.float64(np.float32(P["half"]))
root_min = np.round((g["root_c"].astype(np.float64) - float(g["root_half"])) / pitch).astype(np.int32)
levels = int(round(np.log2(float(g["root_half"]) / np.float64(np.float32(P["half"])))))
ctx.rebase(root_min, levels)
cells = cabi.cells_of(g["centres"], P["half"])
ctx.leaves_update(cells, g["centres"], g["offsets"], g["samples"])
ctx.leaves_set_boxes(cells, g["boxes"])
got, chosen, tie = ctx.query(g["X"], g["init"].copy(), debug=True)
O64 = oraclepy.Oracle(double=True, cov_float=True)
offs = g["offsets"]
gps = [O64.gp_train(2, g["samples"][offs[i]:offs[i + 1]], P["scale"], P["noise"]) for i in range(len(offs) - 1)]
m64 = O64.make_map(2, g["centres"], P["half"], gps, P["search"], P["var_thre"], P["noise"], boxes=g["boxes"])
w64 = m64.test(g["X"], g["init"].astype(np.float64))
ev = g["ncand"] > 0
ef, eg, evr = H._errs(got, g["rows"], 2)
er = H._errs(g["rows"], w64, 2)
idx = np.argsort(-np.where(ev, np.maximum(ef, eg), 0))[:8]
np.set_printoptions(precision=7, linewidth=200)
for i in idx:
    print("row", i, "ncand", g["ncand"][i], "err f/grad/var", ef[i], eg[i], evr[i], "ref-vs-64", er[0][i], er[1][i], er[2][i])
    print("   gpu", got[i]); print("   ref", g["rows"][i]); print("   f64", w64[i])
    # single-leaf evaluations of the three nearest leaves
    for k in (1, 2, 3):
        s = chosen[i, k]
        if s >= 0:
            li = [j for j in range(len(cells)) if ctx.leaf_index(cells[j]) == s][0]
            print("      leaf", li, "alone f64:", gps[li].test(g["X"][i:i + 1].astype(np.float64))[0])
