"""Development aid: a query placed exactly on a sample (r = 0 makes kf2 divide by zero in the reference,
covFnc.cpp:29-33): compare the NaN pattern of the CUDA path with the oracle's."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
from gpismap_b200 import cabi
from oracle import oraclepy
O = oraclepy.Oracle()
rng = np.random.default_rng(2)
P = H.P3
ctx = cabi.Ctx(3)
cells = np.array([[6, -3, 1]], np.int32)
centres = ((2 * cells + 1) * np.float64(np.float32(P["half"]))).astype(np.float32)
s = H.leaf_samples3(60, rng, spread=0.03)
s[:, :3] += centres[0] - np.array([0.3125, -0.1375, 0.0625], np.float32)
ctx.leaves_update(cells, centres, [0, len(s)], s)
x = np.concatenate([s[:5, :3], s[5:8, :3] + np.float32(1e-3)]).astype(np.float32)
got = ctx.query(x)
gp = O.gp_train(3, s, P["scale"], P["noise"])
want = O.make_map(3, centres, P["half"], [gp], P["search"], P["var_thre"], P["noise"]).test(x)
np.set_printoptions(precision=4, linewidth=200)
print("got\n", got); print("want\n", want)
print("nan pattern equal:", np.array_equal(np.isnan(got), np.isnan(want)))
