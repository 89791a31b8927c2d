"""Development aid: histogram of leaf sizes (block rows) of the bench map."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpismap_b200 import cabi, hostapi, synth
m = hostapi.GPisMap3()
for k in range(40):
    dz, pose = synth.frame(k, 40); m.update(dz, pose)
ctx = cabi.Ctx(3, 0, borrowed=m.ctx_handle())
centres, counts = m.leaves()
half = 0.025
cells = np.floor(centres / (2 * half)).astype(np.int32)
ns = []
for c in cells:
    g = ctx.leaf_get(c, want_L=False)
    if g: ns.append(g["n"])
ns = np.array(ns); nb = (ns + 31) // 32
print("leaves", len(ns), "n: min/median/max", ns.min(), int(np.median(ns)), ns.max())
h = np.bincount(nb)
w = np.bincount(nb, weights=nb.astype(float) ** 2)   # share of the solve work
for b in range(len(h)):
    if h[b]: print(f"nb={b:3d} leaves {h[b]:5d}  cumulative {h[:b + 1].sum() / len(nb):.3f}  work share cumulative {w[:b + 1].sum() / w.sum():.3f}")
