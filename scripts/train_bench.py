"""Development aid: leaf-training kernel time vs leaf size (batch of identical-size leaves, one launch)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpismap_b200 import cabi
import helpers
rng = np.random.default_rng(5)
nleaf = int(sys.argv[1]) if len(sys.argv) > 1 else 296
print("k_leaf_train (with the refinement step), leaves per launch", nleaf)
sizes = tuple(int(v) for v in os.environ.get("TB_SIZES", "48,96,144,200,260,330,400,480").split(","))
for N in sizes:
    ctx = cabi.Ctx(dim=3)
    base = helpers.leaf_samples3(N, rng, spread=0.045)
    cells = np.zeros((nleaf, 3), np.int32); cells[:, 0] = np.arange(nleaf) * 4
    centres = (cells.astype(np.float32) + 0.5) * 0.05
    samples = np.tile(base, (nleaf, 1)).astype(np.float32)
    for i in range(nleaf):
        samples[i * N:(i + 1) * N, :3] += centres[i] - np.array([0.3125, -0.1375, 0.0625], np.float32)
    offsets = (np.arange(nleaf + 1) * N).astype(np.int32)
    ts = []
    for rep in range(3):
        ctx.leaves_update(cells, centres, offsets, samples)
        st = ctx.stats(); ts.append(st["last_train_ms"])
    n = st["last_train_sum_n"] // nleaf
    fl = st["last_train_flops"]
    if hasattr(cabi.lib(), "gpis_debug_train_timing"):
        import ctypes as C
        tb = np.zeros(8, np.int64); cabi.lib().gpis_debug_train_timing(tb.ctypes.data_as(C.c_void_p))
        tot = max(1, tb[:5].sum())
        print("   phases A/B/C/D/E %:", " ".join(f"{100 * v / tot:5.1f}" for v in tb[:5]), " cycles/launch", int(tot // 3))
    print(f"N={N:4d} n={n:5d} nb={(n + 31) // 32:3d}  {min(ts):8.3f} ms  {fl / min(ts) / 1e9:8.1f} TFLOP/s")
    del ctx
