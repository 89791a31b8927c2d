"""Development aid: per-warp phase timing of one CTA of k_eval_v3 (library built with -DE3_TIMING=<block>)."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpismap_b200 import cabi, hostapi, synth
m = hostapi.GPisMap3()
for k in range(int(sys.argv[1]) if len(sys.argv) > 1 else 12):
    dz, pose = synth.frame(k, 40); m.update(dz, pose)
ctx = cabi.Ctx(3, 0, borrowed=m.ctx_handle())
X = synth.query_grid(96)
L = cabi.lib()
buf = np.zeros(64 * 32, np.int64)
L.gpis_debug_timing(buf.ctypes.data_as(C.c_void_p))
ctx.query(X)
L.gpis_debug_timing(buf.ctypes.data_as(C.c_void_p))
t = buf.reshape(64, 32)[:8]
names = ["kstar+mean", "wait U_j", "issue+wait tiles", "fma", "elim total", "end barrier", "variance+out", "nb", "sum cnt", "steps"]
print("evals", ctx.stats()["last_query_evals"], "eval ms", ctx.stats()["last_query_eval_ms"])
for i, n in enumerate(names):
    print(f"{n:18s}", " ".join(f"{int(v):9d}" for v in t[:, i]))

print("FMA cycles per quarter-step by active slots (1..4), and per slot:")
for c in range(1, 5):
    cyc = t[:, 16 + c].sum(); cnt = t[:, 24 + c].sum()
    if cnt: print(f"  cnt={c}: steps {int(cnt):6d}  cycles/step {cyc / cnt:8.1f}  cycles/slot {cyc / cnt / c:8.1f}")
