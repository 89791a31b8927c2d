"""BASELINE configs[0]: the reference's whole 2-D demo run (28 laser scans of data/2D/gazebo1.mat, committed as
tests/golden/seq2d_demo.npz) through GPisMap::update, then a 0.05 m test grid over the demo's window through
GPisMap::test — the CUDA path next to the unmodified reference (oracle/_ref) on the box's host cores.
Writes one JSON object (profiles/rNN_config0_2d.json is a copy of it)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpismap_b200 import hostapi
g = dict(np.load(os.path.join(ROOT, "tests", "golden", "seq2d_demo.npz")))
step = float(sys.argv[1]) if len(sys.argv) > 1 else 0.05
xs = np.arange(-5 + step, 20 - step + 1e-9, step); ys = np.arange(-15 + step, 5 - step + 1e-9, step)
xg, yg = np.meshgrid(xs, ys)
X = np.stack([xg.T.ravel(), yg.T.ravel()], 1).astype(np.float32)

def run(m):
    ups = []
    for i in range(g["ranges"].shape[0]):
        t0 = time.perf_counter(); m.update(g["thetas"], g["ranges"][i], g["pose6"][i]); ups.append((time.perf_counter() - t0) * 1e3)
    m.test(X[:1000])
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); rows = m.test(X); ts.append(time.perf_counter() - t0)
    return ups, min(ts), rows

out = {"workload": f"2-D demo: 28 scans x 270 beams, test grid {step} m over [-5,20]x[-15,5] ({len(X)} queries)"}
m = hostapi.GPisMap()
ups, tq, rows = run(m)
out["b200"] = {"update_ms_median": float(np.median(ups[1:])), "update_ms_first": ups[0], "queries_per_s": len(X) / tq, "query_ms": tq * 1e3,
               "leaves": int(m.leaves()[0].shape[0]), "evaluated": int((rows[:, 3] < 1.0).sum())}
m.close()
try:
    from oracle import refpy
    if refpy.available():
        r = refpy.RefMap2()
        ups, tq, rrows = run(r)
        out["reference"] = {"update_ms_median": float(np.median(ups[1:])), "queries_per_s": len(X) / tq, "query_ms": tq * 1e3,
                            "cores": int(refpy.lib().ref_hardware_concurrency()), "leaves": int(r.clusters()[0].shape[0])}
        ev = (rrows[:, 3] < 1.0) & (rows[:, 3] < 1.0)
        out["agreement"] = {"same_evaluated_mask": bool(np.array_equal(rrows[:, 3] < 1.0, rows[:, 3] < 1.0)),
                            "max_abs_f_diff": float(np.abs(rows[ev, 0] - rrows[ev, 0]).max()),
                            "note": "the GPU observation GP rounds differently from the CPU one, so stored samples differ in the last bits; strict parity is checked at leaf level (tests/)"}
except Exception as e:
    out["reference"] = {"error": repr(e)}
print(json.dumps(out))
