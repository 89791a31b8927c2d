"""Where does the HBM term of the query path bind? (SURVEY.md 8d asks for both regimes to be reported.)
Two workloads with small or no leaf work per query, device-resident inputs, CUDA-event times from the library:
  (a) the 2-D demo map (77 leaves, n <= 225) queried on a 4096 x 4096 grid;
  (b) the 3-D 40-frame map queried on a 256^3 grid translated outside the room: every query is an empty-space query
      (candidate search + var_f preset only, 44 B of compulsory traffic per query).
Prints one JSON line."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gpismap_b200 import cabi, hostapi, synth

peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
hbm = float(peaks.get("hbm_gbs", 6650.0))
fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12
out = {"hbm_peak_gbs": hbm, "fp32_peak_tflops": fp32_peak}


def run(ctx, x, w, reps=3):
    xd = torch.from_numpy(x).cuda()
    rd = torch.zeros((x.shape[0], w), dtype=torch.float32, device="cuda")
    best = None
    for _ in range(reps + 1):
        rd.zero_()
        torch.cuda.synchronize()
        ctx.query_device(xd.data_ptr(), x.shape[0], rd.data_ptr())
        s = ctx.stats()
        if best is None or s["last_query_ms"] < best["last_query_ms"]:
            best = s
    ev = int((rd[:, w // 2] < 1.0).sum().item())
    return best, ev


# (a) 2-D demo map
g = dict(np.load(os.path.join(ROOT, "tests", "golden", "seq2d_demo.npz")))
m2 = hostapi.GPisMap()
for i in range(g["ranges"].shape[0]):
    m2.update(g["thetas"], g["ranges"][i], g["pose6"][i])
c2 = cabi.Ctx(2, 0, borrowed=m2.ctx_handle())
xs = np.linspace(-5, 20, 4096, dtype=np.float32) + np.float32(0.0137)
ys = np.linspace(-15, 5, 4096, dtype=np.float32) + np.float32(0.0071)
X2 = np.stack(np.meshgrid(xs, ys, indexing="ij"), -1).reshape(-1, 2).astype(np.float32)
s, ev = run(c2, X2, 6)
sec = s["last_query_ms"] * 1e-3
out["map2d_4096x4096"] = {
    "queries": int(X2.shape[0]), "evaluated": ev, "leaves": s["leaves"], "ms": s["last_query_ms"], "eval_kernel_ms": s["last_query_eval_ms"],
    "queries_per_s": X2.shape[0] / sec,
    "hbm_term": {"compulsory_bytes": s["last_query_bytes_compulsory"], "gbs": s["last_query_bytes_compulsory"] / sec / 1e9,
                 "frac_of_hbm_peak": s["last_query_bytes_compulsory"] / sec / 1e9 / hbm},
    "fp32_term": {"flops": s["last_query_flops"], "tflops": s["last_query_flops"] / sec / 1e12, "frac_of_fp32_peak": s["last_query_flops"] / sec / 1e12 / fp32_peak},
}
m2.close()
# (b) empty-space queries against the 3-D map
m3 = hostapi.GPisMap3()
for k in range(40):
    dz, pose = synth.frame(k, 40)
    m3.update(dz, pose)
c3 = cabi.Ctx(3, 0, borrowed=m3.ctx_handle())
X3 = synth.query_grid(256) + np.array([10.0, 0.0, 0.0], np.float32)
s, ev = run(c3, X3, 8)
sec = s["last_query_ms"] * 1e-3
out["map3d_empty_space_256^3"] = {
    "queries": int(X3.shape[0]), "evaluated": ev, "leaves": s["leaves"], "ms": s["last_query_ms"], "queries_per_s": X3.shape[0] / sec,
    "hbm_term": {"compulsory_bytes": 44.0 * X3.shape[0], "gbs": 44.0 * X3.shape[0] / sec / 1e9, "frac_of_hbm_peak": 44.0 * X3.shape[0] / sec / 1e9 / hbm},
    "note": "125 hash probes per query (5^3 lattice cells around the query box) against the L2-resident table: bound by probe latency, not by HBM",
}
m3.close()
print(json.dumps(out))
