"""BASELINE configs[3] ("large-scale 3D: 1000 synthetic depth frames, batched per-frame leaf Cholesky throughput"):
drives GPisMap3::update over a 1000-frame random-walk trajectory through the box room (synth.walk_frame, seed 1) and
reports, per frame, the dirty leaves, the training kernel time and its Cholesky GFLOP/s, and how the map grows.
Prints one JSON line. The CPU reference cannot run this configuration in any reasonable time (9-16 s per frame, and
its dense factors — 4 n^2 bytes per leaf, 5,316 leaves x ~6.8 MB = 36 GB — exceed what the update loop leaves free);
see profiles/README.md.

    python scripts/config4.py [frames=1000]
"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpismap_b200 import cabi, hostapi, synth

nf = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
m = hostapi.GPisMap3()
ctx = cabi.Ctx(3, 0, borrowed=m.ctx_handle())
upd, trn, leaves, flops, sumn, nsamp, nleaf = [], [], [], [], [], [], []
t_all = time.time()
for k in range(nf):
    dz, pose = synth.walk_frame(k, nf)
    t0 = time.perf_counter()
    m.update(dz, pose)
    upd.append(1e3 * (time.perf_counter() - t0))
    st = ctx.stats()
    ph, cnt, ms = m.timing()
    trn.append(ms); leaves.append(int(cnt[2])); flops.append(st["last_train_flops"] if cnt[2] else 0.0)
    sumn.append(st["last_train_sum_n"] / max(1, st["last_train_leaves"]) if cnt[2] else 0.0)
    if k % 50 == 49 or k == nf - 1:
        nsamp.append(int(m.getAllPoints().shape[0])); nleaf.append(int(ctx.stats()["leaves"]))
# Leaf training overlaps the next frame (gpis_set_train_mode): timing() reports the batch that has completed, i.e. the
# previous frame's. Wait for the last one and shift the series so that trn[k] belongs to frame k.
t0 = time.perf_counter()
ctx.train_wait()
final_wait_ms = 1e3 * (time.perf_counter() - t0)
if int(os.environ.get("GPIS_TRAIN_MODE", "3")) != 0:
    trn = trn[1:] + [ctx.stats()["last_train_ms"]]
st = ctx.stats()
upd, trn, leaves, flops = map(np.asarray, (upd, trn, leaves, flops))
busy = trn > 0
out = {
    "config": f"BASELINE configs[3]: {nf} synthetic 640x480 depth frames, random-walk trajectory (seed 1), default leaf sizes",
    "frames": nf, "wall_s_total_incl_frame_synthesis": time.time() - t_all,
    "update_ms_per_frame": {"median": float(np.median(upd)), "p90": float(np.percentile(upd, 90)), "max": float(upd.max()),
                            "mean_incl_final_train_wait": float((upd.sum() + final_wait_ms) / nf)},
    "train_mode": int(os.environ.get("GPIS_TRAIN_MODE", "3")),
    "trained_leaves_per_frame": {"median": float(np.median(leaves)), "p90": float(np.percentile(leaves, 90)), "max": int(leaves.max()), "total": int(leaves.sum())},
    "train_kernel_ms_per_frame": {"median": float(np.median(trn)), "p90": float(np.percentile(trn, 90))},
    "dirty_leaves_per_s_of_training_kernel": float(leaves.sum() / (trn.sum() * 1e-3)),
    "dirty_leaves_per_s_of_update": float(leaves.sum() / (upd.sum() * 1e-3)),
    "cholesky_tflops_of_training_kernel": float(flops.sum() / (trn.sum() * 1e-3) / 1e12),
    "mean_unknowns_per_trained_leaf": float(np.mean(np.asarray(sumn)[busy])) if busy.any() else 0.0,
    "samples_every_50_frames": nsamp, "leaves_every_50_frames": nleaf,
    "final": {"samples_point_leaves": nsamp[-1], "leaves": st["leaves"], "arena_gb": st["arena_bytes_used"] / 1e9,
              "reference_dense_factor_gb_for_the_same_leaves": None},
}
print(json.dumps(out))
