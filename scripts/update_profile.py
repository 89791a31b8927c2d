"""Development aid: per-phase timing of GPisMap3::update over the bench's synthetic frames."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpismap_b200 import hostapi, synth
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 40
m = hostapi.GPisMap3()
rows = []
frames = [synth.frame(k, nf) for k in range(nf)]   # generated up front: the updates below run back to back
for k in range(nf):
    if k == 2: t_all = time.perf_counter()      # frame 0 creates the context, the first training launch loads K1's module
    dz, pose = frames[k]
    t0 = time.perf_counter(); m.update(dz, pose); dt = time.perf_counter() - t0
    ph, cnt, ms = m.timing()
    rows.append(list(ph * 1e3) + [dt * 1e3, ms] + list(cnt))
from gpismap_b200 import cabi
cabi.Ctx(3, 0, borrowed=m.ctx_handle()).train_wait()   # the last batch may still be in flight (gpis_set_train_mode)
t_all = time.perf_counter() - t_all
print(f"sustained: {1e3 * t_all / (nf - 2):.2f} ms per frame over frames 2..{nf - 1}, back to back, incl. the final training wait "
      f"(GPIS_TRAIN_MODE={os.environ.get('GPIS_TRAIN_MODE', 'default 3')})")
import ctypes as C
L = hostapi.lib()
secs = np.zeros(16); calls = np.zeros(16, np.int64)
L.gm_profile(secs.ctypes.data_as(C.c_void_p), calls.ctypes.data_as(C.c_void_p))
a = np.array(rows)
names = ["preproc", "regressObs", "updateMapPoints", "addNewMeas+eval", "trainActive", "total", "train_kernel", "valid_px", "active", "trained"]
print("median / p90 / max over", nf, "frames (ms)")
for i, n in enumerate(names):
    print(f"{n:18s} {np.median(a[:, i]):10.2f} {np.percentile(a[:, i], 90):10.2f} {a[:, i].max():10.2f}")
if os.environ.get("UPDATE_PROFILE_ROWS"):
    print("per frame: " + " | ".join(names))
    for k, r in enumerate(a): print(k, " ".join(f"{v:8.2f}" for v in r))
gap = a[:, 5] - a[:, :5].sum(1)
print(f"{'total - phases':18s} {np.median(gap):10.2f} {np.percentile(gap, 90):10.2f} {gap.max():10.2f}")

pn = ["gpis_obs_test / gpis_reeval", "uMP: cull", "uMP: collect + gpis_reeval", "uMP: serial apply (incl on-demand obs tests)", "evalPoints: serial insert", "train: dirty set (host path)", "train: sample lists (gpis_samples_set)", "gpis_leaves_train_dirty / _update", "sync_table"]
print("host profile, ms per frame (calls per frame)")
for i, n in enumerate(pn):
    print(f"{n:32s} {secs[i] * 1e3 / nf:9.2f}  ({calls[i] / nf:.1f})")
