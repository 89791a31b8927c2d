"""Development aid: a short GPisMap3 run (overlapped training, batched second generation, staged query copies) for
compute-sanitizer:  compute-sanitizer --tool memcheck python scripts/sanitize_update.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gpismap_b200 import hostapi, synth
m = hostapi.GPisMap3()
for k in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    dz, pose = synth.frame(k, 40)
    m.update(dz, pose)
S = m.all_samples()
X = (S[::29, :3] - S[::29, 3:6] * np.float32(0.01)).astype(np.float32)
xh = torch.from_numpy(X).pin_memory()
rh = torch.zeros((len(X), 8), dtype=torch.float32).pin_memory()
rows = m.test(xh.numpy(), rh.numpy())      # pinned buffers: the staged-copy path of gpis_query
rows2 = m.test(X)                          # pageable buffers
assert np.array_equal(rows, rows2, equal_nan=True)
print("ok", len(S), "samples,", int((rows[:, 4] < 0.3).sum()), "of", len(X), "queries evaluated")
m.close()
