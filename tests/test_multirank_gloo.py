"""world_size-2 logic of the bench's multi-GPU path, on CPU with gloo: the z-plane sharding covers the
grid exactly once, and a replica that received the leader's samples over a broadcast answers its
shard bit-identically to a single-rank run (mock C ABI = oracle)."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, mocklib, outdir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from gpismap_b200 import hostapi, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "map3d.npz")))
    n = torch.zeros(1, dtype=torch.int64)
    if rank == 0:
        payload = torch.from_numpy(g["samples_in"].copy())
        n[0] = payload.shape[0]
    dist.broadcast(n, 0)
    if rank != 0:
        payload = torch.zeros((int(n[0]), 9), dtype=torch.float32)
    dist.broadcast(payload, 0)
    m = hostapi.GPisMap3(libpath=mocklib)
    m.insert_samples(payload.numpy())
    m.train_active()
    G = 12
    planes = list(range(rank, G, world))
    X = np.concatenate([synth.query_grid(G, lo=np.array([-0.1, -0.1, -0.1]), hi=np.array([0.2, 0.2, 0.2]), inflate=0.0,
                                         z_slab=(z, z + 1)) for z in planes])
    rows = m.test(X)
    np.save(os.path.join(outdir, f"rows{rank}.npy"), rows)
    np.save(os.path.join(outdir, f"planes{rank}.npy"), np.array(planes))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import mockbuild
    from gpismap_b200 import hostapi, synth
    mock = mockbuild.build()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, mock, str(tmp_path)), nprocs=2, join=True)
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "map3d.npz")))
    m = hostapi.GPisMap3(libpath=mock)
    m.insert_samples(g["samples_in"])
    m.train_active()
    G = 12
    full = m.test(synth.query_grid(G, lo=np.array([-0.1, -0.1, -0.1]), hi=np.array([0.2, 0.2, 0.2]), inflate=0.0)).reshape(G, G * G, 8)
    seen = []
    for r in range(2):
        rows = np.load(os.path.join(str(tmp_path), f"rows{r}.npy")).reshape(-1, G * G, 8)
        planes = np.load(os.path.join(str(tmp_path), f"planes{r}.npy"))
        seen += planes.tolist()
        assert np.array_equal(rows, full[planes])
    assert sorted(seen) == list(range(G))
