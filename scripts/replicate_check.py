"""K5 check on >= 2 GPUs (torchrun, one rank per GPU): rank 0 maps synthetic frames through the drop-in GPisMap3 and
calls gpis_replicate after every frame; the other ranks follow. After every frame all ranks answer the same query
points and the rows must be bit-identical to rank 0's (the map changes incrementally: retrained leaves, erased leaves,
effective boxes, root box). Prints REPLICATE_OK on rank 0.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/replicate_check.py [frames]
"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from gpismap_b200 import cabi, hostapi, synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 4
gmap = hostapi.GPisMap3(device=local) if rank == 0 else None
ctx = cabi.Ctx(3, local, borrowed=gmap.ctx_handle()) if rank == 0 else cabi.Ctx(3, local)
ids = [cabi.Ctx.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, 0)
ctx.comm_init(rank, world, ids[0])
X = synth.query_grid(96)
ok = True
for k in range(nf):
    if rank == 0:
        dz, pose = synth.frame(k, 40)
        gmap.update(dz, pose)
    dist.barrier()          # the timing below must not include waiting for rank 0's update
    ctx.replicate(0)
    st = ctx.stats()
    rows = ctx.query(X)
    t = torch.from_numpy(rows).cuda()
    ref = t.clone()
    dist.broadcast(ref, 0)
    same = bool(torch.equal(ref.view(torch.int32), t.view(torch.int32)))
    flag = torch.tensor([1 if same else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    ms = torch.tensor([st["last_replicate_ms"]], device="cuda", dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"frame {k}: {st['last_replicate_records']} records, {st['last_replicate_bytes'] / 1e6:.1f} MB, {ms.item():.2f} ms (max over ranks), "
              f"replicas bit-identical: {bool(flag.item())}, evaluated rows {(rows[:, 4] < 1).sum()}", flush=True)
    ok = ok and bool(flag.item())
if rank == 0:
    print("REPLICATE_OK" if ok else "REPLICATE_MISMATCH", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
