"""cProfile of rank 0 in a real torch.distributed run of the sharded slide path (overhead hunting)"""
import sys, os, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from cdnet_b200 import sharded
from tools.run_slide import build_rows
H, W = int(sys.argv[1]), int(sys.argv[2])
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
comm, be = sharded.DistComm(), sharded.CudaBackend()
r0, r1 = sharded.row_partition(H, world)[rank]
rows = build_rows(r0, r1, W, 1)
d = sharded.alloc_shard_buffers(be, rank, world, H, W, 1)
for k, v in rows.items():
    d[k].copy_(be.to_dev(v))
for _ in range(3):
    sharded.postprocess_slide([d], comm, H, W, be)
torch.cuda.synchronize(); dist.barrier()
pr = cProfile.Profile(); pr.enable()
t0 = time.perf_counter()
for _ in range(5):
    sharded.postprocess_slide([d], comm, H, W, be)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 5
pr.disable()
if rank == 0:
    print("ms per slide", 1e3 * dt)
    pstats.Stats(pr).sort_stats("tottime").print_stats(18)
dist.barrier(); dist.destroy_process_group()
