"""phase timing of the sharded slide path with all ranks simulated on one GPU (host-logic overhead)"""
import sys, os, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cdnet_b200 import sharded
from tools.run_slide import build_rows
H = W = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
G = int(sys.argv[2]) if len(sys.argv) > 2 else 8
be = sharded.CudaBackend()
parts = sharded.row_partition(H, G)
shards = [{k: be.to_dev(v) for k, v in build_rows(a, b, W, 1).items()} for a, b in parts]
comm = sharded.SimComm(G)
sharded.postprocess_slide(shards, comm, H, W, be)
torch.cuda.synchronize()
t0 = time.perf_counter()
sharded.postprocess_slide(shards, comm, H, W, be)
torch.cuda.synchronize()
print("total ms", 1e3 * (time.perf_counter() - t0))
pr = cProfile.Profile(); pr.enable()
sharded.postprocess_slide(shards, comm, H, W, be)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
