"""Whole-slide postproc = 1 (watershed) timing, one process per GPU:
    python tools/slide_ws.py [H W [overlap]]            or under torchrun for N > 1
Prints the tools/bench_configs.whole_slide record (verified against the unsharded call first) as one JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from tools import bench_configs as BC  # noqa: E402


def main():
    H = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
    W = int(sys.argv[2]) if len(sys.argv) > 2 else 40000
    overlap = int(sys.argv[3]) if len(sys.argv) > 3 else 128
    postproc = int(os.environ.get("SLIDE_POSTPROC", "1"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        dist.init_process_group("nccl")
    try:
        out = BC.whole_slide(torch, dist, rank, world, H=H, W=W, steps=2, postproc=postproc, overlap=overlap)
        if rank == 0:
            print(json.dumps(out), flush=True)
    finally:
        if world > 1:
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
