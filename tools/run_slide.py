#!/usr/bin/env python
"""Whole-slide post-processing (BASELINE configs[4]) row-sharded over the GPUs of one box.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/run_slide.py --H 40000 --W 40000 [--maps 1] [--steps 3] [--verify]

Every rank builds its own rows of a synthetic prediction map (a few seeded 1000x1000 MoNuSeg-shaped
tiles repeated over the shard), runs cdnet_b200.sharded.postprocess_slide (halo rows, two scalar
reductions and the seam union over torch.distributed/NCCL) and rank 0 prints one JSON line; time = max
over ranks of the CUDA-event time.  --verify (small slides): rank 0 also post-processes the whole slide
on its own GPU and every rank's rows must match bit for bit."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def build_rows(r0, r1, W, n_maps, tile=1000, n_seeds=4):
    from cdnet_b200 import synth
    base = [synth.postproc_inputs(100 + i, tile, tile) for i in range(n_seeds)]
    Hl = r1 - r0
    dcm = np.empty((n_maps, Hl, W), np.uint8)
    prob = np.empty((3, Hl, W), np.float32)
    point = np.empty((1, Hl, W), np.float32)
    for y in range(r0 - r0 % tile, r1, tile):
        for x in range(0, W, tile):
            t = base[((y // tile) * 7 + (x // tile) * 3) % n_seeds]
            ya, yb = max(y, r0), min(y + tile, r1)
            xb = min(x + tile, W)
            sl = (slice(ya - y, yb - y), slice(0, xb - x))
            dcm[:, ya - r0:yb - r0, x:xb] = t["dcm"][:n_maps][(slice(None),) + sl]
            prob[:, ya - r0:yb - r0, x:xb] = t["prob"][(slice(None),) + sl]
            point[:, ya - r0:yb - r0, x:xb] = t["point"][(slice(None),) + sl]
    return dict(dcm=dcm, prob=prob, point=point)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--H", type=int, default=8000)
    ap.add_argument("--W", type=int, default=8000)
    ap.add_argument("--maps", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--verify", action="store_true")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from cdnet_b200 import sharded, api
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl" if world > 1 else "gloo", rank=rank, world_size=world,
                            **({"device_id": torch.device("cuda", local)} if world > 1 else {}))
    comm = sharded.DistComm()
    be = sharded.CudaBackend()
    r0, r1 = sharded.row_partition(a.H, world)[rank]
    rows = build_rows(r0, r1, a.W, a.maps)
    dev_rows = sharded.alloc_shard_buffers(be, rank, world, a.H, a.W, a.maps)
    for k, v in rows.items():
        dev_rows[k].copy_(be.to_dev(v))
    del rows

    def step():
        return sharded.postprocess_slide([dev_rows], comm, a.H, a.W, be, 9, 20, 2)[0]

    out = step()  # warm-up
    torch.cuda.synchronize()
    dist.barrier()
    times = []
    for _ in range(a.steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0.record()
        out = step()
        e1.record()
        torch.cuda.synchronize()
        ms = max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0))
        t = torch.tensor([ms], device="cuda") if world > 1 else torch.tensor([ms])
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(float(t.item()))
    ok = None
    if a.verify:
        full = build_rows(0, a.H, a.W, a.maps) if rank == 0 else None
        if rank == 0:
            lab, _ = api.dam_postprocess_cuda(be.to_dev(full["dcm"])[None], be.to_dev(full["prob"])[None],
                                              be.to_dev(full["point"])[None], 9, 20, 2, 0)
            lab = lab[0].cpu().numpy()
        gathered = [None] * world
        dist.all_gather_object(gathered, out.cpu().numpy())
        if rank == 0:
            got = np.concatenate(gathered, axis=0)
            ok = bool(np.array_equal(got, lab))
    if rank == 0:
        ms = float(np.median(times))
        print(json.dumps({"what": "whole-slide DAM post-proc, row-sharded", "H": a.H, "W": a.W, "n_maps": a.maps,
                          "n_gpus": world, "ms_per_slide": ms, "mpx_per_s": a.H * a.W / 1e6 / (ms * 1e-3),
                          "times_ms": times, "verified_equal_to_single_gpu": ok,
                          "n_labels": int(out.max().item()) if world == 1 else None}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
