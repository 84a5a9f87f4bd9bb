#!/usr/bin/env python
"""bench.py -- headline benchmark of the CDNet geometry hot path on B200.

Workload (BASELINE.json configs[1]): CDNet inference post-processing (test_dam.py:455-563, the
reference's default postproc=0 branch) on 14 synthetic 1000x1000 MoNuSeg-shaped tiles per GPU:
8 TTA direction-argmax maps (uint8) + a 3-class probability map (f32) + a point map (f32) ->
instance label map.  A "step" = one pass over the 14 tiles.  Metric: Mpixel/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   (N > 1, weak scaling:
        every rank post-processes its own 14 tiles; no data-path collective)

`value`  : inputs resident in HBM, CUDA-event time over K steps, max over ranks.
`e2e`    : the same K steps through the host-buffer API (cdnet_b200.api.DamPostprocessPlan.run):
           pinned-host -> device copies of the inputs and device -> host copy of the labels
           inside the timed region.
`roofline`: per-kernel CUDA-event times collected by the library's own profiler (cdnet_profile_*)
           over K further steps; the dominant kernel's algorithmic bytes / its mean duration.
`cpu_baseline`: the oracle port of the reference's path (oracle/restate.py, literal per-instance
           loops, 1 process) on a bounded sample of the same tiles, rank 0, N=1 only.
`--impl reference`: the oracle port on all host cores (one tile per worker process per step).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

TILES, H, W = 14, 1000, 1000
DIRECTION_CLASSES, MIN_AREA, RADIUS, POSTPROC = 9, 20, 2, 0
ALG_BYTES_PER_PX = 28.0  # 8 u8 maps + 3 f32 + 1 f32 in, int32 labels out

# algorithmic bytes per pixel of each kernel at its own boundary (DESIGN.md section "kernels")
KERNEL_BYTES_PER_PX = {
    "k_ddm_codes": 10.0,        # 8 class maps (u8) in, 2-byte code word out
    "k_ddm_codes_simd": 10.0,
    "k_ddm_bits": 10.0,         # bit-sliced form of the same pass
    "k_point_max": 4.0,         # f32 point map in
    "k_point_max4": 4.0,
    "k_boost_inside": 19.0,     # codes 2 + point 4 + prob 12 in, inside mask 1 out
    "k_boost_inside4": 19.0,
    "k_ccl_init_rows": 13.0,    # mask 1 in, parent 4 + two zero-filled planes 8 out
    "k_ccl_strip": 13.0,        # same traffic; the unions happen in shared memory
    "k_ccl_merge": 2.0,         # all levels together read every mask row twice; parent plane touched sparsely
    "k_ccl_merge4": 2.0,
    "k_flatten_fill": 10.0,     # parent 4 in/out, mask 1, state 1
    "k_flatten_fill4": 10.0,
    "k_fill_merge": 1.0,        # state 1 (unions are sparse)
    "k_fill_merge4": 1.0,
    "k_flatten_area": 9.0,      # parent 4 in/out, state 1
    "k_flatten_area4": 9.0,
    "k_keep_large": 10.0,       # state 1 + parent 4 + area 4 in, keep 1 out
    "k_keep_large4": 10.0,
    "k_diag_merge": 2.0,
    "k_diag_merge4": 2.0,
    "k_flatten_count": 9.0,
    "k_flatten_count4": 9.0,
    "k_assign_ids": 5.0,
    "k_relabel": 13.0,          # parent 4 + keep 1 + idmap 4 in, labels 4 out
    "k_relabel4": 13.0,
    "k_label_dilate": 12.0,     # labels 4 in, int64 8 out
    "k_label_dilate4": 12.0,
}


def make_inputs(rank, n_tiles=TILES):
    from cdnet_b200 import synth
    tiles = [synth.postproc_inputs(100 + rank * TILES + i, H, W) for i in range(n_tiles)]
    return tiles


class ClockSampler(threading.Thread):
    """samples nvidia-smi clocks / throttle reasons for one GPU while the timed region runs"""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.proc = index, [], False, None

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.samples.append([s.strip() for s in line.split(",")])
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        try:
            if self.proc:
                self.proc.terminate()
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx.append(float(s[1]))
                for n, v in zip(names, s[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_one_tile(args):
    seed, literal = args
    from cdnet_b200 import synth
    from oracle import restate as O
    d = synth.postproc_inputs(seed, H, W)
    t0 = time.perf_counter()
    O.dam_postprocess(d["prob"], d["point"], d["dcm"], DIRECTION_CLASSES, MIN_AREA, RADIUS, POSTPROC,
                      literal=literal)
    return time.perf_counter() - t0


def cpu_baseline(n_tiles=2):
    ts = [cpu_one_tile((100 + i, True)) for i in range(n_tiles)]
    mpx = n_tiles * H * W / 1e6
    return {"value": mpx / sum(ts), "unit": "Mpixel/s", "cores": 1, "kind": "port",
            "sample": "%d of the %d 1000x1000 tiles, oracle/restate.py dam_postprocess (literal numpy/scipy "
                      "restatement of test_dam.py:455-563, postproc=0), 1 process" % (n_tiles, TILES)}


def run_reference(a):
    """the reference's CPU path (oracle port; the reference is Python and its scikit-image dependency is
    absent, so there is no oracle/_ref build) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, TILES))
    ctx = mp.get_context("spawn")
    with ctx.Pool(workers) as pool:
        jobs = [(100 + i, True) for i in range(workers)]
        for _ in range(a.warmup):
            pool.map(cpu_one_tile, jobs[:workers])
        t0 = time.perf_counter()
        for _ in range(a.steps):
            pool.map(cpu_one_tile, jobs)
        dt = time.perf_counter() - t0
    mpx = a.steps * workers * H * W / 1e6
    val = mpx / dt
    sample = ("%d 1000x1000 tiles per step (one per worker process, %d workers), oracle/restate.py "
              "dam_postprocess literal port, postproc=0" % (workers, workers))
    line = {"impl": "reference", "metric": "Mpixel/s CDNet DAM post-processing (test_dam.py:455-563)",
            "value": val, "unit": "Mpixel/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1e3 * dt / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/f32/f64->int64", "data": "synthetic",
            "config": workload_config(),
            "cpu_baseline": {"value": val, "unit": "Mpixel/s", "cores": workers, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config():
    return {"workload": "configs[1]: CDNet DAM inference post-proc, %d synthetic %dx%d tiles per GPU "
                        "(8 TTA direction-argmax u8 + prob f32[3] + point f32), postproc=%d, min_area=%d, "
                        "radius=%d, direction_classes=%d" % (TILES, H, W, POSTPROC, MIN_AREA, RADIUS,
                                                              DIRECTION_CLASSES),
            "tiles_per_gpu": TILES, "tile": [H, W], "l2_policy": "inputs larger than L2 (336 MB per step)",
            "parallelism": "independent tiles per rank, no collective"}


def extra_paths(torch, api, peak):
    """device-resident timings of the paths around the headline (CUDA events, 3 warm-ups + 10 calls each)"""
    out = {}

    def timed_ms(fn, iters=10, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    try:  # BASELINE configs[2] shape: CPM17-like 500x500 label tiles -> ternary / point / direction targets
        from cdnet_b200 import synth
        n_tiles = 64
        ids = np.stack([synth.as_uint8_label(synth.instance_map(1000 + i, 500, 500, 120))[:, :, 0] for i in range(n_tiles)])
        d_ids = torch.from_numpy(ids).cuda()
        ms = timed_ms(lambda: api.encode_targets_cuda(d_ids, True, 8))
        px = n_tiles * 500 * 500
        out["targets"] = {"workload": "configs[2]-shaped: %d synthetic 500x500 label tiles (~120 nuclei), LabelEncoding "
                                      "path, 8 direction classes, device-resident" % n_tiles,
                          "ms_per_batch": ms, "value": px / 1e6 / (ms * 1e-3), "unit": "Mpixel/s",
                          "alg_bytes_per_px": 12.0, "alg_frac_of_peak": 12.0 * px / (ms * 1e-3) / 1e9 / peak}
        del d_ids
    except Exception as e:  # noqa: BLE001
        out["targets"] = {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}
    try:  # the fused TTA hand-off (csrc/handoff.cu): 8 variants of raw logits -> prob / point / dcm
        B, C = TILES, DIRECTION_CLASSES
        g = torch.Generator(device="cuda").manual_seed(0)
        ml = [torch.randn((B, 3, H, W), device="cuda", generator=g) for _ in range(8)]
        pt = [torch.randn((B, 1, H, W), device="cuda", generator=g) for _ in range(8)]
        dl = [torch.randn((B, C, H, W), device="cuda", generator=g) for _ in range(8)]
        ms = timed_ms(lambda: api.tta_merge_cuda(ml, pt, dl), iters=5, warm=2)
        px = B * H * W
        bpp = 8 * (3 + 1 + C) * 4 + 16 + 8
        out["tta_merge"] = {"workload": "%d tiles of %dx%d, 8 TTA variants, %d direction classes" % (B, H, W, C),
                            "ms_per_batch": ms, "value": px / 1e6 / (ms * 1e-3), "unit": "Mpixel/s",
                            "alg_bytes_per_px": float(bpp), "alg_frac_of_peak": bpp * px / (ms * 1e-3) / 1e9 / peak}
        del ml, pt, dl
    except Exception as e:  # noqa: BLE001
        out["tta_merge"] = {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl != "reference" else a.warmup
    if a.impl == "reference":
        run_reference(a)
        return

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from cdnet_b200 import api, _cabi
    L = _cabi.lib()

    tiles = make_inputs(rank)
    plan = api.DamPostprocessPlan(TILES, H, W, DIRECTION_CLASSES, MIN_AREA, RADIUS, POSTPROC, write_prob=False)
    for i, t in enumerate(tiles):
        plan.h_dcm[i], plan.h_prob[i], plan.h_point[i] = t["dcm"], t["prob"], t["point"]
    plan.run()  # populates the device buffers; first-call checks
    first = plan.h_labels.copy()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        barrier()
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(a.warmup):
        plan.launch_device()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    n0 = api.launch_count()
    ms_dev = timed(plan.launch_device, a.steps)
    launches = api.launch_count() - n0
    # end to end through the host-buffer API
    for _ in range(2):
        plan.run()
    ms_e2e = timed(plan.launch, a.steps)
    clocks = sampler.finish()
    assert np.array_equal(first, plan.h_labels), "results changed between runs"

    # per-kernel CUDA-event times (library profiler), separate pass so the headline is untouched
    L.cdnet_profile_enable(1)
    for _ in range(a.steps):
        plan.launch_device()
    buf = (b"\0" * 65536)
    import ctypes
    cbuf = ctypes.create_string_buffer(65536)
    L.cdnet_profile_report(cbuf, 65536)
    L.cdnet_profile_enable(0)
    kern = {}
    for line in cbuf.value.decode().splitlines():
        name, cnt, tot = line.split("\t")
        base = name.strip("()").split("<")[0]
        k = kern.setdefault(base, [0, 0.0])
        k[0] += int(cnt)
        k[1] += float(tot)
    total_k = sum(v[1] for v in kern.values()) or 1.0
    top = max(kern.items(), key=lambda kv: kv[1][1])
    px_per_launch = TILES * H * W
    peak, peak_src = measured_peak()
    top_name, (top_cnt, top_ms) = top
    # a "launch" of the dominant kernel = everything it does for one step (the row-tree merge is one logical
    # pass split over log2(H) launches); px_per_launch pixels per step
    per_launch_ms = top_ms / a.steps
    bpp = KERNEL_BYTES_PER_PX.get(top_name, ALG_BYTES_PER_PX)
    achieved = bpp * px_per_launch / (per_launch_ms * 1e-3) / 1e9
    traffic = None
    try:
        tj = json.load(open(os.path.join(REPO, "profiles", "ncu_traffic.json")))
        if top_name in tj["bytes_per_px"]:
            traffic = tj["bytes_per_px"][top_name] * px_per_launch  # dram read+write per step, from ncu --set full
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": top_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "alg_bytes_per_px": bpp, "kernel_ms_per_launch": per_launch_ms, "launches_per_step": top_cnt / a.steps,
                "kernel_share_of_step": top_ms / total_k,
                "pipeline_alg_frac": (ALG_BYTES_PER_PX * px_per_launch / (ms_dev / a.steps * 1e-3) / 1e9) / peak,
                "kernels_ms_per_step": {k: round(v[1] / a.steps, 4) for k, v in
                                        sorted(kern.items(), key=lambda kv: -kv[1][1])}}

    mpx_step = world * TILES * H * W / 1e6
    value = mpx_step / (ms_dev / a.steps * 1e-3)
    e2e_val = mpx_step / (ms_e2e / a.steps * 1e-3)
    line = {"metric": "Mpixel/s CDNet DAM post-processing (test_dam.py:455-563)", "value": value,
            "unit": "Mpixel/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_dev / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/f32/f64->int64", "data": "synthetic", "config": workload_config(),
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": "Mpixel/s", "ms_per_step": ms_e2e / a.steps,
                    "h2d_bytes_per_step": int(plan.h2d_bytes), "d2h_bytes_per_step": int(plan.d2h_bytes)},
            "gpu_launches": int(launches), "roofline": roofline}
    if rank == 0:
        if world == 1 and not a.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline()
            # informative extras, AFTER every number of the contract has been taken: the other half of BASELINE's
            # metric (label -> direction maps) and the hand-off kernel.  Each is fenced: a failure is recorded, not raised.
            line["extra"] = extra_paths(torch, api, peak)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
