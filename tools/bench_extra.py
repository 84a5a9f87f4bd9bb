#!/usr/bin/env python
"""Extra (non-contract) timings with the library's per-kernel CUDA-event profiler:
   --what targets : BASELINE configs[2]-shaped target generation (B tiles of HxW, LabelEncoding path)
   --what ws      : DAM post-processing with postproc=1 (watershed chain) on 14 x 1000^2
Prints one JSON line per run."""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def profile(L, fn, steps):
    import torch
    L.cdnet_profile_enable(1)
    for _ in range(steps):
        fn()
    buf = ctypes.create_string_buffer(1 << 16)
    L.cdnet_profile_report(buf, 1 << 16)
    L.cdnet_profile_enable(0)
    kern = {}
    for line in buf.value.decode().splitlines():
        name, cnt, tot = line.split("\t")
        base = name.strip("()").split("<")[0]
        k = kern.setdefault(base, [0, 0.0])
        k[0] += int(cnt)
        k[1] += float(tot)
    return {k: round(v[1] / steps, 4) for k, v in sorted(kern.items(), key=lambda kv: -kv[1][1])}


def timed(fn, steps):
    import torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="targets")
    ap.add_argument("--tiles", type=int, default=64)
    ap.add_argument("--size", type=int, default=500)
    ap.add_argument("--nuclei", type=int, default=120)
    ap.add_argument("--classes", type=int, default=8)
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    import torch
    from cdnet_b200 import api, synth, _cabi
    L = _cabi.lib()
    if a.what == "targets":
        base = [synth.as_uint8_label(synth.instance_map(1000 + i, a.size, a.size, a.nuclei))[:, :, 0] for i in range(8)]
        ids = np.stack([base[i % 8] for i in range(a.tiles)])
        d_ids = torch.from_numpy(ids).cuda()
        fn = lambda: api.encode_targets_cuda(d_ids, True, a.classes)
        for _ in range(2):
            fn()
        ms = timed(fn, a.steps)
        mpx = a.tiles * a.size * a.size / 1e6
        kern = profile(L, fn, a.steps)
        plan = api.EncodeTargetsPlan(a.tiles, a.size, a.size, a.classes)
        plan.h_ids[:] = ids
        t0, p0, d0 = [x.copy() for x in plan.run()]
        ref = fn()
        assert np.array_equal(t0, ref[0].cpu().numpy()) and np.array_equal(d0, ref[2].cpu().numpy())
        sweep = {}
        for ch in (8, 16, 32, 64):
            for _ in range(2):
                plan.launch(chunk=ch)
            sweep[ch] = timed(lambda: plan.launch(chunk=ch), a.steps)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dd = torch.empty_like(plan.t_direction, device="cuda")
        plan.t_direction.copy_(dd, non_blocking=True)
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(5):
            plan.t_direction.copy_(dd, non_blocking=True)
        ev1.record()
        torch.cuda.synchronize()
        d2h_gbs = dd.numel() * 8 * 5 / (ev0.elapsed_time(ev1) * 1e-3) / 1e9
        print(json.dumps({"e2e_chunk_sweep_ms": sweep, "d2h_only_GBps": d2h_gbs}))
        ms_e2e = min(sweep.values())
        print(json.dumps({"what": "targets", "tiles": a.tiles, "size": a.size, "classes": a.classes, "ms": ms,
                          "mpx_per_s": mpx / (ms * 1e-3), "e2e_ms": ms_e2e, "e2e_mpx_per_s": mpx / (ms_e2e * 1e-3),
                          "h2d_bytes": plan.h2d_bytes, "d2h_bytes": plan.d2h_bytes, "kernels_ms": kern}))
    elif a.what == "ws":
        tiles = [synth.postproc_inputs(100 + i, 1000, 1000) for i in range(14)]
        plan = api.DamPostprocessPlan(14, 1000, 1000, 9, 20, 2, 1)
        for i, t in enumerate(tiles):
            plan.h_dcm[i], plan.h_prob[i], plan.h_point[i] = t["dcm"], t["prob"], t["point"]
        plan.run()
        for _ in range(2):
            plan.launch_device()
        ms = timed(plan.launch_device, a.steps)
        print(json.dumps({"what": "dam postproc=1", "ms": ms, "mpx_per_s": 14.0 / (ms * 1e-3),
                          "kernels_ms": profile(L, plan.launch_device, a.steps)}))
    else:
        assert a.what == "metrics", a.what
        # instance metrics (stats_utils.py drop-ins): pair-table reduction for 14 x 1000^2 label pairs on the device,
        # then the full host-buffer call per tile (H2D + kernels + D2H of the table + float64 epilogue), and the
        # oracle port on one tile as the CPU yardstick (the reference itself: 2.4 s get_fast_aji + 1.7 s get_fast_pq
        # per 1000^2 tile with ~700 nuclei, measured in the build container)
        import contextlib
        import io
        from cdnet_b200 import metrics as M
        pairs = [synth.metric_pair(900 + i, 1000, 1000, 700, 1) for i in range(2)]
        t = torch.from_numpy(np.stack([pairs[i % 2][0] for i in range(14)])).cuda()
        p = torch.from_numpy(np.stack([pairs[i % 2][1] for i in range(14)])).cuda()
        cap = 31250
        keys = torch.empty((14, cap), dtype=torch.int64, device="cuda")
        counts = torch.empty((14, cap), dtype=torch.int32, device="cuda")
        n_out = torch.empty((14,), dtype=torch.int32, device="cuda")
        status = torch.empty((14,), dtype=torch.int32, device="cuda")
        ws = torch.empty(L.cdnet_label_pairs_workspace_bytes(14, cap), dtype=torch.uint8, device="cuda")

        def fn():
            rc = L.cdnet_label_pairs(t.data_ptr(), p.data_ptr(), 4, keys.data_ptr(), counts.data_ptr(), n_out.data_ptr(),
                                     status.data_ptr(), 14, 1000, 1000, cap, ws.data_ptr(), ws.numel(),
                                     torch.cuda.current_stream().cuda_stream)
            assert rc == 0, rc
        for _ in range(3):
            fn()
        ms = timed(fn, a.steps)
        kern = profile(L, fn, a.steps)
        assert int(status.max()) == 0
        tt, pp = pairs[0]
        with contextlib.redirect_stdout(io.StringIO()):
            M.get_fast_aji(tt, pp)
            t0 = time.time()
            for _ in range(5):
                aji = M.get_fast_aji(tt, pp)
                pq = M.get_fast_pq(tt, pp)
                dice = M.get_dice_1(tt, pp)
            host_ms = (time.time() - t0) / 5 * 1e3
            t0 = time.time()
            res = M.instance_metrics_cuda(t, p)
            batch_ms = (time.time() - t0) * 1e3
        print(json.dumps({"what": "metrics", "pair_table_ms_14x1000x1000": ms, "mpx_per_s": 14.0 / (ms * 1e-3),
                          "alg_GBps": 14e6 * 8 / (ms * 1e-3) / 1e9, "pairs_per_tile": int(n_out[0]),
                          "dropin_aji_pq_dice_ms_per_tile_host_buffers": host_ms,
                          "instance_metrics_cuda_ms_14_tiles_device_resident": batch_ms,
                          "aji": float(aji[0]), "pq": float(pq[0][2]), "dice": float(dice), "kernels_ms": kern}))


if __name__ == "__main__":
    main()
