"""The other BASELINE.json configs, timed beside bench.py's contract line (its `extra` object, every N).

configs[4]  whole_slide : 40 000 x 40 000 synthetic prediction map (P1: one direction map + prob f32[3] + point f32),
                          row-partitioned over the N ranks, cdnet_b200.sharded.postprocess_slide over NCCL (halo rows,
                          two scalars, three seam rounds); strong scaling: the slide is fixed, N grows.  `verified`:
                          a down-sized slide is post-processed sharded AND on rank 0 alone in the same run and the
                          labels are compared bit for bit on the device.
configs[2]  targets     : 256 CPM17-shaped 500 x 500 label tiles -> ternary / point / direction targets, tiles sharded
                          with sharded.shard_tiles (no collective), device-resident and through EncodeTargetsPlan
                          (pinned host buffers, H2D + D2H inside the timed region).
configs[3]  config3     : 1 024 tiles of 1000 x 1000, 16-direction target transform -> 17-class direction-difference
                          map of the produced direction classes -> 4-connected labelling of the interior mask.
configs[0]  config0     : ONE 1000 x 1000 instance map through the reference-signature drop-in
                          (api.LabelEncoding.__call__, numpy in / numpy out), N = 1 only.
Every figure is a max over ranks of a CUDA-event (or wall-clock, whichever is larger) time.  Each block is fenced by
the caller: a failure is recorded in the JSON, never raised.
"""
import os
import time

import numpy as np

SLIDE_TILE = 1000
SLIDE_YOFF = 333  # the tiling is shifted so that shard boundaries (multiples of H / N) cut through nuclei


def _sync_max(torch, dist, world, ms):
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return float(ms)


def _barrier(torch, dist, world):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def slide_bands(torch, W, n_maps, n_seeds=4):
    """4 device-resident bands [C, 1000, W] per plane: band k = the tiles of slide tile-row k (mod 4)"""
    from cdnet_b200 import synth
    base = [synth.postproc_inputs(100 + i, SLIDE_TILE, SLIDE_TILE) for i in range(n_seeds)]
    dev = {k: [torch.from_numpy(np.ascontiguousarray(b[k][:n_maps] if k == "dcm" else b[k])).cuda() for b in base]
           for k in ("dcm", "prob", "point")}
    bands = []
    for yt in range(n_seeds):
        band = {}
        for k in ("dcm", "prob", "point"):
            parts = []
            for x in range(0, W, SLIDE_TILE):
                t = dev[k][(yt * 7 + (x // SLIDE_TILE) * 3) % n_seeds]
                parts.append(t[..., :min(SLIDE_TILE, W - x)])
            band[k] = torch.cat(parts, dim=-1).contiguous()
        bands.append(band)
    return bands


def fill_slide_rows(bands, dst, r0, r1):
    """rows [r0, r1) of the synthetic slide into the dict of device views dst (dcm / prob / point, rows axis -2)"""
    n = len(bands)
    y = r0
    while y < r1:
        g = y + SLIDE_YOFF
        yt, wy = divmod(g, SLIDE_TILE)
        take = min(SLIDE_TILE - wy, r1 - y)
        for k in ("dcm", "prob", "point"):
            dst[k][:, y - r0:y - r0 + take] = bands[yt % n][k][:, wy:wy + take]
        y += take


def whole_slide(torch, dist, rank, world, H=40000, W=40000, steps=3, verify_hw=(8000, 6000), peak=None, postproc=0,
                overlap=128):
    from cdnet_b200 import api, sharded
    be = sharded.CudaBackend()
    comm = sharded.DistComm() if world > 1 else sharded.SimComm(1)
    out = {"workload": "configs[4]: %dx%d synthetic prediction map (P1: 1 direction map u8 + prob f32[3] + point f32), "
                       "row-partitioned over %d rank(s), sharded.postprocess_slide over %s, postproc=%d%s, min_area=20, "
                       "radius=2" % (H, W, world, "NCCL" if world > 1 else "one process", postproc,
                                     " (watershed; %d overlap rows per seam)" % overlap if postproc else ""),
           "n_gpus": world, "scaling": "strong", "alg_bytes_per_px": 21.0}
    bands = slide_bands(torch, max(W, verify_hw[1]), 1)

    graph_state = {"used": False}

    def run(h, w, n_steps, want_phases, use_graph=False):
        r0, r1 = sharded.row_partition(h, world)[rank]
        bw = [{k: b[k][..., :w] for k in b} for b in bands] if w != bands[0]["dcm"].shape[-1] else bands
        bufs = sharded.alloc_shard_buffers(be, rank, world, h, w, 1)
        fill_slide_rows(bw, bufs, r0, r1)
        eager = lambda tm=None: sharded.postprocess_slide([bufs], comm, h, w, be, 9, 20, 2, timings=tm, postproc=postproc,
                                                          overlap=overlap)[0]
        step = eager
        if use_graph:
            # the whole step (kernels, torch ops, NCCL all-gathers) recorded once in a CUDA graph and replayed
            try:
                plan = sharded.SlidePlan(bufs, comm, h, w, be, 9, 20, 2, postproc=postproc, overlap=overlap)
                step = lambda tm=None: plan.run() if tm is None else eager(tm)
                graph_state["used"] = True
            except Exception as e:  # noqa: BLE001
                graph_state["error"] = "%s: %s" % (type(e).__name__, str(e)[:200])
        lab = step()
        times = []
        for _ in range(n_steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            _barrier(torch, dist, world)
            t0 = time.perf_counter()
            e0.record()
            lab = step()
            e1.record()
            torch.cuda.synchronize()
            ms = max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0))
            times.append(_sync_max(torch, dist, world, ms))
        if graph_state.get("used") and use_graph:
            ref = eager()
            graph_state["equals_eager"] = bool(torch.equal(ref, lab))
            del ref
        phases = None
        if want_phases:
            tm = {}
            _barrier(torch, dist, world)
            step(tm)
            phases = {}
            for k, v in tm.items():
                if world > 1:
                    t = torch.tensor([v], device="cuda", dtype=torch.float64)
                    g = torch.empty((world,), device="cuda", dtype=torch.float64)
                    dist.all_gather_into_tensor(g, t)
                    vals = g.cpu().tolist()
                else:
                    vals = [v]
                phases[k] = {"max": round(max(vals), 3), "min": round(min(vals), 3), "mean": round(sum(vals) / len(vals), 3)}
        return lab, times, phases, (r0, r1)

    # ---- verification on a slide one GPU can also process alone: sharded == single-GPU, bit for bit
    hv, wv = verify_hw
    lab, _, _, (r0, r1) = run(hv, wv, 0, False)
    if world > 1:
        assert hv % world == 0
        full = torch.empty((hv, wv), dtype=lab.dtype, device="cuda")
        dist.all_gather_into_tensor(full, lab.contiguous())
    else:
        full = lab
    ok = None
    if rank == 0:
        whole = {"dcm": torch.empty((1, hv, wv), dtype=torch.uint8, device="cuda"),
                 "prob": torch.empty((3, hv, wv), dtype=torch.float32, device="cuda"),
                 "point": torch.empty((1, hv, wv), dtype=torch.float32, device="cuda")}
        fill_slide_rows([{k: b[k][..., :wv] for k in b} for b in bands], whole, 0, hv)
        single, _ = api.dam_postprocess_cuda(whole["dcm"][None], whole["prob"][None], whole["point"][None], 9, 20, 2,
                                             postproc)
        ok = bool(torch.equal(single[0], full))
        out["verify"] = {"slide": [hv, wv], "n_labels": int(single.max().item()),
                         "reference": "api.dam_postprocess_cuda on rank 0 alone (unsharded tile path)"}
        del whole, single
    del full, lab
    out["verified"] = ok
    torch.cuda.empty_cache()
    # ---- the timed slide
    lab, times, phases, _ = run(H, W, steps, True, use_graph=os.environ.get("CDNET_SLIDE_NO_GRAPH") is None)
    out["cuda_graph"] = graph_state
    ms = float(np.median(times))
    out.update({"slide": [H, W], "ms_per_slide": ms, "times_ms": [round(t, 3) for t in times],
                "value": H * W / 1e6 / (ms * 1e-3), "unit": "Mpixel/s", "phases_ms": phases,
                "phases_note": "one extra step with a device synchronisation after every phase (max / min / mean over ranks); "
                               "the phases contain their halo exchanges / seam all-gathers"})
    if peak:
        out["alg_frac_of_peak"] = 21.0 * H * W / (ms * 1e-3) / 1e9 / (peak * world)
    del lab
    torch.cuda.empty_cache()
    return out


def _label_tiles(n_distinct, size, nuclei, seed0):
    from cdnet_b200 import synth
    return [synth.as_uint8_label(synth.instance_map(seed0 + i, size, size, nuclei))[:, :, 0] for i in range(n_distinct)]


def targets_config2(torch, dist, rank, world, steps=5, peak=None):
    from cdnet_b200 import api, sharded
    n_tiles, size = 256, 500
    lo, hi = sharded.shard_tiles(n_tiles, world, rank)
    base = _label_tiles(8, size, 120, 1000)
    ids = np.stack([base[i % 8] for i in range(lo, hi)])
    d_ids = torch.from_numpy(ids).cuda()
    fn = lambda: api.encode_targets_cuda(d_ids, True, 8)
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _barrier(torch, dist, world)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = _sync_max(torch, dist, world, e0.elapsed_time(e1) / steps)
    plan = api.EncodeTargetsPlan(hi - lo, size, size, 8)
    plan.h_ids[:] = ids
    for _ in range(2):
        plan.launch()
    _barrier(torch, dist, world)
    e0.record()
    for _ in range(steps):
        plan.launch()
    e1.record()
    torch.cuda.synchronize()
    ms_e2e = _sync_max(torch, dist, world, e0.elapsed_time(e1) / steps)
    px = n_tiles * size * size
    out = {"workload": "configs[2]: 256 synthetic CPM17-shaped 500x500 label tiles (~120 nuclei each, 8 distinct maps "
                       "repeated), LabelEncoding path (my_transforms_direction.py:697-885), 8 direction classes, "
                       "tiles sharded over %d rank(s), no collective" % world,
           "n_gpus": world, "scaling": "strong", "tiles_per_rank": hi - lo, "ms_per_batch": ms,
           "value": px / 1e6 / (ms * 1e-3), "unit": "Mpixel/s", "alg_bytes_per_px": 12.0,
           "e2e": {"value": px / 1e6 / (ms_e2e * 1e-3), "unit": "Mpixel/s", "ms_per_batch": ms_e2e,
                   "h2d_bytes_per_step": int(plan.h2d_bytes) * world, "d2h_bytes_per_step": int(plan.d2h_bytes) * world,
                   "api": "EncodeTargetsPlan.launch (pinned host buffers)"}}
    if peak:
        out["alg_frac_of_peak"] = 12.0 * px / (ms * 1e-3) / 1e9 / (peak * world)
    return out


def config3(torch, dist, rank, world, steps=2, chunk=32, peak=None):
    from cdnet_b200 import api, sharded
    n_tiles, size = 1024, 1000
    lo, hi = sharded.shard_tiles(n_tiles, world, rank)
    base = _label_tiles(4, size, 700, 5000)
    d_base = torch.from_numpy(np.stack(base)).cuda()
    idx = torch.arange(chunk, device="cuda") % 4
    d_ids = d_base[idx].contiguous()  # one chunk of label tiles; every chunk of the rank's share re-uses it

    def chunk_pass(n):
        ids = d_ids[:n]
        ternary, point, direction = api.encode_targets_cuda(ids, True, 16)
        ddm = api.ddm_cuda(direction, 17)          # direction classes 0..16 -> 17-class direction-difference map
        lab = api.label_cuda(ternary == 255, 4)    # 4-connected labelling of the interior mask
        return ddm, lab

    def full_pass():
        for a in range(lo, hi, chunk):
            chunk_pass(min(chunk, hi - a))

    chunk_pass(min(chunk, hi - lo))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _barrier(torch, dist, world)
    e0.record()
    for _ in range(steps):
        full_pass()
    e1.record()
    torch.cuda.synchronize()
    ms = _sync_max(torch, dist, world, e0.elapsed_time(e1) / steps)
    px = n_tiles * size * size
    out = {"workload": "configs[3]: 1024 synthetic 1000x1000 label tiles (~700 nuclei, 4 distinct maps repeated), "
                       "16-direction target transform -> 17-class direction-difference map -> 4-connected labelling "
                       "of the interior mask; tiles sharded over %d rank(s) in chunks of %d, device-resident, no "
                       "collective" % (world, chunk),
           "n_gpus": world, "scaling": "strong", "tiles_per_rank": hi - lo, "ms_per_pass": ms,
           "value": px / 1e6 / (ms * 1e-3), "unit": "Mpixel/s", "alg_bytes_per_px": 20.0}
    if peak:
        out["alg_frac_of_peak"] = 20.0 * px / (ms * 1e-3) / 1e9 / (peak * world)
    return out


def config0(torch, cpu=True):
    """configs[0]: the reference's own CPU-runnable case -- one 1000x1000 instance map through the drop-in"""
    from cdnet_b200 import api, synth
    lab3 = synth.as_uint8_label(synth.instance_map(0, 1000, 1000, 700))
    enc = api.LabelEncoding(3, 1, 1)
    imgs = (None, None, lab3)
    enc(imgs)
    t0 = time.perf_counter()
    n = 5
    for _ in range(n):
        res = enc(imgs)
    torch.cuda.synchronize()
    ms = 1e3 * (time.perf_counter() - t0) / n
    out = {"workload": "configs[0]: one synthetic 1000x1000 MoNuSeg-shaped instance map (~700 nuclei) -> ternary / point / "
                       "8-direction maps through api.LabelEncoding(3,1,1).__call__ (numpy in, numpy out, host<->device "
                       "copies inside)", "ms_per_tile": ms, "value": 1.0 / (ms * 1e-3), "unit": "Mpixel/s"}
    del res
    return out
