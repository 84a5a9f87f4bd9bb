"""CPU: the host logic of the multi-GPU paths (cdnet_b200/sharded.py).

* tile sharding: contiguous, complete, balanced;
* whole-slide row partition: (a) all ranks simulated in-process, (b) two real processes over
  torch.distributed with the gloo backend -- both must reproduce the UNSHARDED oracle bit for bit
  (labels incl. raster-order numbering, hole filling and small-object removal across seams).
The per-rank device ops are a numpy stand-in (tests/sharded_numpy_backend.py); the CUDA backend is
checked against the same property on the GPU box (tests/test_gpu_sharded.py) and, here, in 2 and 3 gloo
processes with its kernels running under the SIMT emulator (tests/simt)."""
import os
import sys

import numpy as np
import pytest

from cdnet_b200 import sharded, synth
from oracle import restate as O

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sharded_numpy_backend import NumpyBackend  # noqa: E402


def test_shard_tiles():
    for n in (1, 7, 14, 256, 1024):
        for w in (1, 2, 3, 4, 8):
            parts = [sharded.shard_tiles(n, w, r) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def _slide(seed, H, W, n, n_maps):
    d = synth.postproc_inputs(seed, H, W, n)
    # big structures that straddle many seams: a frame-touching ring with a hole, a long vertical bar
    yy, xx = np.mgrid[0:H, 0:W]
    ring = ((yy - H // 2) ** 2 / (H * 0.42) ** 2 + (xx - W // 2) ** 2 / (W * 0.3) ** 2 <= 1) & \
           ~((yy - H // 2) ** 2 / (H * 0.36) ** 2 + (xx - W // 2) ** 2 / (W * 0.22) ** 2 <= 1)
    bar = (abs(xx - W // 5) < 3) & (yy > 3) & (yy < H - 4)
    big = ring | bar
    d["prob"][1][big] += np.float32(3.0)
    return d["dcm"][:n_maps].copy(), d["prob"], d["point"]


def _split(dcm, prob, point, H, G):
    parts = sharded.row_partition(H, G)
    return [dict(dcm=dcm[:, a:b].copy(), prob=prob[:, a:b].copy(), point=point[:, a:b].copy()) for a, b in parts]


@pytest.mark.parametrize("G", [2, 3, 5])
@pytest.mark.parametrize("n_maps", [8, 1])
def test_slide_simulated_ranks(G, n_maps):
    H, W = 151, 164
    dcm, prob, point = _slide(31, H, W, 40, n_maps)
    ref = O.dam_postprocess(prob.copy(), point, dcm, 9, 20, 2, 0, literal=False)["pred_labeled"]
    outs = sharded.postprocess_slide(_split(dcm, prob, point, H, G), sharded.SimComm(G), H, W, NumpyBackend(), 9, 20, 2)
    got = np.concatenate(outs, axis=0)
    assert got.dtype == ref.dtype
    assert np.array_equal(got, ref), int((got != ref).sum())


def _worker(rank, world, port, H, W, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dcm, prob, point = _slide(32, H, W, 36, 8)
        mine = _split(dcm, prob, point, H, world)[rank]
        comm = sharded.DistComm()
        out = sharded.postprocess_slide([mine], comm, H, W, NumpyBackend(), 9, 20, 2)[0]
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_slide_two_processes_gloo():
    import torch.multiprocessing as mp
    H, W, world = 140, 150, 2
    port = 29500 + (os.getpid() % 2000)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, H, W, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    dcm, prob, point = _slide(32, H, W, 36, 8)
    ref = O.dam_postprocess(prob.copy(), point, dcm, 9, 20, 2, 0, literal=False)["pred_labeled"]
    got = np.concatenate([res[r] for r in range(world)], axis=0)
    assert np.array_equal(got, ref)


def _worker_kernels(rank, world, port, H, W, n_maps, q):
    """like _worker, but with the product's CudaBackend: its kernels run under the SIMT emulator (tests/simt), its
    exchanges over gloo -- the orchestration, the C-ABI calls and the kernels are all the shipped code"""
    import torch.distributed as dist
    from simt import emulated_api
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        with emulated_api():
            dcm, prob, point = _slide(34, H, W, 30, n_maps)
            mine = _split(dcm, prob, point, H, world)[rank]
            be = sharded.CudaBackend()
            out = sharded.postprocess_slide([mine], sharded.DistComm(), H, W, be, 9, 20, 2)[0]
            q.put((rank, be.to_host(out)))
    finally:
        dist.destroy_process_group()


@pytest.mark.simt
@pytest.mark.parametrize("world,n_maps", [(2, 8), (3, 1)])
def test_slide_processes_gloo_with_emulated_kernels(world, n_maps):
    import torch.multiprocessing as mp
    H, W = 132, 140
    port = 31500 + (os.getpid() % 2000) + world
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_kernels, args=(r, world, port, H, W, n_maps, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    dcm, prob, point = _slide(34, H, W, 30, n_maps)
    ref = O.dam_postprocess(prob.copy(), point, dcm, 9, 20, 2, 0, literal=False)["pred_labeled"]
    got = np.concatenate([res[r] for r in range(world)], axis=0)
    assert got.dtype == ref.dtype and np.array_equal(got, ref), int((got != ref).sum())


def test_constant_direction_map_asserts():
    H, W = 60, 64
    dcm, prob, point = _slide(33, H, W, 6, 8)
    dcm[2] = 0
    with pytest.raises(AssertionError):
        sharded.postprocess_slide(_split(dcm, prob, point, H, 2), sharded.SimComm(2), H, W, NumpyBackend(), 9, 20, 2)


@pytest.mark.parametrize("seed,H,W,G", [(51, 64, 40, 8), (52, 97, 131, 6), (53, 33, 300, 4), (54, 120, 64, 7)])
def test_slide_many_ranks_small_shards(seed, H, W, G):
    """thin shards (a few rows each): components cross several seams, holes span many shards"""
    rng = np.random.default_rng(seed)
    dcm, prob, point = _slide(seed, H, W, max(4, H * W // 900), 8)
    # serpentine foreground that winds through every shard, plus random specks on the seams
    yy, xx = np.mgrid[0:H, 0:W]
    snake = ((yy // 3) % 2 == 0) & (xx > 2) & (xx < W - 3)
    link = ((yy % 6) == 3) & (xx >= W - 6) & (xx < W - 3) | ((yy % 6) == 0) & (xx > 2) & (xx <= 5) & (yy > 0)
    prob[1][(snake | link) & (rng.random((H, W)) < 0.97)] += np.float32(3.0)
    ref = O.dam_postprocess(prob.copy(), point, dcm, 9, 20, 2, 0, literal=False)["pred_labeled"]
    outs = sharded.postprocess_slide(_split(dcm, prob, point, H, G), sharded.SimComm(G), H, W, NumpyBackend(), 9, 20, 2)
    got = np.concatenate(outs, axis=0)
    assert np.array_equal(got, ref), int((got != ref).sum())


def _worker_watershed(rank, world, port, H, W, overlap, seed, q):
    """postproc = 1 over gloo with the product's CudaBackend (kernels under the SIMT emulator)"""
    import torch.distributed as dist
    from simt import emulated_api
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        with emulated_api():
            d = synth.postproc_inputs(seed, H, W, 60)
            mine = _split(d["dcm"], d["prob"], d["point"], H, world)[rank]
            be = sharded.CudaBackend()
            try:
                out = sharded.postprocess_slide([mine], sharded.DistComm(), H, W, be, 9, 20, 2, postproc=1,
                                                overlap=overlap)[0]
                q.put((rank, be.to_host(out)))
            except RuntimeError as e:
                q.put((rank, str(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.simt
@pytest.mark.parametrize("world,overlap", [(2, 48), (3, 40), (2, 4)])
def test_slide_watershed_gloo_with_emulated_kernels(world, overlap):
    """overlap 4 is far too small for the nuclei: EVERY rank must raise (the error word is all-gathered), none may
    run ahead into the next collective"""
    import torch.multiprocessing as mp
    H, W = 240, 200
    port = 33500 + (os.getpid() % 2000) + world + overlap
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_watershed, args=(r, world, port, H, W, overlap, 35, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    if overlap < 8:
        assert all(isinstance(v, str) and "overlap" in v for v in res.values()), res
        return
    d = synth.postproc_inputs(35, H, W, 60)
    ref = O.dam_postprocess(d["prob"].copy(), d["point"], d["dcm"], 9, 20, 2, 1, literal=False)["pred_labeled"]
    got = np.concatenate([res[r] for r in range(world)], axis=0)
    assert got.dtype == ref.dtype and np.array_equal(got, ref), int((got != ref).sum())
