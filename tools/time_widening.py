"""Times the kernels added while widening round 1 (csrc/handoff.cu, csrc/training.cu) with CUDA events:
    python tools/time_widening.py [--tiles 14] [--size 1000] > gpurun_out/widening_times.json
Algorithmic GB/s = compulsory bytes at the kernel's own boundary / time."""
import argparse
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from cdnet_b200 import api, training  # noqa: E402


def timed(fn, iters=5, warm=2):
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tiles", type=int, default=14)
    ap.add_argument("--size", type=int, default=1000)
    a = ap.parse_args()
    B, H, W, C = a.tiles, a.size, a.size, 9
    dev = torch.device("cuda", 0)
    px = B * H * W
    res = {"tiles": B, "tile": [H, W], "peak_gbs": 6554.6}
    g = torch.Generator(device=dev).manual_seed(0)
    ml = [torch.randn((B, 3, H, W), device=dev, generator=g) for _ in range(8)]
    pt = [torch.randn((B, 1, H, W), device=dev, generator=g) for _ in range(8)]
    dl = [torch.randn((B, C, H, W), device=dev, generator=g) for _ in range(8)]
    ms = timed(lambda: api.tta_merge_cuda(ml, pt, dl))
    byt = px * (8 * (3 + 1 + C) * 4 + 16 + 8)
    res["k_tta_merge"] = {"ms": ms, "alg_gbs": byt / ms / 1e6, "alg_bytes_per_px": byt / px}

    def torch_path():
        out = []
        for v in range(8):
            p = torch.softmax(ml[v], dim=1)
            d = torch.softmax(dl[v], dim=1)
            d[:, 0] = d[:, 0] * p[:, 0]
            out.append((p, torch.argmax(d, dim=1)))
        return out
    res["torch_softmax_argmax_x8_no_unflip"] = {"ms": timed(torch_path)}
    del ml, pt, dl
    torch.cuda.empty_cache()

    direction = torch.randint(0, C, (B, H, W), device=dev, generator=g)
    target = torch.randint(0, 3, (B, H, W), device=dev, generator=g)
    ms = timed(lambda: training.direction_one_hot_cuda(direction, target, C))
    byt = px * (8 + 8 + 4 * C)
    res["k_dir_one_hot(+minmax)"] = {"ms": ms, "alg_gbs": byt / ms / 1e6, "alg_bytes_per_px": byt / px}
    lab = direction.to(torch.uint8)
    ms = timed(lambda: training.DTOffsetHelper.label_to_vector(lab, C))
    res["k_label_to_vector<u8>"] = {"ms": ms, "alg_gbs": px * 17 / ms / 1e6, "alg_bytes_per_px": 17}
    ang = (torch.rand((B, H, W), device=dev, generator=g) * 360 - 180)
    ms = timed(lambda: training.DTOffsetHelper.align_angle(ang, 8, return_tensor=True))
    res["k_align_angle<f32,f32>"] = {"ms": ms, "alg_gbs": px * 16 / ms / 1e6, "alg_bytes_per_px": 16}
    ms = timed(lambda: training.DTOffsetHelper.angle_to_vector(ang, 8, return_tensor=True))
    res["k_angle_to_vector<f32,f32>"] = {"ms": ms, "alg_gbs": px * 12 / ms / 1e6, "alg_bytes_per_px": 12}
    vec = torch.randn((B, H, W, 2), device=dev, generator=g)
    ms = timed(lambda: training.DTOffsetHelper.vector_to_label(vec, 8, return_tensor=True))
    res["k_vector_to_label<f32>"] = {"ms": ms, "alg_gbs": px * 16 / ms / 1e6, "alg_bytes_per_px": 16}
    ids = (torch.rand((B, H, W), device=dev, generator=g) * 4).to(torch.uint8)
    ms = timed(lambda: training.ternary_label_cuda(ids, 0))
    res["k_ternary_label"] = {"ms": ms, "alg_gbs": px * 2 / ms / 1e6, "alg_bytes_per_px": 2}
    for k, v in res.items():
        if isinstance(v, dict) and "alg_gbs" in v:
            v["frac_of_peak"] = v["alg_gbs"] / res["peak_gbs"]
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
