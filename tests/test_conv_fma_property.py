"""CPU: re-asserts on THIS host that torch's CPU conv2d (what the reference calls at
my_transforms_direction.py:827-830) is bit-identical to the sequential f32 FMA chain the oracle and the
CUDA kernel implement (SURVEY.md section 7 hard-part 2).  A failure here means this host's
torch/oneDNN build sums in a different order -- informative for the reference-vs-host question, it does
not affect CUDA-vs-oracle parity (both use the chain)."""
import numpy as np
import pytest


def test_torch_conv2d_equals_fma_chain():
    torch = pytest.importorskip("torch")
    from oracle import clib, restate as O
    ker = O.sobel_kernels(11)
    rng = np.random.default_rng(0)
    H, W = 96, 120
    yy, xx = np.mgrid[0:H, 0:W]
    img = np.zeros((H, W))
    for _ in range(5):
        cy, cx, r = rng.integers(8, H - 8), rng.integers(8, W - 8), rng.integers(4, 14)
        d = np.sqrt((yy - cy) ** 2 + (xx - cx) ** 2)
        m = d <= r
        img = np.where(m, 1 - d / (d[m].max() + 1e-7), img)
    img32 = img.astype(np.float32)
    ref = torch.nn.functional.conv2d(torch.from_numpy(img32).view(1, 1, H, W),
                                     torch.from_numpy(ker).view(2, 1, 11, 11), padding=5)[0].numpy()
    mine = clib.conv11_fma(img32, ker)
    assert np.array_equal(ref.view(np.uint32), mine.view(np.uint32))
