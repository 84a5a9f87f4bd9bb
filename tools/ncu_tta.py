"""one launch of k_tta_merge (and one of k_ternary_label) for an ncu capture: python tools/ncu_tta.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from cdnet_b200 import api, training  # noqa: E402
B, H, W, C = 6, 1000, 1000, 9
dev = torch.device("cuda", 0)
ml = [torch.randn((B, 3, H, W), device=dev) for _ in range(8)]
pt = [torch.randn((B, 1, H, W), device=dev) for _ in range(8)]
dl = [torch.randn((B, C, H, W), device=dev) for _ in range(8)]
for _ in range(2):
    api.tta_merge_cuda(ml, pt, dl)
torch.cuda.synchronize()
