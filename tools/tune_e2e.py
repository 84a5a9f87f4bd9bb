"""e2e chunk-size sweep for DamPostprocessPlan.launch (tuning aid)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cdnet_b200 import api, synth
tiles = [synth.postproc_inputs(100 + i, 1000, 1000) for i in range(14)]
plan = api.DamPostprocessPlan(14, 1000, 1000, 9, 20, 2, 0, write_prob=True)  # what bench.py times as e2e
for i, t in enumerate(tiles):
    plan.h_dcm[i], plan.h_prob[i], plan.h_point[i] = t["dcm"], t["prob"], t["point"]
plan.run()
for chunk in (1, 2, 3, 4, 7, 14):
    for _ in range(3):
        plan.launch(chunk=chunk)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        plan.launch(chunk=chunk)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("chunk", chunk, "ms", round(ms, 3), "Mpx/s", round(14.0 / ms * 1e3, 1))
