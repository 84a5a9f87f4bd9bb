import torch, sys
sys.path.insert(0, '.')
from cdnet_b200 import api, synth
import numpy as np
d = synth.postproc_inputs(100, 1000, 1000)
B = 14
dcm = torch.from_numpy(d["dcm"])[None].repeat(B,1,1,1).cuda().contiguous()
prob = torch.from_numpy(d["prob"])[None].repeat(B,1,1,1).cuda().contiguous()
point = torch.from_numpy(d["point"])[None].repeat(B,1,1,1).cuda().contiguous()
for i in range(3):
    out, st = api.dam_postprocess_cuda(dcm, prob.clone(), point, 9, 20, 2, 1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
p2 = prob.clone()
e0.record()
out, st = api.dam_postprocess_cuda(dcm, p2, point, 9, 20, 2, 1)
e1.record(); torch.cuda.synchronize()
print("postproc=1 14x1000^2 ms", e0.elapsed_time(e1), "labels", int(out.max()))
