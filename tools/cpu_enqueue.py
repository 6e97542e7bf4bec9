"""CPU time to ENQUEUE one FusedHandStep.step() (python + ctypes + launches), vs its GPU time (tuning helper)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hifihr_b200 as hf
from hifihr_b200.synthetic import synthetic_inputs
B = 64
step = hf.FusedHandStep(B, image_size=224, faces_per_pixel=4, soft=True, texture_size=512, device="cuda")
inp = synthetic_inputs(B, S=224, seed=1)
fcl, prp = hf.get_ndc_fx_fy_cx_cy(inp["Ks"])
d = lambda t: t.cuda().contiguous()
args = (d(inp["pose"]), d(inp["betas"]), d(-fcl), d(prp), d(inp["root_xyz"]), d(inp["light_dir"]), d(inp["light_color"]), d(inp["imgs"]), d(inp["segms_gt"].float()))
for _ in range(5):
    step.step(*args)
torch.cuda.synchronize()
n = 200
t0 = time.perf_counter()
for _ in range(n):
    step.step(*args)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"enqueue {1e6*(t1-t0)/n:.0f} us/step (CPU), total {1e6*(t2-t0)/n:.0f} us/step")
