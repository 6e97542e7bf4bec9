"""Does a host->device copy in flight slow the step's kernels?  Graph-replayed C2 step alone vs with a copy stream
streaming 12.9 MB pinned-host buffers the whole time (tuning / diagnosis helper)."""
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
g = step.capture(*args)
host = torch.empty(12_862_976, dtype=torch.uint8).pin_memory()
devb = torch.empty_like(host, device="cuda")
cs = torch.cuda.Stream()
def run(n, copies_per_step):
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        for _ in range(copies_per_step):
            with torch.cuda.stream(cs):
                devb.copy_(host, non_blocking=True)
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for c in (0, 1, 2, 3, 4):
    print(f"{c} copies of 12.9 MB in flight per step: {run(100, c):.4f} ms/step")
