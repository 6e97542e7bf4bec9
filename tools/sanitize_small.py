"""Small end-to-end steps for compute-sanitizer (memcheck / racecheck): fused step with the tile queue at K = 4 / 8 / 1,
batched hand layer at B = 9 (ragged 128-sample tile), graph-free."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hifihr_b200 as hf
from hifihr_b200.synthetic import synthetic_inputs
for (B, S, K, soft) in ((9, 72, 4, True), (8, 56, 8, True), (8, 48, 1, False)):
    step = hf.FusedHandStep(B, image_size=S, faces_per_pixel=K, soft=soft, texture_size=64, device="cuda")
    inp = synthetic_inputs(B, S=S, seed=3)
    fcl, prp = hf.get_ndc_fx_fy_cx_cy(inp["Ks"])
    d = lambda t: t.cuda().contiguous()
    args = (d(inp["pose"]), d(inp["betas"]), d(-fcl), d(prp), d(inp["root_xyz"]), d(inp["light_dir"]), d(inp["light_color"]),
            d(inp["imgs"]), d(inp["segms_gt"].float()))
    step.step(*args)
    torch.cuda.synchronize()
    step.check_status()
    assert torch.isfinite(step.g_pose).all() and torch.isfinite(step.g_texture).all()
    print("ok", B, S, K, float(step.g_pose.abs().max()))
