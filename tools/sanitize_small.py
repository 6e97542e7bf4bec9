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
# NIMBLE-shaped step (texel-major texture PCA, K = 1 hard walk, fused rasterizer backward) and the face-vertex gather
B, S, T = 3, 64, 64
step = hf.FusedNimbleStep(B, image_size=S, texture_size=T, device="cuda")
g = torch.Generator().manual_seed(5)
pose = torch.cat([torch.randn(B, 3, generator=g) * 0.4, torch.randn(B, 30, generator=g) * 0.5], 1).cuda()
shape, texp = (torch.randn(B, 20, generator=g) * 0.5).cuda(), torch.randn(B, 10, generator=g).cuda()
inp = synthetic_inputs(B, S=S, seed=4)
fcl, prp = hf.get_ndc_fx_fy_cx_cy(inp["Ks"])
root = torch.tensor([[0.0, 0.0, 0.45]]).repeat(B, 1).cuda()
step.step(pose, shape, d(-fcl), d(prp), root, d(inp["light_dir"]), d(inp["light_color"]), d(inp["imgs"]), d(inp["segms_gt"].float()),
          tex_params=texp)
torch.cuda.synchronize()
assert torch.isfinite(step.g_pose).all() and torch.isfinite(step.g_tex_params).all()
from hifihr_b200 import ops
v = torch.randn(B, step.hm.V, 3, device="cuda", requires_grad=True)
ops.FaceVertsFunction.apply(step.topo, v).square().sum().backward()
torch.cuda.synchronize()
print("ok nimble", float(step.g_pose.abs().max()), float(v.grad.abs().max()))
