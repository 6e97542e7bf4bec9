import sys; sys.path.insert(0,'/root/repo')
import torch
import hifihr_b200 as hf
from oracle import p3d, pipeline as P
DEV='cuda'
B,S,K=3,96,4
inp = P.synthetic_inputs(B, S=S, seed=23)
fcl, prp = p3d.ndc_intrinsics(inp["Ks"])
d = lambda t: t.to(DEV).contiguous()
args = (d(inp["pose"]), d(inp["betas"]), d(-fcl), d(prp), d(inp["root_xyz"]), d(inp["light_dir"]), d(inp["light_color"]), d(inp["imgs"]), d(inp["segms_gt"].float()))
step = hf.FusedHandStep(B, image_size=S, faces_per_pixel=K, soft=True, texture_size=64, device=DEV)
step.forward(*args)
runs=[]
for rep in range(4):
    step.backward(args[0],args[1],args[2],args[3],args[4])
    torch.cuda.synchronize()
    runs.append({k: getattr(step,k).clone() for k in ("g_image","face_rec","g_pose","g_betas","g_texture","g_light_dir","g_light_color","g_verts")})
for k in runs[0]:
    print(k, [bool(torch.equal(r[k], runs[0][k])) for r in runs[1:]])
sums=[]
for rep in range(4):
    step.forward(*args); torch.cuda.synchronize(); sums.append(step.sums.clone())
print("sums equal", [bool(torch.equal(s, sums[0])) for s in sums[1:]])
