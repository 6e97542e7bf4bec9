"""Localise an end-to-end gradient discrepancy: per sample, then per stage (dense fragment gradients of the shader
backward, rasterizer backward on identical upstream gradients).  Run on the GPU box."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hifihr_b200 as hf
from hifihr_b200 import ops, _lib as L
from hifihr_b200.mano_assets import load_mano
from oracle import p3d, pipeline as P

DEV = "cuda"
mano = load_mano()
B, S, K, T = 8, 224, 4, 512
lam = dict(texture=1.0, mrgb=1.0, ssim_tex=1.0, sil=1.0, iou=0.5)
inp = P.synthetic_inputs(B, S=S, seed=1234)
step = hf.FusedHandStep(B, image_size=S, faces_per_pixel=K, soft=True, texture_size=T, lambdas=lam, device=DEV)
tex = step.texture.detach().cpu().clone()
fcl, prp = p3d.ndc_intrinsics(inp["Ks"])
d = lambda t: t.to(DEV).contiguous()
ARGS = (d(inp["pose"]), d(inp["betas"]), d(-fcl), d(prp), d(inp["root_xyz"]), d(inp["light_dir"]), d(inp["light_color"]),
        d(inp["imgs"]), d(inp["segms_gt"].float()))     # kept alive: the step caches raw pointers
step.step(*ARGS)
torch.cuda.synchronize()
sel = step.p2f.cpu()
res = {}
for name, dt in (("o32", torch.float32), ("o64", torch.float64)):
    oi = {k: (v.clone().to(dt) if v.is_floating_point() else v.clone()) for k, v in inp.items()}
    for k in ("pose", "betas"):
        oi[k].requires_grad_(True)
    ro = P.render_path(mano, oi, tex.to(dt), image_size=S, K=K, blur_radius=step.blur, soft=True, dtype=dt, pix_to_face=sel)
    for k in ("verts_ndc", "verts_view"):
        ro[k].retain_grad()
    fr = ro["fragments"]
    for t in (fr.zbuf, fr.bary_coords, fr.dists):
        t.retain_grad()
    ro["image"].retain_grad()
    loss, _ = P.total_loss(ro, oi, lam, 1.0)
    loss.backward()
    res[name] = dict(pose=oi["pose"].grad, betas=oi["betas"].grad, ndc=ro["verts_ndc"].grad, view=ro["verts_view"].grad,
                     gz=fr.zbuf.grad, gb=fr.bary_coords.grad, gd=fr.dists.grad, gimg=ro["image"].grad, img=ro["image"].detach())
def re(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).abs().max() / max(1e-30, b.abs().max()))
o64, o32 = res["o64"], res["o32"]
print("per-sample pose-grad error gpu vs o64 | o32 vs o64 | max|g|")
for n in range(B):
    print(n, re(step.g_pose[n], o64["pose"][n]), re(o32["pose"][n], o64["pose"][n]), float(o64["pose"][n].abs().max()))
print("image      ", re(step.image, o64["img"]), re(o32["img"], o64["img"]))
print("g_image    ", re(step.g_image, o64["gimg"]), re(o32["gimg"], o64["gimg"]))
print("g_ndc (acc)", re(step.g_ndc, o64["ndc"]), re(o32["ndc"], o64["ndc"]))
# g_view in the product excludes the ndc->view part (applied in geom_bwd); the oracle's view grad includes everything
for n in range(B):
    print(" g_ndc sample", n, re(step.g_ndc[n], o64["ndc"][n]), re(o32["ndc"][n], o64["ndc"][n]), float(o64["ndc"][n].abs().max()))
# dense fragment gradients from the modular shade backward on the same inputs
sa = step._shade_args
gz, gb, gd = torch.empty_like(step.zbuf), torch.empty_like(step.bary), torch.empty_like(step.dists)
acc2 = torch.zeros_like(step.acc)
sb = L.HfrShadeBwdArgs(sa, L.ptr(step.g_image), L.ptr(gz), L.ptr(gb), L.ptr(gd), None, None, float(step.blur), 1, 1,
                       L.ptr(torch.zeros_like(step.g_view)), L.ptr(torch.zeros_like(step.g_vn)), L.ptr(torch.zeros_like(step.g_texture)),
                       L.ptr(torch.zeros_like(step.g_light_dir)), L.ptr(torch.zeros_like(step.g_light_color)), None, 0, 0)
L.call("hfr_shade_backward", sb)
torch.cuda.synchronize()
print("dense g_zbuf ", re(gz, o64["gz"]), re(o32["gz"], o64["gz"]))
print("dense g_bary ", re(gb, o64["gb"]), re(o32["gb"], o64["gb"]))
print("dense g_dists", re(gd, o64["gd"]), re(o32["gd"], o64["gd"]))
for name, g, key in (("gz", gz, "gz"), ("gd", gd, "gd")):
    diff = (g.cpu().double() - o64[key]).abs()
    idx = int(diff.argmax())
    n, rem = divmod(idx, S * S * K)
    y, rem = divmod(rem, S * K)
    x, k = divmod(rem, K)
    print(f"worst {name}: n={n} y={y} x={x} k={k} gpu={float(g[n,y,x,k])} o32={float(o32[key][n,y,x,k])} o64={float(o64[key][n,y,x,k])}")
    print("  p2f", step.p2f[n, y, x].tolist(), "z", step.zbuf[n, y, x].tolist(), "d", step.dists[n, y, x].tolist())
    print("  image gpu", step.image[n, y, x].tolist(), "o64", o64["img"][n, y, x].tolist())
    print("  g_image gpu", step.g_image[n, y, x].tolist(), "o64", o64["gimg"][n, y, x].tolist())
    print("  gz gpu", gz[n, y, x].tolist(), "o32", o32["gz"][n, y, x].tolist(), "o64", o64["gz"][n, y, x].tolist())
    print("  gd gpu", gd[n, y, x].tolist(), "o32", o32["gd"][n, y, x].tolist(), "o64", o64["gd"][n, y, x].tolist())
    print("  gb gpu", gb[n, y, x].tolist(), "o64", o64["gb"][n, y, x].tolist())
# how many fragments differ noticeably
for name, g, key in (("gz", gz, "gz"), ("gd", gd, "gd"), ("gb", gb, "gb")):
    ref = o64[key]
    diff = (g.cpu().double() - ref).abs()
    thr = 1e-3 * float(ref.abs().max())
    print(name, "fragments off by > 1e-3 max:", int((diff > thr).sum()), "of", int((ref != 0).sum()))
# the face made of the worst vertices of the worst sample
n = max(range(B), key=lambda i: re(step.g_ndc[i], o64["ndc"][i]))
diff = (step.g_ndc[n].cpu().double() - o64["ndc"][n]).abs().amax(1)
top = torch.topk(diff, 5)
print("worst sample", n, "verts", top.indices.tolist(), top.values.tolist())
faces = step.topo.faces_long.cpu()
vs = set(top.indices.tolist()[:3])
fid = [i for i in range(faces.shape[0]) if set(faces[i].tolist()) == vs]
print("face", fid, [faces[i].tolist() for i in fid])
for v in top.indices.tolist()[:3]:
    print("  v", v, "gpu", step.g_ndc[n, v].tolist(), "o64", o64["ndc"][n, v].tolist(), "ndc", step.verts_ndc[n, v].tolist())
if fid:
    f = fid[0] + n * 1538
    where = (step.p2f[n] == f).nonzero().cpu()
    print("fragments of that face:", where.shape[0])
    for (y, x, k) in where.tolist():
        print(f"  y={y} x={x} k={k} z={float(step.zbuf[n,y,x,k]):.7f} d={float(step.dists[n,y,x,k]):.3e} bary={step.bary[n,y,x,k].tolist()}",
              f"gz gpu {float(gz[n,y,x,k]):.4e} o64 {float(o64['gz'][n,y,x,k]):.4e} | gd gpu {float(gd[n,y,x,k]):.4e} o64 {float(o64['gd'][n,y,x,k]):.4e}",
              "| gb gpu", [f"{v:.3e}" for v in gb[n,y,x,k].tolist()], "o64", [f"{v:.3e}" for v in o64['gb'][n,y,x,k].tolist()])
    fv = step.face_verts[f].cpu()
    print("face_verts", [[repr(float(x)) for x in r] for r in fv])
