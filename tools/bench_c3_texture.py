"""C3-shaped measurement (BASELINE config 3: NIMBLE-shaped hand, V~5990, 1024^2 diffuse texture, batch 128, 256x256,
forward+backward) of the render with the texture PCA model sampled inside the shader kernels (TexturesUVPCA)
against the same render with the per-sample maps materialised first (library GEMM + TexturesUV).
usage (GPU box): python tools/bench_c3_texture.py [B] [S] [T]   -> one JSON line"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import hifihr_b200 as hf  # noqa: E402
from hifihr_b200.nimble import MyNIMBLELayer  # noqa: E402
from hifihr_b200.synthetic import synthetic_inputs  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
S = int(sys.argv[2]) if len(sys.argv) > 2 else 256
T = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
DEV = "cuda"
g = torch.Generator().manual_seed(5)
pose = torch.cat([torch.randn(B, 3, generator=g) * 0.4, torch.randn(B, 30, generator=g) * 0.5], 1).to(DEV).requires_grad_(True)
shape = (torch.randn(B, 20, generator=g) * 0.5).to(DEV).requires_grad_(True)
texp = torch.randn(B, 10, generator=g).to(DEV).requires_grad_(True)
inp = synthetic_inputs(B, S=32, seed=6)
root = torch.tensor([[0.0, 0.0, 0.45]]).repeat(B, 1).to(DEV)
fcl, prp = hf.get_ndc_fx_fy_cx_cy(inp["Ks"])
cams = hf.PerspectiveCameras(focal_length=-fcl.to(DEV), principal_point=prp.to(DEV), device=DEV)
lights = hf.DirectionalLights(diffuse_color=inp["light_color"].to(DEV), direction=inp["light_dir"].to(DEV), device=DEV)
rs = hf.RasterizationSettings(image_size=S, blur_radius=0.0, faces_per_pixel=1)
mats = hf.Materials(diffuse_color=((0.8, 0.8, 0.8),), specular_color=((0.2, 0.2, 0.2),), shininess=30, device=DEV)
renderer = hf.MeshRenderer(rasterizer=hf.MeshRasterizer(raster_settings=rs), shader=hf.HardPhongShader(materials=mats, device=DEV))
gimg = torch.randn(B, S, S, 4, generator=g).to(DEV)
res = {}
for fused in (True, False):
    layer = MyNIMBLELayer(True, DEV, shape_ncomp=20, pose_ncomp=30, tex_ncomp=10, tex_size=T, fused_texture=fused).to(DEV)

    def step():
        for t in (pose, shape, texp):
            t.grad = None
        out = layer({"pose_params": pose, "shape_params": shape, "texture_params": texp}, handle_collision=False)
        meshes = out["skin_meshes"]
        meshes.offset_verts_(root[:, None].repeat(1, layer.V, 1).view(B * layer.V, 3))
        img = renderer(meshes, cameras=cams, lights=lights)
        (img * gimg).sum().backward()
        return img

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        img = step()
    e1.record()
    torch.cuda.synchronize()
    res["fused" if fused else "materialised"] = {"ms_per_step": e0.elapsed_time(e1) / n,
                                                  "samples_per_s": B * n / (e0.elapsed_time(e1) * 1e-3),
                                                  "peak_mem_GB": torch.cuda.max_memory_allocated() / 1e9,
                                                  "coverage": float((img[..., 3] > 0).float().mean()),
                                                  "g_texp_absmax": float(texp.grad.abs().max())}
    del layer
    torch.cuda.empty_cache()
print(json.dumps({"workload": f"C3-shaped: NIMBLE-like V=5986 F=11968, B={B}, {S}^2, K=1 HardPhong, texture PCA 10 x {T}^2, fwd+bwd "
                              "(modular path: layer -> rasterize -> shade, autograd bridges)", **res}))
