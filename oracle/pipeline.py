"""TEST INFRASTRUCTURE — the whole hot path on the CPU (torch autograd), assembled from the
oracle pieces in the reference's order (models_res_nimble.py:133-223 + losses.py:355-408):

  ManoLayer -> xyz_from_vertice -> root shift -> NDC camera -> rasterize -> Phong / UV texture ->
  blend -> avg-pool / split -> losses.

Also provides the synthetic inputs of SURVEY.md §8(d) so tests, smoke() and bench.py draw the
same tensors.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

from . import losses as olosses
from . import p3d
from .mano import ManoOracle


def synthetic_inputs(B, S=224, seed=1234, dtype=torch.float32):
    """poses, shapes, camera, lights, target images and masks (CPU tensors)."""
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.rand(*s, generator=g)  # noqa: E731
    n = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
    pose = torch.cat([n(B, 3) * 1.5, n(B, 45) * 0.5], 1)
    betas = n(B, 10) * 0.5
    root_xyz = torch.stack([r(B) * 0.06 - 0.03, r(B) * 0.06 - 0.03, r(B) * 0.2 + 0.55], 1)
    f = r(B) * 80 + 440
    c = 112 + r(B, 2) * 16 - 8
    Ks = torch.zeros(B, 3, 4)
    Ks[:, 0, 0] = f
    Ks[:, 1, 1] = f
    Ks[:, 0, 2] = c[:, 0]
    Ks[:, 1, 2] = c[:, 1]
    Ks[:, 2, 2] = 1
    imgs = r(B, 3, S, S)
    yy, xx = torch.meshgrid(torch.arange(S), torch.arange(S), indexing="ij")
    disc = (((yy - S / 2 + 0.5) ** 2 + (xx - S / 2 + 0.5) ** 2) <= (0.3 * S) ** 2).long()
    seg = disc[None].repeat(B, 1, 1)
    light_color = r(B, 3) * 0.8 + 0.2
    light_dir = n(B, 3)
    out = dict(pose=pose, betas=betas, root_xyz=root_xyz, Ks=Ks, imgs=imgs, segms_gt=seg,
               light_color=light_color, light_dir=light_dir)
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in out.items()}


def mano_uvs(mano):
    xz = np.asarray(mano["v_template"], np.float64)[:, [0, 2]]
    uv = (xz - xz.min(0)) / (xz.max(0) - xz.min(0))
    return torch.tensor(uv.astype(np.float32)), torch.tensor(np.asarray(mano["f"], np.int64))


def synthetic_texture(T=512, seed=20231):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(1, T, T, 3, generator=g)


def render_path(mano, inp, texture, *, image_size=224, aa=1, K=1, blur_radius=0.0, soft=False, sigma=1e-4,
                gamma=1e-4, binarize=False, root_id=9, dtype=torch.float32, mano_oracle=None, c_select=False,
                threads=1, pix_to_face=None, point_light=False):
    """Returns dict with verts, joints, verts_view, verts_ndc, fragments, image (N,H,W,4), re_img, re_sil.
    pix_to_face: use this face selection instead of searching (three-way precision tests give the fp32 and the fp64
    evaluation the SAME fragments, so only arithmetic differs)."""
    orc = mano_oracle or ManoOracle(mano, dtype=dtype)
    verts, jtr = orc(inp["pose"], inp["betas"])
    joints = orc.xyz_from_vertice(verts)
    pred_root = joints[:, root_id:root_id + 1]
    joints_rel = joints - pred_root
    verts_rel = verts - pred_root
    view = (verts + (-pred_root)) + inp["root_xyz"][:, None]
    fcl, prp = p3d.ndc_intrinsics(inp["Ks"])
    ndc = p3d.project_ndc(view, -fcl, prp)
    faces = orc.faces
    B, Fm = verts.shape[0], faces.shape[0]
    S = image_size * aa
    fv = ndc[:, faces].reshape(-1, 3, 3)
    first = [i * Fm for i in range(B)]
    nf = [Fm] * B
    sel = pix_to_face
    if c_select and sel is None:   # CPU baseline: the scalar C rasterizer does the O(P*F) search, torch differentiates the winners
        from . import raster_c
        sel = raster_c.rasterize_naive(fv, first, nf, S, blur_radius, K, perspective_correct=True, threads=threads)[0]
    fr = p3d.rasterize_meshes(fv, first, nf, S, blur_radius, K, perspective_correct=True, pix_to_face=sel)
    uvs, fuv = mano_uvs(mano)
    texels = p3d.sample_textures_uv(fr, texture.to(dtype), fuv, uvs.to(dtype))
    colors = p3d.phong_shading(fr, view, faces, texels, inp["light_dir"], inp["light_color"], point_light=point_light)
    if soft:
        image = p3d.softmax_rgb_blend(colors, fr, sigma, gamma)
    else:
        image = p3d.hard_rgb_blend(colors, fr)
    img = image.permute(0, 3, 1, 2)
    if aa > 1:
        img = F.avg_pool2d(img, kernel_size=aa, stride=aa)
    re_img = img[:, :3]
    re_sil = img[:, 3:4]
    if binarize:
        re_sil = torch.where(re_sil > 0, torch.full_like(re_sil, 255.0), re_sil).detach()
    return dict(verts=verts, jtr=jtr, joints=joints_rel, verts_rel=verts_rel, verts_view=view, verts_ndc=ndc,
                fragments=fr, image=image, re_img=re_img, re_sil=re_sil)


def total_loss(out, inp, lambdas, sil_scale):
    terms = olosses.render_losses(out["re_img"], out["re_sil"], inp["imgs"], inp["segms_gt"], lambdas,
                                  sil_scale=sil_scale)
    return sum(terms.values()), terms
