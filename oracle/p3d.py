"""TEST INFRASTRUCTURE — CPU restatement (torch + autograd) of the PyTorch3D pieces on the hot path.

PARITY UNPINNED: PyTorch3D is a third-party dependency of the reference
(README.md:70-71, un-pinned git HEAD), absent from /root/reference and from this
image.  This file restates its published algorithm as summarised in SURVEY.md
Appendix A (upstream files mirrored: renderer/mesh/rasterize_meshes.py,
csrc/rasterize_meshes/rasterize_meshes_cpu.cpp, csrc/utils/geometry_utils.h,
renderer/mesh/shading.py, renderer/lighting.py, renderer/blending.py,
renderer/mesh/textures.py, structures/meshes.py, ops/interp_face_attrs.py).
Reference call sites that make each piece relevant: models_res_nimble.py:70-96
(settings, HardPhongShader, Materials), :183-198 (camera, lights), :203-211
(offset_verts_, render, avg-pool).

Everything is evaluated op by op in the dtype of the inputs (no fused
multiply-add on the CPU path), so fp32 results are reproducible bit for bit by
a kernel that uses the same operation order without FMA contraction.
"""
from __future__ import annotations

from typing import NamedTuple, Optional

import torch
import torch.nn.functional as F

K_EPS = 1e-8


class Fragments(NamedTuple):
    pix_to_face: torch.Tensor
    zbuf: torch.Tensor
    bary_coords: torch.Tensor
    dists: torch.Tensor


# ----------------------------------------------------------------------------- camera (A.1)
def ndc_intrinsics(Ks: torch.Tensor):
    """models_res_nimble.py:228-235 (hard-coded 224/112)."""
    fx = Ks[:, 0, 0] * 2 / 224.0
    fy = Ks[:, 1, 1] * 2 / 224.0
    px = -(Ks[:, 0, 2] - 112.0) * 2 / 224.0
    py = -(Ks[:, 1, 2] - 112.0) * 2 / 224.0
    return torch.stack([fx, fy], -1), torch.stack([px, py], -1)


def project_ndc(verts_view: torch.Tensor, focal: torch.Tensor, prp: torch.Tensor):
    """PerspectiveCameras(in_ndc) with R=I, T=0: p_h = [X,Y,Z,1]·K, xy = p_h.xy / p_h.w, and
    MeshRasterizer.transform overwrites z with view depth.  `focal` is what the camera
    was given (the reference passes -fcl, models_res_nimble.py:184)."""
    X, Y, Z = verts_view.unbind(-1)
    fx, fy = focal[:, 0:1], focal[:, 1:2]
    px, py = prp[:, 0:1], prp[:, 1:2]
    x = (fx * X + px * Z) / Z
    y = (fy * Y + py * Z) / Z
    return torch.stack([x, y, Z], -1)


# ----------------------------------------------------------------------------- pixel grid (A.3)
def pix_to_ndc(i: torch.Tensor, S1: int, S2: int, dtype):
    rng = (2.0 * S1) / S2 if S1 > S2 else 2.0
    rng_t = torch.tensor(rng, dtype=dtype)
    off = rng_t / 2.0
    return -off + (rng_t * i.to(dtype) + off) / torch.tensor(float(S1), dtype=dtype)


def pixel_centers(H: int, W: int, dtype=torch.float32):
    yi = torch.arange(H)
    xi = torch.arange(W)
    yf = pix_to_ndc(H - 1 - yi, H, W, dtype)
    xf = pix_to_ndc(W - 1 - xi, W, H, dtype)
    return yf, xf


# ----------------------------------------------------------------------------- per (pixel, face) math (A.4)
def _edge(px, py, ax, ay, bx, by):
    return (px - ax) * (by - ay) - (py - ay) * (bx - ax)


def _seg_dist2(px, py, ax, ay, bx, by):
    bax, bay = bx - ax, by - ay
    l2 = bax * bax + bay * bay
    degenerate = l2 <= K_EPS
    l2s = torch.where(degenerate, torch.ones_like(l2), l2)
    t = (bax * (px - ax) + bay * (py - ay)) / l2s
    t = t.clamp(0.0, 1.0)
    qx, qy = ax + t * bax, ay + t * bay
    dx, dy = qx - px, qy - py
    d = dx * dx + dy * dy
    ex, ey = px - bx, py - by
    return torch.where(degenerate, ex * ex + ey * ey, d)


def face_pixel_terms(px, py, v, blur_radius: float, perspective_correct: bool, clip_bary: bool,
                     cull_backfaces: bool = False):
    """v: (...,3,3) NDC face verts; px,py broadcastable to v[...,0,0].
    Returns pz, bary_clip (...,3), signed_dist, valid (bool)."""
    x0, y0, z0 = v[..., 0, 0], v[..., 0, 1], v[..., 0, 2]
    x1, y1, z1 = v[..., 1, 0], v[..., 1, 1], v[..., 1, 2]
    x2, y2, z2 = v[..., 2, 0], v[..., 2, 1], v[..., 2, 2]
    r = float(torch.tensor(blur_radius, dtype=v.dtype).sqrt())
    xmin = torch.minimum(torch.minimum(x0, x1), x2) - r
    xmax = torch.maximum(torch.maximum(x0, x1), x2) + r
    ymin = torch.minimum(torch.minimum(y0, y1), y2) - r
    ymax = torch.maximum(torch.maximum(y0, y1), y2) + r
    zmin = torch.minimum(torch.minimum(z0, z1), z2)
    outside = (px < xmin) | (px > xmax) | (py < ymin) | (py > ymax) | (zmin < K_EPS)
    face_area = _edge(x0, y0, x1, y1, x2, y2)          # EdgeFunction(v0, v1, v2)
    zero_area = (face_area <= K_EPS) & (face_area >= -K_EPS)
    valid = ~outside & ~zero_area
    if cull_backfaces:
        valid = valid & ~(face_area < 0)
    area = _edge(x2, y2, x0, y0, x1, y1) + K_EPS        # EdgeFunction(v2, v0, v1) + eps
    area = torch.where(valid, area, torch.ones_like(area))
    b0 = _edge(px, py, x1, y1, x2, y2) / area
    b1 = _edge(px, py, x2, y2, x0, y0) / area
    b2 = _edge(px, py, x0, y0, x1, y1) / area
    if perspective_correct:
        t0 = b0 * z1 * z2
        t1 = z0 * b1 * z2
        t2 = z0 * z1 * b2
        den = (t0 + t1 + t2).clamp(min=K_EPS)
        b0, b1, b2 = t0 / den, t1 / den, t2 / den
    if clip_bary:
        c0, c1, c2 = b0.clamp(0.0, 1.0), b1.clamp(0.0, 1.0), b2.clamp(0.0, 1.0)
        s = (c0 + c1 + c2).clamp(min=1e-5)
        c0, c1, c2 = c0 / s, c1 / s, c2 / s
    else:
        c0, c1, c2 = b0, b1, b2
    pz = c0 * z0 + c1 * z1 + c2 * z2
    valid = valid & ~(pz < 0)
    e01 = _seg_dist2(px, py, x0, y0, x1, y1)
    e02 = _seg_dist2(px, py, x0, y0, x2, y2)
    e12 = _seg_dist2(px, py, x1, y1, x2, y2)
    pick01 = (e01 <= e02) & (e01 <= e12)
    pick02 = (e02 <= e01) & (e02 <= e12)
    dist = torch.where(pick01, e01, torch.where(pick02, e02, e12))
    inside = (b0 > 0) & (b1 > 0) & (b2 > 0)
    sdist = torch.where(inside, -dist, dist)
    valid = valid & (inside | (dist < blur_radius))
    return pz, torch.stack([c0, c1, c2], -1), sdist, valid


def rasterize_meshes(face_verts: torch.Tensor, mesh_to_face_first_idx, num_faces_per_mesh,
                     image_size, blur_radius: float = 0.0, faces_per_pixel: int = 1,
                     perspective_correct: bool = True, clip_barycentric_coords: Optional[bool] = None,
                     cull_backfaces: bool = False, pixel_chunk: int = 2048,
                     pix_to_face: Optional[torch.Tensor] = None) -> Fragments:
    """Naive O(P·F) rasterizer (A.2-A.4) with autograd through zbuf / bary / dists (A.5).

    A no-grad dense pass picks, per pixel, the K smallest (z, face index); the
    selected (pixel, k) pairs are then re-evaluated differentiably with the same
    formulas, so forward values are those of the dense pass."""
    H, W = (image_size, image_size) if isinstance(image_size, int) else image_size
    K = faces_per_pixel
    if clip_barycentric_coords is None:
        clip_barycentric_coords = blur_radius > 0
    dt = face_verts.dtype
    N = len(num_faces_per_mesh)
    yf, xf = pixel_centers(H, W, dt)
    PY = yf[:, None].expand(H, W).reshape(-1)
    PX = xf[None, :].expand(H, W).reshape(-1)
    P = H * W
    p2f = torch.full((N, P, K), -1, dtype=torch.int64)
    fv_det = face_verts.detach()
    if pix_to_face is not None:      # selection already done (e.g. by the scalar C oracle): only differentiate
        p2f = pix_to_face.reshape(N, P, K).clone()
    for n in range(N if pix_to_face is None else 0):
        f0, nf = int(mesh_to_face_first_idx[n]), int(num_faces_per_mesh[n])
        if nf == 0:
            continue
        v = fv_det[f0:f0 + nf][None]                       # (1,F,3,3)
        for s in range(0, P, pixel_chunk):
            e = min(P, s + pixel_chunk)
            pz, _, _, valid = face_pixel_terms(PX[s:e, None], PY[s:e, None], v, blur_radius,
                                               perspective_correct, clip_barycentric_coords,
                                               cull_backfaces)
            key = torch.where(valid, pz, torch.full_like(pz, float("inf")))
            zs, idx = torch.sort(key, dim=1, stable=True)  # ties -> smaller face index first
            kk = min(K, nf)
            sel = idx[:, :kk] + f0
            sel = torch.where(torch.isinf(zs[:, :kk]), torch.full_like(sel, -1), sel)
            p2f[n, s:e, :kk] = sel
    mask = p2f >= 0
    safe = p2f.clamp(min=0)
    v_sel = face_verts[safe]                               # (N,P,K,3,3) differentiable gather
    pz, bary, sd, _ = face_pixel_terms(PX[None, :, None], PY[None, :, None], v_sel, blur_radius,
                                       perspective_correct, clip_barycentric_coords, cull_backfaces)
    neg = torch.full_like(pz, -1.0)
    zbuf = torch.where(mask, pz, neg)
    dists = torch.where(mask, sd, neg)
    bary = torch.where(mask[..., None], bary, neg[..., None].expand_as(bary))
    return Fragments(p2f.view(N, H, W, K), zbuf.view(N, H, W, K), bary.view(N, H, W, K, 3),
                     dists.view(N, H, W, K))


# ----------------------------------------------------------------------------- attributes / textures (A.6, A.7)
def interpolate_face_attributes(pix_to_face, bary, face_attrs):
    """(N,H,W,K), (N,H,W,K,3), (ΣF,3,D) -> (N,H,W,K,D); zero where pix_to_face < 0."""
    mask = pix_to_face < 0
    idx = pix_to_face.clamp(min=0)
    a = face_attrs[idx]                                    # (N,H,W,K,3,D)
    out = (bary[..., None] * a).sum(-2)
    return out.masked_fill(mask[..., None], 0.0)


def sample_textures_uv(fragments: Fragments, maps: torch.Tensor, faces_uvs: torch.Tensor,
                       verts_uvs: torch.Tensor):
    """TexturesUV.sample_textures.  maps (N or 1,Ht,Wt,C); faces_uvs (F,3) per mesh (same for
    all meshes); verts_uvs (Vt,2).  Returns (N,H,W,K,C)."""
    N, H, W, K = fragments.pix_to_face.shape
    Fm = faces_uvs.shape[0]
    fvu = verts_uvs[faces_uvs]                             # (F,3,2)
    fvu = fvu[None].expand(N, Fm, 3, 2).reshape(N * Fm, 3, 2)
    uv = interpolate_face_attributes(fragments.pix_to_face, fragments.bary_coords, fvu)
    uv = uv.permute(0, 3, 1, 2, 4).reshape(N * K, H, W, 2)
    if maps.shape[0] == 1 and N > 1:
        maps = maps.expand(N, -1, -1, -1)
    C = maps.shape[-1]
    tm = maps.permute(0, 3, 1, 2)[None].expand(K, -1, -1, -1, -1).transpose(0, 1)
    tm = tm.reshape(N * K, C, maps.shape[1], maps.shape[2])
    grid = uv * 2.0 - 1.0
    tm = torch.flip(tm, [2])
    tex = F.grid_sample(tm, grid, mode="bilinear", align_corners=True, padding_mode="border")
    return tex.reshape(N, K, C, H, W).permute(0, 3, 4, 1, 2)


def vertex_normals(verts: torch.Tensor, faces: torch.Tensor):
    """Meshes.verts_normals: area-weighted face normals summed per vertex, normalised (eps 1e-6).
    verts (N,V,3), faces (F,3) shared."""
    v = verts[:, faces]                                    # (N,F,3,3)
    n1 = torch.cross(v[:, :, 2] - v[:, :, 1], v[:, :, 0] - v[:, :, 1], dim=-1)
    n2 = torch.cross(v[:, :, 0] - v[:, :, 2], v[:, :, 1] - v[:, :, 2], dim=-1)
    n0 = torch.cross(v[:, :, 1] - v[:, :, 0], v[:, :, 2] - v[:, :, 0], dim=-1)
    out = torch.zeros_like(verts)
    out = out.index_add(1, faces[:, 1], n1)
    out = out.index_add(1, faces[:, 2], n2)
    out = out.index_add(1, faces[:, 0], n0)
    return F.normalize(out, eps=1e-6, dim=-1)


# ----------------------------------------------------------------------------- lighting (A.8)
def phong_shading(fragments: Fragments, verts_view, faces, texels, light_dir, light_diffuse,
                  light_ambient=(0.5, 0.5, 0.5), light_specular=(0.2, 0.2, 0.2),
                  mat_ambient=(1.0, 1.0, 1.0), mat_diffuse=(0.8, 0.8, 0.8),
                  mat_specular=(0.2, 0.2, 0.2), shininess: float = 30.0, cam_center=None, point_light=False):
    """phong_shading + _apply_lighting + DirectionalLights.diffuse/specular.
    verts_view (N,V,3); faces (F,3); texels (N,H,W,K,3); light_dir/diffuse (N,3).
    point_light=True restates PointLights.diffuse/specular (renderer/lighting.py; the branch of
    models_res_nimble.py:191-198): light_dir is then the light LOCATION and the direction of every
    shaded point is location - point."""
    N, V, _ = verts_view.shape
    dt = verts_view.dtype
    Fm = faces.shape[0]
    t3 = lambda c: torch.as_tensor(c, dtype=dt).view(1, 1, 1, 1, 3)  # noqa: E731
    vn = vertex_normals(verts_view, faces)
    fverts = verts_view[:, faces].reshape(N * Fm, 3, 3)
    fnorm = vn[:, faces].reshape(N * Fm, 3, 3)
    pts = interpolate_face_attributes(fragments.pix_to_face, fragments.bary_coords, fverts)
    nrm = interpolate_face_attributes(fragments.pix_to_face, fragments.bary_coords, fnorm)
    d = light_dir.view(N, 1, 1, 1, 3)
    if point_light:
        d = d - pts
    col = light_diffuse.view(N, 1, 1, 1, 3)
    n_hat = F.normalize(nrm, p=2, dim=-1, eps=1e-6)
    d_hat = F.normalize(d, p=2, dim=-1, eps=1e-6)
    cosang = (n_hat * d_hat).sum(-1)
    diffuse = col * F.relu(cosang)[..., None]
    mask = (cosang > 0).to(dt)
    cc = torch.zeros(3, dtype=dt) if cam_center is None else cam_center
    view = F.normalize(cc.view(1, 1, 1, 1, 3) - pts, p=2, dim=-1, eps=1e-6)
    refl = -d_hat + 2 * (cosang[..., None] * n_hat)
    alpha = F.relu((view * refl).sum(-1)) * mask
    specular = t3(light_specular) * torch.pow(alpha, shininess)[..., None]
    ambient = t3(mat_ambient) * t3(light_ambient)
    diffuse = t3(mat_diffuse) * diffuse
    specular = t3(mat_specular) * specular
    return (ambient + diffuse) * texels + specular


# ----------------------------------------------------------------------------- blending (A.8, A.9)
def hard_rgb_blend(colors, fragments: Fragments, background=(1.0, 1.0, 1.0)):
    is_bg = fragments.pix_to_face[..., 0] < 0
    bg = torch.as_tensor(background, dtype=colors.dtype)
    rgb = torch.where(is_bg[..., None], bg, colors[..., 0, :])
    alpha = (~is_bg).to(colors.dtype)[..., None]
    return torch.cat([rgb, alpha], -1)


def sigmoid_alpha_blend(colors, fragments: Fragments, sigma=1e-4):
    mask = (fragments.pix_to_face >= 0).to(colors.dtype)
    prob = torch.sigmoid(-fragments.dists / sigma) * mask
    alpha = torch.prod(1.0 - prob, dim=-1)
    return torch.cat([colors[..., 0, :], (1.0 - alpha)[..., None]], -1)


def softmax_rgb_blend(colors, fragments: Fragments, sigma=1e-4, gamma=1e-4,
                      background=(1.0, 1.0, 1.0), znear: float = 1.0, zfar: float = 100.0):
    dt = colors.dtype
    eps = 1e-10
    mask = (fragments.pix_to_face >= 0).to(dt)
    bg = torch.as_tensor(background, dtype=dt)
    prob = torch.sigmoid(-fragments.dists / sigma) * mask
    alpha = torch.prod(1.0 - prob, dim=-1)
    z_inv = (zfar - fragments.zbuf) / (zfar - znear) * mask
    z_inv_max = torch.max(z_inv, dim=-1).values[..., None].clamp(min=eps)
    w = prob * torch.exp((z_inv - z_inv_max) / gamma)
    delta = torch.exp((eps - z_inv_max) / gamma).clamp(min=eps)
    denom = w.sum(-1)[..., None] + delta
    wc = (w[..., None] * colors).sum(-2)
    rgb = (wc + delta * bg) / denom
    return torch.cat([rgb, (1.0 - alpha)[..., None]], -1)
