"""No-GPU checks of the kernels' per-element math (the *_math.cuh headers compiled for the host by
tests/host_emul) against the oracle.  Test-only: the product has no CPU path."""
import ctypes as C

import numpy as np
import pytest
import torch

from hifihr_b200 import _lib as L
from oracle import p3d, raster_c
from oracle import pipeline as P
from oracle.mano import rodrigues
from tests.host_emul import load

lib = load()
ptr = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731


def _scene(mano, B=1, S=48, K=4, blur=9.21e-4, soft=True, seed=7):
    inp = P.synthetic_inputs(B, S=S, seed=seed)
    tex = P.synthetic_texture(64)
    out = P.render_path(mano, inp, tex, image_size=S, aa=1, K=K, blur_radius=blur, soft=soft)
    return inp, tex, out


def test_rodrigues_fwd_bwd():
    g = torch.Generator().manual_seed(0)
    v = torch.randn(64, 3, generator=g)
    v[0] = 0
    v[1] = 1e-6
    R = rodrigues(v).reshape(64, 9).numpy()
    vn, Rn = v.numpy().copy(), np.zeros((64, 9), np.float32)
    lib.emul_rodrigues_fwd(ptr(vn), ptr(Rn), 64)
    assert np.abs(Rn - R).max() < 5e-7
    v64 = v.double().requires_grad_(True)
    gR = torch.randn(64, 3, 3, generator=g)
    (rodrigues(v64) * gR.double()).sum().backward()
    gRn, gvn = gR.reshape(64, 9).numpy().copy(), np.zeros((64, 3), np.float32)
    lib.emul_rodrigues_bwd(ptr(vn), ptr(gRn), ptr(gvn), 64)
    assert np.abs(gvn - v64.grad.numpy()).max() < 5e-6


@pytest.mark.parametrize("H,W,K,blur", [(64, 64, 1, 0.0), (40, 56, 4, 9.21e-4)])
def test_raster_math_bit_exact_and_grad(mano, H, W, K, blur):
    inp = P.synthetic_inputs(1, S=64, seed=3)
    out = P.render_path(mano, inp, P.synthetic_texture(16), image_size=64, K=1)
    faces = torch.tensor(np.asarray(mano["f"], np.int64))
    fv = out["verts_ndc"][:, faces].reshape(-1, 3, 3).contiguous()
    Fm = fv.shape[0]
    c = raster_c.rasterize_naive(fv, [0], [Fm], (H, W), blur, K)
    p2f = np.zeros((H, W, K), np.int64)
    zb, ba, ds = np.zeros((H, W, K), np.float32), np.zeros((H, W, K, 3), np.float32), np.zeros((H, W, K), np.float32)
    fvn = fv.numpy()
    lib.emul_raster(ptr(fvn), C.c_int64(Fm), H, W, K, C.c_float(blur), 1, int(blur > 0), 0, ptr(p2f), ptr(zb), ptr(ba), ptr(ds))
    assert (p2f == c[0][0].numpy()).all()
    assert (zb == c[1][0].numpy()).all() and (ba == c[2][0].numpy()).all() and (ds == c[3][0].numpy()).all()
    # backward against fp64 autograd of the oracle
    g = torch.Generator().manual_seed(1)
    f64 = fv.double().clone().requires_grad_(True)
    fr = p3d.rasterize_meshes(f64, [0], [Fm], (H, W), blur, K)
    if not (fr.pix_to_face[0].numpy() == p2f).all():
        pytest.skip("fp64 oracle picks different faces at a tie")
    gz, gb, gd = (torch.randn(fr.zbuf.shape, generator=g), torch.randn(fr.bary_coords.shape, generator=g),
                  torch.randn(fr.dists.shape, generator=g) * 1e-2)
    mk = (fr.pix_to_face >= 0).double()
    ((fr.zbuf * gz * mk).sum() + (fr.bary_coords * gb * mk[..., None]).sum() + (fr.dists * gd * mk).sum()).backward()
    gfv = np.zeros((Fm, 9), np.float32)
    lib.emul_raster_bwd(ptr(fvn), ptr(p2f), ptr(gz[0].numpy().copy()), ptr(gb[0].numpy().copy()),
                        ptr(gd[0].numpy().copy()), H, W, K, 1, int(blur > 0), ptr(gfv))
    ref = f64.grad.reshape(-1, 9).numpy()
    assert np.abs(gfv - ref).max() <= 1e-4 * max(1.0, np.abs(ref).max())


def _shade_args(mano, inp, tex, out, K, soft, keep):
    fr = out["fragments"]
    N, H, W, _ = fr.pix_to_face.shape
    faces = np.ascontiguousarray(np.asarray(mano["f"], np.int32))
    uvs, fuv = P.mano_uvs(mano)
    vn = p3d.vertex_normals(out["verts_view"], torch.tensor(faces.astype(np.int64)))
    arrs = dict(p2f=fr.pix_to_face.numpy().copy(), z=fr.zbuf.detach().numpy().copy(),
                b=fr.bary_coords.detach().numpy().copy(), d=fr.dists.detach().numpy().copy(), faces=faces,
                vv=out["verts_view"].detach().numpy().copy(), vn=vn.detach().numpy().copy(),
                fuv=fuv.numpy().astype(np.int32).copy(), uvs=uvs.numpy().copy(), tex=tex.detach().numpy().copy(),
                ld=inp["light_dir"].numpy().copy(), lc=inp["light_color"].numpy().copy(),
                img=np.zeros((N, H, W, 4), np.float32))
    keep.append(arrs)
    p = L.HfrShadeParams()
    p.N, p.H, p.W, p.K, p.F, p.V = N, H, W, K, faces.shape[0], 778
    p.blend, p.shade = (2 if soft else 0), 1
    p.sigma, p.gamma, p.znear, p.zfar = 1e-4, 1e-4, 1.0, 100.0
    p.background = L.f3((1, 1, 1)); p.light_ambient = L.f3((.5, .5, .5)); p.light_specular = L.f3((.2, .2, .2))
    p.mat_ambient = L.f3((1, 1, 1)); p.mat_diffuse = L.f3((.8, .8, .8)); p.mat_specular = L.f3((.2, .2, .2))
    p.shininess = 30.0
    p.tex_n, p.tex_h, p.tex_w, p.VT = 1, tex.shape[1], tex.shape[2], 778
    a = L.HfrShadeFwdArgs(p, *[arrs[k].ctypes.data for k in ("p2f", "z", "b", "d", "faces", "vv", "vn", "fuv", "uvs",
                                                              "tex", "ld", "lc", "img")])
    return a, arrs


@pytest.mark.parametrize("soft,K,blur", [(True, 4, 9.21e-4), (False, 1, 0.0)])
def test_shade_forward_matches_oracle(mano, soft, K, blur):
    inp, tex, out = _scene(mano, B=2, S=40, K=K, blur=blur, soft=soft)
    keep = []
    a, arrs = _shade_args(mano, inp, tex, out, K, soft, keep)
    lib.emul_shade_fwd(C.byref(a))
    ref = out["image"].detach().numpy()
    assert np.abs(arrs["img"] - ref).max() < 2e-5


def test_blend_backward_matches_autograd(mano):
    inp, tex, out = _scene(mano, B=1, S=40, K=4)
    fr = out["fragments"]
    Pn = 40 * 40
    g = torch.Generator().manual_seed(5)
    # fp32 oracle: with gamma = 1e-4 the exponent (z_inv - z_max)/gamma amplifies fp32 rounding of z_inv to
    # ~1e-3 relative, so the like-for-like comparison is against the fp32 autograd of the same formulas
    colors = torch.rand(1, 40, 40, 4, 3, generator=g).requires_grad_(True)
    z = fr.zbuf.detach().clone().requires_grad_(True)
    d = fr.dists.detach().clone().requires_grad_(True)
    frd = p3d.Fragments(fr.pix_to_face, z, fr.bary_coords.detach(), d)
    img = p3d.softmax_rgb_blend(colors, frd)
    grgba = torch.randn(img.shape, generator=g)
    (img * grgba).sum().backward()
    p = L.HfrShadeParams()
    p.K, p.blend, p.sigma, p.gamma, p.znear, p.zfar = 4, 2, 1e-4, 1e-4, 1.0, 100.0
    p.background = L.f3((1, 1, 1))
    idn = fr.pix_to_face.numpy().reshape(Pn, 4).copy()
    zn, dn = z.detach().float().numpy().reshape(Pn, 4).copy(), d.detach().float().numpy().reshape(Pn, 4).copy()
    cn = colors.detach().float().numpy().reshape(Pn, 4, 3).copy()
    gn = grgba.float().numpy().reshape(Pn, 4).copy()
    gc, gz, gd = np.zeros((Pn, 4, 3), np.float32), np.zeros((Pn, 4), np.float32), np.zeros((Pn, 4), np.float32)
    lib.emul_blend_bwd(C.byref(p), Pn, ptr(idn), ptr(zn), ptr(dn), ptr(cn), ptr(gn), ptr(gc), ptr(gz), ptr(gd))
    mk = (idn >= 0)
    rc = colors.grad.numpy().reshape(Pn, 4, 3) * mk[..., None]
    rz, rd = z.grad.numpy().reshape(Pn, 4), d.grad.numpy().reshape(Pn, 4)
    assert np.abs(gc - rc).max() < 1e-4 * max(1, np.abs(rc).max())
    assert np.abs(gd - rd).max() < 2e-3 * max(1, np.abs(rd).max())
    assert np.abs(gz - rz).max() < 2e-3 * max(1, np.abs(rz).max())


def test_phong_and_texture_backward_match_autograd():
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(11)
    M = 200
    Pp = (torch.randn(M, 3, generator=g, dtype=torch.float64) * 0.1 + torch.tensor([0, 0, 0.6])).requires_grad_(True)
    Nn = torch.randn(M, 3, generator=g, dtype=torch.float64).requires_grad_(True)
    ld = torch.randn(3, generator=g, dtype=torch.float64).requires_grad_(True)
    lc = (torch.rand(3, generator=g, dtype=torch.float64) * 0.8 + 0.2).requires_grad_(True)
    tx = torch.rand(M, 3, generator=g, dtype=torch.float64).requires_grad_(True)
    n_hat = F.normalize(Nn, eps=1e-6, dim=-1)
    d_hat = F.normalize(ld, eps=1e-6, dim=-1)
    cosang = (n_hat * d_hat).sum(-1)
    diffuse = lc * F.relu(cosang)[:, None]
    view = F.normalize(-Pp, eps=1e-6, dim=-1)
    refl = -d_hat + 2 * cosang[:, None] * n_hat
    alpha = F.relu((view * refl).sum(-1)) * (cosang > 0).double()
    spec = 0.2 * torch.pow(alpha, 30.0)[:, None]
    color = (0.5 + 0.8 * diffuse) * tx + 0.2 * spec
    gcol = torch.randn(M, 3, generator=g, dtype=torch.float64)
    (color * gcol).sum().backward()
    p = L.HfrShadeParams()
    p.light_ambient = L.f3((.5, .5, .5)); p.light_specular = L.f3((.2, .2, .2))
    p.mat_ambient = L.f3((1, 1, 1)); p.mat_diffuse = L.f3((.8, .8, .8)); p.mat_specular = L.f3((.2, .2, .2))
    p.shininess = 30.0
    f = lambda t: t.detach().float().numpy().copy()  # noqa: E731
    dh = f(d_hat)
    out, gP, gN, gT = (np.zeros((M, 3), np.float32) for _ in range(4))
    gdh, glc = np.zeros(3, np.float32), np.zeros(3, np.float32)
    lib.emul_phong(C.byref(p), M, ptr(f(Pp)), ptr(f(Nn)), ptr(dh), ptr(f(lc)), ptr(f(tx)), ptr(f(gcol)), ptr(out),
                   ptr(gP), ptr(gN), ptr(gT), ptr(gdh), ptr(glc))
    assert np.abs(out - f(color)).max() < 1e-5
    for got, ref in ((gP, Pp.grad), (gN, Nn.grad), (gT, tx.grad), (glc, lc.grad)):
        assert np.abs(got - ref.numpy()).max() < 1e-4 * max(1, ref.abs().max().item())
    # g_dhat -> g_dir through the normalisation
    dl = ld.detach().norm().item()
    gdir = (gdh - dh * (dh @ gdh)) / dl
    assert np.abs(gdir - ld.grad.numpy()).max() < 1e-4 * max(1, ld.grad.abs().max().item())
    # texture sampling
    Ht, Wt = 12, 16
    tex = torch.rand(1, Ht, Wt, 3, generator=g, dtype=torch.float64).requires_grad_(True)
    uv = (torch.rand(M, 2, generator=g, dtype=torch.float64) * 1.2 - 0.1).requires_grad_(True)
    grid = (uv * 2 - 1).view(1, M, 1, 2)
    tm = torch.flip(tex.permute(0, 3, 1, 2), [2])
    smp = F.grid_sample(tm, grid, mode="bilinear", align_corners=True, padding_mode="border")[0, :, :, 0].T
    gs = torch.randn(M, 3, generator=g, dtype=torch.float64)
    (smp * gs).sum().backward()
    o, guv, gtex = np.zeros((M, 3), np.float32), np.zeros((M, 2), np.float32), np.zeros((Ht, Wt, 3), np.float32)
    lib.emul_tex(ptr(f(tex)), Ht, Wt, M, ptr(f(uv)), ptr(f(gs)), ptr(o), ptr(guv), ptr(gtex))
    assert np.abs(o - f(smp)).max() < 1e-5
    assert np.abs(guv - uv.grad.numpy()).max() < 1e-3 * max(1, uv.grad.abs().max().item())
    assert np.abs(gtex - tex.grad[0].numpy()).max() < 1e-4 * max(1, tex.grad.abs().max().item())


def test_pca_texture_math_matches_autograd():
    """the texture PCA model evaluated at the bilinear taps (shade_math.cuh hfr_texel /
    hfr_tex_fetch / hfr_tex_uv_grad / hfr_tex_param_grad) against grid_sample on the composed map, fp64 autograd."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(23)
    M, Ht, Wt, npc = 150, 10, 14, 5
    mean = torch.rand(Ht, Wt, 3, generator=g, dtype=torch.float64)
    basis = torch.randn(npc, Ht, Wt, 3, generator=g, dtype=torch.float64) * 0.1
    params = torch.randn(npc, generator=g, dtype=torch.float64).requires_grad_(True)
    uv = (torch.rand(M, 2, generator=g, dtype=torch.float64) * 1.2 - 0.1).requires_grad_(True)
    tex = (mean + torch.einsum("k,khwc->hwc", params, basis))[None]
    tm = torch.flip(tex.permute(0, 3, 1, 2), [2])
    smp = F.grid_sample(tm, (uv * 2 - 1).view(1, M, 1, 2), mode="bilinear", align_corners=True, padding_mode="border")[0, :, :, 0].T
    gs = torch.randn(M, 3, generator=g, dtype=torch.float64)
    (smp * gs).sum().backward()
    f = lambda t: np.ascontiguousarray(t.detach().float().numpy())  # noqa: E731
    res = []
    for texel_major in (False, True):   # (npc,Ht,Wt,3), and the texel-major records of HfrShadeParams.tex_basis_stride
        stride = 12 * ((npc + 3) // 4) if texel_major else 0
        bs = f(basis)
        if texel_major:
            bs = np.zeros((Ht * Wt, stride), np.float32)
            bs[:, :3 * npc] = f(basis.permute(1, 2, 0, 3).reshape(Ht * Wt, 3 * npc))
            assert bs.ctypes.data % 16 == 0
        o, guv, gp = np.zeros((M, 3), np.float32), np.zeros((M, 2), np.float32), np.zeros(npc, np.float32)
        lib.emul_tex_pca(ptr(f(mean)), ptr(bs), ptr(f(params)), npc, Ht, Wt, M, ptr(f(uv)), ptr(f(gs)), ptr(o), ptr(guv), ptr(gp),
                         stride)
        assert np.abs(o - f(smp)).max() < 1e-5
        assert np.abs(guv - uv.grad.numpy()).max() < 1e-3 * max(1, uv.grad.abs().max().item())
        assert np.abs(gp - params.grad.numpy()).max() < 1e-4 * max(1, params.grad.abs().max().item())
        res.append((o, guv, gp))
    # the two layouts run the same sums in the same order
    assert all(np.array_equal(a, b) for a, b in zip(*res))
    # the backward's single visit of the taps (hfr_tex_fetch_d + hfr_tex_uv_grad_d): same texel bits, uv gradient re-associated
    o, guv, gp = np.zeros((M, 3), np.float32), np.zeros((M, 2), np.float32), np.zeros(npc, np.float32)
    lib.emul_tex_pca(ptr(f(mean)), ptr(bs), ptr(f(params)), npc, Ht, Wt, M, ptr(f(uv)), ptr(f(gs)), ptr(o), ptr(guv), ptr(gp), -stride)
    assert np.array_equal(o, res[1][0]) and np.array_equal(gp, res[1][2])
    assert np.abs(guv - uv.grad.numpy()).max() < 1e-3 * max(1, uv.grad.abs().max().item())


def test_edge_sign_filter_never_rejects_an_inside_pixel():
    """The K = 1 / blur_radius = 0 walk of the rasterizer rejects a candidate when one edge function is zero or differs in
    sign from the face area (raster_math.cuh hfr_edge_sign_outside) and only then skips the exact coverage math.  Fuzz of
    that claim on the host build of the same header: the filter never fires on a pair the exact math calls inside -
    random faces, pixels on edges and vertices, sliver / tiny / huge faces, denormal-sized edge functions, z down to 1e-6,
    both windings, with and without perspective correction."""
    g = np.random.default_rng(7)
    n = 400_000
    fv = g.uniform(-1.2, 1.2, (n, 3, 3)).astype(np.float32)
    fv[..., 2] = g.uniform(0.05, 3.0, (n, 3)).astype(np.float32)
    k = n // 8
    fv[k:2 * k, :, :2] *= np.float32(1e-3)                       # tiny faces around the origin
    fv[2 * k:3 * k, 2, :2] = fv[2 * k:3 * k, 0, :2] + (fv[2 * k:3 * k, 1, :2] - fv[2 * k:3 * k, 0, :2]) * np.float32(0.5) \
        + g.normal(0, 1e-7, (k, 2)).astype(np.float32)           # slivers: third vertex (almost) on the opposite edge
    fv[3 * k:4 * k, :, :2] *= np.float32(1e-18)                  # edge functions in the denormal range
    fv[4 * k:5 * k, :, 2] = g.uniform(1e-6, 1e-4, (k, 3)).astype(np.float32)   # depths next to the validity threshold
    fv[5 * k:6 * k, :, :2] = np.round(fv[5 * k:6 * k, :, :2] * 8) / 8          # vertices on a coarse grid ...
    pxy = g.uniform(-1.2, 1.2, (n, 2)).astype(np.float32)
    pxy[5 * k:6 * k] = np.round(pxy[5 * k:6 * k] * 8) / 8                      # ... and pixels on the same grid (exact zeros)
    pxy[k:2 * k] *= np.float32(1e-3)
    pxy[3 * k:4 * k] *= np.float32(1e-18)
    w = g.dirichlet([1, 1, 1], 2 * k).astype(np.float32)                       # pixels INSIDE their face (convex combinations)
    pxy[6 * k:8 * k] = np.einsum("nc,ncd->nd", w, fv[6 * k:8 * k, :, :2])
    pxy[7 * k:7 * k + 1000] = fv[7 * k:7 * k + 1000, 0, :2]                    # exactly on a vertex
    fv = np.ascontiguousarray(fv.reshape(n, 9))
    pxy = np.ascontiguousarray(pxy)
    for pc in (0, 1):
        out = np.zeros(n, np.uint8)
        lib.emul_sign_filter(ptr(fv), ptr(pxy), n, pc, ptr(out))
        assert not np.any(out == 3), f"filter rejected {int((out == 3).sum())} inside pairs (pc={pc})"
        # the test is not vacuous: it rejects most outside pairs and the inside class is well populated
        assert (out == 2).sum() > n // 10 and (out == 1).sum() > n // 4
        # what it lets through although the pixel is outside is rare (ties, underflow): the exact math then decides
        assert (out == 0).sum() < n // 20


def test_depth_cull_bound_holds_for_convex_combinations():
    """The fine pass skips a face whose nearest vertex, shrunk by 1e-5 (raster_tile.cuh: zmin * 0.99999), is not in front
    of a pixel's K-th depth, and stops at the first depth bucket with that property.  That is exact only if the depth the
    coverage math would produce can never be smaller: pz must be a convex combination of the vertex depths, up to fp32
    rounding.  Fuzz on the host build of the same math (slivers, depth ratios of 1e4 inside one face, pixels next to edges
    and vertices): the bound holds whenever the barycentrics are renormalised - perspective correction (every camera of the
    reference is a PerspectiveCameras: models_res_nimble.py:183-186) or clamping (blur > 0).

    KNOWN LIMITATION, pinned here: with perspective_correct=False AND blur_radius=0 PyTorch3D's barycentrics are e_i / (area +
    1e-8) and sum to area / (area + 1e-8) < 1, so a face of NDC area below ~1e-3 yields pz slightly BELOW a convex
    combination (a sliver of area 2e-8: half of it) and the cull could drop it where PyTorch3D keeps it.  No path of the
    reference rasterises that way; the functional `rasterize_meshes(..., perspective_correct=False,
    blur_radius=0)` does."""
    g = np.random.default_rng(11)
    n = 300_000
    fv = g.uniform(-1.2, 1.2, (n, 3, 3)).astype(np.float32)
    fv[..., 2] = np.exp(g.uniform(np.log(0.02), np.log(200.0), (n, 3))).astype(np.float32)   # depth ratios up to 1e4
    k = n // 4
    fv[:k, 2, :2] = fv[:k, 0, :2] + (fv[:k, 1, :2] - fv[:k, 0, :2]) * np.float32(0.3) + g.normal(0, 1e-6, (k, 2)).astype(np.float32)
    w = g.dirichlet([0.3, 0.3, 0.3], n).astype(np.float32)          # many pixels close to edges and vertices
    pxy = np.einsum("nc,ncd->nd", w, fv[:, :, :2]).astype(np.float32)
    pxy[k:2 * k] += g.normal(0, 0.05, (k, 2)).astype(np.float32)    # some outside: clamped barycentrics
    fvf = np.ascontiguousarray(fv.reshape(n, 9))
    pxy = np.ascontiguousarray(pxy)
    zmin = fv[..., 2].min(1)
    # faces whose doubled area is within fp32 noise of the validity threshold (|area| <= 1e-8 is invalid) are left out: their
    # edge functions are rounding noise of magnitude 1e-8 .. 1e-7, the reference's own depth there is garbage (all clamped
    # coordinates 0 -> pz = 0)
    d = fv.astype(np.float64)
    fa = (d[:, 2, 0] - d[:, 0, 0]) * (d[:, 1, 1] - d[:, 0, 1]) - (d[:, 2, 1] - d[:, 0, 1]) * (d[:, 1, 0] - d[:, 0, 0])
    regular = np.abs(fa) >= 1e-6

    def violations(pc, clip):
        pz, fl = np.zeros(n, np.float32), np.zeros(n, np.uint8)
        lib.emul_pair_depth(ptr(fvf), ptr(pxy), n, pc, clip, ptr(pz), ptr(fl))
        sel = regular & (fl & 1).astype(bool) & ((fl & 4) != 0) & (((fl & 2) != 0) | bool(clip))     # pairs the cull is applied to
        assert sel.sum() > n // 4
        return sel & ~(pz >= np.float32(0.99999) * zmin)

    for pc, clip in ((1, 0), (1, 1), (0, 1)):
        bad = violations(pc, clip)
        assert not bad.any(), (pc, clip, int(bad.sum()))
    bad = violations(0, 0)
    assert bad.any() and bad[:k].sum() == bad.sum()        # only the slivers of the first block, never a regular face of this set
