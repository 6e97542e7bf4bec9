"""GPU parity tests of the late round-2 changes (through the C-ABI, against the CPU oracle):

* the K = 1 / blur_radius == 0 walk of the rasterizer's fine pass (exact edge-sign rejection before the coverage math):
  Fragments stay bit-exact where pixel centres sit EXACTLY on edges and vertices, for both windings, for a tile that
  holds every face of the mesh, and at the NIMBLE-shaped face count;
* FaceVertsFunction (hfr_face_verts_forward / _backward) against ATen indexing and its autograd;
* the texel-major texture PCA basis against the component-major one (same bits);
* ShadeFunction computing only the gradients autograd asks for.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import p3d, raster_c  # noqa: E402
from oracle import pipeline as P  # noqa: E402
from oracle.mano import ManoOracle  # noqa: E402

DEV = "cuda"
NAMES = ["pix_to_face", "zbuf", "bary_coords", "dists"]


@pytest.fixture(scope="module")
def hf():
    import hifihr_b200
    assert os.path.isfile(hifihr_b200.LIB_PATH), "libhifihr_b200.so missing: the CUDA path is the only path"
    return hifihr_b200


def _check_bit_exact(hf, fv, first, nf, size, K, blur, pc, cull=False):
    ref = raster_c.rasterize_naive(fv, first, nf, size, blur, K, perspective_correct=pc, cull_backfaces=cull, threads=8)
    out = hf.rasterize_meshes(fv.to(DEV), size, blur, K, perspective_correct=pc, clip_barycentric_coords=blur > 0,
                              cull_backfaces=cull, mesh_to_face_first_idx=torch.tensor(first, device=DEV),
                              num_faces_per_mesh=torch.tensor(nf, device=DEV))
    for name, g, r in zip(NAMES, out, ref):
        bad = g.cpu() != r
        assert not bad.any(), f"{name}: {int(bad.sum())} mismatching entries of {bad.numel()}"
    return out


@pytest.mark.parametrize("pc", [True, False])
def test_hard_k1_pixel_centres_on_edges_and_vertices(hf, pc):
    """Triangles whose vertices ARE pixel centres (S = 16: NDC centres (2 i + 1) / 16 - 1 are exact in fp32): many edge
    functions are exactly zero, which the strict `> 0` inside rule must reject, in both windings and depth orders."""
    S = 16
    c = lambda i: (2.0 * i + 1.0) / S - 1.0  # noqa: E731
    g = torch.Generator().manual_seed(5)
    tris = []
    for _ in range(60):
        ij = torch.randint(0, S, (3, 2), generator=g)
        z = 1.0 + torch.rand(3, generator=g) * 2.0
        t = [[c(int(ij[k, 0])), c(int(ij[k, 1])), float(z[k])] for k in range(3)]
        tris.append(t)
        tris.append([t[0], t[2], t[1]])          # the same triangle with the other winding (area < 0)
    # axis-aligned right triangles sharing a diagonal: the diagonal's pixels belong to NEITHER
    tris.append([[c(2), c(2), 1.5], [c(12), c(2), 1.5], [c(2), c(12), 1.5]])
    tris.append([[c(12), c(12), 1.5], [c(2), c(12), 1.5], [c(12), c(2), 1.5]])
    fv = torch.tensor(tris, dtype=torch.float32)
    for K in (1, 2):
        _check_bit_exact(hf, fv, [0], [fv.shape[0]], S, K, 0.0, pc)
    _check_bit_exact(hf, fv, [0], [fv.shape[0]], S, 1, 0.0, pc, cull=True)


def _mano_face_verts(mano, B, seed, S=64):
    inp = P.synthetic_inputs(B, S=S, seed=seed)
    orc = ManoOracle(mano)
    verts, _ = orc(inp["pose"], inp["betas"])
    joints = orc.xyz_from_vertice(verts)
    view = (verts - joints[:, 9:10]) + inp["root_xyz"][:, None]
    fcl, prp = p3d.ndc_intrinsics(inp["Ks"])
    ndc = p3d.project_ndc(view, -fcl, prp)
    fv = ndc[:, orc.faces].reshape(-1, 3, 3).contiguous()
    Fm = orc.faces.shape[0]
    return fv, [i * Fm for i in range(B)], [Fm] * B


@pytest.mark.parametrize("scale,size", [(0.08, 64), (0.3, (40, 72)), (1.0, 224)])
def test_hard_k1_dense_tiles(hf, mano, scale, size):
    """K = 1, blur 0 with every face of the hand inside one or two tiles (several staged batches per tile), a non-square
    image, and the C1 size."""
    fv, first, nf = _mano_face_verts(mano, 2, seed=31)
    fv = fv.clone()
    fv[..., :2] = fv[..., :2] * scale
    out = _check_bit_exact(hf, fv, first, nf, size, 1, 0.0, True)
    assert (out[0] >= 0).any()


def test_hard_k1_nimble_sized_mesh(hf):
    """The NIMBLE-shaped stand-in (11 968 faces, two 8192-face coarse chunks per tile) at K = 1, blur 0."""
    from hifihr_b200.nimble import MyNIMBLELayer
    layer = MyNIMBLELayer(True, DEV, shape_ncomp=20, pose_ncomp=30, tex_ncomp=10, tex_size=64).to(DEV)
    g = torch.Generator().manual_seed(3)
    B = 2
    pose = torch.cat([torch.randn(B, 3, generator=g) * 0.4, torch.randn(B, 30, generator=g) * 0.5], 1).to(DEV)
    shape = (torch.randn(B, 20, generator=g) * 0.5).to(DEV)
    out = layer({"pose_params": pose, "shape_params": shape, "texture_params": torch.zeros(B, 10, device=DEV)}, handle_collision=False)
    meshes = out["skin_meshes"]
    meshes.offset_verts_(torch.tensor([[0.0, 0.0, 0.45]], device=DEV).repeat(B * layer.V, 1))
    inp = P.synthetic_inputs(B, S=96, seed=4)
    fcl, prp = hf.get_ndc_fx_fy_cx_cy(inp["Ks"])
    cams = hf.PerspectiveCameras(focal_length=-fcl.to(DEV), principal_point=prp.to(DEV), device=DEV)
    rast = hf.MeshRasterizer(raster_settings=hf.RasterizationSettings(image_size=96, blur_radius=0.0, faces_per_pixel=1))
    _, ndc = rast.transform(meshes, cameras=cams)
    fv = ndc.reshape(-1, 3)[meshes.faces_packed()].cpu().contiguous()
    Fn = layer.F
    out = _check_bit_exact(hf, fv, [0, Fn], [Fn, Fn], 96, 1, 0.0, True)
    assert 0.02 < (out[0] >= 0).float().mean() < 0.8


def test_face_verts_function_matches_indexing(hf, mano):
    from hifihr_b200 import ops
    from hifihr_b200.structures import topology_for
    faces = torch.tensor(np.asarray(mano["f"], np.int64))
    V = int(np.asarray(mano["v_template"]).shape[0])
    topo = topology_for(faces, V, torch.device(DEV))
    g = torch.Generator().manual_seed(9)
    verts = torch.randn(3, V, 3, generator=g).to(DEV).requires_grad_(True)
    fv = ops.FaceVertsFunction.apply(topo, verts)
    off = (torch.arange(3, device=DEV) * V).view(-1, 1, 1)
    ref_in = verts.detach().clone().requires_grad_(True)
    ref = ref_in.reshape(-1, 3)[(faces.to(DEV)[None] + off).reshape(-1, 3)]
    assert fv.shape == ref.shape == (3 * faces.shape[0], 3, 3) and torch.equal(fv, ref)
    gout = torch.randn(fv.shape, generator=g).to(DEV)
    fv.backward(gout)
    ref.backward(gout)
    err = float((verts.grad - ref_in.grad).abs().max() / ref_in.grad.abs().max())
    assert err < 1e-6, err
    # fixed summation order: bit-reproducible
    v2 = verts.detach().clone().requires_grad_(True)
    ops.FaceVertsFunction.apply(topo, v2).backward(gout)
    assert torch.equal(v2.grad, verts.grad)


def test_texel_major_basis_matches_component_major(hf, mano):
    """hfr_shade_forward / hfr_shade_backward with the packed (T*T, 12*ceil(n/4)) basis against the (n,T,T,3) one: the same
    sums in the same order - identical image, d/d(coefficients) within atomics noise."""
    from hifihr_b200 import ops
    from hifihr_b200 import _lib as L
    B, S, T, n = 2, 64, 32, 10
    fv, first, nf = _mano_face_verts(mano, B, seed=12)
    frags = hf.rasterize_meshes(fv.to(DEV), S, 0.0, 1, perspective_correct=True,
                                mesh_to_face_first_idx=torch.tensor(first, device=DEV), num_faces_per_mesh=torch.tensor(nf, device=DEV))
    orc = ManoOracle(mano)
    faces = orc.faces.to(DEV).to(torch.int32).contiguous()
    V, Fm = int(np.asarray(mano["v_template"]).shape[0]), faces.shape[0]
    g = torch.Generator().manual_seed(2)
    verts_view = (torch.randn(B, V, 3, generator=g) * 0.05 + torch.tensor([0.0, 0.0, 0.5])).to(DEV)
    vn = torch.nn.functional.normalize(torch.randn(B, V, 3, generator=g), dim=-1).to(DEV)
    uvs = torch.rand(V, 2, generator=g).to(DEV)
    mean = torch.rand(1, T, T, 3, generator=g).to(DEV)
    basis = (torch.randn(n, T, T, 3, generator=g) * 0.1).to(DEV)
    tp = torch.randn(B, n, generator=g).to(DEV)
    ld = torch.tensor([[0.0, 0.3, -1.0]] * B, device=DEV)
    lc = torch.full((B, 3), 0.7, device=DEV)
    g_img = torch.randn(B, S, S, 4, generator=g).to(DEV)
    res = []
    for packed in (False, True):
        bs = ops.pack_tex_basis(basis) if packed else basis
        p = ops.shade_params(B, S, S, 1, Fm, V, 0, 1, 1e-4, 1e-4, (1, 1, 1), (.5, .5, .5), (.2, .2, .2), (1, 1, 1),
                             (.8, .8, .8), (.2, .2, .2), 30.0, tex_shape=mean.shape[:3], VT=V, tex_pca=n,
                             tex_basis_stride=bs.shape[1] if packed else 0)
        t = tp.clone().requires_grad_(True)
        img = ops.ShadeFunction.apply(p, frags[0], frags[1], frags[2], frags[3], faces, verts_view, vn, faces, uvs, mean,
                                      ld, lc, bs, t)
        (img * g_img).sum().backward()
        res.append((img.detach(), t.grad))
    assert ops.pack_tex_basis(basis).shape == (T * T, 36) and ops.pack_tex_basis(basis) is ops.pack_tex_basis(basis)
    assert torch.equal(res[0][0], res[1][0])
    err = float((res[0][1] - res[1][1]).abs().max() / res[0][1].abs().max())
    assert err < 1e-5, err
    with pytest.raises(ValueError):          # a stride that does not match 12 * ceil(n / 4)
        p.tex_basis_stride = 32
        ops.ShadeFunction.apply(p, frags[0], frags[1], frags[2], frags[3], faces, verts_view, vn, faces, uvs, mean, ld, lc,
                                ops.pack_tex_basis(basis), tp)
    assert L.HfrShadeParams.tex_basis_stride.offset > L.HfrShadeParams.light_point.offset


def test_shade_backward_only_computes_requested_gradients(hf, mano):
    """A frozen texture / frozen lights get no gradient tensor (and no reductions in the kernel); the other gradients
    are unchanged."""
    model = hf.HandRenderModel(device=DEV, image_size=48, aa_factor=1, faces_per_pixel=2, soft=True, binarize=False,
                               blur_radius=9.21e-4, texture_size=32).to(DEV)
    inp = P.synthetic_inputs(2, S=48, seed=8)
    grads = []
    for frozen in (False, True):
        model.texture.requires_grad_(not frozen)
        model.texture.grad = None
        pose = inp["pose"].to(DEV).requires_grad_(True)
        lcol = inp["light_color"].to(DEV).requires_grad_(not frozen)
        out = model({"pose_params": pose, "shape_params": inp["betas"].to(DEV)},
                    light_params={"colors": lcol, "directions": inp["light_dir"].to(DEV)}, Ks=inp["Ks"].to(DEV),
                    root_xyz=inp["root_xyz"].to(DEV))
        (out["re_img"].square().sum() + out["re_sil"].sum()).backward()
        grads.append(pose.grad.clone())
        assert (model.texture.grad is None) == frozen and (lcol.grad is None) == frozen
    err = float((grads[0] - grads[1]).abs().max() / grads[0].abs().max())
    assert err < 1e-4, err


def test_fused_nimble_step_matches_modular_path(hf):
    """FusedNimbleStep (raw launches, rasterizer backward fused into the shading backward, texel-major texture PCA) against
    the modular autograd path on the same inputs - which test_c3_shaped_losses_and_gradients pins to the oracle: identical
    pix_to_face, loss terms to 1e-5, every gradient to 1e-4 (summation order of the atomics); a CUDA-graph replay of the
    step reproduces the eager one."""
    from hifihr_b200 import ops
    B, T, S = 3, 256, 128
    step = hf.FusedNimbleStep(B, image_size=S, texture_size=T, device=DEV)
    layer = step.layer
    V = layer.V
    g = torch.Generator().manual_seed(21)
    pose = torch.cat([torch.randn(B, 3, generator=g) * 0.4, torch.randn(B, 30, generator=g) * 0.5], 1).to(DEV)
    shape = (torch.randn(B, 20, generator=g) * 0.5).to(DEV)
    texp = torch.randn(B, 10, generator=g).to(DEV)
    inp = P.synthetic_inputs(B, S=S, seed=6)
    root = torch.tensor([[0.0, 0.0, 0.45]]).repeat(B, 1).to(DEV)
    fcl, prp = hf.get_ndc_fx_fy_cx_cy(inp["Ks"])
    focal, prp = (-fcl).to(DEV).contiguous(), prp.to(DEV).contiguous()
    ldir, lcol = inp["light_dir"].to(DEV), inp["light_color"].to(DEV)
    imgs, seg = inp["imgs"].to(DEV), inp["segms_gt"].float().to(DEV)
    # ---- modular path ---------------------------------------------------------------------------
    leaves = [t.clone().requires_grad_(True) for t in (pose, shape, texp)]
    lc = lcol.clone().requires_grad_(True)
    out = layer({"pose_params": leaves[0], "shape_params": leaves[1], "texture_params": leaves[2]}, handle_collision=False)
    cams = hf.PerspectiveCameras(focal_length=focal, principal_point=prp, device=DEV)
    lights = hf.DirectionalLights(diffuse_color=lc, direction=ldir, device=DEV)
    rs = hf.RasterizationSettings(image_size=S, blur_radius=0.0, faces_per_pixel=1)
    mats = hf.Materials(diffuse_color=((0.8, 0.8, 0.8),), specular_color=((0.2, 0.2, 0.2),), shininess=30, device=DEV)
    meshes = out["skin_meshes"]
    meshes.offset_verts_(root[:, None].repeat(1, V, 1).view(B * V, 3))
    frags = hf.MeshRasterizer(raster_settings=rs)(meshes, cameras=cams)
    img = hf.HardPhongShader(materials=mats, device=DEV)(frags, meshes, cameras=cams, lights=lights)
    re_img, re_sil, _ = ops.PoolFunction.apply(img, 1, False, None)
    t = ops.RenderLossFunction.apply(re_img, re_sil, imgs, seg, 1.0, True)
    (t[0] + t[1] + t[2]).backward()
    assert (frags.pix_to_face >= 0).float().mean() > 0.02
    # ---- fused step -----------------------------------------------------------------------------
    args = (pose, shape, focal, prp, root, ldir, lcol, imgs, seg)
    step.step(*args, tex_params=texp)
    torch.cuda.synchronize()
    assert torch.equal(step.p2f, frags.pix_to_face) and torch.equal(step.zbuf, frags.zbuf)
    assert float((step.image - img.detach()).abs().max()) < 1e-6
    terms = step.loss_terms()
    for i in range(3):
        assert abs(float(terms[i]) - float(t[i])) < 1e-5 * max(1.0, abs(float(t[i]))), i
    got = {"pose": step.g_pose.clone(), "shape": step.g_betas.clone(), "tex": step.g_tex_params.clone(),
           "lcol": step.g_light_color.clone()}
    for name, a, b in (("pose", got["pose"], leaves[0].grad), ("shape", got["shape"], leaves[1].grad),
                       ("tex", got["tex"], leaves[2].grad), ("lcol", got["lcol"], lc.grad)):
        err = float((a - b).abs().max() / b.abs().max())
        print("fused nimble", name, err)
        assert err < 1e-4, (name, err)
    with pytest.raises(ValueError):
        step.step(*args)                       # the texture coefficients are a required input of this step
    # ---- graph replay == eager ------------------------------------------------------------------
    graph = step.capture(*args, tex_params=texp)
    step.g_pose.zero_()
    graph.replay()
    torch.cuda.synchronize()
    for name, a in (("pose", step.g_pose), ("shape", step.g_betas), ("tex", step.g_tex_params)):
        err = float((a - got[name]).abs().max() / got[name].abs().max())
        assert err < 1e-5, (name, err)
