"""GPU parity tests: every kernel, through the C-ABI, against the CPU oracle on the same seeded inputs.

Tolerances (fp32 path): verts 1e-6 m abs; pix_to_face / zbuf / bary / dists BIT-EXACT against the
scalar C oracle; images 2e-5 abs; losses 1e-5 rel; gradients 2e-3 of the tensor's max magnitude vs the fp32
oracle autograd (accumulation-order noise of fp32 sums over ~10^4 fragments).
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import losses as olosses  # noqa: E402
from oracle import p3d, raster_c  # noqa: E402
from oracle import pipeline as P  # noqa: E402
from oracle.mano import ManoOracle  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DEV = "cuda"


def rel_err(got, ref):
    ref = ref.detach().cpu().double()
    got = got.detach().cpu().double()
    return float((got - ref).abs().max() / max(1e-12, ref.abs().max()))


@pytest.fixture(scope="module")
def hf():
    import hifihr_b200
    assert os.path.isfile(hifihr_b200.LIB_PATH), "libhifihr_b200.so missing: the CUDA path is the only path"
    from hifihr_b200 import _lib
    assert _lib.lib().hfr_device_ok() == 1, "expected a compute-capability 10.x device"
    return hifihr_b200


# ------------------------------------------------------------------------------------------ MANO
def test_mano_matches_reference_golden(hf):
    z = np.load(os.path.join(GOLD, "mano_reference.npz"))
    layer = hf.ManoLayer(center_idx=9, flat_hand_mean=False, side="right", use_pca=True, ncomps=48)
    pose = torch.tensor(z["pose"], device=DEV, requires_grad=True)
    beta = torch.tensor(z["beta"], device=DEV, requires_grad=True)
    v, j = layer(pose, beta)
    assert (v.detach().cpu() - torch.tensor(z["verts"])).abs().max() < 1e-6
    assert (j.detach().cpu() - torch.tensor(z["joints"])).abs().max() < 1e-6
    ((v * torch.tensor(z["g_verts"], device=DEV)).sum() + (j * torch.tensor(z["g_joints"], device=DEV)).sum()).backward()
    assert rel_err(pose.grad, torch.tensor(z["g_pose"])) < 1e-3
    assert rel_err(beta.grad, torch.tensor(z["g_beta"])) < 1e-3
    v0, j0 = layer(torch.zeros(1, 48, device=DEV), torch.zeros(1, 10, device=DEV))
    assert (v0.cpu() - torch.tensor(z["verts_zero"])).abs().max() < 1e-6
    assert j0[0, 9].abs().max() == 0


def test_mano_modes_match_reference_golden(hf):
    """rot6d / robust rot6d root, rotmat joints, axis-angle without PCA, root_palm, share_betas, th_trans and the
    mean-shape default (my_mano.py:341-385, 459-461, 471-478) against golden vectors of the unmodified reference.
    verts/joints 1e-6 m abs (rotmat: 2e-6, a batched device SVD replaces the reference's per-matrix CPU SVD);
    gradients 1e-3 of the tensor's max magnitude."""
    from oracle.gen_golden import MODE_CASES
    z = np.load(os.path.join(GOLD, "mano_modes_reference.npz"))
    for name, (ckw, _, fkw) in MODE_CASES.items():
        kw = dict(ncomps=48, flat_hand_mean=False, center_idx=9)
        kw.update(ckw)
        layer = hf.ManoLayer(**kw)
        pose = torch.tensor(z[f"{name}.pose"], device=DEV, requires_grad=True)
        beta = torch.tensor(z[f"{name}.beta"], device=DEV, requires_grad=True)
        fw = {}
        if fkw.get("root_palm"):
            fw["root_palm"] = torch.Tensor([1])
        if fkw.get("share_betas"):
            fw["share_betas"] = torch.Tensor([1])
        trans = None
        if fkw.get("trans"):
            trans = torch.tensor(z[f"{name}.trans"], device=DEV, requires_grad=True)
            fw["th_trans"] = trans
        v, j = layer(pose, torch.zeros(1) if fkw.get("mean_shape") else beta, **fw)
        tol = 2e-6 if name == "rotmat" else 1e-6
        assert (v.detach().cpu() - torch.tensor(z[f"{name}.verts"])).abs().max() < tol, name
        assert (j.detach().cpu() - torch.tensor(z[f"{name}.joints"])).abs().max() < tol, name
        ((v * torch.tensor(z[f"{name}.g_verts"], device=DEV)).sum()
         + (j * torch.tensor(z[f"{name}.g_joints"], device=DEV)).sum()).backward()
        assert rel_err(pose.grad, torch.tensor(z[f"{name}.g_pose"])) < 1e-3, name
        if f"{name}.g_beta" in z.files:
            assert rel_err(beta.grad, torch.tensor(z[f"{name}.g_beta"])) < 1e-3, name
        if trans is not None:
            assert rel_err(trans.grad, torch.tensor(z[f"{name}.g_trans"])) < 1e-3, name


def test_mano_vs_oracle_many_samples(hf, mano):
    orc = ManoOracle(mano)
    layer = hf.ManoLayer(center_idx=9, flat_hand_mean=False, ncomps=48)
    inp = P.synthetic_inputs(257, S=8, seed=21)
    v, j = layer(inp["pose"].to(DEV), inp["betas"].to(DEV))
    # truth = the fp64 oracle (the fp32 CPU oracle's own error depends on the host's BLAS: 5e-8 here, 1e-5 seen
    # on one GPU box); tolerance 1e-6 m abs as SURVEY.md §8d states
    o64 = ManoOracle(mano, dtype=torch.float64)
    vo, jo = o64(inp["pose"].double(), inp["betas"].double())
    assert (v.cpu().double() - vo).abs().max() < 1e-6 and (j.cpu().double() - jo).abs().max() < 1e-6
    # mean shape (th_betas omitted) and explicit translation
    v2, j2 = layer(inp["pose"][:4].to(DEV))
    vo2, _ = orc(inp["pose"][:4], torch.zeros(4, 10))
    assert (v2.cpu() - vo2).abs().max() < 1e-6
    tr = torch.randn(4, 3)
    v3, j3 = layer(inp["pose"][:4].to(DEV), inp["betas"][:4].to(DEV), tr.to(DEV))
    vo3, jo3 = orc(inp["pose"][:4], inp["betas"][:4], trans=tr)
    assert (v3.cpu() - vo3).abs().max() < 2e-6 and (j3.cpu() - jo3).abs().max() < 2e-6
    # empty batch
    ve, je = layer(torch.zeros(0, 48, device=DEV), torch.zeros(0, 10, device=DEV))
    assert ve.shape == (0, 778, 3) and je.shape == (0, 21, 3)


def test_mano_state_dict_names(hf):
    layer = hf.ManoLayer(center_idx=9, flat_hand_mean=False, ncomps=48)
    names = set(layer.state_dict().keys())
    for k in ("th_betas", "th_shapedirs", "th_posedirs", "th_v_template", "th_J_regressor", "th_weights", "th_faces",
              "th_hands_mean", "th_comps", "th_selected_comps"):
        assert k in names
    assert layer.th_selected_comps.shape == (45, 45) and layer.th_faces.shape == (1538, 3)
    assert layer.kintree_parents[1:] == [0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11, 0, 13, 14]


# ------------------------------------------------------------------------------------------ NIMBLE-shaped layer
def test_nimble_layer_contract_and_lbs_parity(hf):
    """MyNIMBLELayer: the reference's call contract (models_res_nimble.py:55-57,133-142) on the
    seeded NIMBLE-shaped stand-in; LBS at V=5986 / 20 joints vs the generic fp64 LBS oracle (1e-6 m abs), gradients
    1e-3 rel, per-sample PCA texture, and a HardPhong render of its Meshes against the oracle renderer."""
    from hifihr_b200.nimble import MyNIMBLELayer
    from oracle.lbs import LBSOracle
    B, T, S = 3, 32, 48
    layer = MyNIMBLELayer(True, DEV, shape_ncomp=20, pose_ncomp=30, tex_ncomp=10, tex_size=T).to(DEV)
    d = layer._d
    g = torch.Generator().manual_seed(77)
    pose = torch.cat([torch.randn(B, 3, generator=g) * 0.4 + torch.tensor([0.0, 0.0, 0.0]),
                      torch.randn(B, 30, generator=g) * 0.5], 1)
    shape = torch.randn(B, 20, generator=g) * 0.5
    texp = torch.randn(B, 10, generator=g)
    leaves = [t.clone().to(DEV).requires_grad_(True) for t in (pose, shape, texp)]
    out = layer({"pose_params": leaves[0], "shape_params": leaves[1], "texture_params": leaves[2]}, handle_collision=False)
    V, Fn = layer.V, layer.F
    assert out["verts"].shape == (B, V, 3) and out["nimble_joints"].shape == (B, 25, 3) and out["joints"].shape == (B, 21, 3)
    assert out["mano_verts"].shape == (B, 778, 3) and out["textures"].shape == (B, T, T, 3) and out["rot"].shape == (B, 3)
    assert out["faces"].shape == (Fn, 3) and V == 5986 and Fn == 11968
    with pytest.raises(NotImplementedError):
        layer({"pose_params": leaves[0], "shape_params": leaves[1]}, handle_collision=True)
    orc = LBSOracle(d["v_template"], d["shapedirs"], d["posedirs"], d["J_regressor"], d["weights"], d["parents"],
                    pca_comps=d["pca_comps"], pose_mean=d["pose_mean"], tip_verts=d["tip_verts"], dtype=torch.float64)
    po, so = pose.double().requires_grad_(True), shape.double().requires_grad_(True)
    vo, jo = orc(po, so)
    assert (out["verts"].detach().cpu().double() - vo).abs().max() < 1e-6
    assert (out["nimble_joints"].detach().cpu().double() - jo).abs().max() < 1e-6
    gv, gj = torch.randn(B, V, 3, generator=g), torch.randn(B, 25, 3, generator=g)
    ((vo * gv.double()).sum() + (jo * gj.double()).sum()).backward()
    ((out["verts"] * gv.to(DEV)).sum() + (out["nimble_joints"] * gj.to(DEV)).sum()).backward(retain_graph=True)
    assert rel_err(leaves[0].grad, po.grad) < 1e-3 and rel_err(leaves[1].grad, so.grad) < 1e-3
    # per-sample texture = mean + params @ basis
    tex_o = d["tex_mean"][None] + torch.einsum("bk,khwc->bhwc", texp, d["tex_basis"])
    assert (out["textures"].detach().cpu() - tex_o).abs().max() < 1e-5
    # render the layer's Meshes (per-sample TexturesUV) with HardPhong, K=1, and compare with the oracle renderer
    inp = P.synthetic_inputs(B, S=S, seed=5)
    root = torch.tensor([[0.0, 0.0, 0.45]]).repeat(B, 1)
    fcl, prp = p3d.ndc_intrinsics(inp["Ks"])
    cams = hf.PerspectiveCameras(focal_length=-fcl.to(DEV), principal_point=prp.to(DEV), device=DEV)
    lights = hf.DirectionalLights(diffuse_color=inp["light_color"].to(DEV), direction=inp["light_dir"].to(DEV), device=DEV)
    rs = hf.RasterizationSettings(image_size=S, blur_radius=0.0, faces_per_pixel=1)
    mats = hf.Materials(diffuse_color=((0.8, 0.8, 0.8),), specular_color=((0.2, 0.2, 0.2),), shininess=30, device=DEV)
    renderer = hf.MeshRenderer(rasterizer=hf.MeshRasterizer(raster_settings=rs), shader=hf.HardPhongShader(materials=mats, device=DEV))
    meshes = out["skin_meshes"]
    meshes.offset_verts_(root.to(DEV)[:, None].repeat(1, V, 1).view(B * V, 3))
    img = renderer(meshes, cameras=cams, lights=lights)
    view = vo.float() + root[:, None]
    ndc = p3d.project_ndc(view, -fcl, prp)
    faces = torch.tensor(d["faces"])
    fv = ndc[:, faces].reshape(-1, 3, 3)
    fr = p3d.rasterize_meshes(fv, [i * Fn for i in range(B)], [Fn] * B, S, 0.0, 1, perspective_correct=True,
                              pix_to_face=raster_c.rasterize_naive(fv, [i * Fn for i in range(B)], [Fn] * B, S, 0.0, 1)[0])
    texels = p3d.sample_textures_uv(fr, tex_o, faces, torch.tensor(d["verts_uvs"]))
    img_o = p3d.hard_rgb_blend(p3d.phong_shading(fr, view, faces, texels, inp["light_dir"], inp["light_color"]), fr)
    diff = (img.detach().cpu() - img_o).abs().amax(-1)
    assert (diff > 1e-4).float().mean() < 5e-3      # last-bit vertex differences can flip pixels that sit on an edge
    assert ((img[..., 3] > 0).float().mean() > 0.02)
    img[..., :3].sum().backward()
    assert leaves[2].grad is not None and leaves[2].grad.abs().max() > 0


@pytest.mark.parametrize("K,soft", [(1, False), (3, True)])
def test_pca_texture_sampled_in_the_shader(hf, K, soft):
    """SURVEY 8(f) row 4, second half: the NIMBLE-style texture model (mean + params @ basis) evaluated at the
    bilinear taps inside the shader kernels (TexturesUVPCA) against the same render with the per-sample maps
    materialised first (TexturesUV) and against oracle autograd for d/d(texture params): images 2e-5 abs,
    gradients 1e-3 of the tensor's max."""
    from hifihr_b200.nimble import MyNIMBLELayer
    B, T, S = 3, 32, 48
    g = torch.Generator().manual_seed(91)
    pose = torch.cat([torch.randn(B, 3, generator=g) * 0.4, torch.randn(B, 30, generator=g) * 0.5], 1).to(DEV)
    shape = (torch.randn(B, 20, generator=g) * 0.5).to(DEV)
    texp = torch.randn(B, 10, generator=g)
    inp = P.synthetic_inputs(B, S=S, seed=6)
    root = torch.tensor([[0.0, 0.0, 0.45]]).repeat(B, 1).to(DEV)
    fcl, prp = p3d.ndc_intrinsics(inp["Ks"])
    cams = hf.PerspectiveCameras(focal_length=-fcl.to(DEV), principal_point=prp.to(DEV), device=DEV)
    lights = hf.DirectionalLights(diffuse_color=inp["light_color"].to(DEV), direction=inp["light_dir"].to(DEV), device=DEV)
    blur = 9.21e-4 if soft else 0.0
    rs = hf.RasterizationSettings(image_size=S, blur_radius=blur, faces_per_pixel=K)
    mats = hf.Materials(diffuse_color=((0.8, 0.8, 0.8),), specular_color=((0.2, 0.2, 0.2),), shininess=30, device=DEV)
    shader = (hf.SoftPhongShader if soft else hf.HardPhongShader)(materials=mats, device=DEV)
    renderer = hf.MeshRenderer(rasterizer=hf.MeshRasterizer(raster_settings=rs), shader=shader)
    gimg = torch.randn(B, S, S, 4, generator=g).to(DEV)
    imgs, grads = [], []
    for fused in (True, False):
        layer = MyNIMBLELayer(True, DEV, shape_ncomp=20, pose_ncomp=30, tex_ncomp=10, tex_size=T, fused_texture=fused).to(DEV)
        tp = texp.clone().to(DEV).requires_grad_(True)
        out = layer({"pose_params": pose, "shape_params": shape, "texture_params": tp}, handle_collision=False)
        assert (out["textures"] is None) == fused
        assert isinstance(out["skin_meshes"].textures, hf.TexturesUVPCA) == fused
        meshes = out["skin_meshes"]
        meshes.offset_verts_(root[:, None].repeat(1, layer.V, 1).view(B * layer.V, 3))
        img = renderer(meshes, cameras=cams, lights=lights)
        (img * gimg).sum().backward()
        imgs.append(img.detach())
        grads.append(tp.grad.clone())
        if fused:   # the export path materialises the same maps the unfused layer returns
            maps = meshes.textures.maps_padded()
            d = layer._d
            ref_maps = d["tex_mean"][None] + torch.einsum("bk,khwc->bhwc", texp, d["tex_basis"])
            assert (maps.detach().cpu() - ref_maps).abs().max() < 1e-5
    assert (imgs[0] - imgs[1]).abs().max() < 2e-5
    assert (imgs[0][..., 3] > 0).float().mean() > 0.02
    assert rel_err(grads[0], grads[1]) < 1e-3 and grads[0].abs().max() > 0
    # oracle autograd for d(image)/d(texture params) on the kernel's own Fragments
    d = layer._d
    with torch.no_grad():
        fr_dev = renderer.rasterizer(meshes, cameras=cams)
    fr = p3d.Fragments(fr_dev.pix_to_face.cpu(), fr_dev.zbuf.cpu(), fr_dev.bary_coords.cpu(), fr_dev.dists.cpu()) \
        if hasattr(p3d, "Fragments") else fr_dev
    tpo = texp.clone().requires_grad_(True)
    tex_o = d["tex_mean"][None] + torch.einsum("bk,khwc->bhwc", tpo, d["tex_basis"])
    faces = torch.tensor(d["faces"])
    texels = p3d.sample_textures_uv(fr, tex_o, faces, torch.tensor(d["verts_uvs"]))
    view = meshes.verts_padded().detach().cpu()
    colors = p3d.phong_shading(fr, view, faces, texels, inp["light_dir"], inp["light_color"])
    img_o = p3d.softmax_rgb_blend(colors, fr, 1e-4, 1e-4) if soft else p3d.hard_rgb_blend(colors, fr)
    (img_o * gimg.cpu()).sum().backward()
    assert (imgs[0].cpu() - img_o.detach()).abs().max() < 2e-5
    assert rel_err(grads[0], tpo.grad) < 1e-3


# ------------------------------------------------------------------------------------------ geometry
def test_geometry_forward_backward(hf, mano):
    from hifihr_b200 import ops
    orc = ManoOracle(mano)
    B = 5
    inp = P.synthetic_inputs(B, S=8, seed=4)
    verts_cpu, _ = orc(inp["pose"], inp["betas"])
    verts_cpu = verts_cpu.detach().requires_grad_(True)
    joints = orc.xyz_from_vertice(verts_cpu)
    root = joints[:, 9:10]
    view = (verts_cpu + (-root)) + inp["root_xyz"][:, None]
    fcl, prp = p3d.ndc_intrinsics(inp["Ks"])
    ndc = p3d.project_ndc(view, -fcl, prp)
    vn = p3d.vertex_normals(view, orc.faces)
    layer = hf.MyMANOLayer(True, DEV, shape_ncomp=10, pose_ncomp=48, tex_ncomp=None)
    topo = layer.topology(DEV)
    vg = verts_cpu.detach().to(DEV).requires_grad_(True)
    outs = ops.GeomFunction.apply(topo, vg, 9, inp["root_xyz"].to(DEV), (-fcl).to(DEV), prp.to(DEV), True)
    j_g, rel_g, view_g, ndc_g, vn_g = outs
    assert (j_g.cpu() - (joints - root)).abs().max() < 1e-6
    assert (rel_g.cpu() - (verts_cpu - root)).abs().max() < 1e-6
    assert (view_g.cpu() - view).abs().max() < 1e-6
    assert (ndc_g.cpu() - ndc).abs().max() < 1e-5
    assert (vn_g.cpu() - vn).abs().max() < 2e-4
    g = torch.Generator().manual_seed(0)
    ws = [torch.randn(t.shape, generator=g) for t in (joints, verts_cpu, view, ndc, vn)]
    ((joints - root) * ws[0]).sum().add((verts_cpu - root).mul(ws[1]).sum()).add((view * ws[2]).sum()).add(
        (ndc * ws[3]).sum()).add((vn * ws[4]).sum()).backward()
    (j_g * ws[0].to(DEV)).sum().add((rel_g * ws[1].to(DEV)).sum()).add((view_g * ws[2].to(DEV)).sum()).add(
        (ndc_g * ws[3].to(DEV)).sum()).add((vn_g * ws[4].to(DEV)).sum()).backward()
    assert rel_err(vg.grad, verts_cpu.grad) < 2e-3


# ------------------------------------------------------------------------------------------ rasterizer
def _face_verts(mano, B, seed, S=224):
    inp = P.synthetic_inputs(B, S=8, seed=seed)
    orc = ManoOracle(mano)
    verts, _ = orc(inp["pose"], inp["betas"])
    joints = orc.xyz_from_vertice(verts)
    view = (verts - joints[:, 9:10]) + inp["root_xyz"][:, None]
    fcl, prp = p3d.ndc_intrinsics(inp["Ks"])
    ndc = p3d.project_ndc(view, -fcl, prp)
    fv = ndc[:, orc.faces].reshape(-1, 3, 3).contiguous()
    Fm = orc.faces.shape[0]
    return fv, [i * Fm for i in range(B)], [Fm] * B


@pytest.mark.parametrize("H,W,K,blur,B", [(64, 64, 1, 0.0, 2), (224, 224, 4, 9.21e-4, 2), (48, 80, 2, 9.21e-4, 1),
                                          (96, 96, 3, 9.21e-4, 1), (128, 128, 8, 9.21e-4, 1), (100, 100, 16, 2e-3, 1),
                                          (672, 672, 1, 0.0, 1), (37, 53, 1, 0.0, 3)])
def test_rasterizer_bit_exact(hf, mano, H, W, K, blur, B):
    fv, first, nf = _face_verts(mano, B, seed=H + K)
    ref = raster_c.rasterize_naive(fv, first, nf, (H, W), blur, K, threads=8)
    out = hf.rasterize_meshes(fv.to(DEV), (H, W), blur, K, perspective_correct=True, clip_barycentric_coords=blur > 0,
                              mesh_to_face_first_idx=torch.tensor(first, device=DEV),
                              num_faces_per_mesh=torch.tensor(nf, device=DEV))
    names = ["pix_to_face", "zbuf", "bary_coords", "dists"]
    for name, g, r in zip(names, out, ref):
        bad = (g.cpu() != r)
        assert not bad.any(), f"{name}: {int(bad.sum())} mismatching entries of {bad.numel()}"
    assert out[0].dtype == torch.int64 and out[0].shape == (B, H, W, K)
    cov = out[0][..., 0] >= 0
    assert 0.02 < cov.float().mean() < 0.6          # the hand is in view
    # zbuf sorted ascending over valid slots
    z = out[1].cpu()
    valid = out[0].cpu() >= 0
    zz = torch.where(valid, z, torch.full_like(z, float("inf")))
    assert (zz[..., 1:] >= zz[..., :-1]).all()


def test_rasterizer_golden_and_edge_cases(hf):
    z = np.load(os.path.join(GOLD, "raster_oracle.npz"))
    fv = torch.tensor(z["face_verts"], device=DEV)
    first, nf = torch.zeros(1, dtype=torch.int64, device=DEV), torch.tensor([fv.shape[0]], device=DEV)
    out = hf.rasterize_meshes(fv, 32, 9.21e-4, 2, perspective_correct=True, clip_barycentric_coords=True,
                              mesh_to_face_first_idx=first, num_faces_per_mesh=nf)
    assert (out[0].cpu().numpy() == z["pix_to_face"]).all() and (out[1].cpu().numpy() == z["zbuf"]).all()
    assert (out[2].cpu().numpy() == z["bary"]).all() and (out[3].cpu().numpy() == z["dists"]).all()
    # everything off screen / behind the camera / degenerate -> all -1
    bad = torch.tensor([[[5.0, 5.0, 1.0], [6.0, 5.0, 1.0], [5.0, 6.0, 1.0]],
                        [[-0.9, -0.9, -1.0], [0.9, -0.9, 2.0], [0.0, 0.9, 2.0]],
                        [[0.1, 0.1, 1.0], [0.1, 0.1, 1.0], [0.2, 0.2, 1.0]]], device=DEV)
    o = hf.rasterize_meshes(bad, 16, 1e-3, 2, mesh_to_face_first_idx=first, num_faces_per_mesh=torch.tensor([3], device=DEV))
    assert (o[0] == -1).all() and (o[1] == -1).all() and (o[2] == -1).all() and (o[3] == -1).all()
    # ties: coincident faces resolve to the smaller index, nearer face first (matches the CPU oracle)
    tri = [[-0.9, -0.9, 2.0], [0.9, -0.9, 2.0], [0.0, 0.9, 2.0]]
    near = [[-0.9, -0.9, 1.5], [0.9, -0.9, 1.5], [0.0, 0.9, 1.5]]
    t3 = torch.tensor([tri, tri, near], device=DEV)
    o = hf.rasterize_meshes(t3, 8, 0.0, 3, perspective_correct=True, mesh_to_face_first_idx=first,
                            num_faces_per_mesh=torch.tensor([3], device=DEV))
    r = raster_c.rasterize_naive(t3.cpu(), [0], [3], 8, 0.0, 3)
    assert (o[0].cpu() == r[0]).all()
    # an empty mesh in the batch and an empty batch
    two_first = torch.tensor([0, 3], dtype=torch.int64, device=DEV)
    o = hf.rasterize_meshes(t3, 8, 0.0, 1, mesh_to_face_first_idx=two_first, num_faces_per_mesh=torch.tensor([3, 0], device=DEV))
    assert (o[0][1] == -1).all() and (o[0][0] >= 0).any()
    # argument errors surface as Python exceptions (PyTorch3D raises ValueError for bad settings)
    with pytest.raises(ValueError):
        hf.rasterize_meshes(t3, 8, 0.0, 17, mesh_to_face_first_idx=first, num_faces_per_mesh=torch.tensor([3], device=DEV))
    with pytest.raises(Exception):
        hf.rasterize_meshes(t3.cpu(), 8, 0.0, 1, mesh_to_face_first_idx=first, num_faces_per_mesh=torch.tensor([3], device=DEV))


def test_rasterizer_many_faces_in_one_tile(hf, mano):
    """A far-away hand: all 1538 faces land in one or two tiles (exercises batched list staging)."""
    fv, first, nf = _face_verts(mano, 1, seed=77)
    fv = fv.clone()
    fv[..., :2] = fv[..., :2] * 0.08
    ref = raster_c.rasterize_naive(fv, first, nf, 64, 9.21e-4, 8)
    out = hf.rasterize_meshes(fv.to(DEV), 64, 9.21e-4, 8, perspective_correct=True, clip_barycentric_coords=True,
                              mesh_to_face_first_idx=torch.tensor(first, device=DEV), num_faces_per_mesh=torch.tensor(nf, device=DEV))
    for g, r in zip(out, ref):
        assert (g.cpu() == r).all()


def test_rasterizer_backward(hf, mano):
    fv, first, nf = _face_verts(mano, 1, seed=5)
    H = W = 64
    for K, blur in [(1, 0.0), (4, 9.21e-4)]:
        f_cpu = fv.clone().requires_grad_(True)
        fr = p3d.rasterize_meshes(f_cpu, first, nf, (H, W), blur, K)
        g = torch.Generator().manual_seed(K)
        gz, gb, gd = torch.randn(fr.zbuf.shape, generator=g), torch.randn(fr.bary_coords.shape, generator=g), torch.randn(fr.dists.shape, generator=g) * 1e-2
        mk = (fr.pix_to_face >= 0).float()
        ((fr.zbuf * gz * mk).sum() + (fr.bary_coords * gb * mk[..., None]).sum() + (fr.dists * gd * mk).sum()).backward()
        f_gpu = fv.clone().to(DEV).requires_grad_(True)
        o = hf.rasterize_meshes(f_gpu, (H, W), blur, K, perspective_correct=True, clip_barycentric_coords=blur > 0,
                                mesh_to_face_first_idx=torch.tensor(first, device=DEV), num_faces_per_mesh=torch.tensor(nf, device=DEV))
        mkg = (o[0] >= 0).float()
        ((o[1] * gz.to(DEV) * mkg).sum() + (o[2] * gb.to(DEV) * mkg[..., None]).sum() + (o[3] * gd.to(DEV) * mkg).sum()).backward()
        assert rel_err(f_gpu.grad, f_cpu.grad) < 1e-3


# ------------------------------------------------------------------------------------------ renderer API + shading
def _render_modular(hf, inp, texture, S, aa, K, blur, soft, binarize, requires_grad=True):
    model = hf.HandRenderModel(True, DEV, image_size=S, aa_factor=aa, blur_radius=blur, faces_per_pixel=K, soft=soft,
                               binarize=binarize, texture_size=texture.shape[1]).to(DEV)
    with torch.no_grad():
        model.texture.copy_(texture.to(DEV))
    pose = inp["pose"].to(DEV).requires_grad_(requires_grad)
    betas = inp["betas"].to(DEV).requires_grad_(requires_grad)
    ldir = inp["light_dir"].to(DEV).requires_grad_(requires_grad)
    lcol = inp["light_color"].to(DEV).requires_grad_(requires_grad)
    out = model({"pose_params": pose, "shape_params": betas}, {"colors": lcol, "directions": ldir},
                Ks=inp["Ks"].to(DEV), root_xyz=inp["root_xyz"].to(DEV)[:, None], images=inp["imgs"].to(DEV))
    return model, out, dict(pose=pose, betas=betas, light_dir=ldir, light_color=lcol)


@pytest.mark.parametrize("S,aa,K,blur,soft,binarize,sil_scale", [(48, 1, 4, 9.21e-4, True, False, 1.0),
                                                                 (32, 3, 1, 0.0, False, True, 255.0)])
def test_full_path_modular_vs_oracle(hf, mano, S, aa, K, blur, soft, binarize, sil_scale):
    """hand layer -> joints -> root shift -> camera -> rasterize -> Phong/UV -> blend -> pool -> losses, fwd + bwd,
    through the reference-named objects.  Second case = the reference's own setting (SSAA 3x, K=1, hard, binarised)."""
    B = 2
    inp = P.synthetic_inputs(B, S=S, seed=31)
    tex = P.synthetic_texture(64)
    lam = dict(texture=1.0, mrgb=0.5, ssim_tex=0.7, sil=0.3, iou=0.2)
    # oracle
    oi = {k: v.clone() for k, v in inp.items()}
    for k in ("pose", "betas", "light_dir", "light_color"):
        oi[k].requires_grad_(True)
    tex_o = tex.clone().requires_grad_(True)
    ro = P.render_path(mano, oi, tex_o, image_size=S, aa=aa, K=K, blur_radius=blur, soft=soft, binarize=binarize)
    loss_o, terms_o = P.total_loss(ro, oi, lam, sil_scale)
    loss_o.backward()
    # product
    model, out, leaves = _render_modular(hf, inp, tex, S, aa, K, blur, soft, binarize)
    assert (out["joints"].cpu() - ro["joints"]).abs().max() < 1e-6
    assert (out["mano_verts"].cpu() - ro["verts_rel"]).abs().max() < 1e-6
    assert out["re_img"].shape == (B, 3, S, S) and out["re_sil"].shape == (B, 1, S, S)
    assert out["mano_faces"].shape == (B, 1538, 3) and out["maskRGBs"].shape == (B, 3, S, S)
    # a handful of edge pixels may pick a different face because the projected verts differ in the last bit
    diff_img = (out["re_img"].cpu() - ro["re_img"]).abs().amax(1)
    assert (diff_img > 1e-4).float().mean() < 2e-3
    assert (out["re_sil"].cpu() - ro["re_sil"]).abs().max() < (1e-3 if not binarize else 256)

    class A:
        lambda_texture, lambda_mrgb, lambda_ssim_tex, lambda_silhouette, lambda_iou = (lam["texture"], lam["mrgb"],
                                                                                        lam["ssim_tex"], lam["sil"], lam["iou"])
    lf = hf.LossFunction(sil_scale=sil_scale)
    ld = lf({"imgs": inp["imgs"].to(DEV), "segms_gt": inp["segms_gt"].to(DEV)}, out, ["sil", "iou"], "FreiHAND", A)
    for k in lam:
        assert abs(float(ld[k]) - float(terms_o[k])) < 2e-4 * max(1.0, abs(float(terms_o[k]))), k
    sum(ld.values()).backward()
    assert rel_err(leaves["pose"].grad, oi["pose"].grad) < 2e-2
    assert rel_err(leaves["betas"].grad, oi["betas"].grad) < 2e-2
    assert rel_err(model.texture.grad, tex_o.grad) < 2e-2
    assert rel_err(leaves["light_color"].grad, oi["light_color"].grad) < 2e-2
    assert rel_err(leaves["light_dir"].grad, oi["light_dir"].grad) < 2e-2


def test_shader_forward_backward_on_oracle_fragments(hf, mano):
    """Shading + blending in isolation: identical Fragments in, image and all gradients compared."""
    from hifihr_b200 import ops
    B, S, K = 2, 40, 4
    inp = P.synthetic_inputs(B, S=S, seed=8)
    tex = P.synthetic_texture(32)
    ro = P.render_path(mano, inp, tex, image_size=S, K=K, blur_radius=9.21e-4, soft=True)
    fr = ro["fragments"]
    faces = torch.tensor(np.asarray(mano["f"], np.int64))
    uvs, fuv = P.mano_uvs(mano)
    leaves_o = [t.detach().clone().requires_grad_(True) for t in
                (fr.zbuf, fr.bary_coords, fr.dists, ro["verts_view"], tex, inp["light_dir"], inp["light_color"])]
    z, b, d, vv, tx, ldir, lcol = leaves_o
    fro = p3d.Fragments(fr.pix_to_face, z, b, d)
    img_o = p3d.softmax_rgb_blend(p3d.phong_shading(fro, vv, faces, p3d.sample_textures_uv(fro, tx, fuv, uvs), ldir, lcol), fro)
    g = torch.Generator().manual_seed(2)
    gi = torch.randn(img_o.shape, generator=g)
    (img_o * gi).sum().backward()
    layer = hf.MyMANOLayer(True, DEV, shape_ncomp=10, pose_ncomp=48, tex_ncomp=None)
    leaves_g = [t.detach().clone().to(DEV).requires_grad_(True) for t in leaves_o]
    zg, bg, dg, vvg, txg, ldg, lcg = leaves_g
    meshes = hf.Meshes(vvg, layer.mesh_face, topology=layer.topology(DEV))
    meshes.textures = hf.TexturesUV(txg, fuv.to(DEV), uvs.to(DEV))
    shader = hf.SoftPhongShader(materials=hf.Materials(diffuse_color=((0.8, 0.8, 0.8),), specular_color=((0.2, 0.2, 0.2),), shininess=30))
    img_g = shader(hf.Fragments(fr.pix_to_face.to(DEV), zg, bg, dg), meshes,
                   lights=hf.DirectionalLights(diffuse_color=lcg, direction=ldg, device=DEV))
    assert (img_g.cpu() - img_o).abs().max() < 2e-5
    (img_g * gi.to(DEV)).sum().backward()
    for name, a, o in zip(("zbuf", "bary", "dists", "verts", "texture", "light_dir", "light_color"), leaves_g, leaves_o):
        assert rel_err(a.grad, o.grad) < 5e-3, name
    # silhouette shader
    sil = hf.SoftSilhouetteShader()(hf.Fragments(fr.pix_to_face.to(DEV), zg.detach(), bg.detach(), dg.detach()), meshes)
    sil_o = p3d.sigmoid_alpha_blend(torch.ones_like(fr.bary_coords), fr)
    assert (sil.cpu() - sil_o).abs().max() < 1e-5


# ------------------------------------------------------------------------------------------ losses
@pytest.mark.parametrize("sil_scale", [1.0, 255.0])
def test_losses_forward_backward(hf, sil_scale):
    from hifihr_b200 import ops
    g = torch.Generator().manual_seed(3)
    N, S = 3, 45
    re_img = torch.rand(N, 3, S, S, generator=g)
    re_sil = torch.rand(N, 1, S, S, generator=g) * sil_scale
    imgs = torch.rand(N, 3, S, S, generator=g)
    seg = (torch.rand(N, S, S, generator=g) > 0.5).long()
    lam = dict(texture=1.0, mrgb=2.0, ssim_tex=0.5, sil=0.25, iou=0.75)
    a, b = re_img.clone().requires_grad_(True), re_sil.clone().requires_grad_(True)
    terms = olosses.render_losses(a, b, imgs, seg, lam, sil_scale=sil_scale)
    sum(terms.values()).backward()
    ag, bg = re_img.to(DEV).requires_grad_(True), re_sil.to(DEV).requires_grad_(True)
    t = ops.RenderLossFunction.apply(ag, bg, imgs.to(DEV), seg.float().to(DEV), sil_scale, True)
    w = torch.tensor([lam["texture"], lam["mrgb"], lam["ssim_tex"], lam["sil"], lam["iou"]], device=DEV)
    for i, k in enumerate(("texture", "mrgb", "ssim_tex", "sil", "iou")):
        assert abs(float(t[i]) * lam[k] - float(terms[k])) < 1e-5 * max(1.0, abs(float(terms[k]))), k
    (t * w).sum().backward()
    assert rel_err(ag.grad, a.grad) < 1e-3
    assert rel_err(bg.grad, b.grad) < 1e-3


# ------------------------------------------------------------------------------------------ fused step
def test_fused_step_vs_oracle(hf, mano):
    B, S, K = 3, 64, 4
    inp = P.synthetic_inputs(B, S=S, seed=12)
    lam = dict(texture=1.0, mrgb=1.0, ssim_tex=1.0, sil=1.0, iou=0.5)
    step = hf.FusedHandStep(B, image_size=S, faces_per_pixel=K, soft=True, texture_size=64, lambdas=lam, device=DEV)
    tex = step.texture.detach().cpu().clone()
    oi = {k: v.clone() for k, v in inp.items()}
    for k in ("pose", "betas", "light_dir", "light_color"):
        oi[k].requires_grad_(True)
    tex_o = tex.clone().requires_grad_(True)
    ro = P.render_path(mano, oi, tex_o, image_size=S, K=K, blur_radius=step.blur, soft=True)
    loss_o, terms_o = P.total_loss(ro, oi, lam, 1.0)
    loss_o.backward()
    fcl, prp = p3d.ndc_intrinsics(inp["Ks"])
    d = lambda t: t.to(DEV).contiguous()  # noqa: E731
    args = (d(inp["pose"]), d(inp["betas"]), d(-fcl), d(prp), d(inp["root_xyz"]), d(inp["light_dir"]),
            d(inp["light_color"]), d(inp["imgs"]), d(inp["segms_gt"].float()))
    step.step(*args)
    torch.cuda.synchronize()
    # Fragments from the fused kernel are bit-exact against the C oracle run on the kernel's own face_verts
    fv = step.face_verts.cpu()
    Fm = 1538
    ref = raster_c.rasterize_naive(fv, [i * Fm for i in range(B)], [Fm] * B, S, step.blur, K, threads=8)
    assert (step.p2f.cpu() == ref[0]).all() and (step.zbuf.cpu() == ref[1]).all()
    assert (step.bary.cpu() == ref[2]).all() and (step.dists.cpu() == ref[3]).all()
    terms = step.loss_terms().cpu()
    for i, k in enumerate(("texture", "mrgb", "ssim_tex", "sil", "iou")):
        assert abs(float(terms[i]) * lam[k] - float(terms_o[k])) < 5e-4 * max(1.0, abs(float(terms_o[k]))), k
    assert rel_err(step.g_pose, oi["pose"].grad) < 2e-2
    assert rel_err(step.g_betas, oi["betas"].grad) < 2e-2
    assert rel_err(step.g_texture, tex_o.grad) < 2e-2
    assert rel_err(step.g_light_color, oi["light_color"].grad) < 2e-2
    assert rel_err(step.g_light_dir, oi["light_dir"].grad) < 2e-2
    # running the step again gives the same Fragments (determinism of the forward)
    p2f0, z0 = step.p2f.clone(), step.zbuf.clone()
    step.step(*args)
    torch.cuda.synchronize()
    assert (step.p2f == p2f0).all() and (step.zbuf == z0).all()


@pytest.mark.parametrize("S,aa,B", [(32, 3, 3), (40, 2, 2), (25, 2, 2)])   # 25*2 = 50: border-clipped tiles, unaligned rows
def test_fused_ssaa_step_vs_oracle_and_modular(hf, mano, S, aa, B):
    """SURVEY 8(f) row 1 - the reference's own render setting (672^2 -> 3x3 pool, K=1, hard Phong, binarised
    alpha; models_res_nimble.py:74-96, 208-220) as ONE fused pass, at a small size: Fragments bit-exact vs the
    C oracle at the rasterised resolution, pooled outputs vs the oracle pipeline AND vs the modular kernels
    (raster+shade at full resolution, then hfr_pool_forward), losses and gradients vs oracle autograd."""
    from hifihr_b200 import _lib as L
    from hifihr_b200 import ops
    K, Sr = 1, S * aa
    inp = P.synthetic_inputs(B, S=S, seed=77)
    lam = dict(texture=1.0, mrgb=0.5, ssim_tex=0.7, sil=0.3, iou=0.2)
    step = hf.FusedHandStep(B, image_size=S, faces_per_pixel=K, soft=False, texture_size=64, lambdas=lam, device=DEV,
                            aa_factor=aa, binarize=True, sil_scale=255.0, want_nchw=True)
    tex = step.texture.detach().cpu().clone()
    oi = {k: v.clone() for k, v in inp.items()}
    for k in ("pose", "betas", "light_dir", "light_color"):
        oi[k].requires_grad_(True)
    tex_o = tex.clone().requires_grad_(True)
    ro = P.render_path(mano, oi, tex_o, image_size=S, aa=aa, K=K, blur_radius=0.0, soft=False, binarize=True)
    loss_o, terms_o = P.total_loss(ro, oi, lam, 255.0)
    loss_o.backward()
    fcl, prp = p3d.ndc_intrinsics(inp["Ks"])
    d = lambda t: t.to(DEV).contiguous()  # noqa: E731
    args = (d(inp["pose"]), d(inp["betas"]), d(-fcl), d(prp), d(inp["root_xyz"]), d(inp["light_dir"]),
            d(inp["light_color"]), d(inp["imgs"]), d(inp["segms_gt"].float()))
    step.step(*args)
    torch.cuda.synchronize()
    # Fragments at the rasterised resolution: bit-exact against the C oracle on the kernel's own face_verts
    Fm = 1538
    ref = raster_c.rasterize_naive(step.face_verts.cpu(), [i * Fm for i in range(B)], [Fm] * B, Sr, 0.0, K, threads=8)
    assert step.p2f.shape == (B, Sr, Sr, K)
    assert (step.p2f.cpu() == ref[0]).all() and (step.zbuf.cpu() == ref[1]).all()
    assert (step.bary.cpu() == ref[2]).all() and (step.dists.cpu() == ref[3]).all()
    # modular kernels on the same geometry: full-resolution image, then the pooling kernel
    full = torch.empty(B, Sr, Sr, 4, device=DEV)
    fr2 = [torch.empty_like(t) for t in (step.p2f, step.zbuf, step.bary, step.dists)]
    r = ops.raster_args(step.face_verts, step.mesh_first, step.mesh_nf, Sr, Sr, K, 0.0, True, False, False, *fr2, step.ws)
    sa = ops.shade_fwd_args(step.params, fr2, step.topo.faces, step.verts_view, step.vnormals, step.faces_uvs,
                            step.verts_uvs, step.texture, args[5], args[6], full)
    L.call("hfr_raster_shade_forward", L.HfrRasterShadeArgs(r, sa))
    re_img, re_sil, mask = (torch.empty(B, 3, S, S, device=DEV), torch.empty(B, 1, S, S, device=DEV),
                            torch.empty(B, 3, S, S, device=DEV))
    L.call("hfr_pool_forward", L.HfrPoolArgs(B, S, S, aa, 1, L.ptr(full), L.ptr(args[7]), L.ptr(re_img), L.ptr(re_sil), L.ptr(mask)))
    torch.cuda.synchronize()
    assert (fr2[0] == step.p2f).all() and (fr2[1] == step.zbuf).all()
    assert (step.re_img - re_img).abs().max() < 1e-6 and (step.re_sil == re_sil).all() and (step.mask_rgbs == mask).all()
    assert (step.image[..., :3].permute(0, 3, 1, 2) == step.re_img).all() and (step.image[..., 3:4].permute(0, 3, 1, 2) == step.re_sil).all()
    assert set(step.re_sil.unique().tolist()) <= {0.0, 255.0}
    # against the oracle pipeline (a handful of edge pixels may flip: projected verts differ in the last bit)
    diff_img = (step.re_img.cpu() - ro["re_img"]).abs().amax(1)
    assert (diff_img > 1e-4).float().mean() < 2e-3
    assert ((step.re_sil.cpu() != ro["re_sil"]).float().mean()) < 2e-3
    terms = step.loss_terms().cpu()
    for i, k in enumerate(("texture", "mrgb", "ssim_tex", "sil", "iou")):
        assert abs(float(terms[i]) * lam[k] - float(terms_o[k])) < 5e-4 * max(1.0, abs(float(terms_o[k]))), k
    assert rel_err(step.g_pose, oi["pose"].grad) < 2e-2
    assert rel_err(step.g_betas, oi["betas"].grad) < 2e-2
    assert rel_err(step.g_texture, tex_o.grad) < 2e-2
    assert rel_err(step.g_light_color, oi["light_color"].grad) < 2e-2
    assert rel_err(step.g_light_dir, oi["light_dir"].grad) < 2e-2


def test_fused_ssaa_rejects_bad_sizes(hf):
    from hifihr_b200 import _lib as L
    a = L.HfrRasterShadePoolArgs()
    a.r.N, a.r.H, a.r.W, a.r.K = 0, 50, 50, 1
    a.s.p.N, a.s.p.H, a.s.p.W, a.s.p.K = 0, 50, 50, 1
    a.aa = 3
    with pytest.raises(ValueError):
        L.call("hfr_raster_shade_pool_forward", a)


def test_face_records_match_gather_path(hf):
    """The per-(sample, face) attribute records (hfr_face_attr_forward) are a pure re-layout: the fused step with
    and without them renders the same image bit for bit and produces the same gradients up to atomic order."""
    B, S, K = 2, 64, 4
    inp = P.synthetic_inputs(B, S=S, seed=21)
    fcl, prp = p3d.ndc_intrinsics(inp["Ks"])
    d = lambda t: t.to(DEV).contiguous()  # noqa: E731
    args = (d(inp["pose"]), d(inp["betas"]), d(-fcl), d(prp), d(inp["root_xyz"]), d(inp["light_dir"]),
            d(inp["light_color"]), d(inp["imgs"]), d(inp["segms_gt"].float()))
    steps = [hf.FusedHandStep(B, image_size=S, faces_per_pixel=K, soft=True, texture_size=64, device=DEV, face_records=fr)
             for fr in (True, False)]
    for st in steps:
        st.step(*args)
    torch.cuda.synchronize()
    a, b = steps
    rec = a.face_attr.cpu()
    faces = a.topo.faces.cpu().long()
    assert torch.equal(rec[..., 0:9].reshape(B, -1, 3, 3), a.verts_view.cpu()[:, faces])
    assert torch.equal(rec[..., 9:18].reshape(B, -1, 3, 3), a.vnormals.cpu()[:, faces])
    assert torch.equal(rec[..., 18:24].reshape(B, -1, 3, 2), a.verts_uvs.cpu()[a.faces_uvs.cpu().long()][None].expand(B, -1, -1, -1))
    assert torch.equal(rec[..., 24:27].contiguous().view(torch.int32).long(), faces[None].expand(B, -1, -1))
    assert torch.equal(a.image, b.image) and torch.equal(a.p2f, b.p2f)
    for k in ("g_pose", "g_betas", "g_texture", "g_light_dir", "g_light_color"):
        assert rel_err(getattr(a, k), getattr(b, k)) < 1e-4, k


def test_uint8_target_transport_matches_float(hf):
    """8-bit transport of the targets (imgs_u8 / seg_u8 of HfrLossArgs): the loss kernels apply ToTensor's x / 255
    (and byte != 0 -> 1.0) while loading, so a step fed with bytes equals a step fed with the floats the reference's
    loader would have produced - sums to 1e-6 relative (atomic order), the image gradient and parameter gradients too."""
    B, S, K = 2, 72, 4
    inp = P.synthetic_inputs(B, S=S, seed=8)
    fcl, prp = p3d.ndc_intrinsics(inp["Ks"])
    d = lambda t: t.to(DEV).contiguous()  # noqa: E731
    imgs_u8 = (inp["imgs"] * 255.0).round().to(torch.uint8)
    seg_u8 = inp["segms_gt"].to(torch.uint8)
    common = (d(inp["pose"]), d(inp["betas"]), d(-fcl), d(prp), d(inp["root_xyz"]), d(inp["light_dir"]), d(inp["light_color"]))
    a = hf.FusedHandStep(B, image_size=S, faces_per_pixel=K, soft=True, texture_size=64, device=DEV)
    b = hf.FusedHandStep(B, image_size=S, faces_per_pixel=K, soft=True, texture_size=64, device=DEV)
    a.step(*common, d(imgs_u8.float() / 255.0), d(seg_u8.float()))
    b.step(*common, d(imgs_u8), d(seg_u8))
    torch.cuda.synchronize()
    assert torch.equal(a.image, b.image)
    assert rel_err(b.sums[:5], a.sums[:5]) < 1e-6
    assert rel_err(b.g_image, a.g_image) < 1e-6
    for k in ("g_pose", "g_betas", "g_texture", "g_light_dir", "g_light_color"):
        assert rel_err(getattr(b, k), getattr(a, k)) < 1e-4, k
    # the conversion table is ToTensor's division, bit for bit
    ref = torch.arange(256, dtype=torch.float32) / 255.0
    assert torch.equal((torch.arange(256, dtype=torch.uint8).float() / 255.0), ref)


@pytest.mark.parametrize("K,soft", [(2, True), (3, True), (8, True), (3, False)])
def test_fused_fragments_bit_exact_other_k(hf, K, soft):
    """The fused rasterize+shade kernel at the K values the dispatch maps to other template sizes (K=2 -> 2, K=3 -> 4
    with a spare slot, K=8 -> 8 without the payload cache): all four Fragments tensors bit-exact vs the C oracle."""
    B, S = 2, 56
    inp = P.synthetic_inputs(B, S=S, seed=40 + K)
    fcl, prp = p3d.ndc_intrinsics(inp["Ks"])
    d = lambda t: t.to(DEV).contiguous()  # noqa: E731
    step = hf.FusedHandStep(B, image_size=S, faces_per_pixel=K, soft=soft, texture_size=32, device=DEV)
    step.step(d(inp["pose"]), d(inp["betas"]), d(-fcl), d(prp), d(inp["root_xyz"]), d(inp["light_dir"]),
              d(inp["light_color"]), d(inp["imgs"]), d(inp["segms_gt"].float()))
    torch.cuda.synchronize()
    Fm = 1538
    ref = raster_c.rasterize_naive(step.face_verts.cpu(), [i * Fm for i in range(B)], [Fm] * B, S, step.blur, K, threads=8)
    assert (step.p2f.cpu() == ref[0]).all() and (step.zbuf.cpu() == ref[1]).all()
    assert (step.bary.cpu() == ref[2]).all() and (step.dists.cpu() == ref[3]).all()
    assert (step.p2f >= 0).any() and torch.isfinite(step.image).all()
    for t in (step.g_pose, step.g_betas, step.g_texture):
        assert torch.isfinite(t).all() and t.abs().sum() > 0


def test_full_size_properties_c2(hf, mano):
    """BASELINE config 2 sizes (B=64, 224^2, K=4, soft): size-independent properties + a sampled bit-exact check."""
    B, S, K = 64, 224, 4
    inp = P.synthetic_inputs(B, S=S, seed=1234)
    step = hf.FusedHandStep(B, image_size=S, faces_per_pixel=K, soft=True, texture_size=512, device=DEV)
    fcl, prp = p3d.ndc_intrinsics(inp["Ks"])
    d = lambda t: t.to(DEV).contiguous()  # noqa: E731
    args = (d(inp["pose"]), d(inp["betas"]), d(-fcl), d(prp), d(inp["root_xyz"]), d(inp["light_dir"]),
            d(inp["light_color"]), d(inp["imgs"]), d(inp["segms_gt"].float()))
    step.step(*args)
    torch.cuda.synchronize()
    p2f, z = step.p2f, step.zbuf
    valid = p2f >= 0
    n_idx = torch.arange(B, device=DEV).view(B, 1, 1, 1)
    assert ((p2f[valid] // 1538) == n_idx.expand_as(p2f)[valid]).all()      # packed ids stay inside their mesh
    zz = torch.where(valid, z, torch.full_like(z, float("inf")))
    assert (zz[..., 1:] >= zz[..., :-1]).all()                              # sorted by depth
    assert (valid[..., 1:] <= valid[..., :-1]).all()                        # no holes before filled slots
    assert (step.bary[valid].sum(-1) - 1).abs().max() < 1e-4                # clipped barycentrics sum to 1
    img = step.image
    assert torch.isfinite(img).all() and (img[..., 3] >= 0).all() and (img[..., 3] <= 1).all()
    assert ((img[..., 3] > 0) == valid.any(-1)).all()
    for t in (step.g_pose, step.g_betas, step.g_texture, step.g_light_dir, step.g_light_color):
        assert torch.isfinite(t).all() and t.abs().sum() > 0
    for n in (0, 37, 63):                                                    # sampled bit-exact check
        fv = step.face_verts[n * 1538:(n + 1) * 1538].cpu()
        ref = raster_c.rasterize_naive(fv, [0], [1538], S, step.blur, K, threads=8)
        assert ((p2f[n].cpu() - n * 1538).where(p2f[n].cpu() >= 0, torch.tensor(-1)) == ref[0][0]).all()
        assert (z[n].cpu() == ref[1][0]).all() and (step.dists[n].cpu() == ref[3][0]).all()


@pytest.mark.parametrize("dat_name,key", [("FreiHAND", "freihand"), ("HO3D", "ho3d")])
def test_texture_metrics_match_reference_golden(hf, dat_name, key):
    """evaluation-time PSNR / SSIM / L1 / L2 (train_hrnet.py:149-161) from one forward kernel
    pass, against values computed by the unmodified reference pieces; tolerance 1e-5 relative (fp32 sums)."""
    z = np.load(os.path.join(GOLD, "texture_metrics_reference.npz"))
    t = lambda k: torch.tensor(z[k]).to(DEV)  # noqa: E731
    m = hf.texture_metrics({"imgs": t("imgs"), "segms_gt": t("segms_gt")}, {"re_img": t("re_img"), "re_sil": t("re_sil")},
                           dat_name)
    got = np.array([float(m[k]) for k in ("psnr", "ssim", "l1", "l2")])
    assert np.abs(got - z[key]).max() < 1e-5 * np.maximum(1.0, np.abs(z[key])).max(), (got, z[key])
    # and against the oracle on a second, ragged size (partial loss tiles)
    g = torch.Generator().manual_seed(3)
    N, H = 2, 50
    imgs, re = torch.rand(N, 3, H, H, generator=g), torch.rand(N, 3, H, H, generator=g)
    seg = (torch.rand(N, H, H, generator=g) > 0.6).float()
    sil = (torch.rand(N, 1, H, H, generator=g) > 0.5).float() * 255
    mo = olosses.texture_metrics(re, sil, imgs, seg, dat_name)
    mg = hf.texture_metrics({"imgs": imgs.to(DEV), "segms_gt": seg.to(DEV)}, {"re_img": re.to(DEV), "re_sil": sil.to(DEV)}, dat_name)
    for k in ("psnr", "ssim", "l1", "l2"):
        assert abs(float(mg[k]) - float(mo[k])) < 1e-5 * max(1.0, abs(float(mo[k]))), k


# ------------------------------------------------------------------------------------------ keypoints
@pytest.mark.parametrize("pre", ["l1.", "l2."])
def test_keypoint_losses_match_reference_golden(hf, pre):
    """hfr_keypoint_forward/backward against golden vectors of the unmodified reference functions:
    j2d 2e-4 px abs, terms 1e-5 rel, gradients 1e-4 of the tensor's max (L1 sign flips aside, none at these inputs)."""
    from hifihr_b200 import ops
    z = np.load(os.path.join(GOLD, "keypoint_reference.npz"))
    t = lambda k: torch.tensor(z[pre + k], device=DEV)  # noqa: E731
    j, v = t("joints").requires_grad_(True), t("verts").requires_grad_(True)
    faces = torch.tensor(z["faces"], device=DEV)
    terms, j2d = ops.KeypointLossFunction.apply(j, v, t("root"), t("K"), t("joints_gt"), t("j2d_gt"), t("verts_gt"), None,
                                                faces, 1 if pre == "l2." else 0)
    assert (j2d.detach().cpu() - torch.tensor(z[pre + "j2d"])).abs().max() < 2e-4
    for k in range(7):
        ref = z[pre + "terms"][k]
        assert abs(float(terms[k]) - ref) < 1e-5 * max(1.0, abs(ref)), (k, float(terms[k]), ref)
    w = torch.arange(1, 8, device=DEV, dtype=torch.float32)
    (terms[:7] * w).sum().backward()          # (term 7 is the Laplacian, off here: no neighbour lists given)
    assert rel_err(j.grad, torch.tensor(z[pre + "g_joints"])) < 1e-4
    assert rel_err(v.grad, torch.tensor(z[pre + "g_verts"])) < 1e-4


def test_keypoint_terms_through_loss_function_and_model(hf, mano):
    """LossFunction with the keypoint terms on HandRenderModel outputs (joints / mano_verts / mano_faces), gradients
    flowing through hfr_geom_backward and hfr_mano_backward to pose and shape; oracle = MANO + keypoint restatements."""
    from types import SimpleNamespace
    from oracle import keypoints as KP
    B = 5
    inp = P.synthetic_inputs(B, S=8, seed=31)
    g = torch.Generator().manual_seed(5)
    gt_j, gt_v = torch.randn(B, 21, 3, generator=g) * 0.04, torch.randn(B, 778, 3, generator=g) * 0.04
    gt_2d = torch.rand(B, 21, 2, generator=g) * 224
    K33 = inp["Ks"][:, :, :3].contiguous()
    args = SimpleNamespace(base_loss_fn="L1", lambda_j2d_gt=1e-3, lambda_j3d=10.0, lambda_vert_3d=10.0, lambda_bone_direc=0.5,
                           lambda_bone_direc_3d=0.7, lambda_edge_len=3.0, lambda_mscale=2.0)
    used = ["joint_2d", "joint_3d", "vert_3d", "bone_direc", "bone_direc_3d", "edge_length", "mscale"]
    # ours
    model = hf.HandRenderModel(ifRender=False, device=DEV)
    pose, betas = inp["pose"].to(DEV).requires_grad_(True), inp["betas"].to(DEV).requires_grad_(True)
    out = model({"pose_params": pose, "shape_params": betas})
    ex = dict(joints=gt_j.to(DEV), verts=gt_v.to(DEV), j2d_gt=gt_2d.to(DEV), Ks=K33.to(DEV), root_xyz=inp["root_xyz"].to(DEV))
    out["j2d"] = hf.trans_proj_j2d(out, ex["Ks"], root_xyz=ex["root_xyz"])
    ld = hf.LossFunction()(ex, out, used, "FreiHand", args)
    sum(ld.values()).backward()
    # oracle
    orc = ManoOracle(mano)
    po, bo = inp["pose"].clone().requires_grad_(True), inp["betas"].clone().requires_grad_(True)
    vo, _ = orc(po, bo)
    jo = orc.xyz_from_vertice(vo)
    root = jo[:, 9:10]
    jo, vo = jo - root, vo - root
    j2o = KP.project_joints(jo, K33, inp["root_xyz"])
    assert (out["j2d"].detach().cpu() - j2o.detach()).abs().max() < 2e-3       # pixels
    oo = KP.keypoint_losses(jo, j2o, vo, torch.tensor(np.asarray(mano["f"], np.int64)), gt_j, gt_2d, gt_v)
    lam = dict(joint_2d=1e-3, joint_3d=10.0, vert_3d=10.0, bone_direc=0.5, bone_direc_3d=0.7, edge_length=3.0, mscale=2.0)
    for k in used:
        assert abs(float(ld[k]) - lam[k] * float(oo[k])) < 2e-5 * max(1.0, abs(lam[k] * float(oo[k]))), k
    sum(lam[k] * oo[k] for k in used).backward()
    assert rel_err(pose.grad, po.grad) < 1e-3
    assert rel_err(betas.grad, bo.grad) < 1e-3


# ------------------------------------------------------------------------------------------ multi-GPU
def test_two_gpu_sharded_step_matches_single_process(hf):
    """NCCL, world size 2: batch slices + the two all-reduces reproduce the single-process step
    (tools/multi_gpu_check.py; loss terms 2e-6 abs, gradients 1e-4 of the tensor's max)."""
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (the gloo twin of this test runs on CPU: tests/test_dist_gloo.py)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + os.getpid() % 300), os.path.join(root, "tools", "multi_gpu_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
