"""CPU: pin the oracle (golden vectors from the unmodified reference, known answers, two independent
restatements agreeing) and check the C-ABI library exports what include/hifihr_b200.h declares."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle import losses as olosses
from oracle import p3d, raster_c, ref_mano
from oracle import pipeline as P
from oracle.mano import ManoOracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def test_mano_oracle_matches_reference_golden(mano):
    z = np.load(os.path.join(GOLD, "mano_reference.npz"))
    orc = ManoOracle(mano)
    pose = torch.tensor(z["pose"], requires_grad=True)
    beta = torch.tensor(z["beta"], requires_grad=True)
    v, j = orc(pose, beta)
    assert (v.detach() - torch.tensor(z["verts"])).abs().max() < 1e-7
    assert (j.detach() - torch.tensor(z["joints"])).abs().max() < 1e-7
    ((v * torch.tensor(z["g_verts"])).sum() + (j * torch.tensor(z["g_joints"])).sum()).backward()
    assert (pose.grad - torch.tensor(z["g_pose"])).abs().max() < 2e-5 * np.abs(z["g_pose"]).max()
    assert (beta.grad - torch.tensor(z["g_beta"])).abs().max() < 2e-5 * np.abs(z["g_beta"]).max()


def test_mano_oracle_modes_match_reference_golden(mano):
    """Non-default modes of ManoLayer.forward (my_mano.py:341-385, 459-461, 471-478): golden vectors from the
    unmodified reference (oracle/gen_golden.py::mano_mode_cases)."""
    from oracle.gen_golden import MODE_CASES
    z = np.load(os.path.join(GOLD, "mano_modes_reference.npz"))
    for name, (ckw, _, fkw) in MODE_CASES.items():
        kw = dict(ncomps=48, flat_hand_mean=False, center_idx=9)
        kw.update(ckw)
        orc = ManoOracle(mano, **kw)
        pose = torch.tensor(z[f"{name}.pose"], requires_grad=True)
        beta = torch.tensor(z[f"{name}.beta"], requires_grad=True)
        trans = torch.tensor(z[f"{name}.trans"], requires_grad=True) if fkw.get("trans") else None
        v, j = orc(pose, torch.zeros(1) if fkw.get("mean_shape") else beta, trans=trans,
                   root_palm=bool(fkw.get("root_palm")), share_betas=bool(fkw.get("share_betas")))
        assert (v.detach() - torch.tensor(z[f"{name}.verts"])).abs().max() < 2e-7, name
        assert (j.detach() - torch.tensor(z[f"{name}.joints"])).abs().max() < 2e-7, name
        ((v * torch.tensor(z[f"{name}.g_verts"])).sum() + (j * torch.tensor(z[f"{name}.g_joints"])).sum()).backward()
        assert (pose.grad - torch.tensor(z[f"{name}.g_pose"])).abs().max() < 5e-5 * np.abs(z[f"{name}.g_pose"]).max(), name
        if f"{name}.g_beta" in z.files:
            assert (beta.grad - torch.tensor(z[f"{name}.g_beta"])).abs().max() < 5e-5 * np.abs(z[f"{name}.g_beta"]).max(), name
        if trans is not None:
            assert (trans.grad - torch.tensor(z[f"{name}.g_trans"])).abs().max() < 1e-4, name


def test_keypoint_oracle_matches_reference_golden():
    """oracle/keypoints.py against golden vectors of the unmodified reference functions (trans_proj_j2d / proj_func,
    bone_direction_loss, edge_length_loss, L1 / MSE terms, mscale): terms 1e-6 rel, gradients 1e-5 of the max."""
    from oracle import keypoints as KP
    z = np.load(os.path.join(GOLD, "keypoint_reference.npz"))
    for pre in ("l1.", "l2."):
        t = lambda k: torch.tensor(z[pre + k])  # noqa: E731
        j, v = t("joints").requires_grad_(True), t("verts").requires_grad_(True)
        j2 = KP.project_joints(j, t("K"), t("root"))
        assert (j2.detach() - t("j2d")).abs().max() < 1e-4          # pixels
        o = KP.keypoint_losses(j, j2, v, torch.tensor(z["faces"]), t("joints_gt"), t("j2d_gt"), t("verts_gt"), l2=pre == "l2.")
        terms = [o[k] for k in KP.TERMS]
        for k, x in enumerate(terms):
            assert abs(float(x) - z[pre + "terms"][k]) < 1e-6 * max(1.0, abs(z[pre + "terms"][k])), (pre, KP.TERMS[k])
        sum((k + 1) * x for k, x in enumerate(terms)).backward()
        assert (j.grad - t("g_joints")).abs().max() < 1e-5 * t("g_joints").abs().max()
        assert (v.grad - t("g_verts")).abs().max() < 1e-5 * t("g_verts").abs().max()


@pytest.mark.skipif(not ref_mano.available(), reason="reference tree not present (GPU box)")
def test_keypoint_oracle_vs_live_reference():
    from oracle import keypoints as KP
    lu, fh, tu = ref_mano.reference_keypoint_modules()
    g = torch.Generator().manual_seed(11)
    B = 5
    j, jg = torch.randn(B, 21, 3, generator=g) * 0.05, torch.randn(B, 21, 3, generator=g) * 0.05
    root = torch.tensor([[0.01, -0.02, 0.6]]).repeat(B, 1).view(B, 1, 3)
    K = torch.tensor([[480.0, 0.5, 112.0], [0.0, 470.0, 110.0], [0.0, 0.0, 1.0]]).repeat(B, 1, 1)
    j2 = tu.trans_proj_j2d({"joints": j}, K, root_xyz=root)
    assert (j2 - KP.project_joints(j, K, root)).abs().max() < 1e-4
    assert (fh.proj_func(j + root, K) - KP.project_joints(j, K, root)).abs().max() < 1e-4
    con = torch.rand(B, 21, 1, generator=g)
    j2g = j2 + torch.randn(B, 21, 2, generator=g) * 3
    assert abs(float(lu.bone_direction_loss(j2, j2g, con)) - float(KP.bone_direction(j2, j2g, con))) < 1e-7
    assert abs(float(lu.bone_direction_loss(j, jg, con)) - float(KP.bone_direction(j, jg, con))) < 1e-6
    faces = torch.randint(0, 50, (30, 3), generator=g)
    v, vg = torch.randn(B, 50, 3, generator=g), torch.randn(B, 50, 3, generator=g)
    assert abs(float(lu.edge_length_loss(v, vg, faces[None].repeat(B, 1, 1))) - float(KP.edge_length(v, vg, faces))) < 1e-6


def test_generic_lbs_oracle_matches_mano_oracle(mano):
    """oracle/lbs.py (any skeleton; used for the NIMBLE-shaped layer) restricted to MANO's constants must be the
    pinned MANO oracle."""
    from oracle.lbs import LBSOracle
    from oracle.mano import JOINT_REORDER, PARENTS, TIP_VERTS_RIGHT
    inp = P.synthetic_inputs(6, S=8, seed=3)
    for dt, tol in ((torch.float64, 1e-12), (torch.float32, 1e-6)):
        o = ManoOracle(mano, dtype=dt)
        g = LBSOracle(mano["v_template"], mano["shapedirs"], mano["posedirs"], mano["J_regressor"], mano["weights"],
                      PARENTS, pca_comps=mano["hands_components"][:45], pose_mean=mano["hands_mean"],
                      tip_verts=TIP_VERTS_RIGHT, joint_order=JOINT_REORDER, center_joint=9, dtype=dt)
        v, j = o(inp["pose"].to(dt), inp["betas"].to(dt))
        v2, j2 = g(inp["pose"].to(dt), inp["betas"].to(dt))
        assert (v - v2).abs().max() < tol and (j - j2).abs().max() < tol
        tr = torch.randn(6, 3, dtype=dt)
        v, j = o(inp["pose"].to(dt), inp["betas"].to(dt), trans=tr)
        v2, j2 = g(inp["pose"].to(dt), inp["betas"].to(dt), trans=tr)
        assert (v - v2).abs().max() < tol and (j - j2).abs().max() < tol


def test_mano_known_answers_appendix_c(mano):
    """SURVEY.md Appendix C (values printed by the reference ManoLayer, torch CPU fp32)."""
    orc = ManoOracle(mano)
    v, j = orc(torch.zeros(1, 48), torch.zeros(1, 10))
    assert np.allclose(v[0, 0].numpy(), [0.04507172852754593, -0.01613779366016388, 0.01835748739540577], atol=1e-8)
    assert abs(v.double().sum().item() - (-16.718705629871693)) < 1e-4
    assert np.allclose(j[0, 0].numpy(), [0.09466041624546051, 0.0014789639972150326, 0.0033575384877622128], atol=1e-8)
    assert j[0, 9].abs().max() == 0
    g = torch.Generator().manual_seed(1234)
    pose = torch.randn(4, 48, generator=g) * 0.5
    beta = torch.randn(4, 10, generator=g) * 0.5
    v, j = orc(pose, beta)
    assert abs(v.double().sum().item() - (-70.60127041395754)) < 1e-4
    assert np.allclose(j[3, 20].numpy(), [0.06128855049610138, -0.04273726046085358, 0.000701904296875], atol=1e-7)


@pytest.mark.skipif(not ref_mano.available(), reason="reference tree not present (GPU box)")
def test_mano_oracle_vs_live_reference(mano):
    ref = ref_mano.reference_mano_layer()
    orc = ManoOracle(mano)
    g = torch.Generator().manual_seed(5)
    pose = torch.cat([torch.randn(8, 3, generator=g) * 1.5, torch.randn(8, 45, generator=g) * 0.5], 1)
    beta = torch.randn(8, 10, generator=g) * 0.5
    v, j = ref(pose, beta)
    vo, jo = orc(pose, beta)
    assert (v - vo).abs().max() < 1e-7 and (j - jo).abs().max() < 1e-7


def test_ssim_oracle_matches_reference_golden():
    z = np.load(os.path.join(GOLD, "ssim_reference.npz"))
    a, b = torch.tensor(z["a"]), torch.tensor(z["b"])
    assert abs(float(olosses.ssim(a, b)) - float(z["ssim"])) < 1e-6
    assert abs(float(olosses.ssim(a, a)) - float(z["ssim_same"])) < 1e-6


def test_texture_metrics_oracle_matches_reference_golden():
    """PSNR / SSIM / L1 / L2 of train_hrnet.py:149-161, both mask branches; the golden values come
    from the unmodified utils/pytorch_ssim and the reference's MSE / L1 bodies (oracle/gen_golden_metrics.py)."""
    z = np.load(os.path.join(GOLD, "texture_metrics_reference.npz"))
    t = lambda k: torch.tensor(z[k])  # noqa: E731
    for name, key in (("FreiHAND", "freihand"), ("HO3D", "ho3d")):
        m = olosses.texture_metrics(t("re_img"), t("re_sil"), t("imgs"), t("segms_gt"), name)
        got = np.array([float(m[k]) for k in ("psnr", "ssim", "l1", "l2")])
        assert np.abs(got - z[key]).max() < 1e-5 * np.maximum(1.0, np.abs(z[key])).max(), (name, got, z[key])


def test_raster_two_restatements_agree_and_golden():
    z = np.load(os.path.join(GOLD, "raster_oracle.npz"))
    fv = torch.tensor(z["face_verts"])
    Fm = fv.shape[0]
    c = raster_c.rasterize_naive(fv, [0], [Fm], 32, 9.21e-4, 2)
    t = p3d.rasterize_meshes(fv, [0], [Fm], 32, 9.21e-4, 2)
    assert (c[0] == t.pix_to_face).all() and (c[1] == t.zbuf).all() and (c[2] == t.bary_coords).all() and (c[3] == t.dists).all()
    assert (c[0].numpy() == z["pix_to_face"]).all() and (c[1].numpy() == z["zbuf"]).all()
    assert (c[2].numpy() == z["bary"]).all() and (c[3].numpy() == z["dists"]).all()


def test_raster_known_answers_single_triangle():
    """Analytic cases: one front-facing triangle, z=2 plane, 8x8 image."""
    fv = torch.tensor([[[-0.9, -0.9, 2.0], [0.9, -0.9, 2.0], [0.0, 0.9, 2.0]]])
    p2f, zb, ba, ds = raster_c.rasterize_naive(fv, [0], [1], 8, 0.0, 1, perspective_correct=True)
    cov = p2f[0, ..., 0] >= 0
    assert cov.any() and (zb[0, ..., 0][cov] - 2.0).abs().max() < 1e-6          # constant depth
    assert (ba[0, ..., 0, :][cov].sum(-1) - 1).abs().max() < 1e-6               # barycentrics sum to 1
    assert (ba[0, ..., 0, :][cov] > 0).all() and (ds[0, ..., 0][cov] < 0).all()  # strictly inside, negative dist
    # NDC +x is LEFT and +y is UP: the apex (y=+0.9) must be in the top rows, the base in the bottom rows
    rows = cov.any(1).nonzero().flatten()
    assert cov[rows[0]].sum() < cov[rows[-1]].sum()
    # a pixel centre exactly on an edge belongs to no face (strict >): triangle with an edge through x=0.125
    fv2 = torch.tensor([[[0.125, -1.0, 1.0], [0.125, 1.0, 1.0], [-1.0, 0.0, 1.0]]])
    p2 = raster_c.rasterize_naive(fv2, [0], [1], 8, 0.0, 1)[0][0, ..., 0]
    xf = p3d.pixel_centers(8, 8)[1]
    col = int((xf == 0.125).nonzero()[0])
    assert (p2[:, col] == -1).all()
    # depth order + tie: two coincident faces -> smaller index wins; nearer face wins
    tri = [[-0.9, -0.9, 2.0], [0.9, -0.9, 2.0], [0.0, 0.9, 2.0]]
    near = [[-0.9, -0.9, 1.5], [0.9, -0.9, 1.5], [0.0, 0.9, 1.5]]
    p3 = raster_c.rasterize_naive(torch.tensor([tri, tri, near]), [0], [3], 8, 0.0, 3)[0][0]
    c3 = p3[..., 0] >= 0
    assert (p3[..., 0][c3] == 2).all() and (p3[..., 1][c3] == 0).all() and (p3[..., 2][c3] == 1).all()
    # behind the camera / degenerate faces never rasterize
    bad = torch.tensor([[[-0.9, -0.9, -1.0], [0.9, -0.9, 2.0], [0.0, 0.9, 2.0]],
                        [[0.1, 0.1, 1.0], [0.1, 0.1, 1.0], [0.2, 0.2, 1.0]]])
    assert (raster_c.rasterize_naive(bad, [0], [2], 8, 1e-3, 2)[0] == -1).all()


@pytest.mark.parametrize("pc", [True, False])
def test_raster_restatements_agree_on_pixel_centred_vertices(pc):
    """The scene of tests/test_gpu_round2b.py::test_hard_k1_pixel_centres_on_edges_and_vertices on the CPU side: triangles
    whose vertices ARE pixel centres, both windings (many edge functions exactly zero).  The scalar C restatement and the
    vectorised torch restatement of PyTorch3D's rasterizer agree bit for bit there, and a pixel centre on a shared diagonal
    belongs to neither triangle (strict `> 0`)."""
    S = 16
    c = lambda i: (2.0 * i + 1.0) / S - 1.0  # noqa: E731
    g = torch.Generator().manual_seed(5)
    tris = []
    for _ in range(60):
        ij = torch.randint(0, S, (3, 2), generator=g)
        z = 1.0 + torch.rand(3, generator=g) * 2.0
        t = [[c(int(ij[k, 0])), c(int(ij[k, 1])), float(z[k])] for k in range(3)]
        tris.append(t)
        tris.append([t[0], t[2], t[1]])
    fv = torch.tensor(tris, dtype=torch.float32)
    for K in (1, 3):
        a = raster_c.rasterize_naive(fv, [0], [fv.shape[0]], S, 0.0, K, perspective_correct=pc)
        b = p3d.rasterize_meshes(fv, [0], [fv.shape[0]], S, 0.0, K, perspective_correct=pc)
        assert (a[0] == b.pix_to_face).all() and (a[1] == b.zbuf).all() and (a[2] == b.bary_coords).all() and (a[3] == b.dists).all()
        assert (a[0] >= 0).any()
    two = torch.tensor([[[c(2), c(2), 1.5], [c(12), c(2), 1.5], [c(2), c(12), 1.5]],
                        [[c(12), c(12), 1.5], [c(2), c(12), 1.5], [c(12), c(2), 1.5]]])
    p2f = raster_c.rasterize_naive(two, [0], [2], S, 0.0, 1, perspective_correct=pc)[0][0, ..., 0]
    xs = p3d.pixel_centers(S, S)[1]
    ys = p3d.pixel_centers(S, S)[0]
    on_diag = 0
    for yi in range(S):
        for xi in range(S):
            x, y = float(xs[xi]) if xs.dim() == 1 else float(xs[yi, xi]), float(ys[yi]) if ys.dim() == 1 else float(ys[yi, xi])
            if abs((x - c(12)) + (y - c(2))) < 1e-9 and c(2) < x < c(12):      # on the diagonal x + y = c(2) + c(12), strictly between its ends
                on_diag += 1
                assert int(p2f[yi, xi]) == -1
    assert on_diag >= 5 and (p2f == 0).any() and (p2f == 1).any()


def test_soft_blend_known_answer():
    """One fragment exactly on an edge (dist 0): prob 0.5 -> alpha 0.5; empty pixel -> background, alpha 0."""
    p2f = torch.tensor([[[[0, -1]], [[-1, -1]]]])
    fr = p3d.Fragments(p2f, torch.tensor([[[[0.5, -1.0]], [[-1.0, -1.0]]]]), torch.zeros(1, 2, 1, 2, 3),
                       torch.tensor([[[[0.0, -1.0]], [[-1.0, -1.0]]]]))
    col = torch.zeros(1, 2, 1, 2, 3)
    img = p3d.softmax_rgb_blend(col, fr)
    assert abs(float(img[0, 0, 0, 3]) - 0.5) < 1e-6 and float(img[0, 1, 0, 3]) == 0
    assert (img[0, 1, 0, :3] - 1).abs().max() < 1e-6 and img[0, 0, 0, :3].abs().max() < 1e-3
    sil = p3d.sigmoid_alpha_blend(torch.ones(1, 2, 1, 2, 3), fr)
    assert abs(float(sil[0, 0, 0, 3]) - 0.5) < 1e-6


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "hifihr_b200.h")).read()
    declared = set(re.findall(r"\b(hfr_[a-z_0-9]+)\s*\(", hdr))
    from hifihr_b200 import _lib
    from hifihr_b200.build import build
    lib = ctypes.CDLL(build())
    assert declared == set(_lib.ENTRY_POINTS), declared ^ set(_lib.ENTRY_POINTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.hfr_abi_version() == _lib.ABI_VERSION == 5


def test_ctypes_struct_sizes_match_header():
    """sizeof() of every ctypes mirror against a C translation unit compiled from the header."""
    import subprocess
    import tempfile
    from hifihr_b200 import _lib
    names = ["HfrHandModel", "HfrManoFwdArgs", "HfrManoBwdArgs", "HfrTopology", "HfrGeomFwdArgs", "HfrGeomBwdArgs", "HfrFaceVertsArgs",
             "HfrRasterArgs", "HfrRasterBwdArgs", "HfrShadeParams", "HfrShadeFwdArgs", "HfrShadeBwdArgs",
             "HfrRasterShadeArgs", "HfrRasterShadePoolArgs", "HfrFaceAttrArgs", "HfrPoolArgs", "HfrPoolBwdArgs", "HfrLossArgs", "HfrLossBwdArgs", "HfrKeypointArgs",
             "HfrKeypointBwdArgs"]
    src = '#include <stdio.h>\n#include "hifihr_b200.h"\nint main(){' + "".join(
        f'printf("%zu\\n", sizeof({n}));' for n in names) + "return 0;}"
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(td, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    for n, sz in zip(names, sizes):
        assert ctypes.sizeof(getattr(_lib, n)) == sz, (n, ctypes.sizeof(getattr(_lib, n)), sz)


def test_product_has_no_cpu_path():
    import hifihr_b200
    layer = hifihr_b200.ManoLayer(center_idx=9, flat_hand_mean=False, ncomps=48)
    with pytest.raises(RuntimeError):
        layer(torch.zeros(1, 48), torch.zeros(1, 10))
    with pytest.raises(Exception):
        hifihr_b200.rasterize_meshes(torch.zeros(1, 3, 3), 8, mesh_to_face_first_idx=torch.zeros(1, dtype=torch.int64),
                                     num_faces_per_mesh=torch.ones(1, dtype=torch.int64))
    # the product never imports the oracle
    import sys
    pkg = os.path.join(ROOT, "hifihr_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle|import_module\(['\"]oracle|__import__\(['\"]oracle", src, re.M), f


def test_product_and_oracle_synthetic_inputs_agree():
    from hifihr_b200.synthetic import synthetic_inputs
    a, b = synthetic_inputs(3, S=16, seed=9), P.synthetic_inputs(3, S=16, seed=9)
    assert a.keys() == b.keys() and all(torch.equal(a[k], b[k]) for k in a)
