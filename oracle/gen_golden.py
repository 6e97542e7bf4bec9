"""TEST INFRASTRUCTURE — generate tests/golden/*.npz.

Run in the build container (needs /root/reference for the MANO / SSIM goldens):
    python -m oracle.gen_golden
  mano_reference.npz : outputs + gradients of the UNMODIFIED reference ManoLayer (utils/my_mano.py)
                       on seeded inputs, plus the SURVEY.md Appendix C known answers.
  mano_modes_reference.npz : the same layer in its non-default modes (rot6d / robust rot6d root, rotmat joints,
                       axis-angle without PCA, root_palm, share_betas, th_trans, mean shape), outputs + gradients.
  keypoint_reference.npz : proj_func / trans_proj_j2d, bone_direction_loss, edge_length_loss, L1 / MSE keypoint and
                       vertex terms, mscale — unmodified reference functions, terms + gradients.
  ssim_reference.npz : utils/pytorch_ssim.ssim on seeded images (unmodified module).
  raster_oracle.npz  : Fragments of the scalar C oracle on one small seeded MANO view (regression pin
                       of the restatement; PyTorch3D itself is unavailable -> parity unpinned).
"""
import os

import numpy as np
import torch

from hifihr_b200.mano_assets import load_mano
from oracle import pipeline as P
from oracle import raster_c, ref_mano

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


MODE_CASES = {
    # name: (constructor kwargs, pose shape after B, forward kwargs)
    "rot6d": (dict(root_rot_mode="rot6d", ncomps=45), (51,), {}),
    "rot6d_robust": (dict(root_rot_mode="rot6d", ncomps=30, robust_rot=True), (36,), {}),
    "rotmat": (dict(use_pca=False, joint_rot_mode="rotmat"), (16, 3, 3), {}),
    "nopca": (dict(use_pca=False), (48,), {}),
    "root_palm": (dict(center_idx=0), (48,), dict(root_palm=True)),
    "share_betas": (dict(), (48,), dict(share_betas=True)),
    "trans": (dict(), (48,), dict(trans=True)),
    "mean_shape": (dict(center_idx=None, flat_hand_mean=True, ncomps=6), (9,), dict(mean_shape=True)),
}


def mano_mode_cases(B=3):
    """Run the unmodified reference ManoLayer in every non-default mode; rotproj's hard `.cuda()`
    (utils/manopth/rotproj.py:19) is neutralised for the duration of the call (CPU container)."""
    from unittest import mock
    out = {}
    g = torch.Generator().manual_seed(4321)
    for name, (ckw, pshape, fkw) in MODE_CASES.items():
        layer = ref_mano.reference_mano_layer(**ckw)
        if name == "rotmat":
            pose = torch.eye(3) + 0.3 * torch.randn(B, *pshape, generator=g)
        else:
            pose = torch.randn(B, *pshape, generator=g) * 0.5
        pose.requires_grad_(True)
        beta = (torch.randn(B, 10, generator=g) * 0.5).requires_grad_(True)
        kw = {}
        if fkw.get("root_palm"):
            kw["root_palm"] = torch.Tensor([1])
        if fkw.get("share_betas"):
            kw["share_betas"] = torch.Tensor([1])
        trans = None
        if fkw.get("trans"):
            trans = torch.randn(B, 3, generator=g).requires_grad_(True)
            kw["th_trans"] = trans
        with mock.patch.object(torch.Tensor, "cuda", lambda self, *a, **k: self):
            v, j = layer(pose, torch.zeros(1) if fkw.get("mean_shape") else beta, **kw)
        gv, gj = torch.randn(v.shape, generator=g), torch.randn(j.shape, generator=g)
        ((v * gv).sum() + (j * gj).sum()).backward()
        out.update({f"{name}.pose": pose.detach().numpy(), f"{name}.beta": beta.detach().numpy(),
                    f"{name}.verts": v.detach().numpy(), f"{name}.joints": j.detach().numpy(),
                    f"{name}.g_verts": gv.numpy(), f"{name}.g_joints": gj.numpy(), f"{name}.g_pose": pose.grad.numpy()})
        if beta.grad is not None:
            out[f"{name}.g_beta"] = beta.grad.numpy()
        if trans is not None:
            out[f"{name}.trans"] = trans.detach().numpy()
            out[f"{name}.g_trans"] = trans.grad.numpy()
    return out


def keypoint_cases(B=3):
    """The unmodified reference functions (proj via trans_proj_j2d, bone_direction_loss, edge_length_loss, nn.L1Loss /
    mse_loss, the mscale lines of losses.py:293-299) on seeded MANO-sized inputs; terms + autograd gradients of
    sum_k (k+1) * term_k wrt joints and verts, for base_loss_fn L1 and L2."""
    import torch.nn as nn
    import torch.nn.functional as torch_f
    lu, fh, tu = ref_mano.reference_keypoint_modules()
    mano = load_mano()
    faces = torch.tensor(np.asarray(mano["f"], np.int64))
    g = torch.Generator().manual_seed(777)
    out = {"faces": faces.numpy().astype(np.int32)}
    vt = torch.tensor(np.asarray(mano["v_template"], np.float32))
    for l2 in (0, 1):
        base = torch_f.mse_loss if l2 else nn.L1Loss()
        joints = (torch.randn(B, 21, 3, generator=g) * 0.04).requires_grad_(True)
        joints_gt = torch.randn(B, 21, 3, generator=g) * 0.04
        verts = (vt[None] + torch.randn(B, 778, 3, generator=g) * 0.003).requires_grad_(True)
        verts_gt = vt[None] + torch.randn(B, 778, 3, generator=g) * 0.003
        root = torch.stack([torch.rand(B, generator=g) * 0.06 - 0.03, torch.rand(B, generator=g) * 0.06 - 0.03,
                            torch.rand(B, generator=g) * 0.2 + 0.55], 1).view(B, 1, 3)
        K = torch.zeros(B, 3, 3)
        K[:, 0, 0] = 440 + 80 * torch.rand(B, generator=g)
        K[:, 1, 1] = 440 + 80 * torch.rand(B, generator=g)
        K[:, 0, 2] = 112 + 8 * torch.rand(B, generator=g)
        K[:, 1, 2] = 112 - 8 * torch.rand(B, generator=g)
        K[:, 2, 2] = 1
        j2d = tu.trans_proj_j2d({"joints": joints}, K, root_xyz=root)
        j2d_gt = j2d.detach() + torch.randn(B, 21, 2, generator=g) * 4
        con = torch.ones(B, 21, 1)
        terms = [base(j2d_gt, j2d), base(joints, joints_gt), base(verts, verts_gt),
                 lu.bone_direction_loss(j2d, j2d_gt, con), lu.bone_direction_loss(joints, joints_gt, con),
                 lu.edge_length_loss(verts, verts_gt, faces[None].repeat(B, 1, 1)),
                 nn.L1Loss()(torch.sqrt(torch.sum((joints[:, 9, :] - joints[:, 10, :]) ** 2, 1)),
                             torch.ones(B) * 0.0282)]
        total = sum((k + 1) * t for k, t in enumerate(terms))
        total.backward()
        pre = f"l{l2 + 1}."
        out.update({pre + "joints": joints.detach().numpy(), pre + "joints_gt": joints_gt.numpy(),
                    pre + "verts": verts.detach().numpy(), pre + "verts_gt": verts_gt.numpy(), pre + "root": root.numpy(),
                    pre + "K": K.numpy(), pre + "j2d": j2d.detach().numpy(), pre + "j2d_gt": j2d_gt.numpy(),
                    pre + "terms": np.array([float(t.detach()) for t in terms], np.float64),
                    pre + "g_joints": joints.grad.numpy(), pre + "g_verts": verts.grad.numpy()})
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    assert ref_mano.available(), "reference tree not found"
    layer = ref_mano.reference_mano_layer()
    g = torch.Generator().manual_seed(1234)
    pose = (torch.randn(4, 48, generator=g) * 0.5).requires_grad_(True)
    beta = (torch.randn(4, 10, generator=g) * 0.5).requires_grad_(True)
    v, j = layer(pose, beta)
    gv = torch.randn(v.shape, generator=g)
    gj = torch.randn(j.shape, generator=g)
    ((v * gv).sum() + (j * gj).sum()).backward()
    v0, j0 = layer(torch.zeros(1, 48), torch.zeros(1, 10))
    np.savez_compressed(os.path.join(OUT, "mano_reference.npz"), pose=pose.detach().numpy(), beta=beta.detach().numpy(),
                        verts=v.detach().numpy(), joints=j.detach().numpy(), g_verts=gv.numpy(), g_joints=gj.numpy(),
                        g_pose=pose.grad.numpy(), g_beta=beta.grad.numpy(), verts_zero=v0.detach().numpy(),
                        joints_zero=j0.detach().numpy())
    np.savez_compressed(os.path.join(OUT, "mano_modes_reference.npz"), **mano_mode_cases())
    np.savez_compressed(os.path.join(OUT, "keypoint_reference.npz"), **keypoint_cases())
    ps = ref_mano.reference_ssim()
    a, b = torch.rand(2, 3, 40, 40, generator=g), torch.rand(2, 3, 40, 40, generator=g)
    np.savez_compressed(os.path.join(OUT, "ssim_reference.npz"), a=a.numpy(), b=b.numpy(),
                        ssim=float(ps.ssim(a, b)), ssim_same=float(ps.ssim(a, a)))
    mano = load_mano()
    inp = P.synthetic_inputs(1, S=32, seed=99)
    out = P.render_path(mano, inp, P.synthetic_texture(16), image_size=32, K=1)
    faces = torch.tensor(np.asarray(mano["f"], np.int64))
    fv = out["verts_ndc"][:, faces].reshape(-1, 3, 3).contiguous()
    p2f, zb, ba, ds = raster_c.rasterize_naive(fv, [0], [fv.shape[0]], 32, 9.21e-4, 2)
    np.savez_compressed(os.path.join(OUT, "raster_oracle.npz"), face_verts=fv.numpy(), pix_to_face=p2f.numpy().astype(np.int32),
                        zbuf=zb.numpy(), bary=ba.numpy(), dists=ds.numpy())
    print("wrote", os.listdir(OUT))


if __name__ == "__main__":
    main()
