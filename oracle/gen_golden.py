"""TEST INFRASTRUCTURE — generate tests/golden/*.npz.

Run in the build container (needs /root/reference for the MANO / SSIM goldens):
    python -m oracle.gen_golden
  mano_reference.npz : outputs + gradients of the UNMODIFIED reference ManoLayer (utils/my_mano.py)
                       on seeded inputs, plus the SURVEY.md Appendix C known answers.
  mano_modes_reference.npz : the same layer in its non-default modes (rot6d / robust rot6d root, rotmat joints,
                       axis-angle without PCA, root_palm, share_betas, th_trans, mean shape), outputs + gradients.
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
