"""TEST INFRASTRUCTURE — CPU restatement (torch, autograd) of the keypoint / mesh-regulariser terms next to the
render path:

  * proj_func                       utils/fh_utils.py:30-39
  * trans_proj_j2d (root_xyz path)  utils/traineval_util.py:338-354 (called at train_hrnet.py:83)
  * joint_2d / joint_3d / vert_3d   losses.py:244-265 (base_loss_fn = nn.L1Loss or mse_loss, :239-242)
  * bone_direc / bone_direc_3d      losses.py:268-282 + utils/losses_util.py:217-282
  * edge_length                     losses.py:285-289 + utils/losses_util.py:284-301
  * mscale                          losses.py:293-299

Pinned against the unmodified reference functions (oracle/ref_mano.py::reference_keypoint_modules) in
tests/test_oracle_pins.py and through tests/golden/keypoint_reference.npz.
"""
from __future__ import annotations

import torch

# bone i joins child i+1 to its parent: a chain of four per finger hanging off the wrist (losses_util.py:226-245)
BONE_CHILD = list(range(1, 21))
BONE_PARENT = [0 if (c - 1) % 4 == 0 else c - 1 for c in BONE_CHILD]
TERMS = ("joint_2d", "joint_3d", "vert_3d", "bone_direc", "bone_direc_3d", "edge_length", "mscale")


def project_joints(joints, Ks, root_xyz=None):
    """j2d = (K X).xy / (K X).z with X = joints (+ root_xyz)."""
    X = joints if root_xyz is None else joints + root_xyz.reshape(-1, 1, 3)
    uvw = torch.einsum("brc,bkc->bkr", Ks[:, :3, :3], X)
    return uvw[..., :2] / uvw[..., 2:3]


def base_loss(a, b, l2=False):
    return ((a - b) ** 2).mean() if l2 else (a - b).abs().mean()


def bone_direction(j, j_gt, con=None):
    """mean over (batch, 20 bones) of conf * |unit(bone) - unit(bone_gt)|^2, unit(v) = v / (|v| + 1e-4)."""
    v = j[:, BONE_CHILD] - j[:, BONE_PARENT]
    t = j_gt[:, BONE_CHILD] - j_gt[:, BONE_PARENT]
    v = v / (v.pow(2).sum(-1, keepdim=True).sqrt() + 1e-4)
    t = t / (t.pow(2).sum(-1, keepdim=True).sqrt() + 1e-4)
    d = (v - t).pow(2).sum(-1)
    if con is not None:
        c = con.reshape(con.shape[0], -1)
        d = d * c[:, BONE_PARENT] * c[:, BONE_CHILD]
    return d.mean()


def edge_length(pred, gt, faces):
    """mean over (batch, 3F) of | |edge| - |edge_gt| |; edges (0,1), (0,2), (1,2) of every face."""
    f = faces.long()
    out = []
    for a, b in ((0, 1), (0, 2), (1, 2)):
        d = (pred[:, f[:, a]] - pred[:, f[:, b]]).pow(2).sum(-1).sqrt()
        g = (gt[:, f[:, a]] - gt[:, f[:, b]]).pow(2).sum(-1).sqrt()
        out.append((d - g).abs())
    return torch.cat(out, 1).mean()


def mscale(joints):
    return ((joints[:, 9] - joints[:, 10]).pow(2).sum(1).sqrt() - 0.0282).abs().mean()


def keypoint_losses(joints, j2d, verts, faces, joints_gt=None, j2d_gt=None, verts_gt=None, l2=False, con=None):
    """Unweighted terms (dict); a term is present when its ground truth is."""
    out = {}
    if j2d_gt is not None and j2d is not None:
        out["joint_2d"] = base_loss(j2d_gt, j2d, l2)
        out["bone_direc"] = bone_direction(j2d, j2d_gt, con)
    if joints_gt is not None:
        out["joint_3d"] = base_loss(joints, joints_gt, l2)
        out["bone_direc_3d"] = bone_direction(joints, joints_gt, con)
    if verts_gt is not None and verts is not None:
        out["vert_3d"] = base_loss(verts, verts_gt, l2)
        out["edge_length"] = edge_length(verts, verts_gt, faces)
    out["mscale"] = mscale(joints)
    return out


def laplacian_uniform(verts, faces):
    """'triangle' term: losses.py:422-429 -> utils/losses_util.py:340-364 -> pytorch3d.loss.mesh_laplacian_smoothing
    (method="uniform") on Meshes(verts=(B,V,3), faces=(F,3) shared).  Upstream (PARITY UNPINNED - PyTorch3D is not in
    the container; restated from loss/mesh_laplacian_smoothing.py and ops/laplacian_matrices.py):
        L[i, j] = 1 / deg(i) for every edge (i, j) of the mesh (unique, undirected), L[i, i] = -1,
        loss = sum_n (1 / V_n) sum_v | (L x)_v |_2  /  N."""
    f = faces.long()
    V = verts.shape[1]
    e = torch.cat([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], 0)
    e = torch.unique(torch.sort(e, 1)[0], dim=0)
    A = torch.zeros(V, V, dtype=verts.dtype)
    A[e[:, 0], e[:, 1]] = 1
    A[e[:, 1], e[:, 0]] = 1
    deg = A.sum(1)
    Lm = A * torch.where(deg > 0, 1.0 / deg, deg)[:, None] - torch.eye(V, dtype=verts.dtype)
    lx = torch.einsum("ij,bjc->bic", Lm, verts)
    return (lx.norm(dim=2) / V).sum() / verts.shape[0]
