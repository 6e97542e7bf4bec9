"""TEST INFRASTRUCTURE — CPU restatement (torch, any float dtype, autograd) of the MANO path.

Restates, in a different structure (explicit walk over the kinematic tree),
what the reference computes in
  * ManoLayer.forward            utils/my_mano.py:315-483
  * batch_rodrigues / quat2mat   utils/manopth/rodrigues_layer.py:15-54
  * th_posemap_axisang           utils/manopth/tensutils.py:6-42
  * xyz_from_vertice (+ keypoint reorder)
                                 utils/Freihand_GNN_mano/Freihand_trainer_mano_fullsup.py:175-215
Pinned against the unmodified reference in oracle/ref_mano.py (build container)
and against tests/golden/mano_*.npz.
"""
from __future__ import annotations

import numpy as np
import torch

PARENTS = [-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11, 0, 13, 14]      # kintree_table[0]
TIP_VERTS_RIGHT = [745, 317, 444, 556, 673]                           # my_mano.py:456
JOINT_REORDER = [0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20]  # :465-469
# Freihand_trainer_mano_fullsup.py:177-192
FREI_FROM_MANO16 = {0: 0, 1: 5, 2: 6, 3: 7, 4: 9, 5: 10, 6: 11, 7: 17, 8: 18, 9: 19,
                    10: 13, 11: 14, 12: 15, 13: 1, 14: 2, 15: 3}
FREI_TIPS = {4: 744, 8: 320, 12: 443, 16: 555, 20: 672}


def rodrigues(axisang: torch.Tensor) -> torch.Tensor:
    """(N,3) axis-angle -> (N,3,3).  rodrigues_layer.py:43-54 + quat2mat :15-40."""
    angle = torch.norm(axisang + 1e-8, p=2, dim=1, keepdim=True)
    axis = axisang / angle
    half = angle * 0.5
    q = torch.cat([torch.cos(half), torch.sin(half) * axis], dim=1)
    q = q / q.norm(p=2, dim=1, keepdim=True)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    w2, x2, y2, z2 = w * w, x * x, y * y, z * z
    wx, wy, wz, xy, xz, yz = w * x, w * y, w * z, x * y, x * z, y * z
    R = torch.stack([w2 + x2 - y2 - z2, 2 * xy - 2 * wz, 2 * wy + 2 * xz,
                     2 * wz + 2 * xy, w2 - x2 + y2 - z2, 2 * yz - 2 * wx,
                     2 * xz - 2 * wy, 2 * wx + 2 * yz, w2 - x2 - y2 + z2], dim=1)
    return R.view(-1, 3, 3)


def _unit(v):
    """utils/manopth/rot6d.py:54-60."""
    return v / torch.sqrt((v * v).sum(1, keepdim=True)).clamp_min(1e-8)


def ortho6d_to_matrix(p, robust=False):
    """utils/manopth/rot6d.py:4-25 (Gram-Schmidt) and :27-51 (robust variant); columns are (x, y, z)."""
    a, b = p[:, 0:3], p[:, 3:6]
    if robust:
        a, b = _unit(a), _unit(b)
        mid, orth = _unit(a + b), _unit(a - b)
        x, y = _unit(mid + orth), _unit(mid - orth)
        z = _unit(torch.linalg.cross(x, y))
    else:
        x = _unit(a)
        z = _unit(torch.linalg.cross(x, b))
        y = torch.linalg.cross(z, x)
    return torch.stack([x, y, z], 2)


def project_rotations(M):
    """utils/manopth/rotproj.py:4-23: U V^T of every 3x3 block, third column negated when det < 0."""
    out = []
    for m in M.reshape(-1, 3, 3):
        U, _, V = torch.svd(m)
        R = U @ V.t()
        if torch.det(R) < 0:
            R = torch.cat([R[:, :2], -R[:, 2:]], 1)
        out.append(R)
    return torch.stack(out).view(M.shape)


class ManoOracle:
    def __init__(self, mano: dict, ncomps=48, flat_hand_mean=False, center_idx=9, use_pca=True,
                 dtype=torch.float32, root_rot_mode="axisang", joint_rot_mode="axisang", robust_rot=False):
        self.root_rot_mode, self.joint_rot_mode, self.robust_rot = root_rot_mode, joint_rot_mode, robust_rot
        self.rot = 3 if root_rot_mode == "axisang" else 6
        t = lambda a: torch.tensor(np.asarray(a, dtype=np.float64)).to(dtype)  # noqa: E731
        self.dtype = dtype
        self.center_idx = center_idx
        self.use_pca = use_pca
        self.ncomps = ncomps if use_pca else 45
        self.shapedirs = t(mano["shapedirs"])                 # (778,3,10)
        self.posedirs = t(mano["posedirs"])                   # (778,3,135)
        self.v_template = t(mano["v_template"])               # (778,3)
        self.J_regressor = t(mano["J_regressor"])             # (16,778)
        self.weights = t(mano["weights"])                     # (778,16)
        self.faces = torch.tensor(np.asarray(mano["f"], dtype=np.int64))
        hm = np.zeros(45) if flat_hand_mean else np.asarray(mano["hands_mean"])
        self.hands_mean = t(hm)[None]
        self.selected_comps = t(mano["hands_components"][:ncomps])

    def __call__(self, pose, betas, trans=None, root_palm=False, share_betas=False):
        B = pose.shape[0]
        dt = self.dtype
        if not self.use_pca and self.joint_rot_mode == "rotmat":             # my_mano.py:362-373
            R = project_rotations(pose)
        else:
            hand = pose[:, self.rot:self.rot + self.ncomps]
            if self.use_pca:
                hand = hand @ self.selected_comps
            hand = self.hands_mean + hand
            if self.root_rot_mode == "axisang":                                # :349-354
                R = rodrigues(torch.cat([pose[:, :3], hand], 1).reshape(-1, 3)).view(B, 16, 3, 3)
            else:                                                              # :355-361
                R = torch.cat([ortho6d_to_matrix(pose[:, :6], self.robust_rot)[:, None],
                               rodrigues(hand.reshape(-1, 3)).view(B, 15, 3, 3)], 1)
        if betas is None or betas.numel() == 1:                                # :376-381 mean shape
            betas = torch.zeros(B, 10, dtype=dt)
        elif share_betas:                                                      # :384-385
            betas = betas.mean(0, keepdim=True).expand(B, 10)
        pose_map = (R[:, 1:] - torch.eye(3, dtype=dt)).reshape(B, 135)
        v_shaped = torch.einsum("vck,bk->bvc", self.shapedirs, betas) + self.v_template
        J = torch.einsum("jv,bvc->bjc", self.J_regressor, v_shaped)           # (B,16,3)
        v_posed = v_shaped + torch.einsum("vck,bk->bvc", self.posedirs, pose_map)
        # kinematic chain: G_j = G_parent * [R_j | J_j - J_parent]
        G = [None] * 16
        for j in range(16):
            p = PARENTS[j]
            tj = J[:, j] if p < 0 else J[:, j] - J[:, p]
            Tl = torch.zeros(B, 4, 4, dtype=dt)
            Tl = Tl + torch.eye(4, dtype=dt)
            Tl = torch.cat([torch.cat([R[:, j], tj[:, :, None]], 2),
                            torch.tensor([0, 0, 0, 1.0], dtype=dt).expand(B, 1, 4)], 1)
            G[j] = Tl if p < 0 else G[p] @ Tl
        G = torch.stack(G, 1)                                                  # (B,16,4,4)
        Jh = torch.cat([J, torch.zeros(B, 16, 1, dtype=dt)], 2)
        corr = (G @ Jh[..., None])                                             # (B,16,4,1)
        A = G - torch.cat([torch.zeros(B, 16, 4, 3, dtype=dt), corr], 3)
        T = torch.einsum("bjrc,vj->bvrc", A, self.weights)                     # (B,778,4,4)
        vh = torch.cat([v_posed, torch.ones(B, v_posed.shape[1], 1, dtype=dt)], 2)
        verts = torch.einsum("bvrc,bvc->bvr", T, vh)[..., :3]
        chain = G[:, :, :3, 3]
        if root_palm:                                                          # :459-461
            chain = torch.cat([((verts[:, 95] + verts[:, 22]) / 2)[:, None], chain[:, 1:]], 1)
        jtr = torch.cat([chain, verts[:, TIP_VERTS_RIGHT]], 1)[:, JOINT_REORDER]
        if trans is None or bool(torch.norm(trans) == 0):                      # :471
            if self.center_idx is not None:
                c = jtr[:, self.center_idx:self.center_idx + 1]
                jtr = jtr - c
                verts = verts - c
        else:
            jtr = jtr + trans[:, None]
            verts = verts + trans[:, None]
        return verts, jtr

    def xyz_from_vertice(self, verts):
        """(B,778,3) -> (B,21,3) FreiHAND-ordered joints regressed from POSED verts."""
        j16 = torch.einsum("jv,bvc->bjc", self.J_regressor, verts)
        out = [None] * 21
        for m, k in FREI_FROM_MANO16.items():
            out[k] = j16[:, m]
        for k, v in FREI_TIPS.items():
            out[k] = verts[:, v]
        return torch.stack(out, 1)
