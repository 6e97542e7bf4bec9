"""TEST INFRASTRUCTURE — CPU restatement (torch, autograd) of a generic articulated LBS hand layer.

The same algorithm as ManoLayer.forward (utils/my_mano.py:315-483: pose PCA + mean -> Rodrigues ->
pose map -> shape / pose blendshapes -> joint regression on the SHAPED template -> kinematic chain
-> linear blend skinning -> tip vertices appended to the chain joints), written for any joint count,
parent table, tip list and basis size, so it also covers the NIMBLE-shaped stand-in of
hifihr_b200/nimble.py (the real NIMBLE source is absent -> parity unpinned).
Pinned to the MANO oracle (which is pinned to the reference itself) by
tests/test_oracle_pins.py::test_generic_lbs_oracle_matches_mano_oracle.
"""
from __future__ import annotations

import numpy as np
import torch

from .mano import rodrigues


class LBSOracle:
    def __init__(self, v_template, shapedirs, posedirs, J_regressor, weights, parents, pca_comps=None, pose_mean=None,
                 tip_verts=(), joint_order=None, center_joint=None, dtype=torch.float64):
        t = lambda a: torch.tensor(np.asarray(a, dtype=np.float64)).to(dtype)  # noqa: E731
        self.dtype = dtype
        self.v_template, self.shapedirs, self.posedirs = t(v_template), t(shapedirs), t(posedirs)
        self.J_regressor, self.weights = t(J_regressor), t(weights)
        self.parents = [int(p) for p in parents]
        self.NJ = len(self.parents)
        self.pca = None if pca_comps is None else t(pca_comps)
        self.pose_mean = torch.zeros(3 * (self.NJ - 1), dtype=dtype) if pose_mean is None else t(pose_mean)
        self.tips = [int(v) for v in tip_verts]
        self.joint_order = None if joint_order is None else [int(j) for j in joint_order]
        self.center_joint = center_joint

    def __call__(self, pose, betas, trans=None):
        B, NJ, dt = pose.shape[0], self.NJ, self.dtype
        hand = pose[:, 3:]
        if self.pca is not None:
            hand = hand[:, :self.pca.shape[0]] @ self.pca
        full = torch.cat([pose[:, :3], self.pose_mean[None] + hand], 1)
        R = rodrigues(full.reshape(-1, 3)).view(B, NJ, 3, 3)
        pose_map = (R[:, 1:] - torch.eye(3, dtype=dt)).reshape(B, 9 * (NJ - 1))
        v_shaped = self.v_template + torch.einsum("vck,bk->bvc", self.shapedirs, betas)
        J = torch.einsum("jv,bvc->bjc", self.J_regressor, v_shaped)
        v_posed = v_shaped + torch.einsum("vck,bk->bvc", self.posedirs, pose_map)
        rot, tr = [None] * NJ, [None] * NJ                    # world rotation / translation per joint
        for j, p in enumerate(self.parents):
            if p < 0:
                rot[j], tr[j] = R[:, j], J[:, j]
            else:
                rot[j] = rot[p] @ R[:, j]
                tr[j] = tr[p] + (rot[p] @ (J[:, j] - J[:, p])[..., None])[..., 0]
        rot, tr = torch.stack(rot, 1), torch.stack(tr, 1)     # (B,NJ,3,3), (B,NJ,3)
        tA = tr - (rot @ J[..., None])[..., 0]                # remove the rest joint: x -> rot (x - J) + tr
        Tr = torch.einsum("vj,bjrc->bvrc", self.weights, rot)
        Tt = torch.einsum("vj,bjr->bvr", self.weights, tA)
        verts = (Tr @ v_posed[..., None])[..., 0] + Tt
        jtr = torch.cat([tr, verts[:, self.tips]], 1) if self.tips else tr
        if self.joint_order is not None:
            jtr = jtr[:, self.joint_order]
        if trans is not None:
            verts, jtr = verts + trans[:, None], jtr + trans[:, None]
        elif self.center_joint is not None and self.center_joint >= 0:
            c = jtr[:, self.center_joint:self.center_joint + 1]
            verts, jtr = verts - c, jtr - c
        return verts, jtr
