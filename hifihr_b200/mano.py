"""Drop-in hand layers: ManoLayer / MyMANOLayer with the reference call signatures.

ManoLayer(center_idx, flat_hand_mean, ncomps, side, mano_root, use_pca, root_rot_mode,
joint_rot_mode, robust_rot).forward(th_pose_coeffs, th_betas, th_trans, root_palm,
share_betas) -> (th_verts, th_jtr)                       — utils/my_mano.py:225-483
MyMANOLayer(ifRender, device, shape_ncomp, pose_ncomp, tex_ncomp, use_pose_pca)
.forward(hand_params, handle_collision) -> {'skin_meshes', 'mano_verts'}  — :22-54

Buffer names (`th_*`) match the reference so old state_dicts load.  The forward and
backward run as one sm_100a kernel each (csrc/mano.cu); there is no torch fallback.
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn

from . import ops
from .mano_assets import load_mano
from .structures import Meshes

TIP_VERTS = {"right": [745, 317, 444, 556, 673], "left": [745, 317, 445, 556, 673]}     # my_mano.py:455-458
JOINT_REORDER = [0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20]  # :465-469
# xyz_from_vertice: 16 regressed joints + 5 tip verts in FreiHAND order (Freihand_trainer_mano_fullsup.py:177-192)
FREI_FROM_MANO16 = {0: 0, 1: 5, 2: 6, 3: 7, 4: 9, 5: 10, 6: 11, 7: 17, 8: 18, 9: 19, 10: 13, 11: 14, 12: 15,
                    13: 1, 14: 2, 15: 3}
FREI_TIPS = {4: 744, 8: 320, 12: 443, 16: 555, 20: 672}


PALM_VERTS = (95, 22)                                                                     # my_mano.py:460


# ---- rotation-representation helpers of the non-default modes (tiny (B,6)/(B,16,3,3) device-side tensor
#      algebra ahead of the kernel; autograd carries the kernel's g_rots back through them) -----------------
def _normalize(v):
    """utils/manopth/rot6d.py:54-60 (clamped norm)."""
    mag = torch.sqrt(v.pow(2).sum(1)).clamp_min(1e-8)
    return v / mag[:, None]


def rotation_from_ortho6d(poses):
    """utils/manopth/rot6d.py:4-25: Gram-Schmidt of the two 3-vectors, columns (x, y, z)."""
    x = _normalize(poses[:, 0:3])
    z = _normalize(torch.cross(x, poses[:, 3:6], dim=1))
    y = torch.cross(z, x, dim=1)
    return torch.stack((x, y, z), 2)


def robust_rotation_from_ortho6d(poses):
    """utils/manopth/rot6d.py:27-51: symmetric orthogonalisation of the two predicted directions."""
    x, y = _normalize(poses[:, 0:3]), _normalize(poses[:, 3:6])
    middle, orthmid = _normalize(x + y), _normalize(x - y)
    x, y = _normalize(middle + orthmid), _normalize(middle - orthmid)
    z = _normalize(torch.cross(x, y, dim=1))
    return torch.stack((x, y, z), 2)


def batch_rotprojs(rotmats):
    """utils/manopth/rotproj.py:4-23: nearest rotation U V^T per 3x3 block, last column flipped on reflections
    (batched on the device instead of the reference's per-matrix CPU round trip)."""
    U, _, Vh = torch.linalg.svd(rotmats)
    R = U @ Vh
    flip = torch.where(torch.linalg.det(R) < 0, -1.0, 1.0)
    return torch.cat((R[..., :2], R[..., 2:] * flip[..., None, None]), -1)


def axis_angle_to_matrix(v):
    """rodrigues_layer.batch_rodrigues on a handful of constant vectors (buffer construction only)."""
    ang = torch.norm(v + 1e-8, p=2, dim=1, keepdim=True)
    q = torch.cat((torch.cos(ang * 0.5), torch.sin(ang * 0.5) * (v / ang)), 1)
    q = q / q.norm(p=2, dim=1, keepdim=True)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    return torch.stack((w * w + x * x - y * y - z * z, 2 * x * y - 2 * w * z, 2 * w * y + 2 * x * z,
                        2 * w * z + 2 * x * y, w * w - x * x + y * y - z * z, 2 * y * z - 2 * w * x,
                        2 * x * z - 2 * w * y, 2 * w * x + 2 * y * z, w * w - x * x - y * y + z * z), 1).view(-1, 3, 3)


def frei_out_src():
    src = [0] * 21
    for m, k in FREI_FROM_MANO16.items():
        src[k] = m
    for k, v in FREI_TIPS.items():
        src[k] = -(v + 1)
    return src


class ManoLayer(nn.Module):
    def __init__(self, center_idx=None, flat_hand_mean=True, ncomps=6, side="right", mano_root="mano/models",
                 use_pca=True, root_rot_mode="axisang", joint_rot_mode="axisang", robust_rot=False):
        super().__init__()
        if root_rot_mode not in ("axisang", "rot6d"):
            raise ValueError(f"root_rot_mode must be 'axisang' or 'rot6d', got {root_rot_mode!r}")
        if joint_rot_mode not in ("axisang", "rotmat"):
            raise ValueError(f"joint_rot_mode must be 'axisang' or 'rotmat', got {joint_rot_mode!r}")
        self.center_idx = center_idx
        self.robust_rot = robust_rot
        self.rot = 3 if root_rot_mode == "axisang" else 6           # my_mano.py:258-261
        self.flat_hand_mean = flat_hand_mean
        self.side = side
        self.use_pca = use_pca
        self.joint_rot_mode = joint_rot_mode
        self.root_rot_mode = root_rot_mode
        self.ncomps = ncomps if use_pca else 45
        d = load_mano(mano_root, side)
        self._mano = d
        comps = d["hands_components"]
        mean = np.zeros(comps.shape[1]) if flat_hand_mean else d["hands_mean"].copy()
        f32 = lambda a: torch.tensor(np.asarray(a, np.float64)).float()  # noqa: E731
        self.register_buffer("th_betas", torch.zeros(1, 10))
        self.register_buffer("th_shapedirs", f32(d["shapedirs"]))
        self.register_buffer("th_posedirs", f32(d["posedirs"]))
        self.register_buffer("th_v_template", f32(d["v_template"]).unsqueeze(0))
        self.register_buffer("th_J_regressor", f32(d["J_regressor"]))
        self.register_buffer("th_weights", f32(d["weights"]))
        self.register_buffer("th_faces", torch.tensor(d["f"].astype(np.int32)).long())
        self._rotmat_in = (not use_pca) and joint_rot_mode == "rotmat"
        if not self._rotmat_in:
            self.register_buffer("th_hands_mean", f32(mean).unsqueeze(0))
            self.register_buffer("th_comps", f32(comps))
            self.register_buffer("th_selected_comps", f32(comps[:ncomps]))
        else:   # my_mano.py:304-307: registered by the reference, not used by its forward
            self.register_buffer("th_hands_mean_rotmat", axis_angle_to_matrix(f32(mean).view(15, 3)))
        self.kintree_table = d["kintree_table"]
        parents = list(self.kintree_table[0].tolist())
        self.kintree_parents = parents
        self._parents = [-1] + [int(p) for p in parents[1:]]
        self._mean = mean
        self._consts = {}

    def consts(self, device) -> ops.HandModelConsts:
        key = str(device)
        if key not in self._consts:
            d = self._mano
            self._consts[key] = ops.HandModelConsts(
                v_template=d["v_template"], shapedirs=d["shapedirs"], posedirs=d["posedirs"],
                J_regressor=d["J_regressor"], weights=d["weights"], parents=self._parents,
                pca_comps=d["hands_components"][:self.ncomps] if self.use_pca else None, pose_mean=self._mean,
                tip_verts=TIP_VERTS[self.side], joint_order=JOINT_REORDER,
                center_joint=-1 if self.center_idx is None else int(self.center_idx), palm_verts=PALM_VERTS,
                device=device)
        return self._consts[key]

    def forward(self, th_pose_coeffs, th_betas=torch.zeros(1), th_trans=torch.zeros(1),
                root_palm=torch.Tensor([0]), share_betas=torch.Tensor([0])):
        if not th_pose_coeffs.is_cuda:
            raise RuntimeError("hifihr_b200.ManoLayer runs on CUDA tensors only (no CPU path)")
        hm = self.consts(th_pose_coeffs.device)
        rots = None
        if self._rotmat_in:
            # joint_rot_mode='rotmat' (my_mano.py:362-373): (B,16,3,3) matrices, projected onto SO(3)
            assert th_pose_coeffs.dim() == 4, (
                "When not self.use_pca, th_pose_coeffs should have 4 dims, got {}".format(th_pose_coeffs.dim()))
            assert th_pose_coeffs.shape[2:4] == (3, 3), (
                "When not self.use_pca, th_pose_coeffs have 3x3 matrix for two last dims, got {}".format(
                    th_pose_coeffs.shape[2:4]))
            pose, rots = None, batch_rotprojs(th_pose_coeffs)
        else:
            need = self.rot + (hm.pose_dim - 3)
            pose = th_pose_coeffs[:, :need]
            if pose.shape[1] != need:
                raise ValueError(f"th_pose_coeffs needs at least {need} columns, got {th_pose_coeffs.shape[1]}")
            if self.root_rot_mode == "rot6d":   # my_mano.py:355-361: root from the 6-D representation
                six = th_pose_coeffs[:, :6]
                rots = (robust_rotation_from_ortho6d(six) if self.robust_rot else rotation_from_ortho6d(six))[:, None]
        betas = None
        if th_betas is not None and th_betas.numel() != 1:
            betas = th_betas
            if bool(share_betas):
                betas = betas.mean(0, keepdim=True).expand(betas.shape[0], 10)
        trans = None
        if th_trans is not None and th_trans.numel() != 1:
            # same host check as the reference (my_mano.py:471): all-zero translation means "centre"
            if not bool(torch.norm(th_trans) == 0):
                trans = th_trans
        return ops.ManoFunction.apply(hm, pose, betas, trans, rots, self.rot, bool(root_palm))


class MyMANOLayer(nn.Module):
    def __init__(self, ifRender, device, shape_ncomp=20, pose_ncomp=30, tex_ncomp=10, use_pose_pca=True,
                 mano_root=None):
        super().__init__()
        self.pose_num = pose_ncomp
        self.mesh_num = 778
        self.bases_num = 10
        self.keypoints_num = 16
        self.device = device
        self.mano_layer = ManoLayer(center_idx=9, flat_hand_mean=False, side="right", mano_root=mano_root,
                                    use_pca=use_pose_pca, ncomps=pose_ncomp)
        self.mesh_face = self.mano_layer.th_faces.to(torch.int16)[None]       # (1,1538,3) as my_mano.py:33
        self._topo = {}

    def topology(self, device) -> ops.TopologyConsts:
        key = str(device)
        if key not in self._topo:
            d = self.mano_layer._mano
            self._topo[key] = ops.TopologyConsts(d["f"], 778, J_regressor=d["J_regressor"], out_src=frei_out_src(),
                                                 device=device)
        return self._topo[key]

    def forward(self, hand_params, handle_collision=True):
        verts, _ = self.mano_layer(hand_params["pose_params"], hand_params["shape_params"])
        meshes = Meshes(verts, self.mesh_face, topology=self.topology(verts.device))
        return {"skin_meshes": meshes, "mano_verts": verts}


def xyz_from_vertice(topo: ops.TopologyConsts, verts, root_id=None):
    """(B,778,3) -> (B,21,3) FreiHAND joints regressed from posed verts
    (dense_pose_Trainer.xyz_from_vertice; the reference returns (21,B,3) and the caller permutes).
    With root_id the joints come back root-relative together with the shifted verts."""
    outs = ops.GeomFunction.apply(topo, verts, 9 if root_id is None else root_id, None, None, None, False)
    joints_rel, verts_rel = outs[0], outs[1]
    if root_id is None:
        # un-shift: joints_rel = joints - joints[9]; recover absolute joints from the shift of vertex 0
        shift = (verts - verts_rel)[:, :1]
        return joints_rel + shift
    return joints_rel, verts_rel
