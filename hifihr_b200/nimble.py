"""NIMBLE-shaped hand layer on synthetic data.

The reference's `utils/NIMBLE_model` is an empty submodule and the NIMBLE assets are absent, so
`MyNIMBLELayer` cannot be rebuilt from source.  This module keeps the CALL CONTRACT the reference
uses (models_res_nimble.py:55-57, 133-142, 154-172, 203-205; utils/visualize_util.py:24-27):

    MyNIMBLELayer(ifRender, device, shape_ncomp=20, pose_ncomp=30, tex_ncomp=10)
        .forward(hand_params, handle_collision=False) -> dict with
        'nimble_joints' (B,25,3), 'joints' (B,21,3), 'verts' (B,V,3), 'faces', 'skin_meshes' (Meshes with a
        TexturesUV), 'mano_verts' (B,778,3), 'textures' (B,T,T,3), 'rot' (B,3)

on a seeded NIMBLE-SHAPED stand-in (seed 20231): a closed capsule-like mesh with V = 5986 vertices and
F = 11968 faces, 20 rotating joints (wrist + thumb x3 + four fingers x4) + 5 tips = 25 joints, top-4
skinning weights, 20 shape / 30 pose-PCA / 10 texture-PCA components, per-face UVs, and a texture PCA basis.
Parity for this layer is against the same LBS oracle at NIMBLE sizes, not against NIMBLE (parity unpinned).
The LBS runs in the same sm_100a kernels as MANO (csrc/mano.cu).
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn

from . import ops
from .renderer import TexturesUV, TexturesUVPCA
from .structures import Meshes

SEED = 20231
RINGS, SEGS = 68, 88          # V = 2 + RINGS*SEGS = 5986, F = 2*SEGS*RINGS = 11968
FINGER_LEN = (3, 4, 4, 4, 4)  # thumb has 3 rotating joints, the others 4 -> 19 + wrist = 20


def build_nimble_like(tex_size=1024, n_shape=20, n_pose_pca=30, n_tex=10, seed=SEED):
    """All constants of the stand-in as numpy arrays (deterministic)."""
    rng = np.random.RandomState(seed)
    # ---- mesh: an elongated, flattened UV-sphere ("mitten"): x along the fingers -------------------
    th = (np.arange(RINGS) + 1) * np.pi / (RINGS + 1)
    ph = np.arange(SEGS) * 2 * np.pi / SEGS
    ring = np.stack([np.cos(th)[:, None] * np.ones_like(ph)[None], np.sin(th)[:, None] * np.cos(ph)[None],
                     np.sin(th)[:, None] * np.sin(ph)[None]], -1).reshape(-1, 3)
    v = np.concatenate([[[1.0, 0, 0]], ring, [[-1.0, 0, 0]]], 0) * np.array([0.095, 0.045, 0.018])
    v[:, 0] += 0.02
    V = v.shape[0]
    faces = []
    for s in range(SEGS):
        s1 = (s + 1) % SEGS
        faces.append([0, 1 + s, 1 + s1])
        for r in range(RINGS - 1):
            a, b = 1 + r * SEGS + s, 1 + r * SEGS + s1
            c, d = a + SEGS, b + SEGS
            faces += [[a, c, b], [b, c, d]]
        last = 1 + (RINGS - 1) * SEGS
        faces.append([V - 1, last + s1, last + s])
    faces = np.asarray(faces, np.int64)
    # ---- skeleton: wrist + 5 chains along +x, fanned in y -----------------------------------------
    parents, J = [-1], [[-0.06, 0.0, 0.0]]
    tips_pos = []
    for f, n in enumerate(FINGER_LEN):
        y = (f - 2) * 0.018
        base_x = -0.03 if f == 0 else 0.0
        seg = (0.115 - base_x) / (n + 1)
        p = 0
        for k in range(n):
            parents.append(p)
            J.append([base_x + seg * k, y, 0.0])
            p = len(J) - 1
        tips_pos.append([base_x + seg * n, y, 0.0])
    J = np.asarray(J)
    NJ = J.shape[0]
    assert NJ == 20
    # ---- skinning: softmax(-dist^2) to bone segment midpoints, top-4 ------------------------------
    child_end = np.zeros_like(J)
    ends = {i: [] for i in range(NJ)}
    for j in range(1, NJ):
        ends[parents[j]].append(J[j])
    fi = 0
    for j in range(NJ):
        if ends[j]:
            child_end[j] = np.mean(ends[j], 0)
        else:
            child_end[j] = J[j] + np.array([0.02, 0, 0])
    mid = 0.5 * (J + child_end)
    d2 = ((v[:, None] - mid[None]) ** 2).sum(-1)
    w = np.exp(-d2 / (2 * 0.012 ** 2))
    idx = np.argsort(-w, 1)[:, :4]
    wt = np.zeros_like(w)
    np.put_along_axis(wt, idx, np.take_along_axis(w, idx, 1), 1)
    wt /= wt.sum(1, keepdims=True)
    # ---- joint regressor: each joint = mean of its 24 nearest template verts, shifted exactly onto J is not
    #      needed (J_template = Jreg @ v_template defines the rest joints, as in MANO) ------------------
    Jreg = np.zeros((NJ, V))
    for j in range(NJ):
        nn_ = np.argsort(((v - J[j]) ** 2).sum(1))[:24]
        Jreg[j, nn_] = 1.0 / 24
    tip_verts = [int(np.argmin(((v - np.asarray(t)) ** 2).sum(1) - 1e3 * (v[:, 2] > 0))) for t in tips_pos]
    shapedirs = rng.randn(V, 3, n_shape) * 1e-3
    posedirs = rng.randn(V, 3, 9 * (NJ - 1)) * 1e-3
    pca = rng.randn(n_pose_pca, 3 * (NJ - 1)) * 0.3
    pose_mean = rng.randn(3 * (NJ - 1)) * 0.05
    # ---- UVs: cylindrical unwrap (per-vertex; seam faces stretch across, fine for a stand-in) ------
    u = (np.arctan2(v[:, 2] / 0.018, v[:, 1] / 0.045) + np.pi) / (2 * np.pi)
    vv = (v[:, 0] - v[:, 0].min()) / (v[:, 0].max() - v[:, 0].min())
    verts_uvs = np.stack([u, vv], 1).astype(np.float32)
    # ---- texture PCA: smooth random fields upsampled from 16x16 -----------------------------------
    def smooth(n):
        low = torch.tensor(rng.randn(n, 3, 16, 16), dtype=torch.float32)
        up = torch.nn.functional.interpolate(low, size=(tex_size, tex_size), mode="bilinear", align_corners=True)
        return up.permute(0, 2, 3, 1).contiguous()
    tex_mean = (0.6 + 0.1 * smooth(1))[0]
    tex_basis = 0.05 * smooth(n_tex)
    # ---- skin -> MANO vertex map: 778 seeded vertex picks (stands in for NIMBLE's landmark map) -----
    mano_map = np.sort(rng.choice(V, 778, replace=False))
    # 21 "MANO-order" joints out of the 25: wrist, 4 per finger (thumb: 3 joints + tip, others: last 3 + tip)
    chain_ids, c = [], 1
    for n in FINGER_LEN:
        chain_ids.append(list(range(c, c + n)))
        c += n
    mano21 = [0]
    for f, ids in enumerate(chain_ids):
        mano21 += ids[-3:] + [NJ + f]
    return dict(v_template=v, faces=faces, J_regressor=Jreg, weights=wt, parents=parents, shapedirs=shapedirs,
                posedirs=posedirs, pca_comps=pca, pose_mean=pose_mean, tip_verts=tip_verts, verts_uvs=verts_uvs,
                tex_mean=tex_mean, tex_basis=tex_basis, mano_map=mano_map, mano21=mano21)


class MyNIMBLELayer(nn.Module):
    def __init__(self, ifRender, device, shape_ncomp=20, pose_ncomp=30, tex_ncomp=10, tex_size=1024, fused_texture=False):
        super().__init__()
        self.device = torch.device(device)
        self.ifRender = ifRender
        self.fused_texture = fused_texture   # True: skin_meshes carry a TexturesUVPCA, no per-sample maps are built
        d = build_nimble_like(tex_size, shape_ncomp, pose_ncomp, tex_ncomp)
        self._d = d
        self.V, self.F = d["v_template"].shape[0], d["faces"].shape[0]
        self.register_buffer("faces", torch.tensor(d["faces"]))
        self.register_buffer("verts_uvs", torch.tensor(d["verts_uvs"]))
        self.register_buffer("tex_mean", d["tex_mean"])
        self.register_buffer("tex_basis", d["tex_basis"])
        self.register_buffer("mano_map", torch.tensor(d["mano_map"]))
        self.register_buffer("mano21", torch.tensor(d["mano21"]))
        self._consts, self._topo = {}, {}

    def consts(self, device):
        key = str(device)
        if key not in self._consts:
            d = self._d
            self._consts[key] = ops.HandModelConsts(
                v_template=d["v_template"], shapedirs=d["shapedirs"], posedirs=d["posedirs"], J_regressor=d["J_regressor"],
                weights=d["weights"], parents=d["parents"], pca_comps=d["pca_comps"], pose_mean=d["pose_mean"],
                tip_verts=d["tip_verts"], joint_order=None, center_joint=-1, max_influences=4, device=device)
            self._topo[key] = ops.TopologyConsts(d["faces"], self.V, device=device)
        return self._consts[key], self._topo[key]

    def forward(self, hand_params, handle_collision=False):
        if handle_collision:
            raise NotImplementedError("collision handling is disabled by the reference (models_res_nimble.py:133)")
        pose, shape = hand_params["pose_params"], hand_params["shape_params"]
        hm, topo = self.consts(pose.device)
        rot = pose[:, :3]
        verts, joints25 = ops.ManoFunction.apply(hm, pose[:, :hm.pose_dim], shape, None, None, 3, False)
        out = {"nimble_joints": joints25, "joints": joints25[:, self.mano21], "verts": verts, "faces": self.faces,
               "mano_verts": verts[:, self.mano_map], "rot": rot}
        tex_img = None
        meshes = Meshes(verts, self.faces, topology=topo)
        if self.ifRender and hand_params.get("texture_params") is not None:
            tp = hand_params["texture_params"]                                    # (B, 10)
            T = self.tex_mean.shape[0]
            if self.fused_texture:
                # the shader kernels evaluate mean + params @ basis at the taps of every fragment: the (B,T,T,3) maps
                # never exist; 'textures' is then None - ask for maps_padded() to export them
                meshes.textures = TexturesUVPCA(self.tex_mean, self.tex_basis, tp, self.faces, self.verts_uvs)
            else:
                # per-sample diffuse map = mean + params @ basis (plain library GEMM), as the reference's layer returns it
                tex_img = (self.tex_mean.reshape(1, -1) + tp @ self.tex_basis.reshape(tp.shape[1], -1)).view(-1, T, T, 3)
                meshes.textures = TexturesUV(tex_img, self.faces, self.verts_uvs)
        out["skin_meshes"] = meshes
        out["textures"] = tex_img
        return out
