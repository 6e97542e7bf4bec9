"""Device constants + autograd bridges over the C-ABI (hifihr_b200._lib).

Every function here launches hand-written sm_100a kernels on the current torch
stream.  torch is used for memory, streams and autograd bookkeeping only.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib as L

F32, I32, I64 = torch.float32, torch.int32, torch.int64


def _cu(t, dtype=F32):
    return t.to(dtype).contiguous()


# ------------------------------------------------------------------------------------------------
# constants
# ------------------------------------------------------------------------------------------------
class HandModelConsts:
    """Device-resident constants of an LBS hand model in the layout the kernels want
    (HfrHandModel).  Built from the arrays ManoLayer.__init__ reads (utils/my_mano.py:277-313)."""

    def __init__(self, *, v_template, shapedirs, posedirs, J_regressor, weights, parents, pca_comps=None,
                 pose_mean=None, tip_verts=(), joint_order=None, center_joint=-1, max_influences=8, palm_verts=(-1, -1),
                 device="cuda"):
        v_template = np.asarray(v_template, np.float64)
        shapedirs = np.asarray(shapedirs, np.float64)
        posedirs = np.asarray(posedirs, np.float64)
        Jreg = np.asarray(J_regressor, np.float64)
        weights = np.asarray(weights, np.float64)
        V, NS = v_template.shape[0], shapedirs.shape[2]
        NJ = Jreg.shape[0]
        assert posedirs.shape[2] == 9 * (NJ - 1), "posedirs must have 9*(NJ-1) columns"
        C3 = (3 * V + 3) // 4 * 4
        dirs = np.zeros((NS + 9 * (NJ - 1), C3), np.float32)
        dirs[:NS, :3 * V] = shapedirs.reshape(3 * V, NS).T
        dirs[NS:, :3 * V] = posedirs.reshape(3 * V, -1).T
        vt = np.zeros(C3, np.float32)
        vt[:3 * V] = v_template.reshape(-1)
        nw = int(min(max_influences, max(1, (weights != 0).sum(1).max())))
        order = np.argsort(-np.abs(weights), axis=1)[:, :nw]
        skin_idx = order.T.astype(np.int32).copy()                       # (NW,V)
        skin_w = np.take_along_axis(weights, order, 1).T.astype(np.float32).copy()
        self.V, self.NJ, self.NS, self.NW, self.C3 = V, NJ, NS, nw, C3
        self.NPC = 0 if pca_comps is None else int(np.asarray(pca_comps).shape[0])
        self.NT = len(tip_verts)
        self.center_joint = int(center_joint)
        if joint_order is None:
            joint_order = list(range(NJ + self.NT))
        dev = torch.device(device)
        t = lambda a, dt=F32: torch.as_tensor(np.ascontiguousarray(a)).to(dt).to(dev).contiguous()  # noqa: E731
        self.dirs = t(dirs)
        self.v_template = t(vt)
        self.J_template = t(Jreg @ v_template)
        self.J_shapedirs = t(np.einsum("jv,vck->jck", Jreg, shapedirs))
        self.pca_comps = None if pca_comps is None else t(np.asarray(pca_comps, np.float64))
        self.pose_mean = None if pose_mean is None else t(np.asarray(pose_mean, np.float64).reshape(-1))
        self.parents = t(np.asarray(parents, np.int64), I32)
        self.skin_idx = t(skin_idx, I32)
        self.skin_w = t(skin_w)
        # joint-major view of the same non-zero weights (CSR over joints) for the backward's gA reduction
        jv_ptr, jv_vert, jv_w = [0], [], []
        for j in range(NJ):
            for i in range(nw):
                sel = np.nonzero((skin_idx[i] == j) & (skin_w[i] != 0))[0]
                jv_vert += sel.tolist()
                jv_w += skin_w[i, sel].tolist()
            jv_ptr.append(len(jv_vert))
        self.jv_ptr, self.jv_vert, self.jv_w = t(jv_ptr, I32), t(jv_vert or [0], I32), t(np.asarray(jv_w or [0.0], np.float32))
        self.tip_verts = t(np.asarray(list(tip_verts) or [0], np.int64), I32)
        self.joint_order = t(np.asarray(joint_order, np.int64), I32)
        self.pose_dim = 3 + (self.NPC if self.NPC > 0 else 3 * (NJ - 1))
        self.n_out_joints = NJ + self.NT
        s = L.HfrHandModel()
        s.V, s.NJ, s.NS, s.NPC, s.NW, s.NT, s.center_joint, s.C3 = V, NJ, NS, self.NPC, nw, self.NT, self.center_joint, C3
        s.dirs, s.v_template = self.dirs.data_ptr(), self.v_template.data_ptr()
        s.J_template, s.J_shapedirs = self.J_template.data_ptr(), self.J_shapedirs.data_ptr()
        s.pca_comps = None if self.pca_comps is None else self.pca_comps.data_ptr()
        s.pose_mean = None if self.pose_mean is None else self.pose_mean.data_ptr()
        s.parents, s.skin_idx, s.skin_w = self.parents.data_ptr(), self.skin_idx.data_ptr(), self.skin_w.data_ptr()
        s.tip_verts, s.joint_order = self.tip_verts.data_ptr(), self.joint_order.data_ptr()
        s.palm_verts = (C.c_int32 * 2)(int(palm_verts[0]), int(palm_verts[1]))
        s.jv_ptr, s.jv_vert, s.jv_w = self.jv_ptr.data_ptr(), self.jv_vert.data_ptr(), self.jv_w.data_ptr()
        self.struct = s
        self.device = dev
        # batched tensor-core path: the basis split into tf32 hi / lo parts, pre-tiled in the MMA operand layout (once)
        self.basis_packed = None
        self.batched_min = 8          # below this batch the per-sample kernels are used (no workspace is passed)
        if dev.type == "cuda":
            nbytes = int(L.lib().hfr_mano_packed_basis_bytes(C.byref(s)))
            if nbytes > 0:
                self.basis_packed = torch.empty(nbytes // 4, dtype=F32, device=dev)
                L._pending_devices().add(dev.index if dev.index is not None else torch.cuda.current_device())
                L.call("hfr_mano_pack_basis", s, C.c_void_p(self.basis_packed.data_ptr()))
                s.basis_packed = self.basis_packed.data_ptr()

    def workspace(self, B):
        """Scratch of the batched path for B samples (None = use the per-sample kernels)."""
        if self.basis_packed is None or B < self.batched_min:
            return None
        n = int(L.lib().hfr_mano_workspace_bytes(C.byref(self.struct), int(B)))
        return torch.empty((n + 3) // 4, dtype=F32, device=self.device)


class TopologyConsts:
    """Shared mesh topology + (optional) sparse joint regressor on posed verts (HfrTopology)."""

    def __init__(self, faces, V, J_regressor=None, out_src=None, device="cuda"):
        faces = np.asarray(faces, np.int64)
        F = faces.shape[0]
        inc = [[] for _ in range(V)]
        for f in range(F):
            for c in range(3):
                inc[faces[f, c]].append(f * 4 + c)
        vf_ptr = np.zeros(V + 1, np.int64)
        vf_ptr[1:] = np.cumsum([len(x) for x in inc])
        vf_idx = np.asarray([e for x in inc for e in x] or [0], np.int64)
        vf_nbr = np.stack([faces[vf_idx >> 2, ((vf_idx & 3) + 1) % 3], faces[vf_idx >> 2, ((vf_idx & 3) + 2) % 3]], 1) \
            if F > 0 else np.zeros((1, 2), np.int64)
        dev = torch.device(device)
        t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a)).to(dt).to(dev).contiguous()  # noqa: E731
        self.V, self.F = int(V), int(F)
        self.faces = t(faces, I32)
        self.faces_long = t(faces, I64)
        self.vf_ptr, self.vf_idx, self.vf_nbr = t(vf_ptr, I32), t(vf_idx, I32), t(vf_nbr, I32)
        # vertex neighbourhoods over the unique edges (uniform Laplacian, pytorch3d Meshes.laplacian_packed)
        e = np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]], 0) if F > 0 else np.zeros((0, 2), np.int64)
        e = np.unique(np.sort(e, 1), axis=0)
        both = np.concatenate([e, e[:, ::-1]], 0)
        both = both[np.lexsort((both[:, 1], both[:, 0]))]
        nbr_ptr = np.zeros(V + 1, np.int64)
        np.add.at(nbr_ptr, both[:, 0] + 1, 1)
        self.nbr_ptr, self.nbr_idx = t(np.cumsum(nbr_ptr), I32), t(both[:, 1] if len(both) else np.zeros(1, np.int64), I32)
        s = L.HfrTopology()
        s.V, s.F = self.V, self.F
        s.faces, s.vf_ptr, s.vf_idx = self.faces.data_ptr(), self.vf_ptr.data_ptr(), self.vf_idx.data_ptr()
        s.vf_nbr = self.vf_nbr.data_ptr()
        self.NJR = self.NOUT = 0
        if J_regressor is not None:
            J = np.asarray(J_regressor, np.float64)
            NJR = J.shape[0]
            jr_ptr, jr_col, jr_val = [0], [], []
            for j in range(NJR):
                nz = np.nonzero(J[j])[0]
                jr_col += nz.tolist()
                jr_val += J[j, nz].tolist()
                jr_ptr.append(len(jr_col))
            vj_ptr, vj_row, vj_val = [0], [], []
            for v in range(V):
                nz = np.nonzero(J[:, v])[0]
                vj_row += nz.tolist()
                vj_val += J[nz, v].tolist()
                vj_ptr.append(len(vj_row))
            self.jr_ptr, self.jr_col, self.jr_val = t(jr_ptr, I32), t(jr_col, I32), t(jr_val, F32)
            self.vj_ptr, self.vj_row, self.vj_val = t(vj_ptr, I32), t(vj_row, I32), t(vj_val, F32)
            self.out_src = t(np.asarray(out_src, np.int64), I32)
            self.NJR, self.NOUT = NJR, len(out_src)
            s.NJR, s.NOUT = self.NJR, self.NOUT
            s.jr_ptr, s.jr_col, s.jr_val = self.jr_ptr.data_ptr(), self.jr_col.data_ptr(), self.jr_val.data_ptr()
            s.vj_ptr, s.vj_row, s.vj_val = self.vj_ptr.data_ptr(), self.vj_row.data_ptr(), self.vj_val.data_ptr()
            s.out_src = self.out_src.data_ptr()
        self.struct = s
        self.device = dev


# ------------------------------------------------------------------------------------------------
# raw launches (no autograd) — also used by the fused step
# ------------------------------------------------------------------------------------------------
def mano_forward_raw(hm: HandModelConsts, pose, betas, trans, verts, joints, rots=None, pose_off=3, root_palm=False,
                     workspace=None):
    """rots (B,n,3,3) with n = 1 (matrix-driven root, rot6d mode) or NJ (rotmat mode, pose may be None).
    workspace (hm.workspace(B)) selects the batched tensor-core path."""
    n_rot = 0 if rots is None else rots.shape[1]
    a = L.HfrManoFwdArgs(verts.shape[0], L.ptr(pose, F32, "pose"), L.ptr(betas, F32, "betas"),
                         L.ptr(trans, F32, "trans"), L.ptr(verts, F32), L.ptr(joints, F32), L.ptr(rots, F32, "rots"),
                         n_rot, pose_off, int(root_palm), L.ptr(workspace, F32))
    L.call("hfr_mano_forward", hm.struct, a)


def mano_backward_raw(hm, pose, betas, trans, g_verts, g_joints, g_pose, g_betas, g_trans, rots=None, pose_off=3,
                      root_palm=False, g_rots=None, workspace=None, reuse_forward=False):
    n_rot = 0 if rots is None else rots.shape[1]
    a = L.HfrManoBwdArgs(g_verts.shape[0], L.ptr(pose, F32), L.ptr(betas, F32), L.ptr(trans, F32),
                         L.ptr(g_verts, F32), L.ptr(g_joints, F32), L.ptr(g_pose, F32), L.ptr(g_betas, F32),
                         L.ptr(g_trans, F32), L.ptr(rots, F32), n_rot, pose_off, int(root_palm), L.ptr(g_rots, F32),
                         L.ptr(workspace, F32), int(bool(reuse_forward) and workspace is not None))
    L.call("hfr_mano_backward", hm.struct, a)


def geom_forward_raw(topo, verts, root_out, root_xyz, focal, prp, joints, verts_rel, verts_view, verts_ndc, vnormals,
                     face_verts=None):
    a = L.HfrGeomFwdArgs(verts.shape[0], root_out, L.ptr(verts, F32, "verts"), L.ptr(root_xyz, F32), L.ptr(focal, F32),
                         L.ptr(prp, F32), L.ptr(joints, F32), L.ptr(verts_rel, F32), L.ptr(verts_view, F32),
                         L.ptr(verts_ndc, F32), L.ptr(vnormals, F32), L.ptr(face_verts, F32))
    L.call("hfr_geom_forward", topo.struct, a)


def geom_backward_raw(topo, verts, root_out, root_xyz, focal, prp, g_joints, g_rel, g_view, g_ndc, g_vn, g_verts,
                      face_rec=None, raster_ws=None, status=None, rec_partial=None):
    a = L.HfrGeomBwdArgs(verts.shape[0], root_out, L.ptr(verts, F32), L.ptr(root_xyz, F32), L.ptr(focal, F32),
                         L.ptr(prp, F32), L.ptr(g_joints, F32), L.ptr(g_rel, F32), L.ptr(g_view, F32),
                         L.ptr(g_ndc, F32), L.ptr(g_vn, F32), L.ptr(g_verts, F32), L.ptr(face_rec, F32),
                         L.ptr(raster_ws), L.ptr(status), L.ptr(rec_partial, F32))
    L.call("hfr_geom_backward", topo.struct, a)


def raster_args(face_verts, mesh_first, mesh_nf, H, W, K, blur_radius, perspective_correct, clip_bary, cull,
                pix_to_face, zbuf, bary, dists, workspace, tile_queue=None):
    N = mesh_first.shape[0]
    return L.HfrRasterArgs(N, H, W, K, face_verts.shape[0], L.ptr(face_verts, F32, "face_verts"),
                           L.ptr(mesh_first, I64, "mesh_to_face_first_idx"), L.ptr(mesh_nf, I64, "num_faces_per_mesh"),
                           float(blur_radius), int(perspective_correct), int(clip_bary), int(cull),
                           L.ptr(pix_to_face, I64), L.ptr(zbuf, F32), L.ptr(bary, F32), L.ptr(dists, F32),
                           L.ptr(workspace), L.ptr(tile_queue))


def raster_tile_queue(N, H, W, device):
    """Workspace of the cost-ordered tile queue of the fused rasterize+shade kernel."""
    return torch.empty(int(L.lib().hfr_raster_queue_bytes(int(N), int(H), int(W))), dtype=torch.uint8, device=device)


def raster_workspace(Ftot, device):
    nbytes = int(L.lib().hfr_raster_workspace_bytes(int(Ftot)))
    return torch.empty(nbytes, dtype=torch.uint8, device=device)


def raster_tile_box(ws, Ftot, N):
    """Device address of the per-mesh tile box inside a filled rasterizer workspace (None if absent)."""
    return L.lib().hfr_raster_tile_box(ws.data_ptr(), int(Ftot), int(N))


def shade_params(N, H, W, K, F, V, blend, shade, sigma, gamma, background, light_ambient, light_specular,
                 mat_ambient, mat_diffuse, mat_specular, shininess, tex_shape=(1, 1, 1), VT=0, znear=1.0, zfar=100.0,
                 tex_pca=0, light_point=0, tex_basis_stride=0):
    p = L.HfrShadeParams()
    p.N, p.H, p.W, p.K, p.F, p.V, p.blend, p.shade = N, H, W, K, F, V, blend, shade
    p.sigma, p.gamma, p.znear, p.zfar = float(sigma), float(gamma), float(znear), float(zfar)
    p.background = L.f3(background)
    p.light_ambient, p.light_specular = L.f3(light_ambient), L.f3(light_specular)
    p.mat_ambient, p.mat_diffuse, p.mat_specular = L.f3(mat_ambient), L.f3(mat_diffuse), L.f3(mat_specular)
    p.shininess = float(shininess)
    p.tex_n, p.tex_h, p.tex_w, p.VT = int(tex_shape[0]), int(tex_shape[1]), int(tex_shape[2]), int(VT)
    p.tex_pca = int(tex_pca)
    p.light_point = int(light_point)
    p.tex_basis_stride = int(tex_basis_stride)
    return p


def pack_tex_basis(basis):
    """Texel-major copy of a (n_comp,T,T,3) texture PCA basis: (T*T, 12*ceil(n_comp/4)) floats, component k / channel c
    of a texel at 3k + c, zero padded (HfrShadeParams.tex_basis_stride).  Built once per basis tensor (cached on the
    tensor, keyed by its version counter): a layout change of a constant, like the packed hand-layer basis."""
    b = basis.detach()
    cached = getattr(basis, "_hfr_texel_major", None)
    if cached is not None and cached[0] == (b._version, b.data_ptr()):
        return cached[1]
    n = b.shape[0]
    stride = 12 * ((n + 3) // 4)
    packed = torch.zeros(b.shape[1] * b.shape[2], stride, dtype=F32, device=b.device)
    packed[:, :3 * n] = b.to(F32).permute(1, 2, 0, 3).reshape(-1, 3 * n)
    try:
        basis._hfr_texel_major = ((b._version, b.data_ptr()), packed)
    except AttributeError:
        pass
    return packed


def shade_fwd_args(p, frags, faces, verts_view, vnormals, faces_uvs, verts_uvs, texture, light_dir, light_color, image,
                   face_attr=None, tex_basis=None, tex_params=None):
    p2f, zbuf, bary, dists = frags
    return L.HfrShadeFwdArgs(p, L.ptr(p2f, I64), L.ptr(zbuf, F32), L.ptr(bary, F32), L.ptr(dists, F32),
                             L.ptr(faces, I32), L.ptr(verts_view, F32), L.ptr(vnormals, F32), L.ptr(faces_uvs, I32),
                             L.ptr(verts_uvs, F32), L.ptr(texture, F32), L.ptr(light_dir, F32), L.ptr(light_color, F32),
                             L.ptr(image, F32), L.ptr(face_attr, F32), L.ptr(tex_basis, F32), L.ptr(tex_params, F32))


def face_attr_forward(faces, verts_view, vnormals, faces_uvs, verts_uvs, out):
    """Pack the per-(mesh, face) attribute records the shaders read (hfr_face_attr_forward)."""
    N, V = verts_view.shape[0], verts_view.shape[1]
    a = L.HfrFaceAttrArgs(N, faces.shape[0], V, verts_uvs.shape[0], L.ptr(faces, I32), L.ptr(verts_view, F32),
                          L.ptr(vnormals, F32), L.ptr(faces_uvs, I32), L.ptr(verts_uvs, F32), L.ptr(out, F32))
    L.call("hfr_face_attr_forward", a)


_GAUSS = {}


def gauss_taps(device):
    """The 11 fp32 taps of utils/pytorch_ssim/__init__.py:7-9 (sigma 1.5), built the same way."""
    key = str(device)
    if key not in _GAUSS:
        g = torch.tensor([math.exp(-(x - 5) ** 2 / float(2 * 1.5 ** 2)) for x in range(11)])
        _GAUSS[key] = (g / g.sum()).to(F32).to(device).contiguous()
    return _GAUSS[key]


# ------------------------------------------------------------------------------------------------
# autograd bridges
# ------------------------------------------------------------------------------------------------
class ManoFunction(torch.autograd.Function):
    """pose (B, pose_off + ncomps) or None (all joints matrix-driven), rots (B,1|NJ,3,3) or None."""

    @staticmethod
    def forward(ctx, hm: HandModelConsts, pose, betas, trans, rots=None, pose_off=3, root_palm=False):
        pose = None if pose is None else _cu(pose)
        betas = None if betas is None else _cu(betas)
        trans = None if trans is None else _cu(trans)
        rots = None if rots is None else _cu(rots)
        ref = pose if pose is not None else rots
        B = ref.shape[0]
        verts = torch.empty(B, hm.V, 3, device=ref.device, dtype=F32)
        joints = torch.empty(B, hm.n_out_joints, 3, device=ref.device, dtype=F32)
        ws = hm.workspace(B)      # kept for the backward: pose state and posed rest vertices are not recomputed
        mano_forward_raw(hm, pose, betas, trans, verts, joints, rots, pose_off, root_palm, workspace=ws)
        ctx.hm, ctx.cfg, ctx.ws = hm, (pose_off, bool(root_palm), B), ws
        ctx.save_for_backward(pose, betas, trans, rots)
        return verts, joints

    @staticmethod
    def backward(ctx, g_verts, g_joints):
        pose, betas, trans, rots = ctx.saved_tensors
        hm = ctx.hm
        pose_off, root_palm, B = ctx.cfg
        ref = pose if pose is not None else rots
        g_verts = ref.new_zeros(B, hm.V, 3) if g_verts is None else _cu(g_verts)
        g_joints = None if g_joints is None else _cu(g_joints)
        g_pose = None if pose is None else torch.empty_like(pose)
        g_betas = None if betas is None else torch.empty_like(betas)
        g_trans = None if trans is None else torch.empty_like(trans)
        g_rots = None if rots is None else torch.empty_like(rots)
        mano_backward_raw(hm, pose, betas, trans, g_verts, g_joints, g_pose, g_betas, g_trans, rots, pose_off, root_palm,
                          g_rots, workspace=ctx.ws, reuse_forward=True)
        return None, g_pose, g_betas, g_trans, g_rots, None, None


class GeomFunction(torch.autograd.Function):
    """verts -> (joints, verts_rel, verts_view, verts_ndc, vnormals); see hfr_geom_forward."""

    @staticmethod
    def forward(ctx, topo: TopologyConsts, verts, root_out, root_xyz, focal, prp, want_normals):
        verts = _cu(verts)
        B, V = verts.shape[0], verts.shape[1]
        dev = verts.device
        root_xyz = None if root_xyz is None else _cu(root_xyz.reshape(B, 3))
        focal = None if focal is None else _cu(focal)
        prp = None if prp is None else _cu(prp)
        joints = torch.empty(B, max(topo.NOUT, 1), 3, device=dev, dtype=F32) if root_out >= 0 else None
        rel = torch.empty_like(verts)
        view = torch.empty_like(verts)
        ndc = torch.empty_like(verts) if focal is not None else None
        vn = torch.empty_like(verts) if want_normals else None
        geom_forward_raw(topo, verts, root_out, root_xyz, focal, prp, joints, rel, view, ndc, vn)
        ctx.topo, ctx.root_out = topo, root_out
        ctx.save_for_backward(verts, root_xyz, focal, prp)
        outs = (joints if joints is not None else verts.new_zeros(0), rel, view,
                ndc if ndc is not None else verts.new_zeros(0), vn if vn is not None else verts.new_zeros(0))
        return outs

    @staticmethod
    def backward(ctx, g_joints, g_rel, g_view, g_ndc, g_vn):
        verts, root_xyz, focal, prp = ctx.saved_tensors
        fix = lambda g, ok=True: None if (g is None or not ok or g.numel() == 0) else _cu(g)  # noqa: E731
        g_verts = torch.empty_like(verts)
        gs = (fix(g_joints, ctx.root_out >= 0), fix(g_rel), fix(g_view), fix(g_ndc, focal is not None), fix(g_vn))
        geom_backward_raw(ctx.topo, verts, ctx.root_out, root_xyz, focal, prp, *gs, g_verts)
        del gs                                                   # (kept alive until the launch was enqueued)
        return None, g_verts, None, None, None, None, None


class FaceVertsFunction(torch.autograd.Function):
    """(N,V,3) vertices -> packed (N*F,3,3) face vertices (`verts_packed()[faces_packed()]` of MeshRasterizer.forward);
    the backward is a fixed-order CSR gather instead of ATen's sort-based index_put."""

    @staticmethod
    def forward(ctx, topo: TopologyConsts, verts):
        if verts.dim() != 3 or verts.shape[1] != topo.V or verts.shape[2] != 3:
            raise ValueError(f"verts must be (N, {topo.V}, 3) for this topology, got {tuple(verts.shape)}")
        verts = _cu(verts)
        N = verts.shape[0]
        fv = torch.empty(N * topo.F, 3, 3, dtype=F32, device=verts.device)
        L.call("hfr_face_verts_forward", topo.struct, L.HfrFaceVertsArgs(N, L.ptr(verts, F32), L.ptr(fv, F32), None, None))
        ctx.topo, ctx.N = topo, N
        return fv

    @staticmethod
    def backward(ctx, g_fv):
        g_fv = _cu(g_fv)
        g_verts = torch.empty(ctx.N, ctx.topo.V, 3, dtype=F32, device=g_fv.device)
        L.call("hfr_face_verts_backward", ctx.topo.struct, L.HfrFaceVertsArgs(ctx.N, None, None, L.ptr(g_fv, F32), L.ptr(g_verts, F32)))
        return None, g_verts


class RasterizeFunction(torch.autograd.Function):
    """Same contract as pytorch3d._C.rasterize_meshes / rasterize_meshes_backward."""

    @staticmethod
    def forward(ctx, face_verts, mesh_first, mesh_nf, image_size, blur_radius, K, perspective_correct, clip_bary,
                cull_backfaces):
        face_verts = _cu(face_verts)
        H, W = image_size
        N = mesh_first.shape[0]
        dev = face_verts.device
        p2f = torch.empty(N, H, W, K, dtype=I64, device=dev)
        zbuf = torch.empty(N, H, W, K, dtype=F32, device=dev)
        bary = torch.empty(N, H, W, K, 3, dtype=F32, device=dev)
        dists = torch.empty(N, H, W, K, dtype=F32, device=dev)
        ws = raster_workspace(face_verts.shape[0], dev)
        a = raster_args(face_verts, mesh_first, mesh_nf, H, W, K, blur_radius, perspective_correct, clip_bary,
                        cull_backfaces, p2f, zbuf, bary, dists, ws)
        L.call("hfr_raster_forward", a)
        ctx.cfg = (H, W, K, float(blur_radius), int(perspective_correct), int(clip_bary))
        ctx.save_for_backward(face_verts, p2f)
        ctx.mark_non_differentiable(p2f)
        return p2f, zbuf, bary, dists

    @staticmethod
    def backward(ctx, _g_p2f, g_zbuf, g_bary, g_dists):
        face_verts, p2f = ctx.saved_tensors
        H, W, K, blur, pc, clip = ctx.cfg
        g_fv = torch.zeros_like(face_verts)
        fix = lambda g: None if g is None else _cu(g)  # noqa: E731
        gz, gb, gd = fix(g_zbuf), fix(g_bary), fix(g_dists)      # kept alive until the launch is enqueued
        a = L.HfrRasterBwdArgs(p2f.shape[0], H, W, K, face_verts.shape[0], L.ptr(face_verts, F32), L.ptr(p2f, I64),
                               L.ptr(gz, F32), L.ptr(gb, F32), L.ptr(gd, F32), blur, pc, clip, L.ptr(g_fv, F32))
        L.call("hfr_raster_backward", a)
        del gz, gb, gd
        return g_fv, None, None, None, None, None, None, None, None


class ShadeFunction(torch.autograd.Function):
    """Fragments + mesh attributes + texture + lights -> RGBA image (N,H,W,4)."""

    @staticmethod
    def forward(ctx, params, p2f, zbuf, bary, dists, faces, verts_view, vnormals, faces_uvs, verts_uvs, texture,
                light_dir, light_color, tex_basis=None, tex_params=None):
        N, H, W, K = p2f.shape
        cu = lambda t: None if t is None else _cu(t)  # noqa: E731
        zbuf, bary, dists = cu(zbuf), cu(bary), cu(dists)
        verts_view, vnormals, texture = cu(verts_view), cu(vnormals), cu(texture)
        light_dir, light_color, verts_uvs = cu(light_dir), cu(light_color), cu(verts_uvs)
        tex_basis, tex_params = cu(tex_basis), cu(tex_params)
        image = torch.empty(N, H, W, 4, dtype=F32, device=p2f.device)
        a = shade_fwd_args(params, (p2f, zbuf, bary, dists), faces, verts_view, vnormals, faces_uvs, verts_uvs,
                           texture, light_dir, light_color, image, None, tex_basis, tex_params)
        L.call("hfr_shade_forward", a)
        ctx.params = params
        ctx.pca = tex_basis is not None
        ctx.save_for_backward(p2f, zbuf, bary, dists, faces, verts_view, vnormals, faces_uvs, verts_uvs, texture,
                              light_dir, light_color, image, *([tex_basis, tex_params] if tex_basis is not None else []))
        return image

    @staticmethod
    def backward(ctx, g_image):
        (p2f, zbuf, bary, dists, faces, verts_view, vnormals, faces_uvs, verts_uvs, texture, light_dir,
         light_color, image) = ctx.saved_tensors[:13]
        tex_basis, tex_params = ctx.saved_tensors[13:] if ctx.pca else (None, None)
        p = ctx.params
        g_image = _cu(g_image)
        f = shade_fwd_args(p, (p2f, zbuf, bary, dists), faces, verts_view, vnormals, faces_uvs, verts_uvs, texture,
                           light_dir, light_color, image, None, tex_basis, tex_params)
        # only the gradients autograd asks for are computed: every NULL output drops its reductions from the kernel (a
        # frozen texture / mean map alone is 12 scattered atomics per shaded fragment)
        need = ctx.needs_input_grad
        g_tp = torch.zeros_like(tex_params) if (ctx.pca and need[14]) else None
        e = lambda t, i: torch.empty_like(t) if need[i] else None  # noqa: E731
        g_zbuf, g_bary, g_dists = e(zbuf, 2), e(bary, 3), e(dists, 4)
        z = lambda t, i: None if (t is None or not need[i]) else torch.zeros_like(t)  # noqa: E731
        g_vv, g_vn, g_tex, g_ld, g_lc = z(verts_view, 6), z(vnormals, 7), z(texture, 10), z(light_dir, 11), z(light_color, 12)
        a = L.HfrShadeBwdArgs(f, L.ptr(g_image, F32), L.ptr(g_zbuf, F32), L.ptr(g_bary, F32), L.ptr(g_dists, F32),
                              None, None, 0.0, 1, 0, L.ptr(g_vv, F32), L.ptr(g_vn, F32), L.ptr(g_tex, F32),
                              L.ptr(g_ld, F32), L.ptr(g_lc, F32), None, 0, 0, L.ptr(g_tp, F32))
        L.call("hfr_shade_backward", a)
        return None, None, g_zbuf, g_bary, g_dists, None, g_vv, g_vn, None, None, g_tex, g_ld, g_lc, None, g_tp


class PoolFunction(torch.autograd.Function):
    """(N,H*aa,W*aa,4) -> re_img (N,3,H,W), re_sil (N,1,H,W), maskRGBs; models_res_nimble.py:210-220."""

    @staticmethod
    def forward(ctx, image, aa, binarize, images_in):
        image = _cu(image)
        N, Hi, Wi, _ = image.shape
        H, W = Hi // aa, Wi // aa
        dev = image.device
        re_img = torch.empty(N, 3, H, W, dtype=F32, device=dev)
        re_sil = torch.empty(N, 1, H, W, dtype=F32, device=dev)
        images_in = None if images_in is None else _cu(images_in)
        mask = torch.empty(N, 3, H, W, dtype=F32, device=dev) if images_in is not None else None
        a = L.HfrPoolArgs(N, H, W, aa, int(binarize), L.ptr(image, F32), L.ptr(images_in, F32), L.ptr(re_img, F32),
                          L.ptr(re_sil, F32), L.ptr(mask, F32))
        L.call("hfr_pool_forward", a)
        ctx.cfg = (N, H, W, aa, int(binarize))
        if mask is None:
            mask = image.new_zeros(0)
        ctx.mark_non_differentiable(mask)
        return re_img, re_sil, mask

    @staticmethod
    def backward(ctx, g_img, g_sil, _g_mask):
        N, H, W, aa, binarize = ctx.cfg
        ref = g_img if g_img is not None else g_sil
        g_image = torch.empty(N, H * aa, W * aa, 4, dtype=F32, device=ref.device)
        fix = lambda g: None if g is None else _cu(g)  # noqa: E731
        gi, gs = fix(g_img), fix(g_sil)                          # kept alive until the launch is enqueued
        a = L.HfrPoolBwdArgs(N, H, W, aa, binarize, L.ptr(gi, F32), L.ptr(gs, F32), L.ptr(g_image, F32))
        L.call("hfr_pool_backward", a)
        del gi, gs
        return g_image, None, None, None


class RenderLossFunction(torch.autograd.Function):
    """Five render-dependent loss terms (unweighted) from one pass; returns a (5,) tensor
    [texture, mrgb, ssim_tex, sil, iou] following losses.py:355-378, 399-408."""

    @staticmethod
    def forward(ctx, re_img, re_sil, imgs, seg, sil_scale, want_ssim):
        # grad mode is off inside Function.forward and _cu() may copy (non-contiguous / non-fp32 input), so whether a
        # gradient will be asked for is read from autograd's own bookkeeping, not from the (possibly copied) tensors
        need_grad = bool(ctx.needs_input_grad[0] or ctx.needs_input_grad[1])
        re_img, re_sil, imgs, seg = _cu(re_img), _cu(re_sil), _cu(imgs), _cu(seg)
        N, _, H, W = re_img.shape
        dev = re_img.device
        sums = torch.zeros(L.LOSS_NSUMS + 2 * N, dtype=F32, device=dev)
        dmaps = torch.empty(N, 9, H, W, dtype=F32, device=dev) if (want_ssim and need_grad) else None
        gauss = gauss_taps(dev)
        flags = torch.zeros(N, (H + 3) // 4, (W + 3) // 4, dtype=torch.uint8, device=dev)
        a = L.HfrLossArgs(N, H, W, float(sil_scale), int(want_ssim), int(need_grad), 0, L.ptr(re_img, F32),
                          L.ptr(re_sil, F32), L.ptr(imgs, F32), L.ptr(seg, F32), L.ptr(sums, F32), L.ptr(gauss, F32),
                          L.ptr(dmaps, F32), L.ptr(flags))
        L.call("hfr_loss_forward", a)
        cnt = float(N * 3 * H * W)
        tex = sums[0] / cnt
        mrgb = (sums[2] / cnt - sums[1] / cnt) ** 2
        ssim = 1 - sums[4] / cnt
        sil = sums[3] / float(N * H * W)
        mul, add = sums[L.LOSS_NSUMS:L.LOSS_NSUMS + N], sums[L.LOSS_NSUMS + N:]
        iou = 1 - (mul / (add - mul)).mean()
        ctx.cfg = (N, H, W, float(sil_scale), int(want_ssim))
        ctx.save_for_backward(re_img, re_sil, imgs, seg, sums, dmaps if dmaps is not None else sums.new_zeros(0), flags)
        return torch.stack([tex, mrgb, ssim, sil, iou])

    @staticmethod
    def backward(ctx, g):
        re_img, re_sil, imgs, seg, sums, dmaps, flags = ctx.saved_tensors
        N, H, W, sil_scale, want_ssim = ctx.cfg
        dmaps = dmaps if dmaps.numel() else None
        if want_ssim and dmaps is None:
            raise L.HfrError("RenderLossFunction.backward: the forward kept no SSIM derivative maps (it saw no input that "
                             "required grad); the ssim_tex gradient would be dropped")
        gauss = gauss_taps(re_img.device)
        f = L.HfrLossArgs(N, H, W, sil_scale, want_ssim, 1, 0, L.ptr(re_img, F32), L.ptr(re_sil, F32), L.ptr(imgs, F32),
                          L.ptr(seg, F32), L.ptr(sums, F32), L.ptr(gauss, F32), L.ptr(dmaps, F32), L.ptr(flags))
        g_img, g_sil = torch.empty_like(re_img), torch.empty_like(re_sil)
        w = _cu(g)
        a = L.HfrLossBwdArgs(f, L.ptr(w, F32), L.ptr(gauss, F32), N * 3 * H * W, N, L.ptr(g_img, F32), L.ptr(g_sil, F32))
        L.call("hfr_loss_backward", a)
        return g_img, g_sil, None, None, None, None


class SelfRenderLossFunction(torch.autograd.Function):
    """The self-supervised photometric terms of losses.py:317-340 (unweighted) from one pass: returns (3,)
    [texture_self, mrgb_self, ssim_tex_self] for re_img (N,3,H,W) against maskRGBs (N,3,H,W) with the per-sample
    confidences texture_con (N,).  Gradients reach re_img only (maskRGBs is images * (re_sil > 0), models_res_nimble.py:220)."""

    @staticmethod
    def forward(ctx, re_img, mask_rgbs, texture_con):
        need_grad = bool(ctx.needs_input_grad[0])
        re_img, mask_rgbs, con = _cu(re_img), _cu(mask_rgbs), _cu(texture_con.reshape(-1))
        N, _, H, W = re_img.shape
        dev = re_img.device
        sums = torch.zeros(L.LOSS_NSUMS + 3 * N, dtype=F32, device=dev)
        dmaps = torch.empty(N, 9, H, W, dtype=F32, device=dev) if need_grad else None
        gauss = gauss_taps(dev)
        a = L.HfrLossArgs(N, H, W, 1.0, 1, int(need_grad), 0, L.ptr(re_img, F32), None, L.ptr(mask_rgbs, F32), None,
                          L.ptr(sums, F32), L.ptr(gauss, F32), L.ptr(dmaps, F32), None, 3)
        L.call("hfr_loss_forward", a)
        per = float(3 * H * W)
        c2 = con * con
        norm = c2.sum().reshape(1)
        ns = L.LOSS_NSUMS
        tex = (sums[ns:ns + N] * c2).sum() / (per * norm[0])
        mrgb = ((sums[ns + N:ns + 2 * N] / per - sums[ns + 2 * N:ns + 3 * N] / per).abs() * c2).sum() / norm[0]
        ssim = 1 - sums[L.LOSS_SSIM] / (N * per)
        ctx.cfg = (N, H, W)
        ctx.save_for_backward(re_img, mask_rgbs, con, norm, sums, dmaps if dmaps is not None else sums.new_zeros(0))
        return torch.stack([tex, mrgb, ssim])

    @staticmethod
    def backward(ctx, g):
        re_img, mask_rgbs, con, norm, sums, dmaps = ctx.saved_tensors
        N, H, W = ctx.cfg
        if not dmaps.numel():
            raise L.HfrError("SelfRenderLossFunction.backward: the forward kept no SSIM derivative maps")
        gauss = gauss_taps(re_img.device)
        f = L.HfrLossArgs(N, H, W, 1.0, 1, 1, 0, L.ptr(re_img, F32), None, L.ptr(mask_rgbs, F32), None, L.ptr(sums, F32),
                          L.ptr(gauss, F32), L.ptr(dmaps, F32), None, 3)
        g_img = torch.empty_like(re_img)
        w = torch.zeros(5, dtype=F32, device=re_img.device)
        w[:3] = g
        a = L.HfrLossBwdArgs(f, L.ptr(w, F32), L.ptr(gauss, F32), N * 3 * H * W, N, L.ptr(g_img, F32), None,
                             L.ptr(con, F32), L.ptr(norm, F32))
        L.call("hfr_loss_backward", a)
        return g_img, None, None


# ------------------------------------------------------------------------------------------------
# keypoints + mesh regularisers
# ------------------------------------------------------------------------------------------------
BONE_CHILD = list(range(1, 21))                                       # utils/losses_util.py:226-245: bone i ends at joint i+1
BONE_PARENT = [0 if (c - 1) % 4 == 0 else c - 1 for c in BONE_CHILD]  # ... and starts at the wrist or the previous joint
_BONES = {}


def bone_tables(device):
    key = str(device)
    if key not in _BONES:
        _BONES[key] = (torch.tensor(BONE_PARENT, dtype=I32, device=device), torch.tensor(BONE_CHILD, dtype=I32, device=device))
    return _BONES[key]


def keypoint_args(joints, root_xyz, Ks, verts, faces, joints_gt, j2d_gt, verts_gt, conf, l2, j2d, sums, mscale=True,
                  nbr=None):
    B, NJ = joints.shape[0], joints.shape[1]
    bp, bc = bone_tables(joints.device)
    nb = bp.shape[0] if NJ == 21 else 0          # the bone tables are the 21-joint FreiHAND skeleton's
    V = verts.shape[1] if verts is not None else 0
    F = faces.shape[0] if faces is not None else 0
    return L.HfrKeypointArgs(B, NJ, V, F, int(l2), nb, 9 if (mscale and NJ > 10) else -1, 10, 0.0282,
                             L.ptr(joints, F32, "joints"), L.ptr(root_xyz, F32, "root_xyz"), L.ptr(Ks, F32, "Ks"),
                             L.ptr(verts, F32, "verts"), L.ptr(faces, I32, "faces"), L.ptr(joints_gt, F32, "joints_gt"),
                             L.ptr(j2d_gt, F32, "j2d_gt"), L.ptr(verts_gt, F32, "verts_gt"), L.ptr(conf, F32, "conf"),
                             L.ptr(bp, I32), L.ptr(bc, I32), L.ptr(j2d, F32), L.ptr(sums, F32),
                             L.ptr(nbr[0], I32) if nbr is not None else None, L.ptr(nbr[1], I32) if nbr is not None else None)


class KeypointLossFunction(torch.autograd.Function):
    """(joints, verts) -> (terms (8,) [joint_2d, joint_3d, vert_3d, bone_direc, bone_direc_3d, edge_length, mscale,
    triangle] unweighted, j2d (B,NJ,2)).  A term whose ground truth is None is 0; j2d is empty when Ks is None;
    `nbr` = (nbr_ptr, nbr_idx) of TopologyConsts switches the uniform Laplacian term on."""

    @staticmethod
    def forward(ctx, joints, verts, root_xyz, Ks, joints_gt, j2d_gt, verts_gt, conf, faces, l2, nbr=None):
        c = lambda t: None if t is None else _cu(t)  # noqa: E731
        joints, verts, joints_gt, j2d_gt, verts_gt = c(joints), c(verts), c(joints_gt), c(j2d_gt), c(verts_gt)
        B, NJ = joints.shape[0], joints.shape[1]
        dev = joints.device
        root_xyz = None if root_xyz is None else _cu(root_xyz.reshape(B, 3))
        Ks = None if Ks is None else _cu(Ks[:, :3, :3])
        conf = None if conf is None else _cu(conf.reshape(B, NJ))
        faces = None if faces is None else faces.to(I32).contiguous()
        j2d = torch.empty(B, NJ, 2, dtype=F32, device=dev) if Ks is not None else None
        sums = torch.zeros(L.KP_NSUMS, dtype=F32, device=dev)
        a = keypoint_args(joints, root_xyz, Ks, verts, faces, joints_gt, j2d_gt, verts_gt, conf, l2, j2d, sums, nbr=nbr)
        L.call("hfr_keypoint_forward", a)
        V, F, NB = a.V, a.F, a.NB
        cnt = torch.tensor([B * NJ * 2, B * NJ * 3, max(B * V * 3, 1), max(B * NB, 1), max(B * NB, 1), max(B * F * 3, 1), B,
                            max(B * V, 1)], dtype=F32, device=dev)
        ctx.l2 = int(l2)
        ctx.nbr = nbr
        ctx.save_for_backward(joints, verts, root_xyz, Ks, joints_gt, j2d_gt, verts_gt, conf, faces)
        return sums[:8] / cnt, (j2d if j2d is not None else joints.new_zeros(0))

    @staticmethod
    def backward(ctx, g_terms, g_j2d):
        joints, verts, root_xyz, Ks, joints_gt, j2d_gt, verts_gt, conf, faces = ctx.saved_tensors
        B = joints.shape[0]
        dev = joints.device
        w = torch.zeros(L.KP_NSUMS, dtype=F32, device=dev)
        w[:8] = g_terms
        g_joints = torch.empty_like(joints)
        g_verts = torch.empty_like(verts) if verts is not None else None
        g2 = None if (g_j2d is None or g_j2d.numel() == 0 or Ks is None) else _cu(g_j2d)
        f = keypoint_args(joints, root_xyz, Ks, verts, faces, joints_gt, j2d_gt, verts_gt, conf, ctx.l2, None, w, nbr=ctx.nbr)
        a = L.HfrKeypointBwdArgs(f, L.ptr(w, F32), B, L.ptr(g2, F32), L.ptr(g_joints, F32), L.ptr(g_verts, F32))
        L.call("hfr_keypoint_backward", a)
        return g_joints, g_verts, None, None, None, None, None, None, None, None, None
