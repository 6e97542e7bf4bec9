"""ctypes binding of libhifihr_b200.so (the C-ABI declared in include/hifihr_b200.h).

This is the whole "extension": tensors are unpacked to raw device pointers and the
current CUDA stream; there is no CPU fallback — if the library is missing, or a
tensor is not a contiguous CUDA tensor of the expected dtype, the call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HFR_B200_LIB") or os.path.join(_HERE, "libhifihr_b200.so")   # env: tuning builds only

vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float


class HfrHandModel(C.Structure):
    _fields_ = [("V", i32), ("NJ", i32), ("NS", i32), ("NPC", i32), ("NW", i32), ("NT", i32),
                ("center_joint", i32), ("C3", i32),
                ("dirs", vp), ("v_template", vp), ("J_template", vp), ("J_shapedirs", vp),
                ("pca_comps", vp), ("pose_mean", vp), ("parents", vp), ("skin_idx", vp), ("skin_w", vp),
                ("tip_verts", vp), ("joint_order", vp), ("palm_verts", i32 * 2),
                ("jv_ptr", vp), ("jv_vert", vp), ("jv_w", vp), ("basis_packed", vp)]


class HfrManoFwdArgs(C.Structure):
    _fields_ = [("B", i32), ("pose", vp), ("betas", vp), ("trans", vp), ("verts", vp), ("joints", vp),
                ("rots", vp), ("n_rot_in", i32), ("pose_off", i32), ("root_palm", i32), ("workspace", vp)]


class HfrManoBwdArgs(C.Structure):
    _fields_ = [("B", i32), ("pose", vp), ("betas", vp), ("trans", vp), ("g_verts", vp), ("g_joints", vp),
                ("g_pose", vp), ("g_betas", vp), ("g_trans", vp),
                ("rots", vp), ("n_rot_in", i32), ("pose_off", i32), ("root_palm", i32), ("g_rots", vp),
                ("workspace", vp), ("reuse_forward", i32)]


class HfrTopology(C.Structure):
    _fields_ = [("V", i32), ("F", i32), ("faces", vp), ("vf_ptr", vp), ("vf_idx", vp),
                ("NJR", i32), ("NOUT", i32), ("jr_ptr", vp), ("jr_col", vp), ("jr_val", vp),
                ("vj_ptr", vp), ("vj_row", vp), ("vj_val", vp), ("out_src", vp), ("vf_nbr", vp)]


class HfrGeomFwdArgs(C.Structure):
    _fields_ = [("B", i32), ("root_out", i32), ("verts", vp), ("root_xyz", vp), ("focal", vp), ("prp", vp),
                ("joints", vp), ("verts_rel", vp), ("verts_view", vp), ("verts_ndc", vp), ("vnormals", vp),
                ("face_verts", vp)]


class HfrGeomBwdArgs(C.Structure):
    _fields_ = [("B", i32), ("root_out", i32), ("verts", vp), ("root_xyz", vp), ("focal", vp), ("prp", vp),
                ("g_joints", vp), ("g_verts_rel", vp), ("g_verts_view", vp), ("g_verts_ndc", vp),
                ("g_vnormals", vp), ("g_verts", vp), ("face_rec", vp), ("raster_ws", vp), ("status", vp), ("rec_partial", vp)]


class HfrFaceVertsArgs(C.Structure):
    _fields_ = [("B", i32), ("verts", vp), ("face_verts", vp), ("g_face_verts", vp), ("g_verts", vp)]


class HfrRasterArgs(C.Structure):
    _fields_ = [("N", i32), ("H", i32), ("W", i32), ("K", i32), ("Ftot", i64), ("face_verts", vp),
                ("mesh_first", vp), ("mesh_nfaces", vp), ("blur_radius", f32),
                ("perspective_correct", i32), ("clip_barycentric", i32), ("cull_backfaces", i32),
                ("pix_to_face", vp), ("zbuf", vp), ("bary", vp), ("dists", vp), ("workspace", vp), ("tile_queue", vp)]


class HfrRasterBwdArgs(C.Structure):
    _fields_ = [("N", i32), ("H", i32), ("W", i32), ("K", i32), ("Ftot", i64), ("face_verts", vp),
                ("pix_to_face", vp), ("g_zbuf", vp), ("g_bary", vp), ("g_dists", vp), ("blur_radius", f32),
                ("perspective_correct", i32), ("clip_barycentric", i32), ("g_face_verts", vp)]


class HfrShadeParams(C.Structure):
    _fields_ = [("N", i32), ("H", i32), ("W", i32), ("K", i32), ("F", i32), ("V", i32), ("blend", i32),
                ("shade", i32), ("sigma", f32), ("gamma", f32), ("znear", f32), ("zfar", f32),
                ("background", f32 * 3), ("light_ambient", f32 * 3), ("light_specular", f32 * 3),
                ("mat_ambient", f32 * 3), ("mat_diffuse", f32 * 3), ("mat_specular", f32 * 3),
                ("shininess", f32), ("tex_n", i32), ("tex_h", i32), ("tex_w", i32), ("VT", i32), ("tex_pca", i32), ("light_point", i32),
                ("tex_basis_stride", i32)]


class HfrShadeFwdArgs(C.Structure):
    _fields_ = [("p", HfrShadeParams), ("pix_to_face", vp), ("zbuf", vp), ("bary", vp), ("dists", vp),
                ("faces", vp), ("verts_view", vp), ("vnormals", vp), ("faces_uvs", vp), ("verts_uvs", vp),
                ("texture", vp), ("light_dir", vp), ("light_color", vp), ("image", vp), ("face_attr", vp), ("tex_basis", vp), ("tex_params", vp)]


class HfrFaceAttrArgs(C.Structure):
    _fields_ = [("N", i32), ("F", i32), ("V", i32), ("VT", i32), ("faces", vp), ("verts_view", vp), ("vnormals", vp),
                ("faces_uvs", vp), ("verts_uvs", vp), ("face_attr", vp)]


class HfrShadeBwdArgs(C.Structure):
    _fields_ = [("f", HfrShadeFwdArgs), ("g_image", vp), ("g_zbuf", vp), ("g_bary", vp), ("g_dists", vp),
                ("verts_ndc", vp), ("g_verts_ndc", vp), ("blur_radius", f32), ("perspective_correct", i32),
                ("clip_barycentric", i32), ("g_verts_view", vp), ("g_vnormals", vp), ("g_texture", vp),
                ("g_light_dir", vp), ("g_light_color", vp), ("tile_box", vp), ("pool_aa", i32), ("pool_binarize", i32), ("g_tex_params", vp)]


class HfrShadeBwdTiledArgs(C.Structure):
    _fields_ = [("f", HfrShadeFwdArgs), ("g_image", vp), ("verts_ndc", vp), ("focal", vp), ("blur_radius", f32),
                ("perspective_correct", i32), ("clip_barycentric", i32), ("raster_ws", vp), ("face_rec", vp),
                ("rec_cap", i64), ("light_acc", vp), ("tex_acc", vp), ("g_texture", vp), ("fx_scale", vp), ("status", vp),
                ("pool_aa", i32), ("pool_binarize", i32), ("gmax_bits", vp), ("fix_sums", vp), ("fix_w", vp),
                ("fix_image", vp), ("fix_inv_scale", f32), ("fix_count", i64), ("tile_queue", vp)]


class HfrGradFinishArgs(C.Structure):
    _fields_ = [("tex_acc", vp), ("g_texture", vp), ("n_tex", i64), ("light_acc", vp), ("g_light_dir", vp),
                ("g_light_color", vp), ("N", i32), ("fx_scale", vp), ("gmax_bits", vp)]


class HfrRasterShadeArgs(C.Structure):
    _fields_ = [("r", HfrRasterArgs), ("s", HfrShadeFwdArgs)]


class HfrRasterShadePoolArgs(C.Structure):
    _fields_ = [("r", HfrRasterArgs), ("s", HfrShadeFwdArgs), ("aa", i32), ("binarize", i32), ("images_in", vp),
                ("pooled", vp), ("re_img", vp), ("re_sil", vp), ("mask_rgbs", vp)]


class HfrPoolArgs(C.Structure):
    _fields_ = [("N", i32), ("H", i32), ("W", i32), ("aa", i32), ("binarize", i32), ("image", vp),
                ("images_in", vp), ("re_img", vp), ("re_sil", vp), ("mask_rgbs", vp)]


class HfrPoolBwdArgs(C.Structure):
    _fields_ = [("N", i32), ("H", i32), ("W", i32), ("aa", i32), ("binarize", i32), ("g_re_img", vp),
                ("g_re_sil", vp), ("g_image", vp)]


class HfrLossArgs(C.Structure):
    _fields_ = [("N", i32), ("H", i32), ("W", i32), ("sil_scale", f32), ("want_ssim", i32), ("want_grad", i32), ("nhwc", i32),
                ("re_img", vp), ("re_sil", vp), ("imgs", vp), ("seg", vp), ("sums", vp), ("gauss", vp), ("dmaps", vp),
                ("tile_flags", vp), ("mask_mode", i32), ("imgs_u8", vp), ("seg_u8", vp), ("partials", vp), ("ticket", vp),
                ("dmaps_box", vp), ("dmaps_box_aa", i32)]


class HfrLossBwdArgs(C.Structure):
    _fields_ = [("f", HfrLossArgs), ("w", vp), ("gauss", vp), ("count_global", i64), ("n_global", i32), ("g_re_img", vp), ("g_re_sil", vp),
                ("tex_con", vp), ("self_norm", vp), ("tile_box", vp), ("box_aa", i32), ("skip_mrgb", i32), ("gmax_bits", vp)]


class HfrKeypointArgs(C.Structure):
    _fields_ = [("B", i32), ("NJ", i32), ("V", i32), ("F", i32), ("l2", i32), ("NB", i32), ("scale_a", i32), ("scale_b", i32),
                ("scale_len", f32), ("joints", vp), ("root_xyz", vp), ("Kmat", vp), ("verts", vp), ("faces", vp),
                ("joints_gt", vp), ("j2d_gt", vp), ("verts_gt", vp), ("conf", vp), ("bone_parent", vp), ("bone_child", vp),
                ("j2d", vp), ("sums", vp), ("nbr_ptr", vp), ("nbr_idx", vp)]


class HfrKeypointBwdArgs(C.Structure):
    _fields_ = [("f", HfrKeypointArgs), ("w", vp), ("n_global", i32), ("g_j2d_in", vp), ("g_joints", vp), ("g_verts", vp)]


ABI_VERSION = 5
LOSS_NSUMS = 8
FACE_ATTR_FLOATS = 28
FACE_REC_FLOATS = 18
LOSS_L2 = 5
LOSS_SSIM = 4
KP_NSUMS = 8
KP_TERMS = ("joint_2d", "joint_3d", "vert_3d", "bone_direc", "bone_direc_3d", "edge_length", "mscale", "triangle")
ENTRY_POINTS = [
    "hfr_last_error", "hfr_abi_version", "hfr_device_ok", "hfr_mano_forward", "hfr_mano_backward",
    "hfr_geom_forward", "hfr_geom_backward", "hfr_raster_workspace_bytes", "hfr_raster_forward",
    "hfr_raster_backward", "hfr_raster_tile_box", "hfr_shade_forward", "hfr_shade_backward", "hfr_raster_shade_forward",
    "hfr_raster_shade_pool_forward", "hfr_face_attr_forward", "hfr_pool_forward", "hfr_pool_backward", "hfr_loss_forward", "hfr_loss_backward",
    "hfr_keypoint_forward", "hfr_keypoint_backward", "hfr_shade_backward_tiled", "hfr_grad_finish",
    "hfr_loss_partials_floats", "hfr_mano_packed_basis_bytes", "hfr_mano_pack_basis", "hfr_mano_workspace_bytes",
    "hfr_mano_batched_status", "hfr_raster_queue_bytes", "hfr_geom_rec_partial_floats",
    "hfr_face_verts_forward", "hfr_face_verts_backward",
]

_lib = None


class HfrError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load the CUDA library; raise loudly if it was not built (no fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise HfrError(f"{LIB_PATH} is missing: run `python -m hifihr_b200.build` (or "
                           "__graft_entry__.build()); hifihr_b200 has no CPU or PyTorch fallback")
        _lib = C.CDLL(LIB_PATH)
        _lib.hfr_last_error.restype = C.c_char_p
        _lib.hfr_raster_workspace_bytes.restype = C.c_int64
        _lib.hfr_raster_workspace_bytes.argtypes = [C.c_int64]
        _lib.hfr_raster_tile_box.restype = C.c_void_p
        _lib.hfr_raster_tile_box.argtypes = [C.c_void_p, C.c_int64, C.c_int32]
        _lib.hfr_loss_partials_floats.restype = C.c_int64
        _lib.hfr_loss_partials_floats.argtypes = [C.c_int32, C.c_int32, C.c_int32]
        _lib.hfr_geom_rec_partial_floats.restype = C.c_int64
        _lib.hfr_geom_rec_partial_floats.argtypes = [C.c_void_p, C.c_int32]
        _lib.hfr_raster_queue_bytes.restype = C.c_int64
        _lib.hfr_raster_queue_bytes.argtypes = [C.c_int32, C.c_int32, C.c_int32]
        _lib.hfr_mano_packed_basis_bytes.restype = C.c_int64
        _lib.hfr_mano_packed_basis_bytes.argtypes = [C.c_void_p]
        _lib.hfr_mano_workspace_bytes.restype = C.c_int64
        _lib.hfr_mano_workspace_bytes.argtypes = [C.c_void_p, C.c_int32]
        if _lib.hfr_abi_version() != ABI_VERSION:
            raise HfrError("libhifihr_b200.so ABI version mismatch")
    return _lib


_tls = threading.local()


def _pending_devices():
    if not hasattr(_tls, "devs"):
        _tls.devs = set()
    return _tls.devs


def call(name: str, *structs, device=None):
    """Invoke an entry point on the current torch CUDA stream OF THE DEVICE THAT OWNS THE TENSORS (the reference's
    nn.DataParallel calls forward from one thread per device, train_hrnet.py:560; the launchers themselves never call
    cudaSetDevice).  The device is `device` when given, else the one every tensor unpacked by ptr() since the last
    call lives on (mixing devices raises), else the current device.  Error codes map to exceptions: HFR_EINVAL ->
    ValueError as PyTorch3D raises for bad settings, others -> RuntimeError."""
    L = lib()
    devs = _pending_devices()
    if device is not None:
        dev = torch.device(device).index
        dev = torch.cuda.current_device() if dev is None else dev
    elif len(devs) > 1:
        found = sorted(devs)
        devs.clear()
        raise HfrError(f"{name}: tensors live on different CUDA devices {found}")
    else:
        dev = next(iter(devs)) if devs else torch.cuda.current_device()
    devs.clear()
    cargs = [C.byref(s) if isinstance(s, C.Structure) else s for s in structs]   # structs by address, scalars / pointers by value
    if dev == torch.cuda.current_device():
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        rc = getattr(L, name)(*cargs, stream)
    else:
        with torch.cuda.device(dev):
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            rc = getattr(L, name)(*cargs, stream)
    if rc != 0:
        msg = L.hfr_last_error().decode()
        if rc == 1:
            raise ValueError(msg)
        raise HfrError(f"{name} failed (code {rc}): {msg}")


def ptr(t: torch.Tensor | None, dtype=None, name="tensor"):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise HfrError(f"{name} must be a CUDA tensor: hifihr_b200 has no CPU path")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    _pending_devices().add(t.device.index)
    return t.data_ptr()


def f3(v):
    return (f32 * 3)(*[float(x) for x in v])
