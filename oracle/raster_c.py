"""TEST INFRASTRUCTURE — build + ctypes wrapper for oracle/raster_naive.c (CPU oracle / CPU baseline)."""
import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libraster_naive.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "raster_naive.c")
    if not force and os.path.isfile(_SO) and os.path.getmtime(_SO) >= os.path.getmtime(src):
        return _SO
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", src,
                           "-o", _SO, "-lm"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.hfr_oracle_rasterize_naive.restype = ctypes.c_int
    return _lib


def rasterize_naive(face_verts, mesh_first, mesh_nf, image_size, blur_radius=0.0, faces_per_pixel=1,
                    perspective_correct=True, clip_barycentric_coords=None, cull_backfaces=False, threads=1):
    """Same contract as oracle.p3d.rasterize_meshes (forward only, fp32, CPU)."""
    lib = _load()
    H, W = (image_size, image_size) if isinstance(image_size, int) else image_size
    K = faces_per_pixel
    if clip_barycentric_coords is None:
        clip_barycentric_coords = blur_radius > 0
    fv = np.ascontiguousarray(face_verts.detach().cpu().numpy(), dtype=np.float32)
    mf = np.ascontiguousarray(np.asarray(mesh_first, dtype=np.int64))
    nf = np.ascontiguousarray(np.asarray(mesh_nf, dtype=np.int64))
    N = len(nf)
    p2f = np.empty((N, H, W, K), np.int64)
    zb = np.empty((N, H, W, K), np.float32)
    ba = np.empty((N, H, W, K, 3), np.float32)
    di = np.empty((N, H, W, K), np.float32)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    rc = lib.hfr_oracle_rasterize_naive(P(fv), P(mf), P(nf), N, H, W, K, ctypes.c_float(blur_radius),
                                        int(perspective_correct), int(clip_barycentric_coords),
                                        int(cull_backfaces), int(threads), P(p2f), P(zb), P(ba), P(di))
    if rc != 0:
        raise RuntimeError(f"oracle rasterizer failed rc={rc}")
    return (torch.from_numpy(p2f), torch.from_numpy(zb), torch.from_numpy(ba), torch.from_numpy(di))
