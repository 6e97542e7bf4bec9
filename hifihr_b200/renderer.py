"""Drop-in renderer objects with the PyTorch3D names / signatures the reference uses
(imports at models_res_nimble.py:12-21; construction :74-96; call :183-208):

  RasterizationSettings, MeshRasterizer, Fragments, MeshRenderer, HardPhongShader,
  SoftPhongShader, SoftSilhouetteShader, Materials, DirectionalLights, PointLights,
  PerspectiveCameras, BlendParams, TexturesUV.

Everything that touches pixels runs in the sm_100a kernels of csrc/raster.cu and
csrc/shade.cu; semantics follow SURVEY.md Appendix A.  Assumptions the hot path
guarantees and this layer enforces: cameras have R=I, T=0 (models_res_nimble.py:184-186),
all meshes of a batch share one topology, lights are directional.
"""
from __future__ import annotations

import math
from typing import NamedTuple, Optional, Sequence, Union

import torch
from torch import nn

from . import _lib as L
from . import ops
from .structures import Meshes

F32, I32, I64 = torch.float32, torch.int32, torch.int64


class Fragments(NamedTuple):
    pix_to_face: torch.Tensor
    zbuf: torch.Tensor
    bary_coords: torch.Tensor
    dists: torch.Tensor


class RasterizationSettings:
    def __init__(self, image_size: Union[int, Sequence[int]] = 256, blur_radius: float = 0.0, faces_per_pixel: int = 1,
                 bin_size: Optional[int] = None, max_faces_per_bin: Optional[int] = None,
                 perspective_correct: Optional[bool] = None, clip_barycentric_coords: Optional[bool] = None,
                 cull_backfaces: bool = False, z_clip_value: Optional[float] = None, cull_to_frustum: bool = False):
        self.image_size = image_size
        self.blur_radius = blur_radius
        self.faces_per_pixel = faces_per_pixel
        self.bin_size = bin_size                      # accepted for API parity; the tile size is fixed (16x16)
        self.max_faces_per_bin = max_faces_per_bin    # accepted; tile lists never overflow (batched staging)
        self.perspective_correct = perspective_correct
        self.clip_barycentric_coords = clip_barycentric_coords
        self.cull_backfaces = cull_backfaces
        self.z_clip_value = z_clip_value
        self.cull_to_frustum = cull_to_frustum


class BlendParams(NamedTuple):
    sigma: float = 1e-4
    gamma: float = 1e-4
    background_color: Sequence[float] = (1.0, 1.0, 1.0)


def _rows(x, n, device, name):
    t = torch.as_tensor(x, dtype=F32, device=device)
    if t.dim() == 1:
        t = t[None]
    if t.shape[0] == 1 and n > 1:
        t = t.expand(n, -1)
    if t.shape[0] != n:
        raise ValueError(f"{name}: batch {t.shape[0]} does not match {n}")
    return t


class PerspectiveCameras:
    """NDC perspective camera, R=I and T=0 only (what the reference builds at models_res_nimble.py:184)."""

    def __init__(self, focal_length=1.0, principal_point=((0.0, 0.0),), R=None, T=None, K=None, device="cpu",
                 in_ndc: bool = True, image_size=None):
        if R is not None or T is not None or K is not None or not in_ndc:
            raise NotImplementedError("hifihr_b200 cameras: only in_ndc=True with R=I, T=0 (the reference's setup)")
        self.device = torch.device(device) if not isinstance(focal_length, torch.Tensor) else focal_length.device
        fl = torch.as_tensor(focal_length, dtype=F32, device=self.device)
        if fl.dim() == 0:
            fl = fl.view(1, 1).expand(1, 2)
        elif fl.dim() == 1:
            fl = fl[:, None].expand(-1, 2) if fl.shape[0] != 2 else fl[None]
        self.focal_length = fl
        self.principal_point = torch.as_tensor(principal_point, dtype=F32, device=self.device).reshape(-1, 2)

    def is_perspective(self):
        return True

    def get_camera_center(self):
        return torch.zeros(self.focal_length.shape[0], 3, device=self.device)

    def get_znear(self):
        return None

    def to(self, device):
        self.device = torch.device(device)
        self.focal_length = self.focal_length.to(device)
        self.principal_point = self.principal_point.to(device)
        return self


class DirectionalLights:
    def __init__(self, ambient_color=((0.5, 0.5, 0.5),), diffuse_color=((0.3, 0.3, 0.3),),
                 specular_color=((0.2, 0.2, 0.2),), direction=((0, 1, 0),), device="cpu"):
        self.device = torch.device(device)
        t = lambda x: x if isinstance(x, torch.Tensor) else torch.tensor(x, dtype=F32, device=self.device)  # noqa: E731
        self.ambient_color, self.diffuse_color = t(ambient_color), t(diffuse_color)
        self.specular_color, self.direction = t(specular_color), t(direction)


class PointLights:
    """pytorch3d.renderer.lighting.PointLights as constructed at models_res_nimble.py:191-198 (ifLight=False):
    the light direction of a shaded point is location - point (diffuse / specular of renderer/lighting.py)."""

    def __init__(self, ambient_color=((0.5, 0.5, 0.5),), diffuse_color=((0.3, 0.3, 0.3),),
                 specular_color=((0.2, 0.2, 0.2),), location=((0, 1, 0),), device="cpu"):
        self.device = torch.device(device)
        t = lambda x: x if isinstance(x, torch.Tensor) else torch.tensor(x, dtype=F32, device=self.device)  # noqa: E731
        self.ambient_color, self.diffuse_color = t(ambient_color), t(diffuse_color)
        self.specular_color, self.location = t(specular_color), t(location)


class Materials:
    def __init__(self, ambient_color=((1, 1, 1),), diffuse_color=((1, 1, 1),), specular_color=((1, 1, 1),),
                 shininess=64, device="cpu"):
        self.device = torch.device(device)
        self.ambient_color, self.diffuse_color, self.specular_color = ambient_color, diffuse_color, specular_color
        self.shininess = shininess


def _c3(x):
    """One RGB triple shared by the batch -> python floats (kernel parameter)."""
    t = torch.as_tensor(x, dtype=F32).reshape(-1, 3)
    if t.shape[0] != 1:
        raise NotImplementedError("per-sample ambient/specular/material colours are not on the hot path")
    return [float(v) for v in t[0]]


class TexturesUV:
    """maps (N or 1,Ht,Wt,3), faces_uvs (F,3) or (N,F,3), verts_uvs (Vt,2) or (N,Vt,2) — shared uv layout."""

    def __init__(self, maps, faces_uvs, verts_uvs, padding_mode="border", align_corners=True, sampling_mode="bilinear"):
        if padding_mode != "border" or not align_corners or sampling_mode != "bilinear":
            raise NotImplementedError("TexturesUV: only the PyTorch3D defaults (bilinear, border, align_corners=True)")
        self._maps = maps
        fu = faces_uvs[0] if faces_uvs.dim() == 3 else faces_uvs
        vu = verts_uvs[0] if verts_uvs.dim() == 3 else verts_uvs
        self._faces_uvs = fu.to(I32).contiguous()
        self._verts_uvs = vu.to(F32).contiguous()

    def maps_padded(self):
        return self._maps

    def sample_textures(self, fragments, meshes=None):
        raise NotImplementedError("texel sampling is fused into the shader kernels (hfr_shade_forward)")


class TexturesUVPCA(TexturesUV):
    """UV texture given as a PCA model: per-sample map = mean + params @ basis (the NIMBLE texture model, SURVEY.md
    §8f row 4).  The shader kernels evaluate the model at the four bilinear taps of every fragment, so the
    (N, T, T, 3) per-sample maps are never materialised; gradients flow to `params` (and to `mean`).
    mean (1,T,T,3) or (T,T,3); basis (n_comp,T,T,3); params (N,n_comp)."""

    def __init__(self, mean, basis, params, faces_uvs, verts_uvs):
        mean = mean if mean.dim() == 4 else mean[None]
        super().__init__(mean, faces_uvs, verts_uvs)
        if basis.shape[1:] != mean.shape[1:] or params.shape[1] != basis.shape[0]:
            raise ValueError("TexturesUVPCA: basis must be (n_comp,T,T,3) matching the mean map, params (N,n_comp)")
        self._basis, self._params = basis, params

    def maps_padded(self):
        """The materialised per-sample maps (library GEMM) - for visualisation / export only."""
        n = self._params.shape[1]
        return (self._maps.reshape(1, -1) + self._params @ self._basis.reshape(n, -1)).view(-1, *self._maps.shape[1:])


# ------------------------------------------------------------------------------------------------
class MeshRasterizer(nn.Module):
    def __init__(self, cameras=None, raster_settings=None):
        super().__init__()
        self.cameras = cameras
        self.raster_settings = raster_settings if raster_settings is not None else RasterizationSettings()

    def to(self, device):
        if self.cameras is not None:
            self.cameras = self.cameras.to(device)
        return self

    def transform(self, meshes_world: Meshes, **kwargs):
        """world -> NDC with view-space z (MeshRasterizer.transform); returns (verts_view, verts_ndc)."""
        cameras = kwargs.get("cameras", self.cameras)
        if cameras is None:
            raise ValueError("Cameras must be specified either at initialization or in the forward pass of MeshRasterizer")
        verts = meshes_world.verts_padded()
        N = verts.shape[0]
        focal = _rows(cameras.focal_length, N, verts.device, "focal_length")
        prp = _rows(cameras.principal_point, N, verts.device, "principal_point")
        outs = ops.GeomFunction.apply(meshes_world.topology, verts, -1, None, focal, prp, False)
        return outs[2], outs[3]

    def forward(self, meshes_world: Meshes, **kwargs) -> Fragments:
        rs = kwargs.get("raster_settings", self.raster_settings)
        _, verts_ndc = self.transform(meshes_world, **kwargs)
        N, V = verts_ndc.shape[0], verts_ndc.shape[1]
        if rs.z_clip_value is not None or rs.cull_to_frustum:
            raise NotImplementedError("z_clip_value / cull_to_frustum: the reference never sets them")
        H, W = (rs.image_size, rs.image_size) if isinstance(rs.image_size, int) else tuple(rs.image_size)
        K = rs.faces_per_pixel
        if K < 1 or K > 16:
            raise ValueError(f"faces_per_pixel must be in [1, 16], got {K}")
        cameras = kwargs.get("cameras", self.cameras)
        pc = rs.perspective_correct if rs.perspective_correct is not None else cameras.is_perspective()
        clip = rs.clip_barycentric_coords if rs.clip_barycentric_coords is not None else rs.blur_radius > 0.0
        face_verts = ops.FaceVertsFunction.apply(meshes_world.topology, verts_ndc)   # verts_packed()[faces_packed()]
        p2f, zbuf, bary, dists = ops.RasterizeFunction.apply(
            face_verts, meshes_world.mesh_to_faces_packed_first_idx(), meshes_world.num_faces_per_mesh(), (H, W),
            rs.blur_radius, K, pc, clip, rs.cull_backfaces)
        return Fragments(p2f, zbuf, bary, dists)


def rasterize_meshes(meshes_or_face_verts, image_size=256, blur_radius=0.0, faces_per_pixel=8, bin_size=None,
                     max_faces_per_bin=None, perspective_correct=False, clip_barycentric_coords=False,
                     cull_backfaces=False, z_clip_value=None, cull_to_frustum=False, mesh_to_face_first_idx=None,
                     num_faces_per_mesh=None):
    """Functional form (pytorch3d.renderer.mesh.rasterize_meshes): a Meshes already in NDC, or packed face_verts."""
    if isinstance(meshes_or_face_verts, Meshes):
        m = meshes_or_face_verts
        fv = ops.FaceVertsFunction.apply(m.topology, m.verts_padded())
        first, nf = m.mesh_to_faces_packed_first_idx(), m.num_faces_per_mesh()
    else:
        fv, first, nf = meshes_or_face_verts, mesh_to_face_first_idx, num_faces_per_mesh
    H, W = (image_size, image_size) if isinstance(image_size, int) else tuple(image_size)
    return ops.RasterizeFunction.apply(fv, first, nf, (H, W), blur_radius, faces_per_pixel, perspective_correct,
                                       clip_barycentric_coords, cull_backfaces)


class _ShaderBase(nn.Module):
    blend = L.HFR_BLEND_HARD if hasattr(L, "HFR_BLEND_HARD") else 0
    shade = 1

    def __init__(self, device="cpu", cameras=None, lights=None, materials=None, blend_params=None):
        super().__init__()
        self.lights = lights if lights is not None else DirectionalLights(device=device)
        self.materials = materials if materials is not None else Materials(device=device)
        self.cameras = cameras
        self.blend_params = blend_params if blend_params is not None else BlendParams()

    def to(self, device):
        return self

    def _render(self, fragments: Fragments, meshes: Meshes, **kwargs):
        lights = kwargs.get("lights", self.lights)
        materials = kwargs.get("materials", self.materials)
        bp = kwargs.get("blend_params", self.blend_params)
        p2f = fragments.pix_to_face
        N, H, W, K = p2f.shape
        dev = p2f.device
        topo = meshes.topology
        if self.shade == 1:
            point = isinstance(lights, PointLights)
            tex = meshes.textures
            if tex is None:
                raise ValueError("Meshes does not have textures")
            pca = isinstance(tex, TexturesUVPCA)
            maps = tex._maps if pca else tex.maps_padded()
            basis = ops.pack_tex_basis(tex._basis) if pca else None     # texel-major copy, built once per basis
            params = ops.shade_params(N, H, W, K, topo.F, topo.V, self.blend, 1, bp.sigma, bp.gamma,
                                      bp.background_color, _c3(lights.ambient_color), _c3(lights.specular_color),
                                      _c3(materials.ambient_color), _c3(materials.diffuse_color),
                                      _c3(materials.specular_color), materials.shininess,
                                      tex_shape=maps.shape[:3], VT=tex._verts_uvs.shape[0],
                                      tex_pca=tex._basis.shape[0] if pca else 0, light_point=int(point),
                                      tex_basis_stride=basis.shape[1] if pca else 0)
            ldir = _rows(lights.location, N, dev, "lights.location") if point else _rows(lights.direction, N, dev, "lights.direction")
            lcol = _rows(lights.diffuse_color, N, dev, "lights.diffuse_color")
            return ops.ShadeFunction.apply(params, p2f, fragments.zbuf, fragments.bary_coords, fragments.dists,
                                           topo.faces, meshes.verts_padded(), meshes.verts_normals_padded(),
                                           tex._faces_uvs.to(dev), tex._verts_uvs.to(dev), maps, ldir, lcol,
                                           basis, tex._params if pca else None)
        params = ops.shade_params(N, H, W, K, topo.F, topo.V, self.blend, 0, bp.sigma, bp.gamma, bp.background_color,
                                  (0, 0, 0), (0, 0, 0), (0, 0, 0), (0, 0, 0), (0, 0, 0), 1.0)
        return ops.ShadeFunction.apply(params, p2f, fragments.zbuf, fragments.bary_coords, fragments.dists,
                                       None, None, None, None, None, None, None, None)

    def forward(self, fragments: Fragments, meshes: Meshes, **kwargs):
        return self._render(fragments, meshes, **kwargs)


class HardPhongShader(_ShaderBase):
    blend, shade = 0, 1


class SoftPhongShader(_ShaderBase):
    blend, shade = 2, 1


class SoftSilhouetteShader(_ShaderBase):
    blend, shade = 1, 0

    def __init__(self, blend_params=None):
        super().__init__(blend_params=blend_params)


class MeshRenderer(nn.Module):
    def __init__(self, rasterizer, shader):
        super().__init__()
        self.rasterizer = rasterizer
        self.shader = shader

    def to(self, device):
        self.rasterizer.to(device)
        self.shader.to(device)
        return self

    def forward(self, meshes_world: Meshes, **kwargs):
        fragments = self.rasterizer(meshes_world, **kwargs)
        return self.shader(fragments, meshes_world, **kwargs)
