"""Meshes-lite: the part of pytorch3d.structures.Meshes the hot path touches.

Surface used by the reference: constructor (utils/my_mano.py:44, utils/losses_util.py:360),
`_num_verts_per_mesh` and `offset_verts_` (models_res_nimble.py:203-205), `verts_padded()`
and `__getitem__` (utils/visualize_util.py:25,29); the renderer needs the packed views,
normals and textures.  All meshes of a batch share one topology (true for MANO / NIMBLE);
`faces` may be (F,3) or the reference's repeated (N,F,3).
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops

_TOPO_CACHE: dict = {}


def topology_for(faces: torch.Tensor, V: int, device) -> ops.TopologyConsts:
    """Incidence tables for a faces tensor, cached by content hash (built once per topology)."""
    f = faces.detach()
    if f.dim() == 3:
        f = f[0]
    fn = f.to("cpu", torch.int64).numpy()
    key = (hash(fn.tobytes()), int(V), str(device))
    if key not in _TOPO_CACHE:
        _TOPO_CACHE[key] = ops.TopologyConsts(fn, V, device=device)
    return _TOPO_CACHE[key]


class Meshes:
    def __init__(self, verts, faces, textures=None, topology: ops.TopologyConsts | None = None):
        if isinstance(verts, (list, tuple)):
            verts = torch.stack(list(verts))
        if isinstance(faces, (list, tuple)):
            faces = torch.stack(list(faces))
        if verts.dim() != 3 or verts.shape[-1] != 3:
            raise ValueError("verts must be (N, V, 3)")
        self._verts = verts
        self._faces = faces
        self.textures = textures
        self.device = verts.device
        N, V = verts.shape[0], verts.shape[1]
        self._N, self._V = N, V
        self._topo = topology if topology is not None else topology_for(faces, V, verts.device)
        self._F = self._topo.F
        self._num_verts_per_mesh = torch.full((N,), V, dtype=torch.int64, device=verts.device)
        self._normals = None

    # -- reference-facing surface ------------------------------------------------------------
    def __len__(self):
        return self._N

    def __getitem__(self, idx):
        if isinstance(idx, int):
            idx = [idx]
        return Meshes(self._verts[idx], self._topo.faces_long, self.textures, topology=self._topo)

    def verts_padded(self):
        return self._verts

    def faces_padded(self):
        return self._topo.faces_long[None].expand(self._N, -1, -1)

    def verts_packed(self):
        return self._verts.reshape(-1, 3)

    def faces_packed(self):
        off = (torch.arange(self._N, device=self.device, dtype=torch.int64) * self._V).view(-1, 1, 1)
        return (self._topo.faces_long[None] + off).reshape(-1, 3)

    def num_faces_per_mesh(self):
        return torch.full((self._N,), self._F, dtype=torch.int64, device=self.device)

    def num_verts_per_mesh(self):
        return self._num_verts_per_mesh

    def mesh_to_faces_packed_first_idx(self):
        return torch.arange(self._N, dtype=torch.int64, device=self.device) * self._F

    def offset_verts_(self, vert_offsets_packed):
        """In-place semantic of the reference call (models_res_nimble.py:204-205); autograd-safe."""
        off = vert_offsets_packed
        if off.dim() == 2 and off.shape[0] == self._N * self._V:
            off = off.reshape(self._N, self._V, 3)
        self._verts = self._verts + off
        self._normals = None
        return self

    def offset_verts(self, vert_offsets_packed):
        return Meshes(self._verts, self._faces, self.textures, topology=self._topo).offset_verts_(vert_offsets_packed)

    def update_padded(self, new_verts_padded):
        return Meshes(new_verts_padded, self._faces, self.textures, topology=self._topo)

    def verts_normals_padded(self):
        if self._normals is None:
            outs = ops.GeomFunction.apply(self._topo, self._verts, -1, None, None, None, True)
            self._normals = outs[4]
        return self._normals

    def verts_normals_packed(self):
        return self.verts_normals_padded().reshape(-1, 3)

    def sample_textures(self, fragments):
        if self.textures is None:
            raise ValueError("Meshes does not have textures")
        return self.textures.sample_textures(fragments, self)

    def to(self, device):
        return Meshes(self._verts.to(device), self._faces.to(device), self.textures)

    @property
    def topology(self):
        return self._topo
