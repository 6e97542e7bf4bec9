"""The hot path of Model.forward (models_res_nimble.py:133-223) without the CNN encoders.

HandRenderModel   — modular drop-in built from the reference-named objects (hand layer,
                    xyz_from_vertice, root shift, PerspectiveCameras, MeshRenderer, avg-pool,
                    output dict), differentiable through torch autograd bridges.
FusedHandStep     — the same computation as ONE forward+backward sequence of raw kernel
                    launches on preallocated buffers (10 launches per step, CUDA-graph
                    capturable): MANO -> geometry -> rasterize+shade -> loss | loss' ->
                    shade'+rasterize' -> geometry' -> MANO'.  This is what bench.py times.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
from torch import nn

from . import _lib as L
from . import ops
from .mano import MyMANOLayer, xyz_from_vertice
from .renderer import (BlendParams, DirectionalLights, HardPhongShader, Materials, MeshRasterizer, MeshRenderer,
                       PerspectiveCameras, PointLights, RasterizationSettings, SoftPhongShader, TexturesUV)

F32, I32, I64 = torch.float32, torch.int32, torch.int64


def get_ndc_fx_fy_cx_cy(Ks):
    """models_res_nimble.py:228-235 (the 224/112 constants are the reference's)."""
    ndc_fx = Ks[:, 0, 0] * 2 / 224.0
    ndc_fy = Ks[:, 1, 1] * 2 / 224.0
    ndc_px = -(Ks[:, 0, 2] - 112.0) * 2 / 224.0
    ndc_py = -(Ks[:, 1, 2] - 112.0) * 2 / 224.0
    return torch.stack([ndc_fx, ndc_fy], dim=-1), torch.stack([ndc_px, ndc_py], dim=-1)


def mano_synthetic_uvs(v_template, faces):
    """MANO_RIGHT.pkl has no UVs: planar map of the template's x/z extent (§8d)."""
    xz = np.asarray(v_template, np.float64)[:, [0, 2]]
    uv = (xz - xz.min(0)) / (xz.max(0) - xz.min(0))
    return np.ascontiguousarray(uv, dtype=np.float32), np.ascontiguousarray(faces, dtype=np.int64)


class HandRenderModel(nn.Module):
    """hand layer -> joints -> root shift -> camera -> render -> pool -> outputs (reference order)."""

    def __init__(self, ifRender=True, device="cuda", hand_model="mano", root_id=9, image_size=224, aa_factor=3,
                 blur_radius=0.0, faces_per_pixel=1, soft=False, binarize=True, texture_size=512, mano_root=None,
                 blend_params=None, ifLight=True):
        super().__init__()
        self.ifLight = ifLight      # models_res_nimble.py:187-198: False renders with a default PointLights
        if hand_model != "mano":
            raise NotImplementedError("use hifihr_b200.nimble.MyNIMBLELayer for the NIMBLE-shaped stand-in")
        self.root_id = root_id
        self.ifRender = ifRender
        self.aa_factor = aa_factor
        self.binarize = binarize
        self.ncomps = [10, 48, None]
        self.hand_layer = MyMANOLayer(ifRender, device, shape_ncomp=10, pose_ncomp=48, tex_ncomp=None,
                                      mano_root=mano_root)
        d = self.hand_layer.mano_layer._mano
        uv, fuv = mano_synthetic_uvs(d["v_template"], d["f"])
        self.register_buffer("verts_uvs", torch.tensor(uv))
        self.register_buffer("faces_uvs", torch.tensor(fuv))
        g = torch.Generator().manual_seed(20231)
        self.texture = nn.Parameter(torch.rand(1, texture_size, texture_size, 3, generator=g))
        self.mano_face = self.hand_layer.mesh_face
        if ifRender:
            rs = RasterizationSettings(image_size=image_size * aa_factor, blur_radius=blur_radius,
                                       faces_per_pixel=faces_per_pixel)
            materials = Materials(diffuse_color=((0.8, 0.8, 0.8),), specular_color=((0.2, 0.2, 0.2),), shininess=30,
                                  device=device)
            shader_cls = SoftPhongShader if soft else HardPhongShader
            self.renderer_p3d = MeshRenderer(rasterizer=MeshRasterizer(raster_settings=rs),
                                             shader=shader_cls(materials=materials, device=device,
                                                               blend_params=blend_params))

    def forward(self, hand_params, light_params=None, Ks=None, root_xyz=None, images=None):
        outputs = self.hand_layer(hand_params, handle_collision=False)
        outputs.update(hand_params)
        verts = outputs["mano_verts"]
        dev = verts.device
        topo = self.hand_layer.topology(dev)
        joints, verts_rel = xyz_from_vertice(topo, verts, root_id=self.root_id)
        outputs["joints"] = joints
        outputs["mano_verts"] = verts_rel
        B = verts.shape[0]
        if self.ifRender:
            fcl, prp = get_ndc_fx_fy_cx_cy(Ks)
            cameras = PerspectiveCameras(focal_length=-fcl, principal_point=prp, device=dev)
            if self.ifLight:
                lighting = DirectionalLights(diffuse_color=light_params["colors"], direction=light_params["directions"],
                                             device=dev)
            else:
                lighting = PointLights(device=dev)
            meshes = outputs["skin_meshes"]
            pred_root = (verts - verts_rel)[:, :1]            # = joints[:, root_id] before the shift
            verts_num = verts.shape[1]                        # = meshes._num_verts_per_mesh[0], without a host sync
            meshes.offset_verts_(-pred_root.repeat(1, verts_num, 1).view(verts_num * B, 3))
            meshes.offset_verts_(root_xyz.reshape(B, 1, 3).repeat(1, verts_num, 1).view(verts_num * B, 3))
            meshes.textures = TexturesUV(self.texture, self.faces_uvs, self.verts_uvs)
            rendered = self.renderer_p3d(meshes, cameras=cameras, lights=lighting)
            re_img, re_sil, mask = ops.PoolFunction.apply(rendered, self.aa_factor, self.binarize, images)
            outputs["re_img"], outputs["re_sil"] = re_img, re_sil
            if images is not None:
                outputs["maskRGBs"] = mask
        outputs["mano_faces"] = self.mano_face.to(dev).repeat(B, 1, 1)
        return outputs


class FusedHandStep:
    """Forward + backward of the whole hot path as raw launches on preallocated buffers.

    Inputs (device, fp32, contiguous): pose (B,48), betas (B,10), focal (B,2), prp (B,2), root_xyz (B,3),
    light_dir (B,3), light_color (B,3), imgs (B,3,S,S), seg (B,S,S).  imgs / seg may also be uint8 (the datasets'
    8-bit images and {0,1} masks): the loss kernels then apply ToTensor's x / 255 while loading.  The shared texture (1,T,T,3) and the
    loss weights live in the object.  After `step()`: `sums` holds the loss partial sums, and g_pose, g_betas,
    g_texture, g_light_dir, g_light_color the gradients of  sum_k lambda_k * term_k.
    """

    def __init__(self, B, image_size=224, faces_per_pixel=4, blur_radius=None, sigma=1e-4, gamma=1e-4, soft=True,
                 texture_size=512, lambdas=None, device="cuda", mano_root=None, n_global=None, sil_scale=1.0,
                 aa_factor=1, binarize=False, want_nchw=False, face_records=False, tiled_backward=True,
                 deterministic=True, rec_per_face=12, batched_hand=True, tile_queue=True):
        """aa_factor > 1 selects the SSAA-fused render (the reference's own setting is image_size=224,
        aa_factor=3, faces_per_pixel=1, soft=False, binarize=True, sil_scale=255; models_res_nimble.py:74-96,
        208-220): Fragments are rasterised at image_size*aa_factor, the pooled RGBA (B,S,S,4) is the only image
        that exists, and the backward folds avg_pool2d' into the shading backward.  want_nchw also writes
        re_img / re_sil / maskRGBs in the reference's NCHW layout.  face_records packs one contiguous attribute
        record per (sample, face) each step (one more launch) so the shaders skip two dependent gathers; measured
        neutral on B200 at C2 (backward 370 -> 363 us, forward + 9 us for the packing launch), hence off by default.
        tiled_backward selects the atomics-free backward (hfr_shade_backward_tiled: per-tile face sort, one record per
        (face, tile), fixed-order gather in the geometry backward); with deterministic=True the texture gradient goes
        through 64-bit fixed-point accumulators as the light gradients always do, so a step is bit-reproducible.
        rec_per_face sizes the record store (records = rec_per_face * B * F; a face needs one per 16x16 tile its
        dilated bounding box touches; overflow raises through the status word / NaN gradients)."""
        dev = torch.device(device)
        self.B, self.S, self.K, self.dev = B, image_size, faces_per_pixel, dev
        self.aa, self.binarize = int(aa_factor), bool(binarize)
        self.Sr = image_size * self.aa          # rasterised resolution
        self.soft = soft
        self.blur = (np.log(1.0 / 1e-4 - 1.0) * sigma if soft else 0.0) if blur_radius is None else blur_radius
        self._build_model(dev, mano_root, texture_size)
        lam = dict(texture=1.0, mrgb=1.0, ssim_tex=1.0, sil=1.0, iou=0.0)
        lam.update(lambdas or {})
        self.lambdas = lam
        self.w = torch.tensor([lam["texture"], lam["mrgb"], lam["ssim_tex"], lam["sil"], lam["iou"]], device=dev)
        self.n_global = n_global or B
        self.sil_scale = sil_scale
        V, Fm, S, K, Sr = self.hm.V, self.topo.F, self.S, self.K, self.Sr
        e = lambda *s, dt=F32: torch.empty(*s, dtype=dt, device=dev)  # noqa: E731
        self.verts, self.joints = e(B, V, 3), e(B, max(self.topo.NOUT, 1), 3)
        self.verts_rel, self.verts_view, self.verts_ndc, self.vnormals = e(B, V, 3), e(B, V, 3), e(B, V, 3), e(B, V, 3)
        self.face_verts = e(B * Fm, 3, 3)
        self.face_attr = e(B, Fm, L.FACE_ATTR_FLOATS) if face_records else None
        self.p2f = e(B, Sr, Sr, K, dt=I64)
        self.zbuf, self.bary, self.dists = e(B, Sr, Sr, K), e(B, Sr, Sr, K, 3), e(B, Sr, Sr, K)
        # pooled RGBA when aa > 1.  g_image starts as zeros: the loss backward only writes the tiles that can hold fragments
        self.image, self.g_image = e(B, S, S, 4), torch.zeros(B, S, S, 4, dtype=F32, device=dev)
        self.re_img = self.re_sil = self.mask_rgbs = None
        if want_nchw:
            self.re_img, self.re_sil, self.mask_rgbs = e(B, 3, S, S), e(B, 1, S, S), e(B, 3, S, S)
        self.dmaps = e(B, 9, S, S)
        self.tile_flags = torch.zeros(B, (S + 3) // 4, (S + 3) // 4, dtype=torch.uint8, device=dev)
        # step outputs (loss partial sums, per-sample pose / shape gradients) live in ONE flat buffer so a single
        # device->host copy returns them, and in two alternating sets (flip_outputs) so that copy can overlap the
        # next step on another stream
        self._n_sums = L.LOSS_NSUMS + 2 * B
        self._n_pose, self._n_shape = self.hm.pose_dim, self.hm.NS          # 48 + 10 for MANO
        self._outs = [torch.zeros(self._n_sums + (self._n_pose + self._n_shape + self.n_tex) * B, dtype=F32, device=dev)
                      for _ in range(2)]
        self._out_set = 0
        self._bind_outputs()
        self.ws = ops.raster_workspace(B * Fm, dev)
        self.tile_queue = ops.raster_tile_queue(B, Sr, Sr, dev) if (tile_queue and self.aa == 1) else None
        self.mesh_first = (torch.arange(B, device=dev, dtype=I64) * Fm).contiguous()
        self.mesh_nf = torch.full((B,), Fm, device=dev, dtype=I64)
        # every accumulated gradient lives in one flat buffer so a single memset clears them
        n_tex_acc = self.texture.numel() if self.texture_grad else 0
        n_acc = 3 * B * V * 3 + n_tex_acc + 6 * B
        self.acc = torch.zeros(n_acc, dtype=F32, device=dev)
        o = 0
        def take(n, shape):
            nonlocal o
            t = self.acc[o:o + n].view(*shape)
            o += n
            return t
        self.g_ndc, self.g_view, self.g_vn = take(B * V * 3, (B, V, 3)), take(B * V * 3, (B, V, 3)), take(B * V * 3, (B, V, 3))
        self.g_texture = take(n_tex_acc, self.texture.shape) if self.texture_grad else None
        self.g_light_dir, self.g_light_color = take(3 * B, (B, 3)), take(3 * B, (B, 3))
        self.g_verts = e(B, V, 3)
        self.mano_ws = self.hm.workspace(B) if batched_hand else None   # batched tensor-core hand layer (None: per-sample kernels)
        self.tiled, self.deterministic = bool(tiled_backward), bool(deterministic)
        if self.tiled:
            self.rec_cap = int(rec_per_face) * B * Fm + 4096
            self.face_rec = e(self.rec_cap, L.FACE_REC_FLOATS)
            self.rec_partial = e(int(L.lib().hfr_geom_rec_partial_floats(C.byref(self.topo.struct), B)))
            self.light_acc = torch.zeros(B, 6, dtype=I64, device=dev)
            self.tex_acc = torch.zeros(self.texture.shape, dtype=I64, device=dev) if self.deterministic else None
            self.status = torch.zeros(1, dtype=torch.int32, device=dev)
            self.fx_scale = torch.ones(1, dtype=F32, device=dev)
            self.gmax_bits = torch.zeros(1, dtype=torch.int32, device=dev)
            self._focal = None
        # deterministic loss sums: per-CTA partials + a ticket word (the last CTA adds them in a fixed order)
        self.loss_partials = e(int(L.lib().hfr_loss_partials_floats(B, S, S)) + 8) if self.deterministic else None
        self.loss_ticket = torch.zeros(1, dtype=torch.int32, device=dev) if self.deterministic else None
        self.gauss = ops.gauss_taps(dev)
        self.params = ops.shade_params(B, Sr, Sr, K, Fm, V, 2 if soft else 0, 1, sigma, gamma, (1.0, 1.0, 1.0),
                                       (0.5, 0.5, 0.5), (0.2, 0.2, 0.2), (1.0, 1.0, 1.0), (0.8, 0.8, 0.8),
                                       (0.2, 0.2, 0.2), 30.0, tex_shape=self.texture.shape[:3], VT=self.verts_uvs.shape[0],
                                       tex_pca=self.n_tex, tex_basis_stride=self.tex_basis.shape[1] if self.n_tex else 0)
        # kernels of OURS per step(): mano, geom, [face records], raster setup, raster+shade(+pool), loss | loss', shade'+raster',
        # geom', mano'  (the two torch memsets of the accumulators are not counted)
        # tiled backward: + raster scan, record clear, gradient finish, record gather; batched hand layer: 3 + 3 launches
        # instead of 1 + 1; tile queue: + the ordering pass  (C2 default: 18 kernels per step, as the ncu launch list shows)
        self.launches_per_step = (9 + (1 if face_records else 0) + (4 if self.tiled else 0) + (4 if self.mano_ws is not None else 0)
                                  + (1 if self.tile_queue is not None else 0) + (1 if self.n_tex else 0))
        if self.tiled and not self.deterministic:
            self.g_light_dir.zero_()

    def _build_model(self, dev, mano_root, texture_size):
        """The hand model behind the step: MANO (utils/my_mano.py), its synthetic UV layout and one shared (1,T,T,3) map.
        Subclasses swap the model (FusedNimbleStep); root_out is the output joint the mesh is centred on (-1: none)."""
        self.layer = MyMANOLayer(True, dev, shape_ncomp=10, pose_ncomp=48, tex_ncomp=None, mano_root=mano_root)
        self.hm = self.layer.mano_layer.consts(dev)
        self.topo = self.layer.topology(dev)
        d = self.layer.mano_layer._mano
        uv, fuv = mano_synthetic_uvs(d["v_template"], d["f"])
        self.verts_uvs = torch.tensor(uv, device=dev)
        self.faces_uvs = torch.tensor(fuv, device=dev, dtype=I32)
        g = torch.Generator().manual_seed(20231)
        self.texture = torch.rand(1, texture_size, texture_size, 3, generator=g).to(dev)
        self.root_out = 9
        self.tex_basis, self.n_tex = None, 0       # texture PCA model (texel-major basis, components): FusedNimbleStep
        self.texture_grad = True                   # False: the map is frozen (no gradient buffer, no reductions)

    def _bind_outputs(self):
        o, B, ns = self._outs[self._out_set], self.B, self._n_sums
        self.out = o
        self.sums = o[:ns]
        npz, nsh = self._n_pose, self._n_shape
        self.g_pose = o[ns:ns + npz * B].view(B, npz)
        self.g_betas = o[ns + npz * B:ns + (npz + nsh) * B].view(B, nsh)
        self.g_tex_params = o[ns + (npz + nsh) * B:].view(B, self.n_tex) if self.n_tex else None

    def flip_outputs(self):
        """Switch to the other output set: the next step() writes there, the set just produced stays intact."""
        self._out_set ^= 1
        self._bind_outputs()

    # ---------------------------------------------------------------------------------------
    def forward(self, pose, betas, focal, prp, root_xyz, light_dir, light_color, imgs, seg, tex_params=None):
        B, S, K, Sr = self.B, self.S, self.K, self.Sr
        if (tex_params is not None) != bool(self.n_tex):
            raise ValueError("tex_params: texture PCA coefficients are given exactly when the step has a texture model")
        self._tex_params = tex_params
        # the cached argument structs hold raw pointers: keep the inputs alive until the backward has been enqueued
        self._inputs = (pose, betas, focal, prp, root_xyz, light_dir, light_color, imgs, seg, tex_params)
        self._focal = focal
        ops.mano_forward_raw(self.hm, pose, betas, None, self.verts, None, workspace=self.mano_ws)
        ops.geom_forward_raw(self.topo, self.verts, self.root_out, root_xyz, focal, prp, self.joints, self.verts_rel,
                             self.verts_view, self.verts_ndc, self.vnormals, self.face_verts)
        self.launch_raster_shade(light_dir, light_color, imgs)
        if not self.deterministic:
            self.sums.zero_()
        # targets may arrive as the dataset's 8-bit images / masks: the loss kernels convert while loading
        u8i, u8s = imgs.dtype == torch.uint8, seg.dtype == torch.uint8
        self._loss_args = L.HfrLossArgs(B, S, S, self.sil_scale, 1, 1, 1, L.ptr(self.image), None,
                                        None if u8i else L.ptr(imgs, F32, "imgs"), None if u8s else L.ptr(seg, F32, "seg"),
                                        L.ptr(self.sums), L.ptr(self.gauss), L.ptr(self.dmaps), L.ptr(self.tile_flags), 0,
                                        L.ptr(imgs, torch.uint8, "imgs") if u8i else None,
                                        L.ptr(seg, torch.uint8, "seg") if u8s else None,
                                        L.ptr(self.loss_partials), L.ptr(self.loss_ticket),
                                        ops.raster_tile_box(self.ws, B * self.topo.F, B), self.aa)
        L.call("hfr_loss_forward", self._loss_args)

    def launch_raster_shade(self, light_dir, light_color, imgs=None):
        """setup + rasterize + shade (+ SSAA pool and output split when aa_factor > 1) - two launches."""
        K, Sr = self.K, self.Sr
        if self.face_attr is not None:
            ops.face_attr_forward(self.topo.faces, self.verts_view, self.vnormals, self.faces_uvs, self.verts_uvs,
                                  self.face_attr)
        r = ops.raster_args(self.face_verts, self.mesh_first, self.mesh_nf, Sr, Sr, K, self.blur, True, self.blur > 0,
                            False, self.p2f, self.zbuf, self.bary, self.dists, self.ws, self.tile_queue)
        s = ops.shade_fwd_args(self.params, (self.p2f, self.zbuf, self.bary, self.dists), self.topo.faces,
                               self.verts_view, self.vnormals, self.faces_uvs, self.verts_uvs, self.texture,
                               light_dir, light_color, self.image if self.aa == 1 else None, self.face_attr,
                               self.tex_basis, self._tex_params if self.n_tex else None)
        if self.n_tex:
            # texture PCA model: the texel evaluation is a template parameter of the standalone shader only (DESIGN.md
            # 8), so rasterize and shade are two launches here
            L.call("hfr_raster_forward", r)
            L.call("hfr_shade_forward", s)
        elif self.aa == 1:
            L.call("hfr_raster_shade_forward", L.HfrRasterShadeArgs(r, s))
        else:
            masks = self.re_img is not None and imgs is not None and imgs.dtype == F32   # maskRGBs needs float images
            L.call("hfr_raster_shade_pool_forward",
                   L.HfrRasterShadePoolArgs(r, s, self.aa, int(self.binarize), L.ptr(imgs, F32) if masks else None,
                                            L.ptr(self.image), L.ptr(self.re_img), L.ptr(self.re_sil),
                                            L.ptr(self.mask_rgbs) if masks else None))
        self._shade_args = s

    def launch_shade_backward(self):
        """shade' + blend' + rasterize' (+ avg_pool2d' when aa_factor > 1) in one launch; accumulates into self.acc."""
        B = self.B
        if self.tiled:
            # the fixed-point multiplier of the 64-bit accumulators is derived on the device from max |g_image| (gmax_bits,
            # left by the loss backward); the mean-RGB term's gradient is added here from the (all-reduced) sums
            t = L.HfrShadeBwdTiledArgs(self._shade_args, L.ptr(self.g_image), L.ptr(self.verts_ndc), L.ptr(self._focal, F32, "focal"),
                                       float(self.blur), 1, int(self.blur > 0), L.ptr(self.ws), L.ptr(self.face_rec), self.rec_cap,
                                       L.ptr(self.light_acc), L.ptr(self.tex_acc), None if self.deterministic else L.ptr(self.g_texture),
                                       L.ptr(self.fx_scale), L.ptr(self.status), self.aa if self.aa > 1 else 0, int(self.binarize),
                                       L.ptr(self.gmax_bits), L.ptr(self.sums), L.ptr(self.w), L.ptr(self.image),
                                       1.0 / float(self.sil_scale), self.n_global * 3 * self.S * self.S, L.ptr(self.tile_queue))
            L.call("hfr_shade_backward_tiled", t)
            L.call("hfr_grad_finish", L.HfrGradFinishArgs(L.ptr(self.tex_acc), L.ptr(self.g_texture) if self.deterministic else None,
                                                          self.texture.numel() if self.deterministic else 0, L.ptr(self.light_acc),
                                                          L.ptr(self.g_light_dir), L.ptr(self.g_light_color), B, L.ptr(self.fx_scale),
                                                          L.ptr(self.gmax_bits)))
            return
        sb = L.HfrShadeBwdArgs(self._shade_args, L.ptr(self.g_image), None, None, None, L.ptr(self.verts_ndc),
                               L.ptr(self.g_ndc), float(self.blur), 1, int(self.blur > 0), L.ptr(self.g_view),
                               L.ptr(self.g_vn), L.ptr(self.g_texture), L.ptr(self.g_light_dir),
                               L.ptr(self.g_light_color), ops.raster_tile_box(self.ws, B * self.topo.F, B),
                               self.aa if self.aa > 1 else 0, int(self.binarize), L.ptr(self.g_tex_params))
        L.call("hfr_shade_backward", sb)

    def launch_geom_backward(self, focal, prp, root_xyz):
        if self.tiled:   # gathers d/d(view), d/d(normal) from the (face, tile) records in a fixed order
            ops.geom_backward_raw(self.topo, self.verts, self.root_out, root_xyz, focal, prp, None, None, None, None, None, self.g_verts,
                                  face_rec=self.face_rec, raster_ws=self.ws, status=self.status, rec_partial=self.rec_partial)
        else:
            ops.geom_backward_raw(self.topo, self.verts, self.root_out, root_xyz, focal, prp, None, None, self.g_view, self.g_ndc,
                                  self.g_vn, self.g_verts)

    def launch_loss_backward(self):
        S = self.S
        a = L.HfrLossBwdArgs(self._loss_args, L.ptr(self.w), L.ptr(self.gauss), self.n_global * 3 * S * S,
                             self.n_global, L.ptr(self.g_image), None, None, None,
                             # both shading backwards skip the tiles outside the mesh's tile box, so the loss backward does too
                             ops.raster_tile_box(self.ws, self.B * self.topo.F, self.B),
                             self.aa, 1 if self.tiled else 0, L.ptr(self.gmax_bits) if self.tiled else None)
        L.call("hfr_loss_backward", a)

    def backward(self, pose, betas, focal, prp, root_xyz, shared_grad_hook=None, sums_hook=None):
        """sums_hook(sums) -> work handles (or None): all-reduce of the loss partial sums under data parallelism.  With
        the tiled backward the loss backward does not depend on it (the only global quantity, the mean-RGB scale, is
        applied by the shading backward), so the collective overlaps the loss backward kernel instead of sitting in
        front of it.
        shared_grad_hook(g_texture) -> work handles: called as soon as the gradient of the shared texture is complete
        (after the shade/rasterize backward), waited on after the last kernel of the step is enqueued."""
        # the collective is STARTED before the loss backward is enqueued (it only waits for the loss forward) and the
        # main stream waits for it after: with the tiled backward the two overlap
        pending = (sums_hook(self.sums) or ()) if sums_hook is not None else ()
        if not self.tiled:
            for w in pending:
                w.wait()
            pending = ()
        self.launch_loss_backward()
        for w in pending:
            w.wait()
        if not self.tiled:
            self.acc.zero_()
            if self.n_tex:
                self.g_tex_params.zero_()
        elif not self.deterministic:
            self.g_texture.zero_()
        self.launch_shade_backward()
        works = shared_grad_hook(self.g_texture) if (shared_grad_hook is not None and self.g_texture is not None) else ()
        self.launch_geom_backward(focal, prp, root_xyz)
        ops.mano_backward_raw(self.hm, pose, betas, None, self.g_verts, None, self.g_pose, self.g_betas, None,
                              workspace=self.mano_ws, reuse_forward=True)
        for w in works:
            w.wait()

    def step(self, pose, betas, focal, prp, root_xyz, light_dir, light_color, imgs, seg, tex_params=None):
        self.forward(pose, betas, focal, prp, root_xyz, light_dir, light_color, imgs, seg, tex_params)
        self.backward(pose, betas, focal, prp, root_xyz)

    def capture(self, pose, betas, focal, prp, root_xyz, light_dir, light_color, imgs, seg, shared_grad_hook=None,
                sums_hook=None, tex_params=None):
        """Capture one forward + backward on these (static) input tensors and the CURRENT output set into a CUDA graph;
        `graph.replay()` then enqueues the whole step with one driver call.  The eager step costs ~650 us of host time
        (python + ctypes + 17 launches) against ~900 us of device time at C2 - with several ranks sharing the host's
        cores that is what bounds the step, not the GPU.  The collectives of the hooks are captured with the kernels."""
        cur = torch.cuda.current_stream(self.dev)
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):      # warm-up off the capturing stream (lazy initialisations, allocator)
            for _ in range(2):
                self.forward(pose, betas, focal, prp, root_xyz, light_dir, light_color, imgs, seg, tex_params)
                self.backward(pose, betas, focal, prp, root_xyz, shared_grad_hook=shared_grad_hook, sums_hook=sums_hook)
        cur.wait_stream(side)
        torch.cuda.synchronize(self.dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.forward(pose, betas, focal, prp, root_xyz, light_dir, light_color, imgs, seg, tex_params)
            self.backward(pose, betas, focal, prp, root_xyz, shared_grad_hook=shared_grad_hook, sums_hook=sums_hook)
        return g

    def check_status(self):
        """Host check of the record-store status word (a device->host sync): raises if the store was too small."""
        if self.tiled and int(self.status.item()) != 0:
            raise L.HfrError("FusedHandStep: the (face, tile) record store overflowed; raise rec_per_face")

    def loss_terms(self, sums=None):
        """[texture, mrgb, ssim_tex, sil, iou] (unweighted) from the partial sums (device tensor ops)."""
        s = self.sums if sums is None else sums
        B, S = self.B, self.S
        cnt = float(self.n_global * 3 * S * S)
        mul, add = s[L.LOSS_NSUMS:L.LOSS_NSUMS + B], s[L.LOSS_NSUMS + B:]
        return torch.stack([s[0] / cnt, (s[2] / cnt - s[1] / cnt) ** 2, 1 - s[4] / cnt,
                            s[3] / float(self.n_global * S * S), 1 - (mul / (add - mul)).sum() / self.n_global])


class FusedNimbleStep(FusedHandStep):
    """The same raw-launch step for the NIMBLE-shaped hand (BASELINE configs[2]: V = 5986, F = 11968, 20 shape + 30 pose PCA
    + 10 texture coefficients, 1024^2 texture PCA, K = 1 hard Phong; models_res_nimble.py:133-142 is the layer's
    contract).  Differences from the MANO step: the LBS layer runs the per-sample kernels at NIMBLE size; no joint
    regression / root centring (the reference adds root_xyz only, :203-205 -> geometry with root_out = -1); the texture is
    mean + params @ basis evaluated at the bilinear taps from the texel-major basis (the (B,T,T,3) maps never exist),
    which is a template parameter of the standalone shader - so rasterize and shade are two launches; the backward is
    hfr_shade_backward with the rasterizer backward fused in (no Fragments gradients, no per-face gradient tensor) and
    d/d(texture coefficients) as a step output (g_tex_params); the mean map is frozen.
    `step(pose (B,33), shape (B,20), focal, prp, root_xyz, light_dir, light_color, imgs, seg, tex_params=(B,10))`."""

    def __init__(self, B, image_size=256, texture_size=1024, shape_ncomp=20, pose_ncomp=30, tex_ncomp=10, lambdas=None,
                 device="cuda", n_global=None, sil_scale=1.0):
        self._nimble_cfg = (shape_ncomp, pose_ncomp, tex_ncomp)
        lam = dict(texture=1.0, mrgb=1.0, ssim_tex=1.0, sil=0.0, iou=0.0)
        lam.update(lambdas or {})
        super().__init__(B, image_size=image_size, faces_per_pixel=1, blur_radius=0.0, soft=False, texture_size=texture_size,
                         lambdas=lam, device=device, n_global=n_global, sil_scale=sil_scale, tiled_backward=False,
                         deterministic=True, batched_hand=True, tile_queue=False)
        # LBS, geometry, raster setup + scan + fine pass, shader, loss | loss', shade' + rasterize', geometry', LBS'
        self.launches_per_step = 11 + (4 if self.mano_ws is not None else 0)

    def _build_model(self, dev, mano_root, texture_size):
        from .nimble import MyNIMBLELayer
        ns, npz, nt = self._nimble_cfg
        self.layer = MyNIMBLELayer(True, dev, shape_ncomp=ns, pose_ncomp=npz, tex_ncomp=nt, tex_size=texture_size,
                                   fused_texture=True).to(dev)
        self.hm, self.topo = self.layer.consts(dev)
        self.verts_uvs = self.layer.verts_uvs.to(dev).to(F32).contiguous()
        self.faces_uvs = self.topo.faces                       # the stand-in shares the vertex indexing with its UVs
        self.texture = self.layer.tex_mean.to(dev).to(F32)[None].contiguous()      # (1,T,T,3) mean map
        self.tex_basis = ops.pack_tex_basis(self.layer.tex_basis)                    # texel-major, built once
        self.n_tex = int(self.layer.tex_basis.shape[0])
        self.root_out = -1
        self.texture_grad = False

