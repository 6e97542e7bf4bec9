"""Render-dependent part of the reference's LossFunction with the same call signature
(losses.py:229-453): LossFunction()(examples, outputs, loss_used, dat_name, args) -> dict.

Implemented terms: 'texture', 'mrgb', 'ssim_tex' (losses.py:355-378; computed whenever the
outputs hold re_img and re_sil, as in the reference), 'sil' (:399-403) and 'iou' (:405-408).
All five come out of ONE forward kernel pass (csrc/loss.cu) and one backward pass.
The keypoint / mesh terms next to the render path — 'joint_2d', 'joint_3d',
'vert_3d', 'bone_direc', 'bone_direc_3d', 'edge_length' (losses.py:244-289) and 'mscale' (:293-299) — come out
of one more kernel pair (csrc/keypoint.cu).  `trans_proj_j2d` is utils/traineval_util.py:338-354.
'triangle' (:422-429, uniform Laplacian smoothing) rides in the same kernel pair.  The self-supervised photometric
terms 'texture_self' / 'mrgb_self' / 'ssim_tex_self' (:317-340) are computed, as in the reference, whenever
examples holds 'texture_con' next to rendered outputs (one more pass of the loss kernels in their per-sample-weight
mode).  'mshape' / 'mpose' / 'mtex' (:431-451) and 'scale' (:301-313) are one-line MSE terms on the regressed
parameters, not on the hot path: they are evaluated with plain torch expressions exactly as the reference writes
them.  Terms that need networks or heat-maps (VGG perceptual, hm_integral*, kp_cons, tsa_poses, open_2dj*) raise.
"""
from __future__ import annotations

import torch

from . import ops

_RENDER_TERMS = ("texture", "mrgb", "ssim_tex", "sil", "iou")
_KEYPOINT_TERMS = ("joint_2d", "joint_3d", "vert_3d", "bone_direc", "bone_direc_3d", "edge_length", "mscale", "triangle")
_KP_LAMBDA = dict(joint_2d="lambda_j2d_gt", joint_3d="lambda_j3d", vert_3d="lambda_vert_3d", bone_direc="lambda_bone_direc",
                  bone_direc_3d="lambda_bone_direc_3d", edge_length="lambda_edge_len", mscale="lambda_mscale",
                  triangle="lambda_laplacian")
_PARAM_TERMS = ("mshape", "mpose", "mtex", "scale")          # plain torch, as the reference writes them
_SELF_TERMS = ("texture_self", "mrgb_self", "ssim_tex_self")  # produced whenever texture_con is present


def trans_proj_j2d(outputs, Ks_this, scales=None, is_ortho=False, root_xyz=None, which_joints="joints"):
    """utils/traineval_util.py:338-354 for the perspective / unscaled call the training loop makes
    (train_hrnet.py:83): j2d = proj_func(outputs[which_joints] + root_xyz, Ks)."""
    if scales is not None or is_ortho:
        raise NotImplementedError("trans_proj_j2d: only the perspective, unscaled path (train_hrnet.py:83)")
    j3d = outputs[which_joints]
    _, j2d = ops.KeypointLossFunction.apply(j3d, None, root_xyz, Ks_this, None, None, None, None, None, 0)
    return j2d


def texture_metrics(examples, outputs, dat_name="FreiHAND") -> dict:
    """The evaluation-time texture metrics of train_hrnet.py:149-161 (and compute_texture_metric.py:49-60):
    PSNR, SSIM, L1, L2 between the masked rendering and the masked input, one forward kernel pass
    (csrc/loss.cu, metric mode).  Mask = segms_gt, or re_sil > 0 for dat_name == 'HO3D' (:150-155).
    Returns 0-dim device tensors (the reference calls .item() on each; LPIPS needs a network and is not part
    of this path)."""
    from . import _lib as L
    c = ops._cu
    re_img, re_sil, imgs = c(outputs["re_img"].detach()), c(outputs["re_sil"].detach()), c(examples["imgs"])
    N, _, H, W = re_img.shape
    seg = c(examples["segms_gt"].float()) if "segms_gt" in examples else torch.zeros(N, H, W, device=re_img.device)
    sums = torch.zeros(L.LOSS_NSUMS + 2 * N, dtype=torch.float32, device=re_img.device)
    g = ops.gauss_taps(re_img.device)
    a = L.HfrLossArgs(N, H, W, 1.0, 1, 0, 0, L.ptr(re_img, torch.float32), L.ptr(re_sil, torch.float32),
                      L.ptr(imgs, torch.float32), L.ptr(seg, torch.float32), L.ptr(sums), L.ptr(g), None, None,
                      2 if dat_name == "HO3D" else 1)
    L.call("hfr_loss_forward", a)
    cnt = float(N * 3 * H * W)
    l2 = sums[L.LOSS_L2] / cnt
    return {"psnr": -10 * torch.log10(l2), "ssim": sums[4] / cnt, "l1": sums[0] / cnt, "l2": l2}


_LAP_TOPO = {}


def _laplacian_topology(faces, V):
    """Neighbour lists of a (F,3) face tensor, built once per (F, V, device) and cached: the hand layers emit the same
    topology every step (MANO 1538 faces, NIMBLE-shaped stand-in), and re-deriving it would cost a device->host copy
    per call.  Two different topologies with equal face AND vertex counts on one device are not distinguished."""
    key = (int(faces.shape[0]), int(V), str(faces.device))
    if key not in _LAP_TOPO:
        _LAP_TOPO[key] = ops.TopologyConsts(faces.detach().cpu().numpy(), int(V), device=faces.device)
    return _LAP_TOPO[key]


class LossFunction:
    def __init__(self, sil_scale: float = 255.0):
        # 255: reference mode (re_sil binarised to {0,255}, models_res_nimble.py:219; losses.py:359 divides by 255)
        # 1  : soft-silhouette mode (alpha in [0,1]) used by the north-star configs
        self.sil_scale = float(sil_scale)

    def __call__(self, examples, outputs, loss_used, dat_name, args) -> dict:
        loss_dic = {}
        unknown = [k for k in loss_used if k not in _RENDER_TERMS + _KEYPOINT_TERMS + _PARAM_TERMS + _SELF_TERMS]
        if unknown:
            raise NotImplementedError(f"loss terms outside the render hot path: {unknown}")
        kp = [k for k in loss_used if k in _KEYPOINT_TERMS]
        if kp:
            loss_dic.update(self._keypoint_terms(examples, outputs, kp, args, dat_name))
        loss_dic.update(self._param_terms(examples, outputs, loss_used, dat_name, args))
        if "re_img" in outputs and "re_sil" in outputs and "texture_con" in examples:      # losses.py:317-340
            t = ops.SelfRenderLossFunction.apply(outputs["re_img"], outputs["maskRGBs"], examples["texture_con"])
            loss_dic["texture_self"] = args.lambda_texture * t[0]
            loss_dic["mrgb_self"] = args.lambda_mrgb * t[1]
            loss_dic["ssim_tex_self"] = args.lambda_ssim_tex * t[2]
        if "re_img" in outputs and "re_sil" in outputs:
            seg = examples["segms_gt"].float()
            terms = ops.RenderLossFunction.apply(outputs["re_img"], outputs["re_sil"], examples["imgs"], seg,
                                                 self.sil_scale, True)
            loss_dic["texture"] = args.lambda_texture * terms[0]
            loss_dic["mrgb"] = args.lambda_mrgb * terms[1]
            loss_dic["ssim_tex"] = args.lambda_ssim_tex * terms[2]
            if "sil" in loss_used:
                loss_dic["sil"] = args.lambda_silhouette * terms[3]
            if "iou" in loss_used:
                loss_dic["iou"] = args.lambda_iou * terms[4]
        elif "sil" in loss_used or "iou" in loss_used:
            raise AssertionError("silhouette loss needs rendered sil and gt sil")
        return loss_dic

    @staticmethod
    def _param_terms(examples, outputs, loss_used, dat_name, args):
        """losses.py:301-313 ('scale'), :431-451 ('mshape', 'mpose', 'mtex'): MSE terms on regressed parameters."""
        import torch.nn.functional as torch_f
        d = {}
        if "scale" in loss_used:
            assert ("joints" in outputs) and "scales" in examples, "Using scale as loss but joints not outputted or scales not inputted."
            if dat_name in ("FreiHand", "RHD"):
                cal_scale = torch.sqrt(torch.sum((outputs["joints"][:, 9] - outputs["joints"][:, 10]) ** 2, 1))
                d["scale"] = args.lambda_scale * torch_f.mse_loss(cal_scale, examples["scales"].to(cal_scale.device))
        if "mshape" in loss_used:
            assert "shape_params" in outputs, "Using mshape as loss but shape_params not outputted."
            d["mshape"] = args.lambda_shape * torch_f.mse_loss(outputs["shape_params"], torch.zeros_like(outputs["shape_params"]))
        if "mpose" in loss_used:
            assert "pose_params" in outputs, "Using mpose as loss but pose_params not outputted."
            d["mpose"] = args.lambda_pose * torch_f.mse_loss(outputs["pose_params"], torch.zeros_like(outputs["pose_params"]))
        if "mtex" in loss_used and ("texture_params" in outputs):
            d["mtex"] = args.lambda_tex_reg * torch_f.mse_loss(outputs["texture_params"], torch.zeros_like(outputs["texture_params"]))
        return d

    @staticmethod
    def _keypoint_terms(examples, outputs, used, args, dat_name="FreiHAND"):
        """One kernel pass for every requested keypoint / mesh term; same asserts as losses.py:245-286.  The 2-D
        terms use the projection j2d = proj_func(joints + root_xyz, Ks) (what train_hrnet.py:83 stores in
        outputs['j2d']) fused into the same kernel, so they need examples['Ks'] and examples['root_xyz']."""
        need2d = "joint_2d" in used or "bone_direc" in used
        need3d = "joint_3d" in used or "bone_direc_3d" in used
        needv = "vert_3d" in used or "edge_length" in used
        lap = "triangle" in used
        if need2d and dat_name == "Dart":
            # the reference takes outputs['j2d'] as given; for Dart that is an ORTHOGRAPHIC projection
            # (train_hrnet.py:70-75), which the fused perspective projection of this kernel does not reproduce
            raise NotImplementedError("2-D keypoint terms: dat_name 'Dart' uses an orthographic j2d (train_hrnet.py:70-75); "
                                      "only the unscaled perspective projection of train_hrnet.py:83 is fused here")
        if need2d:
            assert "j2d_gt" in examples and ("j2d" in outputs), "Using joint_2d in losses, but j2d_gt or j2d are not provided."
        if need3d:
            assert "joints" in outputs and "joints" in examples, "Using joint_3d in losses, but joints or joints_gt are not provided."
        if needv:
            assert "mano_verts" in outputs and "verts" in examples, "Using vert_3d in losses, but verts or verts_gt are not provided."
        if "edge_length" in used:
            assert "mano_faces" in outputs, "Using edge_length but verts or faces not outputted."
        if "mscale" in used:
            assert "joints" in outputs, "Using mscale but joints not outputted."
        if need2d and not ("Ks" in examples and "root_xyz" in examples):
            raise NotImplementedError("2-D keypoint terms need examples['Ks'] and examples['root_xyz'] (the projection is fused)")
        l2 = {"L1": 0, "L2": 1}[getattr(args, "base_loss_fn", "L1")]
        vkey, fkey = ("verts", "faces") if ("verts" in outputs and "faces" in outputs) else ("mano_verts", "mano_faces")
        if lap:
            assert fkey in outputs and vkey in outputs, "Using triangle as loss but faces or verts are not outputted."
        faces = outputs["mano_faces"][0] if needv else None
        nbr, pverts = None, (outputs["mano_verts"] if needv else None)
        if lap:
            lf = outputs[fkey][0] if outputs[fkey].dim() == 3 else outputs[fkey]
            topo = _laplacian_topology(lf, outputs[vkey].shape[1])
            nbr = (topo.nbr_ptr, topo.nbr_idx)
            if needv and vkey != "mano_verts":
                raise NotImplementedError("'triangle' on outputs['verts'] cannot be combined with vert_3d / edge_length on mano_verts")
            pverts = outputs[vkey]
        terms, _ = ops.KeypointLossFunction.apply(
            outputs["joints"], pverts,
            examples["root_xyz"] if need2d else None, examples["Ks"] if need2d else None,
            examples["joints"] if need3d else None, examples["j2d_gt"] if need2d else None,
            examples["verts"] if needv else None, None, faces, l2, nbr)
        idx = {k: i for i, k in enumerate(_KEYPOINT_TERMS)}
        return {k: getattr(args, _KP_LAMBDA[k]) * terms[idx[k]] for k in used}
