"""Render-dependent part of the reference's LossFunction with the same call signature
(losses.py:229-453): LossFunction()(examples, outputs, loss_used, dat_name, args) -> dict.

Implemented terms: 'texture', 'mrgb', 'ssim_tex' (losses.py:355-378; computed whenever the
outputs hold re_img and re_sil, as in the reference), 'sil' (:399-403) and 'iou' (:405-408).
All five come out of ONE forward kernel pass (csrc/loss.cu) and one backward pass.
Terms that do not depend on the render (keypoints, regularisers, VGG perceptual) are out of
the hot path's scope and raise if requested.
"""
from __future__ import annotations

import torch

from . import ops

_RENDER_TERMS = ("texture", "mrgb", "ssim_tex", "sil", "iou")


class LossFunction:
    def __init__(self, sil_scale: float = 255.0):
        # 255: reference mode (re_sil binarised to {0,255}, models_res_nimble.py:219; losses.py:359 divides by 255)
        # 1  : soft-silhouette mode (alpha in [0,1]) used by the north-star configs
        self.sil_scale = float(sil_scale)

    def __call__(self, examples, outputs, loss_used, dat_name, args) -> dict:
        loss_dic = {}
        unknown = [k for k in loss_used if k not in _RENDER_TERMS]
        if unknown:
            raise NotImplementedError(f"loss terms outside the render hot path: {unknown}")
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
