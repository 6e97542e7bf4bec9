"""TEST INFRASTRUCTURE — CPU restatement of the render-dependent loss terms.

Follows losses.py:355-378 (texture / mrgb / ssim_tex), :399-408 (sil / iou),
:317-340 (self-supervised variants), utils/losses_util.py:366-378 (IOU / iou)
and utils/pytorch_ssim/__init__.py:7-37,65-73 (SSIM; pinned against the
unmodified module in tests/test_oracle_pins.py where the reference tree exists).

``sil_scale`` is 255 in the reference mode (re_sil binarised to {0,255},
models_res_nimble.py:219) and 1 in the soft-silhouette mode the north-star adds.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def gaussian_window(window_size=11, sigma=1.5, dtype=torch.float32):
    g = torch.tensor([math.exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2))
                      for x in range(window_size)])
    g = (g / g.sum()).to(torch.float32)          # pytorch_ssim builds it in fp32 (torch.Tensor)
    w2 = (g[:, None] @ g[None, :]).float()
    return g.to(dtype), w2.to(dtype)


def ssim(img1, img2, window_size=11, size_average=True):
    C = img1.shape[1]
    _, w2 = gaussian_window(window_size, 1.5, img1.dtype)
    win = w2.expand(C, 1, window_size, window_size).contiguous()
    pad = window_size // 2
    mu1 = F.conv2d(img1, win, padding=pad, groups=C)
    mu2 = F.conv2d(img2, win, padding=pad, groups=C)
    mu1_sq, mu2_sq, mu12 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    s1 = F.conv2d(img1 * img1, win, padding=pad, groups=C) - mu1_sq
    s2 = F.conv2d(img2 * img2, win, padding=pad, groups=C) - mu2_sq
    s12 = F.conv2d(img1 * img2, win, padding=pad, groups=C) - mu12
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    m = ((2 * mu12 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))
    return m.mean() if size_average else m.mean(1).mean(1).mean(1)


def iou_loss(sil, gt):
    """losses_util.py:366-378: 1 - mean_b( Σ a·b / (Σ(a+b) - Σ a·b) )."""
    b = sil.shape[0]
    mul = (sil * gt).reshape(b, -1).sum(1)
    add = (sil + gt).reshape(b, -1).sum(1)
    return 1 - (mul / (add - mul)).mean()


def render_losses(re_img, re_sil, imgs, segms_gt, lambdas: dict, sil_scale: float = 255.0,
                  texture_con=None, masked_rgbs=None):
    """Returns dict of the terms present in `lambdas` (keys: texture, mrgb, ssim_tex, sil, iou,
    texture_self, mrgb_self, ssim_tex_self)."""
    out = {}
    seg = segms_gt.unsqueeze(1).to(re_img.dtype)
    target = seg * imgs                                        # losses.py:357
    rim = re_img * (re_sil / sil_scale).repeat(1, 3, 1, 1)     # :359
    if "texture" in lambdas:
        out["texture"] = lambdas["texture"] * (rim - target).abs().mean()
    if "mrgb" in lambdas:
        out["mrgb"] = lambdas["mrgb"] * (target.mean() - rim.mean()) ** 2
    if "ssim_tex" in lambdas:
        out["ssim_tex"] = lambdas["ssim_tex"] * (1 - ssim(rim, target))
    if "sil" in lambdas:
        out["sil"] = lambdas["sil"] * (re_sil - seg).abs().mean()
    if "iou" in lambdas:
        out["iou"] = lambdas["iou"] * iou_loss(re_sil, seg)
    if texture_con is not None and masked_rgbs is not None:    # losses.py:317-340
        c2 = (texture_con ** 2).view(-1, 1, 1, 1)
        if "texture_self" in lambdas:
            wgt = c2.expand_as(re_img)
            out["texture_self"] = lambdas["texture_self"] * ((re_img - masked_rgbs).abs() * wgt).sum() / wgt.sum()
        if "mrgb_self" in lambdas:
            m1 = re_img.reshape(re_img.shape[0], -1).mean(1)
            m2 = masked_rgbs.reshape(re_img.shape[0], -1).mean(1)
            out["mrgb_self"] = lambdas["mrgb_self"] * ((m1 - m2).abs() * texture_con ** 2).sum() / (texture_con ** 2).sum()
        if "ssim_tex_self" in lambdas:
            out["ssim_tex_self"] = lambdas["ssim_tex_self"] * (1 - ssim(re_img, masked_rgbs))
    return out


def texture_metrics(re_img, re_sil, imgs, segms_gt, dat_name="FreiHAND"):
    """train_hrnet.py:149-161: evaluation-time PSNR / SSIM / L1 / L2 between the masked rendering and the masked
    input (mask = re_sil > 0 for HO3D, segms_gt otherwise).  LPIPS (a network) is outside this path."""
    if dat_name == "HO3D":
        m = (re_sil > 0).to(re_img.dtype).repeat(1, 3, 1, 1)
        target, pred = imgs * m, re_img * m                     # :150-152
    else:
        seg = segms_gt.unsqueeze(1).to(re_img.dtype)
        target, pred = seg * imgs, re_img * seg                 # :154-155
    mse = ((pred - target) ** 2).mean()
    return {"psnr": -10 * torch.log10(mse), "ssim": ssim(pred, target), "l1": (pred - target).abs().mean(), "l2": mse}
