"""TEST INFRASTRUCTURE — golden vectors for the evaluation-time texture metrics.

Run in the build container (needs /root/reference):  python -m oracle.gen_golden_metrics
Writes tests/golden/texture_metrics_reference.npz: seeded re_img / re_sil / imgs / segms_gt and the PSNR / SSIM /
L1 / L2 the reference computes in train_hrnet.py:149-161 — formed here with the UNMODIFIED utils/pytorch_ssim.ssim
and the two-line LossFunction.MSE_loss / L1_loss bodies (losses.py:455-461; LossFunction itself cannot be
constructed here: its __init__ downloads VGG weights and calls .cuda(), losses.py:232), for both mask branches.
"""
import os

import numpy as np
import torch

from oracle import ref_mano

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def mse_loss(pred, label=0):      # losses.py:455-457
    return ((pred.contiguous() - label) ** 2).mean()


def l1_loss(pred, label=0):       # losses.py:459-461
    return torch.abs(pred.contiguous() - label).mean()


def reference_metrics(ps, examples, outputs, dat_name):
    """train_hrnet.py:149-161, statement for statement (LPIPS dropped)."""
    if dat_name == 'HO3D':
        maskRGBs = examples['imgs'].mul((outputs['re_sil'] > 0).float().repeat(1, 3, 1, 1))
        mask_re_img = outputs['re_img'].mul((outputs['re_sil'] > 0).float().repeat(1, 3, 1, 1))
    else:
        maskRGBs = examples['segms_gt'].unsqueeze(1) * examples['imgs']
        mask_re_img = outputs['re_img'] * examples['segms_gt'].unsqueeze(1)
    psnr = -10 * mse_loss(mask_re_img, maskRGBs).log10().item()
    ssim = ps.ssim(mask_re_img, maskRGBs).item()
    l1 = l1_loss(mask_re_img, maskRGBs).mean().item()
    l2 = mse_loss(mask_re_img, maskRGBs).mean().item()
    return np.array([psnr, ssim, l1, l2], np.float64)


def main():
    ps = ref_mano.reference_ssim()
    g = torch.Generator().manual_seed(4242)
    N, S = 3, 72                       # 72 = 2.25 loss tiles per side: interior, edge and partial tiles
    imgs = torch.rand(N, 3, S, S, generator=g)
    re_img = (imgs + 0.15 * torch.randn(N, 3, S, S, generator=g)).clamp(0, 1)
    yy, xx = torch.meshgrid(torch.arange(S), torch.arange(S), indexing="ij")
    seg = torch.stack([(((yy - 30 - 4 * n) ** 2 + (xx - 36) ** 2) <= (14 + 3 * n) ** 2).float() for n in range(N)])
    sil = torch.stack([(((yy - 34) ** 2 + (xx - 30 - 5 * n) ** 2) <= (16 + 2 * n) ** 2).float() for n in range(N)]) * 255.0
    ex, out = {"imgs": imgs, "segms_gt": seg}, {"re_img": re_img, "re_sil": sil[:, None]}
    np.savez_compressed(os.path.join(OUT, "texture_metrics_reference.npz"), imgs=imgs.numpy(), re_img=re_img.numpy(),
                        segms_gt=seg.numpy(), re_sil=sil[:, None].numpy(),
                        freihand=reference_metrics(ps, ex, out, "FreiHAND"), ho3d=reference_metrics(ps, ex, out, "HO3D"))
    print("wrote texture_metrics_reference.npz")


if __name__ == "__main__":
    main()
