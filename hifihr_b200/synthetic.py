"""Synthetic workload of SURVEY.md §8(d) (seeded, CPU tensors): poses, shapes, FreiHAND-like cameras,
lights, target images and disc masks.  bench.py and examples draw their inputs here; the oracle keeps
its own identical generator and tests/test_oracle_pins.py checks the two agree."""
from __future__ import annotations

import torch


def synthetic_inputs(B, S=224, seed=1234, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.rand(*s, generator=g)  # noqa: E731
    n = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
    pose = torch.cat([n(B, 3) * 1.5, n(B, 45) * 0.5], 1)          # global rotation, 45 PCA coefficients
    betas = n(B, 10) * 0.5
    root_xyz = torch.stack([r(B) * 0.06 - 0.03, r(B) * 0.06 - 0.03, r(B) * 0.2 + 0.55], 1)
    f = r(B) * 80 + 440                                            # focal in 224-px units
    c = 112 + r(B, 2) * 16 - 8
    Ks = torch.zeros(B, 3, 4)
    Ks[:, 0, 0] = f
    Ks[:, 1, 1] = f
    Ks[:, 0, 2] = c[:, 0]
    Ks[:, 1, 2] = c[:, 1]
    Ks[:, 2, 2] = 1
    imgs = r(B, 3, S, S)
    yy, xx = torch.meshgrid(torch.arange(S), torch.arange(S), indexing="ij")
    disc = (((yy - S / 2 + 0.5) ** 2 + (xx - S / 2 + 0.5) ** 2) <= (0.3 * S) ** 2).long()
    seg = disc[None].repeat(B, 1, 1)
    light_color = r(B, 3) * 0.8 + 0.2
    light_dir = n(B, 3)
    out = dict(pose=pose, betas=betas, root_xyz=root_xyz, Ks=Ks, imgs=imgs, segms_gt=seg,
               light_color=light_color, light_dir=light_dir)
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in out.items()}
