"""World-size-2 gloo test (CPU) of the multi-GPU host logic: sharding + the two-all-reduce protocol
reproduce the single-process losses exactly.  The per-rank partial sums are formed with the oracle's
formulas (on a GPU box they come from hfr_loss_forward)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hifihr_b200._lib import LOSS_NSUMS
from hifihr_b200.dist import all_reduce_loss_sums, all_reduce_shared_grads, loss_terms_from_sums, shard_range
from oracle import losses as olosses


def _partial_sums(re_img, re_sil, imgs, seg):
    n = re_img.shape[0]
    s = torch.zeros(LOSS_NSUMS + 2 * n)
    segf = seg.unsqueeze(1).float()
    tgt, rim = segf * imgs, re_img * re_sil.repeat(1, 3, 1, 1)
    s[0], s[1], s[2] = (rim - tgt).abs().sum(), rim.sum(), tgt.sum()
    s[3] = (re_sil - segf).abs().sum()
    s[4] = olosses.ssim(rim, tgt, size_average=False).sum() * (3 * re_img.shape[2] * re_img.shape[3])
    s[LOSS_NSUMS:LOSS_NSUMS + n] = (re_sil * segf).reshape(n, -1).sum(1)
    s[LOSS_NSUMS + n:] = (re_sil + segf).reshape(n, -1).sum(1)
    return s


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    N, S = 5, 24
    re_img, re_sil = torch.rand(N, 3, S, S, generator=g), torch.rand(N, 1, S, S, generator=g)
    imgs, seg = torch.rand(N, 3, S, S, generator=g), (torch.rand(N, S, S, generator=g) > 0.5).long()
    lo, hi = shard_range(N, rank, world)
    sums = _partial_sums(re_img[lo:hi], re_sil[lo:hi], imgs[lo:hi], seg[lo:hi])
    all_reduce_loss_sums(sums)
    terms = loss_terms_from_sums(sums, hi - lo, N, S, S)
    shared = torch.full((4,), float(rank + 1))
    all_reduce_shared_grads(shared)
    ref = olosses.render_losses(re_img, re_sil, imgs, seg, dict(texture=1, mrgb=1, ssim_tex=1, sil=1, iou=1), sil_scale=1.0)
    ref = torch.stack([ref[k] for k in ("texture", "mrgb", "ssim_tex", "sil", "iou")])
    q.put((rank, (terms - ref).abs().max().item(), shared.tolist(), (lo, hi)))
    dist.destroy_process_group()


def test_two_rank_protocol_matches_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in procs)
    [p.join(timeout=60) for p in procs]
    assert res[0][3] == (0, 3) and res[1][3] == (3, 5)
    for rank, err, shared, _ in res:
        assert err < 2e-6, (rank, err)
        assert shared == [3.0] * 4


def test_shard_range_covers_batch():
    for n in (1, 7, 64, 4096):
        for w in (1, 2, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
