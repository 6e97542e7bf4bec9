#!/usr/bin/env python
"""Run under torchrun (N ranks, NCCL): the sharded step (batch slices + the two all-reduces of hifihr_b200.dist)
must reproduce the single-process step on the same global batch — loss terms, per-sample pose gradients, and the
all-reduced gradient of the shared texture.  Rank 0 prints one JSON line and exits non-zero on mismatch.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/multi_gpu_check.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    import hifihr_b200 as hf
    from hifihr_b200 import dist as hdist
    from hifihr_b200.synthetic import synthetic_inputs
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    NG, S, K, T = 8 * world + 3, 96, 4, 64          # ragged: shard sizes differ by one
    lam = dict(texture=1.0, mrgb=1.0, ssim_tex=1.0, sil=1.0, iou=0.0)
    inp = synthetic_inputs(NG, S=S, seed=77)
    fcl, prp = hf.get_ndc_fx_fy_cx_cy(inp["Ks"])
    full = [inp["pose"], inp["betas"], -fcl, prp, inp["root_xyz"], inp["light_dir"], inp["light_color"], inp["imgs"],
            inp["segms_gt"].float()]

    def run(lo, hi, reduce):
        t = [x[lo:hi].contiguous().to(dev) for x in full]
        st = hf.FusedHandStep(hi - lo, image_size=S, faces_per_pixel=K, soft=True, texture_size=T, lambdas=lam, device=dev,
                              n_global=NG)
        st.forward(*t)
        if reduce:
            hdist.all_reduce_loss_sums(st.sums)
        st.backward(t[0], t[1], t[2], t[3], t[4], shared_grad_hook=hdist.all_reduce_shared_grads_async if reduce else None)
        torch.cuda.synchronize()
        return st

    lo, hi = hdist.shard_range(NG, rank, world)
    mine = run(lo, hi, True)
    terms = hdist.loss_terms_from_sums(mine.sums, hi - lo, NG, S, S)
    ok, rep = True, {}
    if rank == 0:
        ref = run(0, NG, False)
        tr = ref.loss_terms()
        rep["loss_terms_abs_err"] = float((terms[:4] - tr[:4]).abs().max())
        rep["g_pose_rel_err"] = float((mine.g_pose - ref.g_pose[lo:hi]).abs().max() / ref.g_pose.abs().max())
        rep["g_texture_rel_err"] = float((mine.g_texture - ref.g_texture).abs().max() / ref.g_texture.abs().max())
        # fp32 sums accumulated in a different order (atomics) on different shard boundaries
        ok = rep["loss_terms_abs_err"] < 2e-6 and rep["g_pose_rel_err"] < 1e-4 and rep["g_texture_rel_err"] < 1e-4
        rep.update(world=world, global_batch=NG, shard=[lo, hi], ok=ok)
        print(json.dumps(rep), flush=True)
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    dist.destroy_process_group()
    sys.exit(int(flag.item() != 0))


if __name__ == "__main__":
    main()
