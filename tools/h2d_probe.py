"""Host->device copy bandwidth of the bench's per-step input set with all ranks copying at once (no compute).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_probe.py
Prints one JSON line: per-rank and aggregate GB/s for N concurrent pinned-host -> device streams of 12.9 MB chunks.
It answers whether the host (not the GPUs) bounds the end-to-end number at N = 8."""
import json
import os

import torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
nbytes = 64 * (3 * 224 * 224 + 224 * 224) + 64 * 70 * 4          # the C2 e2e input set: 8-bit targets + masks + parameters
host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
devb = torch.empty(nbytes, dtype=torch.uint8, device=dev)
for _ in range(20):
    devb.copy_(host, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
n = 400
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    devb.copy_(host, non_blocking=True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
gbs = nbytes * n / (ms * 1e-3) / 1e9
t = torch.tensor([gbs], device=dev)
allv = [torch.zeros_like(t) for _ in range(world)]
if world > 1:
    dist.all_gather(allv, t)
else:
    allv = [t]
if rank == 0:
    v = [float(x) for x in allv]
    print(json.dumps({"n_ranks": world, "bytes_per_copy": nbytes, "copies": n, "per_rank_GBps": [round(x, 2) for x in v],
                      "aggregate_GBps": round(sum(v), 2), "ms_per_copy_slowest": round(nbytes / (min(v) * 1e9) * 1e3, 4),
                      "cpus_visible": os.cpu_count(), "affinity": sorted(os.sched_getaffinity(0))[:4] + ["..."]}), flush=True)
if world > 1:
    dist.destroy_process_group()
