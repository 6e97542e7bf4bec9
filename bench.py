#!/usr/bin/env python
"""bench.py — hand renders/sec (forward+backward) of the HiFiHR render hot path on B200.

    python bench.py --gpus N --steps K --warmup W          # ours (hand-written sm_100a kernels)
    python bench.py --impl reference --steps K --warmup W  # the reference CPU path (oracle port) on host cores

Workload (N=1): BASELINE.json configs[1] = "C2": MANO LBS + soft rasterization (K=4, blur 9.21e-4) + Phong x UV
texture (512^2, shared) + softmax blend + sil/texture/mrgb/SSIM losses with full backward, batch 64 per GPU,
224x224, synthetic inputs.  One step = MANO -> geometry -> rasterize+shade -> loss ->
loss' -> shade'+rasterize' -> geometry' -> MANO' over one batch.  N>1: one process per GPU (torchrun), the batch
shards by sample (weak scaling, 64 per GPU); NCCL all-reduces the loss partial sums (needed by the mean-RGB term
before its gradient) and the shared-texture gradient.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

CONFIGS = {
    # name: per-GPU batch B (or global batch G sharded over the GPUs = strong scaling), image size, K, soft, texture size
    # BASELINE.json configs[0]: the reference's own CPU-runnable case - the whole batch of 8 is one step of the CPU arm
    "c1": dict(B=8, S=224, K=1, soft=False, T=512, ref_batch=8,
               desc="C1 MANO + hard Phong/UV texture render + losses, 224^2, K=1, batch 8 (BASELINE configs[0])"),
    "c2": dict(B=64, S=224, K=4, soft=True, T=512, desc="C2 MANO soft-raster K=4 + Phong/UV texture + losses, 224^2, B=64/GPU"),
    # configs[2]: NIMBLE-shaped stand-in (V=5986, F=11968), 10 x 1024^2 texture PCA sampled in the shader, modular path
    "c3": dict(B=128, S=256, K=1, soft=False, T=1024, nimble=True,
               desc="C3 NIMBLE-shaped hand (V=5986, F=11968), 1024^2 PCA texture, 256^2, K=1 hard Phong + photometric losses, B=128/GPU"),
    # the same workload through the modular drop-in API (MyNIMBLELayer -> MeshRenderer -> losses under autograd)
    "c3m": dict(B=128, S=256, K=1, soft=False, T=1024, nimble=True, modular=True,
                desc="C3 (modular autograd path) NIMBLE-shaped hand (V=5986, F=11968), 1024^2 PCA texture, 256^2, K=1 hard Phong + photometric losses, B=128/GPU"),
    "c4": dict(G=4096, S=224, K=4, soft=True, T=512, desc="C4 as C2 with global batch 4096 sharded over the GPUs"),
    "c5": dict(G=256, S=512, K=8, soft=True, T=512, desc="C5 soft raster 512^2 K=8 blur>0, global batch 256 sharded over the GPUs"),
    "r": dict(B=48, S=672, K=1, soft=False, T=512, desc="reference setting 672^2 K=1 hard Phong (no pooling stage), B=48"),
    # the reference's own render config, fused (models_res_nimble.py:74-96, 208-220)
    "rp": dict(B=48, S=224, K=1, soft=False, T=512, aa=3, binarize=True, sil_scale=255.0, tiled=False,
               desc="R reference setting: 672^2 K=1 hard Phong + 3x3 SSAA pool + binarised alpha fused, losses at 224^2, B=48"),
}
LAMBDAS = dict(texture=1.0, mrgb=1.0, ssim_tex=1.0, sil=1.0, iou=0.0)


def algorithmic_bytes_per_sample(S, K, V=778, T=512, B=64, n_params=58, aa=1):
    """Fragments written+read (56 B/pixel/K, at the RASTERISED resolution S*aa), image-side
    traffic (64 B/pixel at the loss resolution S), vertex streams, parameters, shared texture read + grad write
    amortised over the per-GPU batch."""
    P = S * S
    return 56 * P * aa * aa * K + 64 * P + 48 * V + 8 * n_params + 24 * T * T / B


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json copy bandwidth)"
    return 6650.0, "fallback (B200_PROFILING.md)"


_wc_keepalive = []


def pinned_buffer(nbytes, write_combined=False):
    """Pinned host staging buffer.  write_combined: cudaHostAllocWriteCombined (the device reads it over PCIe without
    snooping the CPU caches; the host only ever writes it sequentially) - falls back to torch's pinned allocator."""
    if write_combined:
        try:
            import ctypes as C
            rt = C.CDLL("libcudart.so.12")
            ptr = C.c_void_p()
            if rt.cudaHostAlloc(C.byref(ptr), C.c_size_t(nbytes), C.c_uint(0x04)) == 0 and ptr.value:
                arr = (C.c_ubyte * nbytes).from_address(ptr.value)
                _wc_keepalive.append(arr)
                t = torch.frombuffer(arr, dtype=torch.uint8)
                if t.is_pinned():
                    return t
        except Exception:   # noqa: BLE001
            pass
    return torch.empty(nbytes, dtype=torch.uint8).pin_memory()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args, cfg):
    """The reference's CPU implementation of the path, restated (oracle/): reference-identical MANO in torch,
    scalar C naive rasterizer (all host threads), torch CPU shading / blending / losses, autograd backward.
    PyTorch3D's CPU kernel is restated — upstream binary unavailable in this image."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from hifihr_b200.mano_assets import load_mano
    from oracle import pipeline as P
    from oracle import raster_c
    raster_c.build()
    mano = load_mano()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    Bs = cfg.get("ref_batch", 2)                # bounded sample per step (C1: the configuration's own batch of 8)
    S, K = cfg["S"], cfg["K"]
    blur = 9.21034e-4 if cfg["soft"] else 0.0
    tex = P.synthetic_texture(cfg["T"])

    def one_step(seed):
        inp = P.synthetic_inputs(Bs, S=S, seed=seed)
        inp["pose"].requires_grad_(True)
        inp["betas"].requires_grad_(True)
        t = tex.clone().requires_grad_(True)
        out = P.render_path(mano, inp, t, image_size=S, K=K, blur_radius=blur, soft=cfg["soft"], c_select=True,
                            threads=cores, aa=cfg.get("aa", 1), binarize=cfg.get("binarize", False))
        loss, _ = P.total_loss(out, inp, {k: v for k, v in LAMBDAS.items() if v}, cfg.get("sil_scale", 1.0))
        loss.backward()
        return float(loss.detach())

    for w in range(args.warmup):
        one_step(100 + w)
    t0 = time.perf_counter()
    for k in range(args.steps):
        one_step(200 + k)
    dt = time.perf_counter() - t0
    val = Bs * args.steps / dt
    sample = f"{Bs} samples/step of {cfg['desc']} (fwd+bwd), {args.steps} steps"
    line = {"impl": "reference", "metric": "hand renders/sec (fwd+bwd)", "value": val, "unit": "samples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["desc"], "image_size": S, "faces_per_pixel": K, "sample_batch": Bs},
            "cpu_baseline": {"value": val, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args, cfg):
    import torch.distributed as dist
    import hifihr_b200 as hf
    from hifihr_b200 import _lib as L
    from hifihr_b200 import dist as hdist
    from hifihr_b200.synthetic import synthetic_inputs
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # rank 0 prints ONE JSON line on stdout: NCCL's own output (the version banner it prints at VERSION / WARN /
        # INFO) goes to a per-process file instead
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/hifihr_b200_nccl.%h.%p.log")
        dist.init_process_group("nccl", device_id=dev)
    B = cfg["B"] if "B" in cfg else cfg["G"] // world
    if args.batch:
        B = args.batch
    S, K = cfg["S"], cfg["K"]
    aa = cfg.get("aa", 1)
    step = hf.FusedHandStep(B, image_size=S, faces_per_pixel=K, soft=cfg["soft"], texture_size=cfg["T"],
                            lambdas=LAMBDAS, device=dev, n_global=B * world, sil_scale=cfg.get("sil_scale", 1.0),
                            aa_factor=aa, binarize=cfg.get("binarize", False),
                            face_records=bool(os.environ.get("HFR_FACE_RECORDS")),     # env: A/B tuning only
                            tile_queue=os.environ.get("HFR_TILE_QUEUE", "1") != "0",
                            # K=1 hard rasterization at 672^2 has ~1 fragment per covered pixel: the per-tile sort of the
                            # atomics-free backward costs more than it saves there (605 vs 417 us) - that config keeps the
                            # scatter backward (not bit-reproducible); everything else uses the deterministic tiled backward
                            tiled_backward=cfg.get("tiled", True))
    inp = synthetic_inputs(B, S=S, seed=1234 + rank)
    fcl, prp = hf.get_ndc_fx_fy_cx_cy(inp["Ks"])
    # Target images are 8-bit in the datasets the reference trains on (its loader applies ToTensor = x / 255 on the
    # host, utils/traineval_util.py:26-96): the synthetic U(0,1) images are quantised to 8 bits once, the
    # device-resident arm gets them as the floats ToTensor would produce, the end-to-end arm ships the BYTES (and
    # the {0,1} mask as bytes) and the loss kernels convert while loading - same arithmetic, 4x fewer PCIe bytes.
    imgs_u8 = (inp["imgs"] * 255.0).round().to(torch.uint8)
    seg_u8 = inp["segms_gt"].to(torch.uint8)
    small = [inp["pose"], inp["betas"], -fcl, prp, inp["root_xyz"], inp["light_dir"], inp["light_color"]]
    # the step's inputs live in ONE pinned host buffer (256-byte aligned fields) and cross PCIe as ONE copy per step
    fields = [t.contiguous() for t in small + [imgs_u8, seg_u8]]
    offs, total_b = [], 0
    for t in fields:
        offs.append(total_b)
        total_b += (t.numel() * t.element_size() + 255) // 256 * 256
    host = pinned_buffer(total_b, write_combined=os.environ.get("HFR_E2E_WC", "0") == "1")
    for t, o in zip(fields, offs):
        host[o:o + t.numel() * t.element_size()] = t.view(-1).view(torch.uint8)
    devt = [t.to(dev, non_blocking=True) for t in small + [imgs_u8.float() / 255.0, seg_u8.float()]]
    h2d_bytes = sum(t.numel() * t.element_size() for t in fields)

    def field_views(buf):
        return [buf[o:o + t.numel() * t.element_size()].view(t.dtype).view(t.shape) for t, o in zip(fields, offs)]
    out_host = [torch.empty(step.out.shape, dtype=torch.float32).pin_memory() for _ in range(2)]   # sums + g_pose + g_betas
    d2h_bytes = out_host[0].numel() * 4

    def one_step(tensors):
        pose, betas, focal, prpp, root, ldir, lcol, imgs, seg = tensors
        step.forward(pose, betas, focal, prpp, root, ldir, lcol, imgs, seg)
        # loss partial sums (global means of the mean-RGB term): all-reduced while the loss backward kernel runs;
        # gradient of the shared texture: all-reduced while the geometry / hand-layer backward run (both no-ops at N=1)
        step.backward(pose, betas, focal, prpp, root, shared_grad_hook=hdist.all_reduce_shared_grads_async,
                      sums_hook=hdist.all_reduce_loss_sums_async)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # The step is replayed from a CUDA graph (one driver call per step instead of ~17 launches through python/ctypes:
    # 650 us of host time per eager step would bound the 900 us step as soon as several ranks share the host's cores).
    # HFR_GRAPH=0 times the eager launches instead; a failed capture falls back to them and says so in `config`.
    use_graph = os.environ.get("HFR_GRAPH", "1") != "0"
    graph_note = "eager launches"

    def make_graph(tensors):
        return step.capture(*tensors, shared_grad_hook=hdist.all_reduce_shared_grads_async,
                            sums_hook=hdist.all_reduce_loss_sums_async)

    dev_graph = None
    if use_graph:
        try:
            dev_graph = make_graph(devt)
            graph_note = "CUDA graph replay (one graph launch per step)"
        except Exception as e:   # noqa: BLE001
            graph_note = f"eager launches (graph capture failed: {type(e).__name__})"
            dev_graph = None
            torch.cuda.synchronize()
    run_dev = (lambda t: dev_graph.replay()) if dev_graph is not None else one_step

    # ---- device-resident timing ----------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        run_dev(devt)
    barrier()
    # `windows` windows of EXACTLY args.steps steps each, every one bracketed by barrier + synchronize; the reported
    # time is the median window (max over ranks per window), the spread is printed next to it.  Several windows keep
    # the GPU under load long enough for the clock sampler to see it.
    win_ms = []
    with ClockSampler(local) as clk:
        for _ in range(args.windows):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            e0.record()
            for _ in range(args.steps):
                run_dev(devt)
            e1.record()
            barrier()
            win_ms.append(e0.elapsed_time(e1))
    # ---- end to end: pinned host inputs in (8-bit targets), loss + per-sample grads out, every step ---
    # Double-buffered (NB = 2): step i+1's inputs cross PCIe on a copy stream while step i computes (what a DataLoader
    # with pin_memory + non_blocking does for the reference, train_hrnet.py:375-391 / utils/traineval_util.py:26-96).
    # Every step's inputs cross PCIe inside the timed region.  A third buffer (two steps of prefetch) was measured at
    # N = 8 and made the end-to-end step SLOWER (1.10 -> 1.28 ms): more host->device traffic in flight, not less slack,
    # is what costs there.
    NB = 2
    copy_stream = torch.cuda.Stream(device=dev)
    d2h_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream(dev)
    d2h_done = [torch.cuda.Event() for _ in range(2)]
    raw = [torch.empty(total_b, dtype=torch.uint8, device=dev) for _ in range(NB)]
    bufs = [field_views(r) for r in raw]
    ready = [torch.cuda.Event() for _ in range(NB)]
    consumed = [torch.cuda.Event() for _ in range(NB)]

    def enqueue_copy(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])      # the step that last read this slot has finished
            if os.environ.get("HFR_E2E_NOCOPY") != "1":     # diagnostic only: the e2e number needs the copy
                raw[slot].copy_(host, non_blocking=True)
            ready[slot].record(copy_stream)

    e2e_graphs = {}
    if dev_graph is not None:
        try:
            for r in raw:
                r.copy_(host)
            for slot in range(NB):       # every (input slot, output set) combination the loop can meet
                for _ in (0, 1):
                    e2e_graphs[(slot, step._out_set)] = make_graph(bufs[slot])
                    step.flip_outputs()
        except Exception:   # noqa: BLE001
            e2e_graphs = {}
            torch.cuda.synchronize()

    def e2e_run(nsteps):
        # The step's results (loss sums + per-sample gradients, one flat buffer) go back on a third stream while the
        # next step computes into the other output set; a set is only rewritten after its copy has finished.
        for ev in consumed + d2h_done:
            ev.record(main_stream)
        for j in range(min(NB - 1, nsteps)):
            enqueue_copy(j)
        for i in range(nsteps):
            cur = i % NB
            if i + NB - 1 < nsteps:
                enqueue_copy((i + NB - 1) % NB)
            main_stream.wait_event(ready[cur])
            oset = step._out_set
            main_stream.wait_event(d2h_done[oset])      # this output set was copied out (two steps ago)
            g = e2e_graphs.get((cur, oset))
            if g is not None:
                g.replay()
            else:
                one_step(bufs[cur])
            consumed[cur].record(main_stream)
            done = torch.cuda.Event()
            done.record(main_stream)
            with torch.cuda.stream(d2h_stream):
                d2h_stream.wait_event(done)
                if os.environ.get("HFR_E2E_NOD2H") != "1":  # diagnostic only
                    out_host[oset].copy_(step.out, non_blocking=True)
                d2h_done[oset].record(d2h_stream)
            step.flip_outputs()

    e2e_run(3)
    barrier()
    win_e2e = []
    for _ in range(args.windows):
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        f0.record()
        e2e_run(args.steps)
        f1.record()
        barrier()
        win_e2e.append(f0.elapsed_time(f1))
    one_step(devt)      # rebind the cached launch arguments to the output set that is active now (untimed)
    torch.cuda.synchronize()
    # ---- per-kernel durations (CUDA events around each launch group, same stream) -----------------
    names = ["mano_fwd", "geom_fwd", "raster_shade_fwd", "loss_fwd", "loss_bwd", "shade_raster_bwd", "geom_bwd", "mano_bwd"]
    acc = {n: 0.0 for n in names}
    reps = min(args.steps, 10)
    pose, betas, focal, prpp, root, ldir, lcol, imgs, seg = devt

    def timed(name, fn):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        return name, a, b

    from hifihr_b200 import ops
    for _ in range(reps):
        ev = []
        ev.append(timed("mano_fwd", lambda: ops.mano_forward_raw(step.hm, pose, betas, None, step.verts, None, workspace=step.mano_ws)))
        ev.append(timed("geom_fwd", lambda: ops.geom_forward_raw(step.topo, step.verts, 9, root, focal, prpp, step.joints,
                                                                 step.verts_rel, step.verts_view, step.verts_ndc,
                                                                 step.vnormals, step.face_verts)))
        ev.append(timed("raster_shade_fwd", lambda: step.launch_raster_shade(ldir, lcol, imgs)))
        if not step.deterministic:
            step.sums.zero_()
        ev.append(timed("loss_fwd", lambda: L.call("hfr_loss_forward", step._loss_args)))
        ev.append(timed("loss_bwd", lambda: step.launch_loss_backward()))
        if not step.tiled:
            step.acc.zero_()
        ev.append(timed("shade_raster_bwd", lambda: step.launch_shade_backward()))
        ev.append(timed("geom_bwd", lambda: step.launch_geom_backward(focal, prpp, root)))
        ev.append(timed("mano_bwd", lambda: ops.mano_backward_raw(step.hm, pose, betas, None, step.g_verts, None,
                                                                  step.g_pose, step.g_betas, None, workspace=step.mano_ws,
                                                                  reuse_forward=True)))
        torch.cuda.synchronize()
        for n, a, b in ev:
            acc[n] += a.elapsed_time(b)
    kern_ms = {n: acc[n] / reps for n in names}
    # ---- reduce over ranks (max time) -------------------------------------------------------------
    t = torch.tensor([win_ms, win_e2e], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    win_ms, win_e2e = sorted(float(x) for x in t[0]), sorted(float(x) for x in t[1])
    ms, ms_e2e = win_ms[len(win_ms) // 2], win_e2e[len(win_e2e) // 2]
    if rank == 0:
        total = B * world * args.steps
        value = total / (ms * 1e-3)
        e2e = total / (ms_e2e * 1e-3)
        peak, peak_src = peaks()
        top = max(kern_ms, key=kern_ms.get)
        P_ = S * S
        alg = {  # algorithmic bytes per launch of each big kernel
            "raster_shade_fwd": B * (28 * K * aa * aa + 16) * P_,   # Fragments (rasterised res.) + RGBA written once
            "shade_raster_bwd": B * (28 * K * aa * aa + 16) * P_,   # Fragments + image gradient read once
            "loss_fwd": B * (16 + 12 + 4 + 36) * P_,             # RGBA, target, mask read; 9 derivative maps written
            "loss_bwd": B * (16 + 12 + 4 + 36 + 16) * P_,        # the same read again + image gradient written
        }
        ach = alg.get(top, 0) / (kern_ms[top] * 1e-3) / 1e9
        # DRAM bytes of one launch of that kernel from the committed `ncu --set full` capture (same workload only)
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        kname = {"raster_shade_fwd": "raster_shade_fwd_kernel", "shade_raster_bwd": "shade_bwd_tiled_kernel" if step.tiled else "shade_bwd_kernel",
                 "loss_fwd": "loss_fwd_kernel", "loss_bwd": "loss_bwd_kernel"}.get(top)
        if args.config == "c2" and B == 64 and kname and os.path.isfile(tpath):
            tj = json.load(open(tpath))
            traffic = tj.get("dram_bytes_per_launch", {}).get(kname)
            traffic_src = tj.get("source")
        step_bytes = algorithmic_bytes_per_sample(S, K, T=cfg["T"], B=B, aa=aa) * B
        line = {
            "metric": "hand renders/sec (fwd+bwd)", "value": value, "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak" if "B" in cfg else "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["desc"], "batch_per_gpu": B, "global_batch": B * world, "image_size": S,
                       "ssaa": aa, "faces_per_pixel": K, "blur_radius": step.blur, "texture": cfg["T"],
                       "parallelism": f"dp{world} (batch shards by sample; NCCL all-reduce of loss sums + texture grad)",
                       "launch": graph_note,
                       "l2": f"no flush: per-step working set ({(28 * K * aa * aa + 32) * P_ * B / 1e6:.0f} MB Fragments+images) exceeds the 126 MB L2"},
            "e2e": {"value": e2e, "unit": "samples/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "ms_per_step": ms_e2e / args.steps,
                    "inputs": "one packed pinned buffer per step (one H2D copy): pose/shape/camera/light fp32 + target images and masks as uint8 (x/255 fused into the loss kernels)"},
            "windows": {"n": args.windows, "steps_each": args.steps, "statistic": "median",
                        "ms_per_step": [w / args.steps for w in win_ms],
                        "e2e_ms_per_step": [w / args.steps for w in win_e2e]},
            "gpu_launches": step.launches_per_step * args.steps * args.windows,
            "clocks": clk.summary(),
            "roofline": {"bound": "hbm", "kernel": top, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "algorithmic_bytes_per_launch": alg.get(top, 0), "peak_source": peak_src, "kernel_ms": kern_ms,
                         "step_frac": step_bytes / (ms / args.steps * 1e-3) / 1e9 / peak,
                         "step": {"algorithmic_MB_per_sample": step_bytes / B / 1e6,
                                  "achieved": step_bytes / (ms / args.steps * 1e-3) / 1e9,
                                  "frac": step_bytes / (ms / args.steps * 1e-3) / 1e9 / peak}},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(cfg)
        print(json.dumps(line), flush=True)
    if world > 1:
        # graphs that captured NCCL collectives must be gone before the communicator is; tearing the process group down
        # with them alive was seen to hang, so the ranks meet at a barrier and leave without the destructor
        dev_graph = None
        e2e_graphs.clear()
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)



# ------------------------------------------------------------------------------------------------ C3 (NIMBLE-shaped)
def c3_inputs(B, S, seed):
    """Synthetic C3 inputs: pose (3 + 30 PCA), 20 shape, 10 texture coefficients, camera, lights, targets."""
    from hifihr_b200.synthetic import synthetic_inputs
    g = torch.Generator().manual_seed(seed)
    pose = torch.cat([torch.randn(B, 3, generator=g) * 0.4, torch.randn(B, 30, generator=g) * 0.5], 1)
    shape = torch.randn(B, 20, generator=g) * 0.5
    texp = torch.randn(B, 10, generator=g)
    inp = synthetic_inputs(B, S=S, seed=seed + 1)
    root = torch.tensor([[0.0, 0.0, 0.45]]).repeat(B, 1)
    return pose, shape, texp, inp, root


def c3_reference_step(d, pose, shape, texp, inp, root, S, threads):
    """CPU restatement of the C3 step: generic LBS oracle -> NDC -> scalar C rasterizer (selection) + torch autograd
    (values) -> PCA texture -> Phong -> hard blend -> photometric losses, backward."""
    from oracle import losses as olosses
    from oracle import p3d, raster_c
    from oracle.lbs import LBSOracle
    orc = LBSOracle(d["v_template"], d["shapedirs"], d["posedirs"], d["J_regressor"], d["weights"], d["parents"],
                    pca_comps=d["pca_comps"], pose_mean=d["pose_mean"], tip_verts=d["tip_verts"], dtype=torch.float32)
    B = pose.shape[0]
    po, so, to = pose.clone().requires_grad_(True), shape.clone().requires_grad_(True), texp.clone().requires_grad_(True)
    vo, _ = orc(po, so)
    view = vo + root[:, None]
    fcl, prp = p3d.ndc_intrinsics(inp["Ks"])
    ndc = p3d.project_ndc(view, -fcl, prp)
    faces = torch.tensor(d["faces"])
    Fn = faces.shape[0]
    fv = ndc[:, faces].reshape(-1, 3, 3)
    first, nf = [i * Fn for i in range(B)], [Fn] * B
    sel = raster_c.rasterize_naive(fv.detach(), first, nf, S, 0.0, 1, perspective_correct=True, threads=threads)[0]
    fr = p3d.rasterize_meshes(fv, first, nf, S, 0.0, 1, perspective_correct=True, pix_to_face=sel)
    tex_o = d["tex_mean"][None] + torch.einsum("bk,khwc->bhwc", to, d["tex_basis"])
    texels = p3d.sample_textures_uv(fr, tex_o, faces, torch.as_tensor(d["verts_uvs"]))
    colors = p3d.phong_shading(fr, view, faces, texels, inp["light_dir"], inp["light_color"])
    imo = p3d.hard_rgb_blend(colors, fr).permute(0, 3, 1, 2)
    terms = olosses.render_losses(imo[:, :3], imo[:, 3:4], inp["imgs"], inp["segms_gt"], dict(texture=1.0, mrgb=1.0, ssim_tex=1.0),
                                  sil_scale=1.0)
    sum(terms.values()).backward()
    return float(sum(terms.values()).detach())


def run_c3_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from hifihr_b200.nimble import build_nimble_like
    from oracle import raster_c
    raster_c.build()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    d = {k: (torch.tensor(v) if isinstance(v, __import__("numpy").ndarray) and v.dtype.kind == "f" else v)
         for k, v in build_nimble_like(tex_size=cfg["T"]).items()}
    d = {k: (v.float() if torch.is_tensor(v) else v) for k, v in d.items()}
    Bs, S = 1, cfg["S"]
    for w in range(min(args.warmup, 1)):
        c3_reference_step(d, *c3_inputs(Bs, S, 100 + w), S, cores)
    t0 = time.perf_counter()
    for k in range(args.steps):
        c3_reference_step(d, *c3_inputs(Bs, S, 200 + k), S, cores)
    dt = time.perf_counter() - t0
    val = Bs * args.steps / dt
    sample = f"{Bs} sample/step of {cfg['desc']} (fwd+bwd), {args.steps} steps"
    print(json.dumps({"impl": "reference", "metric": "hand renders/sec (fwd+bwd)", "value": val, "unit": "samples/s",
                      "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": cfg["desc"], "image_size": S, "faces_per_pixel": 1, "sample_batch": Bs},
                      "cpu_baseline": {"value": val, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
                      "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                      "gpu_launches": 0}), flush=True)


def run_c3(args, cfg):
    """BASELINE configs[2] as raw launches (hifihr_b200.FusedNimbleStep): per-sample LBS kernels at NIMBLE size -> geometry
    (root shift, camera, normals, packed face vertices) -> rasterizer -> Phong shader evaluating the 10 x 1024^2 texture
    PCA at the bilinear taps -> photometric losses -> their backward (loss', shade' + rasterize' fused, geometry', LBS'),
    replayed from a CUDA graph.  Gradients: pose (B,33), shape (B,20), texture coefficients (B,10), lights."""
    import torch.distributed as dist
    import hifihr_b200 as hf
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/hifihr_b200_nccl.%h.%p.log")
        dist.init_process_group("nccl", device_id=dev)
    B, S, T = args.batch or cfg["B"], cfg["S"], cfg["T"]
    step = hf.FusedNimbleStep(B, image_size=S, texture_size=T, device=dev, n_global=B * world)
    V, Fn = step.hm.V, step.topo.F
    pose, shape, texp, inp, root = c3_inputs(B, S, 1234 + rank)
    fcl, prp = hf.get_ndc_fx_fy_cx_cy(inp["Ks"])
    small = [pose, shape, -fcl, prp, root, inp["light_dir"], inp["light_color"]]
    imgs_u8 = (inp["imgs"] * 255.0).round().to(torch.uint8)
    seg_u8 = inp["segms_gt"].to(torch.uint8)
    host = [t.contiguous().pin_memory() for t in small + [imgs_u8, seg_u8, texp]]
    h2d_bytes = sum(t.numel() * t.element_size() for t in host)
    out_host = torch.empty(step.out.shape, dtype=torch.float32).pin_memory()    # sums + g_pose + g_shape + g_tex_params
    d2h_bytes = out_host.numel() * 4

    from hifihr_b200 import dist as hdist

    def one_step(t):
        step.forward(*t[:9], tex_params=t[9])
        # the mean-RGB term is a difference of GLOBAL means: its sums are all-reduced before the loss backward (no-op at N=1)
        step.backward(*t[:5], sums_hook=hdist.all_reduce_loss_sums_async)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # device-resident arm: float targets (what ToTensor produces); end-to-end arm: the 8-bit targets cross PCIe and the
    # loss kernels convert while loading (as in run_ours)
    devt = [t.to(dev) for t in small] + [imgs_u8.float().to(dev) / 255.0, seg_u8.float().to(dev), texp.to(dev)]
    e2e_in = [torch.empty_like(h, device=dev) for h in host]
    use_graph = os.environ.get("HFR_GRAPH", "1") != "0"
    graph_note, g_dev, g_e2e = "eager launches", None, None
    if use_graph:
        try:
            g_dev = step.capture(*devt[:9], tex_params=devt[9], sums_hook=hdist.all_reduce_loss_sums_async)
            g_e2e = step.capture(*e2e_in[:9], tex_params=e2e_in[9], sums_hook=hdist.all_reduce_loss_sums_async)
            graph_note = "CUDA graph replay (one graph launch per step)"
        except Exception as e:   # noqa: BLE001
            graph_note = f"eager launches (graph capture failed: {type(e).__name__})"
            g_dev = g_e2e = None
            torch.cuda.synchronize()
    run_dev = (lambda: g_dev.replay()) if g_dev is not None else (lambda: one_step(devt))
    run_e2e = (lambda: g_e2e.replay()) if g_e2e is not None else (lambda: one_step(e2e_in))
    for _ in range(max(args.warmup, 3)):
        run_dev()
    barrier()
    win_ms, win_e2e = [], []
    with ClockSampler(local) as clk:
        for _ in range(args.windows):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            e0.record()
            for _ in range(args.steps):
                run_dev()
            e1.record()
            barrier()
            win_ms.append(e0.elapsed_time(e1))
    for _ in range(args.windows):       # end to end: this step's inputs from pinned host memory, its results back to the host
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(args.steps):
            for d, h in zip(e2e_in, host):
                d.copy_(h, non_blocking=True)
            run_e2e()
            out_host.copy_(step.out, non_blocking=True)
        e1.record()
        barrier()
        win_e2e.append(e0.elapsed_time(e1))
    # per-kernel times of one eager step (CUDA events around each stage)
    names = ["lbs_fwd", "geom_fwd", "raster_fwd", "shade_fwd", "loss_fwd", "loss_bwd", "shade_raster_bwd", "geom_bwd", "lbs_bwd"]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
    from hifihr_b200 import _lib as L
    from hifihr_b200 import ops
    kernel_ms = {k: 0.0 for k in names}
    reps = 5
    po, sh, fo, pp, rt, ld, lc, im, sg, tp = devt
    for _ in range(reps):
        step._tex_params, step._focal = tp, fo
        step._inputs = tuple(devt)
        ev[0].record()
        ops.mano_forward_raw(step.hm, po, sh, None, step.verts, None, workspace=step.mano_ws)
        ev[1].record()
        ops.geom_forward_raw(step.topo, step.verts, step.root_out, rt, fo, pp, step.joints, step.verts_rel, step.verts_view,
                             step.verts_ndc, step.vnormals, step.face_verts)
        ev[2].record()
        r = ops.raster_args(step.face_verts, step.mesh_first, step.mesh_nf, S, S, 1, 0.0, True, False, False, step.p2f, step.zbuf,
                            step.bary, step.dists, step.ws, None)
        L.call("hfr_raster_forward", r)
        ev[3].record()
        step.launch_raster_shade(ld, lc, im)          # rasterizer again + shader: the shader's share is the difference
        ev[4].record()
        step.forward(po, sh, fo, pp, rt, ld, lc, im, sg, tp)   # whole forward (keeps the loss arguments current)
        ev[5].record()
        step.launch_loss_backward()
        ev[6].record()
        step.acc.zero_()
        step.g_tex_params.zero_()
        step.launch_shade_backward()
        ev[7].record()
        step.launch_geom_backward(fo, pp, rt)
        ev[8].record()
        ops.mano_backward_raw(step.hm, po, sh, None, step.g_verts, None, step.g_pose, step.g_betas, None, workspace=step.mano_ws,
                              reuse_forward=True)
        ev[9].record()
        torch.cuda.synchronize()
        t = [ev[i].elapsed_time(ev[i + 1]) for i in range(9)]
        kernel_ms["lbs_fwd"] += t[0] / reps
        kernel_ms["geom_fwd"] += t[1] / reps
        kernel_ms["raster_fwd"] += t[2] / reps
        kernel_ms["shade_fwd"] += max(t[3] - t[2], 0.0) / reps
        kernel_ms["loss_fwd"] += max(t[4] - t[0] - t[1] - t[3], 0.0) / reps
        kernel_ms["loss_bwd"] += t[5] / reps
        kernel_ms["shade_raster_bwd"] += t[6] / reps
        kernel_ms["geom_bwd"] += t[7] / reps
        kernel_ms["lbs_bwd"] += t[8] / reps
    t = torch.tensor([win_ms, win_e2e], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    win_ms, win_e2e = sorted(float(x) for x in t[0]), sorted(float(x) for x in t[1])
    ms, ms_e2e = win_ms[len(win_ms) // 2], win_e2e[len(win_e2e) // 2]
    if rank == 0:
        total = B * world * args.steps
        peak, peak_src = peaks()
        P_ = S * S
        # Fragments w+r, image-side traffic, vertex streams, parameters, texture basis + mean read once per step
        step_bytes = (56 * P_ + 64 * P_ + 48 * V + 8 * 60) * B + 12 * 11 * T * T
        alg_raster = B * 28 * P_
        raster_ms = kernel_ms["raster_fwd"]
        line = {"metric": "hand renders/sec (fwd+bwd)", "value": total / (ms * 1e-3), "unit": "samples/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": cfg["desc"], "batch_per_gpu": B, "global_batch": B * world, "image_size": S,
                           "faces_per_pixel": 1, "texture": T, "verts": V, "faces": Fn, "launch": graph_note,
                           "parallelism": f"dp{world} (FusedNimbleStep, batch shards by sample; no shared-parameter gradient: "
                                          "the texture model is frozen)",
                           "l2": "no flush: the texture basis alone (151 MB texel-major) exceeds the 126 MB L2"},
                "e2e": {"value": total / (ms_e2e * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": h2d_bytes,
                        "d2h_bytes_per_step": d2h_bytes, "ms_per_step": ms_e2e / args.steps,
                        "inputs": "pose/shape/texture coefficients/camera/light fp32 + target images and masks as uint8, pinned"},
                "windows": {"n": args.windows, "steps_each": args.steps, "statistic": "median",
                            "ms_per_step": [w / args.steps for w in win_ms], "e2e_ms_per_step": [w / args.steps for w in win_e2e]},
                "gpu_launches": step.launches_per_step * args.steps * args.windows, "clocks": clk.summary(),
                "roofline": {"bound": "hbm", "kernel": "raster_fwd", "achieved": alg_raster / (raster_ms * 1e-3) / 1e9, "peak": peak,
                             "unit": "GB/s", "frac": alg_raster / (raster_ms * 1e-3) / 1e9 / peak, "traffic": None,
                             "algorithmic_bytes_per_launch": alg_raster, "peak_source": peak_src,
                             "kernel_ms": kernel_ms,
                             "step_frac": step_bytes / (ms / args.steps * 1e-3) / 1e9 / peak,
                             "step": {"algorithmic_MB_per_sample": step_bytes / B / 1e6}}}
        print(json.dumps(line), flush=True)
    if world > 1:
        # graphs must go before the process group (see run_ours)
        del g_dev, g_e2e
        dist.destroy_process_group()


def run_c3_modular(args, cfg):
    """BASELINE configs[2] on the modular path: MyNIMBLELayer (LBS kernels at NIMBLE size) -> MeshRasterizer ->
    HardPhongShader with the 10 x 1024^2 texture PCA evaluated at the bilinear taps -> photometric losses, autograd
    bridges over the C-ABI for the backward (`--config c3m`; `--config c3` runs the fused step, run_c3)."""
    import torch.distributed as dist
    import hifihr_b200 as hf
    from hifihr_b200 import ops
    from hifihr_b200.nimble import MyNIMBLELayer
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/hifihr_b200_nccl.%h.%p.log")
        dist.init_process_group("nccl", device_id=dev)
    B, S, T = args.batch or cfg["B"], cfg["S"], cfg["T"]
    layer = MyNIMBLELayer(True, dev, shape_ncomp=20, pose_ncomp=30, tex_ncomp=10, tex_size=T, fused_texture=True).to(dev)
    V, Fn = layer.V, layer.F
    pose, shape, texp, inp, root = c3_inputs(B, S, 1234 + rank)
    fcl, prp = hf.get_ndc_fx_fy_cx_cy(inp["Ks"])
    small = [pose, shape, texp, -fcl, prp, root, inp["light_dir"], inp["light_color"]]
    imgs_u8 = (inp["imgs"] * 255.0).round().to(torch.uint8)
    seg_u8 = inp["segms_gt"].to(torch.uint8)
    host = [t.contiguous().pin_memory() for t in small + [imgs_u8, seg_u8]]
    h2d_bytes = sum(t.numel() * t.element_size() for t in host)
    rs = hf.RasterizationSettings(image_size=S, blur_radius=0.0, faces_per_pixel=1)
    mats = hf.Materials(diffuse_color=((0.8, 0.8, 0.8),), specular_color=((0.2, 0.2, 0.2),), shininess=30, device=dev)
    renderer = hf.MeshRenderer(rasterizer=hf.MeshRasterizer(raster_settings=rs), shader=hf.HardPhongShader(materials=mats, device=dev))
    out_host = torch.empty(3 + B * (33 + 20 + 10), dtype=torch.float32).pin_memory()
    d2h_bytes = out_host.numel() * 4

    def one_step(tensors):
        po, so, to, focal, prpp, rt, ldir, lcol, imgs, seg = tensors
        leaves = [t.detach().requires_grad_(True) for t in (po, so, to)]
        cams = hf.PerspectiveCameras(focal_length=focal, principal_point=prpp, device=dev)
        lights = hf.DirectionalLights(diffuse_color=lcol, direction=ldir, device=dev)
        out = layer({"pose_params": leaves[0], "shape_params": leaves[1], "texture_params": leaves[2]}, handle_collision=False)
        meshes = out["skin_meshes"]
        meshes.offset_verts_(rt[:, None].repeat(1, V, 1).view(B * V, 3))
        img = renderer(meshes, cameras=cams, lights=lights)
        # the output split of models_res_nimble.py:210-220 (permute + slices; no pooling at this render size) as one kernel
        re_img, re_sil, _ = ops.PoolFunction.apply(img, 1, False, None)
        tgt = imgs.float() / 255.0 if imgs.dtype == torch.uint8 else imgs
        sg = seg.float() if seg.dtype == torch.uint8 else seg
        terms = ops.RenderLossFunction.apply(re_img, re_sil, tgt, sg, 1.0, True)
        loss = terms[0] + terms[1] + terms[2]
        if world > 1:
            loss = loss / world
        loss.backward()
        return torch.cat([terms[:3].detach().reshape(-1)] + [t.grad.reshape(-1) for t in leaves])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    devt = [t.to(dev) for t in small] + [imgs_u8.float().to(dev) / 255.0, seg_u8.float().to(dev)]
    for _ in range(max(args.warmup, 3)):
        one_step(devt)
    win_ms, win_e2e = [], []
    with ClockSampler(local) as clk:
        for _ in range(args.windows):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            e0.record()
            for _ in range(args.steps):
                one_step(devt)
            e1.record()
            barrier()
            win_ms.append(e0.elapsed_time(e1))
    for _ in range(args.windows):       # end to end: this step's inputs from pinned host memory, its results back to the host
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(args.steps):
            res = one_step([h.to(dev, non_blocking=True) for h in host])
            out_host.copy_(res, non_blocking=True)
        e1.record()
        barrier()
        win_e2e.append(e0.elapsed_time(e1))
    # the rasterizer alone (the dominant kernel of this shape)
    with torch.no_grad():
        out = layer({"pose_params": devt[0], "shape_params": devt[1], "texture_params": devt[2]}, handle_collision=False)
        meshes = out["skin_meshes"]
        meshes.offset_verts_(devt[5][:, None].repeat(1, V, 1).view(B * V, 3))
        cams = hf.PerspectiveCameras(focal_length=devt[3], principal_point=devt[4], device=dev)
        for _ in range(2):
            renderer.rasterizer(meshes, cameras=cams)
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        for _ in range(5):
            renderer.rasterizer(meshes, cameras=cams)
        r1.record()
        torch.cuda.synchronize()
        raster_ms = r0.elapsed_time(r1) / 5
    t = torch.tensor([win_ms, win_e2e], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    win_ms, win_e2e = sorted(float(x) for x in t[0]), sorted(float(x) for x in t[1])
    ms, ms_e2e = win_ms[len(win_ms) // 2], win_e2e[len(win_e2e) // 2]
    if rank == 0:
        total = B * world * args.steps
        peak, peak_src = peaks()
        P_ = S * S
        # Fragments w+r, image-side traffic, vertex streams, parameters, texture basis + mean read once per step
        step_bytes = (56 * P_ + 64 * P_ + 48 * V + 8 * 60) * B + 12 * 11 * T * T
        alg_raster = B * 28 * P_
        line = {"metric": "hand renders/sec (fwd+bwd)", "value": total / (ms * 1e-3), "unit": "samples/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": cfg["desc"], "batch_per_gpu": B, "global_batch": B * world, "image_size": S,
                           "faces_per_pixel": 1, "texture": T, "verts": V, "faces": Fn,
                           "parallelism": f"dp{world} (modular path, batch shards by sample)",
                           "l2": "no flush: the texture basis alone (132 MB) exceeds the 126 MB L2"},
                "e2e": {"value": total / (ms_e2e * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": h2d_bytes,
                        "d2h_bytes_per_step": d2h_bytes, "ms_per_step": ms_e2e / args.steps},
                "windows": {"n": args.windows, "steps_each": args.steps, "statistic": "median",
                            "ms_per_step": [w / args.steps for w in win_ms], "e2e_ms_per_step": [w / args.steps for w in win_e2e]},
                "gpu_launches": 14 * args.steps * args.windows, "clocks": clk.summary(),
                "roofline": {"bound": "hbm", "kernel": "raster_fwd", "achieved": alg_raster / (raster_ms * 1e-3) / 1e9, "peak": peak,
                             "unit": "GB/s", "frac": alg_raster / (raster_ms * 1e-3) / 1e9 / peak, "traffic": None,
                             "algorithmic_bytes_per_launch": alg_raster, "peak_source": peak_src,
                             "kernel_ms": {"raster_fwd": raster_ms},
                             "step_frac": step_bytes / (ms / args.steps * 1e-3) / 1e9 / peak,
                             "step": {"algorithmic_MB_per_sample": step_bytes / B / 1e6}}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()

def cpu_baseline(cfg):
    """Oracle port timed on this box's host cores on a bounded sample of the same workload (rank 0, N=1)."""
    from hifihr_b200.mano_assets import load_mano
    from oracle import pipeline as P
    from oracle import raster_c
    raster_c.build()
    mano = load_mano()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    S, K = cfg["S"], cfg["K"]
    blur = 9.21034e-4 if cfg["soft"] else 0.0
    Bs, n, t_used = cfg.get("ref_batch", 2), 0, 0.0
    tex = P.synthetic_texture(cfg["T"])
    while t_used < 12.0 and n < 8:
        inp = P.synthetic_inputs(Bs, S=S, seed=500 + n)
        inp["pose"].requires_grad_(True)
        t = tex.clone().requires_grad_(True)
        t0 = time.perf_counter()
        out = P.render_path(mano, inp, t, image_size=S, K=K, blur_radius=blur, soft=cfg["soft"], c_select=True, threads=cores,
                            aa=cfg.get("aa", 1), binarize=cfg.get("binarize", False))
        loss, _ = P.total_loss(out, inp, {k: v for k, v in LAMBDAS.items() if v}, cfg.get("sil_scale", 1.0))
        loss.backward()
        dt = time.perf_counter() - t0
        if n > 0:
            t_used += dt
        n += 1
    done = max(n - 1, 1)
    return {"value": Bs * done / max(t_used, 1e-9), "unit": "samples/s", "cores": cores, "kind": "port",
            "sample": f"{done} x {Bs} samples of the same workload, fwd+bwd (reference-identical MANO in torch, scalar C naive "
                      f"rasterizer on {cores} threads, torch CPU shading/losses; PyTorch3D CPU kernel restated - upstream binary unavailable)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch override")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--windows", type=int, default=5, help="timing windows of --steps steps each (median reported)")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        if cfg.get("nimble"):
            run_c3_reference(args, cfg)
        else:
            run_reference(args, cfg)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; hifihr_b200 has no CPU path (use --impl reference for the CPU arm)")
        if cfg.get("nimble"):
            (run_c3_modular if cfg.get("modular") else run_c3)(args, cfg)
        else:
            run_ours(args, cfg)


if __name__ == "__main__":
    main()
