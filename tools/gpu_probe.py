"""Quick per-kernel probe for the GPU box: runs every entry point once with prints (flush) so a hang or a
launch failure is attributable.  Not a test; `pytest -m gpu` is the parity suite."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
t00 = time.time()


def say(*a):
    print(f"[{time.time() - t00:7.1f}s]", *a, flush=True)


say("import torch")
import torch  # noqa: E402

say("cuda", torch.cuda.is_available(), torch.cuda.get_device_name(0), "cpus", os.cpu_count(), "torch threads", torch.get_num_threads())
import hifihr_b200 as hf  # noqa: E402
from hifihr_b200 import _lib as L  # noqa: E402
from hifihr_b200.synthetic import synthetic_inputs  # noqa: E402

say("lib", L.lib().hfr_device_ok())
dev = "cuda"
B, S, K = int(os.environ.get("PB", 4)), int(os.environ.get("PS", 64)), 4
inp = synthetic_inputs(B, S=S, seed=1)
layer = hf.ManoLayer(center_idx=9, flat_hand_mean=False, ncomps=48)
pose, betas = inp["pose"].to(dev).requires_grad_(True), inp["betas"].to(dev).requires_grad_(True)
say("mano fwd ...")
v, j = layer(pose, betas)
torch.cuda.synchronize()
say("mano fwd ok", float(v.abs().sum()))
(v.square().sum() + j.sum()).backward()
torch.cuda.synchronize()
say("mano bwd ok", float(pose.grad.abs().sum()))
step = hf.FusedHandStep(B, image_size=S, faces_per_pixel=K, soft=True, texture_size=64, device=dev)
fcl, prp = hf.get_ndc_fx_fy_cx_cy(inp["Ks"])
d = lambda t: t.to(dev).contiguous()  # noqa: E731
args = (d(inp["pose"]), d(inp["betas"]), d(-fcl), d(prp), d(inp["root_xyz"]), d(inp["light_dir"]), d(inp["light_color"]),
        d(inp["imgs"]), d(inp["segms_gt"].float()))
from hifihr_b200 import ops  # noqa: E402
say("mano raw ...")
ops.mano_forward_raw(step.hm, args[0], args[1], None, step.verts, None); torch.cuda.synchronize()
say("geom fwd ...")
ops.geom_forward_raw(step.topo, step.verts, 9, args[4], args[2], args[3], step.joints, step.verts_rel, step.verts_view,
                     step.verts_ndc, step.vnormals, step.face_verts); torch.cuda.synchronize()
say("geom ok", float(step.face_verts.abs().sum()))
r = ops.raster_args(step.face_verts, step.mesh_first, step.mesh_nf, S, S, K, step.blur, True, True, False, step.p2f, step.zbuf,
                    step.bary, step.dists, step.ws)
say("raster fwd ...")
L.call("hfr_raster_forward", r); torch.cuda.synchronize()
say("raster ok cover", float((step.p2f[..., 0] >= 0).float().mean()))
say("fused forward ...")
step.forward(*args); torch.cuda.synchronize()
say("forward ok", step.loss_terms().tolist())
say("fused backward ...")
step.backward(args[0], args[1], args[2], args[3], args[4]); torch.cuda.synchronize()
say("backward ok", float(step.g_pose.abs().sum()), float(step.g_texture.abs().sum()))
t0 = time.time()
for _ in range(10):
    step.step(*args)
torch.cuda.synchronize()
say(f"10 steps B={B} S={S}: {(time.time() - t0) * 100:.2f} ms/step")
say("done")
