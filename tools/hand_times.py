"""Hand layer + geometry stages only, in a loop (tuning helper).
  python tools/hand_times.py [B]                      CUDA-event time of each entry point (warm), per-sample vs batched
  ncu --cache-control none --metrics gpu__time_duration.sum --csv ... python tools/hand_times.py   warm per-kernel times"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import hifihr_b200 as hf  # noqa: E402
from hifihr_b200 import ops  # noqa: E402
from hifihr_b200.synthetic import synthetic_inputs  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
step = hf.FusedHandStep(B, image_size=64, faces_per_pixel=1, soft=False, texture_size=32, device="cuda")
inp = synthetic_inputs(B, S=64, seed=1)
fcl, prp = hf.get_ndc_fx_fy_cx_cy(inp["Ks"])
d = lambda t: t.cuda().contiguous()  # noqa: E731
pose, betas, focal, prpp, root = d(inp["pose"]), d(inp["betas"]), d(-fcl), d(prp), d(inp["root_xyz"])
step.step(pose, betas, focal, prpp, root, d(inp["light_dir"]), d(inp["light_color"]), d(inp["imgs"]), d(inp["segms_gt"].float()))
torch.cuda.synchronize()


def timed(fn, reps=30):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


for ws, name in ((step.mano_ws, "batched"), (None, "per-sample")):
    f = timed(lambda: ops.mano_forward_raw(step.hm, pose, betas, None, step.verts, None, workspace=ws))
    g = timed(lambda: ops.mano_backward_raw(step.hm, pose, betas, None, step.g_verts, None, step.g_pose, step.g_betas, None,
                                            workspace=ws, reuse_forward=True))
    print(f"{name:11s} B={B}: mano_fwd {f:6.1f} us   mano_bwd {g:6.1f} us")
gf = timed(lambda: ops.geom_forward_raw(step.topo, step.verts, 9, root, focal, prpp, step.joints, step.verts_rel, step.verts_view,
                                        step.verts_ndc, step.vnormals, step.face_verts))
gb = timed(lambda: step.launch_geom_backward(focal, prpp, root))
print(f"geom_fwd {gf:6.1f} us   geom_bwd {gb:6.1f} us")
