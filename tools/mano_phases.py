"""Phase timeline of mano_bwd_kernel (clock64 of block 0 at the phase boundaries) from a -DHFR_MANO_TIMING build.
usage (GPU box): bash tools/build_variant.sh mt -DHFR_MANO_TIMING   (here), then
                 HFR_B200_LIB=hifihr_b200/_build/lib_mt.so python tools/mano_phases.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import hifihr_b200 as hf  # noqa: E402
from hifihr_b200 import _lib as L  # noqa: E402
from hifihr_b200.synthetic import synthetic_inputs  # noqa: E402

B = 64
step = hf.FusedHandStep(B, image_size=64, faces_per_pixel=1, soft=False, texture_size=32, device="cuda")
inp = synthetic_inputs(B, S=64, seed=1)
fcl, prp = hf.get_ndc_fx_fy_cx_cy(inp["Ks"])
d = lambda t: t.cuda().contiguous()  # noqa: E731
args = (d(inp["pose"]), d(inp["betas"]), d(-fcl), d(prp), d(inp["root_xyz"]), d(inp["light_dir"]), d(inp["light_color"]),
        d(inp["imgs"]), d(inp["segms_gt"].float()))
for _ in range(5):
    step.step(*args)
torch.cuda.synchronize()
buf = (C.c_longlong * 11)()
assert L.lib().hfr_debug_mano_times(buf, 11) == 0
t = list(buf)
names = ["setup (PCA, Rodrigues, J, chain)", "blend recompute", "load g_verts + sums", "route joint grads",
         "gA (skin backward reductions)", "g_vp = T^T g_v", "transposed contraction", "chain backward", "Rodrigues' + betas", "PCA'"]
mhz = 1965.0
for i in range(10):
    print(f"{names[i]:36s} {(t[i + 1] - t[i]) / mhz:7.2f} us")
print(f"{'total':36s} {(t[10] - t[0]) / mhz:7.2f} us")
