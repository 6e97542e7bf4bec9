"""Phase timeline of the small per-sample kernels of the batched hand layer (-DHFR_MANO_TIMING build).
usage: bash tools/build_variant.sh mt -DHFR_MANO_TIMING ; HFR_B200_LIB=hifihr_b200/_build/lib_mt.so python tools/batched_phases.py"""
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
buf = (C.c_longlong * 48)()
assert L.lib().hfr_debug_batched_times(buf) == 0
t = list(buf)
mhz = 1965.0
for k, name in enumerate(("prep", "skin_bwd", "chain_bwd")):
    row = t[16 * k:16 * k + 16]
    pts = [(i, v) for i, v in enumerate(row) if v]
    pts.sort(key=lambda p: p[1])
    print(name, " ".join(f"[{i}] +{(v - pts[0][1]) / mhz:.2f}us" for i, v in pts))
