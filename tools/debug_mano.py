"""GPU debug: per-sample MANO error vs the fp32/fp64 oracle (not part of the product)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hifihr_b200 as hf
from hifihr_b200.mano_assets import load_mano
from oracle.mano import ManoOracle
from oracle import pipeline as P
mano = load_mano()
inp = P.synthetic_inputs(257, S=8, seed=21)
layer = hf.ManoLayer(center_idx=9, flat_hand_mean=False, ncomps=48)
v, j = layer(inp["pose"].cuda(), inp["betas"].cuda())
v2, j2 = layer(inp["pose"].cuda(), inp["betas"].cuda())
print("deterministic:", torch.equal(v, v2))
o64 = ManoOracle(mano, dtype=torch.float64)
v64, j64 = o64(inp["pose"].double(), inp["betas"].double())
e = (v.cpu().double() - v64).abs().amax(dim=(1, 2))
ej = (j.cpu().double() - j64).abs().amax(dim=(1, 2))
print("max err verts", e.max().item(), "joints", ej.max().item())
top = torch.topk(e, 8)
for val, idx in zip(top.values.tolist(), top.indices.tolist()):
    print(idx, f"{val:.3e}", "joint err %.3e" % ej[idx].item(), "root aa", inp["pose"][idx, :3].tolist(), "angle", inp["pose"][idx, :3].norm().item())
print("median", e.median().item())
