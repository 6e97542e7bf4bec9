"""Batched hand layer (one tensor-core blend contraction per batch, csrc/mano_batched.cu) through the C-ABI:

* against the golden vectors of the UNMODIFIED reference ManoLayer (utils/my_mano.py:315-483), every input mode,
  verts / joints 1e-6 m abs, gradients 1e-3 of the tensor max - the same bar as the per-sample kernels;
* against the fp64 oracle at batch sizes that span several 128-sample MMA tiles and a ragged last tile;
* against the per-sample kernels (same formulas, fp32 FMA contraction): verts 2e-7 m, gradients 1e-5;
* bit-reproducible run to run (fixed-order split-K sum), and no barrier wait ever timed out.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import pipeline as P  # noqa: E402
from oracle.mano import ManoOracle  # noqa: E402

DEV = "cuda"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel_err(got, ref):
    ref = ref.detach().cpu().double()
    got = got.detach().cpu().double()
    return float((got - ref).abs().max() / max(1e-12, ref.abs().max()))


@pytest.fixture(scope="module")
def hf():
    import hifihr_b200
    assert os.path.isfile(hifihr_b200.LIB_PATH), "libhifihr_b200.so missing: the CUDA path is the only path"
    return hifihr_b200


@pytest.fixture(scope="module")
def mano():
    from hifihr_b200.mano_assets import load_mano
    return load_mano()


def _batched(layer, on=True):
    hm = layer.consts(torch.device(DEV))
    assert hm.basis_packed is not None, "the packed basis was not built"
    hm.batched_min = 1 if on else 10 ** 9
    return layer


def _status(hf):
    return int(hf._lib.lib().hfr_mano_batched_status())


def test_batched_matches_reference_golden(hf):
    z = np.load(os.path.join(GOLD, "mano_reference.npz"))
    layer = _batched(hf.ManoLayer(center_idx=9, flat_hand_mean=False, side="right", use_pca=True, ncomps=48))
    pose = torch.tensor(z["pose"], device=DEV, requires_grad=True)
    beta = torch.tensor(z["beta"], device=DEV, requires_grad=True)
    v, j = layer(pose, beta)
    assert (v.detach().cpu() - torch.tensor(z["verts"])).abs().max() < 1e-6
    assert (j.detach().cpu() - torch.tensor(z["joints"])).abs().max() < 1e-6
    ((v * torch.tensor(z["g_verts"], device=DEV)).sum() + (j * torch.tensor(z["g_joints"], device=DEV)).sum()).backward()
    assert rel_err(pose.grad, torch.tensor(z["g_pose"])) < 1e-3
    assert rel_err(beta.grad, torch.tensor(z["g_beta"])) < 1e-3
    assert _status(hf) == 0


def test_batched_modes_match_reference_golden(hf):
    from oracle.gen_golden import MODE_CASES
    z = np.load(os.path.join(GOLD, "mano_modes_reference.npz"))
    for name, (ckw, _, fkw) in MODE_CASES.items():
        kw = dict(ncomps=48, flat_hand_mean=False, center_idx=9)
        kw.update(ckw)
        layer = _batched(hf.ManoLayer(**kw))
        pose = torch.tensor(z[f"{name}.pose"], device=DEV, requires_grad=True)
        beta = torch.tensor(z[f"{name}.beta"], device=DEV, requires_grad=True)
        fw = {}
        if fkw.get("root_palm"):
            fw["root_palm"] = torch.Tensor([1])
        if fkw.get("share_betas"):
            fw["share_betas"] = torch.Tensor([1])
        trans = None
        if fkw.get("trans"):
            trans = torch.tensor(z[f"{name}.trans"], device=DEV, requires_grad=True)
            fw["th_trans"] = trans
        v, j = layer(pose, torch.zeros(1) if fkw.get("mean_shape") else beta, **fw)
        tol = 2e-6 if name == "rotmat" else 1e-6
        assert (v.detach().cpu() - torch.tensor(z[f"{name}.verts"])).abs().max() < tol, name
        assert (j.detach().cpu() - torch.tensor(z[f"{name}.joints"])).abs().max() < tol, name
        ((v * torch.tensor(z[f"{name}.g_verts"], device=DEV)).sum()
         + (j * torch.tensor(z[f"{name}.g_joints"], device=DEV)).sum()).backward()
        assert rel_err(pose.grad, torch.tensor(z[f"{name}.g_pose"])) < 1e-3, name
        if f"{name}.g_beta" in z.files:
            assert rel_err(beta.grad, torch.tensor(z[f"{name}.g_beta"])) < 1e-3, name
        if trans is not None:
            assert rel_err(trans.grad, torch.tensor(z[f"{name}.g_trans"])) < 1e-3, name
    assert _status(hf) == 0


@pytest.mark.parametrize("B", [8, 64, 130, 257])
def test_batched_vs_fp64_oracle_and_per_sample(hf, mano, B):
    inp = P.synthetic_inputs(B, S=8, seed=40 + B)
    g = torch.Generator().manual_seed(B)
    gv, gj = torch.randn(B, 778, 3, generator=g).to(DEV), torch.randn(B, 21, 3, generator=g).to(DEV)
    outs = {}
    for mode in ("batched", "per_sample"):
        layer = _batched(hf.ManoLayer(center_idx=9, flat_hand_mean=False, ncomps=48), on=mode == "batched")
        pose = inp["pose"].to(DEV).requires_grad_(True)
        beta = inp["betas"].to(DEV).requires_grad_(True)
        v, j = layer(pose, beta)
        ((v * gv).sum() + (j * gj).sum()).backward()
        outs[mode] = (v.detach(), j.detach(), pose.grad.clone(), beta.grad.clone())
    o64 = ManoOracle(mano, dtype=torch.float64)
    po, bo = inp["pose"].double().requires_grad_(True), inp["betas"].double().requires_grad_(True)
    vo, jo = o64(po, bo)
    ((vo * gv.cpu().double()).sum() + (jo * gj.cpu().double()).sum()).backward()
    vb, jb, gpb, gbb = outs["batched"]
    assert (vb.cpu().double() - vo.detach()).abs().max() < 1e-6 and (jb.cpu().double() - jo.detach()).abs().max() < 1e-6
    assert rel_err(gpb, po.grad) < 1e-4 and rel_err(gbb, bo.grad) < 1e-4
    vs, js, gps, gbs = outs["per_sample"]
    assert (vb - vs).abs().max() < 2e-7 and (jb - js).abs().max() < 2e-7
    assert rel_err(gpb, gps) < 1e-5 and rel_err(gbb, gbs) < 1e-5
    assert _status(hf) == 0


def test_batched_is_bit_reproducible(hf):
    B = 96
    inp = P.synthetic_inputs(B, S=8, seed=77)
    g = torch.Generator().manual_seed(3)
    gv = torch.randn(B, 778, 3, generator=g).to(DEV)
    runs = []
    for _ in range(3):
        layer = _batched(hf.ManoLayer(center_idx=9, flat_hand_mean=False, ncomps=48))
        pose = inp["pose"].to(DEV).requires_grad_(True)
        beta = inp["betas"].to(DEV).requires_grad_(True)
        v, j = layer(pose, beta)
        (v * gv).sum().backward()
        runs.append((v.detach().clone(), pose.grad.clone(), beta.grad.clone()))
    for r in runs[1:]:
        for a, b in zip(r, runs[0]):
            assert torch.equal(a, b)
