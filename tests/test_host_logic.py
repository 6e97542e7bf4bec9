"""CPU tests of host-side logic that needs no kernel: the LossFunction terms the reference evaluates as one-line
torch expressions (losses.py:301-313, 431-451) and the drop-in's argument checking."""
import pytest
import torch
import torch.nn.functional as F


class _Args:
    lambda_shape, lambda_pose, lambda_tex_reg, lambda_scale = 0.3, 0.2, 0.05, 2.0


def test_parameter_terms_match_reference_expressions():
    from hifihr_b200.losses import LossFunction
    g = torch.Generator().manual_seed(0)
    out = {"shape_params": torch.randn(4, 10, generator=g), "pose_params": torch.randn(4, 48, generator=g),
           "texture_params": torch.randn(4, 10, generator=g), "joints": torch.randn(4, 21, 3, generator=g)}
    ex = {"scales": torch.rand(4, generator=g)}
    ld = LossFunction()(ex, out, ["mshape", "mpose", "mtex", "scale"], "FreiHand", _Args)
    assert torch.allclose(ld["mshape"], 0.3 * F.mse_loss(out["shape_params"], torch.zeros_like(out["shape_params"])))
    assert torch.allclose(ld["mpose"], 0.2 * F.mse_loss(out["pose_params"], torch.zeros_like(out["pose_params"])))
    assert torch.allclose(ld["mtex"], 0.05 * F.mse_loss(out["texture_params"], torch.zeros_like(out["texture_params"])))
    cal = torch.sqrt(torch.sum((out["joints"][:, 9] - out["joints"][:, 10]) ** 2, 1))
    assert torch.allclose(ld["scale"], 2.0 * F.mse_loss(cal, ex["scales"]))
    # 'scale' is only defined for FreiHand / RHD in the reference (losses.py:304-313)
    assert "scale" not in LossFunction()(ex, out, ["scale"], "HO3D", _Args)


def test_unknown_terms_raise_and_dart_2d_is_rejected():
    from hifihr_b200.losses import LossFunction
    with pytest.raises(NotImplementedError):
        LossFunction()({}, {}, ["perceptual"], "FreiHAND", _Args)
    out = {"joints": torch.zeros(1, 21, 3), "j2d": torch.zeros(1, 21, 2)}
    ex = {"j2d_gt": torch.zeros(1, 21, 2), "Ks": torch.eye(3)[None], "root_xyz": torch.zeros(1, 3)}
    with pytest.raises(NotImplementedError):
        LossFunction()(ex, out, ["joint_2d"], "Dart", _Args)


def test_call_rejects_mixed_devices_bookkeeping():
    """ptr() records the device of every tensor it unpacks and call() launches there; CPU tensors raise."""
    from hifihr_b200 import _lib as L
    with pytest.raises(L.HfrError):
        L.ptr(torch.zeros(3), torch.float32, "x")
    assert L.ptr(None) is None
