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


def test_workspace_size_queries_are_host_only_and_consistent():
    """The size queries of the round-2 entry points are plain host arithmetic (no GPU): tile queue, hand-layer
    workspace / packed basis, record partials.  Sizes follow the layouts documented in include/hifihr_b200.h."""
    import ctypes as C
    from hifihr_b200 import _lib as L
    lib = L.lib()
    # tile queue: 16 header words + (1 count + 8 class lists + 1 group list) words per 16x16 tile
    T = 64 * 14 * 14
    assert lib.hfr_raster_queue_bytes(64, 224, 224) == (16 + 10 * T) * 4 + 64
    assert lib.hfr_raster_queue_bytes(1, 17, 33) == (16 + 10 * (2 * 3)) * 4 + 64      # ragged sizes round tiles up
    # MANO-shaped model: V=778, NJ=16, NS=10 -> NK=145, KP=152, NKP16=160, C3=2336 -> C3P=2368, 73 forward tiles, 37 chunks
    m = L.HfrHandModel()
    m.V, m.NJ, m.NS, m.NPC, m.NW, m.NT, m.center_joint, m.C3 = 778, 16, 10, 45, 4, 5, 9, 2336
    fwd, bwd = 73 * 2 * 32 * 152, 37 * 2 * 160 * 64
    assert lib.hfr_mano_packed_basis_bytes(C.byref(m)) == (fwd + bwd + 32) * 4
    w64, w4096 = lib.hfr_mano_workspace_bytes(C.byref(m), 64), lib.hfr_mano_workspace_bytes(C.byref(m), 4096)
    assert 0 < w64 < w4096 and w64 % 4 == 0
    # one 128-sample tile of operands at B=64, 32 at B=4096; split-K shrinks as the batch grows (enough CTAs already)
    assert w4096 < 64 * w64
    t = L.HfrTopology()
    t.V, t.F = 778, 1538
    assert lib.hfr_geom_rec_partial_floats(C.byref(t), 64) == 64 * 3 * 1538 * 6
