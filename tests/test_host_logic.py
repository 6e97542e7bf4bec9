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


def test_pack_tex_basis_layout_and_cache():
    """ops.pack_tex_basis: (n,T,T,3) -> texel-major (T*T, 12*ceil(n/4)) records, component k / channel c at 3k + c, zero
    padded (HfrShadeParams.tex_basis_stride); built once per basis tensor and rebuilt when the tensor is modified."""
    from hifihr_b200 import ops
    g = torch.Generator().manual_seed(1)
    for n in (1, 4, 10):
        b = torch.randn(n, 5, 7, 3, generator=g)
        p = ops.pack_tex_basis(b)
        stride = 12 * ((n + 3) // 4)
        assert p.shape == (35, stride) and p.dtype == torch.float32 and p.is_contiguous()
        for k in range(n):
            assert torch.equal(p[:, 3 * k:3 * k + 3].reshape(5, 7, 3), b[k])
        assert (p[:, 3 * n:] == 0).all()
        assert ops.pack_tex_basis(b) is p            # cached on the tensor
        b.mul_(2.0)                                  # in-place change bumps the version counter
        q = ops.pack_tex_basis(b)
        assert q is not p and torch.equal(q[:, :3].reshape(5, 7, 3), b[0])


def test_fused_step_output_layouts():
    """One flat output buffer per step (a single device->host copy): loss sums, then d/d(pose), d/d(shape) and - for the
    NIMBLE-shaped step - d/d(texture coefficients).  Construction is host logic; launching on CPU tensors raises."""
    import hifihr_b200 as hf
    from hifihr_b200 import _lib as L
    m = hf.FusedHandStep(2, image_size=32, texture_size=16, device="cpu")
    assert m.out.numel() == L.LOSS_NSUMS + 2 * 2 + 2 * (48 + 10)
    assert m.g_pose.shape == (2, 48) and m.g_betas.shape == (2, 10) and m.g_tex_params is None
    assert m.g_texture.shape == (1, 16, 16, 3) and m.root_out == 9
    s = hf.FusedNimbleStep(2, image_size=32, texture_size=16, device="cpu")
    assert s.out.numel() == L.LOSS_NSUMS + 2 * 2 + 2 * (33 + 20 + 10)
    assert s.g_pose.shape == (2, 33) and s.g_betas.shape == (2, 20) and s.g_tex_params.shape == (2, 10)
    assert s.g_texture is None and s.root_out == -1 and not s.tiled          # frozen mean map, no root centring
    assert s.params.tex_pca == 10 and s.params.tex_basis_stride == 36 and s.tex_basis.shape == (16 * 16, 36)
    # views alias the flat buffer and follow flip_outputs()
    a = s.g_tex_params.data_ptr()
    s.flip_outputs()
    assert s.g_tex_params.data_ptr() != a
    s.flip_outputs()
    assert s.g_tex_params.data_ptr() == a
    z = torch.zeros(2, 3)
    with pytest.raises(ValueError):                  # the texture coefficients are a required input of the NIMBLE step
        s.forward(z, z, z, z, z, z, z, z, z)
    with pytest.raises(ValueError):
        m.forward(z, z, z, z, z, z, z, z, z, tex_params=z)
    with pytest.raises(L.HfrError):                  # no CPU path
        s.forward(torch.zeros(2, 33), torch.zeros(2, 20), z, z, z, z, z, torch.zeros(2, 3, 32, 32), torch.zeros(2, 32, 32),
                  tex_params=torch.zeros(2, 10))


def test_face_verts_function_has_no_cpu_path():
    from hifihr_b200 import _lib as L
    from hifihr_b200 import ops
    from hifihr_b200.structures import topology_for
    faces = torch.tensor([[0, 1, 2], [2, 1, 3]])
    topo = topology_for(faces, 4, torch.device("cpu"))
    with pytest.raises(L.HfrError):
        ops.FaceVertsFunction.apply(topo, torch.zeros(1, 4, 3))
