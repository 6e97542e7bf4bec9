"""TEST INFRASTRUCTURE — import the UNMODIFIED reference ManoLayer (build container only).

Follows SURVEY.md Appendix B: chumpy / pytorch3d are absent, so two stub modules
are registered before ``utils.my_mano`` is imported, and ``ready_arguments``
(utils/mano/webuser/smpl_handpca_wrapper_HAND_only.py:22-67) is replaced by a
loader that returns the same raw arrays ``ManoLayer.__init__`` reads
(utils/my_mano.py:277-313).  ``/root/reference`` does not exist on the GPU box,
so this module is used only by ``oracle/gen_golden.py`` and by CPU tests that
skip when the tree is missing.
"""
import os
import pickle
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("HIFIHR_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "utils", "my_mano.py"))


class _Ch:
    def __setstate__(self, st):
        self.__dict__.update(st)

    @property
    def r(self):
        if hasattr(self, "x"):
            return np.asarray(self.x)
        return self.a.r.ravel()[self.idxs].reshape(self.preferred_shape)


class _Select(_Ch):
    pass


class _R:
    def __init__(self, a):
        self.r = np.asarray(a)


class MeshesStub:
    def __init__(self, verts, faces):
        self.verts, self.faces = verts, faces


_mm = None


def _install():
    global _mm
    if _mm is not None:
        return _mm
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    ch, chch, chre = (types.ModuleType(n) for n in ("chumpy", "chumpy.ch", "chumpy.reordering"))
    chch.Ch = ch.Ch = _Ch
    chre.Select = _Select
    ch.ch, ch.reordering = chch, chre
    sys.modules.update({"chumpy": ch, "chumpy.ch": chch, "chumpy.reordering": chre})
    if "pytorch3d" not in sys.modules:
        p3, st, ms = (types.ModuleType(n) for n in
                      ("pytorch3d", "pytorch3d.structures", "pytorch3d.structures.meshes"))
        ms.Meshes = st.Meshes = MeshesStub
        st.meshes = ms
        p3.structures = st
        sys.modules.update({"pytorch3d": p3, "pytorch3d.structures": st,
                            "pytorch3d.structures.meshes": ms})
    import utils.my_mano as mm  # the reference module, unmodified

    def ready_arguments(path):
        d = pickle.load(open(path, "rb"), encoding="latin1")
        o = dict(d)
        o["shapedirs"] = _R(d["shapedirs"].r)
        o["betas"] = _R(np.zeros(10))
        for k in ("posedirs", "v_template", "weights"):
            o[k] = _R(d[k])
        return o

    mm.ready_arguments = ready_arguments
    _mm = mm
    return mm


def reference_mano_layer(**kw):
    """The reference's ManoLayer, constructed as MyMANOLayer does (utils/my_mano.py:35-36)."""
    mm = _install()
    args = dict(center_idx=9, flat_hand_mean=False, side="right",
                mano_root=os.path.join(REF_ROOT, "utils", "mano"), use_pca=True, ncomps=48)
    args.update(kw)
    return mm.ManoLayer(**args)


def reference_ssim():
    """utils/pytorch_ssim imports cleanly as-is."""
    _install()
    import utils.pytorch_ssim as ps
    return ps


class _StubModule(types.ModuleType):
    """Stand-in for plotting / IO packages the reference imports at module level but never touches on this path."""

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        from unittest import mock
        return mock.MagicMock()


def reference_keypoint_modules():
    """(utils.losses_util, utils.fh_utils, utils.traineval_util) of the reference, unmodified: bone_direction_loss /
    edge_length_loss (losses_util.py:217-301), proj_func (fh_utils.py:30-39), trans_proj_j2d (traineval_util.py:338-354).
    Absent third-party modules (pytorch3d.loss, skimage, matplotlib, ...) are stubbed as they are not on this path."""
    _install()
    import importlib
    for _ in range(40):
        try:
            lu = importlib.import_module("utils.losses_util")
            fh = importlib.import_module("utils.fh_utils")
            tu = importlib.import_module("utils.traineval_util")
            return lu, fh, tu
        except ModuleNotFoundError as e:
            sys.modules[e.name] = _StubModule(e.name)
    raise RuntimeError("could not import the reference keypoint modules")
