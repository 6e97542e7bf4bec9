"""MANO asset loading without chumpy.

The reference loads ``MANO_RIGHT.pkl`` through chumpy (``ready_arguments``,
utils/mano/webuser/smpl_handpca_wrapper_HAND_only.py:22-67) and reads only the
fields listed at utils/my_mano.py:277-313.  chumpy is not needed for the
numbers themselves: the pickle holds plain numpy arrays plus one chumpy
``Select`` object (``shapedirs`` = the first 10 of 20 shape columns).  This
module unpickles with two tiny stand-in classes and returns fp64 numpy arrays.

Search order inside ``mano_root``: ``MANO_<SIDE>.pkl`` then ``MANO_<SIDE>.npz``
(the npz is what ``tools/convert_mano.py`` writes so the asset can travel to a
box where the reference tree is absent).
"""
from __future__ import annotations

import io
import os
import pickle

import numpy as np

_FIELDS = ("hands_components", "hands_mean", "shapedirs", "posedirs", "v_template",
           "J_regressor", "weights", "f", "kintree_table")


class _Ch:
    """Stand-in for chumpy.ch.Ch / chumpy.reordering.Select (state only)."""

    def __setstate__(self, st):
        self.__dict__.update(st)

    @property
    def r(self):
        if hasattr(self, "x"):
            return np.asarray(self.x)
        # chumpy Select: gather `idxs` of the flattened parent, reshape
        return self.a.r.ravel()[self.idxs].reshape(self.preferred_shape)


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith("chumpy"):
            return _Ch
        return super().find_class(module, name)


def _arr(v):
    if hasattr(v, "r"):
        v = v.r
    if hasattr(v, "toarray"):
        v = v.toarray()
    return np.asarray(v)


def load_mano_pkl(path: str) -> dict:
    with open(path, "rb") as fh:
        d = _Unpickler(io.BytesIO(fh.read()), encoding="latin1").load()
    out = {k: _arr(d[k]) for k in _FIELDS}
    out["f"] = out["f"].astype(np.int64)
    out["kintree_table"] = out["kintree_table"].astype(np.int64)
    return out


def default_mano_root() -> str:
    """Directory holding the converted asset (git-ignored, travels with gpurun)."""
    env = os.environ.get("HIFIHR_MANO_ROOT")
    if env:
        return env
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    return os.path.join(here, "baseline", "_ref", "mano")


def load_mano(mano_root: str | None = None, side: str = "right") -> dict:
    """Return the MANO arrays (fp64) the layer needs; raises if absent."""
    roots = [mano_root] if mano_root else []
    roots.append(default_mano_root())
    stem = "MANO_RIGHT" if side == "right" else "MANO_LEFT"
    for root in roots:
        pkl = os.path.join(root, stem + ".pkl")
        npz = os.path.join(root, stem + ".npz")
        if os.path.isfile(pkl):
            return load_mano_pkl(pkl)
        if os.path.isfile(npz):
            z = np.load(npz)
            return {k: z[k] for k in _FIELDS}
    raise FileNotFoundError(
        f"{stem}.pkl/.npz not found under {roots}; run tools/convert_mano.py where the "
        "reference tree is available")
