"""Convert the reference's data/MANO_RIGHT.pkl to baseline/_ref/mano/MANO_RIGHT.npz.

Run in the build container (where /root/reference exists).  The npz is
git-ignored but travels to the GPU box with gpurun.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hifihr_b200.mano_assets import default_mano_root, load_mano_pkl  # noqa: E402


def main(src="/root/reference/data/MANO_RIGHT.pkl"):
    out_dir = default_mano_root()
    os.makedirs(out_dir, exist_ok=True)
    d = load_mano_pkl(src)
    np.savez_compressed(os.path.join(out_dir, "MANO_RIGHT.npz"), **d)
    print("wrote", os.path.join(out_dir, "MANO_RIGHT.npz"), {k: v.shape for k, v in d.items()})


if __name__ == "__main__":
    main(*sys.argv[1:])
