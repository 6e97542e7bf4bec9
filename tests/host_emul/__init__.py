"""TEST-ONLY: builds tests/host_emul/emul.cpp (the kernels' math headers compiled for the host)."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_SO = os.path.join(_HERE, "_emul.so")


def load():
    src = os.path.join(_HERE, "emul.cpp")
    csrc = os.path.join(_ROOT, "hifihr_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".cuh")]
    if not os.path.isfile(_SO) or any(os.path.getmtime(d) > os.path.getmtime(_SO) for d in deps):
        flags = ["-DHFR_HAVE_SHADE"] if os.path.isfile(os.path.join(csrc, "shade_math.cuh")) else []
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-x", "c++", src, "-o", _SO,
                               "-I", csrc, "-I", os.path.join(_ROOT, "include")] + flags)
    return ctypes.CDLL(_SO)
