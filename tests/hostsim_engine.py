"""CPU stand-in for the CUDA library, for tests only (see tests/hostsim/hostsim.cpp)."""
import ctypes
import os
import subprocess

from transcriptclean_b200 import engine as E

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim")
_SO = os.path.join(_DIR, "libtc_hostsim.so")
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build():
    csrc = os.path.join(_ROOT, "transcriptclean_b200", "csrc")
    srcs = [os.path.join(_DIR, "hostsim.cpp"), os.path.join(_ROOT, "include", "tc_b200.h")] + \
           sorted(os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cu", ".cuh", ".cpp")))
    if os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in srcs):
        return _SO
    subprocess.check_call(["g++", "-O2", "-g", "-std=c++17", "-x", "c++", "-shared", "-fPIC", "-pthread", "-o", _SO, srcs[0]])
    return _SO


class HostSimEngine(E.Engine):
    def __init__(self):
        lib = ctypes.CDLL(build())
        E._declare(lib)
        self._setup(lib, 0)
