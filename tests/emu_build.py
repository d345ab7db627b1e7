"""Builds the TEST-ONLY CPU emulation of the CUDA kernels (same sources, g++ -DHIPGP_EMU) into
tests/_emu/libhipgp_emu.so and returns it as a ctypes handle with the C-ABI prototypes attached.
Used by the `not gpu` tests to exercise kernel index arithmetic without a GPU.  Never used by the product."""
import ctypes as C
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC_DIR = os.path.join(ROOT, "hipgp_b200", "csrc")
OUT_DIR = os.path.join(ROOT, "tests", "_emu")
OUT = os.path.join(OUT_DIR, "libhipgp_emu.so")


def _stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    srcs = [os.path.join(SRC_DIR, f) for f in os.listdir(SRC_DIR) if f.endswith((".cu", ".cuh", ".h", ".inl"))]
    srcs.append(os.path.join(ROOT, "include", "hipgp_b200.h"))
    return any(os.path.getmtime(s) > t for s in srcs)


def build():
    os.makedirs(OUT_DIR, exist_ok=True)
    if _stale():
        # a reduced length list keeps the CPU build short; other lengths take the generic kernels (still checked)
        cmd = ["g++", "-O1", "-std=c++17", "-DHIPGP_EMU", "-DHIPGP_DEV_SMALL", "-x", "c++", "-fPIC", "-shared", "-pthread",
               os.path.join(SRC_DIR, "plan.cu"), os.path.join(SRC_DIR, "fast_inst.cu"), "-o", OUT]
        subprocess.run(cmd, check=True)
    return OUT


_lib = None


def load():
    global _lib
    if _lib is None:
        from hipgp_b200 import _lib as L
        _lib = L.declare(C.CDLL(build()))
    return _lib
