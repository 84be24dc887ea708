"""ctypes binding of oracle/libcdssynth.so: the synthetic-image generator built for the host alone (oracle/synth_host.cpp).

TEST / BENCH INFRASTRUCTURE.  bench.py's `--impl reference` leg and the cpu_baseline leg take their inputs from here so that
the CPU arm never maps libcdsgpu.so; the images are bit-identical to the ones the GPU library renders."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libcdssynth.so")
_lib = None


def build(force=False):
    srcs = [os.path.join(_HERE, "synth_host.cpp"), os.path.join(_HERE, "..", "colormipsearch_b200", "csrc", "cds_synth.h")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(_SO) < os.path.getmtime(s) for s in srcs):
        subprocess.run(["make", "-C", _HERE, "-B", "libcdssynth.so"], check=True, stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.cdss_synth_rgb.restype = C.c_int
        L.cdss_synth_rgb.argtypes = [C.c_int, C.c_uint64, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_void_p]
        L.cdss_synth_gradient.restype = C.c_int
        L.cdss_synth_gradient.argtypes = [C.c_uint64, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_void_p]
        _lib = L
    return _lib


def synth_rgb(kind, seed, first_index, n, W, H):
    out = np.empty((n, H, W, 3), np.uint8)
    if lib().cdss_synth_rgb(int(kind), int(seed), int(first_index), int(n), W, H, out.ctypes.data):
        raise ValueError("cdss_synth_rgb: bad arguments")
    return out


def synth_gradient(seed, first_index, n, W, H):
    out = np.empty((n, H, W), np.uint16)
    if lib().cdss_synth_gradient(int(seed), int(first_index), int(n), W, H, out.ctypes.data):
        raise ValueError("cdss_synth_gradient: bad arguments")
    return out
