"""ctypes binding of libcoclr_b200.so (include/coclr_b200.h).

There is no CPU path: importing is cheap, but every op raises if the CUDA library is missing or
the tensors are not on a CUDA device.  PyTorch is used only for device memory and streams.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libcoclr_b200.so")
_lib = None


class CoclrError(RuntimeError):
    pass


class Src(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("ld", C.c_int), ("coff", C.c_int), ("C", C.c_int),
                ("T", C.c_int), ("H", C.c_int), ("W", C.c_int),
                ("scale", C.c_void_p), ("shift", C.c_void_p), ("relu", C.c_int)]


class Geom(C.Structure):
    _fields_ = [("kt", C.c_int), ("kh", C.c_int), ("kw", C.c_int),
                ("st", C.c_int), ("sh", C.c_int), ("sw", C.c_int),
                ("pt", C.c_int), ("ph", C.c_int), ("pw", C.c_int),
                ("transposed", C.c_int)]


class Conv(C.Structure):
    _fields_ = [("src", Src), ("g", Geom),
                ("B", C.c_int), ("Td", C.c_int), ("Hd", C.c_int), ("Wd", C.c_int),
                ("Kreal", C.c_int), ("wpk", C.c_void_p), ("wunscale", C.c_void_p),
                ("N", C.c_int), ("BN", C.c_int), ("n_tiles", C.c_int),
                ("dst", C.c_void_p), ("dst_ld", C.c_int), ("dst_coff", C.c_int),
                ("accumulate", C.c_int), ("stats", C.c_void_p),
                ("npass", C.c_int), ("bf16", C.c_int)]


class Wgrad(C.Structure):
    _fields_ = [("src", Src), ("g", Geom), ("dy", Src),
                ("B", C.c_int), ("Td", C.c_int), ("Hd", C.c_int), ("Wd", C.c_int),
                ("Cout", C.c_int), ("Cin_real", C.c_int), ("dw", C.c_void_p),
                ("npass", C.c_int), ("bf16", C.c_int), ("splits", C.c_int)]


class Pack(C.Structure):
    _fields_ = [("w", C.c_void_p), ("Cout", C.c_int), ("Cin", C.c_int), ("taps", C.c_int),
                ("Cpad", C.c_int), ("mode", C.c_int), ("bf16", C.c_int),
                ("wpk", C.c_void_p), ("unscale", C.c_void_p)]


def library_path():
    return _LIB_PATH


def load():
    """Load the CUDA library; raises CoclrError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise CoclrError(
            "libcoclr_b200.so is missing (%s). Build it with `python -m coclr_b200.build` "
            "(needs nvcc with sm_100a support). There is no CPU fallback." % _LIB_PATH)
    lib = C.CDLL(_LIB_PATH)
    from . import _signatures
    _signatures.declare(lib)
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        raise CoclrError("%s failed with code %d" % (what, rc))


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def dptr(t):
    if t is None:
        return None
    if not t.is_cuda:
        raise CoclrError("coclr_b200 ops need CUDA tensors (no CPU fallback)")
    return C.c_void_p(t.data_ptr())


_num_sms = {}


def num_sms(device=None):
    dev = torch.cuda.current_device() if device is None else torch.device(device).index
    if dev not in _num_sms:
        _num_sms[dev] = torch.cuda.get_device_properties(dev).multi_processor_count
    return _num_sms[dev]
