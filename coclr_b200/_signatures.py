"""argtypes/restype of every symbol declared in include/coclr_b200.h (kept in one place so the CPU
test-suite can check that the built library exports all of them)."""
import ctypes as C

P = C.c_void_p
I = C.c_int

# name -> (restype, argtypes)
EXPORTS = {
    "coclr_conv_igemm": (I, [P, I, P]),
    "coclr_conv_packed_bytes": (C.c_size_t, [I, I, C.POINTER(I), C.POINTER(I)]),
    "coclr_conv_wgrad": (I, [P, P]),
    "coclr_pack_weights": (I, [P, P]),
}


def declare(lib):
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
