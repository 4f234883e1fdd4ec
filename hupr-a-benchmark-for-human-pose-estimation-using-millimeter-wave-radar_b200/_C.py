"""ctypes binding of libhupr_b200.so — the C ABI declared in include/hupr_b200.h.

There is deliberately NO fallback: if the shared library is missing or a call fails, a RuntimeError is
raised.  Tensors are passed as raw device pointers + the current CUDA stream.
"""
import ctypes
import os

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "libhupr_b200.so")
_lib = None

# name -> (restype, argtypes); must list every symbol of include/hupr_b200.h
SIGNATURES = {
    "hupr_version": (ctypes.c_int, []),
    "hupr_error_string": (ctypes.c_char_p, [ctypes.c_int]),
    "hupr_fft_cascade_i16": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
}


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "hupr_b200: %s not found — build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  There is no CPU/PyTorch fallback." % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(handle, name)   # AttributeError if the symbol is missing -> loud failure
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def check(code, what):
    if code != 0:
        msg = lib().hupr_error_string(code).decode()
        raise RuntimeError("hupr_b200.%s failed: %s (code %d)" % (what, msg, code))


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr())
