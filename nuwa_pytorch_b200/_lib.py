"""ctypes binding of libnuwa_b200.so (the C-ABI declared in include/nuwa_b200.h).

The product path has NO fallback: if the shared library is missing or a call fails, an exception is
raised.  PyTorch is used only for device memory and streams; raw device pointers cross the boundary.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libnuwa_b200.so")

c_int, c_void_p, c_float, c_ll = ctypes.c_int, ctypes.c_void_p, ctypes.c_float, ctypes.c_longlong

# name -> argtypes (restype is int unless listed in _RESTYPES)
SIGNATURES = {
    "nuwa_abi_version": [],
    "nuwa_launch_count": [],
    "nuwa_strerror": [c_int],
    "nuwa_gemm_bf16": [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p,
                       c_void_p, c_int, c_int, c_int, c_void_p],
    "nuwa_conv2d_nhwc_bf16": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                              c_void_p, c_void_p, c_int, c_int, c_void_p],
}
_RESTYPES = {"nuwa_strerror": ctypes.c_char_p, "nuwa_launch_count": ctypes.c_ulonglong}

_lib = None


class NuwaB200Error(RuntimeError):
    pass


def lib():
    """Load (once) and return the ctypes handle.  Fails loudly when the extension is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NuwaB200Error(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nuwa_pytorch_b200 has no CPU / eager fallback)")
        handle = ctypes.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the symbol is not exported
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, c_int)
        _lib = handle
    return _lib


def check(code, what):
    if code != 0:
        msg = lib().nuwa_strerror(code).decode()
        raise NuwaB200Error(f"{what} failed: {msg} (code {code})")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    assert t.is_cuda, "nuwa_pytorch_b200 kernels need CUDA tensors (there is no CPU path)"
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def launch_count():
    return int(lib().nuwa_launch_count())
