"""Loader for the in-tree CUDA library.  There is no fallback: if libpopnet_b200.so is missing, stale
or built for another ABI, every product entry point raises."""
import ctypes
import os

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpopnet_b200.so")
_lib = None


class PopnetError(RuntimeError):
    pass


def get():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PopnetError(
                "CUDA library %s not found; build it with `python -m popnet_b200.build` "
                "(there is no CPU fallback)" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        _abi.bind(lib)
        if lib.popnet_abi_version() != _abi.ABI_VERSION:
            raise PopnetError("libpopnet_b200.so ABI %d != expected %d; rebuild" %
                              (lib.popnet_abi_version(), _abi.ABI_VERSION))
        _lib = lib
    return _lib


def check(rc, what):
    if rc != _abi.OK:
        extra = ""
        if rc == -4:
            extra = " (cudaError %d)" % get().popnet_last_cuda_error()
        raise PopnetError("%s failed: %s%s" % (what, _abi.STATUS_NAMES.get(rc, rc), extra))
