"""ctypes binding of libopenrl_b200.so (the C-ABI declared in include/openrl_b200.h).

There is NO fallback: if the library is missing or a symbol is absent the import of the
product path fails loudly (`OrlLibraryError`).  Tensors cross the boundary as raw device
pointers (`tensor.data_ptr()`) plus sizes; calls are asynchronous on the current torch stream.
"""
import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libopenrl_b200.so")
HEADER = os.path.join(os.path.dirname(HERE), "include", "openrl_b200.h")


class OrlLibraryError(RuntimeError):
    pass


class OrlError(RuntimeError):
    pass


_c = ctypes
_P = _c.c_void_p
_I = _c.c_int
_D = _c.c_double
_F = _c.c_float
_L = _c.c_longlong

# name -> argtypes  (restype is int unless noted)
_SIGNATURES = {
    "orl_abi_version": [],
    "orl_device_sm_count": [_c.POINTER(_I)],
    "orl_gae": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _D, _D, _I, _P],
}

_lib = None


def declared_symbols():
    """Every function name declared in include/openrl_b200.h."""
    with open(HEADER) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(orl_[a-z0-9_]+)\s*\(", text)))


def load(path=None):
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise OrlLibraryError(
            f"{p} not found: build it with `python -m openrl_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback."
        )
    try:
        lib = ctypes.CDLL(p)
    except OSError as e:  # pragma: no cover
        raise OrlLibraryError(f"cannot load {p}: {e}") from e
    for name, argtypes in _SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise OrlLibraryError(f"{p} does not export {name}") from e
        fn.argtypes = argtypes
        fn.restype = _I
    lib.orl_last_error.restype = _c.c_char_p
    lib.orl_last_error.argtypes = []
    if lib.orl_abi_version() != 1:
        raise OrlLibraryError("ABI version mismatch")
    if path is None:
        _lib = lib
    return lib


def check(code, what=""):
    if code != 0:
        msg = load().orl_last_error().decode(errors="replace")
        raise OrlError(f"{what} failed with code {code}: {msg}")


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def current_stream():
    import torch

    return torch.cuda.current_stream().cuda_stream
