"""ctypes binding of libemip_b200.so (include/emip_b200.h).

There is no CPU or eager-PyTorch fallback: if the library is missing, or a call
fails, the op raises.  The library is built in-tree by ``emip_b200.build``.
"""
import ctypes
import os
import re
import threading

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "_C", "libemip_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_PKG), "include", "emip_b200.h")

_lock = threading.Lock()
_lib = None


class EmipError(RuntimeError):
    pass


def declared_symbols():
    """Names of every function declared in include/emip_b200.h."""
    with open(HEADER_PATH) as f:
        src = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(emip_[a-z0-9_]+)\s*\(", src)))


def lib():
    """Load (once) and return the ctypes handle; raises if the extension is not built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise EmipError(
                        f"{LIB_PATH} not found: build it with `python -m emip_b200.build` "
                        "(there is no CPU fallback for the emip_b200 ops)")
                h = ctypes.CDLL(LIB_PATH)
                h.emip_last_error.restype = ctypes.c_char_p
                _lib = h
    return _lib


def check(code, what):
    if code != 0:
        msg = lib().emip_last_error().decode(errors="replace")
        raise EmipError(f"{what} failed with code {code}: {msg}")


def ptr(t):
    """Device pointer of a tensor (or NULL for None) as c_void_p."""
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def stream_ptr():
    """Current stream of the CURRENT device -- call it inside ``on_device`` (or ``torch.cuda.device(t.device)``)."""
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def on_device(fn):
    """Run ``fn`` with the device of its first CUDA tensor argument current: the C ABI launches on the current device and
    ``stream_ptr()`` hands it that device's current stream, so a tensor on cuda:1 must not be processed under cuda:0."""
    import functools

    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        import torch
        for a in args:
            if isinstance(a, torch.Tensor) and a.is_cuda:
                if a.device.index == torch.cuda.current_device():
                    break
                with torch.cuda.device(a.device):
                    return fn(*args, **kwargs)
        return fn(*args, **kwargs)
    return wrapped


I = ctypes.c_int
LL = ctypes.c_longlong
F = ctypes.c_float
SZ = ctypes.c_size_t
