"""ctypes loader for the C-ABI library ``nbss_b200/lib/libnbss_b200.so`` (declared in ``include/nbss_b200.h``).

There is deliberately NO fallback: if the library is missing or a call returns a non-zero status the caller gets a
``RuntimeError``.  Build it with ``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C nbss_b200/csrc``.
"""
from __future__ import annotations

import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# NBSS_LIB selects another build of the same library (tools/phase_profile.py: lib/libnbss_b200_prof.so, `make prof`)
LIB_PATH = os.environ.get("NBSS_LIB") or os.path.join(_HERE, "lib", "libnbss_b200.so")

_lib = None
_lock = threading.Lock()


class NbssError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    """Load (once) and return the CDLL handle. Raises if the library was not built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise NbssError(
                        f"{LIB_PATH} not found: the CUDA extension is not built (run __graft_entry__.build()). "
                        "nbss_b200 has no CPU / PyTorch fallback."
                    )
                _lib = ctypes.CDLL(LIB_PATH)
    return _lib


_STATUS = {-1: "shape/argument error", -2: "null pointer", -3: "unsupported configuration", -4: "workspace too small"}


def check(status: int, what: str) -> None:
    if status == 0:
        return
    if status < 0:
        raise NbssError(f"{what}: {_STATUS.get(status, 'error')} (status {status})")
    raise NbssError(f"{what}: CUDA error {status}")


def ptr(t) -> ctypes.c_void_p:
    """Raw device/host pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr() -> ctypes.c_void_p:
    import torch

    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def device_of(t):
    """Context: make the CUDA device of tensor `t` current for the enclosed launches (the wrappers take the stream of the
    CURRENT device, so a module called with tensors of another device must switch first; a no-op when it already is, and for
    CPU tensors, which the wrappers reject themselves)."""
    import contextlib

    import torch

    if t is not None and getattr(t, "is_cuda", False) and t.device.index != torch.cuda.current_device():
        return torch.cuda.device(t.device)
    return contextlib.nullcontext()
