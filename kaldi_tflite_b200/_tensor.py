"""Plumbing between user arrays (numpy / torch / DLPack) and raw device pointers."""

import ctypes

import numpy as np
import torch

from . import _native


def device():
    return torch.device("cuda", torch.cuda.current_device())


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def is_numpy_like(x):
    return not isinstance(x, torch.Tensor) and not hasattr(x, "__dlpack__") or isinstance(x, np.ndarray)


def as_device(x, dtype=torch.float32):
    """numpy / list / torch (any device) / DLPack -> contiguous CUDA tensor of `dtype`."""
    _native.require_cuda()
    if isinstance(x, torch.Tensor):
        t = x
    elif isinstance(x, np.ndarray):
        x = np.ascontiguousarray(x)
        if not x.flags.writeable:   # e.g. np.frombuffer views of a wav file: torch wants a writable buffer
            x = x.copy()
        t = torch.from_numpy(x)
    elif hasattr(x, "__dlpack__"):
        t = torch.from_dlpack(x)
    else:
        t = torch.as_tensor(np.asarray(x))
    return t.to(device=device(), dtype=dtype, non_blocking=True).contiguous()


def is_int16(x):
    """True for numpy / torch int16 arrays: raw PCM that stays 16-bit up to the front-end kernel."""
    dt = getattr(x, "dtype", None)
    return dt is torch.int16 or (isinstance(dt, np.dtype) and dt == np.int16)


def ptr(t):
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def host_ptr(a):
    """Pointer to a C-contiguous numpy array (kept alive by the caller)."""
    if a is None:
        return ctypes.c_void_p(0)
    assert a.flags["C_CONTIGUOUS"]
    return ctypes.c_void_p(a.ctypes.data)


def like_input(out, ref):
    """numpy in -> numpy out; everything else stays a CUDA torch tensor (DLPack-exportable)."""
    if isinstance(ref, torch.Tensor) or (hasattr(ref, "__dlpack__") and not isinstance(ref, np.ndarray)):
        return out
    return out.cpu().numpy()


_uniform_offsets = {}


def uniform_offsets(batch, rows):
    """offsets[b] = b * rows on the device; cached per (device, batch, rows) -- treat as read-only."""
    dev = device()
    key = (str(dev), int(batch), int(rows))
    t = _uniform_offsets.get(key)
    if t is None:
        t = torch.arange(batch + 1, device=dev, dtype=torch.int64) * rows
        # (a tensor first produced while a CUDA graph is being captured is only filled when the graph runs: never cache it)
        if not (dev.type == "cuda" and torch.cuda.is_current_stream_capturing()):
            if len(_uniform_offsets) > 64:
                _uniform_offsets.clear()
            _uniform_offsets[key] = t
    return t
