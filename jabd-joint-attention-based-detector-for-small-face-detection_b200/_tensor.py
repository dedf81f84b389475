"""Argument marshalling between the reference's Python signatures and the C-ABI.

Inputs may be CUDA torch tensors (used in place), CPU torch tensors or numpy
arrays (staged to the current CUDA device by an explicit H2D copy), or any
object exporting ``__dlpack__``.  Results come back in the kind of the first
array argument (numpy in -> numpy out, CPU tensor in -> CPU tensor out, CUDA in
-> CUDA out).  Compute always happens in ``libjabd_b200.so`` on the GPU: there
is no host implementation to fall back to, so without a CUDA device every
operator raises.
"""
import ctypes

import numpy as np
import torch

KIND_CUDA, KIND_CPU, KIND_NUMPY = "cuda", "cpu", "numpy"


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("jabd_b200 needs a CUDA device (sm_100a); it has no CPU path")


def device_of(*xs):
    """CUDA device of the first CUDA tensor among xs, else the current device."""
    for x in xs:
        if isinstance(x, torch.Tensor) and x.is_cuda:
            return x.device
    require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def kind_of(x):
    if isinstance(x, np.ndarray):
        return KIND_NUMPY
    if isinstance(x, torch.Tensor):
        return KIND_CUDA if x.is_cuda else KIND_CPU
    return KIND_CUDA


def to_dev(x, device, dtype=torch.float32):
    """Contiguous `dtype` tensor on `device` holding x's values."""
    if isinstance(x, torch.Tensor):
        t = x
    elif isinstance(x, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(x))
    elif hasattr(x, "__dlpack__"):
        t = torch.from_dlpack(x)
    else:
        t = torch.as_tensor(x)
    if t.dtype != dtype or t.device != device:
        t = t.to(device=device, dtype=dtype)
    return t.contiguous()


def like(kind, t):
    """Return CUDA result tensor `t` in the caller's kind."""
    if kind == KIND_NUMPY:
        return t.cpu().numpy()
    if kind == KIND_CPU:
        return t.cpu()
    return t


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def stream_of(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def workspace(nbytes, device):
    """Caller-owned scratch for one call (torch's caching allocator: stream-safe, 512-byte aligned)."""
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def variances_of(v):
    v = list(v)
    if len(v) != 2:
        raise ValueError("variances must hold two values (centre, size)")
    return float(v[0]), float(v[1])
