"""Tensor plumbing between the caller's framework and the raw pointers of the C ABI.

Anything that speaks DLPack (`__dlpack__`: torch, TensorFlow >= 2.x eager tensors, CuPy, JAX) or a
legacy DLPack capsule (`tf.experimental.dlpack.to_dlpack(t)`) is borrowed zero-copy as a torch view;
torch is used purely as the device-memory / stream / allocator layer.  CPU tensors are rejected:
there is no CPU fallback (use the `host_*` functions of `hostapi.py` for host buffers, which still
compute on the GPU).
"""
from __future__ import annotations

import torch


def from_any(x, *, name="tensor") -> torch.Tensor:
    """Borrow `x` as a torch tensor without copying.  A TensorFlow tensor is produced on TensorFlow's own stream,
    which DLPack does not carry: its device work is drained first (see tf_adapter.py)."""
    if isinstance(x, torch.Tensor):
        return x
    if is_tf_tensor(x):
        import tensorflow as tf  # only reachable when the caller already uses TensorFlow

        from .tf_adapter import _sync_tf

        _sync_tf(tf, x)
    if type(x).__name__ == "PyCapsule":
        return torch.utils.dlpack.from_dlpack(x)
    if hasattr(x, "__dlpack__"):
        return torch.from_dlpack(x)
    raise TypeError(f"{name}: expected a torch.Tensor, a DLPack capsule or an object with __dlpack__, "
                    f"got {type(x).__name__}")


def require_cuda(t: torch.Tensor, dtype, *, name="tensor") -> torch.Tensor:
    """Validate device/dtype and return a C-contiguous tensor (a device-side copy if needed)."""
    if not t.is_cuda:
        raise ValueError(f"{name} must live on a CUDA device (got {t.device}); there is no CPU fallback")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype} (got {t.dtype})")
    return t if t.is_contiguous() else t.contiguous()


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def is_tf_tensor(x) -> bool:
    mod = type(x).__module__ or ""
    return mod.startswith("tensorflow")


def to_caller_framework(out: torch.Tensor, like):
    """Return `out` in the framework of `like` (TensorFlow eager tensor in -> TensorFlow tensor out)."""
    if is_tf_tensor(like):
        import tensorflow as tf  # only reachable when the caller already uses TensorFlow

        torch.cuda.current_stream(out.device).synchronize()  # TensorFlow reads on its own stream
        return tf.experimental.dlpack.from_dlpack(torch.utils.dlpack.to_dlpack(out))
    return out
