"""Host-buffer calls: numpy arrays (or pinned CPU torch tensors) in, numpy arrays out, GPU in between.

These bind the `ph_host_*` entry points of libpalhist — the functions a CPU-side caller of the
reference (a `tf.data` map worker, `dataset_utils.py:236-244`; an eager `generator_loss`,
`pix2pix_model.py:242-250`) would call with host memory.  Host<->device copies, chunked and overlapped
with the kernels, happen inside the call.  This is the path bench.py times as `e2e`.
"""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np

from . import _lib
from .histogram import EPSILON, _method_id, _sigma_sqr, tf_linspace

_tls = threading.local()


class HostContext:
    """Per-thread device staging (streams, events, grow-only arena). Not shareable across threads."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        _lib.call("ph_host_ctx_create", int(device), C.byref(self._h))
        self.device = int(device)

    def close(self):
        if self._h:
            _lib.load().ph_host_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def default_context(device: int = 0) -> HostContext:
    ctxs = getattr(_tls, "ctxs", None)
    if ctxs is None:
        ctxs = _tls.ctxs = {}
    if device not in ctxs:
        ctxs[device] = HostContext(device)
    return ctxs[device]


def _np(x, dtype, name):
    if hasattr(x, "numpy") and not isinstance(x, np.ndarray):  # CPU torch tensor (possibly pinned)
        x = x.numpy()
    a = np.asarray(x)
    if a.dtype != dtype:
        raise TypeError(f"{name} must be {np.dtype(dtype).name}, got {a.dtype}")
    if not a.flags.c_contiguous:
        a = np.ascontiguousarray(a)
    return a


def histogram_loss(real_image, fake_image, size=64, method="inverse-quadratic", sigma=0.02, *, impl="auto",
                   want_grad=True, out_grad=None, ctx=None, device=0):
    """Loss and d loss/d fake of pix2pix_model.py:243-245 for host images (B,H,W,3|4) float32.
    Returns (loss: float, grad: np.ndarray | None)."""
    real = _np(real_image, np.float32, "real_image")
    fake = _np(fake_image, np.float32, "fake_image")
    if real.shape != fake.shape or real.ndim != 4 or real.shape[-1] not in (3, 4):
        raise ValueError("real_image and fake_image must both be (B,H,W,3|4)")
    b, h, w, ch = real.shape
    dom = tf_linspace(-3.0, 3.0, int(size))
    loss = np.zeros((1,), np.float32)
    grad = None
    if want_grad:
        grad = out_grad if out_grad is not None else np.empty_like(fake)
        grad = _np(grad, np.float32, "out_grad")
    ctx = ctx or default_context(device)
    _lib.call("ph_host_hist_loss", ctx._h, real.ctypes.data, fake.ctypes.data, b, h * w, ch, dom.ctypes.data,
              int(size), _method_id(method), _sigma_sqr(sigma), EPSILON, _lib.IMPLS[impl], loss.ctypes.data,
              grad.ctypes.data if grad is not None else None)
    return float(loss[0]), grad


def histogram_loss_begin(real_image, fake_image, size=64, method="inverse-quadratic", sigma=0.02, *, impl="auto",
                         ctx=None, device=0) -> float:
    """Phase 1 of a sharded evaluation: upload, both forward passes and the backward kernels at unit scale
    (the gradient depends on the global sum only through the factor 1 / (B sqrt(S)), applied in phase 2);
    returns this shard's sum of squares (all-reduce it over ranks, then call `histogram_loss_finish`)."""
    fake = _np(fake_image, np.float32, "fake_image")
    real = real_image.numpy() if hasattr(real_image, "numpy") and not isinstance(real_image, np.ndarray) else np.asarray(real_image)
    dom = tf_linspace(-3.0, 3.0, int(size))
    ssum = C.c_double(0.0)
    ctx = ctx or default_context(device)
    if real.dtype == np.uint8:
        # sprites straight from the PNG decoder: blacken + normalise happen on the device (dataset_utils.py:66-77)
        real = np.ascontiguousarray(real)
        if real.shape != fake.shape or real.ndim != 4 or real.shape[-1] != 4:
            raise ValueError("uint8 real_image and float32 fake_image must both be (B,H,W,4)")
        b, h, w, _ = real.shape
        _lib.call("ph_host_hist_begin_u8real", ctx._h, real.ctypes.data, fake.ctypes.data, b, h * w, dom.ctypes.data,
                  int(size), _method_id(method), _sigma_sqr(sigma), EPSILON, _lib.IMPLS[impl], C.byref(ssum))
        return float(ssum.value)
    real = _np(real, np.float32, "real_image")
    if real.shape != fake.shape or real.ndim != 4 or real.shape[-1] not in (3, 4):
        raise ValueError("real_image and fake_image must both be (B,H,W,3|4)")
    b, h, w, ch = real.shape
    _lib.call("ph_host_hist_begin", ctx._h, real.ctypes.data, fake.ctypes.data, b, h * w, ch, dom.ctypes.data,
              int(size), _method_id(method), _sigma_sqr(sigma), EPSILON, _lib.IMPLS[impl], C.byref(ssum))
    return float(ssum.value)


def histogram_loss_finish(ssum_global: float, global_batch: int, out_grad=None, *, out_grad_device=None, ctx=None,
                          device=0):
    """Phase 2: loss of the whole batch and the gradient of this shard, downloaded into `out_grad` (host
    float32 array shaped like the shard's fake images) and/or left in `out_grad_device` (a CUDA float32
    tensor of that shape: the generator's backward consumes it on the device)."""
    loss = np.zeros((1,), np.float32)
    grad = _np(out_grad, np.float32, "out_grad") if out_grad is not None else None
    dptr = None
    if out_grad_device is not None:
        if not (out_grad_device.is_cuda and out_grad_device.is_contiguous() and out_grad_device.dtype.is_floating_point
                and out_grad_device.element_size() == 4):
            raise ValueError("out_grad_device must be a contiguous float32 CUDA tensor")
        dptr = out_grad_device.data_ptr()
    ctx = ctx or default_context(device)
    _lib.call("ph_host_hist_finish", ctx._h, float(ssum_global), int(global_batch), loss.ctypes.data,
              grad.ctypes.data if grad is not None else None, dptr)
    return float(loss[0]), grad


def histogram_loss_finish_comm(comm, global_batch: int, out_grad=None, *, out_grad_device=None, ctx=None, device=0):
    """Phase 2 for ranks connected by a `_comm.PeerComm`: the shard's sum of squares that `histogram_loss_begin`
    left on the device is summed over the ranks by one kernel over peer memory (NVLink) — no host round trip and
    no library collective between the phases."""
    loss = np.zeros((1,), np.float32)
    grad = _np(out_grad, np.float32, "out_grad") if out_grad is not None else None
    dptr = None
    if out_grad_device is not None:
        if not (out_grad_device.is_cuda and out_grad_device.is_contiguous() and out_grad_device.dtype.is_floating_point
                and out_grad_device.element_size() == 4):
            raise ValueError("out_grad_device must be a contiguous float32 CUDA tensor")
        dptr = out_grad_device.data_ptr()
    ctx = ctx or default_context(device)
    _lib.call("ph_host_hist_finish_comm", ctx._h, comm.handle, int(global_batch), loss.ctypes.data,
              grad.ctypes.data if grad is not None else None, dptr)
    return float(loss[0]), grad


def histogram_loss_sharded(comm, real_image, fake_image, global_batch: int, size=64, method="inverse-quadratic",
                           sigma=0.02, *, impl="auto", out_grad=None, out_grad_device=None, ctx=None, device=0):
    """One call for this rank's shard of a batch spread over the ranks of `comm` (`_comm.PeerComm`): both phases
    with the sum over ranks taken on the device over peer memory in between — a single host synchronisation.
    `real_image`: float32 (B,H,W,3|4) or uint8 RGBA sprites.  Returns (loss, out_grad)."""
    fake = _np(fake_image, np.float32, "fake_image")
    real = real_image.numpy() if hasattr(real_image, "numpy") and not isinstance(real_image, np.ndarray) else np.asarray(real_image)
    is_u8 = real.dtype == np.uint8
    real = np.ascontiguousarray(real) if is_u8 else _np(real, np.float32, "real_image")
    if real.shape != fake.shape or real.ndim != 4 or real.shape[-1] not in ((4,) if is_u8 else (3, 4)):
        raise ValueError("real_image and fake_image must both be (B,H,W,3|4) (uint8 real images: RGBA)")
    b, h, w, ch = fake.shape
    dom = tf_linspace(-3.0, 3.0, int(size))
    loss = np.zeros((1,), np.float32)
    grad = _np(out_grad, np.float32, "out_grad") if out_grad is not None else None
    dptr = None
    if out_grad_device is not None:
        if not (out_grad_device.is_cuda and out_grad_device.is_contiguous() and out_grad_device.dtype.is_floating_point
                and out_grad_device.element_size() == 4):
            raise ValueError("out_grad_device must be a contiguous float32 CUDA tensor")
        dptr = out_grad_device.data_ptr()
    ctx = ctx or default_context(device)
    _lib.call("ph_host_hist_loss_sharded", ctx._h, comm.handle, real.ctypes.data, 1 if is_u8 else 0, fake.ctypes.data, b,
              h * w, ch, dom.ctypes.data, int(size), _method_id(method), _sigma_sqr(sigma), EPSILON, _lib.IMPLS[impl],
              int(global_batch), loss.ctypes.data, grad.ctypes.data if grad is not None else None, dptr)
    return float(loss[0]), grad


def load_indexed_images(source_image, target_image, palette_ordering="grayness", *, with_one_hot=False,
                        out=None, ctx=None, device=0, seed=None):
    """dataset_utils.py:138-151 for host images (B,H,W,4), int32 (values 0..255) or uint8 (the decoded PNG as it
    is: a quarter of the upload, widened on the device).  Returns (source_indexed, target_indexed, palette
    [, target_one_hot]) as numpy arrays.  `out=(source_indexed, target_indexed, palette)`: caller-owned int32 result
    buffers of shapes (B,H,W,1), (B,H,W,1), (B,256,4) — numpy arrays or pinned CPU torch tensors (page-locked
    results download at PCIe speed; fresh pageable arrays take a staged copy).  `palette_ordering="shuffled"`
    (io_utils.py:56-58) permutes each palette with numpy's `default_rng(seed)`."""
    from .io_utils import PaletteOverflowError, _ordering_id
    from .configuration import MAX_PALETTE_SIZE

    def _is_u8(x):
        return str(getattr(x, "dtype", "")).endswith("uint8")  # numpy or (pinned) CPU torch tensor

    as_u8 = _is_u8(source_image) and _is_u8(target_image)
    dt = np.uint8 if as_u8 else np.int32
    src = _np(source_image, dt, "source_image")
    tgt = _np(target_image, dt, "target_image")
    if src.shape != tgt.shape or src.ndim != 4 or src.shape[-1] != 4:
        raise ValueError("source_image and target_image must both be (B,H,W,4)")
    b, h, w, _ = src.shape
    if out is not None:
        s_idx, t_idx, pal = (_np(o, np.int32, "out") for o in out)
        if s_idx.shape != (b, h, w, 1) or t_idx.shape != (b, h, w, 1) or pal.shape != (b, MAX_PALETTE_SIZE, 4):
            raise ValueError("out buffers must have shapes (B,H,W,1), (B,H,W,1), (B,256,4)")
        if not (s_idx.flags.writeable and t_idx.flags.writeable and pal.flags.writeable):
            raise ValueError("out buffers must be writeable")
    else:
        s_idx = np.empty((b, h, w, 1), np.int32)
        t_idx = np.empty((b, h, w, 1), np.int32)
        pal = np.empty((b, MAX_PALETTE_SIZE, 4), np.int32)
    nc = np.empty((b,), np.int32)
    oh = np.empty((b, h, w, MAX_PALETTE_SIZE), np.float32) if with_one_hot else None
    order = _ordering_id(palette_ordering)
    keys = None
    if palette_ordering == "shuffled":  # independent uniform keys; the kernel ranks the first n of each row
        keys = np.random.default_rng(seed).random((b, MAX_PALETTE_SIZE), dtype=np.float32)
    ctx = ctx or default_context(device)
    _lib.call("ph_host_load_indexed_images_u8" if as_u8 else "ph_host_load_indexed_images", ctx._h, src.ctypes.data,
              tgt.ctypes.data, b, h * w, order, keys.ctypes.data if keys is not None else None, s_idx.ctypes.data,
              t_idx.ctypes.data, pal.ctypes.data, nc.ctypes.data, oh.ctypes.data if oh is not None else None)
    if (nc == _lib.PALETTE_BAD_VALUE).any():
        raise ValueError("colour values must lie in [0, 255]")
    if (nc > MAX_PALETTE_SIZE).any():
        raise PaletteOverflowError(f"more than {MAX_PALETTE_SIZE} unique colours")
    if with_one_hot:
        return s_idx, t_idx, pal, oh
    return s_idx, t_idx, pal
