"""Drop-in for the reference's `histogram.py` — same callables, same argument meaning, CUDA underneath.

    calculate_component_histogram   histogram.py:5-32
    calculate_rgbuv_histogram       histogram.py:36-81
    hellinger_loss                  histogram.py:84-89
    l1_loss / l2_loss               histogram.py:92-97

plus `histogram_loss(real, fake)`: the three lines of `Pix2PixHistogramModel.generator_loss`
(pix2pix_model.py:243-245) as one differentiable call.

Inputs are CUDA tensors of any DLPack-speaking framework (see `_tensor.py`); outputs are torch CUDA
tensors (TensorFlow tensors when the input was one).  Gradients flow through torch autograd:
`calculate_rgbuv_histogram` and `hellinger_loss` are `autograd.Function`s whose backward passes are
the analytic kernels of libpalhist (the reference relies on TF autodiff, pix2pix_model.py:78).
When the batch is sharded over ranks (`group=`), the only exchange is the sum over ranks of the scalar
sum of squares that the Hellinger distance takes over the whole batch (histogram.py:88-89): one 32-thread
kernel over NVLink peer memory (`_comm.py`; an NCCL all-reduce where the ranks cannot map each other's memory).
"""
from __future__ import annotations


import numpy as np
import torch

from . import _lib
from ._tensor import from_any, ptr, require_cuda, stream_ptr, to_caller_framework

EPSILON = 1e-6  # histogram.py:53

_dom_cache: dict = {}


def tf_linspace(start: float, stop: float, num: int) -> np.ndarray:
    """float32 values of `tf.linspace(start, stop, num)` (histogram.py:55): start, start+delta*i, stop."""
    start32, stop32 = np.float32(start), np.float32(stop)
    if num == 1:
        return np.array([start32], dtype=np.float32)
    delta = np.float32((stop32 - start32) / np.float32(num - 1))
    mid = (start32 + delta * np.arange(1, num - 1, dtype=np.float32)).astype(np.float32)
    return np.concatenate([[start32], mid, [stop32]]).astype(np.float32)


def histogram_domain(size: int, device) -> torch.Tensor:
    """(size,) float32 bin centres on `device`, cached."""
    key = (int(size), str(device))
    dom = _dom_cache.get(key)
    if dom is None:
        dom = torch.from_numpy(tf_linspace(-3.0, 3.0, int(size))).to(device)
        _dom_cache[key] = dom
    return dom


MIRROR_FLAG = 16  # PH_IMPL_MIRROR


def _mirror_flag(size: int, sigma, mirror: bool) -> int:
    """PH_IMPL_MIRROR when the bin centres are antisymmetric to within 2e-5 sigma (palhist.h) — `tf.linspace(-3, 3,
    64)` is, to 3.6e-7 — so that dense 64-bin batches run the mirrored-tile forward."""
    if not mirror:
        return 0
    dom = tf_linspace(-3.0, 3.0, int(size)).astype(np.float64)
    return MIRROR_FLAG if float(np.abs(dom + dom[::-1]).max()) <= 2e-5 * float(np.float32(sigma)) else 0


def _method_id(method) -> int:
    try:
        return _lib.METHODS[method]
    except KeyError:
        # the reference silently returns garbage for any other string (histogram.py:22-27)
        raise ValueError(f"method must be 'inverse-quadratic' or 'RBF', got {method!r}") from None


def _sigma_sqr(sigma) -> float:
    s = np.float32(sigma)  # tf.pow(sigma, 2) on a python float is evaluated in float32 (histogram.py:54)
    return float(np.float32(s * s))


def _workspace(batch, npix, bins, impl, device):
    n = _lib.load().ph_hist_workspace_bytes(batch, npix, bins, impl)
    return torch.empty(int(n), dtype=torch.uint8, device=device), int(n)


def _check_image(image):
    if image.dim() != 4 or image.shape[-1] not in (3, 4):
        raise ValueError(f"image_batch must be (batch, H, W, 3|4), got {tuple(image.shape)}")
    return image.shape[0], image.shape[1] * image.shape[2], image.shape[3]


def _forward(image, dom, method_id, sigma_sqr, impl):
    """-> (hist (B,S,S,3), denom (B,)) on image.device; image is contiguous float32 CUDA."""
    b, npix, ch = _check_image(image)
    bins = dom.numel()
    hist = torch.empty((b, bins, bins, 3), dtype=torch.float32, device=image.device)
    denom = torch.empty((b,), dtype=torch.float32, device=image.device)
    if b == 0:
        return hist, denom
    ws, ws_bytes = _workspace(b, npix, bins, impl, image.device)
    with torch.cuda.device(image.device):
        _lib.call("ph_hist_forward", ptr(image), b, npix, ch, ptr(dom), bins, method_id, sigma_sqr, EPSILON,
                  ptr(hist), ptr(denom), ptr(ws), ws_bytes, impl, stream_ptr(image.device))
    return hist, denom


def _backward(image, dom, method_id, sigma_sqr, impl, hist_pred, denom, *, grad_hist=None, hist_true=None,
              ssum=None, global_batch=0, loss_scale=None):
    b, npix, ch = _check_image(image)
    bins = dom.numel()
    grad = torch.empty_like(image)
    if b == 0:
        return grad
    ws, ws_bytes = _workspace(b, npix, bins, impl, image.device)
    with torch.cuda.device(image.device):
        _lib.call("ph_hist_backward", ptr(image), b, npix, ch, ptr(dom), bins, method_id, sigma_sqr, EPSILON,
                  ptr(hist_pred), ptr(denom), ptr(grad_hist), ptr(hist_true), ptr(ssum), int(global_batch),
                  ptr(loss_scale), ptr(grad), ptr(ws), ws_bytes, impl, stream_ptr(image.device))
    return grad


def _forward_ssum(image, dom, method_id, sigma_sqr, impl, hist_true):
    """Forward of the `fake` images fused with their share of the Hellinger sum of squares against `hist_true`
    (one launch, the normalised histogram is read from shared memory): -> (hist, denom, ssum (1,) float64)."""
    b, npix, ch = _check_image(image)
    bins = dom.numel()
    hist = torch.empty((b, bins, bins, 3), dtype=torch.float32, device=image.device)
    denom = torch.empty((b,), dtype=torch.float32, device=image.device)
    ssum = torch.empty((1,), dtype=torch.float64, device=image.device)
    ws, ws_bytes = _workspace(max(b, 1), npix, bins, impl, image.device)
    with torch.cuda.device(image.device):
        _lib.call("ph_hist_forward_ssum", ptr(image), b, npix, ch, ptr(dom), bins, method_id, sigma_sqr, EPSILON,
                  ptr(hist), ptr(denom), ptr(hist_true), ptr(ssum), 0, ptr(ws), ws_bytes, impl,
                  stream_ptr(image.device))
    return hist, denom, ssum


def _ssum(y_true, y_pred):
    out = torch.empty((1,), dtype=torch.float64, device=y_true.device)
    with torch.cuda.device(y_true.device):
        _lib.call("ph_hellinger_ssum", ptr(y_true), ptr(y_pred), y_true.numel(), ptr(out), stream_ptr(y_true.device))
    return out


def _finish(ssum, global_batch):
    loss = torch.empty((), dtype=torch.float32, device=ssum.device)
    with torch.cuda.device(ssum.device):
        _lib.call("ph_hellinger_finish", ptr(ssum), int(global_batch), ptr(loss), stream_ptr(ssum.device))
    return loss


def _reduce_over_ranks(ssum, local_batch, group, global_batch):
    """Sum the one scalar that couples the shards over the ranks (in place); returns the whole-batch size.
    On CUDA tensors of an NCCL group this is the peer-memory kernel of `_comm.py`; otherwise (gloo on the CPU in
    the tests, ranks without P2P access) `torch.distributed.all_reduce`.  Without `global_batch` the shards are
    taken to be equal (pass it when they are not: the last partial batch of an epoch)."""
    if group is None or group is False:
        return local_batch if global_batch is None else int(global_batch)
    import torch.distributed as dist

    pg = None if group is True else group
    comm = None
    if ssum.is_cuda:
        from ._comm import peer_comm

        comm = peer_comm(group, ssum.device)
    if comm is not None:
        comm.allreduce_(ssum)
    else:
        dist.all_reduce(ssum, op=dist.ReduceOp.SUM, group=pg)
    if global_batch is not None:
        return int(global_batch)
    return local_batch * dist.get_world_size(pg)  # equal shards


class _RgbuvHistogramFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, dom, method_id, sigma_sqr, impl):
        hist, denom = _forward(image, dom, method_id, sigma_sqr, impl)
        ctx.save_for_backward(image, dom, hist, denom)
        ctx.conf = (method_id, sigma_sqr, impl)
        return hist

    @staticmethod
    def backward(ctx, grad_hist):
        image, dom, hist, denom = ctx.saved_tensors
        method_id, sigma_sqr, impl = ctx.conf
        grad_hist = require_cuda(grad_hist, torch.float32, name="grad_hist")
        grad = _backward(image, dom, method_id, sigma_sqr, impl, hist, denom, grad_hist=grad_hist)
        return grad, None, None, None, None


class _HellingerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y_true, y_pred, group, global_batch):
        ssum = _ssum(y_true, y_pred)
        gb = _reduce_over_ranks(ssum, y_true.shape[0], group, global_batch)
        ctx.save_for_backward(y_true, y_pred, ssum)
        ctx.gb = gb
        return _finish(ssum, gb)

    @staticmethod
    def backward(ctx, grad_loss):
        y_true, y_pred, ssum = ctx.saved_tensors
        need_true, need_pred = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        g_true = torch.empty_like(y_true) if need_true else None
        g_pred = torch.empty_like(y_pred) if need_pred else None
        scale = grad_loss.to(torch.float32).contiguous()
        with torch.cuda.device(y_true.device):
            _lib.call("ph_hellinger_backward", ptr(y_true), ptr(y_pred), y_true.numel(), ptr(ssum), int(ctx.gb),
                      ptr(scale), ptr(g_true), ptr(g_pred), stream_ptr(y_true.device))
        return g_true, g_pred, None, None


class _HistogramLossFn(torch.autograd.Function):
    """fwd(real) + fwd(fake) + Hellinger, backward to the fake image only (pix2pix_model.py:243-245).

    Sharded batches: the only exchange is the sum over ranks of the sum of squares S, enqueued on the same stream
    right behind the forward kernel that produced this rank's share (a ~3 us peer-memory kernel), so the backward
    kernels read the whole-batch S like the single-device path does — same arithmetic, no correction pass."""

    @staticmethod
    def forward(ctx, real, fake, dom, method_id, sigma_sqr, impl, group, global_batch, dedup_real):
        hist_real, _ = _forward(real, dom, method_id, sigma_sqr, impl | (DEDUP_FLAG if dedup_real else 0))
        hist_fake, denom_fake, ssum = _forward_ssum(fake, dom, method_id, sigma_sqr, impl, hist_real)
        gb = _reduce_over_ranks(ssum, real.shape[0], group, global_batch)
        ctx.save_for_backward(fake, dom, hist_real, hist_fake, denom_fake, ssum)
        ctx.conf = (method_id, sigma_sqr, impl, gb)
        return _finish(ssum, gb)

    @staticmethod
    def backward(ctx, grad_loss):
        fake, dom, hist_real, hist_fake, denom_fake, ssum = ctx.saved_tensors
        method_id, sigma_sqr, impl, gb = ctx.conf
        scale = grad_loss.to(torch.float32).contiguous()
        grad = _backward(fake, dom, method_id, sigma_sqr, impl, hist_fake, denom_fake, hist_true=hist_real,
                         ssum=ssum, global_batch=gb, loss_scale=scale)
        return None, grad, None, None, None, None, None, None, None


# ------------------------------------------------------------------------------------------------
# public API — reference signatures
# ------------------------------------------------------------------------------------------------
DEDUP_FLAG = 8  # PH_IMPL_DEDUP


def _range_checked(run, impl, device, range_check):
    """`range_check="sync"`: evaluate, wait for the kernels, and if the tensor-core forward flagged pixels outside
    its operand range (images far outside [-1, 1], which the reference accepts, histogram.py:58) evaluate again on
    the CUDA-core engine.  `"async"` (default, no synchronisation): such a launch sets the sticky status word
    (`_lib.async_status`) and the next histogram call raises."""
    if range_check not in ("async", "sync"):
        raise ValueError("range_check must be 'async' or 'sync'")
    if range_check == "async" or impl == "simt":
        return run(impl)
    with torch.cuda.device(device):
        torch.cuda.current_stream(device).synchronize()
        _lib.async_status(clear=True)
        out = run(impl)
        torch.cuda.current_stream(device).synchronize()
        if _lib.async_status(clear=True) & _lib.ASYNC_RANGE:
            out = run("simt")
    return out


def calculate_rgbuv_histogram(image_batch, size=64, method="inverse-quadratic", sigma=0.02, *, impl="auto",
                              dedup=False, range_check="async", mirror=True):
    """histogram.py:36-81.  image_batch (B,H,W,3|4) float32 in [-1,1] -> (B,size,size,3), sums to 1 per image.
    `dedup=True`: contract each image's unique colours with their multiplicities (exact; pays off for
    palette images such as dataset sprites, falls back to the dense contraction per image otherwise).
    `range_check`: see `_range_checked` (images outside [-1, 1] on the tensor-core engine).
    `mirror`: dense 64-bin batches share the weight vectors of +x and -x (`_mirror_flag`); False = six vectors per
    pixel around the exact centres."""
    image = require_cuda(from_any(image_batch, name="image_batch"), torch.float32, name="image_batch")
    dom = histogram_domain(size, image.device)
    mid, s2 = _method_id(method), _sigma_sqr(sigma)
    mflag = _mirror_flag(size, sigma, mirror)
    out = _range_checked(lambda eng: _RgbuvHistogramFn.apply(image, dom, mid, s2,
                                                             _lib.IMPLS[eng] | (DEDUP_FLAG if dedup else 0) | mflag),
                         impl, image.device, range_check)
    return to_caller_framework(out, image_batch)


def calculate_component_histogram(component, projection1, projection2, color_intensities, histogram_domain,
                                  method, sigma_sqr, epsilon):
    """histogram.py:5-32.  component/projection* (B,HW); color_intensities (B,HW,1); histogram_domain
    (1,size) -> un-normalised (B,size,size).  Forward only (the differentiable entry is
    `calculate_rgbuv_histogram`)."""
    comp = require_cuda(from_any(component), torch.float32, name="component")
    p1 = require_cuda(from_any(projection1), torch.float32, name="projection1")
    p2 = require_cuda(from_any(projection2), torch.float32, name="projection2")
    inten = require_cuda(from_any(color_intensities), torch.float32, name="color_intensities")
    dom = require_cuda(from_any(histogram_domain), torch.float32, name="histogram_domain").reshape(-1)
    if comp.dim() != 2 or p1.shape != comp.shape or p2.shape != comp.shape or inten.numel() != comp.numel():
        raise ValueError("component/projection1/projection2 must be (batch, HW) and color_intensities (batch, HW, 1)")
    b, npix = comp.shape
    bins = dom.numel()
    out = torch.empty((b, bins, bins), dtype=torch.float32, device=comp.device)
    with torch.cuda.device(comp.device):
        _lib.call("ph_component_histogram", ptr(comp), ptr(p1), ptr(p2), ptr(inten), b, npix, ptr(dom), bins,
                  _method_id(method), float(sigma_sqr), float(epsilon), ptr(out), stream_ptr(comp.device))
    return to_caller_framework(out, component)


def hellinger_loss(y_true, y_pred, *, group=None, global_batch=None):
    """histogram.py:84-89: (1/sqrt 2)·sqrt(sum (sqrt(y_pred)-sqrt(y_true))^2) / batch, one sqrt over the whole
    batch.  `group=True` (default process group) or a ProcessGroup: the batch is sharded over ranks."""
    t = require_cuda(from_any(y_true, name="y_true"), torch.float32, name="y_true")
    p = require_cuda(from_any(y_pred, name="y_pred"), torch.float32, name="y_pred")
    if t.shape != p.shape:
        raise ValueError(f"y_true {tuple(t.shape)} and y_pred {tuple(p.shape)} must have the same shape")
    return to_caller_framework(_HellingerFn.apply(t, p, group, global_batch), y_true)


def _mean_diff(y_true, y_pred, kind):
    t = require_cuda(from_any(y_true), torch.float32, name="y_true")
    p = require_cuda(from_any(y_pred), torch.float32, name="y_pred")
    if t.shape != p.shape:
        raise ValueError("y_true and y_pred must have the same shape")
    out = torch.empty((), dtype=torch.float32, device=t.device)
    with torch.cuda.device(t.device):
        _lib.call("ph_mean_abs_or_sq_diff", ptr(t), ptr(p), t.numel(), kind, ptr(out), stream_ptr(t.device))
    return to_caller_framework(out, y_true)


def l1_loss(y_true, y_pred):
    """histogram.py:92-93 (forward only; unused by the reference's models)."""
    return _mean_diff(y_true, y_pred, 1)


def l2_loss(y_true, y_pred):
    """histogram.py:96-97 (forward only; unused by the reference's models)."""
    return _mean_diff(y_true, y_pred, 2)


def histogram_loss(real_image, fake_image, size=64, method="inverse-quadratic", sigma=0.02, *, group=None,
                   global_batch=None, impl="auto", dedup_real=True, range_check="async", mirror=True):
    """`hellinger_loss(calculate_rgbuv_histogram(real), calculate_rgbuv_histogram(fake))` as one call
    (pix2pix_model.py:243-245); differentiable with respect to `fake_image`.  `dedup_real`: the real images
    come from the dataset and are palette sprites, so their histogram is contracted over unique colours
    (images that are not palette-like are detected on the device and contracted densely).  `mirror`: the fake
    images' forward shares the weight vectors of +x and -x (`calculate_rgbuv_histogram`)."""
    real = require_cuda(from_any(real_image, name="real_image"), torch.float32, name="real_image")
    fake = require_cuda(from_any(fake_image, name="fake_image"), torch.float32, name="fake_image")
    if real.shape != fake.shape:
        raise ValueError("real_image and fake_image must have the same shape")
    if range_check == "sync" and group is not None and group is not False:
        raise ValueError("range_check='sync' re-runs a flagged call on this rank alone; with a sharded batch check "
                         "`_lib.async_status()` after the step instead")
    dom = histogram_domain(size, fake.device)
    mid, s2 = _method_id(method), _sigma_sqr(sigma)
    mflag = _mirror_flag(size, sigma, mirror)
    out = _range_checked(lambda eng: _HistogramLossFn.apply(real, fake, dom, mid, s2, _lib.IMPLS[eng] | mflag, group,
                                                            global_batch, bool(dedup_real)),
                         impl, fake.device, range_check)
    return to_caller_framework(out, fake_image)
