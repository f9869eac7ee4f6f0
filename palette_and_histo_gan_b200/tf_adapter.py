"""TensorFlow-side adapter: the generator-loss term of pix2pix_model.py:242-245 as a
`tf.custom_gradient` around the CUDA kernels, tensors handed over through DLPack without copies.

TensorFlow is NOT a dependency of this package and is not installable in the build image, so this module is
import-safe without it (TensorFlow is imported on first use) and is exercised only as far as that goes
(tests/test_abi.py); the torch path it wraps is what the GPU tests cover.  `tf.experimental.dlpack` works on eager
tensors only: call these from an eager `train_step` (as `Pix2PixIndexedModel.train_step` already is) or through
`tf.py_function` inside the reference's `@tf.function train_step` (pix2pix_model.py:62).

Stream ordering: TensorFlow runs its GPU kernels on its own compute stream and DLPack carries no stream.  Every
hand-over is therefore fenced: before a TensorFlow tensor is borrowed, TensorFlow's device work is drained
(`tf.test.experimental.sync_devices()`, a scalar read-back on versions without it), and before a result goes back,
torch's current stream is synchronised.  Two host synchronisations per call — the price of crossing frameworks
without a shared stream; UNTESTED in the build image (no TensorFlow), see INTEGRATION.md.
"""
from __future__ import annotations

import torch

from . import histogram as _h


def _tf():
    try:
        import tensorflow as tf
    except ImportError as exc:  # pragma: no cover - TensorFlow is absent in the build image
        raise ImportError("palette_and_histo_gan_b200.tf_adapter needs TensorFlow (>= 2.x, eager tensors with DLPack "
                          "support); the torch / DLPack entry points in `histogram` work without it") from exc
    return tf


def _sync_tf(tf, t):
    """All TensorFlow GPU work that produces `t` has finished when this returns."""
    sync = getattr(getattr(tf.test, "experimental", None), "sync_devices", None)
    if sync is not None:
        sync()
    else:  # TF < 2.11: a host read-back of one element orders TensorFlow's stream
        tf.reshape(t, [-1])[:1].numpy()


def _to_torch(t):
    tf = _tf()
    _sync_tf(tf, t)  # the kernels of libpalhist run on torch's stream: TensorFlow must be done writing `t`
    return torch.utils.dlpack.from_dlpack(tf.experimental.dlpack.to_dlpack(t))


def _to_tf(t):
    tf = _tf()
    t = t.contiguous()
    torch.cuda.current_stream(t.device).synchronize()  # TensorFlow must not read before our kernels are done
    return tf.experimental.dlpack.from_dlpack(torch.utils.dlpack.to_dlpack(t))


def histogram_loss(real_image, fake_image, size=64, method="inverse-quadratic", sigma=0.02):
    """`hellinger_loss(calculate_rgbuv_histogram(real), calculate_rgbuv_histogram(fake))` for eager TensorFlow GPU
    tensors in [-1, 1], differentiable with respect to `fake_image` under `tf.GradientTape`."""
    tf = _tf()

    @tf.custom_gradient
    def _loss(real, fake):
        f = _to_torch(fake).requires_grad_(True)
        loss = _h.histogram_loss(_to_torch(real), f, size=size, method=method, sigma=sigma)

        def grad(upstream):
            (g,) = torch.autograd.grad(loss, f, _to_torch(tf.reshape(upstream, [])).to(loss.dtype))
            return None, _to_tf(g)

        return _to_tf(loss.detach().reshape(1))[0], grad

    return _loss(real_image, fake_image)


def calculate_rgbuv_histogram(image_batch, size=64, method="inverse-quadratic", sigma=0.02):
    """histogram.py:36-81 for an eager TensorFlow GPU tensor, differentiable under `tf.GradientTape`."""
    tf = _tf()

    @tf.custom_gradient
    def _hist(image):
        x = _to_torch(image).requires_grad_(True)
        hist = _h.calculate_rgbuv_histogram(x, size=size, method=method, sigma=sigma)

        def grad(upstream):
            (g,) = torch.autograd.grad(hist, x, _to_torch(upstream))
            return _to_tf(g)

        return _to_tf(hist.detach()), grad

    return _hist(image_batch)
