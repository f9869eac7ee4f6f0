"""The TensorFlow leg of the boundary (north_star: "hands TensorFlow tensors through DLPack"): runs only where
TensorFlow with GPU support is installed (it is not in the build image — skipped there)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
tf = pytest.importorskip("tensorflow")

from oracle import histogram_oracle as ho  # noqa: E402


def test_histogram_loss_under_gradient_tape(cuda):
    from palette_and_histo_gan_b200 import tf_adapter

    if not tf.config.list_physical_devices("GPU"):
        pytest.skip("TensorFlow sees no GPU")
    rng = np.random.default_rng(5)
    real = np.tanh(rng.standard_normal((4, 32, 32, 4))).astype(np.float32)
    fake = np.tanh(rng.standard_normal((4, 32, 32, 4))).astype(np.float32)
    ref = ho.hist_loss_and_grad_f64(real, fake)
    with tf.device("/GPU:0"):
        r, f = tf.constant(real), tf.Variable(fake)
        with tf.GradientTape() as tape:
            loss = tf_adapter.histogram_loss(r, f)
        grad = tape.gradient(loss, f)
    assert abs(float(loss) - ref["loss"]) / ref["loss"] < 1e-5
    assert ho.rel_l2(grad.numpy(), ref["grad"]) < 1e-5
    hist = tf_adapter.calculate_rgbuv_histogram(r)
    assert ho.rel_l2(hist.numpy(), ref["hist_real"]) < 1e-5
