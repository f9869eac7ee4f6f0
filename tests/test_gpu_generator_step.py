"""The caller of the hot path (BASELINE.json config 4): a pix2pix "histogram" model step
whose generator loss calls the new kernels (pix2pix_model.py:62-78, 242-250)."""
import numpy as np
import pytest
import torch

from oracle import histogram_oracle as ho

pytestmark = pytest.mark.gpu


def _normalise(u8):
    a = u8.astype(np.float32)
    a = np.where(a[..., 3:4] == 0, 0.0, a)  # blacken_transparent_pixels, dataset_utils.py:11-20
    return (a / 127.5 - 1.0).astype(np.float32)


@pytest.fixture(scope="module")
def G():
    from palette_and_histo_gan_b200 import generator_step

    return generator_step


def test_network_sizes_match_the_reference(G, cuda):
    step = G.Pix2PixHistogramStep(cuda)
    n_gen = sum(p.numel() for p in step.generator.parameters())
    n_disc = sum(p.numel() for p in step.discriminator.parameters())
    # U-Net generator 29.3 M parameters; PatchGAN: 4x4x8x64 + 4x4x64x1 + 1
    assert 29.2e6 < n_gen < 29.4e6
    assert n_disc == 4 * 4 * 8 * 64 + 4 * 4 * 64 + 1
    x = torch.zeros(2, 64, 64, 4, device=cuda)
    assert tuple(step.generator(x).shape) == (2, 64, 64, 4)
    assert tuple(step.discriminator(x, x).shape) == (2, 1, 32, 32)


def test_histogram_term_inside_the_generator_loss(G, cuda, sprites):
    """The histogram term of the generator loss equals the oracle on the generator's own output, and its
    gradient reaches the generator's parameters identically through both engines."""
    src = torch.from_numpy(_normalise(sprites["front"][:6])).to(cuda)
    real = torch.from_numpy(_normalise(sprites["right"][:6])).to(cuda)
    grads = {}
    torch.backends.cudnn.allow_tf32 = False  # fp32 convolutions: the comparison is about the loss kernels
    torch.backends.cuda.matmul.allow_tf32 = False
    for impl in ("tc", "simt"):
        step = G.Pix2PixHistogramStep(cuda, impl=impl, seed=3)
        torch.manual_seed(11)  # same dropout masks for both engines
        fake = step.generator(src)
        pred = step.discriminator(fake, src)
        total, adv, l1, hist = step.generator_loss(pred, fake, real)
        ref = ho.hist_loss_and_grad_f64(real.cpu().numpy(), fake.detach().cpu().numpy())
        assert abs(float(hist.detach()) - ref["loss"]) / ref["loss"] < 1e-5
        assert abs(float(total) - (float(adv) + 30.0 * float(l1) + float(hist))) < 1e-4 * abs(float(total))
        (g_fake,) = torch.autograd.grad(hist, fake, retain_graph=True)
        assert ho.rel_l2(g_fake.cpu().numpy(), ref["grad"]) < 1e-5
        hist.backward()
        grads[impl] = torch.cat([p.grad.flatten() for p in step.generator.parameters() if p.grad is not None])
    a, b = grads["tc"].double(), grads["simt"].double()
    rel = float((a - b).norm() / b.norm())
    assert rel < 1e-3, rel  # through 13 convolution layers; the loss gradients themselves agree to 1e-5 above


def test_train_step_updates_both_networks(G, cuda, sprites):
    src = torch.from_numpy(_normalise(sprites["front"][:8])).to(cuda)
    real = torch.from_numpy(_normalise(sprites["right"][:8])).to(cuda)
    step = G.Pix2PixHistogramStep(cuda)
    before_g = [p.detach().clone() for p in step.generator.parameters()]
    before_d = [p.detach().clone() for p in step.discriminator.parameters()]
    first = {k: float(v) for k, v in step.train_step(src, real).items()}
    assert all(np.isfinite(v) for v in first.values())
    assert first["histogram"] > 0 and first["l1"] > 0
    assert any(not torch.equal(a, b) for a, b in zip(before_g, step.generator.parameters()))
    assert any(not torch.equal(a, b) for a, b in zip(before_d, step.discriminator.parameters()))
    for _ in range(12):
        last = {k: float(v) for k, v in step.train_step(src, real).items()}
    assert all(np.isfinite(v) for v in last.values())
    assert last["l1"] < first["l1"]  # a dozen steps on one batch: the reconstruction term falls
