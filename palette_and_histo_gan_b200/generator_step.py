"""The caller of the histogram loss: one training step of the side2side "histogram" model
(`Pix2PixHistogramModel`, pix2pix_model.py:62-78 `train_step`, :44-57 / :242-250 losses, networks.py:8-98
networks), restated for torch so that the new kernels can be exercised where the reference uses them —
inside the generator loss of a pix2pix step (BASELINE.json config 4).

The networks are ordinary library layers (cuDNN convolutions through torch) and are not part of the hot
path; what this module adds to the path is the wiring the reference has at pix2pix_model.py:242-250:

    real_histogram = histogram.calculate_rgbuv_histogram(real_image)
    fake_histogram = histogram.calculate_rgbuv_histogram(fake_image)
    histogram_loss = histogram.hellinger_loss(real_histogram, fake_histogram)
    total_loss += lambda_histogram * histogram_loss

Images cross the boundary as the reference's NHWC float32 tensors in [-1, 1].

Data parallelism: the Hellinger loss takes one square root over the *global* batch, so every rank evaluates
the same scalar from the all-reduced sum of squares (`histogram_loss(..., group=True)`) and back-propagates
it into its own shard.  DistributedDataParallel *averages* parameter gradients over ranks, which is right for
the per-sample-mean losses (BCE, L1) but would divide the histogram term — a function of all shards — by the
world size; the step therefore scales that term by the world size before `backward()`.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import histogram

IMG_SIZE = 64  # configuration.py:27


def _init(m):
    # tf.random_normal_initializer(0., 0.02) on every kernel (networks.py:7, 24, 41, 55)
    if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
        nn.init.normal_(m.weight, 0.0, 0.02)
        if m.bias is not None:
            nn.init.zeros_(m.bias)


def _instance_norm(channels):
    # tfa InstanceNormalization (networks.py:19, 31): per-sample, per-channel statistics with learnable scale and
    # offset, epsilon 1e-3.  GroupNorm with one group per channel is the same operator and, unlike
    # nn.InstanceNorm2d, accepts the 1x1 bottleneck the reference normalises (output = offset there).
    return nn.GroupNorm(channels, channels, eps=1e-3, affine=True)


def _down(cin, cout, norm=True):
    layers = [nn.Conv2d(cin, cout, 4, stride=2, padding=1, bias=False)]  # k=4, s=2, "same"
    if norm:
        layers.append(_instance_norm(cout))
    layers.append(nn.LeakyReLU(0.3))  # keras LeakyReLU default alpha
    return nn.Sequential(*layers)


def _up(cin, cout, dropout=False):
    layers = [nn.ConvTranspose2d(cin, cout, 4, stride=2, padding=1, bias=False), _instance_norm(cout)]
    if dropout:
        layers.append(nn.Dropout(0.5))
    layers.append(nn.ReLU())
    return nn.Sequential(*layers)


class _SameConv4(nn.Module):
    """Conv2D(k=4, stride 1, padding="same"): TensorFlow pads 1 before and 2 after."""

    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, 4, stride=1, padding=0, bias=True)

    def forward(self, x):
        return self.conv(F.pad(x, (1, 2, 1, 2)))


class UnetGenerator(nn.Module):
    """networks.py:52-98: six stride-2 down blocks (64..512), six up blocks with skip concatenations (the last
    skip is the input itself), final same-padded 4x4 convolution + tanh.  NHWC in / NHWC out."""

    def __init__(self, input_channels=4, output_channels=4):
        super().__init__()
        self.down = nn.ModuleList([
            _down(input_channels, 64, norm=False), _down(64, 128), _down(128, 256), _down(256, 512),
            _down(512, 512), _down(512, 512)])
        self.up = nn.ModuleList([
            _up(512, 512, True), _up(1024, 512, True), _up(1024, 256, True), _up(512, 128), _up(256, 64),
            _up(128, 32)])
        self.last = _SameConv4(32 + input_channels, output_channels)
        self.apply(_init)

    def forward(self, image_nhwc):
        x = image_nhwc.permute(0, 3, 1, 2)
        inputs = x
        skips = []
        for d in self.down:
            x = d(x)
            skips.append(x)
        for u, skip in zip(self.up, list(reversed(skips[:-1])) + [inputs]):
            x = torch.cat([u(x), skip], dim=1)
        return torch.tanh(self.last(x)).permute(0, 2, 3, 1).contiguous()


class PatchDiscriminator(nn.Module):
    """networks.py:39-49: concat(target, source) -> one down block without normalisation -> 1-channel logits."""

    def __init__(self, input_channels=4):
        super().__init__()
        self.down = _down(2 * input_channels, 64, norm=False)
        self.last = _SameConv4(64, 1)
        self.apply(_init)

    def forward(self, target_nhwc, source_nhwc):
        x = torch.cat([target_nhwc, source_nhwc], dim=-1).permute(0, 3, 1, 2)
        return self.last(self.down(x))


class Pix2PixHistogramStep:
    """`Pix2PixHistogramModel.train_step` (pix2pix_model.py:62-78 with the losses of :44-57 and :242-250):
    Adam(2e-4, beta1 0.5) on both networks, BCE-from-logits adversarial terms, `lambda_l1` * L1 and
    `lambda_histogram` * Hellinger histogram loss on the generator (experiments.ipynb uses 30 and 1)."""

    def __init__(self, device, lambda_l1=30.0, lambda_histogram=1.0, *, distributed=False, impl="auto", seed=47):
        torch.manual_seed(seed)
        self.device = torch.device(device)
        self.generator = UnetGenerator().to(self.device)
        self.discriminator = PatchDiscriminator().to(self.device)
        self.lambda_l1, self.lambda_histogram, self.impl = float(lambda_l1), float(lambda_histogram), impl
        self.distributed = bool(distributed)
        self.world = 1
        self.g_module, self.d_module = self.generator, self.discriminator
        if self.distributed:
            import torch.distributed as dist
            from torch.nn.parallel import DistributedDataParallel as DDP

            self.world = dist.get_world_size()
            self.g_module = DDP(self.generator, device_ids=[self.device.index])
            self.d_module = DDP(self.discriminator, device_ids=[self.device.index])
        self.g_opt = torch.optim.Adam(self.generator.parameters(), lr=2e-4, betas=(0.5, 0.999), eps=1e-7)
        self.d_opt = torch.optim.Adam(self.discriminator.parameters(), lr=2e-4, betas=(0.5, 0.999), eps=1e-7)

    def generator_loss(self, fake_predicted, fake_image, real_image, global_batch=None):
        """`global_batch`: the whole batch over all ranks; None = equal shards (local batch x world size)."""
        adversarial = F.binary_cross_entropy_with_logits(fake_predicted, torch.ones_like(fake_predicted))
        l1 = (real_image - fake_image).abs().mean()
        if global_batch is None:
            global_batch = real_image.shape[0] * self.world
        hist = histogram.histogram_loss(real_image, fake_image, group=True if self.distributed else None,
                                        global_batch=global_batch, impl=self.impl)
        total = adversarial + self.lambda_l1 * l1 + self.lambda_histogram * hist
        return total, adversarial, l1, hist

    @staticmethod
    def discriminator_loss(real_predicted, fake_predicted):
        real = F.binary_cross_entropy_with_logits(real_predicted, torch.ones_like(real_predicted))
        fake = F.binary_cross_entropy_with_logits(fake_predicted, torch.zeros_like(fake_predicted))
        return fake + real, real, fake

    def train_step(self, source_image, real_image, global_batch=None):
        """One optimisation step on a (source, real) batch of NHWC float32 images in [-1, 1] (`global_batch`: the
        whole batch over all ranks when the shards are unequal, e.g. the last batch of an epoch).  Returns the loss
        terms as python floats-to-be (0-dim tensors, no host synchronisation)."""
        self.generator.train()
        self.discriminator.train()
        fake_image = self.g_module(source_image)
        # generator update: gradients reach the generator through the discriminator, whose own parameters
        # take their gradient from the discriminator loss only (the reference's two tape.gradient calls)
        for p in self.discriminator.parameters():
            p.requires_grad_(False)
        fake_predicted = self.discriminator(fake_image, source_image)
        g_total, g_adv, g_l1, g_hist = self.generator_loss(fake_predicted, fake_image, real_image, global_batch)
        # the histogram term is one function of every rank's shard; DDP averages gradients (module docstring)
        backward_total = g_total + (self.world - 1) * self.lambda_histogram * g_hist
        self.g_opt.zero_grad(set_to_none=True)
        backward_total.backward()
        for p in self.discriminator.parameters():
            p.requires_grad_(True)
        # discriminator update on the same fake images (detached); one forward over [real; fake] (the
        # discriminator has no batch statistics, so this equals the reference's two calls)
        both = self.d_module(torch.cat([real_image, fake_image.detach()], dim=0),
                             torch.cat([source_image, source_image], dim=0))
        real_predicted, fake_predicted_d = both.chunk(2, dim=0)
        d_total, d_real, d_fake = self.discriminator_loss(real_predicted, fake_predicted_d)
        self.d_opt.zero_grad(set_to_none=True)
        d_total.backward()
        self.g_opt.step()
        self.d_opt.step()
        return {"generator_total": g_total.detach(), "adversarial": g_adv.detach(), "l1": g_l1.detach(),
                "histogram": g_hist.detach(), "discriminator_total": d_total.detach(), "real": d_real.detach(),
                "fake": d_fake.detach()}


__all__ = ["UnetGenerator", "PatchDiscriminator", "Pix2PixHistogramStep", "IMG_SIZE"]
