"""GPU parity tests of the augmentation kernel (dataset_utils.py:80-120), through the C ABI.
The kernel evaluates tf.image.adjust_hue's float32 algorithm without fused multiply-add, so every comparison with
the op-for-op oracle is bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle import augment_oracle as ao
from tests.conftest import sprite_like_batch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def D():
    import palette_and_histo_gan_b200 as pkg

    return pkg.dataset_utils


@pytest.fixture(scope="module")
def reference_augment():
    return dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_augment.npz")))


def dev(a, cuda):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(cuda)


def test_against_the_reference_source_run(D, cuda, reference_augment):
    """`augment_two` of the reference's own source (fixture) on 16 real sprite pairs, batched and one by one."""
    R = reference_augment
    a, b = D.augment_two(dev(R["first"], cuda), dev(R["second"], cuda), hue_delta=torch.from_numpy(R["hue_delta"]),
                         translations=torch.from_numpy(R["translation"]))
    assert np.array_equal(a.cpu().numpy(), R["out_first"]) and np.array_equal(b.cpu().numpy(), R["out_second"])
    for n in range(3):
        a1, b1 = D.augment_two(dev(R["first"][n], cuda), dev(R["second"][n], cuda), hue_delta=float(R["hue_delta"][n]),
                               translations=R["translation"][n].tolist())
        assert a1.shape == (64, 64, 4)
        assert np.array_equal(a1.cpu().numpy(), R["out_first"][n]) and np.array_equal(b1.cpu().numpy(), R["out_second"][n])
    # fused normalize (load_rgba_ds: augmentation, then normalize_two)
    an, _ = D.augment_two(dev(R["first"][:4], cuda), dev(R["second"][:4], cuda), hue_delta=torch.from_numpy(R["hue_delta"][:4]),
                          translations=torch.from_numpy(R["translation"][:4]), should_normalize=True)
    assert np.array_equal(an.cpu().numpy(), R["normalized_first"])


def test_adjust_hue_dense_colours_and_ties(D, cuda):
    rng = np.random.default_rng(11)
    img = rng.integers(0, 256, (8, 32, 48, 4)).astype(np.float32)
    img[0, :, :, 1] = img[0, :, :, 0]            # r == g
    img[1, :, :, 2] = img[1, :, :, 1]            # g == b
    img[2, :, :, :3] = img[2, :, :, :1]          # grey
    img[3] = np.tanh(rng.standard_normal((32, 48, 4))).astype(np.float32)  # [-1, 1] range works the same
    deltas = np.array([-0.5, -0.25, 0.3, 0.4999, 0.0, 1.0, -1.0, 0.013], np.float32)
    out = D.adjust_hue(dev(img, cuda), torch.from_numpy(deltas)).cpu().numpy()
    for n in range(8):
        assert np.array_equal(out[n], ao.augment_hue_rotation(img[n], deltas[n])), n
    assert np.array_equal(out[..., 3], img[..., 3])
    with pytest.raises(ValueError):
        D.adjust_hue(dev(img, cuda), 1.5)


def test_translation_edges_and_rounding(D, cuda):
    rng = np.random.default_rng(3)
    a = rng.integers(0, 256, (6, 20, 12, 4)).astype(np.float32)
    b = rng.integers(0, 256, (6, 20, 12, 4)).astype(np.float32)
    tr = np.array([[0.0, 0.0], [0.5, -0.5], [-1.5, 2.5], [3.49, -2.51], [100.0, 0.0], [-11.4999, 19.4999]], np.float32)
    oa, ob = D.augment_translation((dev(a, cuda), dev(b, cuda)), translations=torch.from_numpy(tr))
    for n in range(6):
        ea, eb = ao.augment_translation((a[n], b[n]), tr[n, 0], tr[n, 1])
        assert np.array_equal(oa[n].cpu().numpy(), ea) and np.array_equal(ob[n].cpu().numpy(), eb), n
    assert not oa[4].any()
    # drawn translations stay inside keras' factors: dx in W*[-0.125, 0.125], dy in H*[-0.15, 0.075]
    g = torch.Generator().manual_seed(47)
    t = D._draw_translations(1000, 64, 64, g).numpy()
    assert t[:, 0].min() >= -8 and t[:, 0].max() <= 8 and t[:, 1].min() >= -9.6001 and t[:, 1].max() <= 4.8001
    d = D._draw_hue_delta(1000, None, g).numpy()
    assert d.min() >= -0.5 and d.max() < 0.5 and abs(d.mean()) < 0.05


def test_probability_gate_and_seeded_rotation(D, cuda):
    rng = np.random.default_rng(9)
    a = sprite_like_batch(rng, 64).astype(np.float32)
    b = sprite_like_batch(rng, 64).astype(np.float32)
    da, db = dev(a, cuda), dev(b, cuda)
    mask = rng.random(64) < 0.8
    deltas = rng.uniform(-0.5, 0.5, 64).astype(np.float32)
    tr = np.stack([rng.uniform(-8, 8, 64), rng.uniform(-9.6, 4.8, 64)], 1).astype(np.float32)
    oa, ob = D.augment_two(da, db, hue_delta=torch.from_numpy(deltas), translations=torch.from_numpy(tr),
                           apply=torch.from_numpy(mask))
    for n in range(64):
        if mask[n]:
            ea, eb = ao.augment_two(a[n], b[n], deltas[n], tr[n, 0], tr[n, 1])
        else:
            ea, eb = a[n], b[n]
        assert np.array_equal(oa[n].cpu().numpy(), ea) and np.array_equal(ob[n].cpu().numpy(), eb), n
    # the wrapper: prob 0 is the identity, prob 1 changes (nearly) every sample, the generator makes it reproducible
    ia, ib = D.create_augmentation_with_prob(0.0)(da, db)
    assert torch.equal(ia, da) and torch.equal(ib, db)
    w1 = D.create_augmentation_with_prob(1.0, generator=torch.Generator().manual_seed(1))(da, db)
    w2 = D.create_augmentation_with_prob(1.0, generator=torch.Generator().manual_seed(1))(da, db)
    assert torch.equal(w1[0], w2[0]) and torch.equal(w1[1], w2[1])
    assert int((w1[0] != da).flatten(1).any(1).sum()) > 56
    # equal seeds, equal rotation (what augment_two relies on, dataset_utils.py:97-99)
    h1 = D.augment_hue_rotation(da, seed=[12, 3456])
    h2 = D.augment_hue_rotation(da, seed=[12, 3456])
    h3 = D.augment_hue_rotation(da, seed=[12, 3457])
    assert torch.equal(h1, h2) and not torch.equal(h1, h3)
    n1, n2 = D.normalize_two(da, db)
    assert np.array_equal(n1.cpu().numpy(), ao.normalize(a))


def test_full_size_properties(D, cuda):
    """Size-independent properties at a loader-sized batch (4096 pairs of 64x64): hue keeps min / max / alpha of
    every pixel, a grey image is a fixed point, translation conserves the multiset of pixels that stay inside."""
    g = torch.Generator(device="cpu").manual_seed(4)
    a = (torch.rand(4096, 64, 64, 4, generator=g) * 255).round().to(cuda)
    b = a.flip(0).contiguous()
    deltas = (torch.rand(4096, generator=g) - 0.5)
    ha = D.adjust_hue(a, deltas)
    assert torch.equal(ha[..., 3], a[..., 3])
    assert torch.equal(ha[..., :3].amax(-1), a[..., :3].amax(-1)) and torch.equal(ha[..., :3].amin(-1), a[..., :3].amin(-1))
    grey = a[..., :1].expand(-1, -1, -1, 4).contiguous()
    assert torch.equal(D.adjust_hue(grey, deltas), grey)
    ta, tb = D.augment_translation((a, b), translations=torch.tensor([3.0, -2.0]))
    assert torch.equal(ta[:, :62, 3:], a[:, 2:, :61]) and torch.equal(tb[:, :62, 3:], b[:, 2:, :61])
    assert not ta[:, 62:].any() and not ta[:, :, :3].any()
    with pytest.raises(ValueError):
        D.augment_two(a, b[:8])
    with pytest.raises(ValueError):
        D.augment_two(a.cpu(), b.cpu())
