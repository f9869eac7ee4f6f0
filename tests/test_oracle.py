"""CPU tests pinning the oracle itself (no GPU): known-answer values from SURVEY.md §4, the committed
golden fixtures, torch autograd, finite differences, and the identities the reference's call sites rely on.
PARITY UNPINNED against TensorFlow itself (not installable in this image); the strongest pin available is the
run of the reference's own source files over a TensorFlow-op shim (tests at the end of this file)."""
import os

import numpy as np
import pytest
import torch

from oracle import histogram_oracle as ho
from oracle import palette_oracle as po
from oracle import torch_port as tp


def test_linspace_matches_tf_formula():
    dom = ho.tf_linspace_f32(-3.0, 3.0, 64)
    assert dom.dtype == np.float32 and dom.shape == (64,)
    assert dom[0] == np.float32(-3.0) and dom[-1] == np.float32(3.0)
    delta = np.float32(np.float32(6.0) / np.float32(63.0))
    assert dom[17] == np.float32(np.float32(-3.0) + delta * np.float32(17.0))
    # TF's linspace is not exactly symmetric
    asym = np.abs(dom + dom[::-1]).max()
    assert 0 < asym < 1e-6
    assert ho.sigma_sqr_f32(0.02) == np.float32(4e-4)


def test_palette_known_answers(sprites, palette_golden):
    """SURVEY.md §4 known-answer values for train/2-front/0 || train/3-right/0, grayness ordering."""
    src, tgt = sprites["front"][0], sprites["right"][0]
    s_idx, t_idx, pal = po.load_indexed_images(src, tgt, "grayness")
    n = int((pal != np.array(po.INVALID_INDEX_COLOR)).any(-1).sum())
    assert n == 50
    assert pal[:4].tolist() == [[0, 0, 0, 0], [30, 30, 30, 255], [106, 14, 14, 255], [81, 32, 30, 255]]
    assert pal[48:50].tolist() == [[255, 255, 235, 255], [255, 255, 243, 255]]
    assert (pal[50:] == np.array([255, 0, 220, 255])).all()
    assert int((s_idx == 0).sum()) == 3406 and int(s_idx.max()) == 49
    assert np.array_equal(po.indexed_to_rgba(s_idx, pal), src.astype(np.int32))
    assert np.array_equal(pal, palette_golden["palette_grayness"][0].astype(np.int32))


def test_palette_golden_fixtures_reproduce(sprites, palette_golden):
    for ordering in ("grayness", "top2bottom", "bottom2top"):
        for i in (1, 7, 63, 64, 107):
            s, t = sprites["front"][i], sprites["right"][i]
            s_idx, t_idx, pal = po.load_indexed_images(s, t, ordering)
            assert np.array_equal(pal, palette_golden[f"palette_{ordering}"][i])
            assert np.array_equal(s_idx, palette_golden[f"src_idx_{ordering}"][i].astype(np.int32))
            assert np.array_equal(t_idx, palette_golden[f"tgt_idx_{ordering}"][i].astype(np.int32))
            assert np.array_equal(po.indexed_to_rgba(t_idx, pal), t.astype(np.int32))


def test_palette_orderings_and_ties():
    # alpha has weight 0: (0,0,0,0) and (0,0,0,255) tie and keep first-occurrence order (stable sort)
    img = np.array([[[0, 0, 0, 255], [9, 9, 9, 255], [0, 0, 0, 0], [1, 1, 1, 255]]], np.int32)
    pal, n = po.extract_palette(img, "grayness")
    assert n == 4 and pal[:4].tolist() == [[0, 0, 0, 255], [0, 0, 0, 0], [1, 1, 1, 255], [9, 9, 9, 255]]
    pal, _ = po.extract_palette(img, "top2bottom")
    assert pal[:4].tolist() == [[0, 0, 0, 255], [9, 9, 9, 255], [0, 0, 0, 0], [1, 1, 1, 255]]
    pal, _ = po.extract_palette(img, "bottom2top")
    assert pal[:4].tolist() == [[1, 1, 1, 255], [0, 0, 0, 0], [9, 9, 9, 255], [0, 0, 0, 255]]
    # single colour: defined as identity
    pal, n = po.extract_palette(np.full((2, 2, 4), 7, np.int32), "grayness")
    assert n == 1 and pal[0].tolist() == [7, 7, 7, 7]
    with pytest.raises(po.PaletteOverflow):
        many = np.stack([np.arange(257) % 256, np.arange(257) // 256, np.zeros(257), np.full(257, 255)], -1)
        po.extract_palette(many.astype(np.int32).reshape(257, 1, 4), "grayness")


def test_rgba_to_indexed_scatter_add_semantics():
    pal, _ = po.extract_palette(np.array([[[1, 2, 3, 255], [4, 5, 6, 255]]], np.int32), "top2bottom")
    # pixel equal to the filler colour while the palette is not full: indices 2..255 add up
    img = np.array([[[255, 0, 220, 255], [4, 5, 6, 255], [7, 7, 7, 7]]], np.int32)
    idx = po.rgba_to_indexed(img, pal)
    assert idx.reshape(-1).tolist() == [sum(range(2, 256)), 1, 0]
    oh = po.one_hot(idx)
    assert oh.shape == (1, 3, 256) and oh[0, 0].sum() == 0 and oh[0, 1, 1] == 1 and oh[0, 2, 0] == 1
    assert po.rgba_to_nearest(img, pal).reshape(-1).tolist()[1] == 1


def test_histogram_identities(hist_golden):
    real = hist_golden["real"][:2]
    h, d = ho.rgbuv_histogram_f64(real)
    assert np.allclose(h.sum(axis=(1, 2, 3)), 1.0, atol=1e-12)
    assert ho.hellinger_loss_f64(h, h) == 0.0
    # alpha is ignored (histogram.py:61)
    changed = real.copy()
    changed[..., 3] = 0.123
    assert np.array_equal(ho.rgbuv_histogram_f64(changed)[0], h)
    # an all-black image: u = v = 0 for every pixel -> every channel is the same rank-1 outer product
    hb, _ = ho.rgbuv_histogram_f64(np.full((1, 8, 8, 4), -1.0, np.float32))
    k0 = 1.0 / (1.0 + ho.tf_linspace_f32(-3, 3, 64).astype(np.float64) ** 2 / float(ho.sigma_sqr_f32()))
    outer = np.outer(k0, k0)
    assert np.allclose(hb[0, :, :, 0], outer / (3 * outer.sum()), rtol=1e-12)
    assert np.allclose(hb[0, :, :, 1], hb[0, :, :, 0]) and np.allclose(hb[0, :, :, 2], hb[0, :, :, 0])


def test_f32_restatement_close_to_f64(hist_golden):
    """the reference's own float32 arithmetic is 2-5e-6 norm-relative from float64."""
    assert ho.rel_l2(hist_golden["hist_fake_f32"], hist_golden["hist_fake"]) < 1e-5
    assert ho.rel_max(hist_golden["hist_real_f32"], hist_golden["hist_real"]) < 1e-5
    l32 = ho.hellinger_loss_f32(hist_golden["hist_real_f32"], hist_golden["hist_fake_f32"])
    assert abs(float(l32) - float(hist_golden["loss"])) / float(hist_golden["loss"]) < 1e-5


def test_golden_loss_and_grad_reproduce(hist_golden):
    res = ho.hist_loss_and_grad_f64(hist_golden["real"][:2], hist_golden["fake"][:2], global_batch=8,
                                    global_ssum=float(hist_golden["ssum"]))
    assert np.allclose(res["grad"], hist_golden["grad"][:2], rtol=1e-10, atol=1e-14)
    assert abs(res["loss"] - float(hist_golden["loss"])) < 1e-14


def test_analytic_gradient_vs_torch_autograd(hist_golden):
    real = hist_golden["real"][:1, ::2, ::2]
    fake = hist_golden["fake"][:1, ::2, ::2]
    res = ho.hist_loss_and_grad_f64(real, fake)
    loss, grad = tp.hist_loss_fwd_bwd(torch.from_numpy(real).double(), torch.from_numpy(fake).double())
    assert abs(float(loss) - res["loss"]) < 1e-13
    assert ho.rel_l2(res["grad"], grad.numpy()) < 1e-12
    assert np.abs(res["grad"][..., 3]).max() == 0.0


def test_analytic_gradient_vs_finite_differences():
    rng = np.random.default_rng(3)
    real = np.tanh(rng.standard_normal((2, 6, 6, 4))).astype(np.float32)
    fake = np.tanh(rng.standard_normal((2, 6, 6, 4))).astype(np.float32)
    res = ho.hist_loss_and_grad_f64(real, fake, size=16)
    h = 2.0 ** -17  # exactly representable in float32, so the perturbed images are exact
    for (b, y, x, c) in [(0, 1, 2, 0), (1, 3, 3, 1), (0, 5, 0, 2), (1, 0, 4, 0)]:
        up, dn = fake.copy(), fake.copy()
        up[b, y, x, c] += np.float32(h)
        dn[b, y, x, c] -= np.float32(h)
        lu = ho.hellinger_loss_f64(ho.rgbuv_histogram_f64(real, 16)[0], ho.rgbuv_histogram_f64(up, 16)[0])
        ld = ho.hellinger_loss_f64(ho.rgbuv_histogram_f64(real, 16)[0], ho.rgbuv_histogram_f64(dn, 16)[0])
        fd = (lu - ld) / (float(up[b, y, x, c]) - float(dn[b, y, x, c]))
        assert abs(fd - res["grad"][b, y, x, c]) < 2e-4 * max(abs(fd), 1e-6) + 1e-9


def test_sharded_loss_equals_whole_batch(hist_golden):
    """shards are coupled only through S and the global batch size."""
    real, fake = hist_golden["real"][:4, ::2, ::2], hist_golden["fake"][:4, ::2, ::2]
    whole = ho.hist_loss_and_grad_f64(real, fake, size=32)
    parts = [ho.hist_loss_and_grad_f64(real[i:i + 2], fake[i:i + 2], size=32) for i in (0, 2)]
    ssum = sum(p["ssum"] for p in parts)
    assert abs(ssum - whole["ssum"]) < 1e-12
    for k, i in enumerate((0, 2)):
        sh = ho.hist_loss_and_grad_f64(real[i:i + 2], fake[i:i + 2], size=32, global_batch=4, global_ssum=ssum)
        assert abs(sh["loss"] - whole["loss"]) < 1e-14
        assert np.allclose(sh["grad"], whole["grad"][i:i + 2], rtol=1e-10, atol=1e-16)


def test_rbf_method_and_unknown_method():
    rng = np.random.default_rng(5)
    img = np.tanh(rng.standard_normal((1, 8, 8, 3))).astype(np.float32)
    h, _ = ho.rgbuv_histogram_f64(img, 16, method="RBF")
    assert np.isclose(h.sum(), 1.0)
    # sigma large enough that exp(-t) never underflows to 0 (sqrt'(0) = inf gives NaN in TF as well)
    loss, grad = tp.hist_loss_fwd_bwd(torch.from_numpy(img).double(), torch.from_numpy(img * 0.5).double(), 16, "RBF", 0.5)
    res = ho.hist_loss_and_grad_f64(img, img * 0.5, 16, "RBF", 0.5)
    assert ho.rel_l2(res["grad"], grad.numpy()) < 1e-10
    with pytest.raises(ValueError):
        ho.rgbuv_histogram_f64(img, 16, method="thresholding")


def test_argmax_indexed_semantics():
    """Indexed-model inference ops (pix2pix_model.py:283-287, 356): first maximum, NaN never selected."""
    p = np.array([[0.1, 0.7, 0.7, 0.2], [np.nan, 0.3, 0.1, 0.3], [np.nan] * 4, [-np.inf] * 4], np.float32)[None]
    idx = po.argmax_indexed(p)
    assert idx.dtype == np.int32 and idx.shape == (1, 4, 1)
    assert idx[0, :, 0].tolist() == [1, 1, 0, 0]
    pal = np.arange(16, dtype=np.int32).reshape(4, 4)
    assert np.array_equal(po.probabilities_to_rgba(p, pal)[0], pal[[1, 1, 0, 0]])


# ---------------------------------------------------------------------------------------------------
# Pinning against the reference's own source (tests/golden/reference_run.npz: /root/reference/histogram.py,
# io_utils.py and dataset_utils.py executed unmodified over the TensorFlow-op shim oracle/ref_shim.py)
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("ordering", ["grayness", "top2bottom", "bottom2top"])
def test_palette_oracle_equals_reference_source(reference_run, ordering):
    """dataset_utils.load_indexed_images (:131-151) -> io_utils.extract_palette / rgba_to_indexed, bit-exact."""
    R = reference_run
    for n in range(R["loader_source"].shape[0]):
        s, t, p = po.load_indexed_images(R["loader_source"][n].astype(np.int32), R["loader_target"][n].astype(np.int32),
                                         ordering)
        assert np.array_equal(p, R[f"palette_{ordering}"][n])
        assert np.array_equal(s, R[f"src_idx_{ordering}"][n]) and np.array_equal(t, R[f"tgt_idx_{ordering}"][n])


def test_palette_helpers_equal_reference_source(reference_run):
    R = reference_run
    n = R["roundtrip_rgba"].shape[0]
    for k in range(n):
        assert np.array_equal(po.indexed_to_rgba(R["tgt_idx_grayness"][k], R["palette_grayness"][k]), R["roundtrip_rgba"][k])
    assert np.array_equal(R["roundtrip_rgba"], R["loader_target"][:n].astype(np.int32))  # io_utils.py:96-103 round trip
    tgt = R["loader_target"][:n].astype(np.float32)
    assert np.array_equal(po.normalize(tgt), R["normalized"])                              # dataset_utils.py:39-48
    oh = po.one_hot(R["tgt_idx_grayness"][:2])[:, ::8, ::8]                                  # pix2pix_model.py:300-301
    assert np.array_equal(oh.reshape(R["one_hot_rows"].shape), R["one_hot_rows"])
    # the decoded sprites are blackened where transparent (dataset_utils.py:11-20, :66-77)
    src = R["loader_source"]
    assert (src[src[..., 3] == 0] == 0).all()


def test_histogram_oracle_equals_reference_source(reference_run, hist_golden):
    """histogram.py:36-81 + :84-97 and autograd through them (the analogue of tape.gradient, pix2pix_model.py:78)."""
    R, G = reference_run, hist_golden
    assert np.array_equal(R["linspace_64"], ho.tf_linspace_f32(-3.0, 3.0, 64))  # bin centres bit for bit
    # float32 restatement against the reference's float32 run; float64 truth within the float32 noise of the reference
    assert ho.rel_l2(ho.rgbuv_histogram_f32(G["fake"]), R["hist_fake"]) < 3e-6
    assert ho.rel_l2(R["hist_real"], G["hist_real"]) < 5e-6 and ho.rel_l2(R["hist_fake"], G["hist_fake"]) < 5e-6
    assert abs(float(R["loss"].reshape(-1)[0]) - float(G["loss"])) / float(G["loss"]) < 1e-6
    # the reference's own float32 gradient is 3e-5 from float64 on these near-black sprites
    assert ho.rel_l2(R["grad"], G["grad"]) < 5e-5
    assert np.abs(R["grad"][..., 3]).max() == 0.0
    assert abs(float(R["l1"]) - ho.l1_loss_f64(R["hist_real"], R["hist_fake"])) < 1e-9
    assert abs(float(R["l2"]) - ho.l2_loss_f64(R["hist_real"], R["hist_fake"])) < 1e-11
    d = R["dense_input"]
    a, _ = ho.rgbuv_histogram_f64(d, size=32)
    assert ho.rel_l2(R["dense_hist_32_iq"], a) < 5e-6
    a, _ = ho.rgbuv_histogram_f64(d, size=64, method="RBF", sigma=0.5)
    assert ho.rel_l2(R["dense_hist_64_rbf"], a) < 5e-6
    raw = ho.raw_histogram_f64(d, ho.tf_linspace_f32(-3.0, 3.0, 16), ho.sigma_sqr_f32(0.02))
    assert ho.rel_l2(R["component_hist"], raw[..., 0]) < 5e-6  # calculate_component_histogram (R, G, B)


def test_torch_port_is_the_reference_source_op_for_op(reference_run, hist_golden):
    """The CPU arm of bench.py (oracle/torch_port.py) reproduces the reference-source run exactly."""
    import torch
    from oracle import torch_port as tp

    loss, grad = tp.hist_loss_fwd_bwd(torch.from_numpy(hist_golden["real"]), torch.from_numpy(hist_golden["fake"]), 64)
    assert abs(float(loss) - float(reference_run["loss"].reshape(-1)[0])) / float(loss) < 1e-6
    assert ho.rel_l2(grad.numpy(), reference_run["grad"]) < 1e-6


def test_tf_shim_op_semantics():
    """The TensorFlow-op behaviours oracle/ref_shim.py assumes (documented TF semantics), stated as executable
    checks so that a reader can see exactly what the reference-source run rests on."""
    import torch
    from oracle import ref_shim

    tf = ref_shim.build()
    # linspace: exact end points, start + delta * k in float32 in between (the formula histogram_oracle restates)
    assert np.array_equal(tf.linspace(-3.0, 3.0, num=64).numpy(), ho.tf_linspace_f32(-3.0, 3.0, 64))
    # UniqueWithCountsV2(axis=[0]): rows in order of first occurrence
    rows = ref_shim.T(np.array([[3, 3], [1, 1], [3, 3], [2, 2], [1, 1]], np.int32))
    uniq, inv, cnt = tf.raw_ops.UniqueWithCountsV2(x=rows, axis=[0])
    assert uniq.numpy().tolist() == [[3, 3], [1, 1], [2, 2]] and inv.numpy().tolist() == [0, 1, 0, 2, 1]
    assert cnt.numpy().tolist() == [2, 2, 1]
    # scatter_nd accumulates duplicates and leaves zeros elsewhere; where() lists coordinates row-major
    out = tf.scatter_nd(ref_shim.T(np.array([[1], [3], [1]], np.int32)), ref_shim.T(np.array([5, 7, 2], np.int32)), [5])
    assert out.numpy().tolist() == [0, 7, 0, 7, 0]
    assert tf.where(ref_shim.T(np.array([[0, 1], [1, 1]], np.int32)) == 1).numpy().tolist() == [[0, 1], [1, 0], [1, 1]]
    # stable ascending argsort, one_hot out of range, repeat with a negative count
    assert tf.argsort(ref_shim.T(np.array([2.0, 1.0, 2.0, 1.0], np.float32)), direction="ASCENDING", stable=True).numpy().tolist() == [1, 3, 0, 2]
    assert tf.one_hot(ref_shim.T(np.array([0, 2, 5, -1], np.int32)), 3).numpy().tolist() == [[1, 0, 0], [0, 0, 1], [0, 0, 0], [0, 0, 0]]
    with pytest.raises((ValueError, RuntimeError)):
        tf.repeat([[255, 0, 220, 255]], [-3], axis=0)
    # python scalars are weakly typed: float32 stays float32; pow(sigma, 2) is the float32 square
    x = ref_shim.T(np.array([0.25], np.float32))
    assert (x * 0.5 + 0.5).dtype == torch.float32
    assert np.float32(tf.pow(0.02, 2).numpy()) == ho.sigma_sqr_f32(0.02)
    # x[::-1] reverses the first axis (io_utils.py:48)
    assert ref_shim.T(np.arange(6, dtype=np.int32).reshape(3, 2))[::-1].numpy().tolist() == [[4, 5], [2, 3], [0, 1]]


@pytest.mark.skipif(not os.path.isdir("/root/reference/datasets"), reason="build container only: needs the reference")
def test_reference_run_fixture_is_reproducible(tmp_path, monkeypatch, reference_run):
    """Re-running the reference's own source over the shim reproduces the committed fixture bit for bit."""
    import shutil
    from oracle import run_reference

    monkeypatch.setattr(run_reference, "OUT", str(tmp_path))
    shutil.copy(os.path.join(os.path.dirname(__file__), "golden", "hist_golden.npz"), tmp_path / "hist_golden.npz")
    run_reference.main()
    fresh = dict(np.load(tmp_path / "reference_run.npz"))
    assert set(fresh) == set(reference_run)
    for k in fresh:
        assert np.array_equal(fresh[k], reference_run[k]), k


# ---------------------------------------------------------------------------------------------------
# Augmentation (dataset_utils.py:80-120)
# ---------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def reference_augment():
    return dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_augment.npz")))


def test_augment_oracle_equals_reference_source(reference_augment):
    """The reference's own `augment_two` (run over the shim, whose image ops are the restated TF algorithms) and a
    direct call of the oracle agree bit for bit: pins the glue (hue on channels 0..2 only, alpha carried, one
    translation shared by the channel-concatenated pair)."""
    from oracle import augment_oracle as ao

    R = reference_augment
    for n in range(R["first"].shape[0]):
        a, b = ao.augment_two(R["first"][n], R["second"][n], R["hue_delta"][n], *R["translation"][n])
        assert np.array_equal(a, R["out_first"][n]) and np.array_equal(b, R["out_second"][n])
    assert np.array_equal(ao.normalize(R["out_first"][:4]), R["normalized_first"])
    assert np.all(np.abs(R["hue_delta"]) <= 0.5)
    assert np.all(np.abs(R["translation"][:, 0]) <= 8.0) and np.all(R["translation"][:, 1] >= -9.6 - 1e-5)
    assert np.all(R["translation"][:, 1] <= 4.8 + 1e-5)


def test_adjust_hue_against_hsv_arithmetic():
    """Independent check of the restated `tf.image.adjust_hue`: the textbook RGB -> HSV -> shift h -> RGB route
    (python's colorsys, float64) gives the same colours to float32 round-off; grey stays grey; a full turn and a
    zero shift are the identity (up to round-off), +delta then -delta returns."""
    import colorsys
    from oracle import augment_oracle as ao

    rng = np.random.default_rng(5)
    x = rng.integers(0, 256, (400, 3)).astype(np.float32)
    x[:20, 1] = x[:20, 0]
    x[20:40, 2] = x[20:40, 1]
    x[40:60] = x[40:60, :1]
    for d in (-0.5, -0.31, 0.0, 0.125, 0.49):
        y = ao.adjust_hue_f32(x, d)
        ref = np.array([colorsys.hsv_to_rgb((colorsys.rgb_to_hsv(*(p / 255.0))[0] + d) % 1.0,
                                            *colorsys.rgb_to_hsv(*(p / 255.0))[1:]) for p in x]) * 255.0
        assert np.abs(y - ref).max() < 5e-4, d
        assert np.array_equal(y[40:60], x[40:60])              # grey pixels: v_max == v_min
        assert np.allclose(np.sort(y, -1)[:, [0, 2]], np.sort(x, -1)[:, [0, 2]])  # min and max are kept
    assert np.abs(ao.adjust_hue_f32(x, 0.0) - x).max() < 2e-4
    assert np.abs(ao.adjust_hue_f32(x, 1.0) - x).max() < 2e-4
    assert np.abs(ao.adjust_hue_f32(ao.adjust_hue_f32(x, 0.2), -0.2) - x).max() < 5e-4


def test_translate_nearest_semantics():
    """ImageProjectiveTransformV3 (NEAREST, CONSTANT 0) with a pure translation: integer shifts move the image,
    x.5 offsets round half away from zero, everything that leaves the frame is zero."""
    from oracle import augment_oracle as ao

    img = np.arange(1, 6 * 5 * 2 + 1, dtype=np.float32).reshape(6, 5, 2)
    out = ao.translate_nearest(img, 2.0, -1.0)        # right by 2, up by 1
    assert np.array_equal(out[:5, 2:], img[1:, :3]) and not out[:, :2].any() and not out[5].any()
    assert np.array_equal(ao.translate_nearest(img, 0.0, 0.0), img)
    # out[x] = in[round(x - 0.5)]: round(-0.5) = -1 (outside), round(0.5) = 1, round(1.5) = 2 ...
    half = ao.translate_nearest(img, 0.5, 0.0)
    assert not half[:, 0].any() and np.array_equal(half[:, 1:4], img[:, 1:4]) and np.array_equal(half[:, 4], img[:, 4])
    assert not ao.translate_nearest(img, 50.0, 0.0).any()
    a, b = ao.augment_translation((img[..., :1], img[..., 1:]), 1.0, 1.0)
    assert np.array_equal(a[1:, 1:, 0], img[:-1, :-1, 0]) and np.array_equal(b[1:, 1:, 0], img[:-1, :-1, 1])


@pytest.mark.skipif(not os.path.isdir("/root/reference/datasets"), reason="build container only: needs the reference")
def test_reference_augment_fixture_is_reproducible(tmp_path, monkeypatch, reference_augment):
    from oracle import run_reference

    monkeypatch.setattr(run_reference, "OUT", str(tmp_path))
    run_reference.main_augment()
    fresh = dict(np.load(tmp_path / "reference_augment.npz"))
    assert set(fresh) == set(reference_augment)
    for k in fresh:
        assert np.array_equal(fresh[k], reference_augment[k]), k


def test_log_table():
    """The float64 log of the kernels' pixel terms (common.cuh: log_pos): the table in the source is the generator's, and
    the formula stays within 3e-15 of long-double logs over the range of x + eps (histogram.py:58-66)."""
    import re
    from oracle import make_log_table as mt
    rows = mt.table()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "palette_and_histo_gan_b200", "csrc", "common.cuh")).read()
    body = src[src.index("PH_LOG_TABLE[128] = {"):]
    body = body[:body.index("};")]
    vals = [float.fromhex(v) for v in re.findall(r"-?0x[0-9a-f.]+p[-+]?\d+", body)]
    assert len(vals) == 256
    assert vals == [v for row in rows for v in row]
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(1e-6, 1.000001, 100000), 10.0 ** rng.uniform(-6, 0, 100000),
                        rng.uniform(0.99, 1.000001, 20000), np.array([1e-6, 1.0, 1.000001, 0.5, 0.25 + 1e-6])])
    err = np.abs(mt.log_pos_numpy(x, rows).astype(np.longdouble) - np.log(x.astype(np.longdouble)))
    assert float(err.max()) < 3e-15, float(err.max())


def test_logf_difference_table():
    """The float32 log differences of the mirrored forward's pixel pass (hist_tc.cu: log_parts / log_diff): the table in
    the source is the generator's; the formula replayed in numpy float32 is within 1e-9 + half an ulp of the long-double
    difference (histogram.py:72-74 takes log(x_a + eps) - log(x_b + eps))."""
    import re
    from oracle import make_log_table as mt
    rows = mt.table_f32()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "palette_and_histo_gan_b200", "csrc", "hist_tc.cu")).read()
    body = src[src.index("PH_LOGF_TABLE[128] = {"):]
    body = body[:body.index("};")]
    vals = [float.fromhex(v) for v in re.findall(r"(-?0x[0-9a-f.]+p[-+]?\d+)f", body)]
    assert len(vals) == 384
    assert vals == [v for row in rows for v in row]
    assert "LN2_HI = 0x1.62e4p-1f, LN2_LO = %rf" % mt.LN2_LO in src
    assert mt.LN2_HI == float.fromhex("0x1.62e4p-1")
    rng = np.random.default_rng(1)
    n = 100000
    x0 = np.concatenate([rng.uniform(1e-6, 1.000001, n), 10.0 ** rng.uniform(-6, 0, n)]).astype(np.float32)
    x1 = rng.permutation(x0)
    d = mt.logdiff_f32_numpy(x0, x1, rows)
    LD = np.longdouble
    ref = np.log(x0.astype(LD)) - np.log(x1.astype(LD))
    err = np.abs(d.astype(LD) - ref)
    ulp = np.spacing(np.abs(ref.astype(np.float32))).astype(LD)
    assert float((err - 0.5 * ulp).max()) < 1e-9, float((err - 0.5 * ulp).max())
