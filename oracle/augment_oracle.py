"""CPU restatement of the reference's augmentation (TEST INFRASTRUCTURE — nothing in the product imports this).

Follows /root/reference/dataset_utils.py:80-120:

    augment_hue_rotation (:80-84)   tf.image.stateless_random_hue(rgb, 0.5, seed) on channels 0..2, alpha kept
    augment_translation  (:87-92)   concat on the channel axis -> keras RandomTranslation((-0.15, 0.075), 0.125,
                                    fill_mode="constant", interpolation="nearest") -> split
    augment_two          (:95-102)  the same hue shift for both images, then one shared translation
    create_augmentation_with_prob (:109-120)  applied when uniform() < prob

The arithmetic lives in TensorFlow 2.9.1 (requirements.txt:99; not vendored, not installable here), restated from
its published algorithm — **parity unpinned** against a real TensorFlow run:

* `tf.image.adjust_hue` on float32 (CPU kernel AdjustHueOp): the pixel is reduced to (h in [0,6), v_min, v_max) by
  ordering its three components (six "categories", ties resolved by the `<` comparisons below), h is shifted by
  6*delta and wrapped with repeated +-6, and the colour is rebuilt from (h, v_min, v_max).  All float32, no fused
  multiply-add.  The value range is irrelevant (the reference applies it to [0,255] floats).
* `stateless_random_hue(image, max_delta, seed)`: delta = stateless uniform in [-max_delta, max_delta) — the draw
  itself (Philox keyed by the seed) is not restated; functions here take `delta` as an argument.
* keras `RandomTranslation`: per image (dx, dy) = (uniform(-0.125, 0.125) * W, uniform(-0.15, 0.075) * H) pixels,
  NOT rounded; `ImageProjectiveTransformV3` with the matrix [1,0,-dx, 0,1,-dy, 0,0], NEAREST, CONSTANT fill 0:
  out[y,x] = in[round(y - dy), round(x - dx)] when that lies inside the image, else 0; round = half away from
  zero (std::round), the coordinate evaluated in float32.
"""
from __future__ import annotations

import numpy as np

F = np.float32
HEIGHT_FACTOR = (-0.15, 0.075)   # dataset_utils.py:89
WIDTH_FACTOR = (-0.125, 0.125)   # dataset_utils.py:89 (a scalar factor f means (-f, f))
MAX_HUE_DELTA = 0.5              # dataset_utils.py:82


def adjust_hue_f32(rgb, delta):
    """`tf.image.adjust_hue(rgb, delta)` for float32 (…,3); scalar python loops kept out: vectorised numpy with
    the same branch structure as the TF CPU kernel."""
    rgb = np.asarray(rgb, F)
    r, g, b = rgb[..., 0], rgb[..., 1], rgb[..., 2]
    # rgb_to_hv_range: category by the kernel's comparison tree
    r_lt_g = r < g
    cat = np.where(r_lt_g,
                   np.where(b < r, 1, np.where(b > g, 3, 2)),
                   np.where(b < g, 0, np.where(b > r, 4, 5))).astype(np.int32)
    vmax = np.choose(cat, [r, g, g, b, b, r]).astype(F)
    vmid = np.choose(cat, [g, r, b, g, r, b]).astype(F)
    vmin = np.choose(cat, [b, b, r, r, g, g]).astype(F)
    span = (vmax - vmin).astype(F)
    flat = vmax == vmin
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = ((vmid - vmin).astype(F) / span).astype(F)
    increase = (cat & 1) == 0
    h = (cat.astype(F) + np.where(increase, ratio, (F(1) - ratio).astype(F))).astype(F)
    h = np.where(flat, F(0), h).astype(F)
    # shift and wrap into [0, 6) by repeated +-6 (the kernel's while loops)
    h = (h + (F(delta) * F(6)).astype(F)).astype(F)
    for _ in range(4):
        h = np.where(h < 0, (h + F(6)).astype(F), h).astype(F)
    for _ in range(4):
        h = np.where(h >= 6, (h - F(6)).astype(F), h).astype(F)
    # hv_range_to_rgb
    cat2 = h.astype(np.int32)
    ratio2 = (h - cat2.astype(F)).astype(F)
    ratio2 = np.where((cat2 & 1) == 0, ratio2, (F(1) - ratio2).astype(F)).astype(F)
    mid = (vmin + (ratio2 * (vmax - vmin).astype(F)).astype(F)).astype(F)
    cat2 = np.clip(cat2, 0, 5)  # `default:` of the switch is category 5
    out_r = np.choose(cat2, [vmax, mid, vmin, vmin, mid, vmax])
    out_g = np.choose(cat2, [mid, vmax, vmax, mid, vmin, vmin])
    out_b = np.choose(cat2, [vmin, vmin, mid, vmax, vmax, mid])
    return np.stack([out_r, out_g, out_b], -1).astype(F)


def augment_hue_rotation(image, delta):
    """dataset_utils.py:80-84 with the random draw replaced by its value: (…,4) float32."""
    image = np.asarray(image, F)
    return np.concatenate([adjust_hue_f32(image[..., 0:3], delta), image[..., 3:4]], -1)


def _round_half_away(x):
    x = np.asarray(x, F).astype(np.float64)  # x +- 0.5 is exact in float64: std::round, not round(x + 0.5f)
    return np.where(x >= 0, np.floor(x + 0.5), np.ceil(x - 0.5)).astype(np.int64)


def translate_nearest(image, dx, dy, fill=0.0):
    """ImageProjectiveTransformV3([1,0,-dx,0,1,-dy,0,0], NEAREST, CONSTANT) on one (H,W,C) image."""
    image = np.asarray(image)
    h, w = image.shape[:2]
    xs = (np.arange(w, dtype=F) + F(-F(dx))).astype(F)   # 1*x + 0*y + (-dx): the zero product adds nothing
    ys = (np.arange(h, dtype=F) + F(-F(dy))).astype(F)
    sx, sy = _round_half_away(xs), _round_half_away(ys)
    okx, oky = (sx >= 0) & (sx < w), (sy >= 0) & (sy < h)
    out = np.full_like(image, fill)
    src = image[np.clip(sy, 0, h - 1)][:, np.clip(sx, 0, w - 1)]
    ok = oky[:, None] & okx[None, :]
    out[ok] = src[ok]
    return out


def augment_translation(images, dx, dy):
    """dataset_utils.py:87-92: the images share one translation (they are concatenated on the channel axis)."""
    cat = np.concatenate([np.asarray(i) for i in images], -1)
    moved = translate_nearest(cat, dx, dy)
    return tuple(np.split(moved, len(images), -1))


def augment_two(first, second, delta, dx, dy):
    """dataset_utils.py:95-102 with the three random draws given."""
    first = augment_hue_rotation(first, delta)
    second = augment_hue_rotation(second, delta)
    return augment_translation((first, second), dx, dy)


def normalize(image):  # dataset_utils.py:39-48
    return ((np.asarray(image, F) / F(127.5)).astype(F) - F(1)).astype(F)
