"""numpy oracle for the palette helpers of the reference (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates `io_utils.py:25-65` (extract_palette), `:78-93` (rgba_to_indexed), `:96-103`
(indexed_to_rgba), the one-hot at `pix2pix_model.py:300-301` and the call-site glue
`dataset_utils.py:131-151`.  Integer work: the bar is bit-exact.

TF semantics assumed (documented behaviour, TF itself is absent here — PARITY UNPINNED):
`UniqueWithCountsV2(axis=[0])` returns rows in order of first occurrence; `argsort(stable=True)` is a
stable ascending sort; `scatter_nd` accumulates duplicate indices and leaves zeros elsewhere;
`one_hot` of an index outside [0, depth) is an all-zero row; the grayness key is a float32
(n,4)x(4,1) product evaluated here as non-fused left-to-right float32.
"""
from __future__ import annotations

import numpy as np

MAX_PALETTE_SIZE = 256  # configuration.py:31
INVALID_INDEX_COLOR = (255, 0, 220, 255)  # configuration.py:32
GRAY_COEFFICIENTS = (0.2989, 0.5870, 0.1140, 0.0)  # io_utils.py:51


class PaletteOverflow(ValueError):
    """More than MAX_PALETTE_SIZE colours: the reference's `tf.repeat` gets a negative count
    (io_utils.py:62) and raises."""


def unique_rows_first_occurrence(rows: np.ndarray) -> np.ndarray:
    rows = np.ascontiguousarray(rows)
    _, first = np.unique(rows, axis=0, return_index=True)
    return rows[np.sort(first)]


def grayness_f32(colors: np.ndarray) -> np.ndarray:
    f = np.float32
    c = colors.astype(f)
    k = [f(v) for v in GRAY_COEFFICIENTS]
    acc = (c[:, 0] * k[0]).astype(f)
    for ch in range(1, c.shape[1]):
        acc = (acc + (c[:, ch] * k[ch]).astype(f)).astype(f)
    return acc


def extract_palette(image, palette_ordering="grayness", channels=4, rng=None):
    """io_utils.py:25-65.  `image` any shape whose size is a multiple of `channels`
    (the caller passes (64,64,8): rows then alternate source/target pixels, dataset_utils.py:142-145)."""
    rows = np.asarray(image).astype(np.int32).reshape(-1, channels)
    if palette_ordering == "top2bottom":
        colors = unique_rows_first_occurrence(rows)
    elif palette_ordering == "bottom2top":
        colors = unique_rows_first_occurrence(rows[::-1])
    elif palette_ordering == "grayness":
        colors = unique_rows_first_occurrence(rows)
        if colors.shape[0] > 1:  # single colour: reference squeezes to rank 0; defined here as identity
            order = np.argsort(grayness_f32(colors), kind="stable")
            colors = colors[order]
    else:  # "shuffled" — io_utils.py:56-58, nondeterministic in the reference
        colors = unique_rows_first_occurrence(rows)
        rng = np.random.default_rng() if rng is None else rng
        colors = colors[rng.permutation(colors.shape[0])]
    n = colors.shape[0]
    if n > MAX_PALETTE_SIZE:
        raise PaletteOverflow(f"{n} unique colours > MAX_PALETTE_SIZE={MAX_PALETTE_SIZE}")
    filler = np.tile(np.asarray(INVALID_INDEX_COLOR, np.int32)[:channels], (MAX_PALETTE_SIZE - n, 1))
    return np.concatenate([colors, filler], axis=0).astype(np.int32), n


def rgba_to_indexed(image, palette):
    """io_utils.py:78-93: idx[n] = sum of every palette row index k whose colour equals pixel n
    (scatter_nd accumulates), 0 when nothing matches."""
    image = np.asarray(image).astype(np.int32)
    palette = np.asarray(palette).astype(np.int32)
    h, w, c = image.shape
    flat = image.reshape(-1, c)
    match = (flat[None, :, :] == palette[:, None, :]).all(-1)  # (K, N)
    ks = np.arange(palette.shape[0], dtype=np.int64)[:, None]
    idx = (match * ks).sum(0).astype(np.int32)
    return idx.reshape(h, w, 1)


def rgba_to_nearest(image, palette):
    """north_star variant: first index of the minimum squared RGBA distance (ties -> lowest k)."""
    image = np.asarray(image).astype(np.int64)
    palette = np.asarray(palette).astype(np.int64)
    h, w, c = image.shape
    flat = image.reshape(-1, c)
    d = ((flat[:, None, :] - palette[None, :, :]) ** 2).sum(-1)
    return np.argmin(d, axis=1).astype(np.int32).reshape(h, w, 1)


def indexed_to_rgba(indexed_image, palette):
    """io_utils.py:96-103."""
    indexed_image = np.asarray(indexed_image)
    palette = np.asarray(palette)
    h, w = indexed_image.shape[:2]
    return palette[indexed_image.reshape(h, w)].reshape(h, w, -1)


def one_hot(indices, depth=MAX_PALETTE_SIZE):
    """`tf.one_hot(idx, depth, axis=-1)` + reshape of pix2pix_model.py:300-301:
    (B,H,W,1) int32 -> (B,H,W,depth) float32; out-of-range index -> all zeros."""
    idx = np.asarray(indices)
    if idx.shape[-1] == 1:
        idx = idx[..., 0]
    out = np.zeros(idx.shape + (depth,), dtype=np.float32)
    ok = (idx >= 0) & (idx < depth)
    pos = np.nonzero(ok)
    out[pos + (idx[ok],)] = 1.0
    return out


def load_indexed_images(source, target, palette_ordering="grayness"):
    """dataset_utils.py:138-151 after PNG decode: shared palette of source||target, two index images."""
    source = np.asarray(source).astype(np.int32)
    target = np.asarray(target).astype(np.int32)
    concatenated = np.concatenate([source, target], axis=-1)
    palette, _ = extract_palette(concatenated, palette_ordering)
    return rgba_to_indexed(source, palette), rgba_to_indexed(target, palette), palette


def blacken_transparent_pixels(image):
    """dataset_utils.py:11-20: alpha == 0 -> whole pixel * 0."""
    image = np.asarray(image)
    mask = image[..., 3:4] == 0
    return np.where(mask, image * 0, image)


def normalize(image):  # dataset_utils.py:39-48
    return (np.asarray(image, np.float32) / np.float32(127.5)) - np.float32(1.0)


def denormalize(image):  # dataset_utils.py:51-60
    return (np.asarray(image, np.float32) + np.float32(1.0)) * np.float32(127.5)


def argmax_indexed(probabilities):
    """`generate` of the indexed model (pix2pix_model.py:283-287): arg-max over the last axis as int32 with a
    trailing singleton axis; first maximum wins, NaNs are never selected (all-NaN / all -inf row -> 0)."""
    p = np.asarray(probabilities, dtype=np.float32)
    clean = np.where(np.isnan(p), -np.inf, p)
    return np.argmax(clean, axis=-1).astype(np.int32)[..., None]


def probabilities_to_rgba(probabilities, palette):
    """`indexed_to_rgba(generate(...), palette)` (pix2pix_model.py:356, 446-447)."""
    idx = argmax_indexed(probabilities)
    pal = np.asarray(palette, dtype=np.int32)
    if idx.ndim == 4 and pal.ndim == 3:
        return np.stack([indexed_to_rgba(idx[b], pal[b]) for b in range(idx.shape[0])])
    if idx.ndim == 4:
        return np.stack([indexed_to_rgba(idx[b], pal) for b in range(idx.shape[0])])
    return indexed_to_rgba(idx, pal)
