"""Generate tests/golden/*.npz from the reference's dataset sprites (TEST INFRASTRUCTURE).

Run in the build container only (needs /root/reference and PIL):

    python -m oracle.make_golden

The GPU box has no /root/reference, so tests read only the committed .npz files.
Sprites are decoded exactly as `dataset_utils.load_image` does (decode_png channels=4,
`blacken_transparent_pixels`, dataset_utils.py:66-77) and stored as uint8.

Fixtures written
  sprites.npz        front/right sprites of train 0..63 and test 0..43 (uint8, (n,64,64,4))
  palette_golden.npz extract_palette/rgba_to_indexed results of the numpy oracle for every
                     front||right pair above, all three deterministic orderings
  hist_golden.npz    float64 histogram / loss / gradient of train/3-right/0..7 vs a seeded
                     tanh-perturbed copy (seed 47 = configuration.py:4), 64 bins
PARITY UNPINNED: goldens come from the numpy restatement, not from TensorFlow.
"""
from __future__ import annotations

import os
import sys

import numpy as np

REF = "/root/reference/datasets/rpg-maker-xp"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def load_sprite(split, direction, number):
    from PIL import Image

    path = os.path.join(REF, split, direction, f"{number}.png")
    img = np.asarray(Image.open(path).convert("RGBA"), dtype=np.uint8)
    assert img.shape == (64, 64, 4), img.shape
    from oracle.palette_oracle import blacken_transparent_pixels

    return blacken_transparent_pixels(img).astype(np.uint8)


def main():
    from oracle import histogram_oracle as ho
    from oracle import palette_oracle as po

    os.makedirs(OUT, exist_ok=True)
    n_train, n_test = 64, 44
    front = np.stack([load_sprite("train", "2-front", i) for i in range(n_train)]
                     + [load_sprite("test", "2-front", i) for i in range(n_test)])
    right = np.stack([load_sprite("train", "3-right", i) for i in range(n_train)]
                     + [load_sprite("test", "3-right", i) for i in range(n_test)])
    np.savez_compressed(os.path.join(OUT, "sprites.npz"), front=front, right=right)

    gold = {}
    for ordering in ("grayness", "top2bottom", "bottom2top"):
        pals, ncols, sidx, tidx = [], [], [], []
        for s, t in zip(front, right):
            cat = np.concatenate([s.astype(np.int32), t.astype(np.int32)], axis=-1)
            pal, n = po.extract_palette(cat, ordering)
            pals.append(pal.astype(np.uint8))
            ncols.append(n)
            sidx.append(po.rgba_to_indexed(s, pal).astype(np.uint8))
            tidx.append(po.rgba_to_indexed(t, pal).astype(np.uint8))
        gold[f"palette_{ordering}"] = np.stack(pals)
        gold[f"ncolors_{ordering}"] = np.asarray(ncols, np.int32)
        gold[f"src_idx_{ordering}"] = np.stack(sidx)
        gold[f"tgt_idx_{ordering}"] = np.stack(tidx)
    np.savez_compressed(os.path.join(OUT, "palette_golden.npz"), **gold)

    # histogram goldens: 8 right-facing sprites vs a tanh-perturbed copy
    rng = np.random.default_rng(47)
    real = po.normalize(right[:8].astype(np.float32))
    noise = rng.standard_normal(real.shape).astype(np.float32)
    fake = np.tanh(np.arctanh(np.clip(real, -0.999, 0.999)) + np.float32(0.35) * noise).astype(np.float32)
    res = ho.hist_loss_and_grad_f64(real, fake, size=64)
    np.savez_compressed(
        os.path.join(OUT, "hist_golden.npz"),
        real=real, fake=fake,
        hist_real=res["hist_real"], hist_fake=res["hist_fake"], denom_fake=res["denom_fake"],
        loss=np.float64(res["loss"]), ssum=np.float64(res["ssum"]), grad=res["grad"],
        hist_real_f32=ho.rgbuv_histogram_f32(real, 64), hist_fake_f32=ho.rgbuv_histogram_f32(fake, 64),
        dom=ho.tf_linspace_f32(-3.0, 3.0, 64),
    )
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    main()
