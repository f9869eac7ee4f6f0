"""Run the reference's OWN source files on the fixture inputs (TEST INFRASTRUCTURE, build container only).

    python -m oracle.run_reference            # writes tests/golden/reference_run.npz

/root/reference/histogram.py, io_utils.py and dataset_utils.py are imported unmodified with `oracle/ref_shim.py`
standing in for the `tensorflow` module (TensorFlow itself is not installable here).  Every line of the
reference's Python executes as written — axis conventions, transposes, broadcasting, the palette / index logic,
the loader glue including PNG decoding and `blacken_transparent_pixels` — and, the shim being backed by torch
tensors, the gradient of the generator-loss term is taken by autograd through the reference's own forward code
(the analogue of `tape.gradient`, pix2pix_model.py:78).  What remains assumed is the behaviour of the individual
TensorFlow ops (listed in ref_shim.py).  The outputs are committed as a fixture; tests/test_oracle.py pins the
numpy/float64 oracle against them and the GPU tests compare the CUDA path with them — nothing reads
/root/reference at test time.
"""
from __future__ import annotations

import os
import sys

import numpy as np

REFERENCE = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
N_LOADER = 40   # sprite pairs pushed through the reference's load_indexed_images
N_INDEX_ONLY = 12


def main():
    import torch

    from oracle import ref_shim

    ref_shim.install(REFERENCE)
    cwd = os.getcwd()
    os.chdir(REFERENCE)  # the loader builds relative dataset paths (configuration.py:6)
    try:
        import dataset_utils as ref_dataset_utils  # noqa: E402  (the reference's modules)
        import histogram as ref_histogram  # noqa: E402
        import io_utils as ref_io_utils  # noqa: E402
        from configuration import DATA_FOLDERS, DIRECTION_FRONT, DIRECTION_RIGHT, DATASET_SIZES

        out = {}
        # ---- palette half: the reference's loader closure, dataset_utils.py:131-151, on real PNG files ----
        for ordering in ("grayness", "top2bottom", "bottom2top"):
            loader = ref_dataset_utils.create_indexed_image_loader(DIRECTION_FRONT, DIRECTION_RIGHT, DATASET_SIZES,
                                                                   "train", ordering)
            load_indexed_images = loader.__closure__ and [c.cell_contents for c in loader.__closure__
                                                          if callable(c.cell_contents)
                                                          and getattr(c.cell_contents, "__name__", "") == "load_indexed_images"][0]
            src_idx, tgt_idx, pals = [], [], []
            for n in range(N_LOADER):
                s, t, p = load_indexed_images(DATA_FOLDERS[0], str(n))
                src_idx.append(s.numpy()); tgt_idx.append(t.numpy()); pals.append(p.numpy())
            out[f"src_idx_{ordering}"] = np.stack(src_idx).astype(np.int32)
            out[f"tgt_idx_{ordering}"] = np.stack(tgt_idx).astype(np.int32)
            out[f"palette_{ordering}"] = np.stack(pals).astype(np.int32)
        # the decoded inputs of those calls (load_image with should_normalize=False, cast to int32)
        src_img, tgt_img = [], []
        for n in range(N_LOADER):
            for side, acc in ((2, src_img), (3, tgt_img)):
                path = os.path.join(DATA_FOLDERS[0], "train", f"{side}-{'front' if side == 2 else 'right'}", f"{n}.png")
                acc.append(ref_dataset_utils.load_image(path, should_normalize=False).numpy().astype(np.uint8))
        out["loader_source"] = np.stack(src_img)
        out["loader_target"] = np.stack(tgt_img)
        # round trip + normalisation helpers on the first few
        rgba = [ref_io_utils.indexed_to_rgba(ref_shim.T(out["tgt_idx_grayness"][n]), ref_shim.T(out["palette_grayness"][n])).numpy()
                for n in range(N_INDEX_ONLY)]
        out["roundtrip_rgba"] = np.stack(rgba).astype(np.int32)
        norm = ref_dataset_utils.normalize(ref_shim.T(out["loader_target"][:N_INDEX_ONLY].astype(np.float32)))
        out["normalized"] = norm.numpy().astype(np.float32)
        out["denormalized"] = ref_dataset_utils.denormalize(norm).numpy().astype(np.float32)
        # one-hot as Pix2PixIndexedModel.train_step builds it (pix2pix_model.py:300-301)
        import tensorflow as tf  # the shim

        idx = ref_shim.T(out["tgt_idx_grayness"][:2])
        oh = tf.reshape(tf.one_hot(idx, 256, axis=-1), [2, 64, 64, -1])
        out["one_hot_rows"] = oh.numpy()[:, ::8, ::8].astype(np.float32)  # a 8x8 sub-grid of pixels keeps the file small

        # ---- histogram half: calculate_rgbuv_histogram + hellinger_loss, and autograd through them ----
        gold = np.load(os.path.join(OUT, "hist_golden.npz"))
        real = ref_shim.T(gold["real"].astype(np.float32))
        fake = ref_shim.T(gold["fake"].astype(np.float32)).requires_grad_(True)
        h_real = ref_histogram.calculate_rgbuv_histogram(real)
        h_fake = ref_histogram.calculate_rgbuv_histogram(fake)
        loss = ref_histogram.hellinger_loss(h_real, h_fake)
        (grad,) = torch.autograd.grad(loss, fake)
        out["hist_real"] = h_real.detach().numpy().astype(np.float32)
        out["hist_fake"] = h_fake.detach().numpy().astype(np.float32)
        out["loss"] = np.float32(loss.detach().numpy())
        out["grad"] = grad.numpy().astype(np.float32)
        out["l1"] = np.float32(ref_histogram.l1_loss(h_real, h_fake).detach().numpy())
        out["l2"] = np.float32(ref_histogram.l2_loss(h_real, h_fake).detach().numpy())
        # other parameters of the reference signature
        rng = np.random.default_rng(47)
        dense = ref_shim.T(np.tanh(rng.standard_normal((2, 16, 16, 4))).astype(np.float32))
        out["dense_input"] = dense.numpy()
        out["dense_hist_32_iq"] = ref_histogram.calculate_rgbuv_histogram(dense, size=32).numpy()
        out["dense_hist_64_rbf"] = ref_histogram.calculate_rgbuv_histogram(dense, size=64, method="RBF", sigma=0.5).numpy()
        r, g, b = [dense.reshape(2, -1, 4)[..., c] * 0.5 + 0.5 for c in range(3)]
        iy = tf.sqrt(r * r + g * g + b * b + 1e-6)[..., None]
        dom = tf.expand_dims(tf.linspace(-3., 3., num=16), 0)
        out["component_hist"] = ref_histogram.calculate_component_histogram(r, g, b, iy, dom, "inverse-quadratic",
                                                                            tf.pow(0.02, 2), 1e-6).numpy()
        out["linspace_64"] = tf.linspace(-3., 3., num=64).numpy()
    finally:
        os.chdir(cwd)
    path = os.path.join(OUT, "reference_run.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


N_AUGMENT = 16  # sprite pairs pushed through the reference's augment_two


def main_augment():
    """dataset_utils.py:80-102 (`augment_two` = hue rotation of both images + one shared translation) run from the
    reference's own source on real sprite pairs -> tests/golden/reference_augment.npz, with the random draws the
    stand-ins made (hue delta, translation in pixels) stored next to the outputs."""
    from oracle import ref_shim

    ref_shim.install(REFERENCE)
    cwd = os.getcwd()
    os.chdir(REFERENCE)
    try:
        import dataset_utils as ref_dataset_utils  # noqa: E402
        from configuration import DATA_FOLDERS

        ref_shim._RNG = np.random.default_rng(47)
        firsts, seconds, out_first, out_second, deltas, trans = [], [], [], [], [], []
        for n in range(N_AUGMENT):
            imgs = []
            for side, name in ((2, "front"), (3, "right")):
                path = os.path.join(DATA_FOLDERS[0], "train", f"{side}-{name}", f"{n}.png")
                imgs.append(ref_dataset_utils.load_image(path, should_normalize=False))
            del ref_shim.DRAWS[:]
            a, b = ref_dataset_utils.augment_two(imgs[0], imgs[1])
            draws = dict((k, v) for k, v in ref_shim.DRAWS)  # both hue calls share the seed, hence the delta
            assert [k for k, _ in ref_shim.DRAWS] == ["hue_delta", "hue_delta", "translation"]
            assert ref_shim.DRAWS[0][1] == ref_shim.DRAWS[1][1]
            firsts.append(imgs[0].numpy()); seconds.append(imgs[1].numpy())
            out_first.append(a.numpy()); out_second.append(b.numpy())
            deltas.append(draws["hue_delta"]); trans.append(draws["translation"])
        out = {"first": np.stack(firsts).astype(np.float32), "second": np.stack(seconds).astype(np.float32),
               "out_first": np.stack(out_first).astype(np.float32), "out_second": np.stack(out_second).astype(np.float32),
               "hue_delta": np.asarray(deltas, np.float32), "translation": np.asarray(trans, np.float32)}
        out["normalized_first"] = np.stack([np.asarray(t.numpy()) for t in
                                            (ref_dataset_utils.normalize_two(ref_shim.T(a), ref_shim.T(b))[0]
                                             for a, b in zip(out["out_first"][:4], out["out_second"][:4]))])
    finally:
        os.chdir(cwd)
    path = os.path.join(OUT, "reference_augment.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "augment":
        sys.exit(main_augment())
    sys.exit(main())
