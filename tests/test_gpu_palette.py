"""GPU parity tests of the palette half, through the C ABI.  Integer work: every comparison is bit-exact."""
import numpy as np
import pytest
import torch

from oracle import palette_oracle as po
from tests.conftest import sprite_like_batch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P():
    import palette_and_histo_gan_b200 as pkg

    return pkg


def dev_i32(a, cuda):
    return torch.from_numpy(np.ascontiguousarray(a).astype(np.int32)).to(cuda)


@pytest.mark.parametrize("ordering", ["grayness", "top2bottom", "bottom2top"])
def test_all_sprite_pairs_match_golden(P, cuda, sprites, palette_golden, ordering):
    src, tgt = dev_i32(sprites["front"], cuda), dev_i32(sprites["right"], cuda)
    s_idx, t_idx, pal = P.dataset_utils.load_indexed_images(src, tgt, ordering)
    assert pal.dtype == torch.int32 and tuple(pal.shape) == (108, 256, 4)
    assert np.array_equal(pal.cpu().numpy(), palette_golden[f"palette_{ordering}"].astype(np.int32))
    assert np.array_equal(s_idx.cpu().numpy()[..., 0], palette_golden[f"src_idx_{ordering}"][..., 0].astype(np.int32))
    assert np.array_equal(t_idx.cpu().numpy()[..., 0], palette_golden[f"tgt_idx_{ordering}"][..., 0].astype(np.int32))
    # round trip (how pix2pix_model.py:446-447 reconstructs targets)
    back = P.io_utils.indexed_to_rgba(t_idx, pal)
    assert torch.equal(back, tgt)
    # the separate calls of dataset_utils.py:142-149 give the same result as the fused glue
    cat = torch.cat([src, tgt], dim=-1)
    pal2, ncol = P.io_utils.extract_palette(cat, ordering, return_counts=True)
    assert torch.equal(pal2, pal)
    assert np.array_equal(ncol.cpu().numpy(), palette_golden[f"ncolors_{ordering}"])
    assert torch.equal(P.io_utils.rgba_to_indexed(src, pal2), s_idx)


def test_single_image_signatures(P, cuda, sprites):
    """Un-batched calls with the reference's exact signatures and shapes."""
    s, t = sprites["front"][3].astype(np.int32), sprites["right"][3].astype(np.int32)
    cat = dev_i32(np.concatenate([s, t], -1), cuda)  # (64,64,8) as dataset_utils.py:142 builds it
    pal = P.io_utils.extract_palette(cat, "grayness")
    epal, _ = po.extract_palette(np.concatenate([s, t], -1), "grayness")
    assert tuple(pal.shape) == (256, 4) and np.array_equal(pal.cpu().numpy(), epal)
    idx = P.io_utils.rgba_to_indexed(dev_i32(s, cuda), pal)
    assert tuple(idx.shape) == (64, 64, 1) and np.array_equal(idx.cpu().numpy(), po.rgba_to_indexed(s, epal))
    rgba = P.io_utils.indexed_to_rgba(idx, pal)
    assert tuple(rgba.shape) == (64, 64, 4) and np.array_equal(rgba.cpu().numpy(), s)


def test_grayness_ties_and_stable_sort(P, cuda):
    # alpha-only ties and the RGB tie d(r,g,b) = (10,-35,154) from SURVEY.md §4
    base = np.array([100, 100, 50, 255])
    tie = base + np.array([10, -35, 154, 0])
    imgs = []
    for order in ([base, tie], [tie, base]):
        px = [[0, 0, 0, 255], order[0], [0, 0, 0, 0], order[1], [0, 0, 0, 7], [255, 255, 255, 255]]
        imgs.append(np.array(px, np.int32).reshape(1, 6, 4))
    batch = np.stack(imgs)
    for ordering in ("grayness", "top2bottom", "bottom2top"):
        pal = P.io_utils.extract_palette(dev_i32(batch, cuda), ordering).cpu().numpy()
        for i in range(2):
            assert np.array_equal(pal[i], po.extract_palette(batch[i], ordering)[0]), ordering


def test_random_images_many_colours(P, cuda):
    rng = np.random.default_rng(5)
    # exactly 256 colours (full palette), 255, 1, and a ragged row count not divisible by the block size
    cases = []
    for ncol, npx in ((256, 64 * 64), (255, 50 * 30), (1, 17), (37, 1000), (200, 8191)):
        cols = rng.integers(0, 256, size=(ncol, 4))
        cols = np.unique(cols, axis=0)
        while cols.shape[0] < ncol:
            cols = np.unique(np.concatenate([cols, rng.integers(0, 256, size=(ncol - cols.shape[0], 4))]), axis=0)
        rng.shuffle(cols)
        px = np.concatenate([cols, cols[rng.integers(0, ncol, size=npx - ncol)]]) if npx > ncol else cols[:npx]
        rng.shuffle(px)
        cases.append(px.astype(np.int32).reshape(1, -1, 4))
    for img in cases:
        for ordering in ("grayness", "top2bottom", "bottom2top"):
            pal, n = P.io_utils.extract_palette(dev_i32(img, cuda), ordering, return_counts=True)
            epal, en = po.extract_palette(img, ordering)
            assert int(n) == en and np.array_equal(pal.cpu().numpy(), epal)
            idx = P.io_utils.rgba_to_indexed(dev_i32(img, cuda), pal)
            assert np.array_equal(idx.cpu().numpy(), po.rgba_to_indexed(img, epal))


def test_overflow_and_bad_values_raise(P, cuda):
    k = np.arange(300)
    many = np.stack([k % 256, k // 256, np.zeros(300), np.full(300, 255)], -1).astype(np.int32).reshape(1, 300, 4)
    with pytest.raises(P.io_utils.PaletteOverflowError):
        P.io_utils.extract_palette(dev_i32(many, cuda), "grayness")
    noise = np.random.default_rng(0).integers(0, 256, size=(2, 64, 64, 4)).astype(np.int32)
    with pytest.raises(P.io_utils.PaletteOverflowError):
        P.io_utils.extract_palette(dev_i32(noise, cuda), "top2bottom")
    bad = np.array([[[0, 0, 0, 255], [0, 300, 0, 255]]], np.int32)
    with pytest.raises(ValueError):
        P.io_utils.extract_palette(dev_i32(bad, cuda), "grayness")


def test_scatter_add_edge_cases(P, cuda):
    """io_utils.py:84-91: duplicates add up, no match gives 0; out-of-range index one-hot is all zero."""
    pal, _ = po.extract_palette(np.array([[[1, 2, 3, 255], [4, 5, 6, 255]]], np.int32), "top2bottom")
    img = np.array([[[255, 0, 220, 255], [4, 5, 6, 255], [7, 7, 7, 7], [1, 2, 3, 255], [-5, 2, 3, 255], [1, 2, 3, 256]]], np.int32)
    idx, oh = P.io_utils.rgba_to_indexed(dev_i32(img, cuda), dev_i32(pal, cuda), with_one_hot=True)
    exp = po.rgba_to_indexed(img, pal)
    assert np.array_equal(idx.cpu().numpy(), exp)
    assert exp.reshape(-1).tolist() == [sum(range(2, 256)), 1, 0, 0, 0, 0]
    assert np.array_equal(oh.cpu().numpy(), po.one_hot(exp))
    # palette rows outside the byte range still match exactly (any int32 follows the reference)
    wide = pal.copy()
    wide[5] = [1000, -3, 70000, 255]
    img2 = np.array([[[1000, -3, 70000, 255], [4, 5, 6, 255]]], np.int32)
    idx2 = P.io_utils.rgba_to_indexed(dev_i32(img2, cuda), dev_i32(wide, cuda))
    assert np.array_equal(idx2.cpu().numpy(), po.rgba_to_indexed(img2, wide))
    # nearest mode: equals exact mode on the reference's domain, differs by design off it
    near = P.io_utils.rgba_to_indexed(dev_i32(img, cuda), dev_i32(pal, cuda), mode="nearest")
    assert np.array_equal(near.cpu().numpy(), po.rgba_to_nearest(img, pal))


def test_one_hot_shapes(P, cuda):
    rng = np.random.default_rng(6)
    idx = rng.integers(-3, 260, size=(2, 9, 7, 1)).astype(np.int32)
    out = P.io_utils.one_hot(dev_i32(idx, cuda))
    assert tuple(out.shape) == (2, 9, 7, 256) and out.dtype == torch.float32
    assert np.array_equal(out.cpu().numpy(), po.one_hot(idx))
    out10 = P.io_utils.one_hot(dev_i32(idx % 10, cuda), depth=10)  # depth not a multiple of 4
    assert np.array_equal(out10.cpu().numpy(), po.one_hot(idx % 10, 10))


def test_cfgB_full_size_properties(P, cuda):
    """cfgB: batch 256 of 64x64 pairs.  Bit-exact against the oracle on 16 spot images, plus the
    round-trip / one-hot-sum properties on all of them."""
    rng = np.random.default_rng(47)
    src = sprite_like_batch(rng, 256).astype(np.int32)
    tgt = src.copy()
    tgt[:, :, ::2] = src[:, :, 1::2]  # target shares most colours with the source, like a real pair
    s_idx, t_idx, pal = P.dataset_utils.load_indexed_images(dev_i32(src, cuda), dev_i32(tgt, cuda), "grayness")
    assert torch.equal(P.io_utils.indexed_to_rgba(s_idx, pal), dev_i32(src, cuda))
    assert torch.equal(P.io_utils.indexed_to_rgba(t_idx, pal), dev_i32(tgt, cuda))
    idx2, oh = P.io_utils.rgba_to_indexed(dev_i32(tgt, cuda), pal, with_one_hot=True)
    assert torch.equal(idx2, t_idx)
    assert float(oh.sum()) == 256 * 64 * 64
    assert torch.equal(oh.argmax(-1, keepdim=True).to(torch.int32), t_idx)
    for i in range(0, 256, 16):
        es, et, ep = po.load_indexed_images(src[i], tgt[i], "grayness")
        assert np.array_equal(pal[i].cpu().numpy(), ep) and np.array_equal(s_idx[i].cpu().numpy(), es)
        assert np.array_equal(t_idx[i].cpu().numpy(), et)


def test_shuffled_ordering_is_a_permutation(P, cuda, sprites):
    s, t = sprites["front"][:4].astype(np.int32), sprites["right"][:4].astype(np.int32)
    s_idx, t_idx, pal = P.dataset_utils.load_indexed_images(dev_i32(s, cuda), dev_i32(t, cuda), "shuffled")
    for i in range(4):
        ep, n = po.extract_palette(np.concatenate([s[i], t[i]], -1), "top2bottom")
        got = pal[i].cpu().numpy()
        assert sorted(map(tuple, got[:n])) == sorted(map(tuple, ep[:n])) and (got[n:] == ep[n:]).all()
    assert torch.equal(P.io_utils.indexed_to_rgba(s_idx, pal), dev_i32(s, cuda))


def test_host_api_matches_device_api(P, cuda, sprites):
    s, t = sprites["front"][:8].astype(np.int32), sprites["right"][:8].astype(np.int32)
    hs, ht, hp, oh = P.hostapi.load_indexed_images(s, t, "grayness", with_one_hot=True)
    for i in range(8):
        es, et, ep = po.load_indexed_images(s[i], t[i], "grayness")
        assert np.array_equal(hs[i], es) and np.array_equal(ht[i], et) and np.array_equal(hp[i], ep)
        assert np.array_equal(oh[i], po.one_hot(et))
    # the decoded PNG as it is (uint8): a quarter of the upload, identical outputs; numpy or pinned torch tensors
    s8, t8 = sprites["front"][:8].astype(np.uint8), sprites["right"][:8].astype(np.uint8)
    us, ut, up = P.hostapi.load_indexed_images(s8, t8, "grayness")
    assert np.array_equal(us, hs) and np.array_equal(ut, ht) and np.array_equal(up, hp)
    us, ut, up = P.hostapi.load_indexed_images(torch.from_numpy(s8).pin_memory(), torch.from_numpy(t8).pin_memory(), "top2bottom")
    es, et, ep = po.load_indexed_images(s[3], t[3], "top2bottom")
    assert np.array_equal(us[3], es) and np.array_equal(ut[3], et) and np.array_equal(up[3], ep)
    with pytest.raises(TypeError):
        P.hostapi.load_indexed_images(s8, t, "grayness")
    # caller-owned (pinned) result buffers
    outs = (torch.empty((8, 64, 64, 1), dtype=torch.int32).pin_memory(), torch.empty((8, 64, 64, 1), dtype=torch.int32).pin_memory(),
            torch.empty((8, 256, 4), dtype=torch.int32).pin_memory())
    rs, rt, rp = P.hostapi.load_indexed_images(s8, t8, "grayness", out=outs)
    assert np.array_equal(outs[0].numpy(), hs) and np.array_equal(outs[1].numpy(), ht) and np.array_equal(outs[2].numpy(), hp)
    assert np.shares_memory(rs, outs[0].numpy())
    with pytest.raises(ValueError):
        P.hostapi.load_indexed_images(s8, t8, "grayness", out=(outs[0], outs[1], torch.empty((8, 255, 4), dtype=torch.int32)))


def test_fused_loader_filler_colour_and_batching(P, cuda):
    """The one-launch loader (dataset_utils.py:138-151) must keep the scatter-add semantics: a pixel equal
    to INVALID_INDEX_COLOR also matches every padding row (io_utils.py:84-91)."""
    rng = np.random.default_rng(8)
    cols = np.array([[0, 0, 0, 0], [255, 0, 220, 255], [10, 20, 30, 255], [255, 255, 255, 255], [9, 9, 9, 9]], np.int32)
    src = cols[rng.integers(0, 5, size=(3, 16, 16))]
    tgt = cols[rng.integers(0, 4, size=(3, 16, 16))]
    for ordering in ("grayness", "top2bottom", "bottom2top"):
        s_idx, t_idx, pal = P.dataset_utils.load_indexed_images(dev_i32(src, cuda), dev_i32(tgt, cuda), ordering)
        for i in range(3):
            es, et, ep = po.load_indexed_images(src[i], tgt[i], ordering)
            assert np.array_equal(pal[i].cpu().numpy(), ep), ordering
            assert np.array_equal(s_idx[i].cpu().numpy(), es) and np.array_equal(t_idx[i].cpu().numpy(), et), ordering
    # un-batched call keeps the reference shapes
    s1, t1, p1 = P.dataset_utils.load_indexed_images(dev_i32(src[0], cuda), dev_i32(tgt[0], cuda), "grayness")
    assert tuple(s1.shape) == (16, 16, 1) and tuple(p1.shape) == (256, 4)


def test_probabilities_to_indexed_and_rgba(P, cuda, sprites):
    """f3 of SURVEY.md §8f: argmax over the softmax channels + palette gather in one kernel, bit-exact."""
    rng = np.random.default_rng(5)
    src, tgt = sprites["front"][:6], sprites["right"][:6]
    _, t_idx, pal = P.dataset_utils.load_indexed_images(dev_i32(src, cuda), dev_i32(tgt, cuda), "grayness")
    # probabilities whose arg-max is the target index, with exact ties, NaNs and -inf sprinkled in
    probs = rng.random((6, 64, 64, 256), dtype=np.float32) * 0.5
    ti = t_idx.cpu().numpy()[..., 0]
    np.put_along_axis(probs, ti[..., None].astype(np.int64), 0.75, axis=-1)
    probs[0, 0, :, 200] = 0.75          # tie with a later (or earlier) channel: first maximum wins
    probs[0, 1, :, 3] = np.nan          # NaN is never selected
    probs[0, 2, :, :] = np.nan          # all-NaN row -> 0
    probs[0, 3, :, :] = -np.inf         # all -inf row -> 0
    probs[0, 4, :, 0] = np.nan
    want_idx = po.argmax_indexed(probs)
    want_rgba = po.probabilities_to_rgba(probs, pal.cpu().numpy())
    got_idx, got_rgba = P.io_utils.probabilities_to_indexed(torch.from_numpy(probs).to(cuda), pal)
    assert got_idx.dtype == torch.int32 and tuple(got_idx.shape) == (6, 64, 64, 1)
    assert np.array_equal(got_idx.cpu().numpy(), want_idx)
    assert np.array_equal(got_rgba.cpu().numpy(), want_rgba)
    # on clean rows the round trip reproduces the target sprite
    assert np.array_equal(got_rgba.cpu().numpy()[1:], tgt[1:].astype(np.int32))
    # index-only call, single image, shared palette, depth that is not a multiple of 128
    only = P.io_utils.probabilities_to_indexed(torch.from_numpy(probs[1]).to(cuda))
    assert np.array_equal(only.cpu().numpy(), want_idx[1])
    odd = rng.standard_normal((2, 5, 7, 37)).astype(np.float32)
    odd[0, 0, 0, 5] = odd[0, 0, 0].max() + 1.0
    odd[0, 0, 0, 30] = odd[0, 0, 0, 5]
    pal1 = pal[0]
    gi, gr = P.io_utils.probabilities_to_indexed(torch.from_numpy(odd).to(cuda), pal1)
    assert np.array_equal(gi.cpu().numpy(), po.argmax_indexed(odd))
    assert np.array_equal(gr.cpu().numpy(), po.probabilities_to_rgba(odd, pal1.cpu().numpy()))


@pytest.mark.parametrize("ordering", ["grayness", "top2bottom", "bottom2top"])
def test_against_the_reference_source_run(P, cuda, reference_run, ordering):
    """Bit-exact against the outputs of the reference's own dataset_utils.load_indexed_images / io_utils.py
    (executed over oracle/ref_shim.py on the dataset's PNG files)."""
    R = reference_run
    src, tgt = dev_i32(R["loader_source"], cuda), dev_i32(R["loader_target"], cuda)
    s_idx, t_idx, pal = P.dataset_utils.load_indexed_images(src, tgt, ordering)
    assert np.array_equal(pal.cpu().numpy(), R[f"palette_{ordering}"])
    assert np.array_equal(s_idx.cpu().numpy(), R[f"src_idx_{ordering}"])
    assert np.array_equal(t_idx.cpu().numpy(), R[f"tgt_idx_{ordering}"])
    if ordering == "grayness":
        n = R["roundtrip_rgba"].shape[0]
        assert np.array_equal(P.io_utils.indexed_to_rgba(t_idx[:n], pal[:n]).cpu().numpy(), R["roundtrip_rgba"])
        oh = P.io_utils.one_hot(t_idx[:2]).cpu().numpy()[:, ::8, ::8]
        assert np.array_equal(oh.reshape(R["one_hot_rows"].shape), R["one_hot_rows"])
        # the uint8 loader prep (blacken + normalise) against the reference's normalize()
        u8 = torch.from_numpy(R["loader_target"][:n]).to(cuda)
        assert np.array_equal(P.dataset_utils.load_image(u8).cpu().numpy(), R["normalized"])


def test_concurrent_host_threads(P, cuda, sprites, palette_golden):
    """The palette ops run inside `tf.data.map(num_parallel_calls=AUTOTUNE)` worker threads in the reference
    (dataset_utils.py:236-244): several host threads call the library at once, each on its own stream.  The entry
    points are re-entrant (caller-owned outputs, no global scratch, thread-local error text)."""
    import threading

    front, right = sprites["front"], sprites["right"]
    n = front.shape[0]
    results, errors = {}, []

    def worker(tid):
        try:
            stream = torch.cuda.Stream(device=cuda)
            with torch.cuda.stream(stream):
                for rep in range(4):
                    lo = (tid * 7 + rep * 13) % (n - 4)
                    src, tgt = dev_i32(front[lo:lo + 4], cuda), dev_i32(right[lo:lo + 4], cuda)
                    s_idx, t_idx, pal = P.dataset_utils.load_indexed_images(src, tgt, "grayness")
                    oh = P.io_utils.one_hot(t_idx)
                    back = P.io_utils.indexed_to_rgba(t_idx, pal)
                    stream.synchronize()
                    results[(tid, rep)] = (lo, s_idx.cpu().numpy(), t_idx.cpu().numpy(), pal.cpu().numpy(),
                                           oh.sum().item(), back.cpu().numpy())
                # an error raised in one thread must not leak its message into another
                if tid == 0:
                    with pytest.raises(Exception):
                        P.io_utils.extract_palette(torch.randint(0, 256, (64, 64, 4), device=cuda, dtype=torch.int32))
        except Exception as exc:  # noqa: BLE001
            errors.append((tid, repr(exc)))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(6)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    assert len(results) == 24
    for (tid, rep), (lo, s_idx, t_idx, pal, oh_sum, back) in results.items():
        for k in range(4):
            ep, _ = po.extract_palette(np.concatenate([front[lo + k], right[lo + k]], -1).astype(np.int32), "grayness")
            assert np.array_equal(pal[k], ep)
            assert np.array_equal(s_idx[k], po.rgba_to_indexed(front[lo + k].astype(np.int32), ep))
            assert np.array_equal(t_idx[k], po.rgba_to_indexed(right[lo + k].astype(np.int32), ep))
            assert np.array_equal(back[k], right[lo + k].astype(np.int32))
        assert oh_sum == 4 * 64 * 64


def test_shuffled_ordering_in_the_kernel_is_seeded_and_consistent(P, cuda, sprites):
    """`palette_ordering="shuffled"` (io_utils.py:56-58) runs inside the fused kernel: the first-occurrence colours are
    permuted by ranking seeded uniform keys.  Same generator seed -> same palettes; the indices refer to the shuffled
    palette (round trip); the permutation is the argsort of the keys the host drew; host API likewise with `seed=`."""
    s, t = sprites["front"][:6].astype(np.int32), sprites["right"][:6].astype(np.int32)
    runs = []
    for _ in range(2):
        g = torch.Generator().manual_seed(123)
        runs.append(P.dataset_utils.load_indexed_images(dev_i32(s, cuda), dev_i32(t, cuda), "shuffled", generator=g))
    assert all(torch.equal(a, b) for a, b in zip(runs[0], runs[1]))
    s_idx, t_idx, pal = runs[0]
    keys = torch.rand((6, 256), generator=torch.Generator().manual_seed(123), dtype=torch.float32).numpy()
    moved = 0
    for i in range(6):
        ep, n = po.extract_palette(np.concatenate([s[i], t[i]], -1), "top2bottom")
        expect = ep.copy()
        expect[:n] = ep[:n][np.argsort(keys[i, :n], kind="stable")]
        assert np.array_equal(pal[i].cpu().numpy(), expect)
        moved += int(not np.array_equal(expect, ep))
    assert moved >= 5  # it is a real permutation
    assert torch.equal(P.io_utils.indexed_to_rgba(s_idx, pal), dev_i32(s, cuda))
    assert torch.equal(P.io_utils.indexed_to_rgba(t_idx, pal), dev_i32(t, cuda))
    # standalone extract_palette and the host-buffer API
    cat = torch.cat([dev_i32(s, cuda), dev_i32(t, cuda)], dim=-1)
    pal2 = P.io_utils.extract_palette(cat, "shuffled", generator=torch.Generator().manual_seed(123))
    assert torch.equal(pal2, pal)
    hs, ht, hp = P.hostapi.load_indexed_images(s, t, "shuffled", seed=7)
    hs2, ht2, hp2 = P.hostapi.load_indexed_images(s.astype(np.uint8), t.astype(np.uint8), "shuffled", seed=7)
    assert np.array_equal(hp, hp2) and np.array_equal(hs, hs2) and np.array_equal(ht, ht2)
    hkeys = np.random.default_rng(7).random((6, 256), dtype=np.float32)
    for i in range(6):
        ep, n = po.extract_palette(np.concatenate([s[i], t[i]], -1), "top2bottom")
        expect = ep.copy()
        expect[:n] = ep[:n][np.argsort(hkeys[i, :n], kind="stable")]
        assert np.array_equal(hp[i], expect)
        assert np.array_equal(po.indexed_to_rgba(hs[i], hp[i]), s[i])


@pytest.mark.parametrize("ordering", ["grayness", "top2bottom", "bottom2top"])
def test_uint8_device_entry_and_large_images(P, cuda, sprites, ordering):
    """`ph_load_indexed_images_u8`: the decoded PNG's uint8 pixels read by the kernel as they are — identical outputs
    to the int32 entry.  Images above 8 192 rows per pair take the kernel variant that re-reads the pixels for the
    index pass instead of keeping their keys in registers; ragged sizes exercise the partially filled batches."""
    s8, t8 = torch.from_numpy(sprites["front"][:20]).to(cuda), torch.from_numpy(sprites["right"][:20]).to(cuda)
    a = P.dataset_utils.load_indexed_images(s8, t8, ordering)
    b = P.dataset_utils.load_indexed_images(s8.to(torch.int32), t8.to(torch.int32), ordering)
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    rng = np.random.default_rng(11)
    for hw in ((96, 96), (70, 61), (3, 5)):
        src = sprite_like_batch(rng, 3, hw=96)[:, :hw[0], :hw[1]].copy()
        tgt = src[:, ::-1].copy()
        for dt in (torch.uint8, torch.int32):
            s_idx, t_idx, pal = P.dataset_utils.load_indexed_images(torch.from_numpy(src).to(cuda).to(dt),
                                                                    torch.from_numpy(tgt).to(cuda).to(dt), ordering)
            for i in range(3):
                es, et, ep = po.load_indexed_images(src[i].astype(np.int32), tgt[i].astype(np.int32), ordering)
                assert np.array_equal(pal[i].cpu().numpy(), ep), (hw, dt)
                assert np.array_equal(s_idx[i].cpu().numpy(), es) and np.array_equal(t_idx[i].cpu().numpy(), et), (hw, dt)


def test_pixel_helpers_are_kernels_and_bit_exact(P, cuda, sprites):
    """dataset_utils.py:11-20, :39-60 as `ph_pixel_map` launches (P6): bit-exact against the numpy restatement, the
    launch counter proves they ran in libpalhist, CPU tensors are refused."""
    D = P.dataset_utils
    raw = sprites["right"][:5].copy()
    raw[0, :3, :3] = [200, 10, 30, 0]                      # non-black transparent pixels
    x = torch.from_numpy(raw.astype(np.float32)).to(cuda)
    P._lib.reset_launch_count()
    blk = D.blacken_transparent_pixels(x)
    nrm = D.normalize(blk)
    den = D.denormalize(nrm)
    assert P._lib.launch_count() == 3
    eb = po.blacken_transparent_pixels(raw.astype(np.float32))
    assert np.array_equal(blk.cpu().numpy(), eb)
    assert np.array_equal(nrm.cpu().numpy(), po.normalize(eb))
    assert np.array_equal(den.cpu().numpy(), po.denormalize(po.normalize(eb)))
    # integer input keeps its dtype through blacken (tf.where), odd element counts through normalize
    bi = D.blacken_transparent_pixels(torch.from_numpy(raw).to(cuda))
    assert bi.dtype == torch.uint8 and np.array_equal(bi.cpu().numpy(), po.blacken_transparent_pixels(raw))
    odd = torch.arange(0, 7, dtype=torch.float32, device=cuda) * 36.5
    assert np.array_equal(D.normalize(odd).cpu().numpy(), po.normalize(odd.cpu().numpy()))
    neg0 = torch.tensor([[5.0, 6.0, 7.0, -0.0]], device=cuda)
    assert float(D.blacken_transparent_pixels(neg0).abs().sum()) == 0.0       # -0.0 == 0 as in tf.where
    with pytest.raises(ValueError):
        D.normalize(torch.zeros(4, 4, 4))
    # same result as the fused uint8 loader
    assert torch.equal(D.load_image(torch.from_numpy(raw).to(cuda)), nrm)
