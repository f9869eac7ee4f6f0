"""CPU tests of the boundary: the C-ABI library loads, exports every symbol include/palhist.h declares,
the ctypes prototypes cover exactly that set, and the host layer validates arguments and refuses to
compute without a GPU (no compute calls are made here)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "palhist.h")


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"PH_API[^;(]*?\b(ph_\w+)\s*\(", text)))


@pytest.fixture(scope="module")
def pkg():
    import __graft_entry__ as g

    if not os.path.exists(os.path.join(ROOT, "palette_and_histo_gan_b200", "libpalhist.so")):
        g.build()
    import palette_and_histo_gan_b200 as p

    return p


def test_header_declares_expected_surface():
    syms = declared_symbols()
    for name in ("ph_hist_forward", "ph_hist_forward_ssum", "ph_hist_backward", "ph_hellinger_ssum", "ph_extract_palette",
                 "ph_rgba_to_indexed", "ph_one_hot", "ph_indexed_to_rgba", "ph_argmax_indexed", "ph_load_indexed_images",
                 "ph_host_hist_loss", "ph_host_load_indexed_images"):
        assert name in syms
    # every entry point cites the reference code it replaces
    text = open(HEADER).read()
    for ref in ("histogram.py:36-81", "histogram.py:5-32", "histogram.py:84-89", "io_utils.py:25-65",
                "io_utils.py:78-93", "io_utils.py:96-103", "pix2pix_model.py:300-301",
                "dataset_utils.py:138-151", "dataset_utils.py:80-102"):
        assert ref in text, ref


def test_library_exports_every_declared_symbol(pkg):
    lib = ctypes.CDLL(pkg._lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in palhist.h but not exported"
    assert lib.ph_abi_version() == pkg._lib.ABI_VERSION == 3


def test_ctypes_prototypes_match_header(pkg):
    assert sorted(pkg._lib.PROTOTYPES) == declared_symbols()
    # argument counts AND kinds agree with the header's parameter lists
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)

    def kind(ctype):
        if ctype in (ctypes.c_void_p, ctypes.c_char_p) or hasattr(ctype, "contents"):
            return "ptr"
        return {ctypes.c_int64: "int64_t", ctypes.c_int: "int", ctypes.c_float: "float",
                ctypes.c_double: "double", ctypes.c_size_t: "size_t"}[ctype]

    for name, (_, args) in pkg._lib.PROTOTYPES.items():
        m = re.search(r"\b%s\s*\(([^)]*)\)" % name, text)
        params = m.group(1).strip()
        plist = [] if params in ("", "void") else [x.strip() for x in params.split(",")]
        assert len(plist) == len(args), (name, plist, args)
        for decl, ctype in zip(plist, args):
            expected = "ptr" if "*" in decl else decl.replace("const ", "").split()[0]
            assert kind(ctype) == expected, (name, decl, ctype)


def test_no_cpu_fallback(pkg):
    """CPU tensors are rejected; nothing in the product package imports the oracle."""
    x = torch.zeros((1, 4, 4, 4), dtype=torch.float32)
    with pytest.raises(ValueError, match="CUDA"):
        pkg.histogram.calculate_rgbuv_histogram(x)
    with pytest.raises(ValueError, match="CUDA"):
        pkg.io_utils.extract_palette(torch.zeros((4, 4, 4), dtype=torch.int32), "grayness")
    out = subprocess.run(["grep", "-rIl", "--include=*.py", "--include=*.cu", "--include=*.cuh", "-E",
                          r"^\s*(from|import)\s+oracle|oracle/", os.path.join(ROOT, "palette_and_histo_gan_b200")],
                         capture_output=True, text=True)
    assert out.stdout.strip() == "", f"product package references the oracle: {out.stdout}"


def test_host_argument_validation(pkg):
    h = pkg.histogram
    with pytest.raises(ValueError):
        h._method_id("thresholding")
    with pytest.raises(ValueError):
        pkg.io_utils._ordering_id("by-count")
    assert h._sigma_sqr(0.02) == float(np.float32(4e-4))
    with pytest.raises(TypeError):
        pkg._tensor.from_any([1, 2, 3])
    with pytest.raises(ValueError):
        pkg.io_utils.extract_palette(torch.zeros((4, 4, 3), dtype=torch.int32), "grayness", channels=3)
    # host API: dtype / shape errors are raised before any device call
    img8 = np.zeros((2, 8, 8, 4), np.uint8)
    with pytest.raises(TypeError):
        pkg.hostapi.load_indexed_images(img8, img8.astype(np.int32))
    with pytest.raises(ValueError):
        pkg.hostapi.load_indexed_images(img8, img8, out=(np.zeros((2, 8, 8, 1), np.int32), np.zeros((2, 8, 8, 1), np.int32),
                                                         np.zeros((2, 255, 4), np.int32)))
    with pytest.raises(ValueError):
        pkg.hostapi.load_indexed_images(img8, img8, "by-count")
    # P6: the pixel helpers refuse CPU tensors (no eager / CPU fallback) and non-RGBA input to blacken
    for fn in (pkg.dataset_utils.blacken_transparent_pixels, pkg.dataset_utils.normalize, pkg.dataset_utils.denormalize):
        with pytest.raises(ValueError):
            fn(torch.zeros(4, 4, 4))
    assert pkg._lib.ORDERINGS["shuffled"] == 3 and pkg._lib.async_status() == 0
    # augmentation: CPU tensors and out-of-range hue shifts are refused
    d = pkg.dataset_utils
    with pytest.raises(ValueError):
        d.augment_two(torch.zeros(8, 8, 4), torch.zeros(8, 8, 4))
    assert d.HEIGHT_FACTOR == (-0.15, 0.075) and d.WIDTH_FACTOR == (-0.125, 0.125) and d.MAX_HUE_DELTA == 0.5
    t = d._draw_translations(64, 64, 64, torch.Generator().manual_seed(0))
    assert t.shape == (64, 2) and float(t[:, 0].abs().max()) <= 8.0 and float(t[:, 1].min()) >= -9.6001


def test_linspace_matches_oracle(pkg):
    from oracle.histogram_oracle import tf_linspace_f32

    for n in (1, 2, 16, 64, 256):
        assert np.array_equal(pkg.histogram.tf_linspace(-3.0, 3.0, n), tf_linspace_f32(-3.0, 3.0, n))


def test_dlpack_ingestion_is_zero_copy(pkg):
    class Foreign:  # any object speaking the DLPack protocol
        def __init__(self, t):
            self.t = t

        def __dlpack__(self, **kw):
            return self.t.__dlpack__(**kw)

        def __dlpack_device__(self):
            return self.t.__dlpack_device__()

    t = torch.arange(12, dtype=torch.float32)
    v = pkg._tensor.from_any(Foreign(t))
    assert v.data_ptr() == t.data_ptr()
    cap = torch.utils.dlpack.to_dlpack(t)
    assert pkg._tensor.from_any(cap).data_ptr() == t.data_ptr()


def test_compute_entry_point_fails_loudly_without_gpu(pkg):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    sm = ctypes.c_int()
    rc = pkg._lib.load().ph_device_info(0, ctypes.byref(sm), None, None)
    assert rc == pkg._lib.PH_ERR_CUDA
    assert "failed" in pkg._lib.last_error()


def test_256_bin_work_plans_cover_every_pixel_once(pkg):
    """Host logic of the dedicated 256-bin kernels (no device needed: the SM count falls back to 148): for a sweep of
    batch sizes and image sizes the forward's pixel slices and the backward's tile ranges cover every pixel exactly
    once, none is empty, slices are whole accumulation chains, the workspace holds every item's partial sums, and the
    cfgE shapes of the 1 / 2 / 4 / 8-GPU runs fill at least 95 % of their last wave."""
    lib = pkg._lib.load()
    out = (ctypes.c_int64 * 4)()
    sms = 148
    for batch in (1, 2, 5, 20, 37, 128, 147, 148, 149, 256, 512, 1024, 4096):
        for npix in (1, 31, 32, 100, 512, 513, 4096, 9216, 65536, 65537, 1 << 20):
            assert lib.ph_hist256_plan(batch, npix, out) == 0
            slices, pps, items, tpi = (int(v) for v in out)
            assert slices >= 1 and pps % 32 == 0
            assert slices * pps >= npix and (slices - 1) * pps < npix            # covered, last slice non-empty
            if slices > 1:
                assert pps % 512 == 0                                             # whole 512-pixel chains
            tiles = -(-npix // 128)
            assert items >= 1 and items * tpi >= tiles and (items - 1) * tpi < tiles
            assert items <= max(32, 1)
            need = batch * slices * 3 * 256 * 256 * 4
            assert lib.ph_hist_workspace_bytes(batch, npix, 256, pkg._lib.IMPLS["tc"]) >= need
    for per_gpu in (1024, 512, 256, 128):                                         # cfgE on 1, 2, 4, 8 GPUs
        assert lib.ph_hist256_plan(per_gpu, 65536, out) == 0
        n_items = per_gpu * int(out[0])
        assert n_items / (-(-n_items // sms) * sms) >= 0.95
        n_items = per_gpu * int(out[2])
        assert n_items / (-(-n_items // sms) * sms) >= 0.93
        assert int(out[2]) >= 5                                                   # G^ streams of <= ~32 images per wave
    assert lib.ph_hist256_plan(0, 64, out) != 0 and lib.ph_hist256_plan(4, 0, out) != 0


def test_tf_adapter_is_import_safe_without_tensorflow(pkg):
    """INTEGRATION.md §2 as code: importable without TensorFlow, a clear ImportError on first use when it is absent."""
    import importlib

    mod = importlib.import_module("palette_and_histo_gan_b200.tf_adapter")
    assert callable(mod.histogram_loss) and callable(mod.calculate_rgbuv_histogram)
    try:
        import tensorflow  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError, match="TensorFlow"):
            mod.histogram_loss(None, None)


def test_round2_entry_points_validate_before_touching_the_device(pkg):
    """Argument validation of the ABI v2 additions happens on the host (no GPU needed): NULL pointers, bad enums,
    misaligned buffers and unconnected communicators are PH_ERR_INVALID with a message, never a crash."""
    lib = pkg._lib.load()
    INVALID = pkg._lib.PH_ERR_INVALID
    buf = ctypes.create_string_buffer(4096)
    addr = ctypes.addressof(buf)
    aligned = (addr + 255) // 256 * 256
    # ph_pixel_map: NULL, unknown op, blacken on a non-RGBA element count
    assert lib.ph_pixel_map(None, 16, 1, aligned, None) == INVALID
    assert lib.ph_pixel_map(aligned, 16, 7, aligned, None) == INVALID and "bad op" in pkg._lib.last_error()
    assert lib.ph_pixel_map(aligned, 6, 0, aligned, None) == INVALID and "multiple of 4" in pkg._lib.last_error()
    # fused loader from uint8 pixels: NULL outputs, bad ordering, misaligned palette
    assert lib.ph_load_indexed_images_u8(aligned, aligned, 1, 16, 2, None, None, aligned, aligned, aligned, None) == INVALID
    assert lib.ph_load_indexed_images_u8(aligned, aligned, 1, 16, 9, None, aligned, aligned, aligned, aligned, None) == INVALID
    assert lib.ph_load_indexed_images_u8(aligned, aligned, 1, 16, 2, None, aligned, aligned, aligned + 4, aligned, None) == INVALID
    # 'shuffled' needs its keys (checked before any launch)
    assert lib.ph_extract_palette(aligned, 0, 16, 3, None, aligned, aligned, None) == INVALID and "shuffle keys" in pkg._lib.last_error()
    assert lib.ph_extract_palette(aligned, 0, 16, 2, None, aligned, aligned, None) == 0  # empty batch: nothing to do
    # communicator: bad rank / world, NULL handles; the sharded host call refuses a NULL communicator
    h = ctypes.c_void_p()
    assert lib.ph_comm_create(0, 3, 2, ctypes.byref(h)) == INVALID
    assert lib.ph_comm_create(0, 0, 99, ctypes.byref(h)) == INVALID
    assert lib.ph_comm_allreduce_sum_f64(None, aligned, 1, None) == INVALID
    assert lib.ph_comm_export(None, aligned) == INVALID and lib.ph_comm_connect(None, aligned) == INVALID
    assert lib.ph_host_hist_finish_comm(None, None, 4, aligned, None, None) == INVALID
    assert lib.ph_host_hist_loss_sharded(None, None, aligned, 0, aligned, 1, 16, 4, aligned, 64, 0, 4e-4, 1e-6, 0, 1, aligned, None,
                                         None) == INVALID
    # the sticky status word does not exist before a device was used: 0, and clearing is harmless
    assert pkg._lib.async_status(0, clear=True) == 0


def test_tf_adapter_is_import_safe_and_fences_both_hand_overs(pkg):
    """TensorFlow is absent here: the adapter must import, raise a clear ImportError on use, and its hand-over helpers
    must order the streams: a sync before borrowing a TensorFlow tensor, a stream synchronise before
    handing a result back."""
    import inspect

    from palette_and_histo_gan_b200 import _tensor, tf_adapter

    try:
        import tensorflow  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError, match="TensorFlow"):
            tf_adapter.histogram_loss(None, None)
    src = inspect.getsource(tf_adapter)
    assert "sync_devices" in src and "_sync_tf(tf, t)" in inspect.getsource(tf_adapter._to_torch)
    assert "synchronize()" in inspect.getsource(tf_adapter._to_tf)
    assert "_sync_tf" in inspect.getsource(_tensor.from_any) and "synchronize()" in inspect.getsource(_tensor.to_caller_framework)
