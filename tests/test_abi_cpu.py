"""CPU-side checks of the C ABI and host logic (no GPU needed)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle import witw_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "witw_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(witw_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from witw_b200 import _lib

    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "missing export " + n
    # the ctypes table covers exactly the header
    assert sorted(_lib.SIGNATURES) == names
    assert _lib.load().witw_version() >= 200


def test_host_grid_and_lut_match_oracle():
    from witw_b200 import _lib, ops

    x, y = ops.polar_grid()
    xr, yr = O.polar_grid()
    assert np.array_equal(x, xr) and np.array_equal(y, yr)
    n = x.size
    idx = np.empty((n, 4), np.int32)
    w = np.empty((n, 4), np.float32)
    _lib.call("witw_bilinear_lut", x.ctypes.data, y.ctypes.data, n, 256, 256, idx.ctypes.data, w.ctypes.data)
    (x0, x1, y0, y1), ws = O.bilinear_lut(xr, yr, 256, 256)
    for col, ref in enumerate((x0, x1, y0, y1)):
        assert np.array_equal(idx[:, col].reshape(x.shape), ref)
    for col, ref in enumerate(ws):
        assert np.array_equal(w[:, col].reshape(x.shape), ref)


def test_norm_table_matches_oracle():
    """witw_norm_lut (host): the 256 possible values of every channel after cvig_fov.py:147, bit for bit."""
    from witw_b200 import _lib

    mean = np.asarray(O.IMG_MEAN, np.float32)
    std = np.asarray(O.IMG_STD, np.float32)
    div = np.full(3, 255.0, np.float32)
    lut = np.empty((3, 256), np.float32)
    _lib.call("witw_norm_lut", div.ctypes.data, mean.ctypes.data, std.ctypes.data, 3, lut.ctypes.data)
    ramp = torch.arange(256, dtype=torch.uint8).view(1, 1, 256).expand(3, 1, 256)
    want = O.image_normalization(ramp).numpy().reshape(3, 256)
    assert np.array_equal(lut, want)


def test_sweep_schedules_stay_within_the_merge_width():
    """Candidate-list counts of both sweeps (host-side schedules, no GPU): 1..64 lists for any problem size."""
    from witw_b200 import _lib

    lib = _lib.load()
    for g, q in ((1, 1), (8, 128), (10000, 10000), (125000, 10000), (1000000, 1), (4096, 2048), (333, 100000)):
        for fn in (lib.witw_match_spec_topk_slots, lib.witw_match_tc_topk_slots):
            n = fn(g, q)
            assert 1 <= n <= 64, (g, q, n)
    assert lib.witw_match_spec_topk_slots(10000, 10000) == 52       # CTA pairs: 13 chunks x 4 lists, the best-balanced split
    try:                                                            # one CTA per tile: 28 chunks x 2 lists
        assert lib.witw_match_spec_variant(1) == 0
        assert lib.witw_match_spec_topk_slots(10000, 10000) == 56
        for g, q in ((1, 1), (1000000, 1), (333, 100000)):
            assert 1 <= lib.witw_match_spec_topk_slots(g, q) <= 64
        assert lib.witw_match_spec_variant(3) != 0                   # rejected, setting unchanged
    finally:
        assert lib.witw_match_spec_variant(2) == 0


def test_polar_plan_u8_structure():
    from witw_b200 import _lib

    lib = _lib.load()
    nbytes = lib.witw_polar_plan_bytes_u8(128, 512, 256)
    assert nbytes == lib.witw_polar_plan_bytes(128, 512, 256)      # same tables, different box geometry
    buf = np.zeros(nbytes, np.uint8)
    _lib.call("witw_polar_plan_build_u8", 128, 512, 256, buf.ctypes.data)
    hdr = buf[:80].view(np.int32)
    assert hdr[1:4].tolist() == [128, 512, 256]
    assert all(x % 16 == 0 for x in hdr[4:8].tolist())               # box starts on 16-byte boundaries of the uint8 rows


def test_polar_plan_structure():
    from witw_b200 import _lib

    lib = _lib.load()
    nbytes = lib.witw_polar_plan_bytes(128, 512, 256)
    assert nbytes > 3 * 4 * 65536 // 2
    buf = np.zeros(nbytes, np.uint8)
    _lib.call("witw_polar_plan_build", 128, 512, 256, buf.ctypes.data)
    hdr = buf[:64].view(np.int32)
    assert hdr[1:4].tolist() == [128, 512, 256]
    assert hdr[12] == 2  # the two clip-quirk pixels are the only exceptions
    # fractions reproduce the float64 grid: fx = fp32(x - floor(x))
    x, y = O.polar_grid()
    pw = int(hdr[13])
    assert pw in (8, 16, 32)
    lut_off = int(buf[56:60].view(np.uint32)[0])
    fx = buf[lut_off: lut_off + 4 * 65536].view(np.float32).reshape(4, 16, 1024)
    q, i, t = 2, 5, 777
    patch, lane = i * 32 + t // 32, t % 32      # warps tile the quadrant with pw x 32/pw patches
    npx = 128 // pw
    row, col = (patch // npx) * (32 // pw) + lane // pw, q * 128 + (patch % npx) * pw + lane % pw
    assert fx[q, i, t] == np.float32(x[row, col] - np.floor(x[row, col]))
    # unsupported geometry is refused, not approximated
    assert lib.witw_polar_plan_bytes(100, 300, 256) == 0
    assert "outside" in _lib.last_error()


def test_operand_size_queries():
    from witw_b200 import _lib

    lib = _lib.load()
    # 360 deg: 30 blocks of 128 B per (item pair, feature row) -> 120 KB per item
    assert lib.witw_gallery_operand_bytes(10000, 64, 64) == 5000 * 64 * 30 * 128
    assert lib.witw_gallery_operand_bytes(10000, 64, 16) == 5000 * 64 * 18 * 128
    assert lib.witw_query_operand_bytes(10000, 64, 64) == 10000 * 4096 * 2
    assert lib.witw_query_operand_bytes(10, 64, 12) == 10 * 64 * 16 * 2
    assert lib.witw_gallery_operand_bytes(8, 3, 16) == 0  # CH not a multiple of 64/sw_pad
    assert 1 <= lib.witw_match_tc_topk_slots(10000, 10000) <= 64


def test_no_cpu_fallback():
    import witw_b200 as W

    ov, su, _ = O.synth_features(4, 4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        W.match(ov, su)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        W.correlation(ov, su)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            W.PolarTransform()({"overhead": torch.zeros(3, 256, 256)})


def test_recall_from_ranks_matches_oracle():
    import witw_b200 as W

    ranks = np.array([1, 1, 2, 7, 11, 300, 5, 1, 3, 64])
    a, b = W.recall_from_ranks(torch.from_numpy(ranks)), O.recall_from_ranks(ranks)
    assert a.keys() == b.keys()
    for k in a:
        assert a[k] == b[k]


def test_install_rebinds_and_restores():
    import types

    import witw_b200 as W

    mod = types.ModuleType("fake_cvig")
    for n in ("bilinear_interpolate", "PolarTransform", "correlation", "crop_overhead", "l2_distance"):
        setattr(mod, n, object())
    orig = W.install(mod)
    assert mod.correlation is W.correlation and mod.PolarTransform is W.PolarTransform
    assert hasattr(mod, "evaluate_ranks") and set(orig) >= {"correlation", "l2_distance"}
    W.uninstall(mod)
    assert mod.correlation is orig["correlation"] and not hasattr(mod, "evaluate_ranks")
    assert not hasattr(mod, "triplet_loss") and not hasattr(mod, "Resize")
    # the loss and the dataset transforms are rebound only where the module has them / on request
    mod.triplet_loss, mod.Resize, mod.ImageNormalization = object(), object(), object()
    orig = W.install(mod, transforms=True)
    assert mod.triplet_loss is W.triplet_loss and mod.Resize is W.Resize and mod.ImageNormalization is W.ImageNormalization
    W.uninstall(mod)
    assert mod.triplet_loss is orig["triplet_loss"] and mod.Resize is orig["Resize"]
    W.install(mod)
    assert mod.Resize is orig["Resize"] and mod.triplet_loss is W.triplet_loss
    W.uninstall(mod)
    with pytest.raises(AttributeError):
        W.install(types.ModuleType("not_cvig"))


def test_shard_bounds_partition():
    import witw_b200 as W

    for n, p in ((10, 4), (1000000, 8), (3, 8), (0, 2)):
        spans = [W.shard_bounds(n, p, r) for r in range(p)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(p - 1))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def test_argument_validation_needs_no_device():
    """The C ABI's error behaviour (status code + witw_last_error message) for bad arguments, checked without a GPU: every
    call below must be refused before any CUDA work is attempted."""
    from witw_b200 import _lib, ops

    lib = _lib.load()
    INVALID, UNSUPPORTED = -1, -2
    # Resize front end
    assert lib.witw_resize_plan_bytes(0, 10, 4, 4, 1) == 0
    assert lib.witw_resize_plan_build(10, 10, 4, 4, 1, None) == INVALID and "null plan" in _lib.last_error()
    plan = ops.resize_plan_host(20, 30, 8, 12, True)
    one = np.ones(3, np.float32)
    assert lib.witw_resize_norm(None, 1, None, 1, 3, None, None, 0, 12, None, None, None, None) == INVALID
    assert lib.witw_resize_norm(None, 1, None, 1, 3, plan.ctypes.data, plan.ctypes.data, 12, 12, None, None, None, None) == INVALID
    assert "column window" in _lib.last_error()
    assert lib.witw_resize_norm(None, 1, None, 1, 9, plan.ctypes.data, plan.ctypes.data, 0, 12, one.ctypes.data, one.ctypes.data,
                                one.ctypes.data, None) == INVALID and "8 channels" in _lib.last_error()
    bad = plan.copy()
    bad[:4] = 0                                   # not a plan: wrong magic
    assert lib.witw_resize_norm(None, 1, None, 1, 3, bad.ctypes.data, bad.ctypes.data, 0, 12, None, None, None, None) == INVALID
    assert lib.witw_resize_norm(None, 1, None, 1, 3, plan.ctypes.data, plan.ctypes.data, 0, 12, None, None, None, None) == INVALID
    assert "null image pointer" in _lib.last_error()
    assert lib.witw_resize_norm(None, 1, None, 0, 3, plan.ctypes.data, plan.ctypes.data, 0, 12, None, None, None, None) == 0   # no planes: nothing to do
    # training slice
    assert lib.witw_triplet_loss_f32(None, 1, 10.0, None, None, None) == UNSUPPORTED and "batch size" in _lib.last_error()
    assert lib.witw_triplet_loss_f32(None, 8, 10.0, None, None, None) == INVALID
    assert lib.witw_match_backward_f32(None, None, None, None, None, None, None, 4, 4, 64, 64, 65, None) == INVALID
    assert lib.witw_match_backward_f32(None, None, None, None, None, None, None, 4, 4, 64, 64, 16, None) == INVALID
    assert "null pointer" in _lib.last_error()
    # peer-memory exchange: sizes and argument checks
    b2, b8 = lib.witw_peer_exchange_bytes(10000, 10, 2), lib.witw_peer_exchange_bytes(10000, 10, 8)
    assert 0 < b2 < b8 and b2 % 256 == 0 and b8 % 256 == 0
    assert b8 >= 2 * (10000 * 4 + 8 * 10000 * 4 + 2 * 8 * 10000 * 10 * 4)           # two copies of thresholds, counts, top-k of 8 ranks
    assert lib.witw_peer_exchange_bytes(10000, 10, 17) == 0 and lib.witw_peer_exchange_bytes(-1, 10, 2) == 0
    assert lib.witw_peer_thresholds(None, None, 0, 10, 100, 10, None, 2, 0, 1, None, None) == INVALID
    assert lib.witw_peer_results(None, None, None, None, 100, 10, None, None, 1, 0, 1, None, None, None, None, None) == INVALID   # world of one
    assert "world 1" in _lib.last_error()
    assert lib.witw_peer_open(None, None) == INVALID
    # matching kernels: shapes the kernels do not cover are refused with a reason
    assert lib.witw_spec_supported(64, 64, 64) == 1 and lib.witw_spec_supported(32, 64, 64) == 0 and lib.witw_spec_supported(64, 32, 16) == 0
    assert lib.witw_gallery_operand_bytes(10, 64, 0) == 0


@pytest.mark.parametrize("name", ["cvig_fov", "cvig_semantic"])
def test_install_on_the_real_reference_module(name):
    """install() against the unmodified reference module (build container only): every rebound callable accepts the
    reference's own positional arguments under the reference's own parameter names (cvig_fov.py:156, 186, 297, 318, 346, 366)."""
    import inspect

    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference tree not mounted")
    import witw_b200 as W

    mod = ref_loader.load(name)
    ref_sigs = {n: inspect.signature(getattr(mod, n)) for n in ("bilinear_interpolate", "correlation", "crop_overhead", "l2_distance", "triplet_loss")}
    ref_call = inspect.signature(mod.PolarTransform.__call__)
    ref_init = inspect.signature(mod.PolarTransform.__init__)
    originals = W.install(mod)
    try:
        assert mod.correlation is W.correlation and mod.PolarTransform is W.PolarTransform and mod.triplet_loss is W.triplet_loss
        for n, ref in ref_sigs.items():
            ours = inspect.signature(getattr(mod, n))
            ref_names = list(ref.parameters)
            assert list(ours.parameters)[: len(ref_names)] == ref_names, (n, ours, ref)
            for p in list(ours.parameters.values())[len(ref_names):]:      # anything extra must be optional
                assert p.default is not inspect.Parameter.empty, (n, p)
            for pn, p in ref.parameters.items():                           # the reference's defaults are kept
                if p.default is not inspect.Parameter.empty:
                    assert ours.parameters[pn].default == p.default, (n, pn)
        assert list(inspect.signature(mod.PolarTransform.__call__).parameters) == list(ref_call.parameters)
        mod.PolarTransform()                                               # the reference constructs it without arguments
        assert all(p.default is not inspect.Parameter.empty for k, p in inspect.signature(mod.PolarTransform.__init__).parameters.items()
                   if k not in ref_init.parameters)
        # heatmap.py and train()/test() also read these; install() must leave them alone
        assert mod.Globals is not None and hasattr(mod, "FOV_DSM") and mod.Resize is not W.Resize
        # the drop-ins refuse CPU tensors loudly (no CPU fallback), as the reference's CPU callers would find out at once
        import torch
        with pytest.raises(RuntimeError):
            mod.correlation(torch.zeros(2, 16, 4, 64), torch.zeros(1, 16, 4, 64))
    finally:
        W.uninstall(mod)
    assert mod.correlation is originals["correlation"] and mod.PolarTransform is originals["PolarTransform"]
    W.install(mod, polar=False)
    assert mod.PolarTransform is originals["PolarTransform"] and mod.correlation is W.correlation
    W.uninstall(mod)


def test_polar_drop_in_explains_itself_inside_a_dataloader_worker():
    """A CPU tensor inside a forked DataLoader worker: a clear error, not 'Cannot re-initialize CUDA in forked subprocess'."""
    import torch
    import torch.utils.data

    import witw_b200 as W

    class Tiles(torch.utils.data.Dataset):
        def __len__(self):
            return 2

        def __getitem__(self, i):
            try:
                W.PolarTransform()({"overhead": torch.zeros(3, 256, 256)})
            except RuntimeError as exc:
                return str(exc)
            return "no error"

    msgs = list(torch.utils.data.DataLoader(Tiles(), batch_size=1, num_workers=1))
    assert all("DataLoader worker" in m[0] and "num_workers=0" in m[0] for m in msgs), msgs
