"""Host-side mirror of the reference's hot-path functions, calling libwitw_b200.so.

Same names, argument meaning, tensor layouts and return types as model/cvig_fov.py
(= model/cvig_semantic.py) of IQTLabs/WITW:

    bilinear_interpolate   cvig_fov.py:156-183
    PolarTransform         cvig_fov.py:186-209
    correlation            cvig_fov.py:297-315
    crop_overhead          cvig_fov.py:318-343
    l2_distance            cvig_fov.py:346-363

plus the fused forms the reference spells as a Python loop:

    match                  correlation -> crop_overhead -> l2_distance   (cvig_fov.py:547-549)
    evaluate_ranks         the rank loop of test()                       (cvig_fov.py:543-552)
    recall_from_ranks      the thresholds that follow                    (cvig_fov.py:553-558)
    baseline_ranks         cvig_baseline.py:453-460
    heatmap_scores         tools/heatmap/heatmap.py:171-177

Everything runs on the CUDA device of its inputs; there is no CPU fallback.  crop_overhead and
l2_distance are differentiable (train() in the reference back-propagates through them,
cvig_fov.py:450-460); correlation is an argmax and carries no gradient, as in the reference;
the fused forms (match, evaluate_ranks, ...) are forward-only and refuse tensors that require
grad under an enabled grad mode rather than silently detaching them.
"""

import ctypes

import numpy as np
import torch

from . import _lib

# model/cvig_fov.py:19-22
SURFACE_HEIGHT_MAX = 128
SURFACE_WIDTH_MAX = 512
OVERHEAD_SIZE = 256

# problems with at least this many (gallery, query) pairs go to the tensor-core kernel
TC_MIN_PAIRS = 1 << 16
# standalone top-k over a materialised matrix: galleries of at least this many rows use threshold + filter + select
SELECT_MIN_ROWS = 8192


# Which tensor-core sweep serves a (gallery, query width): "hankel" = the shift search as one dense contraction
# (csrc/match_tc.cu, 8192*sw_pad FLOP per pair), "spectral" = per-frequency products + inverse FFT in the epilogue
# (csrc/match_spec.cu, 16.9 kFLOP per pair whatever the width; needs C*H == 64).  "auto": spectral when it is
# supported and the query has at least SPEC_MIN_SW columns (the spectral sweep costs the same at every width -- 6.8 ms
# for 10k x 10k against 36.3 ms (360 deg) / 9.6 ms (90 deg) of the dense contraction; narrower queries than that have
# flat spectra whose bf16 rounding costs more accuracy than the dense form's).
TC_IMPL = "auto"
SPEC_MIN_SW = 8


def _pick_impl(impl, ch, w, sw):
    impl = TC_IMPL if impl in (None, "auto") else impl
    if impl == "auto":
        ok = bool(_lib.load().witw_spec_supported(int(ch), int(w), int(sw)))
        return "spectral" if (ok and sw >= SPEC_MIN_SW) else "hankel"
    if impl not in ("hankel", "spectral"):
        raise ValueError("impl must be 'auto', 'hankel' or 'spectral'")
    if impl == "spectral" and not _lib.load().witw_spec_supported(int(ch), int(w), int(sw)):
        raise _lib.WitwError("spectral sweep does not cover CH=%d W=%d sw=%d" % (ch, w, sw))
    return impl


# Bounds on what the fp16 operands of the tensor-core sweeps can have done to a result (csrc/sweep_common.cuh): decisions
# within ERR_SIGMAS standard deviations of the accumulated rounding error are deferred to the fp32 finish (csrc/finish.cu).
ERR_SIGMAS = 5.0
# matrix outputs of the sweep: a pair whose error bound exceeds this fraction of its distance (the near matches) is
# overwritten with its fp32 distance, so every entry is within the north star's 1e-3 relative of the fp32 reference
FIX_REL = 1e-3
# test hook: a fixed capacity of the per-query deferral lists (None: sized from the gallery)
DEFERRAL_CAP = None
# at most this many queries: every pair is evaluated in fp32 from the spectra (heat map, the reference's one-query loop)
EXACT_SMALL_Q = 8
# the fused top-k keeps 16 candidates per query; the exact finish needs a few more candidates than results
TOPK_EXACT_MAX = 12


# ----------------------------------------------------------------------------- helpers
def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _need_cuda(name, *tensors, allow_grad=False):
    for t in tensors:
        if not isinstance(t, torch.Tensor):
            raise TypeError("%s: expected torch tensors, got %r" % (name, type(t)))
        if not t.is_cuda:
            raise RuntimeError(
                "%s: tensor is on %s; witw_b200 runs on a CUDA (sm_100a) device only and has no CPU fallback" % (name, t.device)
            )
        if t.requires_grad and torch.is_grad_enabled() and not allow_grad:
            raise RuntimeError(
                "%s: forward-only kernel got a tensor that requires grad; wrap the call in torch.no_grad() "
                "(training through crop_overhead/l2_distance is not covered by witw_b200)" % name
            )
    dev = tensors[0].device
    for t in tensors[1:]:
        if t.device != dev:
            raise RuntimeError("%s: tensors are on different devices (%s vs %s)" % (name, dev, t.device))
    return dev


def _f32c(t):
    """fp32, contiguous view/copy of a feature tensor (bf16/fp16 inputs are widened)."""
    if t.dtype not in (torch.float32, torch.bfloat16, torch.float16):
        raise TypeError("witw_b200: unsupported dtype %s (fp32, or bf16/fp16 widened to fp32)" % t.dtype)
    return t.detach().to(torch.float32).contiguous()


def _feature_dims(name, overhead_embed, surface_embed):
    if overhead_embed.dim() != 4 or surface_embed.dim() != 4:
        raise ValueError("%s: expected [G,C,H,W] and [Q,C,H,sw] feature maps" % name)
    g, c, h, w = overhead_embed.shape
    q, sc, sh, sw = surface_embed.shape
    if (c, h) != (sc, sh):
        # F.conv2d in the reference raises on a channel mismatch and yields an empty/garbage map on a height mismatch
        raise RuntimeError("%s: feature maps disagree in (C,H): %s vs %s" % (name, (c, h), (sc, sh)))
    if sw > w:
        raise RuntimeError("%s: query width %d exceeds gallery width %d" % (name, sw, w))
    return g, q, c * h, w, sw


# ----------------------------------------------------------------------------- K1
_lut_cache = {}
_plan_cache = {}


def _gather_lut(x, y, src_h, src_w, device):
    """Device copies of the reference's tap indices / weights for sample points (x, y)."""
    x = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
    y = np.ascontiguousarray(np.asarray(y, dtype=np.float64))
    n = x.size
    idx = np.empty((n, 4), dtype=np.int32)
    w = np.empty((n, 4), dtype=np.float32)
    _lib.call("witw_bilinear_lut", x.ctypes.data, y.ctypes.data, n, src_h, src_w, idx.ctypes.data, w.ctypes.data)
    return torch.from_numpy(idx).to(device), torch.from_numpy(w).to(device)


def bilinear_interpolate(im, x, y):
    """Drop-in for cvig_fov.py:156-183.  im [C,H,W] (or [N,C,H,W]) on CUDA, x/y float arrays [h,w].

    Returns [C,h,w] (or [N,C,h,w]) fp32, bit-identical to the reference: taps and weights come
    from the same float64 arithmetic, the blend keeps the reference's order with no FMA contraction.
    """
    x = np.asarray(x)
    y = np.asarray(y)
    assert x.shape == y.shape
    dev = _need_cuda("bilinear_interpolate", im)
    src = _f32c(im)
    lead = src.shape[:-2]
    src_h, src_w = src.shape[-2:]
    with torch.cuda.device(dev):
        idx, w = _gather_lut(x, y, src_h, src_w, dev)
        n_img = int(np.prod(lead)) if len(lead) else 1
        out = torch.empty(tuple(lead) + tuple(x.shape), dtype=torch.float32, device=dev)
        _lib.call("witw_bilinear_gather_f32", src.data_ptr(), out.data_ptr(), idx.data_ptr(), w.data_ptr(),
                  n_img, src_h, src_w, x.size, _stream())
    return out


def polar_grid(h_s=SURFACE_HEIGHT_MAX, w_s=SURFACE_WIDTH_MAX, s_o=OVERHEAD_SIZE):
    """Sample coordinates of cvig_fov.py:197-201 as float64 arrays (x, y) of shape [h_s, w_s]."""
    x = np.empty((h_s, w_s), dtype=np.float64)
    y = np.empty((h_s, w_s), dtype=np.float64)
    _lib.call("witw_polar_grid", h_s, w_s, s_o, x.ctypes.data, y.ctypes.data)
    return x, y


def _polar_plan(h_s, w_s, s_o, device):
    key = (h_s, w_s, s_o, str(device))
    if key not in _plan_cache:
        nbytes = _lib.load().witw_polar_plan_bytes(h_s, w_s, s_o)
        if nbytes == 0:
            _plan_cache[key] = None
        else:
            host = np.zeros(nbytes, dtype=np.uint8)
            _lib.call("witw_polar_plan_build", h_s, w_s, s_o, host.ctypes.data)
            _plan_cache[key] = (host, torch.from_numpy(host).to(device))
    return _plan_cache[key]


def polar_transform(overhead, h_s=SURFACE_HEIGHT_MAX, w_s=SURFACE_WIDTH_MAX, s_o=OVERHEAD_SIZE, exact=False):
    """Batched polar transform: overhead [...,s_o,s_o] on CUDA -> [...,h_s,w_s] fp32.

    exact=False: the staged throughput kernel (weights within 2^-24 of the reference's);
    exact=True, or a geometry outside the staged kernel: the bit-exact gather kernel.
    """
    dev = _need_cuda("polar_transform", overhead)
    src = _f32c(overhead)
    if src.shape[-1] != s_o or src.shape[-2] != s_o:
        raise ValueError("polar_transform: expected %dx%d tiles, got %s" % (s_o, s_o, tuple(src.shape[-2:])))
    lead = tuple(src.shape[:-2])
    n_img = int(np.prod(lead)) if len(lead) else 1
    out = torch.empty(lead + (h_s, w_s), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        plan = None if exact else _polar_plan(h_s, w_s, s_o, dev)
        if plan is None:
            key = ("grid", h_s, w_s, s_o, str(dev))
            if key not in _lut_cache:
                gx, gy = polar_grid(h_s, w_s, s_o)
                _lut_cache[key] = _gather_lut(gx, gy, s_o, s_o, dev)
            idx, w = _lut_cache[key]
            _lib.call("witw_bilinear_gather_f32", src.data_ptr(), out.data_ptr(), idx.data_ptr(), w.data_ptr(),
                      n_img, s_o, s_o, h_s * w_s, _stream())
        else:
            host, devplan = plan
            _lib.call("witw_polar_resample_f32", src.data_ptr(), out.data_ptr(), n_img, host.ctypes.data,
                      devplan.data_ptr(), _stream())
    return out


IMG_MEAN = (0.485, 0.456, 0.406)   # model/cvig_fov.py:24-25
IMG_STD = (0.229, 0.224, 0.225)
_norm_cache = {}


def _norm_table(mean, std, divisor, device):
    """(host, device) copies of the per-channel normalisation table [C,256] fp32 of cvig_fov.py:147."""
    mean = tuple(float(m) for m in mean)
    std = tuple(float(v) for v in std)
    divisor = tuple(float(d) for d in (divisor if hasattr(divisor, "__len__") else [divisor] * len(mean)))
    if not (len(mean) == len(std) == len(divisor)) or not mean:
        raise ValueError("normalized_polar: mean, std and divisor must have one entry per channel")
    key = (mean, std, divisor, str(device))
    if key not in _norm_cache:
        c = len(mean)
        host = np.empty((c, 256), dtype=np.float32)
        m, sd, dv = (np.asarray(v, dtype=np.float32) for v in (mean, std, divisor))
        _lib.call("witw_norm_lut", dv.ctypes.data, m.ctypes.data, sd.ctypes.data, c, host.ctypes.data)
        _norm_cache[key] = (host, torch.from_numpy(host).to(device))
    return _norm_cache[key]


def normalized_polar(overhead_u8, mean=IMG_MEAN, std=IMG_STD, divisor=255.0, exact=False,
                     h_s=SURFACE_HEIGHT_MAX, w_s=SURFACE_WIDTH_MAX, s_o=OVERHEAD_SIZE):
    """ImageNormalization + PolarTransform in one kernel (cvig_fov.py:137-149 then 186-209) for uint8 tiles
    [..., C, s_o, s_o] on CUDA that already have the model's size (Resize, cvig_fov.py:133, is the identity for those;
    other sizes are refused: torchvision's resize is not restated here).  Returns [..., C, h_s, w_s] fp32 =
    ``polar(norm(tile / 255.))``.  One byte per source pixel is read instead of four.

    exact=False: the staged kernel (within 4e-6 absolute of the reference chain); exact=True: every tap through the
    reference's fp32 normalisation table and the reference's blend order, bit-identical to the chain.
    """
    dev = _need_cuda("normalized_polar", overhead_u8)
    if overhead_u8.dtype != torch.uint8:
        raise TypeError("normalized_polar: expected uint8 tiles, got %s" % overhead_u8.dtype)
    if overhead_u8.dim() < 3 or overhead_u8.shape[-1] != s_o or overhead_u8.shape[-2] != s_o:
        raise ValueError("normalized_polar: expected [...,C,%d,%d] tiles (the Resize of the reference is not part of this kernel), got %s"
                         % (s_o, s_o, tuple(overhead_u8.shape)))
    src = overhead_u8.contiguous()
    c = src.shape[-3]
    if c != len(mean):
        raise ValueError("normalized_polar: %d channels but %d means" % (c, len(mean)))
    lead = tuple(src.shape[:-2])
    n_planes = int(np.prod(lead))
    out = torch.empty(lead + (h_s, w_s), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        lut_host, lut_dev = _norm_table(mean, std, divisor, dev)
        plan = None
        if not exact:
            key = ("u8", h_s, w_s, s_o, str(dev))
            if key not in _plan_cache:
                nbytes = _lib.load().witw_polar_plan_bytes_u8(h_s, w_s, s_o)
                if nbytes == 0:
                    _plan_cache[key] = None
                else:
                    host = np.zeros(nbytes, dtype=np.uint8)
                    _lib.call("witw_polar_plan_build_u8", h_s, w_s, s_o, host.ctypes.data)
                    _plan_cache[key] = (host, torch.from_numpy(host).to(dev))
            plan = _plan_cache[key]
        if plan is None:
            key = ("grid", h_s, w_s, s_o, str(dev))
            if key not in _lut_cache:
                gx, gy = polar_grid(h_s, w_s, s_o)
                _lut_cache[key] = _gather_lut(gx, gy, s_o, s_o, dev)
            idx, w = _lut_cache[key]
            _lib.call("witw_bilinear_gather_u8", src.data_ptr(), out.data_ptr(), idx.data_ptr(), w.data_ptr(), n_planes, s_o, s_o,
                      h_s * w_s, lut_dev.data_ptr(), c, _stream())
        else:
            host, devplan = plan
            _lib.call("witw_polar_resample_u8", src.data_ptr(), out.data_ptr(), n_planes, c, lut_host.ctypes.data, lut_dev.data_ptr(),
                      host.ctypes.data, devplan.data_ptr(), _stream())
    return out


# ----------------------------------------------------------------------------- Resize + ImageNormalization (8f item 4)
PANORAMA = {"cvusa": True, "witw": False}     # Globals.path_formats[...]['panorama'], model/cvig_fov.py:36-51
_resize_cache = {}


def resize_plan_host(in_h, in_w, out_h, out_w, antialias):
    """Host copy (uint8 array) of the separable tap tables of one resize geometry; layout: csrc/resize.cu."""
    nbytes = _lib.load().witw_resize_plan_bytes(int(in_h), int(in_w), int(out_h), int(out_w), int(bool(antialias)))
    if nbytes == 0:
        raise ValueError("resize: sizes must be positive, got %dx%d -> %dx%d" % (in_h, in_w, out_h, out_w))
    host = np.zeros(nbytes, dtype=np.uint8)
    _lib.call("witw_resize_plan_build", int(in_h), int(in_w), int(out_h), int(out_w), int(bool(antialias)), host.ctypes.data)
    return host


RESIZE_CACHE_ENTRIES = 256      # plans kept per process (datasets with many distinct raw image sizes); least recently used go first


def _resize_plan(in_h, in_w, out_h, out_w, antialias, device):
    key = (int(in_h), int(in_w), int(out_h), int(out_w), bool(antialias), str(device))
    plan = _resize_cache.pop(key, None)
    if plan is None:
        host = resize_plan_host(in_h, in_w, out_h, out_w, antialias)
        plan = (host, torch.from_numpy(host).to(device))
        while len(_resize_cache) >= RESIZE_CACHE_ENTRIES:
            _resize_cache.pop(next(iter(_resize_cache)))
    _resize_cache[key] = plan       # (re)inserted last: dicts keep insertion order, so the first key is the oldest use
    return plan


def resize_normalize(images, out_h, out_w, antialias=True, mean=None, std=None, divisor=255.0, col_start=0, col_count=None):
    """Bilinear resize (+ optional normalisation) of raw images [..., C, H, W] (uint8 or fp32, CUDA) in one kernel:
    ``torchvision.transforms.functional.resize(img, (out_h, out_w))`` as Resize calls it (cvig_fov.py:119, 131, 133), then
    -- when ``mean`` is given -- ImageNormalization, ``((x / divisor) - mean) / std`` per channel (cvig_fov.py:147;
    cvig_semantic.py:174-175 uses divisor (255, 255, 255, 1, 1)).  antialias=True is what the reference computes with today's
    torchvision (>= 0.17 antialiases tensors by default), antialias=False what its pinned torchvision 0.9.1 computed.
    col_start / col_count: the wrap-around column window of a panorama (cvig_fov.py:120-129).  Returns fp32
    [..., C, out_h, col_count]."""
    dev = _need_cuda("resize_normalize", images)
    if images.dtype not in (torch.uint8, torch.float32):
        raise TypeError("resize_normalize: expected uint8 or float32 images, got %s" % images.dtype)
    if images.dim() < 3:
        raise ValueError("resize_normalize: expected [..., C, H, W] images")
    src = images.detach().contiguous()
    if src.data_ptr() % 16:          # the kernel stages source rows with aligned 16-byte loads from the tensor's base
        src = src.clone()
    c, in_h, in_w = src.shape[-3:]
    lead = tuple(src.shape[:-2])
    n_planes = int(np.prod(lead))
    col_count = int(out_w if col_count is None else col_count)
    norm = [0, 0, 0]
    if mean is not None:
        mean = [float(m) for m in mean]
        std = [float(v) for v in std]
        divisor = [float(d) for d in (divisor if hasattr(divisor, "__len__") else [divisor] * len(mean))]
        if not (len(mean) == len(std) == len(divisor) == c):
            raise ValueError("resize_normalize: %d channels but %d means / %d stds / %d divisors" % (c, len(mean), len(std), len(divisor)))
        norm = [np.asarray(v, dtype=np.float32) for v in (divisor, mean, std)]
    out = torch.empty(lead + (int(out_h), col_count), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        host, devplan = _resize_plan(in_h, in_w, out_h, out_w, antialias, dev)
        _lib.call("witw_resize_norm", src.data_ptr(), int(src.dtype == torch.uint8), out.data_ptr(), n_planes, c, host.ctypes.data,
                  devplan.data_ptr(), int(col_start), col_count, *[v.ctypes.data if mean is not None else 0 for v in norm], _stream())
    return out


def _via_device(t, device, fn, *args, **kwargs):
    """fn(t, ...) on the GPU: a CUDA tensor is used where it is; a CPU tensor (the reference's transforms run on CPU samples)
    is copied to ``device`` (default: the current CUDA device), processed there and returned on the CPU.  No CPU arithmetic."""
    if t.is_cuda:
        return fn(t, *args, **kwargs)
    _refuse_dataloader_worker(fn.__name__)
    if not torch.cuda.is_available():
        raise RuntimeError("%s: no CUDA device; witw_b200 has no CPU fallback" % fn.__name__)
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    return fn(t.to(dev, non_blocking=True), *args, **kwargs).to(t.device)


class Resize(object):
    """Drop-in for cvig_fov.py:100-134 on CUDA tensors ([C,H,W] samples or [N,C,H,W] batches, uint8 or fp32):
    the surface image to 128 x 512 and a window of ``int(fov/360*512)`` columns from a random start with wrap-around
    (panoramas, 'cvusa'), or straight to 128 x that width ('witw'); the overhead image to 256 x 256.  The start column is
    drawn exactly as the reference draws it (``torch.randint(0, 512, ())`` on the CPU generator), once per call: the images
    of a batch share their start column; call it per sample, or use ``resize_normalize(col_start=...)``, for one draw each."""

    def __init__(self, dataset, fov=360, random_orientation=True, antialias=True, device=None):
        self.fov = fov
        self.surface_width = int(self.fov / 360 * SURFACE_WIDTH_MAX)
        self.panorama = PANORAMA[dataset]
        self.random_orientation = random_orientation
        self.antialias = antialias
        self.device = device

    def __call__(self, data):
        if self.panorama:
            start = int(torch.randint(0, SURFACE_WIDTH_MAX, ())) if self.random_orientation else 0
            data["surface"] = _via_device(data["surface"], self.device, resize_normalize, SURFACE_HEIGHT_MAX, SURFACE_WIDTH_MAX,
                                          self.antialias, col_start=start, col_count=self.surface_width)
        else:
            data["surface"] = _via_device(data["surface"], self.device, resize_normalize, SURFACE_HEIGHT_MAX, self.surface_width,
                                          self.antialias)
        data["overhead"] = _via_device(data["overhead"], self.device, resize_normalize, OVERHEAD_SIZE, OVERHEAD_SIZE, self.antialias)
        return data


class ImageNormalization(object):
    """Drop-in for cvig_fov.py:137-149 (divisor 255) / cvig_semantic.py:163-176 (pass divisor=(255, 255, 255, 1, 1) and the
    five-channel mean / std) on CUDA tensors: ``data[key] = ((data[key] / divisor) - mean) / std`` for both images."""

    def __init__(self, mean=IMG_MEAN, std=IMG_STD, divisor=255.0, device=None):
        self.keys = ["surface", "overhead"]
        self.mean, self.std, self.divisor, self.device = mean, std, divisor, device

    def __call__(self, data):
        for key in self.keys:
            h, w = data[key].shape[-2:]
            data[key] = _via_device(data[key], self.device, resize_normalize, h, w, antialias=False, mean=self.mean, std=self.std,
                                    divisor=self.divisor)
        return data


def prepare_pair(surface, overhead, fov=360, panorama=True, start=0, antialias=True, mean=IMG_MEAN, std=IMG_STD, divisor=255.0):
    """The whole transform chain of the reference's datasets (cvig_fov.py:389-392: Resize -> ImageNormalization ->
    PolarTransform) on raw CUDA images, three kernels: resize + normalise the surface image(s), resize + normalise the
    overhead image(s), polar-transform the latter.  Returns {'surface', 'overhead', 'polar'} like the reference's sample
    dict; ``start`` is the panorama's first column (the reference draws it at random, cvig_fov.py:121-124)."""
    sw = int(fov / 360 * SURFACE_WIDTH_MAX)
    if panorama:
        su = resize_normalize(surface, SURFACE_HEIGHT_MAX, SURFACE_WIDTH_MAX, antialias, mean, std, divisor, col_start=start, col_count=sw)
    else:
        su = resize_normalize(surface, SURFACE_HEIGHT_MAX, sw, antialias, mean, std, divisor)
    ov = resize_normalize(overhead, OVERHEAD_SIZE, OVERHEAD_SIZE, antialias, mean, std, divisor)
    return {"surface": su, "overhead": ov, "polar": polar_transform(ov)}


class PolarTransform(object):
    """Drop-in for cvig_fov.py:186-209: ``data['polar'] = polar(data['overhead'])``, other keys kept.

    The reference runs this per sample on CPU inside DataLoader workers.  Here the work is done
    on the GPU: a CUDA tensor (single tile [C,256,256] or a batch [N,C,256,256]) is transformed in
    place on its device; a CPU tensor is copied to ``device`` (default: current CUDA device),
    transformed there and returned on the CPU so the reference's call sites keep working
    (this needs CUDA in the calling process, i.e. num_workers=0 or a spawn context).
    """

    def __init__(self, device=None, exact=False, h_s=SURFACE_HEIGHT_MAX, w_s=SURFACE_WIDTH_MAX, s_o=OVERHEAD_SIZE):
        self.device = device
        self.exact = exact
        self.geom = (h_s, w_s, s_o)

    def __call__(self, data):
        tile = data["overhead"]
        if tile.is_cuda:
            data["polar"] = polar_transform(tile, *self.geom, exact=self.exact)
        else:
            _refuse_dataloader_worker("PolarTransform")
            if not torch.cuda.is_available():
                raise RuntimeError("PolarTransform: no CUDA device; witw_b200 has no CPU fallback")
            dev = torch.device(self.device) if self.device is not None else torch.device("cuda", torch.cuda.current_device())
            data["polar"] = polar_transform(tile.to(dev, non_blocking=True), *self.geom, exact=self.exact).to(tile.device)
        return data


def _refuse_dataloader_worker(name):
    """The reference builds its transforms into the dataset and runs them inside forked DataLoader workers (num_workers=8 /
    12, cvig_fov.py:385, 490, 506), where CUDA cannot be initialised.  Say so, instead of torch's 'Cannot re-initialize CUDA in
    forked subprocess'."""
    import torch.utils.data

    info = torch.utils.data.get_worker_info()
    if info is not None:
        raise RuntimeError(
            "%s: called on a CPU tensor inside DataLoader worker %d -- the witw_b200 drop-in runs on the GPU, and CUDA cannot be "
            "used in the reference's forked worker processes.  Run the loader with num_workers=0 (or multiprocessing_context='spawn'), "
            "or keep the reference's transform in the workers (witw_b200.install(module, polar=False)) and call "
            "witw_b200.polar_transform on the batched data['overhead'] after the loader." % (name, info.id))


# ----------------------------------------------------------------------------- K2/K3
def spectral_rows(rows, row_len):
    """Packed 64-point azimuth spectra [n_rows,64] fp32 of contiguous fp32 feature rows (row_len <= 64 columns,
    zero-padded): the operand of the fp32 finish (csrc/spectral.cu)."""
    n_rows = rows.numel() // row_len
    out = torch.empty((n_rows, 64), dtype=torch.float32, device=rows.device)
    with torch.cuda.device(rows.device):
        _lib.call("witw_spectral_rows_f32", rows.data_ptr(), n_rows, int(row_len), out.data_ptr(), _stream())
    return out


def _gallery_tables(n_items, device, zero=False):
    """(gal_scale [n,64], gal_aux [n,4], crop_inv_norm [n,64]) of include/witw_b200.h for n (padded) items."""
    make = torch.zeros if zero else torch.empty
    return (make(n_items * 64, dtype=torch.float32, device=device), make(n_items * 4, dtype=torch.float32, device=device),
            make(n_items * 64, dtype=torch.float32, device=device))


class GalleryIndex(object):
    """Gallery feature maps prepared for the tensor-core sweeps and their fp32 finish: the fp16 operand of the norm-scaled
    features (azimuth spectra, or Hankel blocks for the dense sweep), the per-item scale / error-bound tables, and -- with
    keep_fp32 -- the fp32 azimuth spectra every exact evaluation works on (csrc/sweep_common.cuh, csrc/finish.cu).

    Built once per (gallery, query width); reused for every query batch.  ``g_offset`` is the
    global index of the first item when the gallery is one shard of a larger one.
    """

    def __init__(self, overhead_embed, surface_width, g_offset=0, keep_fp32=True, impl=None):
        dev = _need_cuda("GalleryIndex", overhead_embed)
        if overhead_embed.dim() != 4:
            raise ValueError("GalleryIndex: expected [G,C,H,W]")
        g, c, h, w = overhead_embed.shape
        self.device, self.G, self.CH, self.W, self.sw = dev, g, c * h, w, int(surface_width)
        self.C, self.H = c, h
        self.g_offset = int(g_offset)
        if w != 64:
            raise _lib.WitwError("GalleryIndex: tensor-core path needs W == 64, got %d" % w)
        ov = _f32c(overhead_embed)
        self.spec = None
        self.impl = _pick_impl(impl, self.CH, w, self.sw)
        with torch.cuda.device(dev):
            if keep_fp32:
                self.spec = torch.empty((g * self.CH, 64), dtype=torch.float32, device=dev)
            if self.impl == "spectral":
                self.operand = torch.empty(_lib.load().witw_spec_gallery_operand_bytes(g, self.CH), dtype=torch.uint8, device=dev)
                self.scale, self.aux, self.crop_inv_norm = _gallery_tables(max((g + 7) // 8 * 8, 8), dev)
                # the finish's fp32 spectra come out of the same pass over the features
                _lib.call("witw_spec_gallery_prep", ov.data_ptr(), g, 0, self.CH, w, self.sw, self.operand.data_ptr(), self.scale.data_ptr(),
                          self.aux.data_ptr(), self.crop_inv_norm.data_ptr(), _ptr(self.spec), _stream())
                return
            nbytes = _lib.load().witw_gallery_operand_bytes(g, self.CH, self.sw)
            if nbytes == 0:
                raise _lib.WitwError("GalleryIndex: " + _lib.last_error())
            self.operand = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            self.scale, self.aux, self.crop_inv_norm = _gallery_tables(max((g + 3) // 4 * 4, 4), dev)
            _lib.call("witw_gallery_prep", ov.data_ptr(), g, self.CH, w, self.sw, self.operand.data_ptr(), self.scale.data_ptr(),
                      self.aux.data_ptr(), self.crop_inv_norm.data_ptr(), _stream())
            if keep_fp32 and g:
                _lib.call("witw_spectral_rows_f32", ov.data_ptr(), g * self.CH, w, self.spec.data_ptr(), _stream())

    def spectral(self):
        """Packed azimuth spectra [G*CH,64] of the fp32 features."""
        if self.spec is None:
            raise ValueError("GalleryIndex: the fp32 spectra were dropped (keep_fp32=False); no fp32 finish")
        return self.spec


class GalleryBuilder(object):
    """Incremental GalleryIndex for the encode loop of test() (cvig_fov.py:519-532).

    The reference grows ``overhead_embed`` with torch.cat per batch (O(n^2) copies).  Here each encoder
    output batch is written once, straight into its slot of the preallocated tensor-core operand, its tables and
    (keep_fp32) the fp32 spectra the exact finish needs (16 KB per item).
    Batches must hold a multiple of 4 items (8 for the spectral sweep: ``batch_multiple``), except the last one.
    keep_raw: also keep the raw fp32 features.
    """

    def __init__(self, capacity, surface_width, channels=16, height=4, width=64, device=None, g_offset=0, keep_fp32=True,
                 keep_raw=False, impl=None):
        if not torch.cuda.is_available():
            raise RuntimeError("GalleryBuilder: no CUDA device; witw_b200 has no CPU fallback")
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.capacity, self.sw = int(capacity), int(surface_width)
        self.C, self.H, self.W, self.CH = channels, height, width, channels * height
        self.g_offset, self.count, self.closed = int(g_offset), 0, False
        if width != 64:
            raise _lib.WitwError("GalleryBuilder: tensor-core path needs W == 64")
        lib = _lib.load()
        self.impl = _pick_impl(impl, self.CH, width, self.sw)
        self.batch_multiple = 8 if self.impl == "spectral" else 4
        with torch.cuda.device(self.device):
            if self.impl == "spectral":
                nbytes = lib.witw_spec_gallery_operand_bytes(self.capacity, self.CH)
                self.pair_bytes = 0
            else:
                nbytes = lib.witw_gallery_operand_bytes(self.capacity, self.CH, self.sw)
                if nbytes == 0:
                    raise _lib.WitwError("GalleryBuilder: " + _lib.last_error())
                self.pair_bytes = lib.witw_gallery_operand_bytes(4, self.CH, self.sw) // 2
            self.operand = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
            self.scale, self.aux, self.crop_inv_norm = _gallery_tables(max((self.capacity + 7) // 8 * 8, 8), self.device, zero=True)
            self.ov = torch.empty((self.capacity, channels, height, width), dtype=torch.float32, device=self.device) if keep_raw else None
            self.spec = torch.empty((self.capacity * self.CH, 64), dtype=torch.float32, device=self.device) if keep_fp32 else None

    def append(self, overhead_embed_part):
        """Add one encoder output batch [n,C,H,W] (CUDA)."""
        _need_cuda("GalleryBuilder.append", overhead_embed_part)
        if self.closed:
            raise RuntimeError("GalleryBuilder.append: a batch that is not a multiple of %d items must be the last one" % self.batch_multiple)
        n = overhead_embed_part.shape[0]
        if tuple(overhead_embed_part.shape[1:]) != (self.C, self.H, self.W):
            raise ValueError("GalleryBuilder.append: expected [n,%d,%d,%d]" % (self.C, self.H, self.W))
        if self.count + n > self.capacity:
            raise ValueError("GalleryBuilder.append: capacity %d exceeded" % self.capacity)
        if n == 0:
            return self
        part = _f32c(overhead_embed_part)
        with torch.cuda.device(self.device):
            if self.ov is not None:
                self.ov[self.count: self.count + n].copy_(part)
            spec_ptr = 0 if self.spec is None else self.spec.data_ptr() + self.count * self.CH * 64 * 4
            tables = (self.scale.data_ptr() + self.count * 64 * 4, self.aux.data_ptr() + self.count * 4 * 4,
                      self.crop_inv_norm.data_ptr() + self.count * 64 * 4)
            if self.impl == "spectral":
                _lib.call("witw_spec_gallery_prep", part.data_ptr(), n, self.count, self.CH, self.W, self.sw, self.operand.data_ptr(),
                          tables[0], tables[1], tables[2], spec_ptr, _stream())
            else:
                if spec_ptr:
                    _lib.call("witw_spectral_rows_f32", part.data_ptr(), n * self.CH, self.W, spec_ptr, _stream())
                _lib.call("witw_gallery_prep", part.data_ptr(), n, self.CH, self.W, self.sw,
                          self.operand.data_ptr() + (self.count // 2) * self.pair_bytes, tables[0], tables[1], tables[2], _stream())
        self.count += n
        self.closed = n % self.batch_multiple != 0
        return self

    def finish(self):
        """The GalleryIndex over everything appended so far."""
        idx = GalleryIndex.__new__(GalleryIndex)
        idx.device, idx.G, idx.CH, idx.W, idx.sw = self.device, self.count, self.CH, self.W, self.sw
        idx.C, idx.H, idx.g_offset, idx.impl = self.C, self.H, self.g_offset, self.impl
        idx.spec = None if self.spec is None else self.spec[: self.count * self.CH]
        idx.operand, idx.scale, idx.aux, idx.crop_inv_norm = self.operand, self.scale, self.aux, self.crop_inv_norm
        return idx


class QueryBatch(object):
    """Query feature maps prepared for the tensor-core sweeps (fp16 operand of the norm-scaled features, per-query sweep
    constants, inverse norms) and, with keep_fp32, their fp32 azimuth spectra for the exact finish."""

    def __init__(self, surface_embed, keep_fp32=True, impl=None):
        dev = _need_cuda("QueryBatch", surface_embed)
        q, c, h, sw = surface_embed.shape
        self.device, self.Q, self.CH, self.sw = dev, q, c * h, sw
        su = _f32c(surface_embed)
        self.spec = None
        self.impl = _pick_impl(impl, self.CH, 64, sw)
        with torch.cuda.device(dev):
            self.inv_norm = torch.empty(max(q, 1), dtype=torch.float32, device=dev)
            self.aux = torch.empty(max(q, 1) * 2, dtype=torch.float32, device=dev)
            if keep_fp32:
                self.spec = torch.empty((q * self.CH, 64), dtype=torch.float32, device=dev)
            if self.impl == "spectral":
                self.operand = torch.empty(_lib.load().witw_spec_query_operand_bytes(q, self.CH), dtype=torch.uint8, device=dev)
                _lib.call("witw_spec_query_prep", su.data_ptr(), q, self.CH, sw, self.operand.data_ptr(), self.aux.data_ptr(),
                          self.inv_norm.data_ptr(), _ptr(self.spec), _stream())
                return
            nbytes = _lib.load().witw_query_operand_bytes(q, self.CH, sw)
            if nbytes == 0:
                raise _lib.WitwError("QueryBatch: " + _lib.last_error())
            self.operand = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            _lib.call("witw_query_prep", su.data_ptr(), q, self.CH, sw, self.operand.data_ptr(), self.aux.data_ptr(),
                      self.inv_norm.data_ptr(), _stream())
            if keep_fp32 and q:
                _lib.call("witw_spectral_rows_f32", su.data_ptr(), q * self.CH, sw, self.spec.data_ptr(), _stream())

    def spectral(self):
        """Packed azimuth spectra [Q*CH,64] of the zero-padded fp32 query rows."""
        if self.spec is None:
            raise ValueError("QueryBatch: the fp32 spectra were dropped (keep_fp32=False); no fp32 finish")
        return self.spec


def tc_supported(ch, w, sw):
    """Shapes the tcgen05 kernel covers: 64 azimuth columns, K = CH*sw_pad a multiple of 64."""
    if w != 64 or sw < 1 or sw > 64:
        return False
    sw_pad = 16 if sw <= 16 else (32 if sw <= 32 else 64)
    return ch % (64 // sw_pad) == 0


def _pick_path(path, g, q, ch, w, sw):
    if path == "auto":
        return "tc" if (tc_supported(ch, w, sw) and g * q >= TC_MIN_PAIRS) else "fp32"
    if path not in ("tc", "fp32"):
        raise ValueError("path must be 'auto', 'tc' or 'fp32'")
    if path == "tc" and not tc_supported(ch, w, sw):
        raise _lib.WitwError("tensor-core path does not cover CH=%d W=%d sw=%d" % (ch, w, sw))
    return path


class Deferral(object):
    """Per-query lists of the (gallery, query) pairs a sweep defers to the fp32 finish (csrc/sweep_common.cuh): list_n [Q]
    counts them, list_g [Q,cap] holds them; a count above cap marks the query for a full fp32 re-evaluation.  The int32
    bookkeeping (rank counts, list_n, qflag, n_flagged) lives in one zero-filled buffer: one fill per evaluation."""

    def __init__(self, n_queries, n_gallery, device, matrix=False, err_sigmas=None, fix_rel=None, cap=None):
        cap = DEFERRAL_CAP if cap is None else cap
        if cap is None:   # room for 3 % (12 % for matrix outputs, where every uncertain argmax is listed) of the gallery per query
            cap = min(max(n_gallery // (8 if matrix else 32), 64), 8192)
        self.cap = int(cap)
        self.err_sigmas = float(ERR_SIGMAS if err_sigmas is None else err_sigmas)
        self.fix_rel = float(FIX_REL if fix_rel is None else fix_rel) if matrix else 0.0
        qp = (max(n_queries, 1) + 3) // 4 * 4
        self.list_g = torch.empty(qp * self.cap, dtype=torch.int32, device=device)
        book = torch.zeros(3 * qp + 4, dtype=torch.int32, device=device)
        self.counts, self.list_n, self.qflag, self.n_flagged = book[:qp], book[qp: 2 * qp], book[2 * qp: 3 * qp], book[3 * qp: 3 * qp + 1]


# Opt-in: an L2 access-policy window over the query operand while a sweep runs (the operand is re-read once per 8 gallery
# items; uploads of the next batch and the gallery operand stream through the same cache).  Raises the device's persisting-L2
# set-aside, a device-wide setting, hence off unless asked for.
L2_WINDOW = False


def sweep_tc(gallery, queries, want_dist=False, want_ori=False, d_true=None, true_idx=None, rank_count=None, topk=0, events=None,
             deferral=None):
    """One tensor-core sweep of a QueryBatch over a GalleryIndex.  Returns a dict with the requested outputs.

    events: optional (start, end) torch.cuda.Event pair recorded around the sweep kernel alone (bench.py's roofline timer).
    deferral: optional Deferral; what fp16 cannot settle is listed there for finish_tc() instead of being decided here, and the
    top-k keys are lower bounds of the exact distances.  Without it: the plain fp16 results."""
    if gallery.CH != queries.CH or gallery.sw != queries.sw or gallery.device != queries.device:
        raise ValueError("sweep_tc: gallery and queries disagree (CH %d/%d, sw %d/%d)" % (gallery.CH, queries.CH, gallery.sw, queries.sw))
    if gallery.impl != queries.impl:
        raise ValueError("sweep_tc: gallery operand is %r but the query operand is %r" % (gallery.impl, queries.impl))
    fn_sweep, fn_slots = ("witw_match_spec", "witw_match_spec_topk_slots") if gallery.impl == "spectral" else ("witw_match_tc", "witw_match_tc_topk_slots")
    dev, g, q = gallery.device, gallery.G, queries.Q
    out = {}
    with torch.cuda.device(dev):
        dist = torch.empty((g, q), dtype=torch.float32, device=dev) if want_dist else None
        ori = torch.empty((g, q), dtype=torch.uint8, device=dev) if want_ori else None
        tk_d = tk_i = None
        slots = 0
        if topk:
            slots = getattr(_lib.load(), fn_slots)(g, q)
            tk_d = torch.empty((slots, q, topk), dtype=torch.float32, device=dev)
            tk_i = torch.empty((slots, q, topk), dtype=torch.int32, device=dev)
        if g > 0 and q > 0:
            args = _lib.SweepArgs(
                gal_op=gallery.operand.data_ptr(), gal_scale=gallery.scale.data_ptr(), gal_aux=gallery.aux.data_ptr(),
                qry_op=queries.operand.data_ptr(), qry_aux=queries.aux.data_ptr(), G=g, Q=q, CH=gallery.CH, sw=gallery.sw,
                g_index_offset=gallery.g_offset, topk=int(topk), dist=_ptr(dist), ori=_ptr(ori), d_true=_ptr(d_true), true_idx=_ptr(true_idx),
                rank_count=_ptr(rank_count), topk_key=_ptr(tk_d), topk_idx=_ptr(tk_i),
                list_g=_ptr(deferral.list_g) if deferral else 0, list_n=_ptr(deferral.list_n) if deferral else 0,
                list_cap=deferral.cap if deferral else 0, err_sigmas=deferral.err_sigmas if deferral else 0.0,
                fix_rel=deferral.fix_rel if deferral else 0.0)
            if L2_WINDOW:
                _lib.call("witw_stream_l2_window", queries.operand.data_ptr(), queries.operand.numel(), _stream())
            if events is not None:
                events[0].record()
            _lib.call(fn_sweep, ctypes.addressof(args), _stream())
            if events is not None:
                events[1].record()
            if L2_WINDOW:
                _lib.call("witw_stream_l2_window", 0, 0, _stream())
        if topk:
            fin_d = torch.empty((q, topk), dtype=torch.float32, device=dev)
            fin_i = torch.empty((q, topk), dtype=torch.int32, device=dev)
            if q > 0:
                _lib.call("witw_topk_merge", tk_d.data_ptr(), tk_i.data_ptr(), slots, q, int(topk), fin_d.data_ptr(), fin_i.data_ptr(), _stream())
            out["topk_dist"], out["topk_idx"] = fin_d, fin_i
    out["dist"], out["ori"] = dist, ori
    return out


def finish_tc(gallery, queries, deferral, d_true=None, rank_count=None, dist=None, ori=None, cand_key=None, cand_idx=None, k_out=0):
    """The fp32 finish of a sweep that ran with ``deferral`` (csrc/finish.cu): settles the deferred pairs (rank decisions
    into rank_count, fp32 values into dist / ori) and re-ranks the merged top-k candidates.  Returns (topk_dist, topk_idx)
    or None.  Queries it cannot finish are flagged in deferral.qflag / n_flagged (see redo_flagged)."""
    dev, q = gallery.device, queries.Q
    td = ti = None
    with torch.cuda.device(dev):
        kc = 0
        if k_out:
            kc = cand_idx.shape[1]
            if gallery.G > 0:
                td = torch.empty((q, k_out), dtype=torch.float32, device=dev)
                ti = torch.empty((q, k_out), dtype=torch.int32, device=dev)
            else:               # an empty gallery: no candidates
                td = torch.full((q, k_out), float("inf"), dtype=torch.float32, device=dev)
                ti = torch.full((q, k_out), -1, dtype=torch.int32, device=dev)
        if gallery.G > 0 and q > 0:
            scratch = torch.empty(_lib.load().witw_finish_scratch_bytes(q, kc), dtype=torch.uint8, device=dev)
            args = _lib.FinishArgs(
                gal_spec=gallery.spectral().data_ptr(), crop_inv_norm=gallery.crop_inv_norm.data_ptr(), qry_spec=queries.spectral().data_ptr(),
                q_inv_norm=queries.inv_norm.data_ptr(), G=gallery.G, Q=q, CH=gallery.CH, g_index_offset=gallery.g_offset,
                list_g=deferral.list_g.data_ptr(), list_n=deferral.list_n.data_ptr(), list_cap=deferral.cap, kc=kc, d_true=_ptr(d_true),
                rank_count=_ptr(rank_count), dist=_ptr(dist), ori=_ptr(ori), cand_key=_ptr(cand_key.contiguous() if k_out else None),
                cand_idx=_ptr(cand_idx.contiguous() if k_out else None), out_dist=_ptr(td), out_idx=_ptr(ti), k_out=int(k_out), reserved=0,
                qflag=deferral.qflag.data_ptr(), n_flagged=deferral.n_flagged.data_ptr(), scratch=scratch.data_ptr())
            _lib.call("witw_finish_spec_f32", ctypes.addressof(args), _stream())
    return (td, ti) if k_out else None


def exact_columns(gallery, queries, q_sel=None, dist=None, ori64=None, ori8=None, ld=None, col_is_q=False, d_true=None, true_idx=None,
                  count_out=None):
    """Exact fp32 distances / orientations of whole query columns from the fp32 spectra (witw_match_columns_spec_f32):
    queries q_sel (int32 tensor, default all) against every item of the gallery."""
    f = queries.Q if q_sel is None else int(q_sel.numel())
    if f == 0 or gallery.G == 0:
        return
    with torch.cuda.device(gallery.device):
        for lo in range(0, f, 65535):
            n = min(65535, f - lo)
            off = 0 if col_is_q else lo
            sel_ptr = 0 if q_sel is None else q_sel.data_ptr() + 4 * lo
            if q_sel is None and lo:
                raise ValueError("exact_columns: more than 65535 columns need an explicit q_sel")
            _lib.call("witw_match_columns_spec_f32", gallery.spectral().data_ptr(), gallery.crop_inv_norm.data_ptr(), queries.spectral().data_ptr(),
                      queries.inv_norm.data_ptr(), gallery.G, gallery.CH, sel_ptr, n,
                      0 if dist is None else dist.data_ptr() + 4 * off, 0 if ori64 is None else ori64.data_ptr() + 8 * off,
                      0 if ori8 is None else ori8.data_ptr() + off, int(f if ld is None else ld), int(bool(col_is_q)), _ptr(d_true), _ptr(true_idx),
                      gallery.g_offset, 0 if count_out is None else count_out.data_ptr() + 4 * lo, _stream())


_index_cache = []          # [(key, features, GalleryIndex)]: the gallery of the reference's one-query loop is prepared once, not per query
INDEX_CACHE_ENTRIES = 2


def _cached_index(overhead_embed, sw):
    """The prepared gallery of a feature tensor, built on first use.  The key is the tensor's storage address, shape, dtype and
    version counter (in-place writes bump it); the entry keeps a reference to the tensor, so its storage cannot be freed and
    the address handed to another tensor while the entry lives."""
    key = (overhead_embed.data_ptr(), tuple(overhead_embed.shape), tuple(overhead_embed.stride()), overhead_embed.dtype,
           overhead_embed._version, int(sw), str(overhead_embed.device))
    for k, _, idx in _index_cache:
        if k == key:
            return idx
    idx = GalleryIndex(overhead_embed, sw)
    _index_cache.insert(0, (key, overhead_embed, idx))
    del _index_cache[INDEX_CACHE_ENTRIES:]
    return idx


def clear_cache():
    """Drop the prepared galleries kept for repeated small-query calls (each holds ~33 KB of device memory per item, and a
    reference to the feature tensor it was built from)."""
    del _index_cache[:]


def match(overhead_embed, surface_embed, path="auto", impl=None):
    """(orientation int64 [G,Q], distance fp32 [G,Q]) = a3 -> a4 -> a5 fused (cvig_fov.py:547-549).

    path 'fp32': exact fp32 kernels; 'tc': the tcgen05 sweep, with every entry fp16 cannot settle (uncertain argmax, small
    distance) overwritten with its fp32 value -- orientations are the fp32 reference's except where fp32 correlations
    themselves tie, distances within 1e-3 relative; 'tc16': the raw fp16 sweep; 'auto': by problem size (a few queries
    against a large gallery -- the heat map, the reference's one-query loop -- are evaluated entirely in fp32 from the
    gallery's spectra, which are prepared once and cached).
    """
    dev = _need_cuda("match", overhead_embed, surface_embed)
    g, q, ch, w, sw = _feature_dims("match", overhead_embed, surface_embed)
    raw = path == "tc16"
    which = _pick_path("tc" if raw else path, g, q, ch, w, sw)
    if path == "auto" and w == 64 and 0 < q <= EXACT_SMALL_Q and g * q >= 1024:
        gallery = _cached_index(overhead_embed, sw)
        queries = QueryBatch(surface_embed, impl=gallery.impl)
        with torch.cuda.device(dev):
            dist = torch.empty((g, q), dtype=torch.float32, device=dev)
            ori = torch.empty((g, q), dtype=torch.int64, device=dev)
            exact_columns(gallery, queries, dist=dist, ori64=ori, ld=q)
        return ori, dist
    if which == "tc":
        impl = _pick_impl(impl, ch, w, sw)
        gallery = GalleryIndex(overhead_embed, sw, keep_fp32=not raw, impl=impl)
        queries = QueryBatch(surface_embed, keep_fp32=not raw, impl=impl)
        if raw:
            res = sweep_tc(gallery, queries, want_dist=True, want_ori=True)
            return res["ori"].to(torch.int64), res["dist"]
        defer = Deferral(q, g, dev, matrix=True)
        res = sweep_tc(gallery, queries, want_dist=True, want_ori=True, deferral=defer)
        finish_tc(gallery, queries, defer, dist=res["dist"], ori=res["ori"])
        if g and q and int(defer.n_flagged):       # lists that overflowed: those columns entirely in fp32
            sel = torch.nonzero(defer.qflag[:q]).to(torch.int32).flatten().contiguous()
            exact_columns(gallery, queries, q_sel=sel, dist=res["dist"], ori8=res["ori"], ld=q, col_is_q=True)
        match.last_flagged = defer.n_flagged
        return res["ori"].to(torch.int64), res["dist"]
    ov, su = _f32c(overhead_embed), _f32c(surface_embed)
    with torch.cuda.device(dev):
        dist = torch.empty((g, q), dtype=torch.float32, device=dev)
        ori = torch.empty((g, q), dtype=torch.int64, device=dev)
        _lib.call("witw_match_f32", ov.data_ptr(), su.data_ptr(), g, q, ch, w, sw, dist.data_ptr(), ori.data_ptr(), 0, _stream())
    return ori, dist


match.last_flagged = None


def correlation_scores(overhead_embed, surface_embed):
    """Un-normalised circular cross-correlation fp32 [G,Q,W] (the conv2d output of cvig_fov.py:308), exact path."""
    dev = _need_cuda("correlation_scores", overhead_embed, surface_embed)
    g, q, ch, w, sw = _feature_dims("correlation_scores", overhead_embed, surface_embed)
    ov, su = _f32c(overhead_embed), _f32c(surface_embed)
    with torch.cuda.device(dev):
        corr = torch.empty((g, q, w), dtype=torch.float32, device=dev)
        _lib.call("witw_match_f32", ov.data_ptr(), su.data_ptr(), g, q, ch, w, sw, 0, 0, corr.data_ptr(), _stream())
    return corr


def correlation(overhead_embed, surface_embed, path="auto"):
    """Drop-in for cvig_fov.py:297-315: orientation int64 [batch_overhead, batch_surface].

    An argmax: no gradient flows through it in the reference either, so tensors that require grad are accepted
    (train() calls it on live encoder outputs, cvig_fov.py:450)."""
    return match(overhead_embed.detach(), surface_embed.detach(), path=path)[0]


def _crop_forward(overhead_embed, orientation, sw):
    dev = overhead_embed.device
    g, c, h, w = overhead_embed.shape
    q = orientation.shape[1]
    ov = _f32c(overhead_embed)
    ori = orientation.detach().to(torch.int64).contiguous()
    with torch.cuda.device(dev):
        out = torch.empty((g, q, c, h, sw), dtype=torch.float32, device=dev)
        _lib.call("witw_crop_gather_f32", ov.data_ptr(), ori.data_ptr(), out.data_ptr(), g, q, c * h, w, sw, _stream())
    return out, ori


class _CropOverheadFn(torch.autograd.Function):
    """crop_overhead with the gradient train() needs (cvig_fov.py:451 feeds the loss through the crop)."""

    @staticmethod
    def forward(ctx, overhead_embed, orientation, sw):
        out, ori = _crop_forward(overhead_embed, orientation, sw)
        ctx.save_for_backward(ori)
        ctx.shape, ctx.sw = tuple(overhead_embed.shape), sw
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (ori,) = ctx.saved_tensors
        g, c, h, w = ctx.shape
        go = _f32c(grad_out)
        with torch.cuda.device(go.device):
            grad_ov = torch.empty(ctx.shape, dtype=torch.float32, device=go.device)
            _lib.call("witw_crop_backward_f32", go.data_ptr(), ori.data_ptr(), grad_ov.data_ptr(), g, ori.shape[1], c * h, w, ctx.sw, _stream())
        return grad_ov, None, None


def crop_overhead(overhead_embed, orientation, surface_width):
    """Drop-in for cvig_fov.py:318-343: [G,Q,C,H,surface_width], rolled by the orientation and cropped.

    The device is taken from the inputs (the reference reads a module-level ``device`` global).  Differentiable
    with respect to ``overhead_embed`` (a scatter back through the roll), as train() requires.
    """
    _need_cuda("crop_overhead", overhead_embed, orientation, allow_grad=True)
    g = overhead_embed.shape[0]
    if overhead_embed.dim() != 4 or orientation.dim() != 2 or orientation.shape[0] != g:
        raise ValueError("crop_overhead: expected [G,C,H,W] features and a [G,Q] orientation")
    sw = int(surface_width)
    if overhead_embed.requires_grad and torch.is_grad_enabled():
        return _CropOverheadFn.apply(overhead_embed, orientation, sw)
    return _crop_forward(overhead_embed, orientation, sw)[0]


def _l2_forward(crop, su, g, q, k):
    with torch.cuda.device(crop.device):
        dist = torch.empty((g, q), dtype=torch.float32, device=crop.device)
        _lib.call("witw_l2_distance_f32", crop.data_ptr(), su.data_ptr(), dist.data_ptr(), g, q, k, _stream())
    return dist


class _L2DistanceFn(torch.autograd.Function):
    """l2_distance with gradients for both arguments (cvig_fov.py:453-460: the triplet loss back-propagates
    into the overhead encoder through the crop and into the surface encoder directly)."""

    @staticmethod
    def forward(ctx, overhead_cropped, surface_embed):
        g, q = overhead_cropped.shape[:2]
        k = int(np.prod(overhead_cropped.shape[2:]))
        crop, su = _f32c(overhead_cropped), _f32c(surface_embed)
        ctx.save_for_backward(crop, su)
        ctx.shapes = (tuple(overhead_cropped.shape), tuple(surface_embed.shape), g, q, k)
        return _l2_forward(crop, su, g, q, k)

    @staticmethod
    def backward(ctx, grad_dist):
        crop, su = ctx.saved_tensors
        crop_shape, su_shape, g, q, k = ctx.shapes
        gd = _f32c(grad_dist)
        need_crop, need_su = ctx.needs_input_grad
        with torch.cuda.device(gd.device):
            grad_crop = torch.empty(crop_shape, dtype=torch.float32, device=gd.device) if need_crop else None
            grad_su = torch.empty(su_shape, dtype=torch.float32, device=gd.device) if need_su else None
            coef = torch.empty((g, q, 2), dtype=torch.float32, device=gd.device) if need_su else None
            _lib.call("witw_l2_distance_backward_f32", crop.data_ptr(), su.data_ptr(), gd.data_ptr(), _ptr(grad_crop), _ptr(grad_su),
                      _ptr(coef), g, q, k, _stream())
        return grad_crop, grad_su


def l2_distance(overhead_cropped, surface_embed):
    """Drop-in for cvig_fov.py:346-363: chord distance fp32 [G,Q] of the L2-normalised maps (no epsilon).
    Differentiable in both arguments."""
    _need_cuda("l2_distance", overhead_cropped, surface_embed, allow_grad=True)
    g, q = overhead_cropped.shape[:2]
    k = int(np.prod(overhead_cropped.shape[2:]))
    if surface_embed.shape[0] != q or int(np.prod(surface_embed.shape[1:])) != k:
        raise RuntimeError("l2_distance: shapes %s and %s do not broadcast" % (tuple(overhead_cropped.shape), tuple(surface_embed.shape)))
    if torch.is_grad_enabled() and (overhead_cropped.requires_grad or surface_embed.requires_grad):
        return _L2DistanceFn.apply(overhead_cropped, surface_embed)
    return _l2_forward(_f32c(overhead_cropped), _f32c(surface_embed), g, q, k)


class _MatchDistanceFn(torch.autograd.Function):
    """correlation -> crop_overhead -> l2_distance as one differentiable op (cvig_fov.py:450-453): the forward is the exact
    fp32 match kernel, the backward works from the features and the orientation -- no [G,Q,C,H,sw] crop in either direction
    (64 MB per 64 x 64 batch at 360 degrees in the reference, plus the same again for its gradient)."""

    @staticmethod
    def forward(ctx, overhead_embed, surface_embed):
        ori, dist = match(overhead_embed.detach(), surface_embed.detach(), path="fp32")
        ctx.save_for_backward(_f32c(overhead_embed), _f32c(surface_embed), ori)
        ctx.mark_non_differentiable(ori)
        return dist, ori

    @staticmethod
    def backward(ctx, grad_dist, _grad_ori):
        ov, su, ori = ctx.saved_tensors
        g, c, h, w = ov.shape
        q, _, _, sw = su.shape
        need_ov, need_su = ctx.needs_input_grad
        gd = _f32c(grad_dist)
        with torch.cuda.device(gd.device):
            grad_ov = torch.empty_like(ov) if need_ov else None
            grad_su = torch.empty_like(su) if need_su else None
            coef = torch.empty((max(g * q, 1), 3), dtype=torch.float32, device=gd.device)
            _lib.call("witw_match_backward_f32", ov.data_ptr(), su.data_ptr(), ori.data_ptr(), gd.data_ptr(), _ptr(grad_ov), _ptr(grad_su),
                      coef.data_ptr(), g, q, c * h, w, sw, _stream())
        return grad_ov, grad_su


def match_distance(overhead_embed, surface_embed):
    """(distance fp32 [G,Q], orientation int64 [G,Q]) of the training step (cvig_fov.py:450-453:
    ``l2_distance(crop_overhead(ov, correlation(ov, su), sw), su)``) as one differentiable call: gradients flow to both
    feature sets through the distance, none through the orientation (an argmax, as in the reference)."""
    _need_cuda("match_distance", overhead_embed, surface_embed, allow_grad=True)
    _feature_dims("match_distance", overhead_embed, surface_embed)
    if torch.is_grad_enabled() and (overhead_embed.requires_grad or surface_embed.requires_grad):
        return _MatchDistanceFn.apply(overhead_embed, surface_embed)
    ori, dist = match(overhead_embed, surface_embed, path="fp32")
    return dist, ori


class _TripletLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, distances, alpha):
        d = _f32c(distances)
        n = d.shape[0]
        with torch.cuda.device(d.device):
            loss = torch.empty(1, dtype=torch.float32, device=d.device)
            grad = torch.empty_like(d) if ctx.needs_input_grad[0] else None
            _lib.call("witw_triplet_loss_f32", d.data_ptr(), n, float(alpha), loss.data_ptr(), _ptr(grad), _stream())
        ctx.save_for_backward(grad)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_loss):
        (grad,) = ctx.saved_tensors
        return (None if grad is None else grad * grad_loss), None


def triplet_loss(distances, alpha=10.):
    """Drop-in for cvig_fov.py:366-382: the soft-margin triplet loss of a [N,N] distance matrix with the matching pairs on
    its diagonal (0-dim tensor); loss and gradient come out of one kernel."""
    _need_cuda("triplet_loss", distances, allow_grad=True)
    if distances.dim() != 2 or distances.shape[0] != distances.shape[1]:
        raise ValueError("triplet_loss: expected a square [N,N] distance matrix, got %s" % (tuple(distances.shape),))
    return _TripletLossFn.apply(distances, alpha)


# ----------------------------------------------------------------------------- K4
def true_match_distances(overhead_embed, surface_embed, true_idx=None):
    """Exact fp32 (distance [Q], orientation [Q]) of each query against its matching gallery item."""
    dev = _need_cuda("true_match_distances", overhead_embed, surface_embed)
    g, q, ch, w, sw = _feature_dims("true_match_distances", overhead_embed, surface_embed)
    ov, su = _f32c(overhead_embed), _f32c(surface_embed)
    with torch.cuda.device(dev):
        pq = torch.arange(q, dtype=torch.int64, device=dev)
        if true_idx is None:
            if q > g:
                raise IndexError("true_match_distances: %d queries but only %d gallery items and no true_idx" % (q, g))
            pg = pq                                   # identity: valid by construction, no device round trip
        else:
            pg = true_idx.to(dev, torch.int64).contiguous()
            if q and (int(pg.max()) >= g or int(pg.min()) < 0):
                raise IndexError("true_match_distances: true index outside the gallery")
        d = torch.empty(q, dtype=torch.float32, device=dev)
        o = torch.empty(q, dtype=torch.int64, device=dev)
        _lib.call("witw_match_pairs_f32", ov.data_ptr(), su.data_ptr(), pg.data_ptr(), pq.data_ptr(), q, ch, w, sw,
                  d.data_ptr(), o.data_ptr(), _stream())
    return d, o


def rank_from_distances(distances, true_idx=None):
    """ranks int64 [Q] from a materialised [G,Q] matrix: #{g: d[g,q] <= d[true(q),q]} (cvig_fov.py:550-552)."""
    dev = _need_cuda("rank_from_distances", distances)
    g, q = distances.shape
    d = _f32c(distances)
    with torch.cuda.device(dev):
        ranks = torch.empty(q, dtype=torch.int64, device=dev)
        ti = None if true_idx is None else true_idx.to(dev, torch.int64).contiguous()
        _lib.call("witw_rank_from_dist_f32", d.data_ptr(), g, q, _ptr(ti), ranks.data_ptr(), _stream())
    return ranks


def topk_from_distances(distances, k, g_offset=0):
    """(dist fp32 [Q,k], idx int32 [Q,k]): the k nearest gallery items of every query column, ascending."""
    dev = _need_cuda("topk_from_distances", distances)
    g, q = distances.shape
    d = _f32c(distances)
    with torch.cuda.device(dev):
        td = torch.empty((q, k), dtype=torch.float32, device=dev)
        ti = torch.empty((q, k), dtype=torch.int32, device=dev)
        if g >= SELECT_MIN_ROWS and k <= 32 and q % 4 == 0 and q > 0 and d.data_ptr() % 16 == 0:
            # large galleries: sample thresholds -> streaming filter -> per-column selection (csrc/rank.cu)
            scratch = torch.empty(_lib.load().witw_topk_select_scratch_bytes(q, int(k)), dtype=torch.uint8, device=dev)
            _lib.call("witw_topk_select_f32", d.data_ptr(), g, q, int(k), td.data_ptr(), ti.data_ptr(), int(g_offset), scratch.data_ptr(), _stream())
            return td, ti
        slices = _lib.load().witw_topk_slices(g, q)
        if slices <= 1:
            _lib.call("witw_topk_from_dist_f32", d.data_ptr(), g, q, int(k), 1, td.data_ptr(), ti.data_ptr(), int(g_offset), _stream())
        else:
            cd = torch.empty((slices, q, k), dtype=torch.float32, device=dev)
            ci = torch.empty((slices, q, k), dtype=torch.int32, device=dev)
            _lib.call("witw_topk_from_dist_f32", d.data_ptr(), g, q, int(k), slices, cd.data_ptr(), ci.data_ptr(), int(g_offset), _stream())
            _lib.call("witw_topk_merge", cd.data_ptr(), ci.data_ptr(), slices, q, int(k), td.data_ptr(), ti.data_ptr(), _stream())
    return td, ti


def evaluate_ranks(overhead_embed, surface_embed, true_idx=None, path="auto", topk=0, exact=True, impl=None):
    """The rank loop of test() (cvig_fov.py:543-552) as one call: ranks int64 [count] on the device.

    Query i matches gallery item i (or true_idx[i]).  The true-match distances are computed in
    exact fp32; on the tensor-core path the gallery sweep counts d[g,q] <= d_true[q] in the GEMM
    epilogue without materialising the [G,Q] matrix; with exact=True (default) every decision the fp16 operands cannot
    settle within their error bound is taken in fp32 and the top-k is re-ranked in fp32, so the results are the fp32
    reference's (see evaluate_ranks_prepared).  With topk > 0 also returns (topk_dist [Q,k], topk_idx [Q,k]).
    """
    dev = _need_cuda("evaluate_ranks", overhead_embed, surface_embed)
    g, q, ch, w, sw = _feature_dims("evaluate_ranks", overhead_embed, surface_embed)
    if true_idx is None and q > g:
        raise ValueError("evaluate_ranks: %d queries but only %d gallery items and no true_idx" % (q, g))
    which = _pick_path(path, g, q, ch, w, sw)
    if which == "fp32":
        _, dist = match(overhead_embed, surface_embed, path="fp32")
        ranks = rank_from_distances(dist, true_idx)
        if topk:
            return (ranks,) + topk_from_distances(dist, topk)
        return ranks
    impl = _pick_impl(impl, ch, w, sw)
    gallery = GalleryIndex(overhead_embed, sw, impl=impl)
    queries = QueryBatch(surface_embed, impl=impl)
    return evaluate_ranks_prepared(gallery, queries, true_idx=true_idx, topk=topk, exact=exact)


def pair_distances_prepared(gallery, queries, pair_g, pair_q):
    """Exact fp32 (distance, orientation) of explicit (local gallery index, query index) pairs on prepared operands,
    through the packed azimuth spectra."""
    dev = gallery.device
    n = pair_g.numel()
    with torch.cuda.device(dev):
        d = torch.empty(n, dtype=torch.float32, device=dev)
        o = torch.empty(n, dtype=torch.int64, device=dev)
        _lib.call("witw_match_pairs_spec_f32", gallery.spectral().data_ptr(), gallery.crop_inv_norm.data_ptr(),
                  queries.spectral().data_ptr(), queries.inv_norm.data_ptr(), pair_g.data_ptr(), pair_q.data_ptr(), n, gallery.CH,
                  d.data_ptr(), o.data_ptr(), _stream())
    return d, o


FALLBACK_COLUMNS = 256      # flagged queries re-done per call of the column kernel (bounds the [G, n] fp32 scratch matrix)


def redo_flagged(gallery, queries, deferral, d_true, t32, counts, topk=0, td=None, ti=None):
    """Queries the finish flagged (deferral list overflowed, or the candidate keys do not prove the top-k complete) are
    evaluated against the whole gallery in fp32: their counts, and their top-k, are replaced.  Returns how many there were."""
    dev = gallery.device
    sel_all = torch.nonzero(deferral.qflag[: queries.Q]).to(torch.int32).flatten().contiguous()
    for lo in range(0, sel_all.numel(), FALLBACK_COLUMNS):
        sel = sel_all[lo: lo + FALLBACK_COLUMNS].contiguous()
        f = sel.numel()
        cnt = torch.zeros(f, dtype=torch.int32, device=dev)
        scratch = torch.empty((gallery.G, f), dtype=torch.float32, device=dev) if topk else None
        exact_columns(gallery, queries, q_sel=sel, dist=scratch, ld=f, d_true=d_true, true_idx=t32, count_out=cnt)
        counts[sel.long()] = cnt
        if topk:
            fd, fi = topk_from_distances(scratch, topk, g_offset=gallery.g_offset)
            td[sel.long()] = fd
            ti[sel.long()] = fi
    return int(sel_all.numel())


class RankEvaluation(object):
    """evaluate_ranks on prepared operands (tensor-core path), in two halves: the constructor enqueues everything on the
    current stream (true-match distances, the sweep, the top-k merge, the fp32 finish) and returns without waiting;
    result() reads back how many queries the finish flagged -- 4 bytes, the one host synchronisation of an evaluation --
    re-does those entirely in fp32, and returns ranks [, topk_dist, topk_idx].  A caller that pipelines evaluations (bench.py)
    constructs step i+1 before it asks step i for its result.

    exact=True (needs the fp32 spectra kept in the operands): the sweep knows, per pair, how far its fp16 operands can have
    moved the distance (csrc/sweep_common.cuh).  Rank decisions inside that slack of the fp32 threshold are taken from the
    fp32 spectra instead, the top-k candidates (kept under lower-bound keys) are re-ranked in fp32, and a query whose
    deferral list overflowed or whose candidate list cannot be proven complete is re-done entirely in fp32 -- so ranks and
    top-k are those of the fp32 reference chain (cvig_fov.py:547-552), not approximations, at any size.
    exact=False: the raw fp16 sweep (ranks may differ where distances tie within ~1e-4).
    """

    def __init__(self, gallery, queries, true_idx=None, topk=0, d_true=None, events=None, exact=True):
        dev = gallery.device
        has_fp32 = gallery.spec is not None and queries.spec is not None
        if d_true is None and not has_fp32:
            raise ValueError("evaluate_ranks_prepared: the fp32 spectra were dropped; pass d_true")
        if exact and not has_fp32:
            raise ValueError("evaluate_ranks_prepared: exact=True needs the fp32 spectra (keep_fp32=True)")
        if exact and topk > TOPK_EXACT_MAX:
            raise ValueError("evaluate_ranks_prepared: the exact fused top-k covers k <= %d (got %d); use exact=False, or "
                             "match() + topk_from_distances() for longer lists" % (TOPK_EXACT_MAX, topk))
        self.gallery, self.queries, self.topk, self.exact = gallery, queries, int(topk), bool(exact)
        self.td = self.ti = self.defer = None
        self.flagged = 0
        nq = queries.Q
        with torch.cuda.device(dev):
            pq = torch.arange(nq, dtype=torch.int64, device=dev)
            if true_idx is None:
                if nq > gallery.G and d_true is None:
                    raise IndexError("evaluate_ranks_prepared: %d queries but only %d gallery items and no true_idx" % (nq, gallery.G))
                pg = pq
            else:
                pg = true_idx.to(dev, torch.int64).contiguous()
            if d_true is None:
                if true_idx is not None and nq and (int(pg.max()) >= gallery.G or int(pg.min()) < 0):
                    raise IndexError("evaluate_ranks_prepared: true index outside the gallery")
                d_true, _ = pair_distances_prepared(gallery, queries, pg, pq)
            self.d_true = d_true
            self.t32 = (pg + gallery.g_offset).to(torch.int32)
            if not self.exact:
                self.counts = torch.zeros(max(nq, 1), dtype=torch.int32, device=dev)
                res = sweep_tc(gallery, queries, d_true=d_true, true_idx=self.t32, rank_count=self.counts, topk=topk, events=events)
                if topk:
                    self.td, self.ti = res["topk_dist"], res["topk_idx"]
                return
            self.defer = Deferral(nq, gallery.G, dev)
            self.counts = self.defer.counts
            res = sweep_tc(gallery, queries, d_true=d_true, true_idx=self.t32, rank_count=self.counts, topk=16 if topk else 0, events=events,
                           deferral=self.defer)
            fin = finish_tc(gallery, queries, self.defer, d_true=d_true, rank_count=self.counts, cand_key=res.get("topk_dist"),
                            cand_idx=res.get("topk_idx"), k_out=topk)
            if topk:
                self.td, self.ti = fin
            self._n_host = torch.empty(1, dtype=torch.int32, pin_memory=True)
            self._n_host.copy_(self.defer.n_flagged, non_blocking=True)
            self._landed = torch.cuda.Event()
            self._landed.record()

    def result(self):
        nq = self.queries.Q
        if self.exact and self.defer is not None:
            self._landed.synchronize()
            n = int(self._n_host[0])
            if n and self.gallery.G > 0 and nq > 0:
                with torch.cuda.device(self.gallery.device):
                    self.flagged = redo_flagged(self.gallery, self.queries, self.defer, self.d_true, self.t32, self.counts, self.topk, self.td, self.ti)
            evaluate_ranks_prepared.last_stats = {"deferred": self.defer.list_n[:nq], "flagged": self.flagged, "list_cap": self.defer.cap}
            self.defer = None
        ranks = self.counts[:nq].to(torch.int64)
        return (ranks, self.td, self.ti) if self.topk else ranks


def evaluate_ranks_prepared(gallery, queries, true_idx=None, topk=0, d_true=None, events=None, exact=True):
    """evaluate_ranks on prepared operands: ``RankEvaluation(...).result()`` (see there).
    ``evaluate_ranks_prepared.last_stats`` = {"deferred": per-query deferral counts (device tensor), "flagged": queries re-done
    entirely in fp32, "list_cap": capacity of a query's list} of the last exact evaluation."""
    return RankEvaluation(gallery, queries, true_idx=true_idx, topk=topk, d_true=d_true, events=events, exact=exact).result()


evaluate_ranks_prepared.last_stats = None


def baseline_ranks(overhead_embed, surface_embed, true_idx=None, return_distances=False):
    """cvig_baseline.py:453-460: Euclidean distances of [N,D] embeddings and the same rank rule."""
    dev = _need_cuda("baseline_ranks", overhead_embed, surface_embed)
    if overhead_embed.dim() != 2 or surface_embed.dim() != 2 or overhead_embed.shape[1] != surface_embed.shape[1]:
        raise ValueError("baseline_ranks: expected [N,D] and [Q,D]")
    n, d = overhead_embed.shape
    q = surface_embed.shape[0]
    ov, su = _f32c(overhead_embed), _f32c(surface_embed)
    with torch.cuda.device(dev):
        dist = torch.empty((n, q), dtype=torch.float32, device=dev)
        ranks = torch.empty(q, dtype=torch.int64, device=dev)
        ti = None if true_idx is None else true_idx.to(dev, torch.int64).contiguous()
        _lib.call("witw_l2_rank_f32", ov.data_ptr(), su.data_ptr(), n, q, d, _ptr(ti), dist.data_ptr(), ranks.data_ptr(), _stream())
    return (ranks, dist) if return_distances else ranks


def recall_from_ranks(ranks):
    """cvig_fov.py:553-558: top-1/5/10/1 % (percent), mean and median rank.  Host arithmetic on the D2H'd ranks."""
    if isinstance(ranks, torch.Tensor):
        ranks = ranks.detach().cpu().numpy()
    ranks = np.asarray(ranks)
    count = ranks.shape[0]
    return {
        "top_one": np.sum(ranks <= 1) / count * 100,
        "top_five": np.sum(ranks <= 5) / count * 100,
        "top_ten": np.sum(ranks <= 10) / count * 100,
        "top_percent": np.sum(ranks * 100 <= count) / count * 100,
        "mean": np.mean(ranks),
        "median": np.median(ranks),
        "count": count,
    }


def heatmap_scores(overhead_embed, surface_embed, output_width_max=64, path="auto"):
    """tools/heatmap/heatmap.py:171-177: one photo against many tiles -> (orientation degrees, dissimilarity, score)."""
    ori, dist = match(overhead_embed, surface_embed, path=path)
    orientations = torch.squeeze(ori) * 360 / output_width_max - 180
    dist = torch.squeeze(dist)
    return orientations, dist, torch.exp(10.0 * (1.0 - dist))
