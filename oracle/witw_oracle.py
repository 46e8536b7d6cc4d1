"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the WITW cross-view retrieval hot path.

A restatement (torch-CPU / numpy, fp32 like the reference, plus an fp64 fused form)
of the seven reference functions on the path.  It is the
checker for the CUDA kernels and the "port" CPU baseline of bench.py; it is never
on the product path.  Parity of this file against the real reference is pinned by
tests/test_oracle_golden.py using fixtures in tests/golden/ that were produced by
tests/golden/make_golden.py from the unmodified reference (imported from
/root/reference in the build container).

Every function cites the reference lines it restates (paths relative to
/root/reference).  The heavy library calls are the same ones the reference makes
(F.conv2d, argmax, linalg.norm, advanced indexing) so CPU timings of this port are
representative of the reference's own PyTorch-CPU path.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# model/cvig_fov.py:19-22 (Globals)
SURFACE_H = 128
SURFACE_W = 512
OVERHEAD = 256


# --------------------------------------------------------------------------- a2 grid
def polar_grid(h_s=SURFACE_H, w_s=SURFACE_W, s_o=OVERHEAD):
    """Sample coordinates of the polar transform, float64 [h_s, w_s] each.

    model/cvig_fov.py:197-201: meshgrid over (w_s, h_s); radius (s_o/2)*(h_s-1-yy)/h_s;
    yy_o = s_o/2 + r*cos(2*pi*xx/w_s); xx_o = s_o/2 - r*sin(2*pi*xx/w_s).
    Evaluation order of the float64 products is kept: ((s_o/2)*(h_s-1-yy))/h_s * trig.
    """
    cols, rows = np.meshgrid(range(w_s), range(h_s))
    half = s_o / 2
    radial = half * (h_s - 1 - rows) / h_s
    ang = 2 * math.pi * cols / w_s
    y_src = half + radial * np.cos(ang)
    x_src = half - radial * np.sin(ang)
    return x_src, y_src


# --------------------------------------------------------------------------- a1 LUT
def bilinear_lut(x, y, src_h, src_w):
    """Clipped tap indices and fp32 weights for bilinear sampling at (x, y).

    model/cvig_fov.py:163-181.  floor, +1, then clip all four integer coordinates to
    the image; the weights are formed in float64 from the *clipped* integers and only
    then rounded to fp32 (torch.FloatTensor).  Returns int arrays x0,x1,y0,y1 and fp32
    arrays wa,wb,wc,wd, each shaped like x.
    """
    x = np.asarray(x)
    y = np.asarray(y)
    assert x.shape == y.shape
    x_lo = np.floor(x).astype(int)
    y_lo = np.floor(y).astype(int)
    x_hi = x_lo + 1
    y_hi = y_lo + 1
    x_lo = np.clip(x_lo, 0, src_w - 1)
    x_hi = np.clip(x_hi, 0, src_w - 1)
    y_lo = np.clip(y_lo, 0, src_h - 1)
    y_hi = np.clip(y_hi, 0, src_h - 1)
    w_a = ((x_hi - x) * (y_hi - y)).astype(np.float32)
    w_b = ((x_hi - x) * (y - y_lo)).astype(np.float32)
    w_c = ((x - x_lo) * (y_hi - y)).astype(np.float32)
    w_d = ((x - x_lo) * (y - y_lo)).astype(np.float32)
    return (x_lo, x_hi, y_lo, y_hi), (w_a, w_b, w_c, w_d)


def bilinear_interpolate(im, x, y):
    """model/cvig_fov.py:156-183.  im [C,H,W] fp32 -> [C,h,w] fp32.

    Blend order of line 183 is kept: ((wa*Ia + wb*Ib) + wc*Ic) + wd*Id, every
    product and sum rounded to fp32 separately.
    """
    (x_lo, x_hi, y_lo, y_hi), (w_a, w_b, w_c, w_d) = bilinear_lut(x, y, im.shape[1], im.shape[2])
    t_a = im[:, y_lo, x_lo]
    t_b = im[:, y_hi, x_lo]
    t_c = im[:, y_lo, x_hi]
    t_d = im[:, y_hi, x_hi]
    w_a, w_b, w_c, w_d = (torch.from_numpy(w).unsqueeze(0) for w in (w_a, w_b, w_c, w_d))
    return w_a * t_a + w_b * t_b + w_c * t_c + w_d * t_d


def polar_transform(overhead, h_s=SURFACE_H, w_s=SURFACE_W, s_o=OVERHEAD):
    """model/cvig_fov.py:186-209 on one tile [C,s_o,s_o] -> [C,h_s,w_s]."""
    x_src, y_src = polar_grid(h_s, w_s, s_o)
    return bilinear_interpolate(overhead, x_src, y_src)


# --------------------------------------------------------------------------- f4 (upstream of a2)
IMG_MEAN = (0.485, 0.456, 0.406)   # model/cvig_fov.py:24-25
IMG_STD = (0.229, 0.224, 0.225)


def image_normalization(img_u8, mean=IMG_MEAN, std=IMG_STD):
    """model/cvig_fov.py:137-149 on one image [C,H,W] uint8 -> fp32: ``norm(data / 255.)`` with torchvision's
    Normalize, i.e. ((x / 255) - mean) / std, every step rounded to fp32."""
    x = img_u8 / 255.
    m = torch.as_tensor(mean, dtype=x.dtype).view(-1, 1, 1)
    sd = torch.as_tensor(std, dtype=x.dtype).view(-1, 1, 1)
    return (x - m) / sd


def normalized_polar(overhead_u8, mean=IMG_MEAN, std=IMG_STD):
    """ImageNormalization then PolarTransform on a uint8 tile that is already 256 x 256 (Resize is the identity there:
    cvig_fov.py:133, 147, 208)."""
    return polar_transform(image_normalization(overhead_u8, mean, std))


# --------------------------------------------------------------------------- f4: Resize (upstream of ImageNormalization)
def resize_taps(in_size, out_size, antialias):
    """Per-output-index taps of the bilinear resize the reference's ``Resize`` performs along one axis
    (model/cvig_fov.py:119, 131, 133: torchvision.transforms.functional.resize on a float tensor, bilinear,
    align_corners=False).  The arithmetic lives in torch (third party, not under /root/reference):

    * antialias=False -- what the reference's pinned torchvision 0.9.1 / torch 1.8.1 (model/requirements.txt) computes:
      ``src = max(fma(scale, i+0.5, -0.5), 0)``, ``i0 = int(src)``, ``i1 = i0 + (i0 < in-1)``, ``l1 = src - i0``, ``l0 = 1 - l1``
      (ATen upsample_bilinear2d, area_pixel_compute_source_index), all in fp32, ``scale = float(in)/out``.
    * antialias=True -- what torchvision >= 0.17 does by default, i.e. the reference as it runs in the build container:
      triangle filter of half-width ``support = max(scale, 1)`` around ``center = scale*(i+0.5)``; taps
      ``[int(center-support+0.5), int(center+support+0.5))`` clipped to the axis, weights ``1-|(j-center+0.5)/max(scale,1)|``
      normalised by their sum (ATen _upsample_bilinear2d_aa, _compute_indices_min_size_weights_aa), fp32.

    Returns (start int64 [out], count int64 [out], weights fp32 [out, kmax]) -- unused weight slots are 0."""
    f32 = np.float32
    scale = f32(in_size) / f32(out_size)
    if not antialias:
        start = np.zeros(out_size, dtype=np.int64)
        count = np.full(out_size, 2, dtype=np.int64)
        wts = np.zeros((out_size, 2), dtype=np.float32)
        for i in range(out_size):
            if in_size == out_size:     # ATen: a scale of exactly 1 is a plain copy
                start[i], count[i], wts[i, 0] = i, 1, 1.0
                continue
            # ATen's builds contract ``scale * (i + 0.5) - 0.5`` into one fused multiply-add (one rounding; measured on the
            # torch of the build container: with an inexact scale such as 100/300 the separately rounded form is off by up
            # to 4e-6 in the weight); the float64 product of two floats is exact, so this is that fma
            src = max(f32(np.float64(scale) * (i + 0.5) - 0.5), f32(0.0))
            i0 = int(src)
            l1 = min(max(f32(src - f32(i0)), f32(0.0)), f32(1.0))
            start[i] = i0
            if i0 < in_size - 1:
                wts[i] = (f32(1.0) - l1, l1)
            else:                       # i1 == i0: both lambdas land on the last sample
                count[i] = 1
                wts[i] = ((f32(1.0) - l1) + l1, 0.0)
        return start, count, wts
    # ATen keeps scale / support / center / the weights in float but writes its 0.5 and 1.0 literals as doubles, so some
    # sub-expressions are evaluated in double before they are rounded back; the float64 casts below mirror that
    f64 = np.float64
    support = f32(scale) if scale >= 1.0 else f32(1.0)
    invscale = f32(1.0 / f64(scale)) if scale >= 1.0 else f32(1.0)
    kmax = int(math.ceil(support)) * 2 + 1
    start = np.zeros(out_size, dtype=np.int64)
    count = np.zeros(out_size, dtype=np.int64)
    wts = np.zeros((out_size, kmax), dtype=np.float32)
    for i in range(out_size):
        center = f32(f64(scale) * (i + 0.5))
        lo = max(int(f64(f32(center - support)) + 0.5), 0)
        n = min(int(f64(f32(center + support)) + 0.5), in_size) - lo
        w = np.zeros(n, dtype=np.float32)
        total = f32(0.0)
        for j in range(n):
            x = abs(f32((f64(f32(f32(j + lo) - center)) + 0.5) * f64(invscale)))
            w[j] = f32(1.0) - x if x < 1.0 else f32(0.0)
            total = f32(total + w[j])
        if total != 0.0:
            w = (w / total).astype(np.float32)
        start[i], count[i] = lo, n
        wts[i, :n] = w
    return start, count, wts


def _resample_axis(x, start, count, wts, axis):
    """One separable pass in fp32: out[i] = sum_j w[i,j] * x[start[i]+j], accumulated left to right."""
    x = np.moveaxis(np.asarray(x, dtype=np.float32), axis, -1)
    out = np.zeros(x.shape[:-1] + (len(start),), dtype=np.float32)
    for i in range(len(start)):
        acc = x[..., start[i]] * wts[i, 0]
        for j in range(1, count[i]):
            acc = (acc + x[..., start[i] + j] * wts[i, j]).astype(np.float32)
        out[..., i] = acc
    return np.moveaxis(out, -1, axis)


def resize_bilinear(img, out_h, out_w, antialias=True):
    """``torchvision.transforms.functional.resize(img, (out_h, out_w))`` on a float image [..., H, W] as the reference calls it
    (model/cvig_fov.py:119, 131, 133).  antialias=True: separable, width first then height with an fp32 intermediate
    (ATen separable_upsample_generic_Nd_kernel_impl).  antialias=False: ``hl0*(wl0*a + wl1*b) + hl1*(wl0*c + wl1*d)``
    (ATen upsample_bilinear2d) -- the same two passes with the rows blended last."""
    x = np.asarray(img, dtype=np.float32)
    in_h, in_w = x.shape[-2:]
    sx, cx, wx = resize_taps(in_w, out_w, antialias)
    sy, cy, wy = resize_taps(in_h, out_h, antialias)
    if antialias and in_w == out_w:
        tmp = x             # ATen skips a pass whose size does not change
    else:
        tmp = _resample_axis(x, sx, cx, wx, -1)
    if antialias and in_h == out_h:
        return torch.from_numpy(np.ascontiguousarray(tmp))
    return torch.from_numpy(np.ascontiguousarray(_resample_axis(tmp, sy, cy, wy, -2)))


def resize_pair(surface, overhead, fov=360, panorama=True, start=0, antialias=True):
    """``Resize.__call__`` (model/cvig_fov.py:117-134) on float images [C,H,W]: the surface image to 128 x 512 and a
    ``surface_width = int(fov/360*512)`` wide window starting at column ``start`` with wrap-around (panorama, 119-129), or
    directly to 128 x surface_width (131); the overhead image to 256 x 256 (133).  ``start`` is the value the reference
    draws with torch.randint (121-124).  Returns (surface, overhead)."""
    sw = int(fov / 360 * SURFACE_W)
    if panorama:
        su = resize_bilinear(surface, SURFACE_H, SURFACE_W, antialias)
        end = start + sw
        if end < SURFACE_W:
            su = su[:, :, start:end]
        else:
            su = torch.cat((su[:, :, start:], su[:, :, : end - SURFACE_W]), dim=2)
    else:
        su = resize_bilinear(surface, SURFACE_H, sw, antialias)
    return su, resize_bilinear(overhead, OVERHEAD, OVERHEAD, antialias)


# --------------------------------------------------------------------------- a3
def correlation_scores(overhead_embed, surface_embed):
    """Un-normalised circular cross-correlation, fp32 [G,Q,W].

    model/cvig_fov.py:299-312: wrap-pad the gallery width by sw-1 columns, conv2d
    with the queries as filters; output height is 1 and is squeezed.
    """
    sw = surface_embed.shape[3]
    wrapped = torch.cat((overhead_embed, overhead_embed[:, :, :, : sw - 1]), dim=3)
    return F.conv2d(wrapped, surface_embed, stride=1).squeeze(-2)


def correlation(overhead_embed, surface_embed):
    """model/cvig_fov.py:297-315: argmax over the azimuth shift, int64 [G,Q], first max wins."""
    return torch.argmax(correlation_scores(overhead_embed, surface_embed), dim=-1)


# --------------------------------------------------------------------------- a4
def crop_overhead(overhead_embed, orientation, surface_width):
    """model/cvig_fov.py:318-343: out[g,q,c,h,k] = ov[g,c,h,(k+ori[g,q]) % W], k < surface_width."""
    n_g, n_q = orientation.shape
    c, h, w = overhead_embed.shape[1:]
    col = (torch.arange(w).view(1, 1, w) + orientation.unsqueeze(-1)) % w          # [G,Q,W]
    col = col.view(n_g, n_q, 1, 1, w).expand(n_g, n_q, c, h, w)
    tiled = overhead_embed.unsqueeze(1).expand(n_g, n_q, c, h, w)
    return torch.gather(tiled, 4, col)[..., :surface_width]


# --------------------------------------------------------------------------- a5
def l2_distance(overhead_cropped, surface_embed):
    """model/cvig_fov.py:346-363: L2-normalise both sides over (C,H,sw); 2*(1 - dot). No epsilon."""
    n_g, n_q = overhead_cropped.shape[:2]
    o = overhead_cropped.reshape(n_g, n_q, -1)
    o = o / torch.linalg.norm(o, ord=2, dim=-1, keepdim=True)
    s = surface_embed.reshape(n_q, -1)
    s = s / torch.linalg.norm(s, ord=2, dim=-1, keepdim=True)
    return 2 * (1 - torch.sum(o * s.unsqueeze(0), dim=2))


def match(overhead_embed, surface_embed):
    """a3 -> a4 -> a5 chained exactly as the callers do (cvig_fov.py:547-549). Returns (orientation, distance)."""
    ori = correlation(overhead_embed, surface_embed)
    crop = crop_overhead(overhead_embed, ori, surface_embed.shape[3])
    return ori, l2_distance(crop, surface_embed)


# --------------------------------------------------------------------------- a6 / a7
def rank_loop(overhead_embed, surface_embed, query_indices=None):
    """model/cvig_fov.py:543-552.  One query at a time against the whole gallery;
    rank = #{g : d[g] <= d[idx]} (ties and self count; NaN compares false)."""
    count = surface_embed.size(0)
    idxs = range(count) if query_indices is None else query_indices
    ranks = np.zeros([len(idxs)], dtype=int)
    for n, idx in enumerate(idxs):
        one = surface_embed[idx : idx + 1]
        _, dist = match(overhead_embed, one)
        dist = dist.squeeze()
        ranks[n] = torch.sum(torch.le(dist, dist[idx])).item()
    return ranks


def baseline_rank_loop(overhead_embed, surface_embed):
    """model/cvig_baseline.py:453-460.  Plain Euclidean distance on [N,D] embeddings, same rank rule."""
    count = surface_embed.size(0)
    ranks = np.zeros([count], dtype=int)
    for idx in range(count):
        diff = overhead_embed - surface_embed[idx : idx + 1]
        dist = torch.pow(torch.sum(torch.pow(diff, 2), dim=1), 0.5)
        ranks[idx] = torch.sum(torch.le(dist, dist[idx])).item()
    return ranks


def recall_from_ranks(ranks):
    """model/cvig_fov.py:553-558 (= cvig_baseline.py:461-466).  Percentages, mean and median rank."""
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


# --------------------------------------------------------------------------- heatmap caller
def heatmap_scores(overhead_embed, surface_embed, output_width_max=64):
    """tools/heatmap/heatmap.py:171-177: one query vs. many tiles -> (degrees, dissimilarity, score)."""
    ori, dist = match(overhead_embed, surface_embed)
    degrees = torch.squeeze(ori) * 360 / output_width_max - 180
    dist = torch.squeeze(dist)
    return degrees, dist, torch.exp(10.0 * (1.0 - dist))


# --------------------------------------------------------------------------- fused identity, fp64
def fused_fp64(overhead_embed, surface_embed):
    """The algebraic identity the CUDA kernels implement, in float64:

        s* = argmax_s corr[g,q,s]
        d  = 2 - 2*corr[g,q,s*] / ( sqrt(sum_{k<sw} colE[g,(s*+k)%W]) * ||su_q|| )

    Used to adjudicate tolerance questions (which of two fp32 answers is nearer the truth)
    and as an independent cross-check of the restatement above.  Returns (corr [G,Q,W] f64,
    orientation int64 [G,Q], distance f64 [G,Q]).
    """
    ov = overhead_embed.double()
    su = surface_embed.double()
    n_g, c, h, w = ov.shape
    n_q, _, _, sw = su.shape
    shift = (torch.arange(w).view(w, 1) + torch.arange(sw).view(1, sw)) % w      # [W, sw]
    windows = ov[:, :, :, shift]                                                # [G,C,H,W,sw]
    corr = torch.einsum("gchsk,qchk->gqs", windows, su)
    ori = torch.argmax(corr, dim=-1)
    col_energy = (ov * ov).sum(dim=(1, 2))                                      # [G,W]
    crop_energy = col_energy[:, shift].sum(-1)                                  # [G,W]
    qn = su.reshape(n_q, -1).norm(dim=1)
    best = torch.gather(corr, 2, ori.unsqueeze(-1)).squeeze(-1)
    cn = torch.sqrt(torch.gather(crop_energy.unsqueeze(1).expand(n_g, n_q, w), 2, ori.unsqueeze(-1)).squeeze(-1))
    dist = 2 - 2 * best / (cn * qn.unsqueeze(0))
    return corr, ori, dist


# --------------------------------------------------------------------------- synthetic data
def synth_features(n_gallery, n_query, fov=360, c=16, h=4, w=64, noise=0.5, seed=1234, planted=True):
    """Synthetic feature maps of the BASELINE shapes.

    Gallery ~ N(0,1)*0.06 (the scale random-init FOV_DSM emits).  With ``planted`` the first
    min(G,Q) queries are the matching gallery item rolled by a per-query azimuth, cropped to
    the field of view and perturbed with noise, so recall is non-degenerate.
    Returns (overhead_embed [G,c,h,w], surface_embed [Q,c,h,sw], planted_shift int64 [Q]).
    """
    sw = int(fov / 360 * 512) // 8
    gen = torch.Generator().manual_seed(seed)
    ov = torch.randn(n_gallery, c, h, w, generator=gen) * 0.06
    su = torch.randn(n_query, c, h, sw, generator=gen) * 0.06
    shifts = torch.randint(0, w, (n_query,), generator=gen)
    if planted:
        n = min(n_gallery, n_query)
        cols = (shifts[:n].view(n, 1) + torch.arange(sw).view(1, sw)) % w     # [n, sw]
        picked = torch.gather(ov[:n], 3, cols.view(n, 1, 1, sw).expand(n, c, h, sw))
        su[:n] = picked + noise * su[:n]
    return ov, su, shifts
