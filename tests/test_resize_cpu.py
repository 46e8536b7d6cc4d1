"""CPU checks of the Resize / ImageNormalization front end: the oracle's restatement of torchvision's
bilinear resize against golden vectors frozen from the unmodified reference's transform chain, and the host-built tap
tables of libwitw_b200 against the oracle's, bit for bit.  No GPU needed."""
import numpy as np
import pytest
import torch

from oracle import witw_oracle as O

torch.set_num_threads(1)

# the oracle accumulates taps left to right with separate roundings; ATen's vectorised kernels contract and reorder some
# of them: measured <= 2 ulp of the 0..255 pixel scale (3.1e-5)
RESIZE_TOL = 6.2e-5


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def redraw_inputs(g):
    gen = torch.Generator().manual_seed(int(g["big_seed"]))
    ov8 = torch.randint(0, 256, (3, 301, 283), generator=gen, dtype=torch.uint8)
    su8 = torch.randint(0, 256, (3, 97, 411), generator=gen, dtype=torch.uint8)
    big8 = torch.randint(0, 256, (3, 750, 750), generator=gen, dtype=torch.uint8)
    assert torch.equal(ov8, t(g["ov_u8"])) and torch.equal(su8, t(g["su_u8"]))
    assert int(big8.long().sum()) == int(g["big_sum"]) and torch.equal(big8[:, :4, :4], t(g["big_corner"]))
    return ov8, su8, big8


def test_resize_chain_matches_reference_golden(golden):
    g = golden("resize")
    ov8, su8, big8 = redraw_inputs(g)
    start = int(g["pano_start"])
    su, ov = O.resize_pair(su8.float(), ov8.float(), fov=90, panorama=True, start=start)
    assert tuple(su.shape) == (3, 128, 128) and tuple(ov.shape) == (3, 256, 256)
    assert (su - t(g["pano_surface"])).abs().max().item() <= RESIZE_TOL
    assert (ov[:, ::3, ::5] - t(g["pano_overhead_sub"])).abs().max().item() <= RESIZE_TOL
    # ... -> ImageNormalization -> PolarTransform (values are /255/std smaller: 2 ulp of 255 -> ~1.1e-6)
    su_n = O.image_normalization(su)
    polar = O.polar_transform(O.image_normalization(ov))
    assert (su_n[:, ::3, ::3] - t(g["pano_surface_norm_sub"])).abs().max().item() <= 2e-6
    assert (polar[:, ::3, ::7] - t(g["pano_polar_sub"])).abs().max().item() <= 2e-6
    # limited-FoV photo (not a panorama) and a 2.9x downsample of the aerial image
    su, ov = O.resize_pair(su8.float(), big8.float(), fov=70, panorama=False)
    assert tuple(su.shape) == (3, 128, 99)
    assert (su[:, ::2, ::3] - t(g["witw_surface_sub"])).abs().max().item() <= RESIZE_TOL
    assert (ov[:, ::5, ::3] - t(g["witw_overhead_sub"])).abs().max().item() <= RESIZE_TOL


def test_resize_without_antialias_matches_pinned_torchvision_call(golden):
    g = golden("resize")
    _, su8, big8 = redraw_inputs(g)
    ov = O.resize_bilinear(big8.float(), 256, 256, antialias=False)
    su = O.resize_bilinear(su8.float(), 128, 512, antialias=False)
    assert (ov[:, ::5, ::3] - t(g["noaa_overhead_sub"])).abs().max().item() <= RESIZE_TOL
    assert (su[:, ::3, ::5] - t(g["noaa_surface_sub"])).abs().max().item() <= RESIZE_TOL


def test_resize_oracle_tracks_torch_interpolate():
    """The third-party arithmetic itself (ATen) is on both machines: random geometries, both variants."""
    import torch.nn.functional as F

    gen = torch.Generator().manual_seed(3)
    for (ih, iw, oh, ow) in [(37, 53, 128, 512), (300, 400, 128, 128), (256, 256, 256, 256), (513, 255, 256, 256), (128, 600, 128, 512)]:
        x = torch.randint(0, 256, (2, ih, iw), generator=gen).float()
        for aa in (True, False):
            ref = F.interpolate(x[None], size=(oh, ow), mode="bilinear", align_corners=False, antialias=aa)[0]
            assert (O.resize_bilinear(x, oh, ow, aa) - ref).abs().max().item() <= RESIZE_TOL, (ih, iw, oh, ow, aa)


def parse_plan(host):
    """Header of csrc/resize.cu's plan blob: magic, geometry, taps per column / row, widest tile, rows per CTA and source-row
    span for uint8 / fp32 sources, table offsets."""
    h = np.frombuffer(host[:80].tobytes(), dtype=np.int32)
    in_h, in_w, out_h, out_w, aa, kx, ky, cols_max = (int(v) for v in h[1:9])
    off = [int(v) for v in h[13:19]]

    def arr(o, n, dt):
        return np.frombuffer(host[o:o + 4 * n].tobytes(), dtype=dt)

    return {"geom": (in_h, in_w, out_h, out_w, aa), "kx": kx, "ky": ky, "cols_max": cols_max,
            "tile_rows": (int(h[9]), int(h[10])), "span": (int(h[11]), int(h[12])),
            "sx": arr(off[0], out_w, np.int32), "cx": arr(off[1], out_w, np.int32), "wx": arr(off[2], out_w * kx, np.float32).reshape(out_w, kx),
            "sy": arr(off[3], out_h, np.int32), "cy": arr(off[4], out_h, np.int32), "wy": arr(off[5], out_h * ky, np.float32).reshape(out_h, ky)}


@pytest.mark.parametrize("geom", [(750, 750, 256, 256), (224, 1232, 128, 512), (225, 225, 256, 256), (97, 411, 128, 99),
                                  (256, 256, 256, 256), (1333, 2000, 128, 512), (3000, 3000, 256, 256), (1, 1, 128, 512)])
@pytest.mark.parametrize("antialias", [True, False])
def test_resize_plan_tables_equal_the_oracles(geom, antialias):
    from witw_b200 import ops

    ih, iw, oh, ow = geom
    p = parse_plan(ops.resize_plan_host(ih, iw, oh, ow, antialias))
    assert p["geom"] == (ih, iw, oh, ow, int(antialias))
    sx, cx, wx = O.resize_taps(iw, ow, antialias)
    sy, cy, wy = O.resize_taps(ih, oh, antialias)
    for name, ref in (("sx", sx), ("cx", cx), ("wx", wx), ("sy", sy), ("cy", cy), ("wy", wy)):
        assert np.array_equal(p[name], ref), name
    # every tile's source rectangle fits the shared memory the launch asks for, and the taps stay inside the image
    for t, esz in ((0, 1), (1, 4)):
        rows = p["tile_rows"][t]
        if rows == 0:
            continue
        assert rows in (1, 2, 4, 8, 16, 32)
        pitch = (p["cols_max"] * esz + 15) // 16 * 16 + 16          # a landing row: one chunk more than the widest rectangle
        assert 2 * p["span"][t] * pitch + rows * ((p["cols_max"] + 3) // 4 * 4) * 4 <= 200 * 1024
        for y0 in range(0, oh, rows):
            y1 = min(y0 + rows, oh) - 1
            assert p["sy"][y1] + p["cy"][y1] - p["sy"][y0] <= p["span"][t]
    assert p["tile_rows"][0] > 0
    for x0 in range(0, ow, 64):
        x1 = min(x0 + 64, ow) - 1
        assert p["sx"][x1] + p["cx"][x1] - p["sx"][x0] <= p["cols_max"]
    assert (np.diff(p["sx"]) >= 0).all() and (np.diff(p["sx"] + p["cx"]) >= 0).all()
    assert (np.diff(p["sy"]) >= 0).all() and (np.diff(p["sy"] + p["cy"]) >= 0).all()
    assert (p["sx"] >= 0).all() and (p["sx"] + p["cx"] <= iw).all() and (p["cx"] >= 1).all()
    assert (p["sy"] >= 0).all() and (p["sy"] + p["cy"] <= ih).all() and (p["cy"] >= 1).all()


def test_resize_plan_rejects_what_the_kernel_cannot_do():
    from witw_b200 import _lib, ops

    with pytest.raises(_lib.WitwError):
        ops.resize_plan_host(256, 9000, 256, 256, True)       # 73 taps per output column
    with pytest.raises(ValueError):
        ops.resize_plan_host(0, 10, 256, 256, True)


def test_plan_driven_resample_tracks_oracle():
    """The kernel's data flow replayed in numpy from the plan tables: per (64-column, tile_rows) tile, stage the source
    rectangle, y pass over four-column groups into an fp32 tile, x pass from that tile.  Checks the tile / rectangle
    bookkeeping; the y-first order moves results by an ulp or two of the pixel scale against the oracle's x-first order."""
    from witw_b200 import ops

    gen = torch.Generator().manual_seed(9)
    ih, iw, oh, ow = 131, 277, 70, 150
    x = torch.randint(0, 256, (ih, iw), generator=gen).float().numpy()
    p = parse_plan(ops.resize_plan_host(ih, iw, oh, ow, True))
    rows = p["tile_rows"][0]
    pitch = (p["cols_max"] + 15) // 16 * 16
    out = np.full((oh, ow), np.nan, np.float32)
    for x0 in range(0, ow, 64):
        x1 = min(x0 + 64, ow)
        c_lo, c_hi = p["sx"][x0], p["sx"][x1 - 1] + p["cx"][x1 - 1]
        n_groups = (c_hi - c_lo + 3) // 4
        assert 4 * n_groups <= pitch
        for y0 in range(0, oh, rows):
            y1 = min(y0 + rows, oh)
            r_lo, r_hi = p["sy"][y0], p["sy"][y1 - 1] + p["cy"][y1 - 1]
            raw = np.zeros((r_hi - r_lo, pitch), np.float32)
            raw[:, : c_hi - c_lo] = x[r_lo:r_hi, c_lo:c_hi]
            mid = np.zeros((y1 - y0, pitch), np.float32)
            for y in range(y0, y1):
                acc = np.zeros(4 * n_groups, np.float32)
                for j in range(p["cy"][y]):
                    acc = (acc + raw[p["sy"][y] - r_lo + j, : 4 * n_groups] * p["wy"][y, j]).astype(np.float32)
                mid[y - y0, : 4 * n_groups] = acc
            for c in range(x0, x1):
                t = mid[:, p["sx"][c] - c_lo:]
                acc = t[:, 0] * p["wx"][c, 0]
                for j in range(1, p["cx"][c]):
                    acc = (acc + t[:, j] * p["wx"][c, j]).astype(np.float32)
                out[y0:y1, c] = acc
    assert np.abs(out - O.resize_bilinear(x, oh, ow, True).numpy()).max() <= RESIZE_TOL


def test_resize_plan_cache_is_bounded(monkeypatch):
    """Datasets with many raw image sizes must not grow the plan cache without bound: least recently used plans go first."""
    from witw_b200 import ops

    monkeypatch.setattr(ops, "RESIZE_CACHE_ENTRIES", 3)
    monkeypatch.setattr(ops, "_resize_cache", {})
    dev = torch.device("cpu")                     # the cache logic is device-agnostic; kernels are not involved
    for i in range(5):
        ops._resize_plan(10 + i, 10, 8, 8, True, dev)
    assert [k[0] for k in ops._resize_cache] == [12, 13, 14]
    first = ops._resize_plan(12, 10, 8, 8, True, dev)
    assert [k[0] for k in ops._resize_cache] == [13, 14, 12]
    assert ops._resize_plan(12, 10, 8, 8, True, dev) is first


def test_random_geometries_plan_oracle_torch_agree():
    """Seeded random sizes (integer and awkward ratios, 1-pixel axes, up- and downsampling): the library's host tables equal
    the oracle's bit for bit, and the oracle's resize tracks ATen's."""
    import torch.nn.functional as F

    from witw_b200 import ops

    rng = np.random.default_rng(20261017)
    sizes = [(1, 7, 3, 64), (12, 9, 1, 1), (300, 100, 100, 300), (96, 512, 32, 128), (2, 2, 255, 257), (511, 513, 128, 512)]
    while len(sizes) < 36:
        ih, iw = (int(v) for v in rng.integers(1, 400, 2))
        oh, ow = (int(v) for v in rng.integers(1, 300, 2))
        if max(ih / oh, iw / ow) < 15:          # the kernel's tap limit (32 per output column) is tested elsewhere
            sizes.append((ih, iw, oh, ow))
    gen = torch.Generator().manual_seed(17)
    for (ih, iw, oh, ow) in sizes:
        x = torch.randint(0, 256, (1, ih, iw), generator=gen).float()
        for aa in (True, False):
            p = parse_plan(ops.resize_plan_host(ih, iw, oh, ow, aa))
            sx, cx, wx = O.resize_taps(iw, ow, aa)
            sy, cy, wy = O.resize_taps(ih, oh, aa)
            for name, ref in (("sx", sx), ("cx", cx), ("wx", wx), ("sy", sy), ("cy", cy), ("wy", wy)):
                assert np.array_equal(p[name], ref), (ih, iw, oh, ow, aa, name)
            ref = F.interpolate(x[None], size=(oh, ow), mode="bilinear", align_corners=False, antialias=aa)[0]
            assert (O.resize_bilinear(x, oh, ow, aa) - ref).abs().max().item() <= RESIZE_TOL, (ih, iw, oh, ow, aa)
