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
    h = np.frombuffer(host[:68].tobytes(), dtype=np.int32)
    in_h, in_w, out_h, out_w, aa, kx, ky, tile_rows, span = (int(v) for v in h[1:10])
    off = [int(v) for v in h[10:16]]

    def arr(o, n, dt):
        return np.frombuffer(host[o:o + 4 * n].tobytes(), dtype=dt)

    return {"geom": (in_h, in_w, out_h, out_w, aa), "kx": kx, "ky": ky, "tile_rows": tile_rows, "span": span,
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
    # every row tile's source rows fit the kernel's shared-memory tile, and the taps stay inside the image
    assert 1 <= p["tile_rows"] <= 32 and p["span"] <= 176
    for y0 in range(0, oh, p["tile_rows"]):
        y1 = min(y0 + p["tile_rows"], oh) - 1
        assert p["sy"][y1] + p["cy"][y1] - p["sy"][y0] <= p["span"]
    assert (p["sx"] >= 0).all() and (p["sx"] + p["cx"] <= iw).all() and (p["cx"] >= 1).all()
    assert (p["sy"] >= 0).all() and (p["sy"] + p["cy"] <= ih).all() and (p["cy"] >= 1).all()


def test_resize_plan_rejects_what_the_kernel_cannot_do():
    from witw_b200 import _lib, ops

    with pytest.raises(_lib.WitwError):
        ops.resize_plan_host(256, 9000, 256, 256, True)       # 73 taps per output column
    with pytest.raises(ValueError):
        ops.resize_plan_host(0, 10, 256, 256, True)


def test_plan_driven_resample_equals_oracle():
    """The kernel's data flow (x pass over the rows a tile needs into an fp32 tile, then the y pass) replayed in numpy from
    the plan tables reproduces the oracle's resize exactly -- the tile / span bookkeeping is what is being checked."""
    from witw_b200 import ops

    gen = torch.Generator().manual_seed(9)
    ih, iw, oh, ow = 131, 277, 64, 100
    x = torch.randint(0, 256, (ih, iw), generator=gen).float().numpy()
    p = parse_plan(ops.resize_plan_host(ih, iw, oh, ow, True))
    out = np.zeros((oh, ow), np.float32)
    for y0 in range(0, oh, p["tile_rows"]):
        y1 = min(y0 + p["tile_rows"], oh)
        r_lo, r_hi = p["sy"][y0], p["sy"][y1 - 1] + p["cy"][y1 - 1]
        tile = np.zeros((r_hi - r_lo, ow), np.float32)
        for c in range(ow):
            acc = x[r_lo:r_hi, p["sx"][c]] * p["wx"][c, 0]
            for j in range(1, p["cx"][c]):
                acc = (acc + x[r_lo:r_hi, p["sx"][c] + j] * p["wx"][c, j]).astype(np.float32)
            tile[:, c] = acc
        for y in range(y0, y1):
            acc = tile[p["sy"][y] - r_lo] * p["wy"][y, 0]
            for j in range(1, p["cy"][y]):
                acc = (acc + tile[p["sy"][y] - r_lo + j] * p["wy"][y, j]).astype(np.float32)
            out[y] = acc
    assert np.array_equal(out, O.resize_bilinear(x, oh, ow, True).numpy())
