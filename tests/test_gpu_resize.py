"""GPU parity of the Resize / ImageNormalization front end (csrc/resize.cu) through the C ABI: against the
oracle on seeded inputs, against the golden vectors frozen from the unmodified reference's transform chain, and -- at the
bench's batch sizes -- through properties (identity geometry, column window = roll of the full resize)."""
import numpy as np
import pytest
import torch

from oracle import witw_oracle as O

pytestmark = pytest.mark.gpu

# fused multiply-adds on the GPU against separately rounded products in the oracle / ATen: a few ulp of the 0..255 scale
RESIZE_TOL = 1.3e-4


@pytest.fixture(scope="module")
def W():
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    import witw_b200
    from witw_b200 import _lib

    _lib.call("witw_device_check")
    return witw_b200


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def redraw_inputs(g):
    gen = torch.Generator().manual_seed(int(g["big_seed"]))
    ov8 = torch.randint(0, 256, (3, 301, 283), generator=gen, dtype=torch.uint8)
    su8 = torch.randint(0, 256, (3, 97, 411), generator=gen, dtype=torch.uint8)
    big8 = torch.randint(0, 256, (3, 750, 750), generator=gen, dtype=torch.uint8)
    assert int(big8.long().sum()) == int(g["big_sum"])
    return ov8, su8, big8


@pytest.mark.parametrize("geom", [(750, 750, 256, 256), (97, 411, 128, 512), (225, 225, 256, 256), (300, 400, 128, 99),
                                  (256, 256, 256, 256), (1333, 1000, 128, 512), (37, 53, 40, 70), (1, 1, 5, 3)])
@pytest.mark.parametrize("antialias", [True, False])
def test_resize_vs_oracle(W, geom, antialias):
    ih, iw, oh, ow = geom
    gen = torch.Generator().manual_seed(ih * 7 + ow)
    img8 = torch.randint(0, 256, (2, 3, ih, iw), generator=gen, dtype=torch.uint8)
    ref = O.resize_bilinear(img8.float(), oh, ow, antialias)
    out8 = W.resize_normalize(img8.cuda(), oh, ow, antialias).cpu()
    outf = W.resize_normalize(img8.float().cuda(), oh, ow, antialias).cpu()
    assert out8.shape == ref.shape and out8.dtype == torch.float32
    assert torch.equal(out8, outf)                      # uint8 pixels convert exactly
    assert (out8 - ref).abs().max().item() <= RESIZE_TOL
    # non-integer fp32 pixels
    imgf = torch.randn(1, ih, iw, generator=gen) * 40 + 100
    ref = O.resize_bilinear(imgf, oh, ow, antialias)
    assert (W.resize_normalize(imgf.cuda(), oh, ow, antialias).cpu() - ref).abs().max().item() <= RESIZE_TOL * 2


def test_transform_chain_vs_reference_golden(W, golden):
    g = golden("resize")
    ov8, su8, big8 = redraw_inputs(g)
    start = int(g["pano_start"])
    # Resize drop-in draws the start column from the same generator state as the reference
    torch.manual_seed(5)
    d = W.Resize("cvusa", fov=90, random_orientation=True)({"surface": su8.cuda(), "overhead": ov8.cuda(), "idx": 3})
    assert set(d) == {"surface", "overhead", "idx"}
    assert (d["surface"].cpu() - t(g["pano_surface"])).abs().max().item() <= RESIZE_TOL
    assert (d["overhead"].cpu()[:, ::3, ::5] - t(g["pano_overhead_sub"])).abs().max().item() <= RESIZE_TOL
    d = W.PolarTransform()(W.ImageNormalization()(d))
    assert (d["surface"].cpu()[:, ::3, ::3] - t(g["pano_surface_norm_sub"])).abs().max().item() <= 4e-6
    assert (d["polar"].cpu()[:, ::3, ::7] - t(g["pano_polar_sub"])).abs().max().item() <= 8e-6
    # the fused form: three kernels from raw images to {'surface', 'overhead', 'polar'}
    f = W.prepare_pair(su8.cuda(), ov8.cuda(), fov=90, panorama=True, start=start)
    assert (f["surface"] - d["surface"]).abs().max().item() <= 1e-6
    assert (f["polar"] - d["polar"]).abs().max().item() <= 4e-6
    # a photo that is not a panorama, and the 2.9x downsample of a CVUSA-sized aerial image
    d = W.Resize("witw", fov=70)({"surface": su8.cuda(), "overhead": big8.cuda()})
    assert tuple(d["surface"].shape) == (3, 128, 99)
    assert (d["surface"].cpu()[:, ::2, ::3] - t(g["witw_surface_sub"])).abs().max().item() <= RESIZE_TOL
    assert (d["overhead"].cpu()[:, ::5, ::3] - t(g["witw_overhead_sub"])).abs().max().item() <= RESIZE_TOL
    # the pinned torchvision's resize (no antialiasing)
    ov = W.resize_normalize(big8.cuda(), 256, 256, antialias=False).cpu()
    su = W.resize_normalize(su8.cuda(), 128, 512, antialias=False).cpu()
    assert (ov[:, ::5, ::3] - t(g["noaa_overhead_sub"])).abs().max().item() <= RESIZE_TOL
    assert (su[:, ::3, ::5] - t(g["noaa_surface_sub"])).abs().max().item() <= RESIZE_TOL


def test_image_normalization_dropin_bit_exact(W):
    gen = torch.Generator().manual_seed(12)
    su8 = torch.randint(0, 256, (3, 128, 200), generator=gen, dtype=torch.uint8)
    ov8 = torch.randint(0, 256, (4, 3, 256, 256), generator=gen, dtype=torch.uint8)     # a batch
    d = W.ImageNormalization()({"surface": su8.cuda(), "overhead": ov8.cuda()})
    assert torch.equal(d["surface"].cpu(), O.image_normalization(su8))
    assert torch.equal(d["overhead"].cpu(), torch.stack([O.image_normalization(x) for x in ov8]))
    # float pixels, as the reference's datasets deliver them (cvig_fov.py:90-91)
    d = W.ImageNormalization()({"surface": su8.float().cuda(), "overhead": ov8[0].float().cuda()})
    assert torch.equal(d["surface"].cpu(), O.image_normalization(su8))
    # a plane that is not a whole number of four-pixel groups takes the general kernel: same bits
    odd = torch.randint(0, 256, (2, 3, 5, 7), generator=gen, dtype=torch.uint8)
    d = W.ImageNormalization()({"surface": odd.cuda(), "overhead": odd.float().cuda()})
    ref_odd = torch.stack([O.image_normalization(x) for x in odd])
    assert torch.equal(d["surface"].cpu(), ref_odd) and torch.equal(d["overhead"].cpu(), ref_odd)
    # cvig_semantic.py:163-176: five channels, only the first three divided by 255
    mean, std = (0.485, 0.456, 0.406, 0.45, 0.45), (0.229, 0.224, 0.225, 0.22, 0.22)
    x = torch.rand(5, 128, 64, generator=gen) * torch.tensor([255, 255, 255, 1, 1.0]).view(5, 1, 1)
    d = W.ImageNormalization(mean, std, divisor=(255, 255, 255, 1, 1))({"surface": x.cuda(), "overhead": x.cuda()})
    ref = x.clone()
    ref[:3] /= 255.
    ref = (ref - torch.tensor(mean).view(5, 1, 1)) / torch.tensor(std).view(5, 1, 1)
    assert torch.equal(d["overhead"].cpu(), ref)


def test_resize_window_batches_and_host_tensors(W):
    gen = torch.Generator().manual_seed(13)
    pano = torch.randint(0, 256, (5, 3, 60, 700), generator=gen, dtype=torch.uint8).cuda()
    full = W.resize_normalize(pano, 128, 512, mean=O.IMG_MEAN, std=O.IMG_STD)
    for start, count in ((0, 512), (500, 128), (511, 512), (17, 1), (256, 256)):
        win = W.resize_normalize(pano, 128, 512, mean=O.IMG_MEAN, std=O.IMG_STD, col_start=start, col_count=count)
        cols = (start + torch.arange(count)) % 512
        assert torch.equal(win, full[..., cols.cuda()]), (start, count)
    # more planes than the grid's z extent: the kernel strides over planes
    many = torch.randint(0, 256, (40000, 1, 6, 9), generator=gen, dtype=torch.uint8)
    out = W.resize_normalize(many.cuda(), 12, 20).cpu()
    for n in (0, 32767, 32768, 39999):
        assert (out[n] - O.resize_bilinear(many[n].float(), 12, 20)).abs().max().item() <= RESIZE_TOL
    # CPU samples (the reference's transforms run on CPU tensors): computed on the GPU, returned on the CPU
    d = W.Resize("witw", fov=360)({"surface": pano[0].cpu(), "overhead": pano[1].cpu()})
    assert not d["surface"].is_cuda and tuple(d["overhead"].shape) == (3, 256, 256)
    assert torch.equal(d["surface"], W.resize_normalize(pano[0], 128, 512).cpu())


def test_resize_errors_and_empty(W):
    x = torch.zeros(3, 10, 10, dtype=torch.int32, device="cuda")
    with pytest.raises(TypeError):
        W.resize_normalize(x, 5, 5)
    with pytest.raises(RuntimeError):
        W.resize_normalize(torch.zeros(3, 10, 10), 5, 5)                       # CPU tensor: no fallback
    with pytest.raises(ValueError):
        W.resize_normalize(torch.zeros(2, 10, 10, device="cuda"), 5, 5, mean=O.IMG_MEAN, std=O.IMG_STD)
    with pytest.raises(W.WitwError):
        W.resize_normalize(torch.zeros(3, 10, 10, device="cuda"), 5, 5, col_start=5)
    assert tuple(W.resize_normalize(torch.zeros(0, 3, 10, 10, device="cuda"), 5, 7).shape) == (0, 3, 5, 7)


def test_prepare_pair_at_batch_scale(W):
    """256 CVUSA-sized pairs (BASELINE configs[0]'s count): the identity-geometry and linearity properties that do not
    need the oracle at this size, plus spot checks against it."""
    gen = torch.Generator().manual_seed(14)
    ov8 = torch.randint(0, 256, (256, 3, 750, 750), generator=gen, dtype=torch.uint8)
    su8 = torch.randint(0, 256, (256, 3, 224, 1232), generator=gen, dtype=torch.uint8)
    out = W.prepare_pair(su8.cuda(), ov8.cuda(), fov=360, panorama=True, start=100)
    assert tuple(out["polar"].shape) == (256, 3, 128, 512) and tuple(out["surface"].shape) == (256, 3, 128, 512)
    for n in (0, 131, 255):
        su, ov = O.resize_pair(su8[n].float(), ov8[n].float(), fov=360, panorama=True, start=100)
        assert (out["surface"][n].cpu() - O.image_normalization(su)).abs().max().item() <= 4e-6
        assert (out["polar"][n].cpu() - O.polar_transform(O.image_normalization(ov))).abs().max().item() <= 8e-6
    # constant images stay constant (weights sum to one) up to rounding
    flat = torch.full((2, 3, 500, 640), 200, dtype=torch.uint8, device="cuda")
    r = W.resize_normalize(flat, 256, 256)
    assert (r - 200.0).abs().max().item() <= 5e-5
