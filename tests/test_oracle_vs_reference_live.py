"""The oracle against the UNMODIFIED reference, live, on seeded random inputs -- beyond the frozen golden vectors.
Runs only where the reference tree is mounted (the build container); the GPU box has no /root/reference and skips it."""
import numpy as np
import pytest
import torch

from oracle import ref_loader
from oracle import witw_oracle as O

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not mounted")

torch.set_num_threads(1)


@pytest.fixture(scope="module")
def cvig():
    return ref_loader.load("cvig_fov")


def test_polar_and_bilinear_bit_exact(cvig):
    gen = torch.Generator().manual_seed(101)
    tile = torch.randn(5, 256, 256, generator=gen) * 3
    assert torch.equal(O.polar_transform(tile), cvig.PolarTransform()({"overhead": tile.clone()})["polar"])
    for _ in range(4):
        h, w = (int(v) for v in torch.randint(2, 40, (2,), generator=gen))
        im = torch.randn(2, h, w, generator=gen)
        x = (torch.rand(6, 11, generator=gen, dtype=torch.float64) * (w + 6) - 3).numpy()
        y = (torch.rand(6, 11, generator=gen, dtype=torch.float64) * (h + 6) - 3).numpy()
        assert torch.equal(O.bilinear_interpolate(im, x, y), cvig.bilinear_interpolate(im, x, y))


def test_match_chain_random_shapes(cvig):
    gen = torch.Generator().manual_seed(102)
    for trial in range(8):
        g, q = (int(v) for v in torch.randint(1, 12, (2,), generator=gen))
        c, h = (int(v) for v in torch.randint(1, 6, (2,), generator=gen))
        w = int(torch.randint(4, 70, (1,), generator=gen))
        sw = int(torch.randint(1, w + 1, (1,), generator=gen))
        ov = torch.randn(g, c, h, w, generator=gen)
        su = torch.randn(q, c, h, sw, generator=gen)
        if trial % 3 == 0:                       # exact ties: a periodic gallery item
            ov[0] = ov[0][..., :1].expand(-1, -1, w).clone()
        ori_ref = cvig.correlation(ov, su)
        crop_ref = cvig.crop_overhead(ov, ori_ref, sw)
        d_ref = cvig.l2_distance(crop_ref, su)
        assert torch.equal(O.correlation(ov, su), ori_ref)
        assert torch.equal(O.crop_overhead(ov, ori_ref, sw), crop_ref)
        d = O.l2_distance(crop_ref, su)
        assert torch.allclose(d, d_ref, rtol=0, atol=2e-6, equal_nan=True)


def test_rank_loop_and_recall(cvig):
    ov, su, _ = O.synth_features(20, 20, fov=180, noise=20.0, seed=103)
    count = su.size(0)
    ranks = np.zeros([count], dtype=int)
    for idx in range(count):                     # cvig_fov.py:543-552, calling the reference's functions
        one = torch.unsqueeze(su[idx, :], 0)
        ori = cvig.correlation(ov, one)
        d = torch.squeeze(cvig.l2_distance(cvig.crop_overhead(ov, ori, one.shape[3]), one))
        ranks[idx] = torch.sum(torch.le(d, d[idx])).item()
    assert len(set(ranks.tolist())) > 3
    assert np.array_equal(O.rank_loop(ov, su), ranks)


def test_transform_chain_random_sizes(cvig):
    """Resize -> ImageNormalization -> PolarTransform with the container's torchvision (antialiased), inexact scale factors
    included; both dataset kinds."""
    gen = torch.Generator().manual_seed(104)
    for trial in range(4):
        oh, ow = (int(v) for v in torch.randint(200, 420, (2,), generator=gen))
        sh, sw_px = int(torch.randint(60, 200, (1,), generator=gen)), int(torch.randint(300, 700, (1,), generator=gen))
        ov8 = torch.randint(0, 256, (3, oh, ow), generator=gen, dtype=torch.uint8)
        su8 = torch.randint(0, 256, (3, sh, sw_px), generator=gen, dtype=torch.uint8)
        fov = (360, 90, 70, 180)[trial]
        if trial % 2 == 0:
            torch.manual_seed(trial)
            state = torch.get_rng_state()
            start = int(torch.randint(0, 512, ()))
            torch.set_rng_state(state)
            d = cvig.Resize("cvusa", fov=fov, random_orientation=True)({"surface": su8.float(), "overhead": ov8.float()})
            su, ov = O.resize_pair(su8.float(), ov8.float(), fov=fov, panorama=True, start=start)
        else:
            d = cvig.Resize("witw", fov=fov)({"surface": su8.float(), "overhead": ov8.float()})
            su, ov = O.resize_pair(su8.float(), ov8.float(), fov=fov, panorama=False)
        assert su.shape == d["surface"].shape and ov.shape == d["overhead"].shape
        assert (su - d["surface"]).abs().max().item() <= 6.2e-5
        assert (ov - d["overhead"]).abs().max().item() <= 6.2e-5
        d = cvig.PolarTransform()(cvig.ImageNormalization()(d))
        assert (O.polar_transform(O.image_normalization(ov)) - d["polar"]).abs().max().item() <= 2e-6


def test_baseline_rank_loop_live():
    base = ref_loader.load("cvig_baseline")
    assert base is not None
    gen = torch.Generator().manual_seed(105)
    ov = torch.randn(30, 64, generator=gen)
    su = ov + 3.0 * torch.randn(30, 64, generator=gen)
    ranks = np.zeros([30], dtype=int)
    for idx in range(30):                        # cvig_baseline.py:453-460
        one = torch.unsqueeze(su[idx, :], 0)
        dd = torch.pow(torch.sum(torch.pow(ov - one, 2), dim=1), 0.5)
        ranks[idx] = torch.sum(torch.le(dd, dd[idx])).item()
    assert np.array_equal(O.baseline_rank_loop(ov, su), ranks)
