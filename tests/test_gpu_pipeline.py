"""BASELINE configs[0] end to end on the GPU: 256 synthetic ground / aerial pairs, 360-degree panoramas,
a random-init VGG16-based encoder, recall@1 / @1 %.  The encoder (cuDNN convolutions) is outside the hot path and is
rebuilt here from torchvision as FOV_DSM builds it (model/cvig_fov.py:248-294: vgg16.features[:23] + three convolutions,
dropout inactive in eval mode); everything on the path -- polar transform, orientation-searched distance, rank loop --
runs through witw_b200 and is compared with the oracle's rank loop on the very same embeddings."""
import numpy as np
import pytest
import torch

from oracle import witw_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def W():
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    import witw_b200
    from witw_b200 import _lib

    _lib.call("witw_device_check")
    return witw_b200


def make_encoder():
    import torchvision

    torch.manual_seed(1234)
    feats = torchvision.models.vgg16(weights=None).features[:23]
    tail = [torch.nn.Conv2d(512, 256, 3, (2, 1), padding=1), torch.nn.ReLU(inplace=True),
            torch.nn.Conv2d(256, 64, 3, (2, 1), padding=1), torch.nn.ReLU(inplace=True), torch.nn.Conv2d(64, 16, 3, padding=1)]
    for m in tail:
        if isinstance(m, torch.nn.Conv2d):
            torch.nn.init.xavier_uniform_(m.weight)
            torch.nn.init.zeros_(m.bias)
    return torch.nn.Sequential(*feats, *tail).eval()


def test_config0_pipeline_256_pairs(W):
    n = 256
    gen = torch.Generator().manual_seed(5)
    # aerial tiles after ImageNormalization: smooth random fields (white noise gives a random-init VGG near-constant features)
    coarse = torch.randn(n, 3, 16, 16, generator=gen) * 2
    tiles = torch.nn.functional.interpolate(coarse, size=(256, 256), mode="bilinear", align_corners=False)
    shifts = torch.randint(0, 64, (n,), generator=gen)
    enc = make_encoder().cuda()
    with torch.no_grad():
        polar = W.PolarTransform()({"overhead": tiles.cuda()})["polar"]      # [n,3,128,512]
        # the ground panorama of pair i: its aerial polar image seen from another heading, plus sensor noise
        ground = torch.stack([torch.roll(polar[i], -8 * int(shifts[i]), dims=2) for i in range(n)])
        ground = ground + 1.5 * torch.randn(ground.shape, generator=gen).cuda()
        ov = torch.cat([enc(polar[i:i + 32]) for i in range(0, n, 32)])
        su = torch.cat([enc(ground[i:i + 32]) for i in range(0, n, 32)])
    assert tuple(ov.shape) == (n, 16, 4, 64) and tuple(su.shape) == (n, 16, 4, 64)
    # polar transform of the batch against the oracle on a few tiles
    for i in (0, 100, 255):
        assert (polar[i].cpu() - O.polar_transform(tiles[i])).abs().max().item() <= 4e-6

    ov_c, su_c = ov.cpu(), su.cpu()
    want = O.rank_loop(ov_c, su_c)                                           # cvig_fov.py:543-552 on the same embeddings
    _, ref = O.match(ov_c, su_c)
    thr = torch.diagonal(ref).unsqueeze(0)
    tie = ((ref - thr).abs() <= 3e-6 * ref.abs().clamp_min(1.0)).sum(0).numpy() - 1
    ref_recall = O.recall_from_ranks(want)
    for path in ("fp32", "tc"):
        ranks = W.evaluate_ranks(ov, su, path=path).cpu().numpy()
        assert np.all(np.abs(ranks - want) <= tie), path
        got = W.recall_from_ranks(ranks)
        if int(tie.sum()) == 0:
            assert got == ref_recall
        assert abs(got["top_one"] - ref_recall["top_one"]) <= 100.0 * tie.sum() / n
    # the rebound names run the reference's own loop body (correlation -> crop_overhead -> l2_distance per query)
    one = su[7:8]
    ori = W.correlation(ov, one)
    d = torch.squeeze(W.l2_distance(W.crop_overhead(ov, ori, one.shape[3]), one))
    assert int(torch.sum(torch.le(d, d[7]))) - want[7] in range(-int(tie[7]), int(tie[7]) + 1)
    # recall is not degenerate on this data: the planted heading is found for most pairs
    ori_all, _ = W.match(ov, su, path="fp32")
    assert (torch.diagonal(ori_all).cpu() == shifts).float().mean().item() >= 0.5
    assert ref_recall["top_one"] >= 20.0 and len(set(want.tolist())) > 5


def test_heatmap_sweep_pipeline_2048_tiles(W):
    """tools/heatmap/heatmap.py:113-187 as one streamed pipeline: 2 048 raw uint8 tiles in pinned host batches -> fused
    normalise + polar kernel -> overhead encoder -> gallery operand (no torch.cat) -> one photo scored against all tiles in
    fp32.  With the bit-exact polar kernel the encoder sees the oracle's own polar images, so orientation, dissimilarity and
    score must be the reference chain's (heatmap.py:171-177 through oracle.heatmap_scores) to fp32 round-off; with the fast
    polar kernel (4e-6 from the oracle's images) the picture stays the same."""
    n, bs = 2048, 64
    gen = torch.Generator().manual_seed(9)
    coarse = torch.rand(n, 3, 16, 16, generator=gen) * 255
    tiles = torch.nn.functional.interpolate(coarse, size=(256, 256), mode="bilinear", align_corners=False).round().clamp(0, 255).to(torch.uint8)
    photo = torch.randint(0, 256, (3, 300, 411), generator=gen, dtype=torch.uint8)
    enc = make_encoder().cuda()
    host_batches = [tiles[i: i + bs].pin_memory() for i in range(0, n, bs)]
    deg, dis, score = W.heatmap_sweep(host_batches, n, photo, enc, enc, fov=90, exact_polar=True)
    assert tuple(deg.shape) == (n,) and tuple(dis.shape) == (n,) and tuple(score.shape) == (n,)
    # the oracle's chain on the host: ImageNormalization -> PolarTransform per tile (cvig_fov.py:147, 186-209), the same
    # encoder on the same batches, then heatmap.py:171-177
    with torch.no_grad():
        polar = torch.stack([O.normalized_polar(t) for t in tiles[:512]])
        ov_ref = torch.cat([enc(polar[i: i + bs].cuda()) for i in range(0, 512, bs)]).cpu()
        surface = W.resize_normalize(photo.cuda().unsqueeze(0), 128, 128, True, W.ops.IMG_MEAN, W.ops.IMG_STD)
        su = enc(surface).cpu()
    rdeg, rdis, rscore = O.heatmap_scores(ov_ref, su)
    from parity_helpers import assert_orientation_is_the_references
    ori = ((deg[:512].cpu() + 180) * 64 / 360).round().long().view(-1, 1)
    same = assert_orientation_is_the_references(ov_ref, su, ori, O.correlation(ov_ref, su)).view(-1)
    assert same.float().mean().item() >= 0.99
    assert (dis[:512].cpu() - rdis)[same].abs().max().item() <= 5e-6
    assert ((score[:512].cpu() - rscore)[same].abs() <= 1e-4 * rscore[same].abs()).all()
    # default (fast polar kernel): same picture
    deg2, dis2, score2 = W.heatmap_sweep(iter(host_batches), n, photo, enc, enc, fov=90)
    agree = deg2 == deg
    assert agree.float().mean().item() >= 0.98 and (dis2 - dis)[agree].abs().max().item() <= 2e-3
    with pytest.raises(ValueError):
        W.heatmap_sweep(host_batches[:3], n, photo, enc, enc, fov=90)


def test_streamed_polar_five_channel_tiles(W):
    """BASELINE configs[4], first half, at reduced count: 5-channel uint8 tiles stream from pinned host memory through the
    fused normalise + polar kernel batch by batch; every batch equals the oracle chain (cvig_semantic.py:163-176 divides only
    the image channels by 255)."""
    gen = torch.Generator().manual_seed(4)
    tiles = torch.randint(0, 256, (96, 5, 256, 256), generator=gen, dtype=torch.uint8)
    mean, std, div = (0.485, 0.456, 0.406, 0.5, 0.5), (0.229, 0.224, 0.225, 0.25, 0.25), (255.0, 255.0, 255.0, 1.0, 1.0)
    batches = [tiles[i: i + 32].pin_memory() for i in range(0, 96, 32)]
    outs = [p.cpu() for p in W.streamed_polar(batches, mean, std, div)]
    assert len(outs) == 3 and tuple(outs[0].shape) == (32, 5, 128, 512)
    for idx in (0, 40, 95):
        x = tiles[idx].float() / torch.tensor(div).view(5, 1, 1)
        ref = O.polar_transform((x - torch.tensor(mean).view(5, 1, 1)) / torch.tensor(std).view(5, 1, 1))
        got = outs[idx // 32][idx % 32]
        assert (got - ref).abs().max().item() <= 2e-4 * max(1.0, float(ref.abs().max()))
