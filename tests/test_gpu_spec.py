"""GPU parity tests of the spectral tensor-core sweep (csrc/match_spec.cu: per-frequency products on tcgen05 with fp16
spectra of the norm-scaled features, inverse FFT + argmax + distance + rank count + top-k in the epilogue) and of its
fp32 finish (csrc/finish.cu) against the oracle."""
import numpy as np
import pytest
import torch

from oracle import witw_oracle as O
from parity_helpers import assert_orientation_is_the_references, check_exact_results, hard_features_cuda, reference_columns

pytestmark = pytest.mark.gpu

KAPPA = 16.0     # csrc/sweep_common.cuh: kSpecKappa


@pytest.fixture(scope="module")
def W():
    assert torch.cuda.is_available()
    import witw_b200

    return witw_b200


@pytest.fixture(autouse=True)
def _spectral_sweep(W):
    W.ops.TC_IMPL = "spectral"          # also for widths where "auto" would pick the dense contraction
    yield
    W.ops.TC_IMPL = "auto"
    W.ops.DEFERRAL_CAP = None


def spec_model(ov, su):
    """What the raw sweep computes, in float64: spectra of the fp32 rows scaled by 16 / norm and rounded to fp16,
    per-frequency products summed over the feature rows, inverse real FFT; norms of the fp32 inputs.
    Returns (normalised correlation [G,Q,64], orientation, distance)."""
    G, Q, sw = ov.shape[0], su.shape[0], su.shape[3]
    pad = torch.zeros(Q, su.shape[1], su.shape[2], 64, dtype=torch.float64)
    pad[..., :sw] = su.double()
    gn = ov.double().reshape(G, -1).norm(dim=1)
    qn = su.double().reshape(Q, -1).norm(dim=1)
    So = torch.fft.rfft(ov.double(), dim=3).reshape(G, -1, 33) * (KAPPA / gn).view(G, 1, 1)
    Sq = torch.fft.rfft(pad, dim=3).reshape(Q, -1, 33) * (KAPPA / qn).view(Q, 1, 1)

    def rnd(z):
        return torch.complex(z.real.float().half().double(), z.imag.float().half().double())

    So, Sq = rnd(So), rnd(Sq)
    P = torch.einsum("grf,qrf->gqf", So, Sq.conj())
    corr = torch.fft.irfft(P, n=64, dim=2) / KAPPA ** 2             # corr / (||ov|| ||su||)
    ori = torch.argmax(corr, -1)
    shift = (torch.arange(64).view(64, 1) + torch.arange(sw).view(1, sw)) % 64
    cn = torch.sqrt((ov.double() ** 2).sum((1, 2))[:, shift].sum(-1))
    best = torch.gather(corr, 2, ori.unsqueeze(-1)).squeeze(-1)
    ratio = gn.view(G, 1) / torch.gather(cn, 1, ori.reshape(G, -1)).reshape(ori.shape)
    return corr, ori, 2 - 2 * best * ratio


@pytest.mark.parametrize("fov,G,Q", [(360, 203, 300), (180, 36, 16), (90, 130, 70), (70, 64, 257), (6, 20, 9), (360, 8, 128), (360, 1, 1)])
def test_spec_match_vs_oracle(W, fov, G, Q):
    ov, su, _ = O.synth_features(G, Q, fov=fov, noise=1.0, seed=fov + G)
    # tier 1: the raw fp16 sweep against the float64 model of the same arithmetic: fp32 transform / accumulation round-off only
    ori, dist = W.match(ov.cuda(), su.cuda(), path="tc16")
    ori, dist = ori.cpu(), dist.cpu()
    assert tuple(ori.shape) == (G, Q) and ori.dtype == torch.int64
    corr, m_ori, m_dist = spec_model(ov, su)
    diff = ori != m_ori
    a = torch.gather(corr, 2, ori.unsqueeze(-1)).squeeze(-1)
    b = torch.gather(corr, 2, m_ori.unsqueeze(-1)).squeeze(-1)
    assert bool((((a - b).abs() <= 4e-6) | ~diff).all())            # differs only between shifts the model itself ties
    assert (dist.double() - m_dist)[~diff].abs().max().item() <= 2e-5
    # tier 2: the finished sweep against the fp32 reference chain: the reference's orientation, distances within the
    # north star's 1e-3 relative at every field of view, on every pair
    ref_ori, ref = O.match(ov, su)
    ori, dist = W.match(ov.cuda(), su.cuda(), path="tc")
    ori, dist = ori.cpu(), dist.cpu()
    same = assert_orientation_is_the_references(ov, su, ori, ref_ori)
    rel = ((dist - ref).abs() / ref.abs())[same]
    assert rel.max().item() <= 1e-3, rel.max().item()
    assert (dist - ref)[same].abs().max().item() <= (6e-4 if su.shape[3] >= 8 else 1.5e-3)
    if G * Q >= 1000:
        assert same.float().mean().item() >= 0.995


@pytest.mark.parametrize("fov,n,noise", [(360, 300, 25.0), (180, 260, 12.0), (90, 260, 10.0)])
def test_spec_raw_sweep_ranks_and_topk(W, fov, n, noise):
    """exact=False: the raw fp16 decisions.  Ranks differ from the reference only by pairs within 3e-4 of the threshold;
    the fused top-k is a sort of the kernel's own distance matrix."""
    ov, su, _ = O.synth_features(n, n, fov=fov, noise=noise, seed=17)
    ranks, td, ti = W.evaluate_ranks(ov.cuda(), su.cuda(), path="tc", topk=5, exact=False)
    ranks = ranks.cpu().numpy()
    ref_ori, ref = O.match(ov, su)
    want = (ref <= torch.diagonal(ref).unsqueeze(0)).sum(0).numpy()
    assert len(set(want.tolist())) > 5
    band = ((ref - torch.diagonal(ref).unsqueeze(0)).abs() <= (3e-4 if fov == 360 else 1.2e-2)).sum(0).numpy() - 1
    assert np.all(np.abs(ranks - want) <= band)
    assert np.mean(ranks == want) >= 0.9
    _, dist = W.match(ov.cuda(), su.cuda(), path="tc16")
    sd = torch.sort(dist.t().cpu(), dim=1, stable=True)
    assert torch.equal(td.cpu(), sd.values[:, :5]) and torch.equal(ti.cpu().long(), sd.indices[:, :5])


@pytest.mark.parametrize("fov,n,noise", [(360, 300, 25.0), (180, 280, 12.0), (90, 260, 10.0), (45, 300, 6.0)])
def test_spec_exact_finish_matches_fp32_reference(W, fov, n, noise):
    """exact=True on the spectral sweep: ranks and top-k are the fp32 reference's (cvig_fov.py:547-552)."""
    ov, su, _ = O.synth_features(n, n, fov=fov, noise=noise, seed=23)
    ranks, td, ti = W.evaluate_ranks(ov.cuda(), su.cuda(), path="tc", topk=5)
    ref_ori, ref = O.match(ov, su)
    want = check_exact_results(ref, ranks.cpu().numpy(), td, ti, 5)
    assert len(set(want.tolist())) > 5
    stats = W.ops.evaluate_ranks_prepared.last_stats
    deferred = stats["deferred"].cpu()
    assert int(deferred.max()) <= stats["list_cap"] and int(deferred.sum()) > 0 and stats["flagged"] == 0
    # the deferral is sparse: the error bound of the fp16 operands is ~1e-4 of a distance
    assert int(deferred.sum()) <= 0.03 * n * n


@pytest.mark.parametrize("fov", [360, 90])
def test_spec_exact_finish_survives_list_overflow(W, fov):
    """Deferral lists that are too small (forced here: 2 entries per query) must not cost exactness: the overflowing
    queries are re-done entirely in fp32 (the re-check list used to overflow silently)."""
    n = 400
    ov, su, _ = O.synth_features(n, n, fov=fov, noise=25.0 if fov == 360 else 10.0, seed=31)
    W.ops.DEFERRAL_CAP = 2
    ranks, td, ti = W.evaluate_ranks(ov.cuda(), su.cuda(), path="tc", topk=5)
    stats = W.ops.evaluate_ranks_prepared.last_stats
    assert stats["flagged"] > 0 and int(stats["deferred"].max()) > 2
    ref_ori, ref = O.match(ov, su)
    check_exact_results(ref, ranks.cpu().numpy(), td, ti, 5)
    # the matrix outputs take the same way out
    ori, dist = W.match(ov.cuda(), su.cuda(), path="tc")
    assert int(W.ops.match.last_flagged) > 0
    same = assert_orientation_is_the_references(ov, su, ori.cpu(), ref_ori)
    assert ((dist.cpu() - ref).abs() / ref.abs())[same].max().item() <= 1e-3


def test_spec_topk_with_duplicate_items_is_proven_or_redone(W):
    """Twenty identical gallery items tie exactly: the 16 candidate keys cannot prove the top 5 complete, the query is
    flagged and re-done in fp32, and the result is the stable sort's (lowest indices first)."""
    ov, su, _ = O.synth_features(200, 40, fov=360, noise=0.5, seed=8)
    ov[100:120] = ov[3]
    ranks, td, ti = W.evaluate_ranks(ov.cuda(), su.cuda(), path="tc", topk=5)
    assert W.ops.evaluate_ranks_prepared.last_stats["flagged"] >= 1
    ref = O.match(ov, su)[1]
    sd = torch.sort(ref.t(), dim=1, stable=True)
    tic = ti.cpu().long()
    assert torch.equal(tic[3], sd.indices[3, :5]) and tic[3].tolist() == [3, 100, 101, 102, 103]
    check_exact_results(ref, ranks.cpu().numpy(), td, ti, 5)
    with pytest.raises(ValueError):
        W.evaluate_ranks(ov.cuda(), su.cuda(), path="tc", topk=13)


def test_spec_properties_at_scale(W):
    """4k x 2k sweep: planted matches are rank 1 with the planted orientation; rolling the gallery moves the orientation;
    two gallery shards (the first one ending inside a group of 8) add up to the unsharded counts."""
    G, Q = 4096, 2048
    ov, su, sh = O.synth_features(G, Q, fov=180, noise=0.5, seed=4)
    ovc, suc = ov.cuda(), su.cuda()
    ranks = W.evaluate_ranks(ovc, suc, path="tc")
    assert int((ranks != 1).sum()) == 0
    ori, dist = W.match(ovc, suc, path="tc")
    assert torch.equal(torch.diagonal(ori[:Q]).cpu(), sh)
    ori2, dist2 = W.match(torch.roll(ovc, 5, dims=3), suc, path="tc")
    flips = ori2 != (ori + 5) % 64
    assert flips.float().mean().item() <= 1e-4                        # only exact fp32 near-ties may land elsewhere
    assert (dist2 - dist)[~flips].abs().max().item() <= 4e-4
    d_true, _ = W.true_match_distances(ovc, suc)
    parts = []
    t32 = torch.arange(Q, dtype=torch.int32, device="cuda")
    for lo, hi in ((0, 1500), (1500, G)):
        cnt = torch.zeros(Q, dtype=torch.int32, device="cuda")
        W.sweep_tc(W.GalleryIndex(ovc[lo:hi], 32, g_offset=lo), W.QueryBatch(suc), d_true=d_true, true_idx=t32, rank_count=cnt)
        parts.append(cnt)
    whole = torch.zeros(Q, dtype=torch.int32, device="cuda")
    W.sweep_tc(W.GalleryIndex(ovc, 32), W.QueryBatch(suc), d_true=d_true, true_idx=t32, rank_count=whole)
    assert torch.equal(parts[0] + parts[1], whole)


@pytest.mark.parametrize("fov,noise", [(360, 25.0), (90, 10.0)])
def test_spec_baseline_size_hard_data_equals_the_oracle(W, fov, noise):
    """BASELINE configs[1] / configs[2] at full size, 10k x 10k, on data where ranks are non-trivial: the ranks and the
    top-10 of a 64-query subset equal the oracle's rank loop (cvig_fov.py:545-552, oracle.rank_loop / oracle.match run on the
    host against the whole gallery) except where fp32 distances tie to 3e-6; nothing was dropped on the way."""
    G = Q = 10000
    sw = int(fov / 360 * 512) // 8
    ov, su, _ = hard_features_cuda(G, Q, sw, noise, seed=5)
    ranks, td, ti = W.evaluate_ranks(ov, su, path="tc", topk=10)
    stats = W.ops.evaluate_ranks_prepared.last_stats
    deferred = stats["deferred"].cpu()
    assert int(deferred.max()) <= stats["list_cap"] and stats["flagged"] == 0
    assert 0 < int(deferred.sum()) <= 0.02 * G * Q
    r = ranks.cpu().numpy()
    assert len(set(r.tolist())) > 1000 and r.max() > 2000              # ranks all over the gallery
    sub = torch.arange(0, Q, Q // 64)[:64]
    ovh, suh = ov.cpu(), su.cpu()
    ref = reference_columns(ovh, suh, sub)
    check_exact_results(ref, r[sub.numpy()], td[sub.cuda()], ti[sub.cuda()], 10, true_rows=sub)
    assert np.array_equal(O.rank_loop(ovh, suh, query_indices=sub.tolist()[:4]), (ref <= ref[sub, torch.arange(64)].unsqueeze(0)).sum(0).numpy()[:4])


def test_spec_baseline_size_10k_x_10k(W):
    """BASELINE configs[1] at full size through size-independent properties: every (gallery, query) pair is visited
    exactly once, planted matches are rank 1 / top-1 with the planted orientation, distances are finite and in [0, 4]."""
    G = Q = 10000
    sw = 64
    ov, su, shifts = hard_features_cuda(G, Q, sw, 0.5, seed=11)
    gal, qry = W.GalleryIndex(ov, sw), W.QueryBatch(su)
    assert gal.impl == "spectral" and qry.impl == "spectral"
    inf = torch.full((Q,), float("inf"), device="cuda")
    cnt = torch.zeros(Q, dtype=torch.int32, device="cuda")
    W.sweep_tc(gal, qry, d_true=inf, rank_count=cnt)
    assert int(cnt.min()) == G and int(cnt.max()) == G
    cnt.zero_()
    W.sweep_tc(gal, qry, d_true=-inf, rank_count=cnt)
    assert int(cnt.abs().max()) == 0
    ranks, td, ti = W.evaluate_ranks_prepared(gal, qry, topk=10)
    assert int((ranks != 1).sum()) == 0
    assert torch.equal(ti[:, 0].long(), torch.arange(Q, device="cuda"))
    assert bool((td[:, 1:] >= td[:, :-1]).all()) and bool(torch.isfinite(td).all())
    assert float(td.min()) >= 0.0 and float(td.max()) <= 4.0
    d_true, o_true = W.true_match_distances(ov, su)
    assert torch.equal(o_true, shifts)
    # the sweep's own orientation and distance on the matches, from a strip of the matrix
    res = W.sweep_tc(W.GalleryIndex(ov[:512], sw), W.QueryBatch(su[:512]), want_dist=True, want_ori=True)
    assert torch.equal(torch.diagonal(res["ori"]).long(), shifts[:512])
    assert (torch.diagonal(res["dist"]) - d_true[:512]).abs().max().item() <= 2e-4


def test_spec_gallery_builder_matches_one_shot_prep(W):
    """Encode-loop plumbing (cvig_fov.py:519-532) on the spectral operand: batches appended one by one give the same
    operand, tables and ranks as preparing the concatenated gallery at once."""
    ov, su, _ = O.synth_features(158, 158, fov=180, noise=6.0, seed=21)
    ovc, suc = ov.cuda(), su.cuda()
    whole = W.GalleryIndex(ovc, 32)
    b = W.GalleryBuilder(200, 32)
    assert b.impl == "spectral" and b.batch_multiple == 8
    for lo, hi in ((0, 64), (64, 128), (128, 158)):
        b.append(ovc[lo:hi])
    built = b.finish()
    assert built.G == 158
    assert torch.equal(built.operand[: whole.operand.numel()], whole.operand)
    for name, width in (("scale", 64), ("aux", 4), ("crop_inv_norm", 64)):
        assert torch.equal(getattr(built, name)[: 160 * width], getattr(whole, name)[: 160 * width]), name
    assert torch.equal(built.spec, whole.spec)
    r1 = W.evaluate_ranks_prepared(whole, W.QueryBatch(suc))
    r2 = W.evaluate_ranks_prepared(built, W.QueryBatch(suc))
    assert torch.equal(r1, r2)
    with pytest.raises(RuntimeError):
        b.append(ovc[:8])


def test_spec_operands_do_not_depend_on_the_feature_scale(W):
    """The operands are norm-scaled before they are rounded to fp16: features a million times larger or smaller (fp16 would
    overflow / flush them) give the same orientations and distances."""
    ov, su, _ = O.synth_features(96, 130, fov=90, noise=2.0, seed=6)
    base = W.match(ov.cuda(), su.cuda(), path="tc16")
    for k_ov, k_su in ((1e6, 1e-6), (2.0 ** -20, 2.0 ** 12)):
        ori, dist = W.match((ov * k_ov).cuda(), (su * k_su).cuda(), path="tc16")
        assert (ori != base[0]).float().mean().item() <= 2e-3
        assert (dist - base[1])[ori == base[0]].abs().max().item() <= 1e-4


def test_spec_nan_and_zero_inputs(W):
    """A zero-norm query / gallery item gives NaN distances (no epsilon in cvig_fov.py:351-361), orientation 0 and rank
    contributions of 0; other pairs are unaffected."""
    ov, su, _ = O.synth_features(64, 40, fov=360, noise=1.0, seed=2)
    su[7] = 0.0
    ov[11] = 0.0
    ref_ori, ref = O.match(ov, su)
    for path in ("tc16", "tc"):
        ori, dist = W.match(ov.cuda(), su.cuda(), path=path)
        ori, dist = ori.cpu(), dist.cpu()
        assert bool(torch.isnan(dist[:, 7]).all()) and bool(torch.isnan(ref[:, 7]).all())
        assert bool(torch.isnan(dist[11]).all()) and bool(torch.isnan(ref[11]).all())
        assert bool((ori[11] == ref_ori[11]).all()) and bool((ori[:, 7] == ref_ori[:, 7]).all())
        keep = torch.ones(64, 40, dtype=torch.bool)
        keep[:, 7] = False
        keep[11] = False
        assert (dist[keep] - ref[keep]).abs().max().item() <= 3e-4
    ranks = W.evaluate_ranks(ov.cuda(), su.cuda(), path="tc").cpu().numpy()
    want = (ref <= torch.diagonal(ref[:40]).unsqueeze(0)).sum(0).numpy()
    assert np.array_equal(ranks, want) and ranks[7] == 0 and ranks[11] == 0


def test_spec_heatmap_sweep_one_query_many_tiles(W):
    """tools/heatmap/heatmap.py:171-177 shape: one photo against a swept grid of tiles -- evaluated entirely in fp32
    (witw_match_columns_spec_f32), so orientation, dissimilarity and score are the reference's."""
    ov, su, sh = O.synth_features(1200, 1, fov=70, noise=0.3, seed=3)
    rdeg, rdis, rscore = O.heatmap_scores(ov, su)
    deg, dis, score = W.heatmap_scores(ov.cuda(), su.cuda())
    assert tuple(deg.shape) == (1200,) and tuple(dis.shape) == (1200,)
    ori = ((deg.cpu() + 180) * 64 / 360).round().long().view(-1, 1)
    same = assert_orientation_is_the_references(ov, su, ori, O.correlation(ov, su)).view(-1)
    assert same.float().mean().item() >= 0.999
    assert (dis.cpu() - rdis)[same].abs().max().item() <= 5e-6
    assert ((score.cpu() - rscore)[same].abs() <= 1e-4 * rscore[same].abs()).all()
    assert int(torch.argmin(dis)) == 0 and float(deg[0]) == float(sh[0]) * 360 / 64 - 180
    # the finished tensor-core sweep gives the same picture
    deg2, dis2, _ = W.heatmap_scores(ov.cuda(), su.cuda(), path="tc")
    assert (deg2.cpu() == rdeg)[same].all() and (dis2.cpu() - rdis)[same].abs().max().item() <= 6e-4


def test_spec_sharded_single_process_matches_unsharded(W):
    """witw_b200/sharded.py on the spectral sweep without a process group: whole gallery as one shard == evaluate_ranks;
    a true match outside the shard is never counted by index."""
    from witw_b200.sharded import CudaLocal, evaluate_ranks_sharded
    ov, su, _ = O.synth_features(600, 300, fov=180, noise=10.0, seed=12)
    perm = torch.randperm(600, generator=torch.Generator().manual_seed(2))
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(600)
    ovp, true_idx = ov[perm].cuda(), inv[:300].cuda()
    want = W.evaluate_ranks(ovp, su.cuda(), true_idx=true_idx, path="tc", topk=5)
    got = evaluate_ranks_sharded(ovp, su.cuda(), 0, 600, true_idx=true_idx, topk=5, local=CudaLocal(path="tc"))
    for a, b in zip(want, got):
        assert torch.equal(a, b)
    ref = O.match(ov[perm], su)[1]
    thr = ref[inv[:300], torch.arange(300)]
    local = CudaLocal(path="tc")
    parts = [local.sweep(ovp[lo:hi], su.cuda(), thr.cuda(), true_idx, lo, 5) for lo, hi in ((0, 288), (288, 600))]
    counts = parts[0][0] + parts[1][0]
    tie = ((ref - thr.unsqueeze(0)).abs() <= 3e-6).sum(0) - 1
    assert bool(((counts.cpu() - want[0].cpu()).abs() <= tie).all())


def test_spec_random_shapes_against_oracle(W):
    """Seeded random problem shapes (gallery not a multiple of 8, queries not a multiple of 128, several work chunks, query
    widths 8..64): distances / orientations of the finished sweep and exact-finish ranks against the fp32 reference chain."""
    rng = np.random.default_rng(2024)
    for trial in range(10):
        G = int(rng.integers(9, 1200))
        Q = int(rng.integers(1, 300))
        sw = int(rng.integers(8, 65))
        gen = torch.Generator().manual_seed(trial)
        ov = torch.randn(G, 16, 4, 64, generator=gen) * 0.06
        su = torch.randn(Q, 16, 4, sw, generator=gen) * 0.06
        n = min(G, Q)
        sh = torch.randint(0, 64, (n,), generator=gen)
        cols = (sh.view(n, 1) + torch.arange(sw).view(1, sw)) % 64
        su[:n] = torch.gather(ov[:n], 3, cols.view(n, 1, 1, sw).expand(n, 16, 4, sw)) + 6.0 * su[:n]
        ref_ori, ref = O.match(ov, su)
        ori, dist = W.match(ov.cuda(), su.cuda(), path="tc")
        same = assert_orientation_is_the_references(ov, su, ori.cpu(), ref_ori)
        assert ((dist.cpu() - ref).abs() / ref.abs())[same].max().item() <= 1e-3, (G, Q, sw)
        if Q <= G:
            ranks = W.evaluate_ranks(ov.cuda(), su.cuda(), path="tc").cpu().numpy()
            check_exact_results(ref, ranks, None, None, 0)


def test_one_query_calls_reuse_the_prepared_gallery_safely(W):
    """The reference's rank loop calls correlation(overhead_embed, one_query) once per query (cvig_fov.py:545-549): the
    gallery is prepared on the first call and reused -- but never for other data: an in-place change or a new tensor (even
    one that could have landed at the same address) is prepared afresh."""
    ov, su, _ = O.synth_features(1500, 3, fov=90, noise=2.0, seed=13)
    ov2, _, _ = O.synth_features(1500, 3, fov=90, noise=2.0, seed=14)
    W.ops.clear_cache()
    a = ov.cuda()
    o1 = W.correlation(a, su[:1].cuda())
    assert len(W.ops._index_cache) == 1
    o1b = W.correlation(a, su[1:2].cuda())
    assert len(W.ops._index_cache) == 1                                  # reused
    def same(got, feats, q):
        return assert_orientation_is_the_references(feats, su[q: q + 1], got.cpu(), O.correlation(feats, su[q: q + 1])).float().mean().item() >= 0.999

    assert same(o1, ov, 0) and same(o1b, ov, 1)
    a.copy_(ov2.cuda())                                                  # in place: same address, new version
    assert same(W.correlation(a, su[:1].cuda()), ov2, 0)
    del a
    b = ov.cuda()                                                        # a new tensor of the same shape
    assert same(W.correlation(b, su[:1].cuda()), ov, 0)
    W.ops.clear_cache()
    assert len(W.ops._index_cache) == 0


@pytest.mark.parametrize("fov,G,Q", [(360, 9, 130), (360, 17, 64), (90, 1003, 129), (180, 4099, 515), (360, 8, 1), (360, 24, 200)])
def test_spec_cta_pair_tiling_equals_one_cta_per_tile(W, fov, G, Q):
    """The CTA-pair sweep (tcgen05 cta_group::2, a pair shares the gallery operand; csrc/match_spec.cu) and the one-CTA-per-tile
    sweep run the same operands through the same accumulation order: every raw output is bit-identical -- on ragged sizes too
    (an odd number of 8-item groups leaves the pair's second CTA a group of zeros past the operand's end)."""
    from witw_b200 import _lib, ops

    ov, su, _ = O.synth_features(G, Q, fov=fov, noise=3.0, seed=G + Q)
    gal, qry = ops.GalleryIndex(ov.cuda(), su.shape[3]), ops.QueryBatch(su.cuda())
    pq = torch.arange(Q, device="cuda") % G
    d_true, _ = ops.pair_distances_prepared(gal, qry, pq, torch.arange(Q, device="cuda"))
    outs = []
    try:
        for variant in (1, 2):
            _lib.call("witw_match_spec_variant", variant)
            for want_ori in (True, False):             # with / without the shift of the maximum (the two epilogues)
                cnt = torch.zeros(Q, dtype=torch.int32, device="cuda")
                r = ops.sweep_tc(gal, qry, want_dist=True, want_ori=want_ori, d_true=d_true, true_idx=pq.to(torch.int32), rank_count=cnt, topk=16)
                torch.cuda.synchronize()
                outs.append((r["dist"].view(torch.int32), r["ori"], cnt, r["topk_dist"].view(torch.int32), r["topk_idx"]))
    finally:
        _lib.call("witw_match_spec_variant", 2)
    for a, b in ((outs[0], outs[2]), (outs[1], outs[3])):
        for x, y in zip(a, b):
            assert (x is None and y is None) or torch.equal(x, y)
    # and the no-argmax epilogue of a full panorama gives the distances of the argmax epilogue
    if su.shape[3] == 64:
        assert torch.equal(outs[2][0], outs[3][0]) and torch.equal(outs[2][2], outs[3][2])
