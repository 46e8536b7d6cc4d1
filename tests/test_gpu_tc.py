"""GPU parity tests of the dense tcgen05 sweep (csrc/match_tc.cu: fp16 operands of the norm-scaled features, fp32
accumulation) and of its fp32 finish against the oracle."""
import numpy as np
import pytest
import torch

from oracle import witw_oracle as O
from parity_helpers import assert_orientation_is_the_references, check_exact_results, hard_features_cuda, reference_columns

pytestmark = pytest.mark.gpu

KAPPA = 256.0    # csrc/sweep_common.cuh: kDenseKappa


@pytest.fixture(scope="module")
def W():
    assert torch.cuda.is_available()
    import witw_b200

    return witw_b200


@pytest.fixture(autouse=True)
def _hankel_sweep(W):
    """This module tests the dense-contraction sweep (csrc/match_tc.cu); tests/test_gpu_spec.py covers the spectral one."""
    W.ops.TC_IMPL = "hankel"
    yield
    W.ops.TC_IMPL = "auto"
    W.ops.DEFERRAL_CAP = None


def fp16_model(ov, su):
    """What the raw sweep computes, in float64: correlation of the features scaled by 256 / norm and rounded to fp16,
    norms of the fp32 inputs.  Returns (normalised correlation, orientation, distance)."""
    G, Q = ov.shape[0], su.shape[0]
    gn = ov.double().reshape(G, -1).norm(dim=1)
    qn = su.double().reshape(Q, -1).norm(dim=1)
    o16 = (ov.double() * (KAPPA / gn).view(G, 1, 1, 1)).float().half().float()
    s16 = (su.double() * (KAPPA / qn).view(Q, 1, 1, 1)).float().half().float()
    corr = O.fused_fp64(o16, s16)[0] / KAPPA ** 2
    ori = torch.argmax(corr, -1)
    w, sw = ov.shape[3], su.shape[3]
    shift = (torch.arange(w).view(w, 1) + torch.arange(sw).view(1, sw)) % w
    cn = torch.sqrt((ov.double() ** 2).sum((1, 2))[:, shift].sum(-1))                       # [G,W]
    best = torch.gather(corr, 2, ori.unsqueeze(-1)).squeeze(-1)
    ratio = gn.view(G, 1) / torch.gather(cn, 1, ori.reshape(G, -1)).reshape(ori.shape)
    return corr, ori, 2 - 2 * best * ratio


@pytest.mark.parametrize("fov,G,Q", [(360, 203, 300), (90, 130, 70), (70, 64, 257), (180, 36, 16), (6, 20, 9)])
def test_tc_match_vs_oracle(W, fov, G, Q):
    ov, su, _ = O.synth_features(G, Q, fov=fov, noise=1.0, seed=fov)
    # tier 1: the raw fp16 sweep against the float64 model of the same arithmetic -> only accumulation-order noise
    ori, dist = W.match(ov.cuda(), su.cuda(), path="tc16")
    ori, dist = ori.cpu(), dist.cpu()
    assert tuple(ori.shape) == (G, Q) and ori.dtype == torch.int64
    corr, m_ori, m_dist = fp16_model(ov, su)
    diff = ori != m_ori
    a = torch.gather(corr, 2, ori.unsqueeze(-1)).squeeze(-1)
    b = torch.gather(corr, 2, m_ori.unsqueeze(-1)).squeeze(-1)
    assert bool((((a - b).abs() <= 4e-6 + 5e-5 * b.abs()) | ~diff).all())
    # the tensor core's fp32 accumulation loses up to an ulp of the running sum in each of its 256 steps (K = 4096): strongly
    # correlated pairs come out up to 1.7e-5 of their correlation low (measured; csrc/sweep_common.cuh budgets 5e-5)
    assert bool(((dist.double() - m_dist).abs() <= 4e-6 + 1e-4 * (1 - m_dist / 2).abs())[~diff].all())
    # tier 2: the finished sweep against the fp32 reference chain: the reference's orientation (up to fp32 ties), distances
    # within the north star's 1e-3 relative at every field of view
    ref_ori, ref = O.match(ov, su)
    ori, dist = W.match(ov.cuda(), su.cuda(), path="tc")
    ori, dist = ori.cpu(), dist.cpu()
    same = assert_orientation_is_the_references(ov, su, ori, ref_ori)
    rel = ((dist - ref).abs() / ref.abs())[same]
    assert rel.max().item() <= 1e-3, rel.max().item()
    assert (dist - ref)[same].abs().max().item() <= (6e-4 if su.shape[3] >= 8 else 1.5e-3)
    if G * Q >= 1000:
        assert same.float().mean().item() >= 0.995


@pytest.mark.parametrize("fov,n,noise", [(360, 300, 25.0), (90, 260, 10.0)])
def test_tc_raw_sweep_ranks_and_topk(W, fov, n, noise):
    """exact=False: the raw fp16 decisions."""
    ov, su, _ = O.synth_features(n, n, fov=fov, noise=noise, seed=17)
    ranks, td, ti = W.evaluate_ranks(ov.cuda(), su.cuda(), path="tc", topk=5, exact=False)
    ranks = ranks.cpu().numpy()
    ref_ori, ref = O.match(ov, su)
    want = (ref <= torch.diagonal(ref).unsqueeze(0)).sum(0).numpy()
    assert len(set(want.tolist())) > 5                                   # non-degenerate ranks
    # identical counts except for gallery items whose fp32 distance ties the threshold within the fp16 error
    band = ((ref - torch.diagonal(ref).unsqueeze(0)).abs() <= (3e-4 if fov == 360 else 1.2e-2)).sum(0).numpy() - 1
    assert np.all(np.abs(ranks - want) <= band)
    assert np.mean(ranks == want) >= 0.9
    # fused top-k agrees with a sort of the kernel's own distance matrix
    _, dist = W.match(ov.cuda(), su.cuda(), path="tc16")
    sd = torch.sort(dist.t().cpu(), dim=1, stable=True)
    assert torch.equal(td.cpu(), sd.values[:, :5]) and torch.equal(ti.cpu().long(), sd.indices[:, :5])


def test_tc_properties_at_scale(W):
    """Size-independent properties on a 4k x 2k sweep: planted matches are rank 1 with the planted
    orientation; rolling the gallery item moves the orientation and leaves the distance unchanged;
    two gallery shards add up to the unsharded counts."""
    G, Q = 4096, 2048
    ov, su, sh = O.synth_features(G, Q, fov=90, noise=0.5, seed=4)
    ovc, suc = ov.cuda(), su.cuda()
    ranks = W.evaluate_ranks(ovc, suc, path="tc")
    assert int((ranks != 1).sum()) == 0
    ori, dist = W.match(ovc, suc, path="tc16")
    assert torch.equal(torch.diagonal(ori[:Q]).cpu(), sh)
    rolled = torch.roll(ovc, 5, dims=3)
    ori2, dist2 = W.match(rolled, suc, path="tc16")
    flips = ori2 != (ori + 5) % 64                                        # same products in the same order; the item norms are summed in
    assert flips.float().mean().item() <= 1e-4                            # another order, which can move an operand element by an fp16 ulp
    assert (dist2 - dist)[~flips].abs().max().item() <= 4e-5                # an element one fp16 ulp off moves a matched pair this far
    d_true, _ = W.true_match_distances(ovc, suc)
    parts = []
    t32 = torch.arange(Q, dtype=torch.int32, device="cuda")
    for lo, hi in ((0, 1500), (1500, G)):
        cnt = torch.zeros(Q, dtype=torch.int32, device="cuda")
        W.sweep_tc(W.GalleryIndex(ovc[lo:hi], 16, g_offset=lo), W.QueryBatch(suc), d_true=d_true, true_idx=t32, rank_count=cnt)
        parts.append(cnt)
    assert torch.equal((parts[0] + parts[1]).long(), ranks)


@pytest.mark.parametrize("fov", [360, 90])
def test_tc_baseline_size_10k_x_10k(W, fov):
    """BASELINE configs[1]/[2] at full size through size-independent properties:
    every (gallery, query) pair is visited exactly once (count with an infinite threshold == G),
    planted matches are rank 1 / top-1 with the planted orientation, distances are finite and in [0, 4]."""
    G = Q = 10000
    sw = int(fov / 360 * 512) // 8
    ov, su, shifts = hard_features_cuda(G, Q, sw, 0.5, seed=11)
    gal, qry = W.GalleryIndex(ov, sw), W.QueryBatch(su)
    assert gal.impl == "hankel"
    inf = torch.full((Q,), float("inf"), device="cuda")
    cnt = torch.zeros(Q, dtype=torch.int32, device="cuda")
    W.sweep_tc(gal, qry, d_true=inf, rank_count=cnt)
    assert int(cnt.min()) == G and int(cnt.max()) == G                    # coverage: each pair exactly once
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
    assert (td[:, 0] - d_true).abs().max().item() <= 5e-6                 # the finish's fp32 distances on the matches
    rec = W.recall_from_ranks(ranks)
    assert rec["top_one"] == 100.0 and rec["top_percent"] == 100.0 and rec["count"] == Q


def test_tc_baseline_size_hard_data_equals_the_oracle(W):
    """The dense sweep at 10k x 10k, 90 degrees, on data with ranks all over the gallery: ranks and top-10 of a 32-query
    subset equal the oracle's (cvig_fov.py:545-552) except where fp32 distances tie to 3e-6; nothing dropped."""
    G = Q = 10000
    ov, su, _ = hard_features_cuda(G, Q, 16, 10.0, seed=5)
    ranks, td, ti = W.evaluate_ranks(ov, su, path="tc", topk=10)
    stats = W.ops.evaluate_ranks_prepared.last_stats
    assert int(stats["deferred"].max()) <= stats["list_cap"] and stats["flagged"] == 0
    r = ranks.cpu().numpy()
    assert len(set(r.tolist())) > 1000
    sub = torch.arange(0, Q, Q // 32)[:32]
    ref = reference_columns(ov.cpu(), su.cpu(), sub)
    check_exact_results(ref, r[sub.numpy()], td[sub.cuda()], ti[sub.cuda()], 10, true_rows=sub)


def test_tc_semantic_config_shapes(W):
    """BASELINE configs[4] shapes at reduced count: 5-channel tiles through the polar transform, 90-degree
    queries against a larger gallery than queries (Q != G), explicit true_idx."""
    tiles = torch.randn(6, 5, 256, 256, device="cuda")
    polar = W.polar_transform(tiles)
    assert tuple(polar.shape) == (6, 5, 128, 512)
    assert (polar - W.polar_transform(tiles, exact=True)).abs().max().item() <= 4e-6
    ov, su, sh = O.synth_features(3000, 500, fov=90, noise=0.5, seed=8)
    perm = torch.randperm(3000, generator=torch.Generator().manual_seed(1))
    ovp = ov[perm]                                                        # query i now matches gallery item inv[i]
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(3000)
    ranks = W.evaluate_ranks(ovp.cuda(), su.cuda(), true_idx=inv[:500].cuda(), path="tc")
    assert int((ranks != 1).sum()) == 0


def test_gallery_builder_matches_one_shot_prep(W):
    """Encode-loop plumbing (cvig_fov.py:519-532): batches appended one by one give the same operand, crop norms
    and ranks as preparing the concatenated gallery at once."""
    ov, su, _ = O.synth_features(158, 158, fov=90, noise=6.0, seed=21)
    ovc, suc = ov.cuda(), su.cuda()
    whole = W.GalleryIndex(ovc, 16)
    b = W.GalleryBuilder(200, 16)
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
        b.append(ovc[:4])          # the 30-item batch closed the builder


def test_heatmap_sweep_one_query_many_tiles(W):
    """tools/heatmap/heatmap.py:171-177 shape: one photo against a swept grid of tiles, every path."""
    ov, su, sh = O.synth_features(1200, 1, fov=70, noise=0.3, seed=3)
    rdeg, rdis, rscore = O.heatmap_scores(ov, su)
    ref_ori = O.correlation(ov, su)
    for path in ("fp32", "auto", "tc"):
        deg, dis, score = W.heatmap_scores(ov.cuda(), su.cuda(), path=path)
        assert tuple(deg.shape) == (1200,) and tuple(dis.shape) == (1200,)
        ori = ((deg.cpu() + 180) * 64 / 360).round().long().view(-1, 1)
        same = assert_orientation_is_the_references(ov, su, ori, ref_ori).view(-1)
        assert same.float().mean().item() >= 0.999
        assert (dis.cpu() - rdis)[same].abs().max().item() <= (6e-4 if path == "tc" else 5e-6)
        assert int(torch.argmin(dis)) == 0 and float(deg[0]) == float(sh[0]) * 360 / 64 - 180


@pytest.mark.parametrize("row_len", [64, 16, 13, 1])
def test_spectral_rows_match_rfft(W, row_len):
    """witw_spectral_rows_f32: packed 64-point spectra of (zero-padded) rows against numpy's float64 rfft."""
    gen = torch.Generator().manual_seed(row_len)
    x = torch.randn(517, row_len, generator=gen)
    spec = W.ops.spectral_rows(x.cuda(), row_len).cpu().numpy().astype(np.float64)
    pad = np.zeros((517, 64))
    pad[:, :row_len] = x.numpy()
    ref = np.fft.rfft(pad, axis=1)
    want = np.empty((517, 64))
    want[:, 0::2], want[:, 1::2] = ref.real[:, :32], ref.imag[:, :32]
    want[:, 1] = ref.real[:, 32]
    scale = np.abs(ref).max()
    assert np.abs(spec - want).max() <= 4e-7 * scale


@pytest.mark.parametrize("fov,G,Q", [(360, 96, 300), (90, 130, 70), (70, 64, 257), (6, 20, 9)])
def test_spectral_pair_distances_vs_oracle(W, fov, G, Q):
    """Exact fp32 pairs through the correlation theorem (csrc/spectral.cu) against the reference chain
    (cvig_fov.py:547-549): every (gallery, query) pair of a small problem."""
    ov, su, _ = O.synth_features(G, Q, fov=fov, noise=2.0, seed=31)
    ov[3] = ov[2]                                             # exact ties between neighbouring items
    ref_ori, ref = O.match(ov, su)
    gallery, queries = W.GalleryIndex(ov.cuda(), su.shape[3]), W.QueryBatch(su.cuda())
    pg = torch.arange(G).repeat_interleave(Q).cuda()
    pq = torch.arange(Q).repeat(G).cuda()
    d, o = W.ops.pair_distances_prepared(gallery, queries, pg, pq)
    d, o = d.cpu().view(G, Q), o.cpu().view(G, Q)
    diff = o != ref_ori
    assert (d - ref)[~diff].abs().max().item() <= 5e-6
    if bool(diff.any()):                                      # only between shifts whose float64 scores tie to fp32 round-off
        corr = O.fused_fp64(ov, su)[0]
        a = torch.gather(corr, 2, o.unsqueeze(-1)).squeeze(-1)
        b = torch.gather(corr, 2, ref_ori.unsqueeze(-1)).squeeze(-1)
        scale = corr.abs().amax(-1)
        assert ((a - b).abs()[diff] <= 2e-6 * scale[diff]).all()
    assert diff.float().mean().item() <= 1e-3


@pytest.mark.parametrize("fov,n,noise", [(360, 300, 25.0), (90, 260, 10.0), (70, 200, 8.0)])
def test_tc_exact_finish_matches_fp32_reference(W, fov, n, noise):
    """exact=True: rank decisions the fp16 operands cannot settle are taken in fp32 and the top-k re-ranked in fp32, so
    ranks and top-k are the fp32 reference's (cvig_fov.py:547-552) -- differences only where fp32 distances themselves tie."""
    ov, su, _ = O.synth_features(n, n, fov=fov, noise=noise, seed=23)
    ranks, td, ti = W.evaluate_ranks(ov.cuda(), su.cuda(), path="tc", topk=5)       # exact=True is the default
    ref_ori, ref = O.match(ov, su)
    want = check_exact_results(ref, ranks.cpu().numpy(), td, ti, 5)
    assert len(set(want.tolist())) > 5
    stats = W.ops.evaluate_ranks_prepared.last_stats
    deferred = stats["deferred"].cpu()
    assert int(deferred.max()) <= stats["list_cap"] and int(deferred.sum()) > 0 and stats["flagged"] == 0
    assert int(deferred.sum()) <= 0.03 * n * n


def test_tc_exact_finish_survives_list_overflow(W):
    n = 400
    ov, su, _ = O.synth_features(n, n, fov=90, noise=10.0, seed=31)
    W.ops.DEFERRAL_CAP = 2
    ranks, td, ti = W.evaluate_ranks(ov.cuda(), su.cuda(), path="tc", topk=5)
    assert W.ops.evaluate_ranks_prepared.last_stats["flagged"] > 0
    check_exact_results(O.match(ov, su)[1], ranks.cpu().numpy(), td, ti, 5)


def test_sharded_cuda_local_single_process(W):
    """witw_b200/sharded.py on the CUDA kernels without a process group: the whole gallery as one shard gives the ranks
    and top-k of evaluate_ranks; a slice of it (true matches partly outside the slice) gives that slice's counts."""
    from witw_b200.sharded import CudaLocal, evaluate_ranks_sharded
    ov, su, _ = O.synth_features(600, 300, fov=90, noise=10.0, seed=12)
    perm = torch.randperm(600, generator=torch.Generator().manual_seed(2))
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(600)
    ovp, true_idx = ov[perm].cuda(), inv[:300].cuda()
    want = W.evaluate_ranks(ovp, su.cuda(), true_idx=true_idx, path="tc", topk=5)
    got = evaluate_ranks_sharded(ovp, su.cuda(), 0, 600, true_idx=true_idx, topk=5, local=CudaLocal(path="tc"))
    for a, b in zip(want, got):
        assert torch.equal(a, b)
    # two shards evaluated one after the other with the exchanged thresholds: counts add up, candidates merge
    ref = O.match(ov[perm], su)[1]
    thr = ref[inv[:300], torch.arange(300)]
    local = CudaLocal(path="tc")
    parts = [local.sweep(ovp[lo:hi], su.cuda(), thr.cuda(), true_idx, lo, 5) for lo, hi in ((0, 288), (288, 600))]
    counts = parts[0][0] + parts[1][0]
    tie = ((ref - thr.unsqueeze(0)).abs() <= 3e-6).sum(0) - 1
    assert bool(((counts.cpu() - want[0].cpu()).abs() <= tie).all())
    td, ti = local.merge(torch.stack([parts[0][1], parts[1][1]]), torch.stack([parts[0][2], parts[1][2]]), 5)
    assert torch.equal(ti, want[2]) and (td - want[1]).abs().max().item() <= 1e-6


def test_tc_random_shapes_exact_finish_against_oracle(W):
    """Seeded random problem shapes on the dense sweep, narrow to full query widths: exact-finish ranks equal the fp32
    reference chain's up to fp32 round-off ties."""
    rng = np.random.default_rng(77)
    for trial in range(8):
        G = int(rng.integers(40, 900))
        Q = int(rng.integers(1, min(G, 260) + 1))
        sw = int(rng.integers(4, 65))
        gen = torch.Generator().manual_seed(100 + trial)
        ov = torch.randn(G, 16, 4, 64, generator=gen) * 0.06
        su = torch.randn(Q, 16, 4, sw, generator=gen) * 0.06
        sh = torch.randint(0, 64, (Q,), generator=gen)
        cols = (sh.view(Q, 1) + torch.arange(sw).view(1, sw)) % 64
        su = torch.gather(ov[:Q], 3, cols.view(Q, 1, 1, sw).expand(Q, 16, 4, sw)) + 6.0 * su
        ref = O.match(ov, su)[1]
        ranks = W.evaluate_ranks(ov.cuda(), su.cuda(), path="tc").cpu().numpy()
        check_exact_results(ref, ranks, None, None, 0)
