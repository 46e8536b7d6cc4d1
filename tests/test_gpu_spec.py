"""GPU parity tests of the spectral tensor-core sweep (csrc/match_spec.cu: per-frequency products on tcgen05 with bf16
spectra, inverse FFT + argmax + distance + rank count + top-k in the epilogue) against the oracle."""
import numpy as np
import pytest
import torch

from oracle import witw_oracle as O

pytestmark = pytest.mark.gpu


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


def spec_model(ov, su):
    """What the kernel computes, in float64: spectra of the fp32 rows rounded to bf16 (the query's scaled by 1/64),
    per-frequency products summed over the feature rows, inverse real FFT; norms of the fp32 inputs."""
    G, Q, sw = ov.shape[0], su.shape[0], su.shape[3]
    pad = torch.zeros(Q, su.shape[1], su.shape[2], 64, dtype=torch.float64)
    pad[..., :sw] = su.double()
    So = torch.fft.rfft(ov.double(), dim=3).reshape(G, -1, 33)
    Sq = torch.fft.rfft(pad, dim=3).reshape(Q, -1, 33) / 64.0

    def rnd(z):
        return torch.complex(z.real.float().bfloat16().double(), z.imag.float().bfloat16().double())

    So, Sq = rnd(So), rnd(Sq)
    P = torch.einsum("grf,qrf->gqf", So, Sq.conj())
    corr = torch.fft.irfft(P, n=64, dim=2) * 64.0                    # irfft divides by 64; the 1/64 is already in Sq
    ori = torch.argmax(corr, -1)
    w = 64
    shift = (torch.arange(w).view(w, 1) + torch.arange(sw).view(1, sw)) % w
    cn = torch.sqrt((ov.double() ** 2).sum((1, 2))[:, shift].sum(-1))
    qn = su.double().reshape(Q, -1).norm(dim=1)
    best = torch.gather(corr, 2, ori.unsqueeze(-1)).squeeze(-1)
    dist = 2 - 2 * best / (torch.gather(cn, 1, ori.reshape(G, -1)).reshape(ori.shape) * qn.unsqueeze(0))
    return corr, ori, dist


@pytest.mark.parametrize("fov,G,Q", [(360, 203, 300), (180, 36, 16), (90, 130, 70), (70, 64, 257), (6, 20, 9), (360, 8, 128), (360, 1, 1)])
def test_spec_match_vs_oracle(W, fov, G, Q):
    ov, su, _ = O.synth_features(G, Q, fov=fov, noise=1.0, seed=fov + G)
    ori, dist = W.match(ov.cuda(), su.cuda(), path="tc")
    ori, dist = ori.cpu(), dist.cpu()
    assert tuple(ori.shape) == (G, Q) and ori.dtype == torch.int64
    # tier 1: against the float64 model of the same arithmetic (bf16 spectra): fp32 transform round-off only
    corr, m_ori, m_dist = spec_model(ov, su)
    diff = ori != m_ori
    a = torch.gather(corr, 2, ori.unsqueeze(-1)).squeeze(-1)
    b = torch.gather(corr, 2, m_ori.unsqueeze(-1)).squeeze(-1)
    assert bool((((a - b).abs() <= 1e-4 * corr.abs().amax(-1)) | ~diff).all())
    assert diff.float().mean().item() <= 0.01
    assert (dist.double() - m_dist)[~diff].abs().max().item() <= 2e-4
    # tier 2: against the fp32 reference chain -> north-star tolerance where the orientation agrees
    ref_ori, ref = O.match(ov, su)
    same = ori == ref_ori
    rel = ((dist - ref).abs() / ref.abs())[same]
    assert rel.max().item() <= (1e-3 if fov == 360 else 4e-3), rel.max().item()
    assert (dist - ref).abs()[same].max().item() <= (2e-3 if su.shape[3] >= 8 else 4e-3)   # one-column queries: flat spectra
    assert same.float().mean().item() >= 0.98
    c32 = O.fused_fp64(ov, su)[0]
    a = torch.gather(c32, 2, ori.unsqueeze(-1)).squeeze(-1)
    b = torch.gather(c32, 2, ref_ori.unsqueeze(-1)).squeeze(-1)
    assert bool((((a - b).abs() <= 2e-2 * c32.abs().amax(-1)) | same).all())


@pytest.mark.parametrize("fov,n,noise", [(360, 300, 25.0), (180, 260, 12.0), (90, 260, 10.0)])
def test_spec_evaluate_ranks_vs_oracle(W, fov, n, noise):
    ov, su, _ = O.synth_features(n, n, fov=fov, noise=noise, seed=17)
    ranks, td, ti = W.evaluate_ranks(ov.cuda(), su.cuda(), path="tc", topk=5, exact=False)
    ranks = ranks.cpu().numpy()
    ref_ori, ref = O.match(ov, su)
    want = (ref <= torch.diagonal(ref).unsqueeze(0)).sum(0).numpy()
    assert len(set(want.tolist())) > 5
    band = ((ref - torch.diagonal(ref).unsqueeze(0)).abs() <= 2e-3).sum(0).numpy() - 1
    assert np.all(np.abs(ranks - want) <= band)
    assert np.mean(ranks == want) >= 0.8
    # fused top-k agrees with a sort of the kernel's own distance matrix
    _, dist = W.match(ov.cuda(), su.cuda(), path="tc")
    sd = torch.sort(dist.t().cpu(), dim=1, stable=True)
    assert torch.equal(td.cpu(), sd.values[:, :5]) and torch.equal(ti.cpu().long(), sd.indices[:, :5])


@pytest.mark.parametrize("fov,n,noise", [(360, 300, 25.0), (180, 280, 12.0), (90, 260, 10.0)])
def test_spec_exact_finish_matches_fp32_reference(W, fov, n, noise):
    """exact=True on the spectral sweep: ranks and top-k are the fp32 reference's (cvig_fov.py:547-552)."""
    ov, su, _ = O.synth_features(n, n, fov=fov, noise=noise, seed=23)
    ranks, td, ti = W.evaluate_ranks(ov.cuda(), su.cuda(), path="tc", topk=5)
    ranks = ranks.cpu().numpy()
    ref_ori, ref = O.match(ov, su)
    thr = torch.diagonal(ref).unsqueeze(0)
    want = (ref <= thr).sum(0).numpy()
    assert len(set(want.tolist())) > 5
    tie = ((ref - thr).abs() <= 3e-6).sum(0).numpy() - 1
    assert np.all(np.abs(ranks - want) <= tie)
    assert np.mean(ranks == want) >= 0.98
    appended, dropped = W.ops.evaluate_ranks_prepared.last_recheck.cpu().tolist()
    assert dropped == 0 and appended > 0
    tdc, tic = td.cpu(), ti.cpu().long()
    assert (tdc - torch.gather(ref.t(), 1, tic)).abs().max().item() <= 5e-6
    assert bool((tdc[:, 1:] >= tdc[:, :-1]).all())
    sd = torch.sort(ref.t(), dim=1, stable=True)
    assert torch.equal(tic[:, 0], sd.indices[:, 0])
    assert (tic == sd.indices[:, :5]).float().mean().item() >= (0.99 if fov == 360 else 0.9)


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
    assert flips.float().mean().item() <= 0.02                        # the rolled spectra round differently
    assert (dist2 - dist)[~flips].abs().max().item() <= 2e-3
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


def test_spec_baseline_size_10k_x_10k(W):
    """BASELINE configs[1] at full size through size-independent properties: every (gallery, query) pair is visited
    exactly once, planted matches are rank 1 / top-1 with the planted orientation, distances are finite and in [0, 4]."""
    G = Q = 10000
    sw = 64
    gen = torch.Generator(device="cuda").manual_seed(11)
    ov = torch.randn(G, 16, 4, 64, device="cuda", generator=gen) * 0.06
    shifts = torch.randint(0, 64, (Q,), device="cuda", generator=gen)
    cols = (shifts.view(Q, 1) + torch.arange(sw, device="cuda").view(1, sw)) % 64
    su = torch.gather(ov, 3, cols.view(Q, 1, 1, sw).expand(Q, 16, 4, sw)) + 0.03 * torch.randn(Q, 16, 4, sw, device="cuda", generator=gen)
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
    # the bf16 sweep's own orientation and distance on the matches, from a strip of the matrix
    res = W.sweep_tc(W.GalleryIndex(ov[:512], sw), W.QueryBatch(su[:512]), want_dist=True, want_ori=True)
    assert torch.equal(torch.diagonal(res["ori"]).long(), shifts[:512])
    assert (torch.diagonal(res["dist"]) - d_true[:512]).abs().max().item() <= 2e-3


def test_spec_gallery_builder_matches_one_shot_prep(W):
    """Encode-loop plumbing (cvig_fov.py:519-532) on the spectral operand: batches appended one by one give the same
    operand, crop norms and ranks as preparing the concatenated gallery at once."""
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
    assert torch.equal(built.crop_inv_norm[: 160 * 64], whole.crop_inv_norm[: 160 * 64])
    r1 = W.evaluate_ranks_prepared(whole, W.QueryBatch(suc))
    r2 = W.evaluate_ranks_prepared(built, W.QueryBatch(suc))
    assert torch.equal(r1, r2)
    with pytest.raises(RuntimeError):
        b.append(ovc[:8])


def test_spec_nan_and_zero_inputs(W):
    """A zero-norm query gives NaN distances (no epsilon in cvig_fov.py:351-361) and rank 0; other queries are unaffected."""
    ov, su, _ = O.synth_features(64, 40, fov=360, noise=1.0, seed=2)
    su[7] = 0.0
    ori, dist = W.match(ov.cuda(), su.cuda(), path="tc")
    ref_ori, ref = O.match(ov, su)
    assert bool(torch.isnan(dist[:, 7]).all()) and bool(torch.isnan(ref[:, 7]).all())
    keep = torch.ones(40, dtype=torch.bool)
    keep[7] = False
    assert (dist.cpu()[:, keep] - ref[:, keep]).abs().max().item() <= 2e-3


def test_spec_heatmap_sweep_one_query_many_tiles(W):
    """tools/heatmap/heatmap.py:171-177 shape on the spectral sweep: one photo against a swept grid of tiles."""
    ov, su, sh = O.synth_features(1200, 1, fov=70, noise=0.3, seed=3)
    rdeg, rdis, rscore = O.heatmap_scores(ov, su)
    deg, dis, score = W.heatmap_scores(ov.cuda(), su.cuda(), path="tc")
    assert tuple(deg.shape) == (1200,) and tuple(dis.shape) == (1200,)
    same = deg.cpu() == rdeg
    assert same.float().mean().item() >= 0.98
    assert (dis.cpu() - rdis)[same].abs().max().item() <= 2e-3
    assert (score.cpu() - rscore)[same].abs().max().item() <= 2e-2 * float(rscore.max())
    assert int(torch.argmin(dis)) == 0 and float(deg[0]) == float(sh[0]) * 360 / 64 - 180


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
    widths 8..64): distances / orientations of the sweep and exact-finish ranks against the fp32 reference chain."""
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
        same = ori.cpu() == ref_ori
        assert same.float().mean().item() >= 0.97, (G, Q, sw)
        assert (dist.cpu() - ref)[same].abs().max().item() <= 3e-3, (G, Q, sw)
        if Q <= G:
            ranks = W.evaluate_ranks(ov.cuda(), su.cuda(), path="tc").cpu().numpy()
            thr = torch.diagonal(ref[:Q]).unsqueeze(0)
            want = (ref <= thr).sum(0).numpy()
            tie = ((ref - thr).abs() <= 3e-6).sum(0).numpy() - 1
            assert np.all(np.abs(ranks - want) <= tie), (G, Q, sw)
