"""GPU parity tests (run with -m gpu on the B200 box): CUDA path through the C ABI vs the
golden vectors frozen from the reference and vs the CPU oracle."""
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


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def cu(a):
    return (t(a) if isinstance(a, np.ndarray) else a).cuda()


# ------------------------------------------------------------------ K1
def test_polar_exact_kernel_bit_exact(W, golden):
    g = golden("polar")
    out = W.polar_transform(cu(g["tile"]), exact=True)
    assert torch.equal(out.cpu(), t(g["polar"]))


def test_polar_fast_kernel(W, golden):
    g = golden("polar")
    ref = t(g["polar"])
    out = W.polar_transform(cu(g["tile"])).cpu()
    assert out.shape == ref.shape and out.dtype == torch.float32
    # weights are within 2^-24 of the reference's, |tile| < 6: a few fp32 ulps of the tile's scale
    err = (out - ref).abs().max().item()
    assert err <= 4e-6, err
    assert float(out[:, 0, 0].abs().max()) == 0.0 and float(out[:, 0, 384].abs().max()) == 0.0
    rel = ((out - ref).abs() / ref.abs().clamp_min(1e-2)).max().item()
    assert rel <= 1e-3


def test_polar_fast_kernel_many_planes(W):
    # more planes than persistent CTAs x pipeline stages: exercises the TMA ring wrap-around and batch layout
    gen = torch.Generator().manual_seed(3)
    tiles = torch.randn(70, 5, 256, 256, generator=gen)
    out = W.polar_transform(tiles.cuda())
    exact = W.polar_transform(tiles.cuda(), exact=True)
    assert out.shape == (70, 5, 128, 512)
    assert (out - exact).abs().max().item() <= 4e-6
    for n, c in ((0, 0), (33, 2), (69, 4)):
        assert torch.equal(exact[n, c].cpu(), O.polar_transform(tiles[n, c: c + 1])[0])


def test_polar_transform_dropin_contract(W, golden):
    g = golden("polar")
    data = {"overhead": cu(g["tile"]), "surface": 7}
    out = W.PolarTransform(exact=True)(data)
    assert out is data and out["surface"] == 7 and torch.equal(out["polar"].cpu(), t(g["polar"]))
    cpu = W.PolarTransform(exact=True)({"overhead": t(g["tile"])})  # CPU in -> CPU out, computed on the GPU
    assert not cpu["polar"].is_cuda and torch.equal(cpu["polar"], t(g["polar"]))


def test_bilinear_interpolate_generic_bit_exact(W, golden):
    g = golden("bilinear")
    out = W.bilinear_interpolate(cu(g["im"]), g["x"], g["y"])
    assert torch.equal(out.cpu(), t(g["out"]))


# ------------------------------------------------------------------ K2/K3, exact fp32 path
CASES = ["fov360", "fov90", "fov70", "fov180", "fov6", "ties", "zeronorm", "c8h2"]


def _check_match(ori, dist, corr64, ref_ori, ref_dist, atol):
    """orientation must agree except where the two candidate shifts tie to round-off; distance within atol."""
    ori, dist = ori.cpu(), dist.cpu()
    assert ori.dtype == torch.int64 and dist.dtype == torch.float32
    diff = ori != ref_ori
    if diff.any():
        a = torch.gather(corr64, 2, ori.unsqueeze(-1)).squeeze(-1)
        b = torch.gather(corr64, 2, ref_ori.unsqueeze(-1)).squeeze(-1)
        scale = corr64.abs().amax(-1)
        assert bool((((a - b).abs() <= 1e-5 * scale) | ~diff).all()), "orientation differs beyond a round-off tie"
    assert torch.equal(torch.isnan(dist), torch.isnan(ref_dist))
    ok = ~torch.isnan(ref_dist) & ~diff
    assert (dist[ok] - ref_dist[ok]).abs().max().item() <= atol
    return int(diff.sum())


@pytest.mark.parametrize("name", CASES)
def test_match_fp32_vs_golden(W, golden, name):
    g = golden("match")
    ov, su = t(g[name + "_ov"]), t(g[name + "_su"])
    ori, dist = W.match(ov.cuda(), su.cuda(), path="fp32")
    corr64 = O.fused_fp64(ov, su)[0]
    flips = _check_match(ori, dist, corr64, t(g[name + "_ori"]), t(g[name + "_dist"]), atol=5e-6)
    if name == "ties":
        assert flips == 0 and torch.equal(ori.cpu()[2], torch.zeros(6, dtype=torch.int64))
    assert torch.equal(W.correlation(ov.cuda(), su.cuda(), path="fp32"), ori)
    sc = W.correlation_scores(ov.cuda(), su.cuda()).cpu()
    ref_sc = O.correlation_scores(ov, su)
    assert torch.allclose(sc, ref_sc, rtol=0, atol=2e-5 * float(ref_sc.abs().max()) + 1e-12, equal_nan=True)


@pytest.mark.parametrize("name", ["fov90", "ties"])
def test_crop_and_l2_standalone(W, golden, name):
    g = golden("match")
    ov, su, ori = t(g[name + "_ov"]), t(g[name + "_su"]), t(g[name + "_ori"])
    crop = W.crop_overhead(ov.cuda(), ori.cuda(), su.shape[3])
    assert tuple(crop.shape) == (ov.shape[0], su.shape[0], 16, 4, su.shape[3])
    assert torch.equal(crop.cpu(), t(g[name + "_crop"]))          # a pure gather: bit-exact
    dist = W.l2_distance(crop, su.cuda())
    assert torch.allclose(dist.cpu(), t(g[name + "_dist"]), rtol=0, atol=3e-6)


def test_match_pairs_and_ragged_tiles(W):
    ov, su, sh = O.synth_features(37, 45, fov=90, noise=2.0, seed=5)   # not multiples of the 4x32 CTA tile
    ori, dist = W.match(ov.cuda(), su.cuda(), path="fp32")
    ref_ori, ref = O.match(ov, su)
    _check_match(ori, dist, O.fused_fp64(ov, su)[0], ref_ori, ref, atol=5e-6)
    idx = torch.arange(37).flip(0)
    d, o = W.true_match_distances(ov.cuda(), su[:37].cuda(), idx.cuda())
    assert torch.allclose(d.cpu(), ref[idx, torch.arange(37)], rtol=0, atol=5e-6)
    assert torch.equal(o.cpu(), ref_ori[idx, torch.arange(37)])


def test_empty_inputs(W):
    ov, su, _ = O.synth_features(4, 3)
    o, d = W.match(ov[:0].cuda(), su.cuda(), path="fp32")
    assert tuple(o.shape) == (0, 3) and tuple(d.shape) == (0, 3)
    o, d = W.match(ov.cuda(), su[:0].cuda(), path="fp32")
    assert tuple(o.shape) == (4, 0)
    assert tuple(W.rank_from_distances(torch.zeros(0, 5).cuda()).shape) == (5,)
    # the tensor-core path on empty sides: shapes as the reference's, nothing launched on nothing
    o, d = W.match(ov.cuda(), su[:0].cuda(), path="tc")
    assert tuple(o.shape) == (4, 0) and tuple(d.shape) == (4, 0)
    o, d = W.match(ov[:0].cuda(), su.cuda(), path="tc")
    assert tuple(o.shape) == (0, 3)


def test_errors(W):
    ov, su, _ = O.synth_features(4, 3)
    with pytest.raises(RuntimeError):
        W.match(ov.cuda(), su[:, :8].cuda())                      # channel mismatch (conv2d raises in the reference)
    x = ov.cuda().requires_grad_(True)
    with pytest.raises(RuntimeError, match="forward-only"):
        W.match(x, su.cuda())
    with torch.no_grad():
        W.match(x, su.cuda(), path="fp32")
    with pytest.raises(W.WitwError):
        W.match(torch.zeros(2, 16, 4, 48).cuda(), torch.zeros(2, 16, 4, 8).cuda(), path="tc")   # W != 64


# ------------------------------------------------------------------ K4
@pytest.mark.parametrize("name", ["r360", "r90"])
def test_evaluate_ranks_fp32_vs_golden(W, golden, name):
    g = golden("ranks")
    ranks = W.evaluate_ranks(cu(g[name + "_ov"]), cu(g[name + "_su"]), path="fp32")
    assert ranks.dtype == torch.int64
    assert np.array_equal(ranks.cpu().numpy(), g[name + "_ranks"])
    rec = W.recall_from_ranks(ranks)
    assert rec == O.recall_from_ranks(g[name + "_ranks"])


def test_rank_from_distances_and_topk(W):
    gen = torch.Generator().manual_seed(9)
    for (G, Q) in ((1000, 516), (257, 33), (5, 4)):
        d = torch.rand(G, Q, generator=gen)
        d[G // 2, :] = d[0, :]                 # exact ties
        d[min(3, G - 1), 1] = float("nan")     # NaN compares false
        true_idx = torch.randint(0, G, (Q,), generator=gen)
        thr = d[true_idx, torch.arange(Q)]
        want = (d <= thr.unsqueeze(0)).sum(0)
        assert torch.equal(W.rank_from_distances(d.cuda(), true_idx.cuda()).cpu(), want)
        k = min(7, G)
        dd = torch.where(torch.isnan(d), torch.full_like(d, float("inf")), d)
        td, ti = W.topk_from_distances(d.cuda(), k)
        ref_d = torch.sort(dd.t(), dim=1, stable=True)
        assert torch.equal(td.cpu(), ref_d.values[:, :k])
        finite = torch.isfinite(ref_d.values[:, :k])                 # a NaN never enters a list: slot stays (inf, -1)
        assert torch.equal(ti.cpu().long()[finite], ref_d.indices[:, :k][finite])
        assert bool((ti.cpu()[~finite] == -1).all())


@pytest.mark.parametrize("n_lists,Q,k", [(64, 300, 16), (40, 257, 10), (3, 33, 128), (1, 5, 4)])
def test_topk_merge_kway(W, n_lists, Q, k):
    """witw_topk_merge: k-way merge of sorted candidate lists with ties (lower list wins) and (+inf, -1) padding."""
    from witw_b200 import _lib
    gen = torch.Generator().manual_seed(n_lists)
    d = (torch.randint(0, 50, (n_lists, Q, k), generator=gen).float() / 8)     # many exact ties
    idx = torch.arange(n_lists * k, dtype=torch.int32).view(n_lists, 1, k).expand(n_lists, Q, k).clone()
    d, order = torch.sort(d, dim=2, stable=True)
    idx = torch.gather(idx, 2, order)
    idx, _ = torch.sort(idx, dim=2)                                             # ascending index inside a list, as the kernels emit
    fill = torch.randint(0, k + 1, (n_lists, Q), generator=gen)                 # lists filled to a random depth
    pad = torch.arange(k).view(1, 1, k) >= fill.unsqueeze(-1)
    d[pad], idx[pad] = float("inf"), -1
    dc, ic = d.cuda().contiguous(), idx.cuda().contiguous()
    out_d = torch.empty((Q, k), dtype=torch.float32, device="cuda")
    out_i = torch.empty((Q, k), dtype=torch.int32, device="cuda")
    _lib.call("witw_topk_merge", dc.data_ptr(), ic.data_ptr(), n_lists, Q, k, out_d.data_ptr(), out_i.data_ptr(),
              torch.cuda.current_stream().cuda_stream)
    flat_d = d.permute(1, 0, 2).reshape(Q, n_lists * k)                         # list-major: a stable sort prefers the lower list
    flat_i = idx.permute(1, 0, 2).reshape(Q, n_lists * k)
    ref = torch.sort(flat_d, dim=1, stable=True)
    want_d = ref.values[:, :k]
    want_i = torch.gather(flat_i, 1, ref.indices[:, :k])
    assert torch.equal(out_d.cpu(), want_d)
    assert torch.equal(out_i.cpu()[torch.isfinite(want_d)], want_i[torch.isfinite(want_d)])
    assert bool((out_i.cpu()[~torch.isfinite(want_d)] == -1).all())


def test_baseline_ranks_vs_golden(W, golden):
    g = golden("baseline")
    ranks, dist = W.baseline_ranks(cu(g["ov"].astype(np.float32)), cu(g["su"].astype(np.float32)), return_distances=True)
    ov, su = t(g["ov"].astype(np.float32)), t(g["su"].astype(np.float32))
    ref = torch.cdist(ov.double(), su.double())
    assert torch.allclose(dist.cpu().double(), ref, rtol=1e-6)
    # rank ties at fp32 round-off aside, the counts are the reference's
    assert np.array_equal(ranks.cpu().numpy(), g["ranks"])


def test_heatmap_scores(W):
    ov, su, _ = O.synth_features(50, 1, fov=70, noise=0.5, seed=2)
    deg, dis, score = W.heatmap_scores(ov.cuda(), su.cuda(), path="fp32")
    rdeg, rdis, rscore = O.heatmap_scores(ov, su)
    assert torch.equal(deg.cpu(), rdeg) and torch.allclose(dis.cpu(), rdis, atol=5e-6) and torch.allclose(score.cpu(), rscore, rtol=1e-4)


def test_install_on_namespace_runs_rank_loop(W, golden):
    """The reference's rank-loop body (cvig_fov.py:545-552) runs unchanged on the rebound names."""
    import types

    cvig = types.ModuleType("cvig_like")
    for n in ("bilinear_interpolate", "PolarTransform", "correlation", "crop_overhead", "l2_distance"):
        setattr(cvig, n, None)
    W.install(cvig)
    g = golden("ranks")
    overhead_embed, surface_embed = cu(g["r90_ov"]), cu(g["r90_su"])
    count = surface_embed.size(0)
    ranks = np.zeros([count], dtype=int)
    for idx in range(count):
        this_surface_embed = torch.unsqueeze(surface_embed[idx, :], 0)
        orientation_estimate = cvig.correlation(overhead_embed, this_surface_embed)
        overhead_cropped_all = cvig.crop_overhead(overhead_embed, orientation_estimate, this_surface_embed.shape[3])
        distances = torch.squeeze(cvig.l2_distance(overhead_cropped_all, this_surface_embed))
        ranks[idx] = torch.sum(torch.le(distances, distances[idx])).item()
    assert np.array_equal(ranks, g["r90_ranks"])


def test_train_step_gradients_match_reference_autograd(W):
    """The forward/backward slice of train() (cvig_fov.py:450-460): correlation -> crop_overhead -> l2_distance ->
    triplet loss -> backward, on the rebound names, against torch autograd through the oracle's functions."""
    def triplet_loss(distances, alpha=10.0):                     # cvig_fov.py:366-382
        n = distances.shape[0]
        m = torch.diagonal(distances)
        a = torch.sum(torch.log(1.0 + torch.exp(alpha * (m - distances))))
        b = torch.sum(torch.log(1.0 + torch.exp(alpha * (m.unsqueeze(1) - distances))))
        return (a + b) / (2.0 * n * (n - 1))

    gen = torch.Generator().manual_seed(5)
    wts = torch.rand(12, 12, generator=gen)                      # a well-conditioned second loss: weighted sum of distances

    def grads(ov, su, dtype, loss_fn, impl):
        o = ov.to(dtype).clone().requires_grad_(True)
        s = su.to(dtype).clone().requires_grad_(True)
        if impl is O:
            ori = O.correlation(o, s)
            dist = O.l2_distance(O.crop_overhead(o, ori, su.shape[3]), s)
        else:
            o, s = ov.cuda().requires_grad_(True), su.cuda().requires_grad_(True)
            ori = W.correlation(o, s, path="fp32")
            dist = W.l2_distance(W.crop_overhead(o, ori, su.shape[3]), s)
        loss = loss_fn(dist)
        loss.backward()
        return ori.cpu(), loss.item(), o.grad.detach().cpu().double(), s.grad.detach().cpu().double()

    for fov in (360, 90):
        ov, su, _ = O.synth_features(12, 12, fov=fov, noise=1.0, seed=fov + 1)
        for loss_fn in (triplet_loss, lambda d: (d * wts.to(d.device, d.dtype)).sum()):
            ori32, l32, go32, gs32 = grads(ov, su, torch.float32, loss_fn, O)     # the reference's own fp32 autograd
            ori64, l64, go64, gs64 = grads(ov, su, torch.float64, loss_fn, O)     # the truth it approximates
            ori, l, go, gs = grads(ov, su, torch.float32, loss_fn, W)
            assert torch.equal(ori, ori32)
            assert abs(l - l64) <= 1e-5 * abs(l64)
            for got, r32, r64 in ((go, go32, go64), (gs, gs32, gs64)):
                assert got.shape == r64.shape
                # as close to the float64 gradient as fp32 arithmetic allows: within 3x the reference's own fp32 error
                # (the saturated triplet loss is ill-conditioned: fp32 autograd is itself ~2e-3 off), or 1e-5 relative
                allowed = max(3.0 * (r32 - r64).abs().max().item(), 1e-5 * r64.abs().max().item())
                assert (got - r64).abs().max().item() <= allowed
    ov_g, su_g = ov.cuda().requires_grad_(True), su.cuda().requires_grad_(True)
    with pytest.raises(RuntimeError, match="forward-only"):
        W.evaluate_ranks(ov_g, su_g)


def test_fused_train_step_matches_reference_autograd(W):
    """match_distance (no [G,Q,C,H,sw] crop in either direction) + the one-kernel triplet_loss against torch autograd through
    the oracle's functions, fp32 and float64; and against the unfused rebound names, which share the arithmetic."""
    def triplet_loss(distances, alpha=10.0):                     # cvig_fov.py:366-382
        n = distances.shape[0]
        m = torch.diagonal(distances)
        a = torch.sum(torch.log(1.0 + torch.exp(alpha * (m - distances))))
        b = torch.sum(torch.log(1.0 + torch.exp(alpha * (m.unsqueeze(1) - distances))))
        return (a + b) / (2.0 * n * (n - 1))

    def oracle_grads(ov, su, dtype, alpha):
        o = ov.to(dtype).clone().requires_grad_(True)
        s = su.to(dtype).clone().requires_grad_(True)
        dist = O.l2_distance(O.crop_overhead(o, O.correlation(o, s), su.shape[3]), s)
        loss = triplet_loss(dist, alpha)
        loss.backward()
        return dist.detach(), loss.item(), o.grad.double(), s.grad.double()

    for fov, n, alpha in ((360, 16, 10.0), (90, 13, 10.0), (180, 8, 2.0)):
        ov, su, _ = O.synth_features(n, n, fov=fov, noise=1.0, seed=fov + 7)
        d32, l32, go32, gs32 = oracle_grads(ov, su, torch.float32, alpha)
        d64, l64, go64, gs64 = oracle_grads(ov, su, torch.float64, alpha)
        o, s = ov.cuda().requires_grad_(True), su.cuda().requires_grad_(True)
        dist, ori = W.match_distance(o, s)
        assert not ori.requires_grad and torch.equal(ori.cpu(), O.correlation(ov, su))
        assert (dist.detach().cpu() - d32).abs().max().item() <= 5e-6
        loss = W.triplet_loss(dist, alpha)
        assert loss.dim() == 0 and abs(loss.item() - l64) <= 1e-5 * abs(l64)
        loss.backward()
        for got, r32, r64 in ((o.grad.cpu().double(), go32, go64), (s.grad.cpu().double(), gs32, gs64)):
            allowed = max(3.0 * (r32 - r64).abs().max().item(), 1e-5 * r64.abs().max().item())
            assert (got - r64).abs().max().item() <= allowed
        # the loss kernel alone against torch on the same matrix (values and gradient), incl. a non-default alpha
        # (float64 torch is the yardstick: fp32 autograd adds the diagonal's +-alpha/z halves to the small terms before they
        # cancel and loses ~0.2 % of the diagonal gradient; the kernel never forms them)
        dm = dist.detach().double().clone().requires_grad_(True)
        triplet_loss(dm, alpha).backward()
        dk = dist.detach().clone().requires_grad_(True)
        lk = W.triplet_loss(dk, alpha)
        (3.0 * lk).backward()                                     # a scaled upstream gradient
        assert (dk.grad.double() - 3.0 * dm.grad).abs().max().item() <= 3e-6 * 3.0 * dm.grad.abs().max().item()
        # gradient with respect to one side only
        s2 = su.cuda().requires_grad_(True)
        d2, _ = W.match_distance(ov.cuda(), s2)
        d2.sum().backward()
        assert s2.grad is not None and torch.isfinite(s2.grad).all()
    # no gradient requested: plain forward
    d3, o3 = W.match_distance(ov.cuda(), su.cuda())
    assert not d3.requires_grad and torch.equal(o3.cpu(), O.correlation(ov, su))
    with pytest.raises(ValueError):
        W.triplet_loss(torch.zeros(3, 4, device="cuda"))


# ----------------------------------------------------------------------------- f4: uint8 -> normalised polar
def test_normalized_polar_exact_matches_reference_chain(W, golden):
    """exact=True: bit-identical to ImageNormalization -> PolarTransform of the unmodified reference (golden prep.npz)."""
    g = golden("prep")
    tile = torch.from_numpy(g["tile_u8"]).cuda()
    out = W.normalized_polar(tile, exact=True)
    assert out.dtype == torch.float32 and tuple(out.shape) == (3, 128, 512)
    assert torch.equal(out.cpu(), torch.from_numpy(g["polar"]))


def test_normalized_polar_fast_path(W, golden):
    """The staged uint8 kernel: within 4e-6 of the reference chain, exactly 0 at the two zero-weight pixels, batched and
    equal to polar_transform(normalised fp32 tile) to the same tolerance."""
    g = golden("prep")
    tile = torch.from_numpy(g["tile_u8"]).cuda()
    out = W.normalized_polar(tile)
    ref = torch.from_numpy(g["polar"])
    assert (out.cpu() - ref).abs().max().item() <= 4e-6
    assert float(out[:, 0, 0].abs().max()) == 0.0 and float(out[:, 0, 384].abs().max()) == 0.0
    gen = torch.Generator().manual_seed(3)
    batch = torch.randint(0, 256, (37, 3, 256, 256), generator=gen, dtype=torch.uint8)
    got = W.normalized_polar(batch.cuda()).cpu()
    want = torch.stack([O.normalized_polar(b) for b in batch[:5]])
    assert (got[:5] - want).abs().max().item() <= 4e-6
    exact = W.normalized_polar(batch.cuda(), exact=True).cpu()
    assert torch.equal(exact[:5], want)
    assert (got - exact).abs().max().item() <= 4e-6


def test_normalized_polar_rejects_other_inputs(W):
    with pytest.raises(TypeError):
        W.normalized_polar(torch.zeros(3, 256, 256).cuda())
    with pytest.raises(ValueError):
        W.normalized_polar(torch.zeros(3, 750, 750, dtype=torch.uint8).cuda())     # needs the reference's Resize first
    with pytest.raises(RuntimeError):
        W.normalized_polar(torch.zeros(3, 256, 256, dtype=torch.uint8))            # CPU tensor: no fallback


@pytest.mark.parametrize("G,Q,k,levels", [(10000, 512, 10, 0), (16384, 260, 7, 50), (9000, 64, 32, 3), (8192, 1000, 1, 0)])
def test_topk_from_distances_large_gallery_thresholded(W, G, Q, k, levels):
    """Galleries of >= 8192 rows take the two-pass form (strided sample pass -> admission thresholds -> full pass): same
    result as a stable sort, including ties (quantised distances) broken by the lower gallery index, NaN / +inf never entering."""
    gen = torch.Generator().manual_seed(G + k)
    d = torch.rand(G, Q, generator=gen)
    if levels:
        d = torch.floor(d * levels) / levels                       # many exact ties, also at the k-th value
    d[5, :] = float("nan")
    d[7, ::3] = float("inf")
    d[11, 1] = -0.0
    d[13, 2] = -1e-7
    td, ti = W.topk_from_distances(d.cuda(), k, g_offset=3)
    clean = torch.where(torch.isnan(d), torch.full_like(d, float("inf")), d)
    sd = torch.sort(clean.t(), dim=1, stable=True)
    assert torch.equal(td.cpu(), sd.values[:, :k])
    assert torch.equal(ti.cpu().long(), sd.indices[:, :k] + 3)
