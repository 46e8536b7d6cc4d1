"""Pins oracle/witw_oracle.py to the golden vectors frozen from the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import witw_oracle as O

torch.set_num_threads(1)


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def test_polar_grid_matches_reference(golden):
    g = golden("polar")
    x, y = O.polar_grid()
    assert np.array_equal(x[::5, ::7], g["grid_x_sub"])
    assert np.array_equal(y[::5, ::7], g["grid_y_sub"])
    assert np.array_equal(x[0], g["grid_x_row0"]) and np.array_equal(y[0], g["grid_y_row0"])
    # the two pixels the reference maps onto coordinate 255.0 exactly
    assert y[0, 0] == 255.0 and x[0, 384] == 255.0


def test_polar_transform_bit_exact(golden):
    g = golden("polar")
    out = O.polar_transform(t(g["tile"]))
    assert out.dtype == torch.float32 and tuple(out.shape) == (2, 128, 512)
    assert torch.equal(out, t(g["polar"]))
    # clip-after-+1 quirk: all four weights vanish there
    assert float(out[:, 0, 0].abs().max()) == 0.0 and float(out[:, 0, 384].abs().max()) == 0.0


def test_image_normalization_and_polar_chain_bit_exact(golden):
    """uint8 tile -> ImageNormalization -> PolarTransform of the unmodified reference (cvig_fov.py:137-149, 186-209)."""
    g = golden("prep")
    tile = t(g["tile_u8"])
    assert tile.dtype == torch.uint8
    norm = O.image_normalization(tile)
    assert norm.dtype == torch.float32 and torch.equal(norm[:, ::37, :], t(g["norm_rows"]))
    out = O.normalized_polar(tile)
    assert torch.equal(out, t(g["polar"]))
    assert float(out[:, 0, 0].abs().max()) == 0.0 and float(out[:, 0, 384].abs().max()) == 0.0


def test_bilinear_generic_bit_exact(golden):
    g = golden("bilinear")
    out = O.bilinear_interpolate(t(g["im"]), g["x"], g["y"])
    assert torch.equal(out, t(g["out"]))


CASES = ["fov360", "fov90", "fov70", "fov180", "fov6", "ties", "zeronorm", "c8h2"]


@pytest.mark.parametrize("name", CASES)
def test_match_chain(golden, name):
    g = golden("match")
    ov, su = t(g[name + "_ov"]), t(g[name + "_su"])
    ori, dist = O.match(ov, su)
    assert ori.dtype == torch.int64 and tuple(ori.shape) == (ov.shape[0], su.shape[0])
    assert torch.equal(ori, t(g[name + "_ori"]))
    ref = t(g[name + "_dist"])
    assert torch.equal(torch.isnan(dist), torch.isnan(ref))
    assert torch.allclose(dist, ref, rtol=0, atol=2e-6, equal_nan=True)
    if name + "_crop" in g:
        crop = O.crop_overhead(ov, ori, su.shape[3])
        assert torch.equal(crop, t(g[name + "_crop"]))


@pytest.mark.parametrize("name", CASES)
def test_fused_identity_fp64(golden, name):
    """The identity the kernels implement agrees with the reference chain."""
    g = golden("match")
    ov, su = t(g[name + "_ov"]), t(g[name + "_su"])
    corr, ori, dist = O.fused_fp64(ov, su)
    ref_ori, ref = t(g[name + "_ori"]), t(g[name + "_dist"])
    # orientation may differ only where the two best shifts tie to fp32 round-off
    diff = ori != ref_ori
    if diff.any():
        top = torch.gather(corr, 2, ori.unsqueeze(-1)).squeeze(-1)
        alt = torch.gather(corr, 2, ref_ori.unsqueeze(-1)).squeeze(-1)
        scale = corr.abs().amax(dim=-1)
        assert bool((((top - alt).abs() <= 1e-5 * scale) | ~diff).all())
    ok = ~torch.isnan(ref) & ~diff
    assert torch.allclose(dist[ok].float(), ref[ok], rtol=0, atol=5e-6)
    if name == "zeronorm":
        assert bool(torch.isnan(dist[3]).all()) and bool(torch.isnan(dist[:, 1]).all())


@pytest.mark.parametrize("name", ["r360", "r90"])
def test_rank_loop(golden, name):
    g = golden("ranks")
    ranks = O.rank_loop(t(g[name + "_ov"]), t(g[name + "_su"]))
    assert np.array_equal(ranks, g[name + "_ranks"])
    rec = O.recall_from_ranks(ranks)
    assert rec["count"] == len(ranks) and rec["top_percent"] <= rec["top_one"] + 100
    assert rec["top_one"] <= rec["top_five"] <= rec["top_ten"]


def test_baseline_rank_loop(golden):
    g = golden("baseline")
    ranks = O.baseline_rank_loop(t(g["ov"].astype(np.float32)), t(g["su"].astype(np.float32)))
    assert np.array_equal(ranks, g["ranks"])


def test_synth_planted_recovers_shift():
    ov, su, sh = O.synth_features(16, 16, fov=90, noise=0.3, seed=5)
    ori, dist = O.match(ov, su)
    assert torch.equal(torch.diagonal(ori), sh)
    assert bool((torch.diagonal(dist) < 0.5).all())
