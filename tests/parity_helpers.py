"""Helpers shared by the GPU parity tests of the two tensor-core sweeps (test_gpu_spec.py, test_gpu_tc.py)."""
import numpy as np
import torch

from oracle import witw_oracle as O


def assert_orientation_is_the_references(ov, su, ori, ref_ori, tol=4e-6):
    """ori equals the reference's argmax except where the float64 correlations of the two shifts tie to fp32 round-off
    (tol of the norm product): there neither fp32 evaluation is wrong."""
    c64 = O.fused_fp64(ov, su)[0]
    scale = ov.double().reshape(ov.shape[0], -1).norm(dim=1).view(-1, 1) * su.double().reshape(su.shape[0], -1).norm(dim=1).view(1, -1)
    a = torch.gather(c64, 2, ori.unsqueeze(-1)).squeeze(-1)
    b = torch.gather(c64, 2, ref_ori.unsqueeze(-1)).squeeze(-1)
    same = ori == ref_ori
    ok = same | ((a - b).abs() <= tol * scale) | ~torch.isfinite(scale * 0 + a + b)
    assert bool(ok.all()), "orientation differs from the reference beyond fp32 ties at %d pairs" % int((~ok).sum())
    return same


def check_exact_results(ref, ranks, td, ti, k, true_rows=None):
    """ranks / top-k against the fp32 reference matrix ref [G,Q]: equal except where fp32 distances tie to 3e-6."""
    G, Q = ref.shape
    rows = torch.arange(Q) if true_rows is None else true_rows
    thr = ref[rows, torch.arange(Q)].unsqueeze(0)
    want = (ref <= thr).sum(0).numpy()
    tie = ((ref - thr).abs() <= 3e-6).sum(0).numpy() - 1
    assert np.all(np.abs(ranks - want) <= tie), "ranks differ beyond fp32 ties at %d queries" % int(np.sum(np.abs(ranks - want) > tie))
    if td is None:
        return want
    tdc, tic = td.cpu(), ti.cpu().long()
    assert (tdc - torch.gather(ref.t(), 1, tic)).abs().max().item() <= 5e-6        # returned distances are the fp32 ones
    assert bool((tdc[:, 1:] >= tdc[:, :-1]).all())
    sd = torch.sort(ref.t(), dim=1, stable=True)
    # every index that differs from the reference's sorted order is a tie: the two items' fp32 distances agree to 3e-6
    mism = tic != sd.indices[:, :k]
    assert ((tdc - sd.values[:, :k]).abs()[mism] <= 3e-6).all(), "top-k index mismatch that is not a tie"
    assert mism.float().mean().item() <= 0.01
    return want


def hard_features_cuda(G, Q, sw, noise, seed):
    """Feature maps generated on the device: every query is its gallery item rolled, cropped and buried in noise, so that
    ranks are spread over the whole gallery (the regime where rank decisions sit inside the distribution's bulk)."""
    gen = torch.Generator(device="cuda").manual_seed(seed)
    ov = torch.randn(G, 16, 4, 64, device="cuda", generator=gen) * 0.06
    shifts = torch.randint(0, 64, (Q,), device="cuda", generator=gen)
    cols = (shifts.view(Q, 1) + torch.arange(sw, device="cuda").view(1, sw)) % 64
    su = torch.gather(ov[:Q], 3, cols.view(Q, 1, 1, sw).expand(Q, 16, 4, sw)) + noise * 0.06 * torch.randn(Q, 16, 4, sw, device="cuda", generator=gen)
    return ov, su, shifts


def reference_columns(ov, su, query_indices):
    """Distances [G, n] of the reference chain on the host for the given queries, one query at a time as test() runs it
    (cvig_fov.py:545-549; a [G,64] crop of a 10k gallery would need 10 GB)."""
    return torch.cat([O.match(ov, su[i: i + 1])[1] for i in query_indices.tolist()], dim=1)
