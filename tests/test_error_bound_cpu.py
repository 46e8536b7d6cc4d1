"""CPU check of the error bound the tensor-core sweeps work with (csrc/sweep_common.cuh): a float64 model of the spectral
sweep's arithmetic -- spectra of the norm-scaled features rounded to fp16 -- never leaves the per-pair bound
e = err_sigmas * sqrt(2 * 2^-20 / 12) * (2/64) * ||O||_4 ||S||_4, decisions outside the slack are the exact ones, and the
top-k keys prove the candidate list complete.  Gaussian, sparse and heavy-tailed features, full and limited field of view."""
import os
import re

import numpy as np
import pytest
import torch

from oracle import witw_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KAPPA, ROUND_SIGMA, ACC_FLOOR, ERR_SIGMAS = 16.0, 3.9867e-4, 2e-6, 5.0


def test_constants_are_the_headers():
    text = open(os.path.join(ROOT, "witw_b200", "csrc", "sweep_common.cuh")).read()
    assert float(re.search(r"kSpecKappa = ([0-9.e+-]+)f", text).group(1)) == KAPPA
    assert float(re.search(r"kRoundSigma = ([0-9.e+-]+)f", text).group(1)) == ROUND_SIGMA
    assert float(re.search(r"kAccFloor = ([0-9.e+-]+)f", text).group(1)) == ACC_FLOOR
    assert abs(ROUND_SIGMA - np.sqrt(2 * 2.0 ** -20 / 12)) < 1e-8
    from witw_b200 import ops
    assert ops.ERR_SIGMAS == ERR_SIGMAS


def fp16(x):
    return x.astype(np.float32).astype(np.float16).astype(np.float64)


def sweep_model(ov, su):
    """(normalised exact correlation, modelled fp16-sweep correlation, per-pair bound e, scale ratio r[g,s])."""
    G, Q, sw = ov.shape[0], su.shape[0], su.shape[3]
    corr = O.fused_fp64(ov, su)[0].numpy()
    ovn = ov.double().numpy().reshape(G, 64, 64)
    sun = np.zeros((Q, 64, 64))
    sun[:, :, :sw] = su.double().numpy().reshape(Q, 64, sw)
    gn, qn = np.sqrt((ovn ** 2).sum((1, 2))), np.sqrt((sun ** 2).sum((1, 2)))
    shift = (np.arange(64)[:, None] + np.arange(sw)[None, :]) % 64
    r = gn[:, None] / np.sqrt((ovn ** 2).sum(1)[:, shift].sum(-1))
    On, Sn = np.fft.rfft(ovn, axis=2) / gn[:, None, None], np.fft.rfft(sun, axis=2) / qn[:, None, None]

    def quartic(X):     # 4-norm of the packed half spectrum: bins 0 and 32 are real
        return ((np.abs(X[:, :, 1:32]) ** 4).sum((1, 2)) + (X[:, :, 0].real ** 4).sum(1) + (X[:, :, 32].real ** 4).sum(1)) ** 0.25

    e = np.maximum(ERR_SIGMAS * ROUND_SIGMA * (2 / 64) * quartic(On)[:, None] * quartic(Sn)[None, :], ACC_FLOOR)
    Or, Sr = fp16(On.real * KAPPA) + 1j * fp16(On.imag * KAPPA), fp16(Sn.real * KAPPA) + 1j * fp16(Sn.imag * KAPPA)
    c = np.fft.irfft(np.einsum("grf,qrf->gqf", Or, np.conj(Sr)), n=64, axis=2) / KAPPA ** 2
    return corr / (gn[:, None, None] * qn[None, :, None]), c, e, r


@pytest.mark.parametrize("kind", ["gauss", "sparse", "heavy"])
@pytest.mark.parametrize("fov,noise", [(360, 25.0), (90, 10.0), (45, 6.0)])
def test_bound_holds_and_decisions_outside_the_slack_are_exact(kind, fov, noise):
    G, Q, k = 640, 160, 10
    ov, su, _ = O.synth_features(G, Q, fov=fov, noise=noise, seed=5)
    gen = torch.Generator().manual_seed(1)
    if kind == "sparse":
        ov = ov * (torch.rand(ov.shape, generator=gen) < 0.05)
        su = su * (torch.rand(su.shape, generator=gen) < 0.2)
    elif kind == "heavy":
        ov = ov * torch.exp(1.5 * torch.randn(ov.shape, generator=gen))
        su = su * torch.exp(1.5 * torch.randn(su.shape, generator=gen))
    c_true, c, e, r = sweep_model(ov, su)
    assert (np.abs(c - c_true) <= 0.7 * e[..., None]).all()            # the bound, with room to spare
    dist = O.fused_fp64(ov, su)[2].numpy()
    arg, best = c.argmax(-1), c.max(-1)
    amb = (c >= (best - 2 * e)[..., None]).sum(-1) > 1
    rs = np.take_along_axis(np.broadcast_to(r[:, None, :], c.shape), arg[..., None], 2)[..., 0]
    d = 2 - 2 * best * rs
    rmax, rmin = r.max(1)[:, None], r.min(1)[:, None]
    slack = np.where(amb, 2 * np.abs(best) * (rmax - rmin) + 6 * e * rmax, 2 * e * rs)
    assert (np.abs(d - dist) <= slack).all()
    assert ((arg == dist_arg(ov, su)) | amb).all()                      # an unambiguous argmax is the exact one
    dt = dist[np.arange(Q), np.arange(Q)]
    inband = np.abs(d - dt[None, :]) <= slack
    assert (((d <= dt[None, :]) == (dist <= dt[None, :])) | inband).all()
    if kind == "gauss":
        assert inband.mean() <= 0.01                                    # and the deferral stays sparse
    key = d - slack
    assert (key <= dist).all()
    order = np.argsort(key, axis=0, kind="stable")[:16]
    for q in range(Q):
        ex = dist[order[:, q], q]
        kth = np.sort(ex)[k - 1]
        if key[order[-1, q], q] > kth:                                   # proven: the list holds the whole top k
            assert np.array_equal(np.sort(ex)[:k], np.sort(dist[:, q])[:k])


def dist_arg(ov, su):
    return O.fused_fp64(ov, su)[1].numpy()


def test_release_library_has_no_debug_switches():
    """The shipped library must not read WITW_* debug variables (r1: a stray variable gave a faster, wrong kernel)."""
    from witw_b200 import _lib

    blob = open(_lib.LIB_PATH, "rb").read()
    for name in (b"WITW_SPEC_DEBUG", b"WITW_TC_FULL_B", b"WITW_TC_CG", b"WITW_POLAR_PW"):
        assert name not in blob, name
    for src in ("match_spec.cu", "match_tc.cu", "polar.cu"):
        text = open(os.path.join(ROOT, "witw_b200", "csrc", src)).read()
        for m in re.finditer(r"getenv", text):
            head = text[: m.start()]
            assert head.rfind("#ifdef WITW_DEBUG_HOOKS") > head.rfind("#endif"), src
