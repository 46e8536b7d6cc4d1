"""CPU check of the arithmetic the spectral sweep (csrc/match_spec.cu) is built on: the operand layout of its header
comment, one GEMM per packed frequency slot, and the generated inverse FFT (tools/gen_ifft64.py) reproduce the reference's
circular cross-correlation (model/cvig_fov.py:297-312, restated in oracle.witw_oracle.fused_fp64)."""
import os
import sys

import numpy as np
import torch

from oracle import witw_oracle as O

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import gen_ifft64  # noqa: E402


def _generated_ifft(re, im):
    e = gen_ifft64.Emitter()
    r = [e.inp("re[%d]" % f, re[f]) for f in range(32)]
    i = [e.inp("im[%d]" % f, im[f]) for f in range(32)]
    return np.array([e.vals[o] for o in gen_ifft64.build(e, r, i)], dtype=np.float64)


def test_generated_ifft_matches_numpy():
    rng = np.random.default_rng(0)
    for _ in range(3):
        sig = rng.standard_normal(64)
        spec = np.fft.rfft(sig) / 64.0
        re = np.concatenate([[spec[0].real], spec[1:32].real])
        im = np.concatenate([[spec[32].real], spec[1:32].imag])
        assert np.abs(_generated_ifft(re, im) - sig).max() <= 2e-6 * np.abs(sig).max()


def test_generated_header_is_current():
    """witw_b200/csrc/ifft64_gen.cuh is what tools/gen_ifft64.py writes."""
    path = os.path.normpath(gen_ifft64.OUT)
    before = open(path).read()
    gen_ifft64.main()
    assert open(path).read() == before


def test_slot_gemm_layout_reproduces_correlation():
    G, Q, fov = 5, 3, 90
    ov, su, _ = O.synth_features(G, Q, fov=fov, noise=1.0, seed=9)
    sw = su.shape[3]
    corr = O.fused_fp64(ov, su)[0].numpy()                                   # [G,Q,64]
    ovr = ov.double().reshape(G, 64, 64).numpy()                             # [item][feature row][column]
    sur = np.zeros((Q, 64, 64))
    sur[:, :, :sw] = su.double().reshape(Q, 64, sw).numpy()
    So = np.fft.rfft(ovr, axis=2)                                            # [G,64,33]
    Sq = np.fft.rfft(sur, axis=2) / 64.0
    # A[q][slot] = [Re S | Im S] (slot 0: [S_0 | S_32]);  B[item][slot][c]: c=0 [Re O | Im O], c=1 [Im O | -Re O]
    A = np.zeros((Q, 32, 128))
    B = np.zeros((G, 32, 2, 128))
    for f in range(1, 32):
        A[:, f, :64], A[:, f, 64:] = Sq[:, :, f].real, Sq[:, :, f].imag
        B[:, f, 0, :64], B[:, f, 0, 64:] = So[:, :, f].real, So[:, :, f].imag
        B[:, f, 1, :64], B[:, f, 1, 64:] = So[:, :, f].imag, -So[:, :, f].real
    A[:, 0, :64], A[:, 0, 64:] = Sq[:, :, 0].real, Sq[:, :, 32].real
    B[:, 0, 0, :64] = So[:, :, 0].real
    B[:, 0, 1, 64:] = So[:, :, 32].real
    D = np.einsum("qfk,gfck->gqfc", A, B)                                    # the 64 TMEM columns of every pair
    for g in range(G):
        for q in range(Q):
            x = _generated_ifft(D[g, q, :, 0], D[g, q, :, 1])
            assert np.abs(x - corr[g, q]).max() <= 3e-6 * np.abs(corr[g, q]).max()


def test_split_radix_recursion_at_every_size():
    """The real-data split-radix recursion behind the codelet (gen_ifft64.hermitian_to_real) against numpy's unnormalised inverse
    real FFT at every power of two it passes through, with its operation counts (the 64-point one is what the epilogue issues)."""
    rng = np.random.default_rng(3)
    ops = {}
    for n in (2, 4, 8, 16, 32, 64):
        sig = rng.standard_normal(n)
        spec = np.fft.rfft(sig) / n
        e = gen_ifft64.Emitter()
        X = []
        for k in range(n // 2 + 1):
            re = e.inp("re%d" % k, spec[k].real)
            im = None if k in (0, n // 2) else e.inp("im%d" % k, spec[k].imag)
            X.append((re, im))
        out = gen_ifft64.hermitian_to_real(e, n, X)
        got = np.array([e.vals[o] for o in out], dtype=np.float64)
        assert np.abs(got - sig).max() <= 2e-6 * np.abs(sig).max(), n
        ops[n] = e.ops
    assert ops[64] == 487 and ops[2] == 2 and all(ops[2 * n] > 2 * ops[n] for n in (2, 4, 8, 16, 32))
    assert ops[64] < 612          # the first codelet: a 32-point complex radix-2 transform of the folded spectrum
