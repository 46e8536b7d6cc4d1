#!/usr/bin/env python3
"""Generates witw_b200/csrc/ifft64_gen.cuh: a straight-line, register-resident inverse real FFT of 64 points.

Input: a Hermitian spectrum in the packed form of csrc/spectral.cu -- re[0] = P_0, im[0] = P_32 (both real),
(re[f], im[f]) = P_f for f = 1..31 -- already scaled by 1/64.  Output: x[s] = sum_{f=0}^{63} P_f e^{+2 pi i f s / 64},
s = 0..63 (P_{64-f} = conj P_f), i.e. numpy.fft.irfft of the unscaled spectrum.

Method: real-data split radix (hermitian_to_real below): 487 operations; the first version (the half-length complex sequence
z[n] = x[2n] + i x[2n+1] through a 32-point radix-2 FFT) needed 612.
The script checks the generated operation list in float32 against numpy.fft.irfft before writing the file.

Usage: python tools/gen_ifft64.py   (rewrites the header in place; the header is committed)
"""
import math
import os

import numpy as np

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "witw_b200", "csrc", "ifft64_gen.cuh")


class Emitter(object):
    def __init__(self):
        self.lines = []
        self.vals = {}
        self.n = 0
        self.ops = 0

    def _new(self, expr, val):
        name = "t%d" % self.n
        self.n += 1
        self.ops += 1
        self.lines.append("  const float %s = %s;" % (name, expr))
        self.vals[name] = np.float32(val)
        return name

    @staticmethod
    def lit(c):
        return "%.9ef" % c

    def inp(self, name, val):
        self.vals[name] = np.float32(val)
        return name

    def add(self, a, b):
        return self._new("%s + %s" % (a, b), self.vals[a] + self.vals[b])

    def sub(self, a, b):
        return self._new("%s - %s" % (a, b), self.vals[a] - self.vals[b])

    def mulc(self, a, c):
        return self._new("%s * %s" % (a, self.lit(c)), self.vals[a] * np.float32(c))

    def fmac(self, a, c, b):  # a*c + b, one rounding
        v = np.float32(np.float64(self.vals[a]) * np.float64(np.float32(c)) + np.float64(self.vals[b]))
        return self._new("fmaf(%s, %s, %s)" % (a, self.lit(c), b), v)



class PackedEmitter(Emitter):
    """Same operation list on float2 operands: two independent transforms in the .x and .y halves, one packed
    FADD2 / FMUL2 / FFMA2 (sm_100 f32x2) instruction per operation."""

    def _new(self, expr, val):
        name = "t%d" % self.n
        self.n += 1
        self.ops += 1
        self.lines.append("  const float2 %s = %s;" % (name, expr))
        self.vals[name] = np.float32(val)
        return name

    def add(self, a, b):
        return self._new("__fadd2_rn(%s, %s)" % (a, b), self.vals[a] + self.vals[b])

    def sub(self, a, b):
        return self._new("__ffma2_rn(%s, neg1, %s)" % (b, a), self.vals[a] - self.vals[b])

    def mulc(self, a, c):
        return self._new("__fmul2_rn(%s, make_float2(%s, %s))" % (a, self.lit(c), self.lit(c)), self.vals[a] * np.float32(c))

    def fmac(self, a, c, b):
        v = np.float32(np.float64(self.vals[a]) * np.float64(np.float32(c)) + np.float64(self.vals[b]))
        return self._new("__ffma2_rn(%s, make_float2(%s, %s), %s)" % (a, self.lit(c), self.lit(c), b), v)


def cmul(e, v, ang):
    """(v.re + i v.im) e^{i ang}: two multiplications and two fused multiply-adds."""
    c, s = math.cos(ang), math.sin(ang)
    p = e.mulc(v[1], -s)
    re = e.fmac(v[0], c, p)              # v.re c - v.im s
    q = e.mulc(v[1], c)
    im = e.fmac(v[0], s, q)              # v.re s + v.im c
    return (re, im)


def hermitian_to_real(e, n, X):
    """Unnormalised inverse DFT x[m] = sum_{k<n} X[k] e^{+2 pi i k m / n} of a Hermitian spectrum given as X[0..n/2] (pairs of
    variable names; X[0] and X[n/2] are real: their second entry is None).  Real-data split radix, decimation in time of the
    output: the even outputs are the half-size transform of X[k] + X[k + n/2], the outputs 4m+1 and 4m+3 the quarter-size
    transforms of w^k (U + iV) and w^{3k} (U - iV), U = X[k] - X[k + n/2], V = X[k + n/4] - X[k + 3n/4], w = e^{2 pi i / n};
    X[n - j] = conj X[j] brings every index back into 0..n/2.  n (5 n / 2 - 6 at the top level) fewer operations per level than
    transforming the half-length complex sequence."""
    if n == 1:
        return [X[0][0]]
    if n == 2:
        return [e.add(X[0][0], X[1][0]), e.sub(X[0][0], X[1][0])]
    h, q = n // 2, n // 4
    E = [None] * (q + 1)
    E[0] = (e.add(X[0][0], X[h][0]), None)
    for k in range(1, q):
        a, b = X[k], X[h - k]
        E[k] = (e.add(a[0], b[0]), e.sub(a[1], b[1]))
    E[q] = (e.mulc(X[q][0], 2.0), None)
    even = hermitian_to_real(e, h, E)
    # odd outputs
    Y = [None] * (q // 2 + 1)
    Z = [None] * (q // 2 + 1)
    u0 = e.sub(X[0][0], X[h][0])                      # U real, iV = -2 Im X[n/4]
    Y[0] = (e.fmac(X[q][1], -2.0, u0), None)
    Z[0] = (e.fmac(X[q][1], 2.0, u0), None)
    for k in range(1, q // 2):
        a, b, c, d = X[k], X[h - k], X[k + q], X[q - k]
        ur, ui = e.sub(a[0], b[0]), e.add(a[1], b[1])     # U = X[k] - conj X[n/2 - k]
        vr, vi = e.sub(c[0], d[0]), e.add(c[1], d[1])     # V = X[k + n/4] - conj X[n/4 - k]
        s_ = (e.sub(ur, vi), e.add(ui, vr))               # U + iV
        t_ = (e.add(ur, vi), e.sub(ui, vr))               # U - iV
        Y[k] = cmul(e, s_, 2 * math.pi * k / n)
        Z[k] = cmul(e, t_, 2 * math.pi * 3 * k / n)
    if q >= 2:                                            # k = n/8: V = -conj U, both results are real
        k = q // 2
        a, b = X[k], X[h - k]
        ur, ui = e.sub(a[0], b[0]), e.add(a[1], b[1])
        Y[k] = (e.mulc(e.sub(ur, ui), math.sqrt(2.0)), None)
        Z[k] = (e.mulc(e.add(ur, ui), -math.sqrt(2.0)), None)
    o1 = hermitian_to_real(e, q, Y)
    o3 = hermitian_to_real(e, q, Z)
    out = [None] * n
    for m in range(h):
        out[2 * m] = even[m]
    for m in range(q):
        out[4 * m + 1] = o1[m]
        out[4 * m + 3] = o3[m]
    return out


def build(e, re, im):
    """re/im: lists of 32 input variable names.  Returns list of 64 output variable names."""
    X = [(re[0], None)] + [(re[f], im[f]) for f in range(1, 32)] + [(im[0], None)]
    return hermitian_to_real(e, 64, X)


def main():
    rng = np.random.default_rng(7)
    worst = 0.0
    bodies = {}
    for cls in (Emitter, PackedEmitter):
        for trial in range(4):
            sig = rng.standard_normal(64)
            spec = np.fft.rfft(sig) / 64.0
            re_v = [spec[0].real] + [spec[f].real for f in range(1, 32)]
            im_v = [spec[32].real] + [spec[f].imag for f in range(1, 32)]
            e = cls()
            re = [e.inp("re[%d]" % f, re_v[f]) for f in range(32)]
            im = [e.inp("im[%d]" % f, im_v[f]) for f in range(32)]
            out = build(e, re, im)
            got = np.array([e.vals[o] for o in out], dtype=np.float64)
            worst = max(worst, np.abs(got - sig).max() / np.abs(sig).max())
        assert worst < 2e-6, worst
        bodies[cls] = ("\n".join(e.lines), "\n".join("  x[%d] = %s;" % (s, o) for s, o in enumerate(out)))
    text = (
        "// GENERATED by tools/gen_ifft64.py -- do not edit.  %d fp32 operations, checked against numpy.fft.irfft\n"
        "// (worst relative error %.1e in float32).\n"
        "// x[s] = sum_f P_f e^{+2 pi i f s/64} for the packed Hermitian spectrum re[0]=P_0, im[0]=P_32, (re[f],im[f])=P_f.\n"
        "#pragma once\n\nnamespace witw {\n\n"
        "__device__ __forceinline__ void ifft64_hermitian(const float (&re)[32], const float (&im)[32], float (&x)[64]) {\n"
        "%s\n%s\n}\n\n"
        "// Two transforms at once, one in each half of the float2 operands (FADD2 / FMUL2 / FFMA2).\n"
        "__device__ __forceinline__ void ifft64_hermitian_x2(const float2 (&re)[32], const float2 (&im)[32], float2 (&x)[64]) {\n"
        "  const float2 neg1 = make_float2(-1.0f, -1.0f);\n"
        "%s\n%s\n}\n\n}  // namespace witw\n" % ((e.ops, worst) + bodies[Emitter] + bodies[PackedEmitter])
    )
    with open(OUT, "w") as f:
        f.write(text)
    print("wrote %s: %d ops, worst rel err %.2e" % (os.path.normpath(OUT), e.ops, worst))


if __name__ == "__main__":
    main()
