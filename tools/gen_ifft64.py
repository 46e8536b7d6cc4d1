#!/usr/bin/env python3
"""Generates witw_b200/csrc/ifft64_gen.cuh: a straight-line, register-resident inverse real FFT of 64 points.

Input: a Hermitian spectrum in the packed form of csrc/spectral.cu -- re[0] = P_0, im[0] = P_32 (both real),
(re[f], im[f]) = P_f for f = 1..31 -- already scaled by 1/64.  Output: x[s] = sum_{f=0}^{63} P_f e^{+2 pi i f s / 64},
s = 0..63 (P_{64-f} = conj P_f), i.e. numpy.fft.irfft of the unscaled spectrum.

Method: the 64 real outputs are the 32 complex points z[n] = x[2n] + i x[2n+1] of
    Z[k] = (P_k + conj P_{32-k}) + i w^k (P_k - conj P_{32-k}),   w = e^{2 pi i / 64},
followed by a 32-point radix-2 decimation-in-time inverse FFT with literal twiddles (trivial ones folded).
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


def brev5(i):
    return int("{:05b}".format(i)[::-1], 2)


def build(e, re, im):
    """re/im: lists of 32 input variable names.  Returns list of 64 output variable names."""
    z = [None] * 32
    # k = 0: (P0 + P32, P0 - P32)
    z[0] = (e.add(re[0], im[0]), e.sub(re[0], im[0]))
    # k = 16: (2 a16, -2 b16)
    z[16] = (e.mulc(re[16], 2.0), e.mulc(im[16], -2.0))
    for k in range(1, 16):
        kk = 32 - k
        ar = e.add(re[k], re[kk])
        ai = e.sub(im[k], im[kk])
        dr = e.sub(re[k], re[kk])
        di = e.add(im[k], im[kk])
        c, s = math.cos(2 * math.pi * k / 64), math.sin(2 * math.pi * k / 64)
        # T = D * w^k
        p = e.mulc(di, -s)
        tr = e.fmac(dr, c, p)            # dr*c - di*s
        q = e.mulc(di, c)
        ti = e.fmac(dr, s, q)            # dr*s + di*c
        # Z_k = A + iT = (ar - ti, ai + tr);  Z_{32-k} = conj(A) + i conj(T) = (ar + ti, tr - ai)
        z[k] = (e.sub(ar, ti), e.add(ai, tr))
        z[kk] = (e.add(ar, ti), e.sub(tr, ai))
    x = [z[brev5(i)] for i in range(32)]
    m = 2
    while m <= 32:
        half = m // 2
        for base in range(0, 32, m):
            for j in range(half):
                u, v = x[base + j], x[base + j + half]
                ang = 2 * math.pi * j / m
                c, s = math.cos(ang), math.sin(ang)
                if j == 0:
                    o1 = (e.add(u[0], v[0]), e.add(u[1], v[1]))
                    o2 = (e.sub(u[0], v[0]), e.sub(u[1], v[1]))
                elif 4 * j == m:  # w = i: t = (-v.im, v.re)
                    o1 = (e.sub(u[0], v[1]), e.add(u[1], v[0]))
                    o2 = (e.add(u[0], v[1]), e.sub(u[1], v[0]))
                elif abs(abs(c) - abs(s)) < 1e-12:
                    sg = 1.0 if c * s > 0 else -1.0
                    # t.re = c*(v.re - sg*v.im), t.im = c*(v.im + sg*v.re)
                    d = e.sub(v[0], v[1]) if sg > 0 else e.add(v[0], v[1])
                    f = e.add(v[1], v[0]) if sg > 0 else e.sub(v[1], v[0])
                    o1 = (e.fmac(d, c, u[0]), e.fmac(f, c, u[1]))
                    o2 = (e.fmac(d, -c, u[0]), e.fmac(f, -c, u[1]))
                else:
                    p = e.mulc(v[1], -s)
                    tr = e.fmac(v[0], c, p)
                    q = e.mulc(v[1], c)
                    ti = e.fmac(v[0], s, q)
                    o1 = (e.add(u[0], tr), e.add(u[1], ti))
                    o2 = (e.sub(u[0], tr), e.sub(u[1], ti))
                x[base + j], x[base + j + half] = o1, o2
        m *= 2
    out = []
    for n in range(32):
        out += [x[n][0], x[n][1]]
    return out


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
