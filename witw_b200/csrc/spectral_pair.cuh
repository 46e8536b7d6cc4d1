// Warp-collective exact fp32 evaluation of one (gallery item, query) pair from packed azimuth spectra (layout:
// spectral.cu): the building block of witw_match_pairs_spec_f32, the per-query finish and the column kernel (finish.cu).
// All of them call this one function, so the same pair always gets the same bits -- the rank rule compares distances
// that come from different launches (cvig_fov.py:552: d[g] <= d[idx]).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace witw {

// (cos, sin)(2 pi m / 64), m = 0..63, into shared memory; call with at least 64 threads, then __syncthreads().
__device__ __forceinline__ void spectral_twiddles(float2* tw) {
  if (threadIdx.x < 64) {
    float s, c;
    sincospif((float)threadIdx.x / 32.0f, &s, &c);
    tw[threadIdx.x] = make_float2(c, s);
  }
}

struct PairMax { float best; int arg; };

// go / qo: the item's / the query's spectrum, already offset by the lane (lane = frequency slot); rows are 32 float2 apart.
// qo may point to shared memory.  P_f = sum_ch O_f conj(S_f); then every lane evaluates the inverse transform at shifts
// `lane` and `lane + 32` from the 32 broadcast P_f; warp argmax (first maximum, NaN is the maximum -- torch.argmax).
// Every lane returns the same (max correlation, shift).
__device__ __forceinline__ PairMax spectral_pair_eval(const float2* __restrict__ go, const float2* qo, int CH, const float2* tw, int lane) {
  float a = 0.f, b = 0.f, c = 0.f;  // sum o.x s.x, sum o.y s.y, sum (o.y s.x - o.x s.y)
#pragma unroll 8
  for (int ch = 0; ch < CH; ++ch) {
    const float2 o = __ldg(go + ch * 32), s = qo[ch * 32];
    a = fmaf(o.x, s.x, a);
    b = fmaf(o.y, s.y, b);
    c = fmaf(o.y, s.x, c);
    c = fmaf(-o.x, s.y, c);
  }
  const float p0 = __shfl_sync(0xffffffffu, a, 0), p32 = __shfl_sync(0xffffffffu, b, 0);
  const float re = a + b, im = c;
  const float base = p0 + ((lane & 1) ? -p32 : p32);
  float lo = 0.f, hi = 0.f;
#pragma unroll
  for (int f = 1; f < 32; ++f) {
    const float fr = __shfl_sync(0xffffffffu, re, f), fi = __shfl_sync(0xffffffffu, im, f);
    const float2 t = tw[(f * lane) & 63];
    const float term = fr * t.x - fi * t.y;
    lo += term;
    hi += (f & 1) ? -term : term;
  }
  const float c_lo = (base + 2.0f * lo) * (1.0f / 64.0f), c_hi = (base + 2.0f * hi) * (1.0f / 64.0f);
  PairMax r;
  r.best = c_lo;
  r.arg = lane;
  if (c_hi > r.best || (c_hi != c_hi && r.best == r.best)) { r.best = c_hi; r.arg = lane + 32; }
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, r.best, m);
    const int oa = __shfl_xor_sync(0xffffffffu, r.arg, m);
    const bool take = (ob > r.best) || (ob != ob && r.best == r.best) || (ob == r.best && oa < r.arg) || (ob != ob && r.best != r.best && oa < r.arg);
    if (take) { r.best = ob; r.arg = oa; }
  }
  return r;
}

}  // namespace witw
