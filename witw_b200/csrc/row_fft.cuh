// Warp-level 64-point real FFT of one feature row (one complex point per lane), shared by spectral.cu and match_spec.cu.
#pragma once
#include <cuda_runtime.h>

namespace witw {

__device__ __forceinline__ int brev5(int n) { return (int)(__brev((unsigned)n) >> 27); }

// The 64 real samples of a row are 32 complex points z[n] = x[2n] + i x[2n+1], one per lane; a 5-stage
// decimation-in-frequency FFT over the lanes (shuffles), then the real-input split.  Result: lane f holds the packed
// spectrum slot f -- (Re X_f, Im X_f) for f = 1..31, (X_0, X_32) for f = 0.
struct RowFft {
  float2 tw[5];  // stage twiddles exp(-2 pi i (lane mod h) / 2h), h = 16, 8, 4, 2, 1
  float ws, wc;  // exp(-2 pi i lane / 64) = (wc, -ws)
  int src_k, src_m;
  __device__ __forceinline__ void init(int lane) {
#pragma unroll
    for (int st = 0; st < 5; ++st) {
      const int h = 16 >> st;
      float s, c;
      sincospif((float)(lane & (h - 1)) / (float)h, &s, &c);
      tw[st] = make_float2(c, -s);
    }
    sincospif((float)lane / 32.0f, &ws, &wc);
    src_k = brev5(lane);
    src_m = brev5((32 - lane) & 31);
  }
  __device__ __forceinline__ float2 run(float2 z, int lane) const {
#pragma unroll
    for (int st = 0; st < 5; ++st) {
      const int h = 16 >> st;
      const float ox = __shfl_xor_sync(0xffffffffu, z.x, h), oy = __shfl_xor_sync(0xffffffffu, z.y, h);
      if (lane & h) {
        const float dx = ox - z.x, dy = oy - z.y;
        z.x = dx * tw[st].x - dy * tw[st].y;
        z.y = dx * tw[st].y + dy * tw[st].x;
      } else {
        z.x += ox;
        z.y += oy;
      }
    }
    // lane n now holds Z[bitrev(n)].  X_k = E_k + W^k O_k with E = (Z_k + conj Z_{32-k})/2, O = -i (Z_k - conj Z_{32-k})/2
    const float ax = __shfl_sync(0xffffffffu, z.x, src_k), ay = __shfl_sync(0xffffffffu, z.y, src_k);
    const float bx = __shfl_sync(0xffffffffu, z.x, src_m), by = -__shfl_sync(0xffffffffu, z.y, src_m);
    const float ex = 0.5f * (ax + bx), ey = 0.5f * (ay + by);
    const float ox = 0.5f * (ay - by), oy = -0.5f * (ax - bx);
    float2 X;
    X.x = ex + (wc * ox + ws * oy);
    X.y = ey + (wc * oy - ws * ox);
    if (lane == 0) { X.x = ax + ay; X.y = ax - ay; }  // X_0 and X_32
    return X;
  }
};

}  // namespace witw
