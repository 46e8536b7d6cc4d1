// Exact fp32 orientation-searched distance of explicit (gallery, query) pairs in the azimuth-frequency
// domain.  Used for the "exact finish" of the tensor-core sweep (true-match thresholds, re-checks of
// near-threshold rank decisions, re-ranking of top-k candidates): the same quantities as
//   corr[s] = sum_{ch,k<sw} ov[ch,(s+k)%64] * su[ch,k]        model/cvig_fov.py:297-312
//   ori     = first argmax_s corr[s]                           model/cvig_fov.py:313
//   dist    = 2*(1 - corr[ori]/(||crop||*||su||))              model/cvig_fov.py:318-363
// but with the circular correlation evaluated through the correlation theorem.  With O = rfft(ov row),
// S = rfft(zero-padded su row) (64 columns -> 33 bins),
//   corr = irfft( sum_ch O_ch * conj(S_ch) )
// costs 64 rows x 32 complex MACs + one 64-point inverse per pair (~13k MACs) instead of 262k, and in
// fp32 it is closer to the float64 value than a 4096-term fp32 dot product (measured: 1.5e-7 vs 3.4e-7
// of the norm product).  The whole-gallery sweep stays a dense tensor-core contraction (match_tc.cu);
// this file only serves the few pairs whose result must not carry bf16 rounding.
//
// Packed spectrum of one 64-column row: 32 float2 slots; slot f (1..31) = (Re X_f, Im X_f),
// slot 0 = (X_0, X_32) (both real).
#include <algorithm>

#include "common.cuh"
#include "row_fft.cuh"
#include "spectral_pair.cuh"

namespace witw {

// One warp per row (RowFft, row_fft.cuh).
__global__ void __launch_bounds__(256)
spectral_rows_kernel(const float* __restrict__ x, int64_t n_rows, int row_len, float* __restrict__ spec) {
  const int lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * 8;
  RowFft fft;
  fft.init(lane);
  const bool full = row_len == 64 && ((uintptr_t)x & 7) == 0;
  for (int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); row < n_rows; row += stride) {
    float2 z;
    if (full) {
      z = __ldg(reinterpret_cast<const float2*>(x + row * 64) + lane);
    } else {  // query rows narrower than the gallery: zero-padded to 64 columns
      const float* r = x + row * row_len;
      const int j = 2 * lane;
      z.x = j < row_len ? r[j] : 0.f;
      z.y = j + 1 < row_len ? r[j + 1] : 0.f;
    }
    reinterpret_cast<float2*>(spec + row * 64)[lane] = fft.run(z, lane);
  }
}

// One warp per pair (spectral_pair_eval); distance from the fp32 norm tables of the prep kernels.
__global__ void __launch_bounds__(256)
spectral_pairs_kernel(const float2* __restrict__ gal_spec, const float* __restrict__ crop_inv_norm,
                      const float2* __restrict__ qry_spec, const float* __restrict__ q_inv_norm,
                      const int64_t* __restrict__ pair_g, const int64_t* __restrict__ pair_q, int64_t n_pairs,
                      const int32_t* __restrict__ n_pairs_dev, int CH, float* __restrict__ dist, int64_t* __restrict__ ori) {
  __shared__ float2 tw[64];
  spectral_twiddles(tw);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t p = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  int64_t n = n_pairs;
  if (n_pairs_dev != nullptr) n = min(n, (int64_t)*n_pairs_dev);
  if (p >= n) return;
  const int64_t g = pair_g[p], q = pair_q[p];
  const PairMax r = spectral_pair_eval(gal_spec + g * CH * 32 + lane, qry_spec + q * CH * 32 + lane, CH, tw, lane);
  if (lane == 0) {
    if (dist != nullptr) dist[p] = 2.0f * (1.0f - r.best * crop_inv_norm[g * 64 + r.arg] * q_inv_norm[q]);
    if (ori != nullptr) ori[p] = r.arg;
  }
}

int launch_pairs_spec(const float* gal_spec, const float* crop_inv_norm, const float* qry_spec, const float* q_inv_norm,
                      const int64_t* pair_g, const int64_t* pair_q, int64_t n_pairs, const int32_t* n_pairs_dev, int CH,
                      float* dist, int64_t* ori, witw_stream_t stream) {
  WITW_REQUIRE(n_pairs > 0 && n_pairs < (1ll << 33), WITW_ERR_INVALID, "spectral pairs: bad pair count %lld", (long long)n_pairs);
  WITW_REQUIRE((((uintptr_t)gal_spec | (uintptr_t)qry_spec) & 7) == 0, WITW_ERR_INVALID, "spectral pairs: spectra must be 8-byte aligned");
  spectral_pairs_kernel<<<(unsigned)ceil_div<int64_t>(n_pairs, 8), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float2*>(gal_spec), crop_inv_norm, reinterpret_cast<const float2*>(qry_spec), q_inv_norm, pair_g, pair_q,
      n_pairs, n_pairs_dev, CH, dist, ori);
  WITW_LAUNCH_CHECK();
  return WITW_OK;
}

}  // namespace witw

using namespace witw;

extern "C" int witw_spectral_rows_f32(const float* x, int64_t n_rows, int row_len, float* spec, witw_stream_t stream) {
  WITW_REQUIRE(n_rows >= 0 && row_len >= 1 && row_len <= 64, WITW_ERR_INVALID, "witw_spectral_rows_f32: rows of 1..64 columns (got %d)", row_len);
  if (n_rows == 0) return WITW_OK;
  WITW_REQUIRE(x && spec, WITW_ERR_INVALID, "witw_spectral_rows_f32: null pointer");
  WITW_REQUIRE(((uintptr_t)spec & 7) == 0, WITW_ERR_INVALID, "witw_spectral_rows_f32: output must be 8-byte aligned");
  const int64_t blocks = std::min<int64_t>(ceil_div<int64_t>(n_rows, 8), (int64_t)sm_count() * 32);
  spectral_rows_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(x, n_rows, row_len, spec);
  WITW_LAUNCH_CHECK();
  return WITW_OK;
}

extern "C" int witw_match_pairs_spec_f32(const float* gal_spec, const float* crop_inv_norm, const float* qry_spec,
                                         const float* q_inv_norm, const int64_t* pair_g, const int64_t* pair_q, int64_t n_pairs,
                                         int CH, float* dist, int64_t* ori, witw_stream_t stream) {
  WITW_REQUIRE(CH > 0 && n_pairs >= 0, WITW_ERR_INVALID, "witw_match_pairs_spec_f32: bad shape");
  if (n_pairs == 0) return WITW_OK;
  WITW_REQUIRE(gal_spec && crop_inv_norm && qry_spec && q_inv_norm && pair_g && pair_q, WITW_ERR_INVALID, "witw_match_pairs_spec_f32: null pointer");
  return launch_pairs_spec(gal_spec, crop_inv_norm, qry_spec, q_inv_norm, pair_g, pair_q, n_pairs, nullptr, CH, dist, ori, stream);
}
