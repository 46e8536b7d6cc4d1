// Backward kernels of crop_overhead / l2_distance so the reference's train() (model/cvig_fov.py:447-460:
// encoders -> correlation -> crop_overhead -> l2_distance -> triplet_loss -> backward) runs on the rebound names.
// correlation() itself is not differentiable (argmax, cvig_fov.py:313).  Training batches are small
// (G = Q = 64), so these are plain, deterministic (atomic-free) kernels.
//
//   crop[g,q,ch,k] = ov[g,ch,(k+ori[g,q])%W]                      =>  d ov[g,ch,j] = sum_q [k=(j-ori)%W < sw] d crop[g,q,ch,k]
//   dist[g,q] = 2*(1 - <o,s>/(|o||s|)),  o = crop[g,q,:], s = su[q,:]
//     d o = gd * (-2) * ( s/(|o||s|) - <o,s> o/(|o|^3 |s|) )
//     d s = sum_g gd * (-2) * ( o/(|o||s|) - <o,s> s/(|o||s|^3) )
#include "common.cuh"

namespace witw {

__global__ void __launch_bounds__(256)
crop_backward_kernel(const float* __restrict__ grad_out, const int64_t* __restrict__ ori, float* __restrict__ grad_ov, int64_t G,
                     int64_t Q, int CH, int W, int sw) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // over [G, CH, W]
  if (i >= G * CH * W) return;
  const int j = (int)(i % W);
  const int ch = (int)((i / W) % CH);
  const int64_t g = i / ((int64_t)W * CH);
  float acc = 0.f;
  for (int64_t q = 0; q < Q; ++q) {
    int sh = (int)(ori[g * Q + q] % W);
    if (sh < 0) sh += W;
    int k = j - sh;
    if (k < 0) k += W;
    if (k < sw) acc += grad_out[((g * Q + q) * CH + ch) * sw + k];
  }
  grad_ov[i] = acc;
}

__global__ void __launch_bounds__(256)
l2_backward_pairs_kernel(const float* __restrict__ crop, const float* __restrict__ su, const float* __restrict__ grad_dist,
                         float* __restrict__ grad_crop, float* __restrict__ coef, int64_t G, int64_t Q, int64_t K) {
  // one warp per (g, q): the pair's scalars, the crop gradient, and the two coefficients the query gradient needs
  const int64_t gq = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (gq >= G * Q) return;
  const int lane = threadIdx.x & 31;
  const int64_t q = gq % Q;
  const float* o = crop + gq * K;
  const float* s = su + q * K;
  float dot = 0.f, oe = 0.f, se = 0.f;
  for (int64_t i = lane; i < K; i += 32) {
    const float a = o[i], b = s[i];
    dot = fmaf(a, b, dot);
    oe = fmaf(a, a, oe);
    se = fmaf(b, b, se);
  }
  for (int m = 16; m > 0; m >>= 1) {
    dot += __shfl_xor_sync(0xffffffffu, dot, m);
    oe += __shfl_xor_sync(0xffffffffu, oe, m);
    se += __shfl_xor_sync(0xffffffffu, se, m);
  }
  const float no = sqrtf(oe), ns = sqrtf(se), gd = grad_dist[gq];
  const float a_s = -2.0f * gd / (no * ns);              // multiplies s in d o, and o in d s
  const float b_o = 2.0f * gd * dot / (no * oe * ns);    // multiplies o in d o
  const float c_s = 2.0f * gd * dot / (no * ns * se);    // multiplies s in d s
  if (grad_crop != nullptr)
    for (int64_t i = lane; i < K; i += 32) grad_crop[gq * K + i] = fmaf(a_s, s[i], b_o * o[i]);
  if (coef != nullptr && lane == 0) { coef[2 * gq] = a_s; coef[2 * gq + 1] = c_s; }
}

__global__ void __launch_bounds__(256)
l2_backward_query_kernel(const float* __restrict__ crop, const float* __restrict__ su, const float* __restrict__ coef,
                         float* __restrict__ grad_su, int64_t G, int64_t Q, int64_t K) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // over [Q, K]
  if (i >= Q * K) return;
  const int64_t q = i / K, k = i - q * K;
  float acc = 0.f, csum = 0.f;
  for (int64_t g = 0; g < G; ++g) {
    const int64_t gq = g * Q + q;
    acc = fmaf(coef[2 * gq], crop[gq * K + k], acc);
    csum += coef[2 * gq + 1];
  }
  grad_su[i] = fmaf(csum, su[i], acc);
}

}  // namespace witw

using namespace witw;

extern "C" int witw_crop_backward_f32(const float* grad_out, const int64_t* ori, float* grad_ov, int64_t G, int64_t Q, int CH, int W,
                                      int sw, witw_stream_t stream) {
  WITW_REQUIRE(G >= 0 && Q >= 0 && CH > 0 && W > 0 && sw > 0 && sw <= W, WITW_ERR_INVALID, "witw_crop_backward_f32: bad shape");
  if (G == 0) return WITW_OK;
  WITW_REQUIRE(grad_ov && (Q == 0 || (grad_out && ori)), WITW_ERR_INVALID, "witw_crop_backward_f32: null pointer");
  const int64_t n = G * CH * W;
  WITW_REQUIRE(ceil_div<int64_t>(n, 256) < (1ll << 31), WITW_ERR_UNSUPPORTED, "witw_crop_backward_f32: gallery too large");
  crop_backward_kernel<<<(unsigned)ceil_div<int64_t>(n, 256), 256, 0, as_stream(stream)>>>(grad_out, ori, grad_ov, G, Q, CH, W, sw);
  WITW_LAUNCH_CHECK();
  return WITW_OK;
}

extern "C" int witw_l2_distance_backward_f32(const float* crop, const float* su, const float* grad_dist, float* grad_crop, float* grad_su,
                                             float* coef_scratch, int64_t G, int64_t Q, int64_t K, witw_stream_t stream) {
  WITW_REQUIRE(G >= 0 && Q >= 0 && K > 0, WITW_ERR_INVALID, "witw_l2_distance_backward_f32: bad shape");
  if (Q == 0) return WITW_OK;
  WITW_REQUIRE(su && (G == 0 || (crop && grad_dist)), WITW_ERR_INVALID, "witw_l2_distance_backward_f32: null pointer");
  WITW_REQUIRE(!grad_su || coef_scratch, WITW_ERR_INVALID, "witw_l2_distance_backward_f32: grad_su needs the [G,Q,2] scratch");
  if (G > 0) {
    const int64_t blocks = ceil_div<int64_t>(G * Q, 8);
    WITW_REQUIRE(blocks < (1ll << 31), WITW_ERR_UNSUPPORTED, "witw_l2_distance_backward_f32: too many pairs");
    l2_backward_pairs_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(crop, su, grad_dist, grad_crop, coef_scratch, G, Q, K);
    WITW_LAUNCH_CHECK();
  }
  if (grad_su != nullptr) {
    l2_backward_query_kernel<<<(unsigned)ceil_div<int64_t>(Q * K, 256), 256, 0, as_stream(stream)>>>(crop, su, coef_scratch, grad_su, G, Q, K);
    WITW_LAUNCH_CHECK();
  }
  return WITW_OK;
}
