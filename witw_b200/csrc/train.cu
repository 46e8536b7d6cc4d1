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

// ------------------------------------------------------------------------------------------
// The same backward without the [G,Q,CH,sw] crop (64 MB per 64 x 64 batch at 360 degrees in the reference): the pair
// scalars come straight from the rolled gallery rows, the two gradients from the features and three coefficients per pair.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
match_coef_kernel(const float* __restrict__ ov, const float* __restrict__ su, const int64_t* __restrict__ ori,
                  const float* __restrict__ grad_dist, float* __restrict__ coef, int64_t G, int64_t Q, int CH, int W, int sw) {
  const int64_t gq = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   // one warp per (g, q)
  if (gq >= G * Q) return;
  const int lane = threadIdx.x & 31;
  const int64_t g = gq / Q, q = gq - g * Q;
  int sh = (int)(ori[gq] % W);
  if (sh < 0) sh += W;
  const float* o = ov + g * (int64_t)CH * W;
  const float* s = su + q * (int64_t)CH * sw;
  float dot = 0.f, oe = 0.f, se = 0.f;
  for (int i = lane; i < CH * sw; i += 32) {
    const int ch = i / sw, k = i - ch * sw;
    int j = k + sh;
    if (j >= W) j -= W;
    const float a = o[ch * W + j], b = s[i];
    dot = fmaf(a, b, dot);
    oe = fmaf(a, a, oe);
    se = fmaf(b, b, se);
  }
  for (int m = 16; m > 0; m >>= 1) {
    dot += __shfl_xor_sync(0xffffffffu, dot, m);
    oe += __shfl_xor_sync(0xffffffffu, oe, m);
    se += __shfl_xor_sync(0xffffffffu, se, m);
  }
  if (lane == 0) {
    const float no = sqrtf(oe), ns = sqrtf(se), gd = grad_dist[gq];
    coef[3 * gq + 0] = -2.0f * gd / (no * ns);              // multiplies s in d o, and o in d s
    coef[3 * gq + 1] = 2.0f * gd * dot / (no * oe * ns);    // multiplies o in d o
    coef[3 * gq + 2] = 2.0f * gd * dot / (no * ns * se);    // multiplies s in d s
  }
}

__global__ void __launch_bounds__(256)
match_backward_ov_kernel(const float* __restrict__ ov, const float* __restrict__ su, const int64_t* __restrict__ ori,
                         const float* __restrict__ coef, float* __restrict__ grad_ov, int64_t G, int64_t Q, int CH, int W, int sw) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // over [G, CH, W]
  if (i >= G * CH * W) return;
  const int j = (int)(i % W);
  const int ch = (int)((i / W) % CH);
  const int64_t g = i / ((int64_t)W * CH);
  float acc = 0.f, bsum = 0.f;
  for (int64_t q = 0; q < Q; ++q) {
    const int64_t gq = g * Q + q;
    int sh = (int)(ori[gq] % W);
    if (sh < 0) sh += W;
    int k = j - sh;
    if (k < 0) k += W;
    if (k < sw) {
      acc = fmaf(coef[3 * gq], su[(q * CH + ch) * sw + k], acc);
      bsum += coef[3 * gq + 1];
    }
  }
  grad_ov[i] = fmaf(bsum, ov[i], acc);
}

__global__ void __launch_bounds__(256)
match_backward_su_kernel(const float* __restrict__ ov, const float* __restrict__ su, const int64_t* __restrict__ ori,
                         const float* __restrict__ coef, float* __restrict__ grad_su, int64_t G, int64_t Q, int CH, int W, int sw) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // over [Q, CH, sw]
  if (i >= Q * CH * sw) return;
  const int k = (int)(i % sw);
  const int ch = (int)((i / sw) % CH);
  const int64_t q = i / ((int64_t)sw * CH);
  float acc = 0.f, csum = 0.f;
  for (int64_t g = 0; g < G; ++g) {
    const int64_t gq = g * Q + q;
    int sh = (int)(ori[gq] % W);
    if (sh < 0) sh += W;
    int j = k + sh;
    if (j >= W) j -= W;
    acc = fmaf(coef[3 * gq], ov[(g * CH + ch) * W + j], acc);
    csum += coef[3 * gq + 2];
  }
  grad_su[i] = fmaf(csum, su[i], acc);
}

// triplet_loss (cvig_fov.py:366-382) and its gradient in one CTA: batches are 64 x 64.
//   loss = ( sum_ij log(1 + exp(alpha (d_jj - d_ij))) + sum_ij log(1 + exp(alpha (d_ii - d_ij))) ) / (2 N (N - 1))
__global__ void __launch_bounds__(1024)
triplet_loss_kernel(const float* __restrict__ d, int N, float alpha, float* __restrict__ loss, float* __restrict__ grad) {
  extern __shared__ float sh[];            // [N] off-diagonal column sums of A, [N] off-diagonal row sums of B, [32] reduction
  float* col_a = sh;
  float* row_b = sh + N;
  float* red = sh + 2 * N;
  const int tid = threadIdx.x;
  const float z = 2.0f * (float)N * (float)(N - 1);
  for (int i = tid; i < 2 * N; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  float part = 0.f;
  // column j owned by one thread at a time keeps the sums deterministic: thread t handles columns / rows t, t + T, ...
  for (int j = tid; j < N; j += blockDim.x) {
    const float mj = d[(size_t)j * N + j];
    float ca = 0.f, rb = 0.f;
    for (int i = 0; i < N; ++i) {
      const float xa = alpha * (mj - d[(size_t)i * N + j]);       // surface -> overhead term, element (i, j)
      const float xb = alpha * (mj - d[(size_t)j * N + i]);       // overhead -> surface term, element (j, i)
      part += logf(1.0f + expf(xa)) + logf(1.0f + expf(xb));
      if (i != j) {       // the diagonal's own two halves cancel in the gradient; summing them here would drown the small terms
        ca += 1.0f / (1.0f + expf(-xa));
        rb += 1.0f / (1.0f + expf(-xb));
      }
    }
    col_a[j] = ca;
    row_b[j] = rb;
  }
  for (int m = 16; m > 0; m >>= 1) part += __shfl_xor_sync(0xffffffffu, part, m);
  if ((tid & 31) == 0) red[tid >> 5] = part;
  __syncthreads();
  if (tid == 0) {
    float total = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) total += red[w];
    *loss = total / z;
  }
  if (grad == nullptr) return;
  for (int e = tid; e < N * N; e += blockDim.x) {
    const int i = e / N, j = e - i * N;
    const float dij = d[e];
    const float a = 1.0f / (1.0f + expf(-alpha * (d[(size_t)j * N + j] - dij)));
    const float b = 1.0f / (1.0f + expf(-alpha * (d[(size_t)i * N + i] - dij)));
    const float gr = (i == j) ? alpha * (col_a[i] + row_b[i]) : -alpha * (a + b);
    grad[e] = gr / z;
  }
}

}  // namespace witw

using namespace witw;

extern "C" int witw_match_backward_f32(const float* ov, const float* su, const int64_t* ori, const float* grad_dist, float* grad_ov,
                                       float* grad_su, float* coef_scratch, int64_t G, int64_t Q, int CH, int W, int sw,
                                       witw_stream_t stream) {
  WITW_REQUIRE(G >= 0 && Q >= 0 && CH > 0 && W > 0 && sw > 0 && sw <= W, WITW_ERR_INVALID, "witw_match_backward_f32: bad shape");
  cudaStream_t st = as_stream(stream);
  if (G == 0 || Q == 0) {       // no pairs: both gradients are zero
    if (grad_ov && G > 0) WITW_CUDA(cudaMemsetAsync(grad_ov, 0, sizeof(float) * (size_t)G * CH * W, st));
    if (grad_su && Q > 0) WITW_CUDA(cudaMemsetAsync(grad_su, 0, sizeof(float) * (size_t)Q * CH * sw, st));
    return WITW_OK;
  }
  WITW_REQUIRE(ov && su && ori && grad_dist && coef_scratch, WITW_ERR_INVALID, "witw_match_backward_f32: null pointer");
  const int64_t pair_blocks = ceil_div<int64_t>(G * Q, 8), ov_blocks = ceil_div<int64_t>(G * CH * W, 256),
                su_blocks = ceil_div<int64_t>(Q * CH * sw, 256);
  WITW_REQUIRE(pair_blocks < (1ll << 31) && ov_blocks < (1ll << 31) && su_blocks < (1ll << 31), WITW_ERR_UNSUPPORTED,
               "witw_match_backward_f32: problem too large");
  match_coef_kernel<<<(unsigned)pair_blocks, 256, 0, st>>>(ov, su, ori, grad_dist, coef_scratch, G, Q, CH, W, sw);
  WITW_LAUNCH_CHECK();
  if (grad_ov) {
    match_backward_ov_kernel<<<(unsigned)ov_blocks, 256, 0, st>>>(ov, su, ori, coef_scratch, grad_ov, G, Q, CH, W, sw);
    WITW_LAUNCH_CHECK();
  }
  if (grad_su) {
    match_backward_su_kernel<<<(unsigned)su_blocks, 256, 0, st>>>(ov, su, ori, coef_scratch, grad_su, G, Q, CH, W, sw);
    WITW_LAUNCH_CHECK();
  }
  return WITW_OK;
}

extern "C" int witw_triplet_loss_f32(const float* dist, int N, float alpha, float* loss, float* grad_dist, witw_stream_t stream) {
  WITW_REQUIRE(N >= 2 && N <= 4096, WITW_ERR_UNSUPPORTED, "witw_triplet_loss_f32: batch size %d outside 2..4096", N);
  WITW_REQUIRE(dist && loss, WITW_ERR_INVALID, "witw_triplet_loss_f32: null pointer");
  const size_t smem = sizeof(float) * (2 * (size_t)N + 32);
  triplet_loss_kernel<<<1, 1024, smem, as_stream(stream)>>>(dist, N, alpha, loss, grad_dist);
  WITW_LAUNCH_CHECK();
  return WITW_OK;
}

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
