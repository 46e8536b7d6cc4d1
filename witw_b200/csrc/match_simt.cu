// Exact fp32 orientation-searched distance on CUDA cores (a3+a4+a5 fused), plus the
// standalone a4 / a5 kernels kept for API parity with the reference's three-step chain.
//
//   corr[g,q,s] = sum_{ch,k<sw} ov[g,ch,(s+k)%W] * su[q,ch,k]      model/cvig_fov.py:297-312
//   ori[g,q]    = first argmax_s corr[g,q,s]                       model/cvig_fov.py:313
//   dist[g,q]   = 2*(1 - corr[g,q,ori]/(||crop||*||su_q||))        model/cvig_fov.py:318-363
//
// This is the path for small problems (training batches, one-query heat maps), for shapes
// the tensor-core kernel does not cover (W != 64) and for the exact true-match distances of
// the rank evaluation; the 10k x 10k sweeps go through match_tc.cu.
#include "common.cuh"

namespace witw {

constexpr int kSimtThreads = 256;
constexpr int kGT = 4;   // gallery items per CTA
constexpr int kQT = 32;  // queries per CTA
constexpr int kQPerThread = 8;

// Shared-memory carve-up (floats):
//   ov_rows  [kGT][W + sw - 1]   wrap-padded rows of the current feature row ch
//   su_rows  [sw][kQT]           current feature row of the 32 queries, k-major
//   col_e    [kGT][W]            column energies  sum_ch ov^2
//   q_e      [kQT]               query energies   sum_{ch,k} su^2
//   corr     [kGT][kQT][W]       only after the main loop (aliases ov_rows/su_rows? no: separate)
template <int NS>  // shifts per thread: 1 (W <= 64) or 2 (W <= 128)
__global__ void __launch_bounds__(kSimtThreads)
match_tile_kernel(const float* __restrict__ ov, const float* __restrict__ su, int64_t G, int64_t Q, int CH, int W, int sw,
                  float* __restrict__ dist, int64_t* __restrict__ ori, float* __restrict__ corr_out) {
  extern __shared__ float smem[];
  const int wpad = W + sw - 1;
  float* ov_rows = smem;                     // kGT * wpad
  float* su_rows = ov_rows + ((kGT * wpad + 3) & ~3);  // sw * kQT, 16-byte aligned for float4 reads
  float* col_e = su_rows + sw * kQT;         // kGT * W
  float* q_e = col_e + kGT * W;              // kQT
  float* corr = q_e + kQT;                   // kGT * kQT * W

  const int tid = threadIdx.x;
  const int s_lane = tid & 63;
  const int qgrp = tid >> 6;  // 0..3 -> queries 8*qgrp .. 8*qgrp+7
  const int64_t g0 = (int64_t)blockIdx.y * kGT;
  const int64_t q0 = (int64_t)blockIdx.x * kQT;

  float acc[NS][kGT][kQPerThread];
#pragma unroll
  for (int a = 0; a < NS; ++a)
#pragma unroll
    for (int g = 0; g < kGT; ++g)
#pragma unroll
      for (int j = 0; j < kQPerThread; ++j) acc[a][g][j] = 0.f;

  for (int i = tid; i < kGT * W; i += kSimtThreads) col_e[i] = 0.f;
  if (tid < kQT) q_e[tid] = 0.f;
  float q_part = 0.f;  // thread (q = tid/8, part = tid%8) accumulates part of ||su_q||^2

  for (int ch = 0; ch < CH; ++ch) {
    __syncthreads();  // previous iteration's readers are done
    for (int i = tid; i < kGT * wpad; i += kSimtThreads) {
      const int g = i / wpad, j = i - g * wpad;
      const int64_t gg = g0 + g;
      ov_rows[i] = gg < G ? ov[(gg * CH + ch) * W + (j % W)] : 0.f;
    }
    for (int i = tid; i < kQT * sw; i += kSimtThreads) {
      const int q = i / sw, k = i - q * sw;
      const int64_t qq = q0 + q;
      su_rows[k * kQT + q] = qq < Q ? su[(qq * CH + ch) * sw + k] : 0.f;
    }
    __syncthreads();
    for (int i = tid; i < kGT * W; i += kSimtThreads) {
      const int g = i / W, j = i - g * W;
      const float v = ov_rows[g * wpad + j];
      col_e[i] = fmaf(v, v, col_e[i]);
    }
    {
      const int q = tid >> 3, part = tid & 7;
      for (int k = part; k < sw; k += 8) {
        const float v = su_rows[k * kQT + q];
        q_part = fmaf(v, v, q_part);
      }
    }
    for (int k = 0; k < sw; ++k) {
      const float4 s0 = *reinterpret_cast<const float4*>(&su_rows[k * kQT + qgrp * kQPerThread]);
      const float4 s1 = *reinterpret_cast<const float4*>(&su_rows[k * kQT + qgrp * kQPerThread + 4]);
      const float sv[kQPerThread] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
      for (int a = 0; a < NS; ++a) {
        const int s = s_lane + 64 * a;
        if (NS == 1 || s < W) {
#pragma unroll
          for (int g = 0; g < kGT; ++g) {
            const float o = ov_rows[g * wpad + min(s, W - 1) + k];
#pragma unroll
            for (int j = 0; j < kQPerThread; ++j) acc[a][g][j] = fmaf(o, sv[j], acc[a][g][j]);
          }
        }
      }
    }
  }
  // ||su_q||^2: reduce the 8 partials of each query (lanes 8q..8q+7 of a warp)
  q_part += __shfl_xor_sync(0xffffffffu, q_part, 1);
  q_part += __shfl_xor_sync(0xffffffffu, q_part, 2);
  q_part += __shfl_xor_sync(0xffffffffu, q_part, 4);
  if ((tid & 7) == 0) q_e[tid >> 3] = q_part;

#pragma unroll
  for (int a = 0; a < NS; ++a) {
    const int s = s_lane + 64 * a;
    if (s < W) {
#pragma unroll
      for (int g = 0; g < kGT; ++g)
#pragma unroll
        for (int j = 0; j < kQPerThread; ++j) corr[(g * kQT + qgrp * kQPerThread + j) * W + s] = acc[a][g][j];
    }
  }
  __syncthreads();

  if (corr_out != nullptr) {
    for (int i = tid; i < kGT * kQT * W; i += kSimtThreads) {
      const int s = i % W, gq = i / W;
      const int64_t gg = g0 + gq / kQT, qq = q0 + gq % kQT;
      if (gg < G && qq < Q) corr_out[(gg * Q + qq) * W + s] = corr[i];
    }
  }
  if (tid < kGT * kQT) {
    const int g = tid / kQT, q = tid % kQT;
    const int64_t gg = g0 + g, qq = q0 + q;
    if (gg < G && qq < Q) {
      const float* c = corr + (size_t)tid * W;
      // torch.argmax: first maximum wins; a NaN is the maximum
      float best = c[0];
      int arg = 0;
      for (int s = 1; s < W; ++s) {
        const float v = c[s];
        if (v > best || (v != v && best == best)) { best = v; arg = s; }
      }
      float cn2 = 0.f;
      for (int k = 0; k < sw; ++k) {
        int j = arg + k;
        j -= (j >= W) ? W : 0;
        cn2 += col_e[g * W + j];
      }
      const float d = 2.0f * (1.0f - best / (sqrtf(cn2) * sqrtf(q_e[q])));
      if (dist != nullptr) dist[gg * Q + qq] = d;
      if (ori != nullptr) ori[gg * Q + qq] = arg;
    }
  }
}

// one CTA per explicit (gallery, query) pair; 128 threads, thread s owns shift s
__global__ void __launch_bounds__(128)
match_pairs_kernel(const float* __restrict__ ov, const float* __restrict__ su, const int64_t* __restrict__ pair_g,
                   const int64_t* __restrict__ pair_q, int CH, int W, int sw, float* __restrict__ dist,
                   int64_t* __restrict__ ori) {
  extern __shared__ float smem[];
  float* o = smem;            // CH * W
  float* s = o + CH * W;      // CH * sw
  float* col_e = s + CH * sw; // W
  float* cbuf = col_e + W;    // W
  __shared__ float q_e_sh;
  const int tid = threadIdx.x;
  const int64_t g = pair_g[blockIdx.x], q = pair_q[blockIdx.x];
  for (int i = tid; i < CH * W; i += 128) o[i] = ov[g * CH * W + i];
  for (int i = tid; i < CH * sw; i += 128) s[i] = su[q * CH * sw + i];
  __syncthreads();
  if (tid < W) {
    float e = 0.f;
    for (int ch = 0; ch < CH; ++ch) e = fmaf(o[ch * W + tid], o[ch * W + tid], e);
    col_e[tid] = e;
    float a0 = 0.f, a1 = 0.f;
    for (int ch = 0; ch < CH; ++ch) {
      const float* orow = o + ch * W;
      const float* srow = s + ch * sw;
      int k = 0;
      for (; k + 1 < sw; k += 2) {
        int j0 = tid + k, j1 = tid + k + 1;
        j0 -= (j0 >= W) ? W : 0;
        j1 -= (j1 >= W) ? W : 0;
        a0 = fmaf(orow[j0], srow[k], a0);
        a1 = fmaf(orow[j1], srow[k + 1], a1);
      }
      if (k < sw) {
        int j0 = tid + k;
        j0 -= (j0 >= W) ? W : 0;
        a0 = fmaf(orow[j0], srow[k], a0);
      }
    }
    cbuf[tid] = a0 + a1;
  }
  // ||su||^2 by warp 3 (threads 96..127) so it overlaps the correlation of the others when W <= 96
  if (tid >= 96) {
    float e = 0.f;
    for (int i = tid - 96; i < CH * sw; i += 32) e = fmaf(s[i], s[i], e);
    for (int m = 16; m > 0; m >>= 1) e += __shfl_xor_sync(0xffffffffu, e, m);
    if (tid == 96) q_e_sh = e;
  }
  __syncthreads();
  if (tid == 0) {
    float best = cbuf[0];
    int arg = 0;
    for (int i = 1; i < W; ++i) {
      const float v = cbuf[i];
      if (v > best || (v != v && best == best)) { best = v; arg = i; }
    }
    float cn2 = 0.f;
    for (int k = 0; k < sw; ++k) {
      int j = arg + k;
      j -= (j >= W) ? W : 0;
      cn2 += col_e[j];
    }
    if (dist != nullptr) dist[blockIdx.x] = 2.0f * (1.0f - best / (sqrtf(cn2) * sqrtf(q_e_sh)));
    if (ori != nullptr) ori[blockIdx.x] = arg;
  }
}

// Register-tiled variant for the reference geometry (W == 64, sw % 4 == 0, CH % 8 == 0): thread (slice, s4) owns the
// four shifts 4*s4..4*s4+3 over one eighth of the feature rows; per 4 query columns it reads 12 gallery values
// (two aligned 16-byte shared loads from the wrap-padded row) and 4 query values (one broadcast 16-byte load) for
// 16 FMAs.  ~9x the throughput of match_pairs_kernel, which stays as the generic-shape path.
__global__ void __launch_bounds__(128)
match_pairs_w64_kernel(const float* __restrict__ ov, const float* __restrict__ su, const int64_t* __restrict__ pair_g,
                       const int64_t* __restrict__ pair_q, int64_t n_pairs, const int32_t* __restrict__ n_pairs_dev, int group,
                       int CH, int sw, float* __restrict__ dist, int64_t* __restrict__ ori) {
  // one CTA per `group` consecutive pairs that share their query (group == 1: no such promise needed): the query is
  // staged once, the gallery items stream through
  extern __shared__ __align__(16) float smem[];
  constexpr int W = 64, WP = 128;
  float* o = smem;                 // CH * 128 : row followed by its own copy (circular wrap)
  float* s = o + CH * WP;          // CH * sw
  float* part = s + CH * sw;       // 8 * 64 partial correlations
  float* col_part = part + 8 * W;  // 2 * 64 partial column energies
  __shared__ float q_e_sh;
  const int64_t p0 = (int64_t)blockIdx.x * group;
  if (n_pairs_dev != nullptr && p0 >= (int64_t)*n_pairs_dev) return;
  if (p0 >= n_pairs) return;
  const int tid = threadIdx.x;
  const int64_t q = pair_q[p0];
  const float4* qsrc = reinterpret_cast<const float4*>(su + q * CH * sw);
  for (int i = tid; i < CH * sw / 4; i += 128) reinterpret_cast<float4*>(s)[i] = qsrc[i];
  const int s4 = tid & 15, slice = tid >> 4;
  const int rows = CH >> 3;
  for (int c = 0; c < group && p0 + c < n_pairs; ++c) {
    const int64_t g = pair_g[p0 + c];
    __syncthreads();  // previous item fully consumed (and, first time round, nothing yet)
    const float4* gsrc = reinterpret_cast<const float4*>(ov + g * CH * W);
    for (int i = tid; i < CH * 16; i += 128) {  // 16 float4 per row, written twice
      const float4 v = gsrc[i];
      const int ch = i >> 4, c4 = i & 15;
      reinterpret_cast<float4*>(o + ch * WP)[c4] = v;
      reinterpret_cast<float4*>(o + ch * WP + W)[c4] = v;
    }
    __syncthreads();
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int r = 0; r < rows; ++r) {
      const int ch = slice * rows + r;
      const float* orow = o + ch * WP + 4 * s4;
      const float* srow = s + ch * sw;
      for (int k = 0; k < sw; k += 4) {
        const float4 a = *reinterpret_cast<const float4*>(orow + k);
        const float4 b = *reinterpret_cast<const float4*>(orow + k + 4);
        const float4 cq = *reinterpret_cast<const float4*>(srow + k);
        const float w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        const float x[4] = {cq.x, cq.y, cq.z, cq.w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int d = 0; d < 4; ++d) acc[d] = fmaf(w[d + j], x[j], acc[d]);
      }
    }
#pragma unroll
    for (int d = 0; d < 4; ++d) part[slice * W + 4 * s4 + d] = acc[d];
    {  // column energies (two halves of the rows)
      const int j = tid & 63, half = tid >> 6;
      float e = 0.f;
      for (int ch = half * (CH >> 1); ch < (half + 1) * (CH >> 1); ++ch) e = fmaf(o[ch * WP + j], o[ch * WP + j], e);
      col_part[half * W + j] = e;
    }
    if (c == 0 && tid >= 96) {  // query energy, once
      float e = 0.f;
      for (int i = tid - 96; i < CH * sw; i += 32) e = fmaf(s[i], s[i], e);
      for (int m = 16; m > 0; m >>= 1) e += __shfl_xor_sync(0xffffffffu, e, m);
      if (tid == 96) q_e_sh = e;
    }
    __syncthreads();
    if (tid < 32) {  // warp 0: finish the 8-way sums, then a warp argmax with first-maximum tie-break
      float c0 = 0.f, c1 = 0.f;
#pragma unroll
      for (int sl = 0; sl < 8; ++sl) { c0 += part[sl * W + tid]; c1 += part[sl * W + tid + 32]; }
      float best = c0;
      int arg = tid;
      if (c1 > best || (c1 != c1 && best == best)) { best = c1; arg = tid + 32; }
      for (int m = 16; m > 0; m >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, m);
        const int oa = __shfl_xor_sync(0xffffffffu, arg, m);
        const bool take = (ob > best) || (ob != ob && best == best) || (ob == best && oa < arg) || (ob != ob && best != best && oa < arg);
        if (take) { best = ob; arg = oa; }
      }
      if (tid == 0) {
        float cn2 = 0.f;
        for (int k = 0; k < sw; ++k) { const int j = (arg + k) & 63; cn2 += col_part[j] + col_part[W + j]; }
        if (dist != nullptr) dist[p0 + c] = 2.0f * (1.0f - best / (sqrtf(cn2) * sqrtf(q_e_sh)));
        if (ori != nullptr) ori[p0 + c] = arg;
      }
    }
  }
}

__global__ void __launch_bounds__(256)
crop_gather_kernel(const float* __restrict__ ov, const int64_t* __restrict__ ori, float* __restrict__ out, int64_t G,
                   int64_t Q, int CH, int W, int sw) {
  // one CTA per (g, q): out[g,q,ch,k] = ov[g,ch,(k + ori[g,q]) % W]
  const int64_t gq = blockIdx.x;
  const int64_t g = gq / Q;
  int sh = (int)(ori[gq] % W);
  if (sh < 0) sh += W;
  const float* src = ov + g * CH * W;
  float* dst = out + gq * CH * sw;
  for (int i = threadIdx.x; i < CH * sw; i += blockDim.x) {
    const int ch = i / sw, k = i - ch * sw;
    int j = k + sh;
    j -= (j >= W) ? W : 0;
    dst[i] = src[ch * W + j];
  }
}

__global__ void __launch_bounds__(256)
l2_distance_kernel(const float* __restrict__ crop, const float* __restrict__ su, float* __restrict__ dist, int64_t G,
                   int64_t Q, int64_t K) {
  // one warp per (g, q)
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
  if (lane == 0) dist[gq] = 2.0f * (1.0f - dot / (sqrtf(oe) * sqrtf(se)));
}

}  // namespace witw

using namespace witw;

static int check_match_shape(const char* fn, int64_t G, int64_t Q, int CH, int W, int sw) {
  WITW_REQUIRE(G >= 0 && Q >= 0 && CH > 0 && W > 0 && sw > 0, WITW_ERR_INVALID, "%s: bad shape G=%lld Q=%lld CH=%d W=%d sw=%d", fn,
               (long long)G, (long long)Q, CH, W, sw);
  WITW_REQUIRE(sw <= W, WITW_ERR_INVALID, "%s: query width %d exceeds gallery width %d", fn, sw, W);
  WITW_REQUIRE(W <= 128, WITW_ERR_UNSUPPORTED, "%s: gallery width %d > 128 is not supported", fn, W);
  return WITW_OK;
}

extern "C" int witw_match_f32(const float* ov, const float* su, int64_t G, int64_t Q, int CH, int W, int sw, float* dist,
                              int64_t* ori, float* corr, witw_stream_t stream) {
  int rc = check_match_shape("witw_match_f32", G, Q, CH, W, sw);
  if (rc != WITW_OK) return rc;
  if (G == 0 || Q == 0) return WITW_OK;
  WITW_REQUIRE(ov && su, WITW_ERR_INVALID, "witw_match_f32: null input");
  const int64_t gx = ceil_div<int64_t>(Q, kQT), gy = ceil_div<int64_t>(G, kGT);
  WITW_REQUIRE(gy <= 65535, WITW_ERR_UNSUPPORTED, "witw_match_f32: gallery of %lld items exceeds the fp32 path (use the tensor-core path)", (long long)G);
  const size_t smem = sizeof(float) * ((size_t)((kGT * (W + sw - 1) + 3) & ~3) + (size_t)sw * kQT + (size_t)kGT * W + kQT + (size_t)kGT * kQT * W);
  if (W <= 64) {
    WITW_CUDA(cudaFuncSetAttribute(match_tile_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    match_tile_kernel<1><<<dim3((unsigned)gx, (unsigned)gy), kSimtThreads, smem, as_stream(stream)>>>(ov, su, G, Q, CH, W, sw, dist, ori, corr);
  } else {
    WITW_CUDA(cudaFuncSetAttribute(match_tile_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    match_tile_kernel<2><<<dim3((unsigned)gx, (unsigned)gy), kSimtThreads, smem, as_stream(stream)>>>(ov, su, G, Q, CH, W, sw, dist, ori, corr);
  }
  WITW_LAUNCH_CHECK();
  return WITW_OK;
}

static int launch_pairs(const float* ov, const float* su, const int64_t* pair_g, const int64_t* pair_q, int64_t n_pairs,
                        const int32_t* n_pairs_dev, int CH, int W, int sw, float* dist, int64_t* ori, witw_stream_t stream,
                        int group = 1) {
  const bool fast = W == 64 && sw % 4 == 0 && CH % 8 == 0 && (((uintptr_t)ov | (uintptr_t)su) & 15) == 0;
  if (fast) {
    const size_t smem = sizeof(float) * ((size_t)CH * 128 + (size_t)CH * sw + 8 * 64 + 2 * 64);
    if (smem <= 200 * 1024) {
      WITW_CUDA(cudaFuncSetAttribute(match_pairs_w64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      match_pairs_w64_kernel<<<(unsigned)ceil_div<int64_t>(n_pairs, group), 128, smem, as_stream(stream)>>>(ov, su, pair_g, pair_q, n_pairs, n_pairs_dev,
                                                                                                      group, CH, sw, dist, ori);
      WITW_LAUNCH_CHECK();
      return WITW_OK;
    }
  }
  WITW_REQUIRE(n_pairs_dev == nullptr, WITW_ERR_UNSUPPORTED, "device-side pair count needs the W == 64 kernel");
  const size_t smem = sizeof(float) * ((size_t)CH * W + (size_t)CH * sw + 2 * (size_t)W);
  WITW_REQUIRE(smem <= 200 * 1024, WITW_ERR_UNSUPPORTED, "witw_match_pairs_f32: feature map of %d rows does not fit shared memory", CH);
  WITW_CUDA(cudaFuncSetAttribute(match_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  match_pairs_kernel<<<(unsigned)n_pairs, 128, smem, as_stream(stream)>>>(ov, su, pair_g, pair_q, CH, W, sw, dist, ori);
  WITW_LAUNCH_CHECK();
  return WITW_OK;
}

extern "C" int witw_match_pairs_f32(const float* ov, const float* su, const int64_t* pair_g, const int64_t* pair_q,
                                    int64_t n_pairs, int CH, int W, int sw, float* dist, int64_t* ori, witw_stream_t stream) {
  int rc = check_match_shape("witw_match_pairs_f32", 1, 1, CH, W, sw);
  if (rc != WITW_OK) return rc;
  if (n_pairs == 0) return WITW_OK;
  WITW_REQUIRE(ov && su && pair_g && pair_q && n_pairs > 0 && n_pairs < (1ll << 31), WITW_ERR_INVALID, "witw_match_pairs_f32: bad arguments");
  return launch_pairs(ov, su, pair_g, pair_q, n_pairs, nullptr, CH, W, sw, dist, ori, stream);
}

extern "C" int witw_crop_gather_f32(const float* ov, const int64_t* ori, float* out, int64_t G, int64_t Q, int CH, int W, int sw,
                                    witw_stream_t stream) {
  WITW_REQUIRE(G >= 0 && Q >= 0 && CH > 0 && W > 0 && sw > 0 && sw <= W, WITW_ERR_INVALID, "witw_crop_gather_f32: bad shape");
  if (G == 0 || Q == 0) return WITW_OK;
  WITW_REQUIRE(ov && ori && out, WITW_ERR_INVALID, "witw_crop_gather_f32: null pointer");
  WITW_REQUIRE(G * Q < (1ll << 31), WITW_ERR_UNSUPPORTED, "witw_crop_gather_f32: %lld pairs is too many to materialise", (long long)(G * Q));
  crop_gather_kernel<<<(unsigned)(G * Q), 256, 0, as_stream(stream)>>>(ov, ori, out, G, Q, CH, W, sw);
  WITW_LAUNCH_CHECK();
  return WITW_OK;
}

extern "C" int witw_l2_distance_f32(const float* crop, const float* su, float* dist, int64_t G, int64_t Q, int64_t K,
                                    witw_stream_t stream) {
  WITW_REQUIRE(G >= 0 && Q >= 0 && K > 0, WITW_ERR_INVALID, "witw_l2_distance_f32: bad shape");
  if (G == 0 || Q == 0) return WITW_OK;
  WITW_REQUIRE(crop && su && dist, WITW_ERR_INVALID, "witw_l2_distance_f32: null pointer");
  const int64_t blocks = ceil_div<int64_t>(G * Q, 8);
  WITW_REQUIRE(blocks < (1ll << 31), WITW_ERR_UNSUPPORTED, "witw_l2_distance_f32: too many pairs");
  l2_distance_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(crop, su, dist, G, Q, K);
  WITW_LAUNCH_CHECK();
  return WITW_OK;
}
