// K4 -- rank by counting and top-k over a materialised [gallery, query] distance matrix,
// plus the baseline Euclidean variant.  Replaces the per-query body of the rank loops in
// model/cvig_fov.py:550-552 and model/cvig_baseline.py:458-460:
//     rank[q] = #{ g : d[g,q] <= d[true(q), q] }     (ties and the match itself count,
//                                                     NaN compares false)
// HBM-bound: 4*G*Q bytes read once.  Columns are queries, so a warp reads 32 (or 128 with
// float4) consecutive queries of one gallery row per request; counts stay in registers and
// leave the SM as one atomic per (query, gallery slice).
#include "common.cuh"

namespace witw {

constexpr int kRankThreads = 128;

// float4 path: Q % 4 == 0, 16-byte aligned base
__global__ void __launch_bounds__(kRankThreads)
rank_count_vec4_kernel(const float* __restrict__ dist, int64_t G, int64_t Q, const int64_t* __restrict__ true_idx,
                       unsigned long long* __restrict__ ranks, int64_t rows_per_block) {
  const int64_t q4 = (int64_t)blockIdx.x * kRankThreads + threadIdx.x;  // index of a group of 4 queries
  if (q4 * 4 >= Q) return;
  const int64_t q = q4 * 4;
  float thr[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int64_t t = true_idx ? true_idx[q + j] : (q + j);
    thr[j] = (t >= 0 && t < G) ? dist[t * Q + q + j] : __int_as_float(0x7fc00000);  // NaN: nothing counts
  }
  const int64_t g0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t g1 = min(g0 + rows_per_block, G);
  unsigned c0 = 0, c1 = 0, c2 = 0, c3 = 0;
  const float4* p = reinterpret_cast<const float4*>(dist + q);
  const int64_t stride4 = Q >> 2;
  int64_t g = g0;
  for (; g + 4 <= g1; g += 4) {  // 4 independent 16-byte loads in flight per thread
    const float4 a = __ldcs(p + g * stride4), b = __ldcs(p + (g + 1) * stride4);
    const float4 c = __ldcs(p + (g + 2) * stride4), d = __ldcs(p + (g + 3) * stride4);
    c0 += (a.x <= thr[0]) + (b.x <= thr[0]) + (c.x <= thr[0]) + (d.x <= thr[0]);
    c1 += (a.y <= thr[1]) + (b.y <= thr[1]) + (c.y <= thr[1]) + (d.y <= thr[1]);
    c2 += (a.z <= thr[2]) + (b.z <= thr[2]) + (c.z <= thr[2]) + (d.z <= thr[2]);
    c3 += (a.w <= thr[3]) + (b.w <= thr[3]) + (c.w <= thr[3]) + (d.w <= thr[3]);
  }
  for (; g < g1; ++g) {
    const float4 a = __ldcs(p + g * stride4);
    c0 += (a.x <= thr[0]);
    c1 += (a.y <= thr[1]);
    c2 += (a.z <= thr[2]);
    c3 += (a.w <= thr[3]);
  }
  if (c0) atomicAdd(ranks + q, (unsigned long long)c0);
  if (c1) atomicAdd(ranks + q + 1, (unsigned long long)c1);
  if (c2) atomicAdd(ranks + q + 2, (unsigned long long)c2);
  if (c3) atomicAdd(ranks + q + 3, (unsigned long long)c3);
}

__global__ void __launch_bounds__(kRankThreads)
rank_count_scalar_kernel(const float* __restrict__ dist, int64_t G, int64_t Q, const int64_t* __restrict__ true_idx,
                         unsigned long long* __restrict__ ranks, int64_t rows_per_block) {
  const int64_t q = (int64_t)blockIdx.x * kRankThreads + threadIdx.x;
  if (q >= Q) return;
  const int64_t t = true_idx ? true_idx[q] : q;
  const float thr = (t >= 0 && t < G) ? dist[t * Q + q] : __int_as_float(0x7fc00000);
  const int64_t g0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t g1 = min(g0 + rows_per_block, G);
  unsigned c = 0;
  for (int64_t g = g0; g < g1; ++g) c += (__ldcs(dist + g * Q + q) <= thr);
  if (c) atomicAdd(ranks + q, (unsigned long long)c);
}

// Euclidean distance matrix of the baseline model: one warp per (gallery row, 8 queries held in shared memory)
constexpr int kL2QT = 8;
__global__ void __launch_bounds__(256)
l2_dist_kernel(const float* __restrict__ ov, const float* __restrict__ su, int64_t N, int64_t Q, int64_t D,
               float* __restrict__ dist) {
  extern __shared__ float sq[];  // kL2QT * D
  const int64_t q0 = (int64_t)blockIdx.x * kL2QT;
  for (int64_t i = threadIdx.x; i < kL2QT * D; i += blockDim.x) {
    const int64_t q = q0 + i / D;
    sq[i] = q < Q ? su[q * D + i % D] : 0.f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t n = (int64_t)blockIdx.y * 8 + warp; n < N; n += (int64_t)gridDim.y * 8) {
    float acc[kL2QT];
#pragma unroll
    for (int j = 0; j < kL2QT; ++j) acc[j] = 0.f;
    const float* o = ov + n * D;
    for (int64_t d = lane; d < D; d += 32) {
      const float v = o[d];
#pragma unroll
      for (int j = 0; j < kL2QT; ++j) {
        const float t = v - sq[j * D + d];
        acc[j] = fmaf(t, t, acc[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < kL2QT; ++j) {
      float a = acc[j];
      for (int m = 16; m > 0; m >>= 1) a += __shfl_xor_sync(0xffffffffu, a, m);
      if (lane == 0 && q0 + j < Q) dist[n * Q + q0 + j] = sqrtf(a);  // pow(sum, 0.5)
    }
  }
}

// ------------------------------------------------------------------------------------------
// top-k: one thread per query column keeps an ascending list of k (distance, index) in shared
// memory, laid out [k][threads] so a warp's accesses to one list slot hit 32 different banks.
// The gallery is cut into row slices for parallelism; short slices never warm their lists up (a slice of 167 rows
// inserts 23 % of its elements), so a first pass over a strided sample of the rows (row_stride > 1) leaves candidate
// lists whose k-th entries bound every column's k-th smallest distance from above, and the full pass (tau_lists set)
// starts each list with that bound as its admission threshold: ~1 % of the elements are inserted instead of 23 %.
// ------------------------------------------------------------------------------------------
constexpr int kTopkThreads = 64;

template <int VEC>  // adjacent query columns per thread: 4 (float4 loads) or 1
__global__ void __launch_bounds__(kTopkThreads)
topk_columns_kernel(const float* __restrict__ dist, int64_t G, int64_t Q, int k, float* __restrict__ out_d,
                    int32_t* __restrict__ out_i, int32_t g_offset, int64_t rows_per_slice, int64_t row_stride,
                    const float* tau_lists, int tau_slices) {
  // blockIdx.x: 64*VEC query columns (a warp reads 32*VEC adjacent queries of one gallery row);
  // blockIdx.y: gallery slice, whose candidate list goes to out[slice][q][k]
  extern __shared__ unsigned char raw[];
  constexpr int LW = VEC * kTopkThreads;                            // lists per CTA
  float* ld = reinterpret_cast<float*>(raw);                       // [k][VEC][kTopkThreads]
  int32_t* li = reinterpret_cast<int32_t*>(ld + (size_t)k * LW);
  const int t = threadIdx.x;
  const int64_t q0 = ((int64_t)blockIdx.x * kTopkThreads + t) * VEC;
  const float inf = __int_as_float(0x7f800000);
  for (int j = 0; j < k; ++j)
#pragma unroll
    for (int c = 0; c < VEC; ++c) { ld[j * LW + c * kTopkThreads + t] = inf; li[j * LW + c * kTopkThreads + t] = -1; }
  if (q0 >= Q) return;
  const int64_t g0 = (int64_t)blockIdx.y * rows_per_slice;
  const int64_t g1 = min(g0 + rows_per_slice, G);
  float worst[VEC], bound[VEC];
  int filled[VEC];
#pragma unroll
  for (int c = 0; c < VEC; ++c) {
    // admission threshold: d < worst.  Any full list's k-th entry is >= the column's k-th smallest distance, so the
    // smallest such entry of the sample pass (or of a slice that already finished: the slots are reused, every value
    // ever stored there is a valid bound) admits exactly the elements <= it; an unfilled list gives +inf.
    float tau = inf;
    if (tau_lists != nullptr && q0 + c < Q)
      for (int sl = 0; sl < tau_slices; ++sl) tau = fminf(tau, __ldcg(tau_lists + ((int64_t)sl * Q + q0 + c) * k + (k - 1)));
    bound[c] = tau < inf ? __uint_as_float(__float_as_uint(tau) + (tau >= 0.f ? 1u : 0xffffffffu)) : inf;  // next float above tau
    if (tau == 0.f) bound[c] = __uint_as_float(1u);
    worst[c] = bound[c];
    filled[c] = 0;
  }
  auto insert = [&](int c, float d, int64_t g) {
    // the caller checked d < worst[c]: strict '<' keeps the earlier (lower) gallery index on ties; NaN, +inf never enter
    float* cd = ld + c * kTopkThreads + t;
    int32_t* ci = li + c * kTopkThreads + t;
    int j = filled[c] < k ? filled[c] : k - 1;
    while (j > 0 && cd[(j - 1) * LW] > d) {
      cd[j * LW] = cd[(j - 1) * LW];
      ci[j * LW] = ci[(j - 1) * LW];
      --j;
    }
    cd[j * LW] = d;
    ci[j * LW] = (int32_t)(g * row_stride) + g_offset;
    if (filled[c] < k) ++filled[c];
    if (filled[c] == k) worst[c] = fminf(bound[c], cd[(k - 1) * LW]);
  };
  int64_t g = g0;
  if constexpr (VEC == 4) {
    constexpr int U = 4;  // 4 x 16 bytes in flight per thread
    const float4* p = reinterpret_cast<const float4*>(dist + q0);
    const int64_t stride4 = (Q >> 2) * row_stride;
    for (; g + U <= g1; g += U) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) v[u] = __ldcs(p + (g + u) * stride4);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (v[u].x < worst[0]) insert(0, v[u].x, g + u);
        if (v[u].y < worst[1]) insert(1, v[u].y, g + u);
        if (v[u].z < worst[2]) insert(2, v[u].z, g + u);
        if (v[u].w < worst[3]) insert(3, v[u].w, g + u);
      }
    }
    for (; g < g1; ++g) {
      const float4 v = __ldcs(p + g * stride4);
      if (v.x < worst[0]) insert(0, v.x, g);
      if (v.y < worst[1]) insert(1, v.y, g);
      if (v.z < worst[2]) insert(2, v.z, g);
      if (v.w < worst[3]) insert(3, v.w, g);
    }
  } else {
    constexpr int U = 8;
    const float* p = dist + q0;
    const int64_t stride1 = Q * row_stride;
    for (; g + U <= g1; g += U) {
      float v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) v[u] = __ldcs(p + (g + u) * stride1);
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (v[u] < worst[0]) insert(0, v[u], g + u);
    }
    for (; g < g1; ++g) {
      const float v = __ldcs(p + g * stride1);
      if (v < worst[0]) insert(0, v, g);
    }
  }
#pragma unroll
  for (int c = 0; c < VEC; ++c) {
    if (q0 + c >= Q) break;
    float* od = out_d + ((int64_t)blockIdx.y * Q + q0 + c) * k;
    int32_t* oi = out_i + ((int64_t)blockIdx.y * Q + q0 + c) * k;
    for (int j = 0; j < k; ++j) {
      od[j] = ld[j * LW + c * kTopkThreads + t];
      oi[j] = li[j * LW + c * kTopkThreads + t];
    }
  }
}

// k-way merge of sorted candidate lists, one warp per query.  Lane l owns the heads of lists l and l + 32; k times the
// warp takes the smallest head (ties: the lower list, i.e. the lower gallery indices, as a stable sort would) and its
// owner advances.  Padding (index < 0) and NaN / +inf distances end a list.
constexpr int kMergeWarps = 8;

__global__ void __launch_bounds__(kMergeWarps * 32)
topk_merge_kernel(const float* __restrict__ cd, const int32_t* __restrict__ ci, int n_lists, int64_t Q, int k,
                  float* __restrict__ out_d, int32_t* __restrict__ out_i) {
  const int lane = threadIdx.x & 31;
  const int64_t q = (int64_t)blockIdx.x * kMergeWarps + (threadIdx.x >> 5);
  if (q >= Q) return;
  const float inf = __int_as_float(0x7f800000);
  const float* pd[2];
  const int32_t* pi[2];
  float hd[2];
  int32_t hi[2];
  int pos[2] = {0, 0};
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int l = lane + 32 * h;
    pd[h] = cd + ((size_t)min(l, n_lists - 1) * Q + q) * k;
    pi[h] = ci + ((size_t)min(l, n_lists - 1) * Q + q) * k;
    hd[h] = inf;
    hi[h] = -1;
    if (l < n_lists) {
      const float d = pd[h][0];
      const int32_t i = pi[h][0];
      if (i >= 0 && d < inf) { hd[h] = d; hi[h] = i; }
    }
  }
  for (int j = 0; j < k; ++j) {
    // this lane's better head: list `lane` wins ties against list `lane + 32`
    const int mine = (hd[1] < hd[0]) ? 1 : 0;
    float bd = mine ? hd[1] : hd[0];
    int bl = lane + 32 * mine;
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
      const float od = __shfl_xor_sync(0xffffffffu, bd, m);
      const int ol = __shfl_xor_sync(0xffffffffu, bl, m);
      if (od < bd || (od == bd && ol < bl)) { bd = od; bl = ol; }
    }
    const bool any = bd < inf;
    const int owner = bl & 31, which = bl >> 5;
    int32_t widx = -1;
    if (lane == owner && any) {
      widx = which ? hi[1] : hi[0];
      // advance the winning list
      const int np = (which ? pos[1] : pos[0]) + 1;
      float nd = inf;
      int32_t ni = -1;
      if (np < k) {
        const float d = (which ? pd[1] : pd[0])[np];
        const int32_t i = (which ? pi[1] : pi[0])[np];
        if (i >= 0 && d < inf) { nd = d; ni = i; }
      }
      if (which) { pos[1] = np; hd[1] = nd; hi[1] = ni; } else { pos[0] = np; hd[0] = nd; hi[0] = ni; }
    }
    widx = __shfl_sync(0xffffffffu, widx, owner);
    if (lane == 0) {
      out_d[q * k + j] = any ? bd : inf;
      out_i[q * k + j] = any ? widx : -1;
    }
  }
}

}  // namespace witw

using namespace witw;

extern "C" int witw_rank_from_dist_f32(const float* dist, int64_t G, int64_t Q, const int64_t* true_idx, int64_t* ranks,
                                       witw_stream_t stream) {
  WITW_REQUIRE(G >= 0 && Q >= 0, WITW_ERR_INVALID, "witw_rank_from_dist_f32: bad shape");
  if (Q == 0) return WITW_OK;
  WITW_REQUIRE(ranks && (dist || G == 0), WITW_ERR_INVALID, "witw_rank_from_dist_f32: null pointer");
  WITW_CUDA(cudaMemsetAsync(ranks, 0, sizeof(int64_t) * Q, as_stream(stream)));
  if (G == 0) return WITW_OK;
  const bool vec = (Q % 4 == 0) && (((uintptr_t)dist & 15) == 0);
  const int64_t cols = vec ? Q / 4 : Q;
  const int64_t bx = ceil_div<int64_t>(cols, kRankThreads);
  // slice the gallery so the grid is a few waves of the machine
  int64_t by = std::max<int64_t>(1, std::min<int64_t>(ceil_div<int64_t>(G, 64), (int64_t)sm_count() * 16 / bx));
  by = std::min<int64_t>(by, 65535);
  const int64_t rows = ceil_div<int64_t>(G, by);
  by = ceil_div<int64_t>(G, rows);
  WITW_REQUIRE(bx < (1ll << 31), WITW_ERR_INVALID, "witw_rank_from_dist_f32: too many queries");
  auto* r = reinterpret_cast<unsigned long long*>(ranks);
  if (vec)
    rank_count_vec4_kernel<<<dim3((unsigned)bx, (unsigned)by), kRankThreads, 0, as_stream(stream)>>>(dist, G, Q, true_idx, r, rows);
  else
    rank_count_scalar_kernel<<<dim3((unsigned)bx, (unsigned)by), kRankThreads, 0, as_stream(stream)>>>(dist, G, Q, true_idx, r, rows);
  WITW_LAUNCH_CHECK();
  return WITW_OK;
}

extern "C" int witw_l2_rank_f32(const float* ov, const float* su, int64_t N, int64_t Q, int64_t D, const int64_t* true_idx,
                                float* dist, int64_t* ranks, witw_stream_t stream) {
  WITW_REQUIRE(N >= 0 && Q >= 0 && D > 0, WITW_ERR_INVALID, "witw_l2_rank_f32: bad shape");
  if (Q == 0) return WITW_OK;
  WITW_REQUIRE(dist != nullptr, WITW_ERR_INVALID, "witw_l2_rank_f32: dist_dev [N,Q] is required (it is the workspace the ranks are counted from)");
  WITW_REQUIRE((ov && su) || N == 0, WITW_ERR_INVALID, "witw_l2_rank_f32: null input");
  const size_t smem = sizeof(float) * kL2QT * (size_t)D;
  WITW_REQUIRE(smem <= 200 * 1024, WITW_ERR_UNSUPPORTED, "witw_l2_rank_f32: embedding dimension %lld too large", (long long)D);
  if (N > 0) {
    WITW_CUDA(cudaFuncSetAttribute(l2_dist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t bx = ceil_div<int64_t>(Q, kL2QT);
    int64_t by = std::max<int64_t>(1, std::min<int64_t>(ceil_div<int64_t>(N, 8), (int64_t)sm_count() * 8 / bx));
    by = std::min<int64_t>(by, 65535);
    WITW_REQUIRE(bx < (1ll << 31), WITW_ERR_INVALID, "witw_l2_rank_f32: too many queries");
    l2_dist_kernel<<<dim3((unsigned)bx, (unsigned)by), 256, smem, as_stream(stream)>>>(ov, su, N, Q, D, dist);
    WITW_LAUNCH_CHECK();
  }
  if (ranks != nullptr) return witw_rank_from_dist_f32(dist, N, Q, true_idx, ranks, stream);
  return WITW_OK;
}

extern "C" int witw_topk_slices(int64_t G, int64_t Q) {
  // enough (query block, gallery slice) CTAs for ~16 per SM (~8 for the thresholded two-pass form of large galleries, whose
  // per-slice cost is the list set-up and write-back rather than insertions), slices of at least 64 rows, at most 64
  // (merge limit)
  const int64_t bx = ceil_div<int64_t>(std::max<int64_t>(Q, 1), kTopkThreads * 4);
  int64_t s = ceil_div<int64_t>((int64_t)sm_count() * (G >= 8192 ? 8 : 16), bx);
  s = std::min<int64_t>(s, std::max<int64_t>(1, G / 64));
  return (int)std::max<int64_t>(1, std::min<int64_t>(s, 64));
}

extern "C" int witw_topk_from_dist_f32(const float* dist, int64_t G, int64_t Q, int k, int n_slices, float* topk_dist,
                                       int32_t* topk_idx, int32_t g_offset, witw_stream_t stream) {
  WITW_REQUIRE(G >= 0 && Q >= 0 && k > 0 && k <= 128, WITW_ERR_INVALID, "witw_topk_from_dist_f32: bad shape (k must be 1..128)");
  WITW_REQUIRE(n_slices >= 1 && n_slices <= 64, WITW_ERR_INVALID, "witw_topk_from_dist_f32: n_slices must be 1..64");
  if (Q == 0) return WITW_OK;
  WITW_REQUIRE(topk_dist && topk_idx && (dist || G == 0), WITW_ERR_INVALID, "witw_topk_from_dist_f32: null pointer");
  const int64_t rows = ceil_div<int64_t>(std::max<int64_t>(G, 1), n_slices);
  const bool vec = (Q % 4 == 0) && (((uintptr_t)dist & 15) == 0) && k <= 32;
  const size_t smem = vec ? (size_t)k * kTopkThreads * 4 * 8 : (size_t)k * kTopkThreads * 8;
  const unsigned bx = (unsigned)ceil_div<int64_t>(Q, kTopkThreads * (vec ? 4 : 1));
  auto launch = [&](int64_t g_rows, int slices, int64_t rows_per, int64_t stride, const float* tau, int tau_slices) -> int {
    if (vec) {
      WITW_CUDA(cudaFuncSetAttribute(topk_columns_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      topk_columns_kernel<4><<<dim3(bx, (unsigned)slices), kTopkThreads, smem, as_stream(stream)>>>(
          dist, g_rows, Q, k, topk_dist, topk_idx, g_offset, rows_per, stride, tau, tau_slices);
    } else {
      WITW_CUDA(cudaFuncSetAttribute(topk_columns_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      topk_columns_kernel<1><<<dim3(bx, (unsigned)slices), kTopkThreads, smem, as_stream(stream)>>>(
          dist, g_rows, Q, k, topk_dist, topk_idx, g_offset, rows_per, stride, tau, tau_slices);
    }
    WITW_LAUNCH_CHECK();
    return WITW_OK;
  };
  // sample pass: 8 slices over a strided sample of the rows (about G/16 of them, at least 1024) written to candidate slots
  // 0..7, merged into slot 8: its k-th entry is the k-th smallest of the whole sample, which admits ~k/n_samples of the
  // elements in the full pass.  The full pass overwrites these slots with its own lists afterwards; whatever a reader finds
  // at a k-th entry is the k-th entry of some full list (or +inf), i.e. a valid bound.
  const int tau_slices = 8;
  if (n_slices > tau_slices && G >= 8192 && rows >= 4 * (int64_t)k) {
    const int64_t n_samples = std::max<int64_t>(1024, G / 16);
    const int64_t stride = std::max<int64_t>(1, G / n_samples);
    const int64_t n_rows = G / stride;
    int rc = launch(n_rows, tau_slices, ceil_div<int64_t>(n_rows, tau_slices), stride, nullptr, 0);
    if (rc != WITW_OK) return rc;
    float* slot8_d = topk_dist + (int64_t)tau_slices * Q * k;
    int32_t* slot8_i = topk_idx + (int64_t)tau_slices * Q * k;
    rc = witw_topk_merge(topk_dist, topk_idx, tau_slices, Q, k, slot8_d, slot8_i, stream);
    if (rc != WITW_OK) return rc;
    return launch(G, n_slices, rows, 1, slot8_d, 1);
  }
  return launch(G, n_slices, rows, 1, nullptr, 0);
}

// ------------------------------------------------------------------------------------------
// top-k of large galleries by threshold + filter + select (witw_topk_select_f32).  The thread-per-column lists above
// diverge on every insertion, and a warp row of 128 columns almost always holds one; so for G >= 8192 the full pass only
// FILTERS: an element at or below the column's threshold tau (an upper bound of its k-th smallest distance, see
// topk_groupmin_kernel) is appended to the column's survivor buffer with one atomic, and a warp per column then selects
// the k smallest survivors in (distance, index) order.  About 4 k (ln k + 0.6) survivors per column are expected from a
// quarter-of-the-rows sample; a column whose buffer overflows (massive ties at tau) is selected from the whole column
// instead -- slow, exact.
// ------------------------------------------------------------------------------------------
constexpr int kFilterThreads = 128;
constexpr int kStash = 4;           // per-thread, per-column survivor slots of the filter pass

// Threshold of a column = max over k disjoint row groups of the group's minimum: k distinct elements lie at or below
// it, so it bounds the k-th smallest distance from above -- from min / max reductions alone, no lists.  Sampled row
// i * row_stride belongs to group i % k; CTA (x, j, s) reduces group j over slice s of the sample.
__global__ void __launch_bounds__(kFilterThreads)
topk_groupmin_kernel(const float* __restrict__ dist, int64_t n_rows, int64_t row_stride, int64_t Q, int k, int64_t rows_per_slice,
                     float* __restrict__ partial /* [slices][k][Q] */) {
  const int64_t q0 = ((int64_t)blockIdx.x * kFilterThreads + threadIdx.x) * 4;
  if (q0 >= Q) return;
  const int j = blockIdx.y;
  const int64_t i0 = (int64_t)blockIdx.z * rows_per_slice, i1 = min(i0 + rows_per_slice, n_rows);
  const float inf = __int_as_float(0x7f800000);
  float4 m = make_float4(inf, inf, inf, inf);
  const float4* p = reinterpret_cast<const float4*>(dist + q0);
  const int64_t stride4 = (Q >> 2) * row_stride;
  int64_t i = i0 + ((j - i0 % k) % k + k) % k;   // first sampled row of group j in this slice
  constexpr int U = 4;
  for (; i + (U - 1) * k < i1; i += U * k) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = __ldcs(p + (i + (int64_t)u * k) * stride4);
#pragma unroll
    for (int u = 0; u < U; ++u) { m.x = fminf(m.x, v[u].x); m.y = fminf(m.y, v[u].y); m.z = fminf(m.z, v[u].z); m.w = fminf(m.w, v[u].w); }
  }
  for (; i < i1; i += k) {
    const float4 v = __ldcs(p + i * stride4);
    m.x = fminf(m.x, v.x); m.y = fminf(m.y, v.y); m.z = fminf(m.z, v.z); m.w = fminf(m.w, v.w);
  }
  *reinterpret_cast<float4*>(partial + ((int64_t)blockIdx.z * k + j) * Q + q0) = m;
}

__global__ void __launch_bounds__(256)
topk_tau_kernel(const float* __restrict__ partial, int n_slices, int k, int64_t Q, float* __restrict__ tau) {
  const int64_t q = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (q >= Q) return;
  const float inf = __int_as_float(0x7f800000);
  float t = -inf;
  for (int j = 0; j < k; ++j) {
    float m = inf;
    for (int s = 0; s < n_slices; ++s) m = fminf(m, partial[((int64_t)s * k + j) * Q + q]);
    t = fmaxf(t, m);
  }
  tau[q] = t;
}

__global__ void __launch_bounds__(kFilterThreads)
topk_filter_kernel(const float* __restrict__ dist, int64_t G, int64_t Q, const float* __restrict__ tau,
                   int64_t rows_per_slice, int cap, int32_t* __restrict__ cnt, float* __restrict__ buf_d,
                   int32_t* __restrict__ buf_i) {
  const int64_t q0 = ((int64_t)blockIdx.x * kFilterThreads + threadIdx.x) * 4;
  if (q0 >= Q) return;
  const float inf = __int_as_float(0x7f800000);
  float bound[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {  // admit d <= tau, written as d < (next float above tau); a NaN or missing tau admits every finite value
    const bool live = q0 + c < Q;
    const float t = live ? tau[q0 + c] : 0.f;
    bound[c] = t < inf ? __uint_as_float(__float_as_uint(t) + (t >= 0.f ? 1u : 0xffffffffu)) : inf;
    if (t == 0.f) bound[c] = __uint_as_float(1u);
    if (!live) bound[c] = -inf;
  }
  const int64_t g0 = (int64_t)blockIdx.y * rows_per_slice;
  const int64_t g1 = min(g0 + rows_per_slice, G);
  // Survivors are first kept in the thread's own slots of shared memory (kStash per column; a slice holds one or two per
  // column) and flushed with one atomic per column at the end: the atomic-per-survivor of the first version stalled the warp
  // for the returned position on 78 % of the rows.  A column that fills its slots falls back to that path.
  __shared__ float stash_d[kStash][4][kFilterThreads];
  __shared__ int32_t stash_i[kStash][4][kFilterThreads];
  int n[4] = {0, 0, 0, 0};
  const int tid = threadIdx.x;
  auto keep = [&](int c, float d, int64_t g) {
    if (n[c] < kStash) {
      stash_d[n[c]][c][tid] = d;
      stash_i[n[c]][c][tid] = (int32_t)g;
      ++n[c];
    } else {
      const int32_t pos = atomicAdd(cnt + q0 + c, 1);
      if (pos < cap) {
        buf_d[(q0 + c) * cap + pos] = d;
        buf_i[(q0 + c) * cap + pos] = (int32_t)g;
      }
    }
  };
  constexpr int U = 4;
  const float4* p = reinterpret_cast<const float4*>(dist + q0);
  const int64_t stride4 = Q >> 2;
  int64_t g = g0;
  for (; g + U <= g1; g += U) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = __ldcs(p + (g + u) * stride4);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (v[u].x < bound[0]) keep(0, v[u].x, g + u);
      if (v[u].y < bound[1]) keep(1, v[u].y, g + u);
      if (v[u].z < bound[2]) keep(2, v[u].z, g + u);
      if (v[u].w < bound[3]) keep(3, v[u].w, g + u);
    }
  }
  for (; g < g1; ++g) {
    const float4 v = __ldcs(p + g * stride4);
    if (v.x < bound[0]) keep(0, v.x, g);
    if (v.y < bound[1]) keep(1, v.y, g);
    if (v.z < bound[2]) keep(2, v.z, g);
    if (v.w < bound[3]) keep(3, v.w, g);
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if (n[c] > 0) {
      const int32_t pos = atomicAdd(cnt + q0 + c, n[c]);
      for (int j = 0; j < n[c]; ++j) {
        if (pos + j < cap) {
          buf_d[(q0 + c) * cap + pos + j] = stash_d[j][c][tid];
          buf_i[(q0 + c) * cap + pos + j] = stash_i[j][c][tid];
        }
      }
    }
  }
}

// (d, i) < (e, j) in the order of a stable ascending sort
__device__ __forceinline__ bool pair_less(float d, int32_t i, float e, int32_t j) { return d < e || (d == e && i < j); }

// Smallest (d, i) pair strictly after (last_d, last_i) among m entries of a column -- survivors from the buffer, or the
// column itself (whole == true) -- reduced over the warp.  NaN and +inf never qualify.
__device__ __forceinline__ void select_next(bool whole, const float* __restrict__ col, int64_t Q, const float* __restrict__ bd_buf,
                                            const int32_t* __restrict__ bi_buf, int64_t m, int lane, bool first, float last_d,
                                            int32_t last_i, float* out_d, int32_t* out_i) {
  const float inf = __int_as_float(0x7f800000);
  float bd = inf;
  int32_t bi = 0x7fffffff;
  for (int64_t e = lane; e < m; e += 32) {
    const float d = whole ? __ldg(col + e * Q) : bd_buf[e];
    const int32_t i = whole ? (int32_t)e : bi_buf[e];
    if (d < inf && (first || pair_less(last_d, last_i, d, i)) && pair_less(d, i, bd, bi)) { bd = d; bi = i; }
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    const float od = __shfl_xor_sync(0xffffffffu, bd, s);
    const int32_t oi = __shfl_xor_sync(0xffffffffu, bi, s);
    if (pair_less(od, oi, bd, bi)) { bd = od; bi = oi; }
  }
  *out_d = bd;
  *out_i = bi;
}

__global__ void __launch_bounds__(256)
topk_select_kernel(const float* __restrict__ dist, int64_t G, int64_t Q, int k, int cap, const int32_t* __restrict__ cnt,
                   float* __restrict__ buf_d, int32_t* __restrict__ buf_i, int32_t g_offset,
                   float* __restrict__ out_d, int32_t* __restrict__ out_i) {
  const int lane = threadIdx.x & 31;
  const int64_t q = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (q >= Q) return;
  const float inf = __int_as_float(0x7f800000);
  float* bd_buf = buf_d + q * cap;
  int32_t* bi_buf = buf_i + q * cap;
  const float* col = dist + q;
  int64_t m = cnt[q];
  bool whole = false;
  if (m > cap) {
    // The buffer holds only the first `cap` survivors.  Their k-th smallest is still an upper bound of the column's
    // k-th smallest (a subset's order statistics are no smaller), and a much tighter one: rescan the column once with
    // it, refilling the buffer.  Only if that overflows too (> cap ties) is every pick taken from the whole column.
    float td = -inf;
    int32_t ti = -1;
    bool first = true;
    for (int j = 0; j < k; ++j) {
      select_next(false, col, Q, bd_buf, bi_buf, cap, lane, first, td, ti, &td, &ti);
      first = false;
      if (!(td < inf)) break;
    }
    __syncwarp();
    int64_t n2 = 0;
    constexpr int RU = 8;  // column loads in flight per lane (the column is strided by Q: one sector per element)
    for (int64_t e0 = 0; e0 < G; e0 += 32 * RU) {
      float v[RU];
#pragma unroll
      for (int u = 0; u < RU; ++u) {
        const int64_t e = e0 + 32 * u + lane;
        v[u] = e < G ? __ldg(col + e * Q) : inf;
      }
#pragma unroll
      for (int u = 0; u < RU; ++u) {
        const bool keep = v[u] <= td && v[u] < inf;
        const unsigned mask = __ballot_sync(0xffffffffu, keep);
        const int64_t pos = n2 + __popc(mask & ((1u << lane) - 1u));
        if (keep && pos < cap) { bd_buf[pos] = v[u]; bi_buf[pos] = (int32_t)(e0 + 32 * u + lane); }
        n2 += __popc(mask);
      }
    }
    __syncwarp();
    m = n2;
    if (n2 > cap) { whole = true; m = G; }
  }
  float last_d = -inf;
  int32_t last_i = -1;
  bool first = true;
  for (int j = 0; j < k; ++j) {
    float bd;
    int32_t bi;
    select_next(whole, col, Q, bd_buf, bi_buf, m, lane, first, last_d, last_i, &bd, &bi);
    if (lane == 0) {
      out_d[q * k + j] = bd;
      out_i[q * k + j] = bd < inf ? bi + g_offset : -1;
    }
    if (!(bd < inf)) {  // exhausted: pad the rest
      for (int r = j + 1 + lane; r < k; r += 32) { out_d[q * k + r] = inf; out_i[q * k + r] = -1; }
      break;
    }
    last_d = bd; last_i = bi; first = false;
  }
}

extern "C" int witw_topk_merge(const float* cand_dist, const int32_t* cand_idx, int n_lists, int64_t Q, int k, float* topk_dist,
                               int32_t* topk_idx, witw_stream_t stream) {
  WITW_REQUIRE(n_lists > 0 && n_lists <= 64 && Q >= 0 && k > 0 && k <= 128, WITW_ERR_INVALID, "witw_topk_merge: bad shape (n_lists 1..64, k 1..128)");
  if (Q == 0) return WITW_OK;
  WITW_REQUIRE(cand_dist && cand_idx && topk_dist && topk_idx, WITW_ERR_INVALID, "witw_topk_merge: null pointer");
  WITW_REQUIRE(ceil_div<int64_t>(Q, kMergeWarps) < (1ll << 31), WITW_ERR_INVALID, "witw_topk_merge: too many queries");
  topk_merge_kernel<<<(unsigned)ceil_div<int64_t>(Q, kMergeWarps), kMergeWarps * 32, 0, as_stream(stream)>>>(cand_dist, cand_idx, n_lists, Q, k,
                                                                                                       topk_dist, topk_idx);
  WITW_LAUNCH_CHECK();
  return WITW_OK;
}

static int select_cap(int k) { return std::max(128, 48 * k); }

constexpr int kTauSlices = 4;

extern "C" size_t witw_topk_select_scratch_bytes(int64_t Q, int k) {
  const int64_t q = std::max<int64_t>(Q, 1);
  // counters, survivor buffers (distance + index), group minima [slices][k][Q], thresholds [Q]
  return (size_t)(q * 4 + q * select_cap(k) * 8 + (int64_t)kTauSlices * k * q * 4 + q * 4 + 1024);
}

extern "C" int witw_topk_select_f32(const float* dist, int64_t G, int64_t Q, int k, float* topk_dist, int32_t* topk_idx,
                                    int32_t g_offset, void* scratch, witw_stream_t stream) {
  WITW_REQUIRE(G >= 1024 && Q >= 0 && k > 0 && k <= 32, WITW_ERR_UNSUPPORTED, "witw_topk_select_f32: needs G >= 1024 and 1 <= k <= 32");
  WITW_REQUIRE(Q % 4 == 0 && (((uintptr_t)dist & 15) == 0), WITW_ERR_UNSUPPORTED, "witw_topk_select_f32: Q must be a multiple of 4 and dist 16-byte aligned");
  if (Q == 0) return WITW_OK;
  WITW_REQUIRE(dist && topk_dist && topk_idx && scratch && (((uintptr_t)scratch & 15) == 0), WITW_ERR_INVALID, "witw_topk_select_f32: null or misaligned pointer");
  const int cap = select_cap(k);
  char* base = reinterpret_cast<char*>(scratch);
  auto take = [&](size_t bytes) { char* p = base; base += (bytes + 255) & ~(size_t)255; return p; };
  int32_t* cnt = reinterpret_cast<int32_t*>(take((size_t)Q * 4));
  float* buf_d = reinterpret_cast<float*>(take((size_t)Q * cap * 4));
  int32_t* buf_i = reinterpret_cast<int32_t*>(take((size_t)Q * cap * 4));
  float* partial = reinterpret_cast<float*>(take((size_t)kTauSlices * k * Q * 4));
  float* tau = reinterpret_cast<float*>(take((size_t)Q * 4));
  // 1. thresholds from every fourth row: k group minima per column, their maximum
  const int64_t stride = 4;
  const int64_t n_rows = G / stride;
  const int64_t bx = ceil_div<int64_t>(Q, kFilterThreads * 4);
  const int64_t rows_tau = ceil_div<int64_t>(n_rows, kTauSlices);
  topk_groupmin_kernel<<<dim3((unsigned)bx, (unsigned)k, kTauSlices), kFilterThreads, 0, as_stream(stream)>>>(dist, n_rows, stride, Q, k, rows_tau, partial);
  WITW_LAUNCH_CHECK();
  topk_tau_kernel<<<(unsigned)ceil_div<int64_t>(Q, 256), 256, 0, as_stream(stream)>>>(partial, kTauSlices, k, Q, tau);
  WITW_LAUNCH_CHECK();
  // 2. filter the whole matrix into the survivor buffers
  WITW_CUDA(cudaMemsetAsync(cnt, 0, (size_t)Q * 4, as_stream(stream)));
  int64_t slices = std::max<int64_t>(1, std::min<int64_t>(ceil_div<int64_t>((int64_t)sm_count() * 12, bx), G / 64));
  slices = std::min<int64_t>(slices, 65535);
  const int64_t rows = ceil_div<int64_t>(G, slices);
  topk_filter_kernel<<<dim3((unsigned)bx, (unsigned)ceil_div<int64_t>(G, rows)), kFilterThreads, 0, as_stream(stream)>>>(
      dist, G, Q, tau, rows, cap, cnt, buf_d, buf_i);
  WITW_LAUNCH_CHECK();
  // 3. k smallest survivors per column
  topk_select_kernel<<<(unsigned)ceil_div<int64_t>(Q, 8), 256, 0, as_stream(stream)>>>(dist, G, Q, k, cap, cnt, buf_d, buf_i, g_offset,
                                                                                      topk_dist, topk_idx);
  WITW_LAUNCH_CHECK();
  return WITW_OK;
}
