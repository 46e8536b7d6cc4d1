// The fp32 finish of the tensor-core sweeps (match_spec.cu, match_tc.cu): everything the fp16 sweep could not settle
// within its error bound (sweep_common.cuh) is settled here from the fp32 azimuth spectra, so that ranks, top-k and
// matrix outputs are those of the fp32 reference chain (correlation -> crop_overhead -> l2_distance, cvig_fov.py:297-363;
// rank rule cvig_fov.py:552).  Three small kernels behind one entry point (witw_finish_spec_f32):
//
//   finish_scan_kernel   prefix sum of the per-query deferral counts (so that the deferred pairs of all queries form one
//                        task list); a query whose list overflowed is flagged and contributes no tasks: the caller re-does
//                        it with columns_kernel.
//   finish_pairs_kernel  one warp per task, persistent grid: every deferred pair and the first k_out top-k candidates of
//                        every query are evaluated exactly.  Pending rank decisions are added to rank_count, matrix entries
//                        overwritten with their fp32 values, candidate distances stored.  (A CTA per query with the query's
//                        spectrum staged in shared memory was 3x slower: two passes of 8 + 2 busy warps, a block-wide
//                        barrier per round and 4 CTAs per SM; the spectra of a query's tasks hit in L1 / L2 anyway.)
//   finish_topk_kernel   one warp per query: the k_out exact distances bound the k_out-th smallest distance of the gallery;
//                        further candidates are evaluated only if their lower-bound key can still reach it (rare), the
//                        evaluated candidates are ranked by (distance, index), and the query is flagged when the keys do not
//                        prove that no item outside the candidate list belongs to the top k_out.
//   columns_kernel       exact fp32 distances / orientations of selected queries against every gallery item: the whole
//                        answer for a handful of queries (heat map, the reference's one-query loop) and the fallback for
//                        flagged queries.
#include <algorithm>
#include <cstring>

#include "common.cuh"
#include "spectral_pair.cuh"
#include "sweep_common.cuh"

namespace witw {

constexpr int kFinThreads = 256;
constexpr int kFinWarps = kFinThreads / 32;
constexpr int kFinMaxCand = 32;

struct FinishParams {
  const float2* gal_spec;        // [G][CH][32]
  const float* crop_inv_norm;    // [G][64]
  const float2* qry_spec;        // [Q][CH][32]
  const float* q_inv_norm;       // [Q]
  int64_t G, Q;
  int CH;
  int32_t g_offset;
  const int32_t* list_g;         // [Q][list_cap] or null
  const int32_t* list_n;         // [Q]
  int32_t list_cap;
  const float* d_true;           // [Q] or null
  int32_t* rank_count;           // [Q] or null
  float* dist;                   // [G][Q] or null
  uint8_t* ori;                  // [G][Q] or null
  const float* cand_key;         // [Q][kc] or null
  const int32_t* cand_idx;       // [Q][kc] global indices, -1 = empty
  int kc, k_out;
  float* out_dist;               // [Q][k_out]
  int32_t* out_idx;
  int32_t* qflag;                // [Q], zeroed by the caller: bit 0 = list overflow, bit 1 = top-k not proven
  int32_t* n_flagged;            // [1], zeroed by the caller
  int32_t* offsets;              // scratch [Q + 1]: first task of every query's deferred pairs
  float* cand_exact;             // scratch [Q][kc]: exact distances of the candidates (+inf: not evaluated / invalid)
};

// One CTA: offsets[q] = number of deferred pairs of the queries before q (overflowed lists count as empty and are flagged).
__global__ void __launch_bounds__(1024)
finish_scan_kernel(const FinishParams P) {
  __shared__ int32_t warp_sum[32];
  __shared__ int32_t carry_sh;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_sh = 0;
  __syncthreads();
  for (int64_t q0 = 0; q0 < P.Q; q0 += 1024) {
    const int64_t q = q0 + threadIdx.x;
    int32_t n = 0;
    if (q < P.Q && P.list_g != nullptr) {
      n = P.list_n[q];
      if (n > P.list_cap) {       // the sweep could not record everything it deferred for this query
        n = 0;
        P.qflag[q] |= 1;
        atomicAdd(P.n_flagged, 1);
      }
    }
    int32_t incl = n;
    for (int m = 1; m < 32; m <<= 1) {
      const int32_t o = __shfl_up_sync(0xffffffffu, incl, m);
      if (lane >= m) incl += o;
    }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int32_t w = warp_sum[lane];
      for (int m = 1; m < 32; m <<= 1) {
        const int32_t o = __shfl_up_sync(0xffffffffu, w, m);
        if (lane >= m) w += o;
      }
      warp_sum[lane] = w;
    }
    __syncthreads();
    const int32_t base = carry_sh + (warp ? warp_sum[warp - 1] : 0);
    if (q < P.Q) P.offsets[q] = base + incl - n;
    __syncthreads();
    if (threadIdx.x == 1023) carry_sh = base + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) P.offsets[P.Q] = carry_sh;
}

__global__ void __launch_bounds__(kFinThreads)
finish_pairs_kernel(const FinishParams P) {
  __shared__ float2 tw[64];
  spectral_twiddles(tw);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t n_list = P.offsets[P.Q];
  const int64_t n_tasks = n_list + (int64_t)P.Q * P.k_out;
  const int64_t stride = (int64_t)gridDim.x * kFinWarps;
  for (int64_t t = (int64_t)blockIdx.x * kFinWarps + (threadIdx.x >> 5); t < n_tasks; t += stride) {
    int64_t q, g;
    uint32_t entry = 0;
    int j = -1;
    if (t < n_list) {             // a deferred pair: find its query (last q with offsets[q] <= t)
      int64_t lo = 0, hi = P.Q;
      while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (P.offsets[mid] <= t) lo = mid; else hi = mid;
      }
      q = lo;
      entry = (uint32_t)P.list_g[q * P.list_cap + (t - P.offsets[q])];
      g = entry & kTagMask;
    } else {                      // one of the first k_out top-k candidates of a query
      const int64_t c = t - n_list;
      q = c / P.k_out;
      j = (int)(c - q * P.k_out);
      const int32_t idx = P.cand_idx[q * P.kc + j];
      g = (int64_t)idx - P.g_offset;
      if (idx < 0 || g < 0 || g >= P.G) {
        if (lane == 0) P.cand_exact[q * P.kc + j] = __int_as_float(0x7f800000);
        continue;
      }
    }
    const PairMax r = spectral_pair_eval(P.gal_spec + g * P.CH * 32 + lane, P.qry_spec + q * P.CH * 32 + lane, P.CH, tw, lane);
    if (lane == 0) {
      const float d = 2.0f * (1.0f - r.best * P.crop_inv_norm[g * 64 + r.arg] * P.q_inv_norm[q]);
      if (j >= 0) {
        P.cand_exact[q * P.kc + j] = (d == d) ? d : __int_as_float(0x7f800000);      // NaN never enters a top-k
      } else {
        if ((entry & kTagRank) && d <= P.d_true[q]) atomicAdd(P.rank_count + q, 1);
        if (P.dist) P.dist[g * P.Q + q] = d;
        if (P.ori) P.ori[g * P.Q + q] = (uint8_t)r.arg;
      }
    }
  }
}

// One warp per query, lane = candidate.
__global__ void __launch_bounds__(kFinThreads)
finish_topk_kernel(const FinishParams P) {
  __shared__ float2 tw[64];
  spectral_twiddles(tw);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t q = (int64_t)blockIdx.x * kFinWarps + (threadIdx.x >> 5);
  if (q >= P.Q) return;
  const float inf = __int_as_float(0x7f800000);
  const bool have = lane < P.kc;
  const float key = have ? P.cand_key[q * P.kc + lane] : inf;
  const int32_t idx = have ? P.cand_idx[q * P.kc + lane] : -1;
  float ex = (lane < P.k_out) ? P.cand_exact[q * P.kc + lane] : inf;
  // the k_out exact distances bound the k_out-th smallest distance of the gallery from above
  float bound = (lane < P.k_out) ? ex : -inf;
  for (int m = 16; m > 0; m >>= 1) bound = fmaxf(bound, __shfl_xor_sync(0xffffffffu, bound, m));
  const int64_t g_mine = (int64_t)idx - P.g_offset;
  const bool want = have && lane >= P.k_out && idx >= 0 && g_mine >= 0 && g_mine < P.G && key <= bound;
  unsigned todo = __ballot_sync(0xffffffffu, want);
  while (todo) {                  // rare: a candidate beyond the first k_out whose lower bound can still reach the top
    const int j = __ffs(todo) - 1;
    todo &= todo - 1;
    const int64_t g = __shfl_sync(0xffffffffu, g_mine, j);
    const PairMax r = spectral_pair_eval(P.gal_spec + g * P.CH * 32 + lane, P.qry_spec + q * P.CH * 32 + lane, P.CH, tw, lane);
    const float d = 2.0f * (1.0f - r.best * P.crop_inv_norm[g * 64 + r.arg] * P.q_inv_norm[q]);
    if (lane == j) ex = (d == d) ? d : inf;
  }
  // position of every evaluated candidate in (distance, index) order
  int pos = 0;
  for (int o = 0; o < P.kc; ++o) {
    const float od = __shfl_sync(0xffffffffu, ex, o);
    const int32_t oi = __shfl_sync(0xffffffffu, idx, o);
    pos += (od < ex || (od == ex && oi < idx)) ? 1 : 0;
  }
  const bool valid = ex < inf;
  const int n_valid = __popc(__ballot_sync(0xffffffffu, valid));
  if (valid && pos < P.k_out) {
    P.out_dist[q * P.k_out + pos] = ex;
    P.out_idx[q * P.k_out + pos] = idx;
  }
  if (lane >= n_valid && lane < P.k_out) {
    P.out_dist[q * P.k_out + lane] = inf;
    P.out_idx[q * P.k_out + lane] = -1;
  }
  // the k_out-th smallest evaluated distance; every item outside the list has a key >= the list's last key, and its exact
  // distance is >= its key: the list is proven complete when that key exceeds it
  float kth = (valid && pos == P.k_out - 1) ? ex : -inf;
  for (int m = 16; m > 0; m >>= 1) kth = fmaxf(kth, __shfl_xor_sync(0xffffffffu, kth, m));
  if (n_valid < P.k_out) kth = inf;
  const float last_key = __shfl_sync(0xffffffffu, key, P.kc - 1);
  const int32_t last_idx = __shfl_sync(0xffffffffu, idx, P.kc - 1);
  if (lane == 0 && last_idx >= 0 && last_key <= kth && !(P.qflag[q] & 1)) {
    P.qflag[q] |= 2;
    atomicAdd(P.n_flagged, 1);
  }
}

constexpr int kColItems = 32;   // gallery items per CTA of the column kernel

__device__ __forceinline__ void stage_query(float2* sq, const float2* __restrict__ src, int n2) {
  const float4* s4 = reinterpret_cast<const float4*>(src);
  float4* d4 = reinterpret_cast<float4*>(sq);
  for (int i = threadIdx.x; i < n2 / 2; i += blockDim.x) d4[i] = __ldg(s4 + i);
}

struct ColumnParams {
  const float2* gal_spec;
  const float* crop_inv_norm;
  const float2* qry_spec;
  const float* q_inv_norm;
  int64_t G;
  int CH;
  const int32_t* q_sel;          // [F] query indices, or null for 0..F-1
  float* dist;                   // element (g, column) at dist[g * ld + column], or null
  int64_t* ori64;                // same indexing, or null
  uint8_t* ori8;                 // same indexing, or null
  int64_t ld;
  int col_is_q;                  // column = the query index (writing into a [G,Q] matrix) instead of 0..F-1
  const float* d_true;           // [Q] thresholds for count_out, or null
  const int32_t* true_idx;       // [Q] global index of the match (counted by index), or null
  int32_t g_offset;
  int32_t* count_out;            // [F] += #{g : d <= d_true}, or null
};

__global__ void __launch_bounds__(kFinThreads)
columns_kernel(const ColumnParams P) {
  extern __shared__ __align__(16) float2 sq[];
  __shared__ float2 tw[64];
  const int f = blockIdx.y;
  const int64_t q = P.q_sel ? P.q_sel[f] : f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  spectral_twiddles(tw);
  stage_query(sq, P.qry_spec + q * P.CH * 32, P.CH * 32);
  __syncthreads();
  const float qin = P.q_inv_norm[q];
  const int64_t col = P.col_is_q ? q : f;
  const float dtrue = (P.count_out && P.d_true) ? P.d_true[q] : __int_as_float(0x7fc00000);
  const int64_t self_g = (P.count_out && P.true_idx) ? (int64_t)P.true_idx[q] - P.g_offset : -1;
  const int64_t g0 = (int64_t)blockIdx.x * kColItems;
  const int64_t g1 = min(g0 + kColItems, P.G);
  int cnt = 0;
  for (int64_t g = g0 + warp; g < g1; g += kFinWarps) {
    const PairMax r = spectral_pair_eval(P.gal_spec + g * P.CH * 32 + lane, sq + lane, P.CH, tw, lane);
    if (lane == 0) {
      const float d = 2.0f * (1.0f - r.best * P.crop_inv_norm[g * 64 + r.arg] * qin);
      if (P.dist) P.dist[g * P.ld + col] = d;
      if (P.ori64) P.ori64[g * P.ld + col] = r.arg;
      if (P.ori8) P.ori8[g * P.ld + col] = (uint8_t)r.arg;
      if (g == self_g) cnt += (dtrue == dtrue) ? 1 : 0;
      else cnt += (d <= dtrue) ? 1 : 0;
    }
  }
  if (P.count_out && lane == 0 && cnt) atomicAdd(P.count_out + f, cnt);
}

}  // namespace witw

using namespace witw;

extern "C" size_t witw_finish_scratch_bytes(int64_t Q, int kc) {
  const int64_t q = std::max<int64_t>(Q, 1);
  return (size_t)(((q + 1) * 4 + 15) / 16 * 16 + q * std::max(kc, 1) * 4);
}

extern "C" int witw_finish_spec_f32(const witw_finish_args* a, witw_stream_t stream) {
  WITW_REQUIRE(a != nullptr, WITW_ERR_INVALID, "witw_finish_spec_f32: null arguments");
  WITW_REQUIRE(a->G >= 0 && a->Q >= 0 && a->CH > 0 && a->list_cap >= 0, WITW_ERR_INVALID, "witw_finish_spec_f32: bad shape");
  if (a->Q == 0 || a->G == 0) return WITW_OK;
  WITW_REQUIRE(a->gal_spec && a->crop_inv_norm && a->qry_spec && a->q_inv_norm && a->qflag && a->n_flagged && a->scratch, WITW_ERR_INVALID,
               "witw_finish_spec_f32: null pointer");
  WITW_REQUIRE(a->list_cap == 0 || (a->list_g && a->list_n), WITW_ERR_INVALID, "witw_finish_spec_f32: deferral list buffers missing");
  WITW_REQUIRE(a->kc >= 0 && a->kc <= kFinMaxCand && (a->kc == 0 || (a->k_out >= 1 && a->k_out <= a->kc && a->cand_key && a->cand_idx && a->out_dist && a->out_idx)),
               WITW_ERR_INVALID, "witw_finish_spec_f32: need 1 <= k_out <= kc <= %d and the candidate / output buffers", kFinMaxCand);
  WITW_REQUIRE((((uintptr_t)a->gal_spec | (uintptr_t)a->qry_spec) & 7) == 0 && ((uintptr_t)a->scratch & 15) == 0, WITW_ERR_INVALID,
               "witw_finish_spec_f32: spectra must be 8-byte, scratch 16-byte aligned");
  WITW_REQUIRE(a->Q < (1ll << 31) && a->Q * (int64_t)std::max(a->list_cap, 1) < (1ll << 31), WITW_ERR_INVALID, "witw_finish_spec_f32: too many queries / list entries");
  FinishParams P;
  std::memset(&P, 0, sizeof(P));
  P.gal_spec = reinterpret_cast<const float2*>(a->gal_spec); P.crop_inv_norm = a->crop_inv_norm;
  P.qry_spec = reinterpret_cast<const float2*>(a->qry_spec); P.q_inv_norm = a->q_inv_norm;
  P.G = a->G; P.Q = a->Q; P.CH = a->CH; P.g_offset = a->g_index_offset;
  P.list_g = a->list_cap > 0 ? a->list_g : nullptr; P.list_n = a->list_n; P.list_cap = a->list_cap;
  P.d_true = a->d_true; P.rank_count = a->rank_count; P.dist = a->dist; P.ori = a->ori;
  P.cand_key = a->cand_key; P.cand_idx = a->cand_idx; P.kc = a->kc; P.k_out = a->kc > 0 ? a->k_out : 0; P.out_dist = a->out_dist; P.out_idx = a->out_idx;
  P.qflag = a->qflag; P.n_flagged = a->n_flagged;
  P.offsets = reinterpret_cast<int32_t*>(a->scratch);
  P.cand_exact = reinterpret_cast<float*>(reinterpret_cast<char*>(a->scratch) + ((a->Q + 1) * 4 + 15) / 16 * 16);
  WITW_REQUIRE(P.list_g == nullptr || (P.d_true && P.rank_count) || P.dist || P.ori, WITW_ERR_INVALID,
               "witw_finish_spec_f32: deferred pairs need somewhere to go (d_true + rank_count, or dist / ori)");
  finish_scan_kernel<<<1, 1024, 0, as_stream(stream)>>>(P);
  WITW_LAUNCH_CHECK();
  const int grid = std::max(1, sm_count()) * 8;
  finish_pairs_kernel<<<grid, kFinThreads, 0, as_stream(stream)>>>(P);
  WITW_LAUNCH_CHECK();
  if (P.kc > 0) {
    finish_topk_kernel<<<(unsigned)ceil_div<int64_t>(a->Q, kFinWarps), kFinThreads, 0, as_stream(stream)>>>(P);
    WITW_LAUNCH_CHECK();
  }
  return WITW_OK;
}

extern "C" int witw_match_columns_spec_f32(const float* gal_spec, const float* crop_inv_norm, const float* qry_spec, const float* q_inv_norm,
                                           int64_t G, int CH, const int32_t* q_sel, int64_t F, float* dist, int64_t* ori64, uint8_t* ori8,
                                           int64_t ld, int col_is_q, const float* d_true, const int32_t* true_idx, int32_t g_index_offset,
                                           int32_t* count_out, witw_stream_t stream) {
  WITW_REQUIRE(G >= 0 && F >= 0 && CH > 0 && ld >= 0, WITW_ERR_INVALID, "witw_match_columns_spec_f32: bad shape");
  if (G == 0 || F == 0) return WITW_OK;
  WITW_REQUIRE(gal_spec && crop_inv_norm && qry_spec && q_inv_norm, WITW_ERR_INVALID, "witw_match_columns_spec_f32: null pointer");
  WITW_REQUIRE(!count_out || d_true, WITW_ERR_INVALID, "witw_match_columns_spec_f32: count_out needs d_true");
  WITW_REQUIRE((((uintptr_t)gal_spec | (uintptr_t)qry_spec) & 15) == 0, WITW_ERR_INVALID, "witw_match_columns_spec_f32: spectra must be 16-byte aligned");
  WITW_REQUIRE(F <= 65535, WITW_ERR_UNSUPPORTED, "witw_match_columns_spec_f32: at most 65535 columns per call (got %lld)", (long long)F);
  const int64_t gx = ceil_div<int64_t>(G, kColItems);
  WITW_REQUIRE(gx < (1ll << 31), WITW_ERR_INVALID, "witw_match_columns_spec_f32: gallery too large");
  ColumnParams P;
  std::memset(&P, 0, sizeof(P));
  P.gal_spec = reinterpret_cast<const float2*>(gal_spec); P.crop_inv_norm = crop_inv_norm;
  P.qry_spec = reinterpret_cast<const float2*>(qry_spec); P.q_inv_norm = q_inv_norm;
  P.G = G; P.CH = CH; P.q_sel = q_sel; P.dist = dist; P.ori64 = ori64; P.ori8 = ori8; P.ld = ld; P.col_is_q = col_is_q;
  P.d_true = d_true; P.true_idx = true_idx; P.g_offset = g_index_offset; P.count_out = count_out;
  const size_t smem = (size_t)CH * 32 * sizeof(float2);
  WITW_REQUIRE(smem <= 200 * 1024, WITW_ERR_UNSUPPORTED, "witw_match_columns_spec_f32: a spectrum of %d rows does not fit shared memory", CH);
  WITW_CUDA(cudaFuncSetAttribute(columns_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  columns_kernel<<<dim3((unsigned)gx, (unsigned)F), kFinThreads, smem, as_stream(stream)>>>(P);
  WITW_LAUNCH_CHECK();
  return WITW_OK;
}
