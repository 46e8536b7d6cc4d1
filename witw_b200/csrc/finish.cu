// The fp32 finish of the tensor-core sweeps (match_spec.cu, match_tc.cu): everything the fp16 sweep could not settle
// within its error bound (sweep_common.cuh) is settled here from the fp32 azimuth spectra, so that ranks, top-k and
// matrix outputs are those of the fp32 reference chain (correlation -> crop_overhead -> l2_distance, cvig_fov.py:297-363;
// rank rule cvig_fov.py:552).
//
//   finish_kernel    one CTA per query.  The query's fp32 spectrum is staged in shared memory once; then
//                    (1) the pairs of the query's deferral list are evaluated exactly: pending rank decisions are added
//                        to rank_count, matrix entries are overwritten with their fp32 values;
//                    (2) the sweep's top-k candidates (ascending lower-bound keys) are re-ranked: the first k_out exactly,
//                        then only those whose key can still reach the k_out-th exact distance; the query is flagged when
//                        the keys do not prove that no item outside the candidate list belongs to the top k.
//                    A query whose list overflowed is flagged and skipped: the caller re-does it with columns_kernel.
//   columns_kernel   exact fp32 distances / orientations of selected queries against every gallery item: the whole
//                    answer for a handful of queries (heat map, the reference's one-query loop) and the fallback for
//                    flagged queries.
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
};

__device__ __forceinline__ void stage_query(float2* sq, const float2* __restrict__ src, int n2) {
  const float4* s4 = reinterpret_cast<const float4*>(src);
  float4* d4 = reinterpret_cast<float4*>(sq);
  for (int i = threadIdx.x; i < n2 / 2; i += blockDim.x) d4[i] = __ldg(s4 + i);
}

__global__ void __launch_bounds__(kFinThreads)
finish_kernel(const FinishParams P) {
  extern __shared__ __align__(16) float2 sq[];      // the query's spectrum, CH x 32 slots
  __shared__ float2 tw[64];
  __shared__ float ex[kFinMaxCand];                  // exact distances of the candidates (+inf: not evaluated / invalid)
  __shared__ float bound;
  __shared__ int cnt_sh;
  const int64_t q = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_list = P.list_g ? P.list_n[q] : 0;
  if (n_list > P.list_cap) {                         // the sweep could not record everything it deferred
    if (threadIdx.x == 0) { P.qflag[q] |= 1; atomicAdd(P.n_flagged, 1); }
    return;
  }
  if (n_list == 0 && P.kc == 0) return;
  spectral_twiddles(tw);
  if (threadIdx.x == 0) cnt_sh = 0;
  stage_query(sq, P.qry_spec + q * P.CH * 32, P.CH * 32);
  __syncthreads();
  const float qin = P.q_inv_norm[q];
  const float inf = __int_as_float(0x7f800000);

  // ---- (1) deferred pairs
  if (n_list > 0) {
    const float dtrue = P.d_true ? P.d_true[q] : __int_as_float(0x7fc00000);
    int cnt = 0;
    for (int i = warp; i < n_list; i += kFinWarps) {
      const uint32_t entry = (uint32_t)P.list_g[q * P.list_cap + i];
      const int64_t g = entry & kTagMask;
      const PairMax r = spectral_pair_eval(P.gal_spec + g * P.CH * 32 + lane, sq + lane, P.CH, tw, lane);
      if (lane == 0) {
        const float d = 2.0f * (1.0f - r.best * P.crop_inv_norm[g * 64 + r.arg] * qin);
        if (entry & kTagRank) cnt += (d <= dtrue) ? 1 : 0;
        if (P.dist) P.dist[g * P.Q + q] = d;
        if (P.ori) P.ori[g * P.Q + q] = (uint8_t)r.arg;
      }
    }
    if (lane == 0 && cnt) atomicAdd(&cnt_sh, cnt);
  }

  // ---- (2) top-k candidates
  if (P.kc > 0) {
    const float* key = P.cand_key + q * P.kc;
    const int32_t* idx = P.cand_idx + q * P.kc;
    for (int round = 0; round < 2; ++round) {
      const int j0 = round == 0 ? 0 : P.k_out, j1 = round == 0 ? P.k_out : P.kc;
      const float reach = round == 0 ? inf : bound;
      for (int j = j0 + warp; j < j1; j += kFinWarps) {
        const int64_t g = (int64_t)idx[j] - P.g_offset;
        float d = inf;
        if (idx[j] >= 0 && g >= 0 && g < P.G && key[j] <= reach) {
          const PairMax r = spectral_pair_eval(P.gal_spec + g * P.CH * 32 + lane, sq + lane, P.CH, tw, lane);
          d = 2.0f * (1.0f - r.best * P.crop_inv_norm[g * 64 + r.arg] * qin);
          if (!(d == d)) d = inf;                    // NaN never enters a top-k
        }
        if (lane == 0) ex[j] = d;
      }
      __syncthreads();
      if (round == 0 && threadIdx.x == 0) {          // k_out exact distances bound the k_out-th smallest of the gallery
        float m = -inf;
        for (int j = 0; j < P.k_out; ++j) m = fmaxf(m, ex[j]);
        bound = m;
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      // selection sort of the evaluated candidates by (distance, index); the first k_out are the answer
      unsigned used = 0;
      float kth = inf;
      for (int r = 0; r < P.k_out; ++r) {
        int bj = -1;
        for (int j = 0; j < P.kc; ++j) {
          if ((used >> j) & 1u) continue;
          if (!(ex[j] < inf)) continue;
          if (bj < 0 || ex[j] < ex[bj] || (ex[j] == ex[bj] && idx[j] < idx[bj])) bj = j;
        }
        if (bj >= 0) used |= 1u << bj;
        P.out_dist[q * P.k_out + r] = bj >= 0 ? ex[bj] : inf;
        P.out_idx[q * P.k_out + r] = bj >= 0 ? idx[bj] : -1;
        kth = bj >= 0 ? ex[bj] : inf;
      }
      // every item outside the list has a key >= the list's last key; its exact distance is >= its key
      if (idx[P.kc - 1] >= 0 && key[P.kc - 1] <= kth) { P.qflag[q] |= 2; atomicAdd(P.n_flagged, 1); }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && cnt_sh && P.rank_count) P.rank_count[q] += cnt_sh;   // the sweep has finished: no atomics needed
}

constexpr int kColItems = 64;   // gallery items per CTA of the column kernel

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

extern "C" int witw_finish_spec_f32(const witw_finish_args* a, witw_stream_t stream) {
  WITW_REQUIRE(a != nullptr, WITW_ERR_INVALID, "witw_finish_spec_f32: null arguments");
  WITW_REQUIRE(a->G >= 0 && a->Q >= 0 && a->CH > 0 && a->list_cap >= 0, WITW_ERR_INVALID, "witw_finish_spec_f32: bad shape");
  if (a->Q == 0 || a->G == 0) return WITW_OK;
  WITW_REQUIRE(a->gal_spec && a->crop_inv_norm && a->qry_spec && a->q_inv_norm && a->qflag && a->n_flagged, WITW_ERR_INVALID,
               "witw_finish_spec_f32: null pointer");
  WITW_REQUIRE(a->list_cap == 0 || (a->list_g && a->list_n), WITW_ERR_INVALID, "witw_finish_spec_f32: deferral list buffers missing");
  WITW_REQUIRE(a->kc >= 0 && a->kc <= kFinMaxCand && (a->kc == 0 || (a->k_out >= 1 && a->k_out <= a->kc && a->cand_key && a->cand_idx && a->out_dist && a->out_idx)),
               WITW_ERR_INVALID, "witw_finish_spec_f32: need 1 <= k_out <= kc <= %d and the candidate / output buffers", kFinMaxCand);
  WITW_REQUIRE((((uintptr_t)a->gal_spec | (uintptr_t)a->qry_spec) & 15) == 0, WITW_ERR_INVALID, "witw_finish_spec_f32: spectra must be 16-byte aligned");
  WITW_REQUIRE(a->Q < (1ll << 31), WITW_ERR_INVALID, "witw_finish_spec_f32: too many queries");
  FinishParams P;
  std::memset(&P, 0, sizeof(P));
  P.gal_spec = reinterpret_cast<const float2*>(a->gal_spec); P.crop_inv_norm = a->crop_inv_norm;
  P.qry_spec = reinterpret_cast<const float2*>(a->qry_spec); P.q_inv_norm = a->q_inv_norm;
  P.G = a->G; P.Q = a->Q; P.CH = a->CH; P.g_offset = a->g_index_offset;
  P.list_g = a->list_cap > 0 ? a->list_g : nullptr; P.list_n = a->list_n; P.list_cap = a->list_cap;
  P.d_true = a->d_true; P.rank_count = a->rank_count; P.dist = a->dist; P.ori = a->ori;
  P.cand_key = a->cand_key; P.cand_idx = a->cand_idx; P.kc = a->kc; P.k_out = a->k_out; P.out_dist = a->out_dist; P.out_idx = a->out_idx;
  P.qflag = a->qflag; P.n_flagged = a->n_flagged;
  const size_t smem = (size_t)a->CH * 32 * sizeof(float2);
  WITW_REQUIRE(smem <= 200 * 1024, WITW_ERR_UNSUPPORTED, "witw_finish_spec_f32: a spectrum of %d rows does not fit shared memory", a->CH);
  WITW_CUDA(cudaFuncSetAttribute(finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  finish_kernel<<<(unsigned)a->Q, kFinThreads, smem, as_stream(stream)>>>(P);
  WITW_LAUNCH_CHECK();
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
