// K2/K3 -- orientation-searched distance on the 5th-generation tensor cores (tcgen05).
// Replaces the correlation -> crop_overhead -> l2_distance chain of model/cvig_fov.py:297-363
// for whole-gallery sweeps, fused with the rank counting of model/cvig_fov.py:552.
//
// GEMM view:  D[q, (g,s)] = sum_{ch,k} su[q,ch,k] * ov[g,ch,(s+k) % 64]
//   M = queries (plain K-major bf16 tile, TMA, 128B swizzle)
//   N = (gallery item, azimuth shift): for one item the B operand is the 64 x K Hankel matrix
//       of its feature rows.  It is never materialised: gallery_prep stores, per pair of items
//       and feature row, "blocks" of 8 rows x 16 bytes
//            block b, row r  =  v_{r%2}[4b + r/2 .. 4b + r/2 + 7]        (v = row, circular)
//       and a NO-SWIZZLE K-major UMMA descriptor with SBO = 128 B (next 8 N-rows = next block)
//       and LBO = 256 B (next 8 K-columns = two blocks on) makes N-row n = 2s + item read
//       v_item[s + k]: overlapping addresses give all 64 shifts of two items from 30 blocks
//       (3.75 KB) instead of a 16 KB expanded tile.
//   K = (feature row, column) = CH * sw_pad, fp32 accumulation in TMEM.
// The shifts of one (query, item) pair sit in 64 TMEM columns of one lane, so the epilogue's
// argmax over the shift is register-local: tcgen05.ld, running max, then
//   dist = 2 - 2 * corr_max * crop_inv_norm[g, s*] * q_inv_norm[q]
// and, optionally, the rank count against the true-match distance and a per-query top-k.
//
// Operands are fp16, norm-scaled, and the epilogue defers what fp16 cannot settle (sweep_common.cuh).
//
// cta_group::2 (default): a CTA pair computes a 256-query x 4-item tile per accumulator stage
// (UMMA 256x256x16); each CTA stages its own 128 queries and 2 items.  Warp roles per CTA:
// 0 TMA producer (cta_group::2 copies signal the leader's barrier), 1 MMA issuer (leader), 2 TMEM allocator,
// 4-7 epilogue.  Two accumulator stages (2 x 256 TMEM columns) overlap epilogue and MMA.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "sweep_common.cuh"
#include "tc_common.cuh"

namespace witw {

void* get_encode_tiled();  // polar.cu

constexpr int kTcThreads = 256;
constexpr int kTcMaxStages = 8;
constexpr int kABytes = 128 * 128;  // 128 queries x 64 bf16
constexpr int kW = 64;              // azimuth columns of the gallery feature map

struct TcGeom {
  int sw_pad;   // 16, 32 or 64: query columns padded with zeros
  int nkap;     // K=16 steps per feature row  = sw_pad / 16
  int cpb;      // feature rows per 64-wide K block = 64 / sw_pad
  int bpc;      // 128-byte blocks stored per (item pair, feature row)
  int sbo, lbo; // B descriptor strides in bytes
  int kstep;    // bytes between consecutive K=16 steps of one feature row
  int b_bytes;  // B bytes per K block per CTA
  int kblocks;  // CH * sw_pad / 64
};

// Layout / CTA-mode experiments exist only in builds with -DWITW_DEBUG_HOOKS (tools/ probes); the shipped library never
// reads the environment: Hankel operand layout, cta_group::2.
#ifdef WITW_DEBUG_HOOKS
static bool full_b_layout() {
  static int v = -1;
  if (v < 0) { const char* e = std::getenv("WITW_TC_FULL_B"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}
static int cta_group_mode() {
  static int v = -1;
  if (v < 0) { const char* e = std::getenv("WITW_TC_CG"); v = (e && e[0] == '1') ? 1 : 2; }
  return v;
}
#else
static bool full_b_layout() { return false; }
static int cta_group_mode() { return 2; }
#endif

static bool make_geom(int CH, int sw, TcGeom* g) {
  if (sw < 1 || sw > 64 || CH < 1) return false;
  g->sw_pad = sw <= 16 ? 16 : (sw <= 32 ? 32 : 64);
  g->nkap = g->sw_pad / 16;
  g->cpb = 64 / g->sw_pad;
  if (CH % g->cpb != 0) return false;
  if (full_b_layout()) {  // debug layout: every (t,u,kappa) block stored separately
    g->bpc = 32 * g->nkap; g->sbo = 128; g->lbo = 16 * 128; g->kstep = 32 * 128;
  } else {                // Hankel layout: block index t + 2u + 4kappa
    g->bpc = 18 + 4 * (g->nkap - 1); g->sbo = 128; g->lbo = 256; g->kstep = 512;
  }
  g->b_bytes = g->cpb * g->bpc * 128;
  g->kblocks = CH * g->sw_pad / 64;
  return true;
}

// ------------------------------------------------------------------------------------------
// operand preparation
// ------------------------------------------------------------------------------------------
constexpr int kPrepRows = 16;  // feature rows of one item pair per CTA

__global__ void __launch_bounds__(256)
gallery_blocks_kernel(const float* __restrict__ ov, int64_t G, int CH, int bpc, int full, const float4* __restrict__ gal_aux,
                      uint4* __restrict__ out) {
  // CTA = (item pair, 16 feature rows).  Phase 1: the 2 x 16 rows are read once (coalesced float4), scaled by the item's
  // kappa / ||ov_g|| (gal_aux[g].w, launch_item_stats), rounded to fp16 once and laid out twice over in shared memory
  // (circular wrap).  Phase 2: every 16-byte row of every block is eight consecutive fp16 of one of those rows; the
  // stores are one contiguous run of the operand.
  __shared__ __align__(16) unsigned short e[2][kPrepRows][136];  // [item][row][0..127 = row twice over]
  const int64_t pair = blockIdx.y;
  const int ch0 = blockIdx.x * kPrepRows;
  const int n_rows = min(kPrepRows, CH - ch0);
  const int tid = threadIdx.x;
  for (int i = tid; i < 2 * kPrepRows * 16; i += 256) {  // one float4 each
    const int item_l = i / (kPrepRows * 16), rem = i - item_l * (kPrepRows * 16);
    const int row = rem >> 4, c4 = rem & 15;
    const int64_t item = 2 * pair + item_l;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (item < G && row < n_rows) {
      v = __ldg(reinterpret_cast<const float4*>(ov + (item * CH + ch0 + row) * kW) + c4);
      const float sc = __ldg(reinterpret_cast<const float*>(gal_aux + item) + 3);
      v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
    }
    const __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
    const uint2 w = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
    *reinterpret_cast<uint2*>(&e[item_l][row][4 * c4]) = w;
    *reinterpret_cast<uint2*>(&e[item_l][row][64 + 4 * c4]) = w;
  }
  __syncthreads();
  const int per_row = bpc * 8;
  uint4* dst = out + ((pair * CH + ch0) * (int64_t)bpc) * 8;
  for (int o = tid; o < n_rows * per_row; o += 256) {
    const int row = o / per_row, rem = o - row * per_row;
    const int b = rem >> 3, r = rem & 7;
    int vb = b;  // virtual Hankel block index t + 2u + 4kappa
    if (full) { const int t = b & 15, u = (b >> 4) & 1, kap = b >> 5; vb = t + 2 * u + 4 * kap; }
    const unsigned short* src = &e[r & 1][row][4 * vb + (r >> 1)];
    uint4 w;
    w.x = (uint32_t)src[0] | ((uint32_t)src[1] << 16);
    w.y = (uint32_t)src[2] | ((uint32_t)src[3] << 16);
    w.z = (uint32_t)src[4] | ((uint32_t)src[5] << 16);
    w.w = (uint32_t)src[6] | ((uint32_t)src[7] << 16);
    dst[o] = w;
  }
}

// Per-item tables (both sweeps; CTA = one item, thread = one azimuth column):
//   crop_inv_norm[g,s] = 1 / ||crop(ov_g, s)||                       -- the fp32 finish (cvig_fov.py:351)
//   gal_scale[g,s]     = ||ov_g|| / (||crop(ov_g, s)|| unit)          -- the sweeps (sweep_common.cuh)
//   gal_aux[g]         = (max_s gal_scale, max - min, rounding scale, kappa / ||ov_g||)
// The rounding scale is written here for the dense operand (dense_sigma > 0: 4-norm of the normalised features); the
// spectral prep kernel writes its own.  Items past G get zeros.  A zero-norm item: operand scale 0, gal_scale NaN.
__global__ void __launch_bounds__(64)
item_stats_kernel(const float* __restrict__ ov, int64_t G, int CH, int sw, float kappa, float unit, float dense_sigma,
                  float* __restrict__ gal_scale, float4* __restrict__ gal_aux, float* __restrict__ crop_inv_norm) {
  __shared__ float col_e[kW];
  __shared__ float red[3][2];
  const int64_t g = blockIdx.x;
  const int j = threadIdx.x;
  if (g >= G) {
    gal_scale[g * kW + j] = 0.f;
    if (crop_inv_norm) crop_inv_norm[g * kW + j] = 0.f;
    if (j == 0) gal_aux[g] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  float e = 0.f, f4 = 0.f;
  for (int ch = 0; ch < CH; ++ch) {
    const float v = ov[(g * CH + ch) * kW + j];
    const float v2 = v * v;
    e += v2;
    f4 = fmaf(v2, v2, f4);
  }
  col_e[j] = e;
  float te = e, t4 = f4;
  for (int m = 16; m > 0; m >>= 1) { te += __shfl_xor_sync(0xffffffffu, te, m); t4 += __shfl_xor_sync(0xffffffffu, t4, m); }
  if ((j & 31) == 0) { red[0][j >> 5] = te; red[1][j >> 5] = t4; }
  __syncthreads();
  const float norm = sqrtf(red[0][0] + red[0][1]);
  float c = 0.f;
  for (int k = 0; k < sw; ++k) c += col_e[(j + k) & 63];
  const float cin = 1.0f / sqrtf(c);
  const float scl = norm * cin / unit;
  gal_scale[g * kW + j] = scl;
  if (crop_inv_norm) crop_inv_norm[g * kW + j] = cin;
  float hi = scl, lo = scl;
  for (int m = 16; m > 0; m >>= 1) { hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, m)); lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, m)); }
  __syncthreads();
  if ((j & 31) == 0) { red[0][j >> 5] = hi; red[2][j >> 5] = lo; }
  __syncthreads();
  if (j == 0) {
    hi = fmaxf(red[0][0], red[0][1]);
    lo = fminf(red[2][0], red[2][1]);
    const float quart = sqrtf(sqrtf(red[1][0] + red[1][1]));
    const bool live = norm > 0.f;
    gal_aux[g] = make_float4(hi, hi - lo, live ? dense_sigma * unit * quart / norm : 0.f, live ? kappa / norm : 0.f);
  }
}

// CTA = one query: K-major fp16 rows of the norm-scaled features, zero-padded to sw_pad columns; q_inv_norm[q] for the fp32
// finish and qry_aux[q] = (1 or NaN, 4-norm of the normalised features) for the sweep.
__global__ void __launch_bounds__(128)
query_prep_kernel(const float* __restrict__ su, int64_t Q, int CH, int sw, int sw_pad, __half* __restrict__ out,
                  float2* __restrict__ qry_aux, float* __restrict__ q_inv_norm) {
  __shared__ float red[2][4];
  const int64_t q = blockIdx.x;
  const float* src = su + q * CH * sw;
  __half* dst = out + q * CH * sw_pad;
  float e = 0.f, f4 = 0.f;
  for (int i = threadIdx.x; i < CH * sw; i += blockDim.x) {
    const float v2 = src[i] * src[i];
    e += v2;
    f4 = fmaf(v2, v2, f4);
  }
  for (int m = 16; m > 0; m >>= 1) { e += __shfl_xor_sync(0xffffffffu, e, m); f4 += __shfl_xor_sync(0xffffffffu, f4, m); }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = e; red[1][threadIdx.x >> 5] = f4; }
  __syncthreads();
  const float norm = sqrtf(red[0][0] + red[0][1] + red[0][2] + red[0][3]);
  const float scale = norm > 0.f ? kDenseKappa / norm : 0.f;
  for (int i = threadIdx.x; i < CH * sw_pad; i += blockDim.x) {
    const int ch = i / sw_pad, k = i - ch * sw_pad;
    dst[i] = __float2half_rn(k < sw ? src[ch * sw + k] * scale : 0.f);
  }
  if (threadIdx.x == 0) {
    q_inv_norm[q] = 1.0f / norm;
    const float quart = sqrtf(sqrtf(red[1][0] + red[1][1] + red[1][2] + red[1][3]));
    qry_aux[q] = make_float2(norm > 0.f ? 1.0f : __int_as_float(0x7fc00000), norm > 0.f ? quart / norm : 0.f);
  }
}

constexpr int kTopkMax = kSweepTopk;

struct TcParams {
  SweepOut out;
  int n_qtiles, n_chunks, groups_per_chunk, n_groups;
  int kblocks, cpb, bpc, nkap;
  int sbo, lbo, kstep, b_bytes;
  int pair_rows;                 // CH * bpc: 128-byte blocks per item pair
  int stage_rows;                // cpb * bpc: blocks per K block per CTA
  int n_stages;                  // depth of the operand ring (<= kTcMaxStages)
};

// AMB: the epilogue also tracks the runner-up accumulator of every pair (the ambiguity test of sweep_common.cuh: cropped
// queries, orientation output).  Two more instructions per accumulator; compiled out of the plain variant, whose epilogue
// must drain 256 columns per accumulator stage while the MMAs of the next stage run.
template <int CG, bool AMB>
__global__ void __launch_bounds__(kTcThreads, 1)
match_tc_kernel(const __grid_constant__ CUtensorMap q_map, const __grid_constant__ CUtensorMap g_map, const TcParams P) {
  constexpr int UM = 128 * CG;           // UMMA M
  constexpr int UN = 128 * CG;           // UMMA N  (= accumulator columns per stage)
  constexpr int IG = 2 * CG;             // gallery items per group
  constexpr uint32_t kTmemCols = 2 * UN;
  // kind::f16: D fp32 (bit 4), A and B fp16 (format fields 0), both K-major
  constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(UN >> 3) << 17) | ((uint32_t)(UM >> 4) << 24);

  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t b_stride = (uint32_t)((P.b_bytes + 127) & ~127);
  unsigned char* a_base = smem;                                 // kTcStages * 16 KB, 1024-aligned
  const uint32_t kTcStages = (uint32_t)P.n_stages;
  unsigned char* b_base = smem + kTcStages * kABytes;           // n_stages * b_stride
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_base + kTcStages * b_stride);
  uint64_t* full = bars;                         // [n_stages]
  uint64_t* empty = bars + kTcMaxStages;         // [n_stages]
  uint64_t* tmem_full = bars + 2 * kTcMaxStages; // [2]
  uint64_t* tmem_empty = tmem_full + 2;       // [2]          (leader only)
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t cta_rank = 0;
  if constexpr (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
  const bool leader = cta_rank == 0;
  const int unit = blockIdx.x / CG, n_units = gridDim.x / CG;
  const int n_items = P.n_chunks * P.n_qtiles;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&q_map) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&g_map) : "memory");
  }
  if (warp == 1 && lane == 0) {
    if (s2u(smem) & 1023u) __trap();  // SWIZZLE_128B operand tiles need a 1024-byte aligned base
    for (uint32_t s = 0; s < kTcStages; ++s) {
      bar_init(s2u(&full[s]), 1);
      bar_init(s2u(&empty[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      bar_init(s2u(&tmem_full[a]), 1);
      bar_init(s2u(&tmem_empty[a]), 4 * CG);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    if constexpr (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s2u(tmem_holder)), "r"(kTmemCols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s2u(tmem_holder)), "r"(kTmemCols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if constexpr (CG == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  __syncthreads();   // (with CG == 2 the cluster barrier already orders this; compute-sanitizer's racecheck only knows this one)
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ===================== TMA producer (whole warp runs the loop, one elected lane issues) =====================
    // Every stage's bytes -- this CTA's 128 queries and 2 items, and the peer's -- are accounted on the
    // LEADER's full barrier: the leader arms it with the pair's total, the peer only issues its copies.
    const uint32_t stage_tx = (uint32_t)CG * (uint32_t)(kABytes + P.b_bytes);
    const uint32_t full0 = CG == 2 ? map_to_cta(s2u(&full[0]), 0) : s2u(&full[0]);
    uint32_t s = 0, ph = 0;
    for (int item = unit; item < n_items; item += n_units) {
      const int qt = item / P.n_chunks, chunk = item - qt * P.n_chunks;
      const int q_row0 = qt * UM + (int)cta_rank * 128;
      const int grp0 = chunk * P.groups_per_chunk;
      const int grp1 = min(grp0 + P.groups_per_chunk, P.n_groups);
      for (int grp = grp0; grp < grp1; ++grp) {
        int g_row = (grp * CG + (int)cta_rank) * P.pair_rows;
        for (int kb = 0; kb < P.kblocks; ++kb, g_row += P.stage_rows) {
          bar_wait(s2u(&empty[s]), ph ^ 1);
          if (elect_one()) {
            const uint32_t fb = full0 + s * 8;
            if (leader) bar_expect_tx(s2u(&full[s]), stage_tx);
            tma_2d<CG>(s2u(a_base) + s * kABytes, &q_map, fb, kb * 64, q_row0);
            tma_2d<CG>(s2u(b_base) + s * b_stride, &g_map, fb, 0, g_row);
          }
          __syncwarp();
          if (++s == kTcStages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ===================== MMA issuer (leader CTA; warp-uniform loop, one elected lane issues) =====================
      // Per K block: one barrier wait, four tcgen05.mma (K = 16 each) whose descriptors are the stage-0
      // descriptors plus precomputed 16-byte-unit offsets, one commit.  Everything else is hoisted.
      const uint64_t a_desc0 = make_desc(s2u(a_base), 16, 1024, 2 /*SWIZZLE_128B*/);
      const uint64_t b_desc0 = make_desc(s2u(b_base), P.lbo, P.sbo, 0 /*no swizzle*/);
      uint32_t b_off[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c_in = j / P.nkap, kap = j - c_in * P.nkap;
        b_off[j] = (uint32_t)(c_in * P.bpc * 128 + kap * P.kstep) >> 4;
      }
      const uint32_t a_step = kABytes >> 4, b_step = b_stride >> 4;
      uint32_t s = 0, ph = 0, acc_it = 0;
      for (int item = unit; item < n_items; item += n_units) {
        const int chunk = item % P.n_chunks;
        const int grp0 = chunk * P.groups_per_chunk;
        const int grp1 = min(grp0 + P.groups_per_chunk, P.n_groups);
        for (int grp = grp0; grp < grp1; ++grp, ++acc_it) {
          const uint32_t acc = acc_it & 1, acc_ph = (acc_it >> 1) & 1;
          bar_wait(s2u(&tmem_empty[acc]), acc_ph ^ 1);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + acc * UN;
          for (int kb = 0; kb < P.kblocks; ++kb) {
            bar_wait(s2u(&full[s]), ph);
            tc_fence_after();
            if (elect_one()) {
              const uint64_t da = a_desc0 + (uint64_t)(s * a_step);
              const uint64_t db = b_desc0 + (uint64_t)(s * b_step);
#pragma unroll
              for (int j = 0; j < 4; ++j)  // 32 bytes (2 units) along K per step inside the 128B-swizzled query rows
                umma_bf16<CG>(tmem_d, da + (uint64_t)(2 * j), db + (uint64_t)b_off[j], kIdesc, (kb | j) != 0 ? 1u : 0u);
              umma_commit<CG>(s2u(&empty[s]));  // frees the stage in both CTAs when these MMAs retire
              if (kb == P.kblocks - 1) umma_commit<CG>(s2u(&tmem_full[acc]));
            }
            __syncwarp();
            if (++s == kTcStages) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: 4 warps = 128 TMEM lanes = 128 queries =====================
    const SweepOut& S = P.out;
    const int wq = warp & 3;
    const int row = wq * 32 + lane;
    const uint32_t lane_field = (uint32_t)(wq * 32) << 16;
    uint32_t acc_it = 0;
    for (int item = unit; item < n_items; item += n_units) {
      const int qt = item / P.n_chunks, chunk = item - qt * P.n_chunks;
      const int64_t q = (int64_t)qt * UM + cta_rank * 128 + row;
      SweepQuery qc;
      qc.load(S, q);
      int cnt = 0;
      float td[kTopkMax];
      int32_t ti[kTopkMax];
#pragma unroll
      for (int j = 0; j < kTopkMax; ++j) { td[j] = __int_as_float(0x7f800000); ti[j] = -1; }
      const int grp0 = chunk * P.groups_per_chunk;
      const int grp1 = min(grp0 + P.groups_per_chunk, P.n_groups);
      for (int grp = grp0; grp < grp1; ++grp, ++acc_it) {
        const uint32_t acc = acc_it & 1, acc_ph = (acc_it >> 1) & 1;
        bar_wait(s2u(&tmem_full[acc]), acc_ph);
        tc_fence_after();
        float best[IG], second[IG];   // second: the largest accumulator of any other shift (the ambiguity test needs it)
        int arg[IG];
#pragma unroll
        for (int i = 0; i < IG; ++i) { best[i] = -__int_as_float(0x7f800000); second[i] = best[i]; arg[i] = 0; }
#pragma unroll
        for (int c = 0; c < UN / 32; ++c) {  // 32 columns = 16 shifts x 2 items
          uint32_t v[32];
          tmem_ld32(tmem_base + lane_field + acc * UN + c * 32, v);
          tmem_ld_wait();
          const int h = c >> 2;              // which CTA's item pair these columns belong to
          const int s0 = (c & 3) * 16;
#pragma unroll
          for (int ds = 0; ds < 16; ++ds) {
#pragma unroll
            for (int gi = 0; gi < 2; ++gi) {
              const float x = __uint_as_float(v[2 * ds + gi]);
              const int i = 2 * h + gi;
              if constexpr (AMB) second[i] = fmaxf(second[i], fminf(x, best[i]));
              if (x > best[i]) { best[i] = x; arg[i] = s0 + ds; }  // strict '>' in ascending shift order: first maximum
            }
          }
        }
        // accumulator stage is drained into registers: hand it back to the MMA issuer
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CG == 1 || leader) bar_arrive_local(s2u(&tmem_empty[acc]));
          else bar_arrive_remote(s2u(&tmem_empty[acc]), 0);   // (CTA-scope release: no GPU-wide fence per group)
        }
        // pin the tail below the release of the accumulator stage (see match_spec.cu: the compiler otherwise delays the arrive)
#pragma unroll
        for (int i = 0; i < IG; ++i) asm volatile("" : "+f"(best[i]), "+f"(second[i]), "+r"(arg[i]));
#pragma unroll
        for (int i = 0; i < IG; ++i) {
          const int64_t g = (int64_t)grp * IG + i;
          if (g < S.G && qc.ok) {
            const float e = sweep_err(S, __ldg(reinterpret_cast<const float*>(S.gal_aux + g) + 2), qc, best[i]);
            const bool amb = AMB && second[i] >= best[i] - 2.0f * e;
            sweep_pair(S, qc, g, q, best[i], arg[i], amb, __ldg(S.gal_scale + g * kW + arg[i]), e, cnt, td, ti);
          }
        }
      }
      if (qc.ok) {
        if (S.rank_count && cnt) atomicAdd(S.rank_count + q, cnt);
        if (S.topk > 0) {
          float* od = S.topk_key + ((int64_t)chunk * S.Q + q) * S.topk;
          int32_t* oi = S.topk_idx + ((int64_t)chunk * S.Q + q) * S.topk;
#pragma unroll
          for (int j = 0; j < kTopkMax; ++j)
            if (j < S.topk) { od[j] = td[j]; oi[j] = ti[j]; }
        }
      }
    }
  }

  // ===================== teardown =====================
  tc_fence_before();
  if constexpr (CG == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  } else {
    __syncthreads();
  }
  if (warp == 2) {
    tc_fence_after();
    if constexpr (CG == 1)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// Per-item tables for G_pad >= G items (zeros past G); also used by the spectral sweep's prep (match_spec.cu).
int launch_item_stats(const float* ov, int64_t G, int64_t G_pad, int CH, int sw, float kappa, float unit, float dense_sigma,
                      float* gal_scale, float* gal_aux, float* crop_inv_norm, witw_stream_t stream) {
  item_stats_kernel<<<(unsigned)G_pad, 64, 0, as_stream(stream)>>>(ov, G, CH, sw, kappa, unit, dense_sigma, gal_scale,
                                                                   reinterpret_cast<float4*>(gal_aux), crop_inv_norm);
  WITW_LAUNCH_CHECK();
  return WITW_OK;
}

// Work split shared by the kernel launch and witw_match_tc_topk_slots().
struct TcSchedule { int cg, n_units, n_qtiles, n_groups, groups_per_chunk, n_chunks; };

static TcSchedule make_schedule(int64_t G, int64_t Q) {
  TcSchedule s;
  s.cg = cta_group_mode();
  s.n_units = std::max(1, sm_count() / s.cg);
  const int64_t tq = 128 * s.cg, ig = 2 * s.cg;
  s.n_qtiles = (int)ceil_div<int64_t>(std::max<int64_t>(Q, 1), tq);
  s.n_groups = (int)ceil_div<int64_t>(std::max<int64_t>(G, 1), ig);
  // about 32 work items per unit for balance, chunks of at least 4 groups, at most 64 chunks
  // (the top-k merge takes up to 64 candidate lists)
  int64_t want_chunks = ceil_div<int64_t>((int64_t)s.n_units * 32, s.n_qtiles);
  want_chunks = std::max<int64_t>(1, std::min<int64_t>(want_chunks, 64));
  want_chunks = std::min<int64_t>(want_chunks, ceil_div<int64_t>(s.n_groups, 4));
  s.groups_per_chunk = (int)ceil_div<int64_t>(s.n_groups, std::max<int64_t>(want_chunks, 1));
  s.n_chunks = (int)ceil_div<int64_t>(s.n_groups, s.groups_per_chunk);
  return s;
}

}  // namespace witw

using namespace witw;

extern "C" size_t witw_gallery_operand_bytes(int64_t G, int CH, int sw) {
  TcGeom g;
  if (G < 0 || !make_geom(CH, sw, &g)) { set_error(WITW_ERR_UNSUPPORTED, "witw_gallery_operand_bytes: unsupported CH=%d sw=%d", CH, sw); return 0; }
  const int64_t pairs = ceil_div<int64_t>(G, 4) * 2;
  return (size_t)std::max<int64_t>(pairs, 2) * CH * g.bpc * 128;
}

extern "C" size_t witw_query_operand_bytes(int64_t Q, int CH, int sw) {
  TcGeom g;
  if (Q < 0 || !make_geom(CH, sw, &g)) { set_error(WITW_ERR_UNSUPPORTED, "witw_query_operand_bytes: unsupported CH=%d sw=%d", CH, sw); return 0; }
  return (size_t)std::max<int64_t>(Q, 1) * CH * g.sw_pad * 2;
}

extern "C" int witw_gallery_prep(const float* ov, int64_t G, int CH, int W, int sw, void* gal_op, float* gal_scale, float* gal_aux,
                                 float* crop_inv_norm, witw_stream_t stream) {
  TcGeom g;
  WITW_REQUIRE(W == kW, WITW_ERR_UNSUPPORTED, "witw_gallery_prep: the tensor-core path needs W == 64 (got %d)", W);
  WITW_REQUIRE(G >= 0 && make_geom(CH, sw, &g), WITW_ERR_UNSUPPORTED, "witw_gallery_prep: unsupported CH=%d sw=%d", CH, sw);
  if (G == 0) return WITW_OK;
  WITW_REQUIRE(ov && gal_op && gal_scale && gal_aux, WITW_ERR_INVALID, "witw_gallery_prep: null pointer");
  WITW_REQUIRE(((uintptr_t)gal_op & 127) == 0 && ((uintptr_t)gal_aux & 15) == 0, WITW_ERR_INVALID,
               "witw_gallery_prep: operand buffer must be 128-byte, gal_aux 16-byte aligned");
  const int64_t g4 = ceil_div<int64_t>(G, 4) * 4;
  WITW_REQUIRE(((uintptr_t)ov & 15) == 0, WITW_ERR_INVALID, "witw_gallery_prep: features must be 16-byte aligned");
  int rc = launch_item_stats(ov, G, g4, CH, sw, kDenseKappa, kDenseUnit, kRoundSigma, gal_scale, gal_aux, crop_inv_norm, stream);
  if (rc != WITW_OK) return rc;
  for (int64_t p0 = 0; p0 < g4 / 2; p0 += 65535) {  // gridDim.y limit
    const int64_t np = std::min<int64_t>(65535, g4 / 2 - p0);
    gallery_blocks_kernel<<<dim3((unsigned)ceil_div(CH, kPrepRows), (unsigned)np), 256, 0, as_stream(stream)>>>(
        ov + p0 * 2 * CH * kW, G - 2 * p0, CH, g.bpc, full_b_layout() ? 1 : 0, reinterpret_cast<const float4*>(gal_aux) + 2 * p0,
        reinterpret_cast<uint4*>(gal_op) + p0 * CH * g.bpc * 8);
    WITW_LAUNCH_CHECK();
  }
  return WITW_OK;
}

extern "C" int witw_query_prep(const float* su, int64_t Q, int CH, int sw, void* qry_op, float* qry_aux, float* q_inv_norm,
                               witw_stream_t stream) {
  TcGeom g;
  WITW_REQUIRE(Q >= 0 && make_geom(CH, sw, &g), WITW_ERR_UNSUPPORTED, "witw_query_prep: unsupported CH=%d sw=%d", CH, sw);
  if (Q == 0) return WITW_OK;
  WITW_REQUIRE(su && qry_op && qry_aux && q_inv_norm, WITW_ERR_INVALID, "witw_query_prep: null pointer");
  WITW_REQUIRE(Q < (1ll << 31) && ((uintptr_t)qry_aux & 7) == 0, WITW_ERR_INVALID, "witw_query_prep: too many queries or qry_aux not 8-byte aligned");
  query_prep_kernel<<<(unsigned)Q, 128, 0, as_stream(stream)>>>(su, Q, CH, sw, g.sw_pad, reinterpret_cast<__half*>(qry_op),
                                                              reinterpret_cast<float2*>(qry_aux), q_inv_norm);
  WITW_LAUNCH_CHECK();
  return WITW_OK;
}

extern "C" int witw_match_tc_topk_slots(int64_t G, int64_t Q) { return make_schedule(G, Q).n_chunks; }

extern "C" int witw_match_tc(const witw_sweep_args* a, witw_stream_t stream) {
  WITW_REQUIRE(a != nullptr, WITW_ERR_INVALID, "witw_match_tc: null arguments");
  const int64_t G = a->G, Q = a->Q;
  const int CH = a->CH, sw = a->sw;
  TcGeom geo;
  WITW_REQUIRE(G >= 0 && Q >= 0 && make_geom(CH, sw, &geo), WITW_ERR_UNSUPPORTED, "witw_match_tc: unsupported CH=%d sw=%d", CH, sw);
  if (G == 0 || Q == 0) return WITW_OK;
  SweepOut out;
  int rc = fill_sweep_out("witw_match_tc", a, kDenseUnit, kDenseAccRel, &out);
  if (rc != WITW_OK) return rc;
  const void* qry_op = a->qry_op;
  const void* gal_op = a->gal_op;
  WITW_REQUIRE(((uintptr_t)qry_op & 15) == 0 && ((uintptr_t)gal_op & 15) == 0, WITW_ERR_INVALID, "witw_match_tc: operands must be 16-byte aligned");
  rc = witw_device_check();
  if (rc != WITW_OK) return rc;

  const TcSchedule sch = make_schedule(G, Q);
  typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  auto encode = reinterpret_cast<EncodeTiledFn>(get_encode_tiled());
  WITW_REQUIRE(encode != nullptr, WITW_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const int64_t ktot = (int64_t)CH * geo.sw_pad;
  CUtensorMap qmap;
  const cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)Q};
  const cuuint64_t strides[1] = {(cuuint64_t)ktot * 2};
  const cuuint32_t box[2] = {64, 128};
  const cuuint32_t estr[2] = {1, 1};
  CUresult cr = encode(&qmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(qry_op), dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  WITW_REQUIRE(cr == CUDA_SUCCESS, WITW_ERR_CUDA, "cuTensorMapEncodeTiled(query operand) failed with CUresult %d", (int)cr);

  // gallery operand as a [total_blocks][64] fp16 matrix of 128-byte blocks; one box = the blocks of one K block
  const int64_t pairs = ceil_div<int64_t>(G, 4) * 2;
  const int64_t total_rows = pairs * CH * geo.bpc;
  WITW_REQUIRE(total_rows < (1ll << 31), WITW_ERR_UNSUPPORTED, "witw_match_tc: gallery of %lld items is too large for one sweep", (long long)G);
  WITW_REQUIRE(geo.cpb * geo.bpc <= 256, WITW_ERR_UNSUPPORTED, "witw_match_tc: operand stage of %d blocks exceeds a TMA box", geo.cpb * geo.bpc);
  CUtensorMap gmap;
  const cuuint64_t gdims[2] = {64, (cuuint64_t)total_rows};
  const cuuint64_t gstrides[1] = {128};
  const cuuint32_t gbox[2] = {64, (cuuint32_t)(geo.cpb * geo.bpc)};
  cr = encode(&gmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(gal_op), gdims, gstrides, gbox, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  WITW_REQUIRE(cr == CUDA_SUCCESS, WITW_ERR_CUDA, "cuTensorMapEncodeTiled(gallery operand) failed with CUresult %d", (int)cr);

  TcParams P;
  std::memset(&P, 0, sizeof(P));
  P.out = out;
  P.n_qtiles = sch.n_qtiles; P.n_chunks = sch.n_chunks; P.groups_per_chunk = sch.groups_per_chunk; P.n_groups = sch.n_groups;
  P.kblocks = geo.kblocks; P.cpb = geo.cpb; P.bpc = geo.bpc; P.nkap = geo.nkap;
  P.sbo = geo.sbo; P.lbo = geo.lbo; P.kstep = geo.kstep; P.b_bytes = geo.b_bytes;
  P.pair_rows = CH * geo.bpc;
  P.stage_rows = geo.cpb * geo.bpc;

  const uint32_t b_stride = (uint32_t)((geo.b_bytes + 127) & ~127);
  const size_t fixed = (2 * kTcMaxStages + 4) * 8 + 16;
  int n_stages = (int)std::min<size_t>(kTcMaxStages, (227 * 1024 - fixed) / (kABytes + b_stride));
  WITW_REQUIRE(n_stages >= 2, WITW_ERR_UNSUPPORTED, "witw_match_tc: operand stage of %u bytes does not fit shared memory twice", kABytes + b_stride);
  P.n_stages = n_stages;
  const size_t smem = (size_t)n_stages * (kABytes + b_stride) + fixed;
  const int n_items = sch.n_chunks * sch.n_qtiles;
  const int units = std::min(sch.n_units, n_items);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(units * sch.cg));
  cfg.blockDim = dim3(kTcThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = as_stream(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)sch.cg;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  auto kernel = sch.cg == 2 ? (out.need_amb ? match_tc_kernel<2, true> : match_tc_kernel<2, false>)
                            : (out.need_amb ? match_tc_kernel<1, true> : match_tc_kernel<1, false>);
  WITW_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  WITW_CUDA(cudaLaunchKernelEx(&cfg, kernel, qmap, gmap, P));
  WITW_LAUNCH_CHECK();
  return WITW_OK;
}
