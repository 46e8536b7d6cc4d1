// K2/K3, azimuth-frequency form -- the orientation-searched distance sweep through the correlation theorem,
// with the per-frequency products on the 5th-generation tensor cores (tcgen05) and the inverse transform,
// argmax, distance, rank count and top-k fused in the epilogue.  Same results contract as match_tc.cu
// (correlation -> crop_overhead -> l2_distance of model/cvig_fov.py:297-363 + the rank rule of :552), about
// 15x fewer tensor FLOPs at 360 degrees:
//
//   corr[g,q,s] = sum_{r,k} ov[g,r,(s+k)%64] su[q,r,k]  =  irfft_s( P[g,q,f] ),   P_f = sum_r O[g,r,f] conj(S[q,r,f])
//
// with O, S the 64-point spectra of the feature rows (r = (channel, height) row, 64 of them; query rows
// zero-padded).  33 frequencies x 64 rows of complex MACs = 16.9 kFLOP per pair instead of 524 kFLOP.
//
// GEMM view, one per packed frequency slot f (slot 0 carries the two real bins 0 and 32):
//   D_f[q, (item, c)] = sum_k A_f[q,k] B_f[(item,c),k],   k = (half, r), K = 128
//   A_f[q]       = [ Re S_f(r) | Im S_f(r) ] / 64                      (slot 0: [ S_0 | S_32 ] / 64)
//   B_f[item,0]  = [ Re O_f(r) |  Im O_f(r) ]   -> Re P_f              (slot 0: [ O_0 | 0 ]   -> P_0)
//   B_f[item,1]  = [ Im O_f(r) | -Re O_f(r) ]   -> Im P_f              (slot 0: [ 0 | O_32 ]  -> P_32)
// Per slot a CTA's accumulator is 128 TMEM lanes x 16 columns, column 4 (item / 2) + 2 c + item % 2 of an 8-item group, and
// the 32 slots fill the 512 TMEM columns: the 64 accumulators of one (query, item) pair are 64 columns of one lane, so the
// inverse real FFT (ifft64_gen.cuh, 487 fp32 operations) and the maximum over the shift are register-local; the epilogue
// transforms two items at once in the two halves of packed f32x2 operations.  TMEM holds 1 024 pairs: a CTA's tile.
//
// Default tiling: CTA pairs (tcgen05 cta_group::2, M = 128, N = 32).  A CTA stages 64 queries (its half of M) and one 8-item
// group (its half of N; the pair's MMA reads both CTAs' halves) per slot, 20 KB, and holds 64 queries x 16 items in TMEM --
// 640 B of operands per pair instead of the 1 152 B of one CTA per 128-query x 8-item tile (kept as variant 1: same operand
// layouts, bit-identical results).  Per CTA (persistent, 16 warps): warps 0-1 TMA producers, warps 2-6 MMA issuers in the
// leader CTA (warp k owns every fifth slot; a single issuing warp's barrier round trip, ~430 cycles per stage, paced the
// first version), warps 8-15 epilogue (two warps per TMEM lane quarter, four items each); 10-stage operand ring, a stage is
// one slot.  What bounds it (measured with tools/ring_roof.py): TMA + MMAs alone 3.1 ms per 10k x 10k (the
// L2 -> SM delivery of the operands), the epilogue alone 2.8 ms, the kernel 4.1 - 4.4 ms -- TMEM holds exactly one tile, so
// the first half of a tile's epilogue cannot overlap the next tile's MMAs.
//
// Operands are fp16 spectra of the norm-scaled features (sweep_common.cuh): 8x finer than bf16 at the same cost, and
// with a per-pair bound on what the rounding can do to a result, so that the epilogue knows which decisions it may
// take itself and which it must defer to the fp32 finish (finish.cu).
#include <cuda.h>
#include <cuda_fp16.h>
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "ifft64_gen.cuh"
#include "row_fft.cuh"
#include "sweep_common.cuh"
#include "tc_common.cuh"

namespace witw {

void* get_encode_tiled();                                                                         // polar.cu

constexpr int kSpThreads = 512;
constexpr int kSpStages = 6;           // one CTA per tile: stages of 36 KB, one issuer warp per stage
#ifndef SPEC_NS_PAIR
#define SPEC_NS_PAIR 10
#define SPEC_NP_PAIR 2
#define SPEC_NI_PAIR 5
#endif
constexpr int kSpStagesPair = SPEC_NS_PAIR;   // CTA pairs: stages of 20 KB
constexpr int kSpProducersPair = SPEC_NP_PAIR, kSpIssuersPair = SPEC_NI_PAIR;   // control warps; both counts divide the ring depth
constexpr int kSpABytes = 2 * 128 * 128;  // 2 K halves x 128 queries x 64 fp16 (one CTA per tile)
constexpr int kSpABytesPair = kSpABytes / 2;  // CTA pair: each CTA stages 64 of the tile's 128 queries
constexpr int kSpBBytes = 2 * 16 * 128;   // 2 K halves x 8 items x (Re, Im) x 64 fp16
constexpr int kSpSlots = 32;
constexpr int kSpItems = 8;            // gallery items per tile (= per operand group)
constexpr int kSpCH = 64;              // feature rows (C*H) the operand layout is built for
constexpr int kSpTopkMax = kSweepTopk;
constexpr int kSpMaxChunks = 32;       // two candidate lists per chunk (four with CTA pairs); witw_topk_merge takes up to 64

// Timing experiments (wrong results by design) exist only in builds with -DWITW_DEBUG_HOOKS (tools/ probes): WITW_SPEC_DEBUG
// bit 0 = no inverse FFT, bit 1 = no TMEM loads in the epilogue, bit 2 = no tcgen05.mma (barriers only), bit 3 = no
// query-stage loads, bit 4 = chunk-major work order, bit 5 = no operand ring at all (the epilogue and the per-tile handshake
// alone), bit 7 = no per-tile handshake and no epilogue (the ring alone), bit 8 = only half of every gallery stage is loaded.
// The shipped library never reads the environment.
#ifdef WITW_DEBUG_HOOKS
static int spec_debug() {
  static int v = -1;
  if (v < 0) { const char* e = std::getenv("WITW_SPEC_DEBUG"); v = e ? std::atoi(e) : 0; }
  return v;
}
#define SPEC_DBG(P, bit) (((P).debug & (bit)) != 0)
#else
static int spec_debug() { return 0; }
#define SPEC_DBG(P, bit) false
#endif

// ------------------------------------------------------------------------------------------
// operand preparation
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint2 pack4_f16(float a, float b, float c, float d) {
  const __half2 lo = __floats2half2_rn(a, b), hi = __floats2half2_rn(c, d);
  return make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
}

// sum over the threads of a 256-thread CTA; red: 8 floats of shared memory
__device__ __forceinline__ float block_sum_256(float v, float* red) {
  for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) t += red[w];
  return t;
}

// sum over the packed half spectrum (32 slots x 64 rows in shared memory) of |X|^4; slot 0 holds the two real bins
__device__ __forceinline__ float spectrum_quartic(const float (*sre)[kSpCH + 1], const float (*sim)[kSpCH + 1]) {
  float s4 = 0.f;
  for (int idx = threadIdx.x; idx < kSpSlots * kSpCH; idx += 256) {
    const int slot = idx >> 6, row = idx & 63;
    const float re = sre[slot][row], im = sim[slot][row];
    if (slot == 0) {
      s4 += re * re * re * re + im * im * im * im;
    } else {
      const float m2 = re * re + im * im;
      s4 += m2 * m2;
    }
  }
  return s4;
}

// CTA = one gallery item.  One pass over the item's fp32 features gives everything the sweeps and the fp32 finish need:
// the 64 row spectra (warp-level FFTs), the column energies behind crop_inv_norm[g,s] = 1/||crop(ov_g,s)|| and
// gal_scale[g,s] = ||ov_g|| / (||crop(ov_g,s)|| unit), the item's norm (operand scale kappa / ||ov_g||), its rounding scale
// (sweep_common.cuh), the fp32 spectra of the finish, and the item's 128 operand rows of 128 bytes (slot, K half, Re/Im)
// in its group's tiles.  A warp loads its eight rows before it transforms the first one: with one 256-byte load in flight per
// warp the kernel ran at 2.8 TB/s.  Items past G (the rest of the last group) get zeros.
__global__ void __launch_bounds__(256)
spec_gallery_prep_kernel(const float* __restrict__ ov, int64_t G, int64_t g_first, int sw, __half* __restrict__ out,
                         float* __restrict__ gal_scale, float4* __restrict__ gal_aux, float* __restrict__ crop_inv_norm,
                         float* __restrict__ spec_out) {
  __shared__ float sre[kSpSlots][kSpCH + 1], sim[kSpSlots][kSpCH + 1];  // [slot][feature row]
  __shared__ float col_part[8][64];
  __shared__ float col_e[64];
  __shared__ float red[8];
  __shared__ float stat[4];           // norm, max scale, min scale
  const int64_t g_local = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool live = g_local < G;
  RowFft fft;
  fft.init(lane);
  float2 z[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    z[i] = make_float2(0.f, 0.f);
    if (live) z[i] = __ldg(reinterpret_cast<const float2*>(ov + (g_local * kSpCH + warp + 8 * i) * 64) + lane);
  }
  float e0 = 0.f, e1 = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    e0 = fmaf(z[i].x, z[i].x, e0);
    e1 = fmaf(z[i].y, z[i].y, e1);
    const float2 X = fft.run(z[i], lane);
    sre[lane][warp + 8 * i] = X.x;
    sim[lane][warp + 8 * i] = X.y;
  }
  col_part[warp][2 * lane] = e0;
  col_part[warp][2 * lane + 1] = e1;
  __syncthreads();
  if (threadIdx.x < 64) {             // column energies, the item's norm
    float e = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) e += col_part[w][threadIdx.x];
    col_e[threadIdx.x] = e;
    float t = e;
    for (int m = 16; m > 0; m >>= 1) t += __shfl_xor_sync(0xffffffffu, t, m);
    if (lane == 0) red[warp] = t;
  }
  if (spec_out != nullptr && live) {  // the fp32 spectra the exact finish works on (layout of witw_spectral_rows_f32)
    float2* dst = reinterpret_cast<float2*>(spec_out + g_local * (kSpCH * 64));
    for (int idx = threadIdx.x; idx < kSpCH * 32; idx += 256) dst[idx] = make_float2(sre[idx & 31][idx >> 5], sim[idx & 31][idx >> 5]);
  }
  __syncthreads();
  const float norm = sqrtf(red[0] + red[1]);
  if (threadIdx.x < 64) {             // crop norms of the 64 shifts and the scale table of the sweep
    const int j = threadIdx.x;
    float c = 0.f;
    const int j0 = sw == 64 ? 0 : j;    // a full panorama's crop is the whole item at every shift: one summation order, one value
    for (int k = 0; k < sw; ++k) c += col_e[(j0 + k) & 63];
    const float cin = live ? 1.0f / sqrtf(c) : 0.f;
    const float scl = live ? norm * cin / kSpecUnit : 0.f;
    gal_scale[g_local * 64 + j] = scl;
    if (crop_inv_norm != nullptr) crop_inv_norm[g_local * 64 + j] = cin;
    float hi = scl, lo = scl;
    for (int m = 16; m > 0; m >>= 1) { hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, m)); lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, m)); }
    if (lane == 0) { stat[warp] = hi; stat[2 + warp] = lo; }
  }
  const float scale = (live && norm > 0.f) ? kSpecKappa / norm : 0.f;     // a zero-norm item: a zero operand, NaN scale table
  // rounding scale of the item (sweep_common.cuh): 4-norm of the normalised half spectrum
  const float s4 = block_sum_256(spectrum_quartic(sre, sim), red);        // (its barriers also publish stat[])
  if (threadIdx.x == 0) {
    const float hi = fmaxf(stat[0], stat[1]), lo = fminf(stat[2], stat[3]);
    gal_aux[g_local] = make_float4(hi, hi - lo, kRoundSigma * (2.0f / 64.0f) * kSpecUnit * sqrtf(sqrtf(s4)) * (scale / kSpecKappa), scale);
  }
  // the item's 128 operand rows: thread = (slot, eight feature rows); its 8 spectrum values give 16 bytes of each of the four
  // operand rows (K half x Re/Im output) of the slot, so every value is loaded, scaled and converted once
  const int64_t g = g_first + g_local;
  const int64_t group = g >> 3;
  const int i = (int)(g & 7);
  {
    const int slot = threadIdx.x >> 3, r0 = (threadIdx.x & 7) * 8;
    float re[8], im[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { re[j] = sre[slot][r0 + j] * scale; im[j] = sim[slot][r0 + j] * scale; }
    const int64_t row0 = (group * kSpSlots + slot) * 32 + 4 * (i >> 1) + (i & 1);   // K half 0, output 0; + 2 = output 1; + 16 = K half 1
    uint4 w[4];   // [Re | Im] -> Re P,  [Im | -Re] -> Im P;  slot 0: [O_0 | 0] -> P_0,  [0 | O_32] -> P_32
    const uint2 a0 = pack4_f16(re[0], re[1], re[2], re[3]), a1 = pack4_f16(re[4], re[5], re[6], re[7]);
    const uint2 b0 = pack4_f16(im[0], im[1], im[2], im[3]), b1 = pack4_f16(im[4], im[5], im[6], im[7]);
    const uint2 n0 = pack4_f16(-re[0], -re[1], -re[2], -re[3]), n1 = pack4_f16(-re[4], -re[5], -re[6], -re[7]);
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    w[0] = make_uint4(a0.x, a0.y, a1.x, a1.y);                                      // half 0, output 0
    w[1] = slot == 0 ? zero : make_uint4(b0.x, b0.y, b1.x, b1.y);                   // half 0, output 1
    w[2] = slot == 0 ? zero : make_uint4(b0.x, b0.y, b1.x, b1.y);                   // half 1, output 0
    w[3] = slot == 0 ? make_uint4(b0.x, b0.y, b1.x, b1.y) : make_uint4(n0.x, n0.y, n1.x, n1.y);   // half 1, output 1
    *reinterpret_cast<uint4*>(out + (row0) * 64 + r0) = w[0];
    *reinterpret_cast<uint4*>(out + (row0 + 2) * 64 + r0) = w[1];
    *reinterpret_cast<uint4*>(out + (row0 + 16) * 64 + r0) = w[2];
    *reinterpret_cast<uint4*>(out + (row0 + 18) * 64 + r0) = w[3];
  }
}

// CTA = one query: spectra of the zero-padded rows scaled by kappa / ||su_q||, the fp32 inverse norm of the query and its
// sweep constants qry_aux[q] = (1 or NaN, 4-norm of the normalised half spectrum).  Operand layout: [query tile of
// 128][slot][K half: Re(r) | Im(r)][query row][64 fp16], so the two K-major tiles of one slot of one query tile are 32
// contiguous KB (one TMA box).  Queries past Q in the last tile are written as zeros.
__global__ void __launch_bounds__(256)
spec_query_prep_kernel(const float* __restrict__ su, int64_t Q, int sw, __half* __restrict__ out, float2* __restrict__ qry_aux,
                       float* __restrict__ q_inv_norm, float* __restrict__ spec_out) {
  __shared__ float sre[kSpSlots][kSpCH + 1], sim[kSpSlots][kSpCH + 1];
  __shared__ float red[8];
  const int64_t q = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool live = q < Q;
  float e = 0.f;
  if (live) {
    RowFft fft;
    fft.init(lane);
    float2 z[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {     // all eight rows of this warp are in flight before the first transform
      const float* r = su + (q * kSpCH + warp + 8 * i) * sw;
      const int j = 2 * lane;
      z[i].x = j < sw ? __ldg(r + j) : 0.f;
      z[i].y = j + 1 < sw ? __ldg(r + j + 1) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      e = fmaf(z[i].x, z[i].x, e);
      e = fmaf(z[i].y, z[i].y, e);
      const float2 X = fft.run(z[i], lane);
      sre[lane][warp + 8 * i] = X.x;
      sim[lane][warp + 8 * i] = X.y;
    }
  }
  const float energy = block_sum_256(e, red);     // also the barrier between the transforms and their readers
  const float norm = sqrtf(energy);
  const float scale = (live && norm > 0.f) ? kSpecKappa / norm : 0.f;
  if (live) {
    const float s4 = block_sum_256(spectrum_quartic(sre, sim), red);
    if (threadIdx.x == 0) {
      q_inv_norm[q] = 1.0f / norm;
      qry_aux[q] = make_float2(norm > 0.f ? 1.0f : __int_as_float(0x7fc00000), sqrtf(sqrtf(s4)) * (scale / kSpecKappa));
    }
    if (spec_out != nullptr) {
      float2* dst = reinterpret_cast<float2*>(spec_out + q * (kSpCH * 64));
      for (int idx = threadIdx.x; idx < kSpCH * 32; idx += 256) dst[idx] = make_float2(sre[idx & 31][idx >> 5], sim[idx & 31][idx >> 5]);
    }
  }
  __half* tile = out + (q >> 7) * (int64_t)(2 * kSpSlots * 128 * 64) + (q & 127) * 64;
  {   // thread = (slot, eight feature rows): 16 bytes of the slot's Re row and of its Im row
    const int slot = threadIdx.x >> 3, r0 = (threadIdx.x & 7) * 8;
    uint4 wr = make_uint4(0u, 0u, 0u, 0u), wi = wr;
    if (live) {
      const float* a = &sre[slot][r0];
      const float* b = &sim[slot][r0];
      const uint2 a0 = pack4_f16(a[0] * scale, a[1] * scale, a[2] * scale, a[3] * scale), a1 = pack4_f16(a[4] * scale, a[5] * scale, a[6] * scale, a[7] * scale);
      const uint2 b0 = pack4_f16(b[0] * scale, b[1] * scale, b[2] * scale, b[3] * scale), b1 = pack4_f16(b[4] * scale, b[5] * scale, b[6] * scale, b[7] * scale);
      wr = make_uint4(a0.x, a0.y, a1.x, a1.y);
      wi = make_uint4(b0.x, b0.y, b1.x, b1.y);
    }
    *reinterpret_cast<uint4*>(tile + (int64_t)(2 * slot) * (128 * 64) + r0) = wr;
    *reinterpret_cast<uint4*>(tile + (int64_t)(2 * slot + 1) * (128 * 64) + r0) = wi;
  }
}

// ------------------------------------------------------------------------------------------
// PTX used only here
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float fmax3(float a, float b, float c) {   // FMNMX3; NaN operands are dropped like fmaxf drops them
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(taddr));
}

struct SpParams {
  SweepOut out;
  int n_qtiles, n_chunks, groups_per_chunk, n_groups;
  int debug;
  int one;                       // always 1; opaque to the compiler (see the epilogue)
};

// CG == 1: one CTA per 128-query x 8-item tile (tcgen05.mma.cta_group::1, M = 128, N = 16).
// CG == 2: a CTA pair per 128-query x 16-item tile (cta_group::2, M = 128, N = 32): each CTA stages 64 queries (the A rows
// of its half of M) and its own group of 8 items (its half of N, which the pair's MMA reads from both CTAs), and its TMEM holds
// D[m, n] at lane m + 64 (n / 16), column n % 16 -- lanes 0-63 = its 64 queries x the leader's group, lanes 64-127 = the same
// queries x the peer's group.  Still 1 024 pairs x 64 accumulators per CTA, but the ring carries 20 KB per slot instead of 36.
// EPI: what the epilogue extracts from a pair's 64 correlations.  0: the maximum alone (full panoramas, distances / ranks / top-k
// only: the crop is the whole item at every shift) -- 32 three-input maxima per item.  1: the maximum and its first shift
// (orientation output or cropped queries -- the crop norm depends on the shift -- without error bounds): compare and two selects
// per correlation.  2: with error bounds, the maximum, then one pass for the shifts within 2e of it (their count decides whether
// the pair is ambiguous, the lowest of them is the shift).
template <int CG, int EPI>
__global__ void __launch_bounds__(kSpThreads, 1)
match_spec_kernel(const __grid_constant__ CUtensorMap q_map, const __grid_constant__ CUtensorMap g_map, const SpParams P) {
  constexpr uint32_t kTmemCols = 512;
  constexpr int kA = CG == 2 ? kSpABytesPair : kSpABytes;   // query bytes per stage and CTA
  constexpr int NS = CG == 2 ? kSpStagesPair : kSpStages;   // ring depth
  constexpr int NP = CG == 2 ? kSpProducersPair : 2;        // producer warps; warp p owns the stages = p (mod NP)
  constexpr int NI = CG == 2 ? kSpIssuersPair : kSpStages;  // issuer warps; warp k owns the stages = k (mod NI)
  static_assert(NP + NI <= 8, "eight control warps");
  // kind::f16: D fp32 (bit 4), A and B fp16 (format fields 0), both K-major, N = 16 per CTA, M = 128
  constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)((16 * CG) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* a_base = smem;                                  // NS x kA, 1024-aligned
  unsigned char* b_base = smem + NS * kA;                        // NS x 4 KB, 1024-aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_base + NS * kSpBBytes);
  uint64_t* full = bars;                      // [NS]   (CG == 2: the leader's are used)
  uint64_t* empty = bars + NS;                // [NS]
  uint64_t* tmem_full = bars + 2 * NS;
  uint64_t* tmem_empty = tmem_full + 1;       //               (CG == 2: the leader's is used)
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t cta_rank = 0;
  if constexpr (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
  const bool leader = cta_rank == 0;
  const int unit = blockIdx.x / CG, n_units = gridDim.x / CG;
  const int n_work = P.n_chunks * P.n_qtiles;
  const bool chunk_major = SPEC_DBG(P, 16);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&q_map) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&g_map) : "memory");
  }
  if (warp == 1 && lane == 0) {
    if (s2u(smem) & 1023u) __trap();  // SWIZZLE_128B operand tiles need a 1024-byte aligned base
    for (uint32_t s = 0; s < NS; ++s) {
      bar_init(s2u(&full[s]), 1);
      bar_init(s2u(&empty[s]), 1);
    }
    bar_init(s2u(tmem_full), NI);          // one tcgen05.commit per issuing warp
    bar_init(s2u(tmem_empty), 8 * CG);     // one arrive per epilogue warp (of both CTAs)
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
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  // 512 threads start with 128 registers each; the eight control warps give theirs up so that an epilogue thread can
  // hold the 128 accumulators of two items plus the transform's temporaries (256 x 48 + 256 x 208 = 64 K registers; at 40 the
  // control loops spill their loop counters)
#ifndef SPEC_NREG_LO
#define SPEC_NREG_LO 48
#define SPEC_NREG_HI 208
#endif
#define SPEC_STR2(x) #x
#define SPEC_STR(x) SPEC_STR2(x)
  if (warp < 8) asm volatile("setmaxnreg.dec.sync.aligned.u32 " SPEC_STR(SPEC_NREG_LO) ";");
  else asm volatile("setmaxnreg.inc.sync.aligned.u32 " SPEC_STR(SPEC_NREG_HI) ";");

  if (warp < NP) {
    // ===================== TMA producers: warp p loads every NP-th slot of this CTA's slot sequence =====================
    // CG == 2: every stage's bytes -- this CTA's 64 queries and 8 items, and the peer's -- are accounted on the LEADER's full
    // barrier: the leader arms it with the pair's total, the peer only issues its copies.
    const uint32_t par = (uint32_t)warp;
    const uint32_t full0 = CG == 2 ? map_to_cta(s2u(&full[0]), 0) : s2u(&full[0]);
    uint32_t c0 = 0;         // sequence number of the current tile's slot 0
    for (int w = unit; w < n_work && !SPEC_DBG(P, 32); w += n_units) {
      const int qt = chunk_major ? w % P.n_qtiles : w / P.n_chunks;
      const int chunk = chunk_major ? w / P.n_qtiles : w - qt * P.n_chunks;
      const int q_row = qt * (2 * kSpSlots * 128);   // first operand row of this query tile
      const int grp0 = chunk * P.groups_per_chunk;
      const int grp1 = min(grp0 + P.groups_per_chunk, P.n_groups);
      for (int grp = grp0; grp < grp1; ++grp, c0 += kSpSlots) {
        const int my_grp = grp * CG + (int)cta_rank;   // (past the last group: the copy delivers zeros)
        for (uint32_t slot = (par + NP - c0 % NP) % NP; slot < kSpSlots; slot += NP) {
          const uint32_t c = c0 + slot, s = c % NS, ph = (c / NS) & 1u;
          bar_wait(s2u(&empty[s]), ph ^ 1);
          if (elect_one()) {
            const uint32_t fb = full0 + s * 8;
            if (SPEC_DBG(P, 8)) {
              if (leader) bar_expect_tx(s2u(&full[s]), (uint32_t)(CG * kSpBBytes));
            } else {
              if (leader) bar_expect_tx(s2u(&full[s]), (uint32_t)(CG * (kA + (SPEC_DBG(P, 256) ? kSpBBytes / 2 : kSpBBytes))));
              if constexpr (CG == 1) {
                tma_2d<1>(s2u(a_base) + s * kA, &q_map, fb, 0, q_row + 256 * (int)slot);
              } else {   // K halves of this CTA's 64 queries: rows [64 rank, +64) of each 128-row half
                tma_2d<2>(s2u(a_base) + s * kA, &q_map, fb, 0, q_row + 256 * (int)slot + 64 * (int)cta_rank);
                tma_2d<2>(s2u(a_base) + s * kA + kA / 2, &q_map, fb, 0, q_row + 256 * (int)slot + 128 + 64 * (int)cta_rank);
              }
            }
            tma_2d<CG>(s2u(b_base) + s * kSpBBytes, &g_map, fb, 0, (my_grp * kSpSlots + (int)slot) * 32);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp < NP + NI) {
    // ===================== MMA issuers: warp NP + k owns every NI-th slot of this CTA's slot sequence =====================
    // (CG == 2: the leader's warps issue for the pair)
    if (leader) {
    const uint32_t k = (uint32_t)(warp - NP);
    uint32_t n_tiles = 0;
    for (int w = unit; w < n_work; w += n_units) {
      const int chunk = chunk_major ? w / P.n_qtiles : w % P.n_chunks;
      const int grp0 = chunk * P.groups_per_chunk;
      n_tiles += (uint32_t)(min(grp0 + P.groups_per_chunk, P.n_groups) - grp0);
    }
    const uint64_t da0 = make_desc(s2u(a_base), 16, 1024, 2 /*SWIZZLE_128B*/);
    const uint64_t db0 = make_desc(s2u(b_base), 16, 1024, 2 /*SWIZZLE_128B*/);
    const uint32_t full0 = s2u(&full[0]), empty0 = s2u(&empty[0]);
    uint32_t seen_tile = 0xffffffffu;
    for (uint32_t c = k; c < n_tiles * kSpSlots; c += NI) {
      const uint32_t tile_it = c / kSpSlots, slot = c % kSpSlots;
      const uint32_t stg = c % NS, ph = (c / NS) & 1u;
      const uint32_t empty_k = empty0 + stg * 8;
      const uint64_t da = da0 + (uint64_t)(stg * (kA >> 4)), db = db0 + (uint64_t)(stg * (kSpBBytes >> 4));
      if (tile_it != seen_tile && !SPEC_DBG(P, 128)) {  // first slot of a new tile: the epilogue must have drained the previous one
        seen_tile = tile_it;
        bar_wait(s2u(tmem_empty), (tile_it & 1) ^ 1);
      }
      if (!SPEC_DBG(P, 32)) bar_wait(full0 + stg * 8, ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t tmem_d = tmem_base + slot * 16u;
        if (!SPEC_DBG(P, 4 | 32)) {
#pragma unroll
          for (int h = 0; h < 2; ++h)    // K halves: tiles of kA / 2 (queries) / 2 KB (items)
#pragma unroll
            for (int j = 0; j < 4; ++j)  // 32 bytes (2 descriptor units) along K per step inside the 128B-swizzled rows
              umma_bf16<CG>(tmem_d, da + (uint64_t)(h * (kA >> 5) + 2 * j), db + (uint64_t)(h * (kSpBBytes >> 5) + 2 * j), kIdesc,
                            (h | j) != 0 ? 1u : 0u);
        }
        if (!SPEC_DBG(P, 32)) umma_commit<CG>(empty_k);          // frees the stage (in both CTAs) when these MMAs retire
        if ((c + NI) / kSpSlots != tile_it && !SPEC_DBG(P, 128)) umma_commit<CG>(s2u(tmem_full));  // this warp's last slot of the tile
      }
      __syncwarp();
    }
    }
  } else if (warp >= 8) {
    // ===================== epilogue: warps 8-11 items 0-3, warps 12-15 items 4-7 of each group =====================
    const SweepOut& S = P.out;
    const int wq = warp & 3;
    const int ihalf = (warp - 8) >> 2;
    const int row = wq * 32 + lane;
    const int qrow = CG == 2 ? (int)cta_rank * 64 + (row & 63) : row;   // query of this lane within the 128-query tile
    const int gsel = CG == 2 ? row >> 6 : 0;                           // which group of the pair's two this lane holds
    const uint32_t lane_field = (uint32_t)(wq * 32) << 16;
    uint32_t tile_it = 0;
    for (int w = unit; w < n_work && !SPEC_DBG(P, 128); w += n_units) {
      const int qt = chunk_major ? w % P.n_qtiles : w / P.n_chunks;
      const int chunk = chunk_major ? w / P.n_qtiles : w - qt * P.n_chunks;
      const int64_t q = (int64_t)qt * 128 + qrow;
      SweepQuery qc;
      qc.load(S, q);
      int cnt = 0;
      float td[kSpTopkMax];
      int32_t ti[kSpTopkMax];
#pragma unroll
      for (int j = 0; j < kSpTopkMax; ++j) { td[j] = __int_as_float(0x7f800000); ti[j] = -1; }
      const int grp0 = chunk * P.groups_per_chunk;
      const int grp1 = min(grp0 + P.groups_per_chunk, P.n_groups);
      for (int grp = grp0; grp < grp1; ++grp, ++tile_it) {
        bar_wait(s2u(tmem_full), tile_it & 1);
        tc_fence_after();
        float bestv[4], sclv[4];
        int argv[4];   // bit 8: another shift lies within 2e of the maximum
#pragma unroll 1
        for (int pp = 0; pp < 2; ++pp) {
          const int i0 = ihalf * 4 + pp * 2;      // items i0 and i0 + 1 in the .x / .y halves
          const int64_t g0 = (int64_t)(grp * CG + gsel) * kSpItems + i0;
          // rounding scales of the two items for the ambiguity test (broadcast loads, consumed after the transform; rows past
          // G are zeros)
          float ag0 = 0.f, ag1 = 0.f;
          if (S.need_amb && (CG == 1 || g0 < S.G)) {   // (a pair's second group may lie past the tables)
            ag0 = __ldg(reinterpret_cast<const float*>(S.gal_aux + g0) + 2);
            ag1 = __ldg(reinterpret_cast<const float*>(S.gal_aux + g0 + 1) + 2);
          }
          float2 re[32], im[32], x[64];
          const uint32_t t0 = tmem_base + lane_field + (uint32_t)(2 * i0);
          if (SPEC_DBG(P, 2)) {
#pragma unroll
            for (int f = 0; f < 32; ++f) { re[f] = make_float2((float)(f + lane), (float)f); im[f] = make_float2((float)(f - pp), 1.f); }
          } else {
#pragma unroll
            for (int f = 0; f < 32; ++f) {
              uint32_t a, b, c, d;
              tmem_ld4(t0 + 16u * f, a, b, c, d);
              re[f] = make_float2(__uint_as_float(a), __uint_as_float(b));
              im[f] = make_float2(__uint_as_float(c), __uint_as_float(d));
            }
            tmem_ld_wait();
          }
          if (pp == 1) {  // this warp's share of the accumulators is in registers: hand TMEM back to the MMA issuers
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (CG == 1 || leader) bar_arrive_local(s2u(tmem_empty));
              else bar_arrive_remote(s2u(tmem_empty), 0);
            }
          }
          // The rest of the pair's work is pure arithmetic on the registers just loaded.  Left in the same basic block, ptxas
          // predicates the barrier arrive above and sinks it below the whole transform (nothing depends on it), so the MMA
          // issuers wait ~1000 instructions longer for every tile: 7.9 instead of 7.1 ms per 10k x 10k sweep (measured; an
          // empty asm that ties the registers orders the PTX but not ptxas's schedule).  A branch on a kernel parameter that
          // is always 1 gives the arithmetic a basic block of its own, after the arrive.
          if (P.one == 0) continue;
          if (SPEC_DBG(P, 1)) {
#pragma unroll
            for (int f = 0; f < 32; ++f) { x[2 * f] = re[f]; x[2 * f + 1] = im[f]; }
          } else {
            ifft64_hermitian_x2(re, im, x);
          }
          float best[2] = {-__int_as_float(0x7f800000), -__int_as_float(0x7f800000)};
          int arg[2] = {0, 0};
          if constexpr (EPI == 1) {
#pragma unroll
            for (int sft = 0; sft < 64; ++sft) {  // strict '>' in ascending shift order: first maximum
              if (x[sft].x > best[0]) { best[0] = x[sft].x; arg[0] = sft; }
              if (x[sft].y > best[1]) { best[1] = x[sft].y; arg[1] = sft; }
            }
          } else {   // the maximum alone: 3-input maxima
#pragma unroll
            for (int sft = 0; sft < 64; sft += 2) {
              best[0] = fmax3(best[0], x[sft].x, x[sft + 1].x);
              best[1] = fmax3(best[1], x[sft].y, x[sft + 1].y);
            }
          }
          if constexpr (EPI == 2) {
            // Shift and ambiguity in one pass: the shifts within 2e of the maximum are the candidates for the exact argmax.  One
            // candidate: it is the maximum's shift.  Several: the pair is ambiguous -- its slack covers every candidate's crop norm
            // (sweep_common.cuh) and matrix outputs are overwritten in fp32 -- and the lowest candidate stands in for the shift.
            const float thr0 = best[0] - 2.0f * sweep_err(S, ag0, qc, best[0]), thr1 = best[1] - 2.0f * sweep_err(S, ag1, qc, best[1]);
            int n0 = 0, n1 = 0;
#pragma unroll
            for (int sft = 63; sft >= 0; --sft) {
              const bool c0 = x[sft].x >= thr0, c1 = x[sft].y >= thr1;
              n0 += c0 ? 1 : 0;
              n1 += c1 ? 1 : 0;
              arg[0] = c0 ? sft : arg[0];
              arg[1] = c1 ? sft : arg[1];
            }
            arg[0] |= (n0 > 1) ? 256 : 0;
            arg[1] |= (n1 > 1) ? 256 : 0;
          }
#pragma unroll
          for (int e = 0; e < 2; ++e) {  // the scale gather is issued now and consumed after the other pair's transform
            const int64_t g = g0 + e;
            const float scl = (g < S.G) ? __ldg(S.gal_scale + g * 64 + (arg[e] & 63)) : 0.f;
            // static indices (the pair loop is not unrolled: two copies of the transform would not fit the registers)
            if (pp == 0) { bestv[e] = best[e]; argv[e] = arg[e]; sclv[e] = scl; }
            else { bestv[2 + e] = best[e]; argv[2 + e] = arg[e]; sclv[2 + e] = scl; }
          }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int64_t g = (int64_t)(grp * CG + gsel) * kSpItems + ihalf * 4 + e;
          if (g < S.G && qc.ok) {
            const float err = sweep_err(S, __ldg(reinterpret_cast<const float*>(S.gal_aux + g) + 2), qc, bestv[e]);   // L1-resident by now
            sweep_pair(S, qc, g, q, bestv[e], argv[e] & 63, (argv[e] & 256) != 0, sclv[e], err, cnt, td, ti);
          }
        }
      }
      if (qc.ok) {
        if (S.rank_count && cnt) atomicAdd(S.rank_count + q, cnt);
        if (S.topk > 0) {
          const int64_t slot = (int64_t)chunk * (2 * CG) + gsel * 2 + ihalf;
          float* od = S.topk_key + (slot * S.Q + q) * S.topk;
          int32_t* oi = S.topk_idx + (slot * S.Q + q) * S.topk;
#pragma unroll
          for (int j = 0; j < kSpTopkMax; ++j)
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

struct SpSchedule { int cg, n_units, n_qtiles, n_groups, groups_per_chunk, n_chunks; };   // n_groups: groups of 8 * cg items

// 2 = CTA pairs (the default), 1 = one CTA per tile; witw_match_spec_variant() switches (tools/, A/B measurements).
static int g_spec_cg = 2;

// Static schedule: work item w = (query tile, gallery chunk) goes to CTA w % n_units.  The number of chunks (<= 32: two
// candidate lists per chunk, witw_topk_merge takes 64) is chosen so that the busiest CTA carries as little more than the
// average as possible -- 10k x 10k: 28 chunks of 45 groups = 98.8 % balance against 93.4 % for the nearest "16 items per
// CTA" choice (30 chunks: two CTAs get a 17th item and the other 146 wait for them).
static double spec_balance(int n_groups, int n_qtiles, int n_chunks_want, int n_units, int* gpc_out, int* n_chunks_out) {
  const int gpc = (int)ceil_div<int64_t>(n_groups, n_chunks_want);
  const int nc = (int)ceil_div<int64_t>(n_groups, gpc);
  const int last = n_groups - (nc - 1) * gpc;           // groups of the last chunk
  // CTA u handles items u, u + n_units, ...; item w has chunk w % nc (query-tile-major order)
  const int64_t n_work = (int64_t)nc * n_qtiles;
  int64_t worst = 0;
  const int n_cta = (int)std::min<int64_t>(n_units, n_work);
  for (int u = 0; u < n_cta; ++u) {
    int64_t load = 0;
    for (int64_t w = u; w < n_work; w += n_units) load += (w % nc == nc - 1) ? last : gpc;
    worst = std::max(worst, load);
  }
  *gpc_out = gpc;
  *n_chunks_out = nc;
  return (double)n_groups * n_qtiles / ((double)worst * n_units);
}

static SpSchedule make_spec_schedule(int64_t G, int64_t Q) {
  static thread_local int64_t memo_g = -1, memo_q = -1;
  static thread_local SpSchedule memo;
  const int cg = (g_spec_cg == 2 && sm_count() >= 2) ? 2 : 1;
  if (G == memo_g && Q == memo_q && memo.cg == cg) return memo;
  SpSchedule s;
  s.cg = cg;
  s.n_units = std::max(1, sm_count() / cg);
  s.n_qtiles = (int)ceil_div<int64_t>(std::max<int64_t>(Q, 1), 128);
  s.n_groups = (int)ceil_div<int64_t>(std::max<int64_t>(G, 1), kSpItems * cg);
  double best = -1.0;
  s.groups_per_chunk = s.n_groups;
  s.n_chunks = 1;
  const int64_t n_work_cap = (int64_t)1 << 22;  // bound the search for huge query counts
  for (int c = 1; c <= kSpMaxChunks / cg && c <= s.n_groups; ++c) {
    if ((int64_t)c * s.n_qtiles > n_work_cap && c > 1) break;
    int gpc, nc;
    const double e = spec_balance(s.n_groups, s.n_qtiles, c, s.n_units, &gpc, &nc);
    if (e > best + 1e-9 || (e > best - 1e-9 && nc > s.n_chunks)) { best = e; s.groups_per_chunk = gpc; s.n_chunks = nc; }
  }
  memo = s; memo_g = G; memo_q = Q;
  return s;
}

}  // namespace witw

using namespace witw;

extern "C" int witw_spec_supported(int CH, int W, int sw) { return (CH == kSpCH && W == 64 && sw >= 1 && sw <= 64) ? 1 : 0; }

extern "C" size_t witw_spec_gallery_operand_bytes(int64_t G, int CH) {
  if (G < 0 || CH != kSpCH) { set_error(WITW_ERR_UNSUPPORTED, "witw_spec_gallery_operand_bytes: the spectral sweep needs C*H == %d (got %d)", kSpCH, CH); return 0; }
  return (size_t)std::max<int64_t>(ceil_div<int64_t>(G, kSpItems), 1) * kSpSlots * kSpBBytes;
}

extern "C" size_t witw_spec_query_operand_bytes(int64_t Q, int CH) {
  if (Q < 0 || CH != kSpCH) { set_error(WITW_ERR_UNSUPPORTED, "witw_spec_query_operand_bytes: the spectral sweep needs C*H == %d (got %d)", kSpCH, CH); return 0; }
  return (size_t)ceil_div<int64_t>(std::max<int64_t>(Q, 1), 128) * 128 * kSpSlots * 128 * 2;
}

extern "C" int witw_spec_gallery_prep(const float* ov, int64_t G, int64_t g_first, int CH, int W, int sw, void* gal_op,
                                      float* gal_scale, float* gal_aux, float* crop_inv_norm, float* spec_out, witw_stream_t stream) {
  WITW_REQUIRE(witw_spec_supported(CH, W, sw), WITW_ERR_UNSUPPORTED, "witw_spec_gallery_prep: needs C*H == 64, W == 64, 1 <= sw <= 64 (got %d, %d, %d)", CH, W, sw);
  WITW_REQUIRE(G >= 0 && g_first >= 0, WITW_ERR_INVALID, "witw_spec_gallery_prep: negative size");
  if (G == 0) return WITW_OK;
  WITW_REQUIRE(ov && gal_op && gal_scale && gal_aux, WITW_ERR_INVALID, "witw_spec_gallery_prep: null pointer");
  WITW_REQUIRE(((uintptr_t)gal_op & 127) == 0 && ((uintptr_t)ov & 7) == 0 && ((uintptr_t)gal_aux & 15) == 0, WITW_ERR_INVALID,
               "witw_spec_gallery_prep: operand must be 128-byte, gal_aux 16-byte, features 8-byte aligned");
  // items up to the end of the last group are written (zeros past G) so a partial group never exposes stale bytes
  const int64_t g_end = ceil_div<int64_t>(g_first + G, kSpItems) * kSpItems;
  const int64_t n = g_end - g_first;
  WITW_REQUIRE(n < (1ll << 31), WITW_ERR_INVALID, "witw_spec_gallery_prep: too many items in one call");
  WITW_REQUIRE(((uintptr_t)spec_out & 7) == 0, WITW_ERR_INVALID, "witw_spec_gallery_prep: spectra must be 8-byte aligned");
  spec_gallery_prep_kernel<<<(unsigned)n, 256, 0, as_stream(stream)>>>(ov, G, g_first, sw, reinterpret_cast<__half*>(gal_op), gal_scale,
                                                                     reinterpret_cast<float4*>(gal_aux), crop_inv_norm, spec_out);
  WITW_LAUNCH_CHECK();
  return WITW_OK;
}

extern "C" int witw_spec_query_prep(const float* su, int64_t Q, int CH, int sw, void* qry_op, float* qry_aux, float* q_inv_norm,
                                    float* spec_out, witw_stream_t stream) {
  WITW_REQUIRE(witw_spec_supported(CH, 64, sw), WITW_ERR_UNSUPPORTED, "witw_spec_query_prep: needs C*H == 64 and 1 <= sw <= 64 (got %d, %d)", CH, sw);
  WITW_REQUIRE(Q >= 0 && Q < (1ll << 31), WITW_ERR_INVALID, "witw_spec_query_prep: bad query count");
  if (Q == 0) return WITW_OK;
  WITW_REQUIRE(su && qry_op && qry_aux && q_inv_norm, WITW_ERR_INVALID, "witw_spec_query_prep: null pointer");
  WITW_REQUIRE(((uintptr_t)qry_op & 127) == 0 && ((uintptr_t)qry_aux & 7) == 0, WITW_ERR_INVALID,
               "witw_spec_query_prep: operand must be 128-byte, qry_aux 8-byte aligned");
  const int64_t q_pad = ceil_div<int64_t>(Q, 128) * 128;
  WITW_REQUIRE(((uintptr_t)spec_out & 7) == 0, WITW_ERR_INVALID, "witw_spec_query_prep: spectra must be 8-byte aligned");
  spec_query_prep_kernel<<<(unsigned)q_pad, 256, 0, as_stream(stream)>>>(su, Q, sw, reinterpret_cast<__half*>(qry_op),
                                                                        reinterpret_cast<float2*>(qry_aux), q_inv_norm, spec_out);
  WITW_LAUNCH_CHECK();
  return WITW_OK;
}

extern "C" int witw_match_spec_topk_slots(int64_t G, int64_t Q) {
  const SpSchedule s = make_spec_schedule(G, Q);
  return s.n_chunks * 2 * s.cg;
}

extern "C" int witw_match_spec_variant(int cta_group) {
  WITW_REQUIRE(cta_group == 1 || cta_group == 2, WITW_ERR_INVALID, "witw_match_spec_variant: 1 (one CTA per tile) or 2 (CTA pairs), got %d", cta_group);
  g_spec_cg = cta_group;
  return WITW_OK;
}

extern "C" int witw_match_spec(const witw_sweep_args* a, witw_stream_t stream) {
  WITW_REQUIRE(a != nullptr, WITW_ERR_INVALID, "witw_match_spec: null arguments");
  const int64_t G = a->G, Q = a->Q;
  WITW_REQUIRE(G >= 0 && Q >= 0 && witw_spec_supported(a->CH, 64, a->sw), WITW_ERR_UNSUPPORTED, "witw_match_spec: unsupported CH=%d sw=%d", a->CH, a->sw);
  if (G == 0 || Q == 0) return WITW_OK;
  SweepOut out;
  int rc = fill_sweep_out("witw_match_spec", a, kSpecUnit, kSpecAccRel, &out);
  if (rc != WITW_OK) return rc;
  WITW_REQUIRE(((uintptr_t)a->qry_op & 127) == 0 && ((uintptr_t)a->gal_op & 127) == 0, WITW_ERR_INVALID, "witw_match_spec: operands must be 128-byte aligned");
  rc = witw_device_check();
  if (rc != WITW_OK) return rc;

  const SpSchedule sch = make_spec_schedule(G, Q);
  typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  auto encode = reinterpret_cast<EncodeTiledFn>(get_encode_tiled());
  WITW_REQUIRE(encode != nullptr, WITW_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const cuuint32_t estr[2] = {1, 1};
  // query spectra: [query tiles x 32 slots x 2 K halves x 128 rows][64] fp16; one box = the 256 rows of one slot of one tile
  CUtensorMap qmap;
  const int64_t q_rows = ceil_div<int64_t>(Q, 128) * (2 * kSpSlots * 128);
  WITW_REQUIRE(q_rows < (1ll << 31), WITW_ERR_UNSUPPORTED, "witw_match_spec: %lld queries are too many for one sweep", (long long)Q);
  const cuuint64_t qdims[2] = {64, (cuuint64_t)q_rows};
  const cuuint64_t qstrides[1] = {128};
  const cuuint32_t qbox[2] = {64, sch.cg == 2 ? 64u : 256u};   // CTA pair: one K half of a CTA's 64 queries per copy
  CUresult cr = encode(&qmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(a->qry_op), qdims, qstrides, qbox, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  WITW_REQUIRE(cr == CUDA_SUCCESS, WITW_ERR_CUDA, "cuTensorMapEncodeTiled(query spectra) failed with CUresult %d", (int)cr);
  // gallery spectra: [groups x 32 slots x 2 K halves x 16 rows][64] fp16; one box = the 32 rows of one slot
  CUtensorMap gmap;
  const int64_t total_rows = ceil_div<int64_t>(G, kSpItems) * kSpSlots * 32;
  WITW_REQUIRE(total_rows < (1ll << 31), WITW_ERR_UNSUPPORTED, "witw_match_spec: gallery of %lld items is too large for one sweep", (long long)G);
  const cuuint64_t gdims[2] = {64, (cuuint64_t)total_rows};
  const cuuint64_t gstrides[1] = {128};
  const cuuint32_t gbox[2] = {64, (spec_debug() & 256) ? 16u : 32u};   // (hooks build, bit 8: what would half the gallery bytes buy?)
  cr = encode(&gmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(a->gal_op), gdims, gstrides, gbox, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  WITW_REQUIRE(cr == CUDA_SUCCESS, WITW_ERR_CUDA, "cuTensorMapEncodeTiled(gallery spectra) failed with CUresult %d", (int)cr);

  SpParams P;
  std::memset(&P, 0, sizeof(P));
  P.out = out;
  P.n_qtiles = sch.n_qtiles; P.n_chunks = sch.n_chunks; P.groups_per_chunk = sch.groups_per_chunk; P.n_groups = sch.n_groups;
  P.debug = spec_debug();
  P.one = 1;
  const int ns = sch.cg == 2 ? kSpStagesPair : kSpStages;
  const size_t smem = (size_t)ns * ((sch.cg == 2 ? kSpABytesPair : kSpABytes) + kSpBBytes) + (2 * ns + 2) * 8 + 16;
  const int grid = std::min(sch.n_units, sch.n_chunks * sch.n_qtiles) * sch.cg;
  const int epi = (a->sw < 64 || a->ori != nullptr) ? (out.need_amb ? 2 : 1) : 0;
  auto kernel = sch.cg == 2 ? (epi == 2 ? match_spec_kernel<2, 2> : epi == 1 ? match_spec_kernel<2, 1> : match_spec_kernel<2, 0>)
                            : (epi == 2 ? match_spec_kernel<1, 2> : epi == 1 ? match_spec_kernel<1, 1> : match_spec_kernel<1, 0>);
  WITW_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid, 1, 1);
  cfg.blockDim = dim3(kSpThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = as_stream(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)sch.cg;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  WITW_CUDA(cudaLaunchKernelEx(&cfg, kernel, qmap, gmap, P));
  WITW_LAUNCH_CHECK();
  return WITW_OK;
}
