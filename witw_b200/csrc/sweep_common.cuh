// What the two tensor-core sweeps (match_spec.cu, match_tc.cu) share: the operand scaling, the bound on what the
// fp16 operands can do to a result, and the tail of the epilogue that turns (max correlation, shift) of one
// (gallery item, query) pair into a distance, a rank decision, a top-k candidate -- or into an entry of the query's
// deferral list when fp16 arithmetic cannot settle it (finish.cu settles those in fp32).
//
// Operands.  Both feature sets are normalised before they are rounded to fp16:
//     gallery operand = fp16( kappa * F(ov_g) / ||ov_g|| ),   query operand = fp16( kappa * F(su_q) / ||su_q|| )
// (F = 64-point row spectra for the spectral sweep, the identity for the dense one), so every item uses the same
// part of fp16's range whatever the scale of the encoder's output, and the accumulator of shift s is
//     acc[s] = unit * c[s],   c[s] = corr[g,q,s] / (||ov_g|| ||su_q||)  in [-1, 1],   unit = 64 kappa^2 | kappa^2.
// The distance of cvig_fov.py:351-361 is then   d = 2 - 2 acc[s*] gal_scale[g,s*] qry_aux[q].x   with
//     gal_scale[g,s] = ||ov_g|| / (||crop(ov_g, s)|| unit)     (= 1/unit for a full panorama)
//     qry_aux[q].x   = 1, or NaN for a zero-norm query (the reference divides by the norm: no epsilon)
//
// Error bound.  Rounding to fp16 moves a stored value v by an error of variance <= (2^-10 v)^2 / 12.  The errors
// of the terms of c[s] are independent, so   var(c[s]) <= (2^-20/12) * 2 * w^2 * sum_i |o_i|^2 |s_i|^2
// <= (2^-20/12) * 2 * w^2 * ||o||_4^2 ||s||_4^2   (Cauchy-Schwarz; w = 2/64 for the half-spectrum sums, 1 for the
// direct form).  The prep kernels store  gal_aux[g].z = sqrt(2 * 2^-20/12) * w * ||o_g||_4 * unit  and
// qry_aux[q].y = ||s_q||_4 (4-norms of the normalised operands), and a pair's bound is
//     e(g,q) = max( err_sigmas * gal_aux[g].z * qry_aux[q].y,  kAccFloor * unit ) + acc_rel |acc|     [accumulator units]
// gal_aux[g] = (max_s gal_scale[g,:], max - min of gal_scale[g,:], the rounding scale above, the operand's kappa / ||ov_g||).
// err_sigmas standard deviations of a bound on the standard deviation (default 5; on Gaussian features the bound
// itself is ~2.8x the measured deviation), floored for the fp32 inverse transform, plus a term for the tensor core's
// fp32 accumulation: every accumulation step can lose an ulp of the running sum (measured on B200: the 256 sequential
// steps of the dense sweep's K = 4096 move a strongly correlated pair by 1.7e-5 of its accumulator, always downwards --
// truncation, not rounding; the spectral sweep accumulates 8 steps per frequency slot).
//
// Consequences.  Let s* be the sweep's argmax.  The exact argmax lies among the shifts with acc[s] >= acc[s*] - 2e.
//   * unambiguous pair (no other shift in that window): |d_exact - d| <= 2 e gal_scale[g,s*] =: slack
//   * ambiguous pair: the exact result may use another shift's crop norm:
//         slack = 2 |acc[s*]| (max_s gal_scale[g,:] - min_s gal_scale[g,:]) + 6 e max_s gal_scale[g,:]
//   * a rank decision d <= d_true[q] (cvig_fov.py:552) is final when |d - d_true| > slack; otherwise the pair goes on
//     the query's list (tag kTagRank) and finish.cu decides it from the fp32 spectra
//   * a top-k candidate is kept under the key d - slack, a lower bound of its exact distance; finish.cu re-ranks
//     the candidates in fp32 and proves from the keys that no item outside the list can belong to the top k
//   * matrix outputs (dist / ori): ambiguous pairs and pairs with slack > fix_rel * d (the bound no longer guarantees
//     the relative tolerance: the near matches) go on the list too (tag 0) and are overwritten with their fp32 values
// A query whose list overflows is re-done entirely in fp32 by the caller (ops.py), so nothing is ever decided in fp16
// inside the slack.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "common.cuh"

namespace witw {

constexpr float kSpecKappa = 16.0f;                       // spectral operands: typical component ~1.4, largest possible 128
constexpr float kSpecUnit = 64.0f * kSpecKappa * kSpecKappa;
constexpr float kDenseKappa = 256.0f;                     // dense operands: typical element ~4, largest possible 256
constexpr float kDenseUnit = kDenseKappa * kDenseKappa;
constexpr float kRoundSigma = 3.9867e-4f;                 // sqrt(2 * 2^-20 / 12)
constexpr float kAccFloor = 2e-6f;                        // fp32 inverse transform / table round-off, in units of c[s]
constexpr float kSpecAccRel = 2e-6f;                      // accumulation loss relative to |acc|: 8 steps per slot, x4 margin
constexpr float kDenseAccRel = 5e-5f;                     // 256 steps: measured 1.7e-5, x3 margin
constexpr int kSweepTopk = 16;                            // register-resident candidates per query and gallery chunk

constexpr uint32_t kTagRank = 0x80000000u;                // list entry: the rank decision of this pair is pending
constexpr uint32_t kTagMask = 0x3fffffffu;                // local gallery index of a list entry

// What a sweep produces; identical for both kernels (filled from witw_sweep_args).
struct SweepOut {
  const float* gal_scale;        // [G_pad][64]
  const float4* gal_aux;         // [G_pad]
  const float2* qry_aux;         // [Q]
  float* dist;                   // [G][Q] or null
  uint8_t* ori;                  // [G][Q] or null
  const float* d_true;           // [Q] or null
  const int32_t* true_idx;       // [Q] or null
  int32_t* rank_count;           // [Q] or null
  float* topk_key;               // [n_slots][Q][topk] or null
  int32_t* topk_idx;
  int32_t* list_g;               // [Q][list_cap] or null
  int32_t* list_n;               // [Q]
  int64_t G, Q;
  int32_t topk;
  int32_t g_offset;
  int32_t list_cap;
  float err_sigmas;              // 0: no bounds -- every decision is taken from the fp16 result, keys are distances
  float acc_floor;               // kAccFloor * unit
  float acc_rel;                 // accumulation loss of the tensor core relative to |acc|
  float fix_rel;
  int need_amb;                  // the epilogue must look for a second shift within 2e of the maximum
};

// Per-query constants of an epilogue thread.
struct SweepQuery {
  float qfac, bq, dtrue;
  int32_t self_g;                // local index of the query's match (counted by index: d[idx] <= d[idx] in the reference), or -1
  bool ok;
  __device__ __forceinline__ void load(const SweepOut& P, int64_t q) {
    ok = q < P.Q;
    const float2 a = ok ? P.qry_aux[q] : make_float2(0.f, 0.f);
    qfac = a.x;
    bq = a.y;
    dtrue = (ok && P.d_true) ? P.d_true[q] : __int_as_float(0x7fc00000);
    self_g = (ok && P.true_idx) ? P.true_idx[q] - P.g_offset : -1;
  }
};

// Bound e on the error of the accumulators of a pair (item with rounding scale ag = gal_aux[g].z, this query) whose
// largest accumulator is acc.
__device__ __forceinline__ float sweep_err(const SweepOut& P, float ag, const SweepQuery& qc, float acc) {
  return fmaxf(P.err_sigmas * ag * qc.bq, P.acc_floor) + P.acc_rel * fabsf(acc);
}

__device__ __forceinline__ void sweep_list_append(const SweepOut& P, int64_t q, uint32_t entry) {
  const int32_t pos = atomicAdd(P.list_n + q, 1);        // a count above list_cap marks the query for the fp32 fallback
  if (pos < P.list_cap) P.list_g[q * P.list_cap + pos] = (int32_t)entry;
}

// The tail of the epilogue for one pair: acc = accumulator at the sweep's argmax, amb = another shift within 2e of it,
// scale = gal_scale[g, arg], e = sweep_err().
__device__ __forceinline__ void sweep_pair(const SweepOut& P, const SweepQuery& qc, int64_t g, int64_t q, float acc, int arg, bool amb,
                                           float scale, float e, int& cnt, float (&td)[kSweepTopk], int32_t (&ti)[kSweepTopk]) {
  const float d = 2.0f * (1.0f - acc * scale * qc.qfac);
  float slack = 0.f;
  const bool bounded = P.err_sigmas > 0.f;
  if (bounded) {
    slack = 2.0f * e * scale;
    if (amb) {  // rare: the exact argmax may be another shift, with another crop norm
      const float2 ms = __ldg(reinterpret_cast<const float2*>(P.gal_aux + g));
      slack = 2.0f * fabsf(acc) * ms.y + 6.0f * e * ms.x;
    }
  }
  uint32_t entry = 0;
  bool listed = false;
  if (P.dist) P.dist[g * P.Q + q] = d;
  if (P.ori) P.ori[g * P.Q + q] = (uint8_t)arg;
  if (bounded && (P.dist || P.ori) && (amb || slack > P.fix_rel * d)) listed = true;
  if (P.rank_count) {
    if ((int32_t)g == qc.self_g) {
      cnt += (qc.dtrue == qc.dtrue) ? 1 : 0;
    } else if (bounded && P.list_g && fabsf(d - qc.dtrue) <= slack) {
      entry |= kTagRank;
      listed = true;
    } else {
      cnt += (d <= qc.dtrue) ? 1 : 0;
    }
  }
  if (listed && P.list_g) sweep_list_append(P, q, entry | (uint32_t)g);
  if (P.topk > 0) {
    float ck = d - slack;                                // NaN never enters: the comparisons below are false
    if (ck < td[kSweepTopk - 1]) {
      int32_t ci = (int32_t)g + P.g_offset;
#pragma unroll
      for (int j = 0; j < kSweepTopk; ++j) {             // ascending register list, strict '<': the earlier item wins ties
        if (ck < td[j]) {
          const float t0 = td[j]; const int32_t t1 = ti[j];
          td[j] = ck; ti[j] = ci; ck = t0; ci = t1;
        }
      }
    }
  }
}

// Host: validate the arguments of a sweep and translate them for the kernels.  unit: see above.
inline int fill_sweep_out(const char* fn, const witw_sweep_args* a, float unit, float acc_rel, SweepOut* o) {
  WITW_REQUIRE(a->gal_op && a->gal_scale && a->gal_aux && a->qry_op && a->qry_aux, WITW_ERR_INVALID, "%s: null operand", fn);
  WITW_REQUIRE(a->topk >= 0 && a->topk <= kSweepTopk, WITW_ERR_UNSUPPORTED, "%s: fused top-k supports k <= %d (got %d)", fn, kSweepTopk, a->topk);
  WITW_REQUIRE(a->topk == 0 || (a->topk_key && a->topk_idx), WITW_ERR_INVALID, "%s: top-k buffers missing", fn);
  WITW_REQUIRE(!a->rank_count || a->d_true, WITW_ERR_INVALID, "%s: rank_count needs d_true", fn);
  WITW_REQUIRE(a->err_sigmas >= 0.f && a->list_cap >= 0 && (a->list_cap == 0 || (a->list_g && a->list_n && a->err_sigmas > 0.f)), WITW_ERR_INVALID,
               "%s: deferral lists need list_g, list_n and err_sigmas > 0", fn);
  WITW_REQUIRE(a->G < (1ll << 30) && a->Q < (1ll << 31), WITW_ERR_INVALID, "%s: sizes exceed 2^30 items / 2^31 queries", fn);
  WITW_REQUIRE((((uintptr_t)a->gal_aux & 15) | ((uintptr_t)a->qry_aux & 7)) == 0, WITW_ERR_INVALID, "%s: gal_aux must be 16-byte, qry_aux 8-byte aligned", fn);
  o->gal_scale = a->gal_scale;
  o->gal_aux = reinterpret_cast<const float4*>(a->gal_aux);
  o->qry_aux = reinterpret_cast<const float2*>(a->qry_aux);
  o->dist = a->dist; o->ori = a->ori; o->d_true = a->d_true; o->true_idx = a->true_idx; o->rank_count = a->rank_count;
  o->topk_key = a->topk_key; o->topk_idx = a->topk_idx;
  o->list_g = a->list_cap > 0 ? a->list_g : nullptr; o->list_n = a->list_n;
  o->G = a->G; o->Q = a->Q; o->topk = a->topk; o->g_offset = a->g_index_offset; o->list_cap = a->list_cap;
  o->err_sigmas = a->err_sigmas; o->acc_floor = kAccFloor * unit; o->acc_rel = acc_rel; o->fix_rel = a->fix_rel;
  o->need_amb = (a->err_sigmas > 0.f && (a->sw < 64 || a->ori != nullptr)) ? 1 : 0;
  return WITW_OK;
}

}  // namespace witw
