// K1 -- aerial polar transform (bilinear resample of overhead tiles into panorama-aligned
// polar images).  Replaces model/cvig_fov.py:156-209 (bilinear_interpolate + PolarTransform).
//
// Two device paths:
//   * witw_bilinear_gather_f32: generic geometry, taps and fp32 weights from a host-built
//     float64 table (exactly the reference's arithmetic) -> bit-exact with the reference.
//   * witw_polar_resample_f32: the throughput path.  One persistent CTA per SM, pinned to one
//     azimuth quadrant.  The quadrant's source window (a 132x129-float box, 1.6 % larger
//     than the pixels it needs) is streamed through a 3-stage TMA ring in shared memory;
//     each of the 1024 threads keeps the sample table of its 16 output pixels in registers
//     for the whole launch (no table traffic), gathers 4 taps per pixel from shared memory
//     and stores with lanes on consecutive output columns (128-byte coalesced rows).
//     HBM traffic is the algorithmic minimum: every source byte is read once (plus the box
//     margin), every output byte written once.
#include <cuda.h>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace witw {

// ------------------------------------------------------------------------------------------
// host side: grid and table (float64, mirrors the reference's evaluation order)
// ------------------------------------------------------------------------------------------
static void polar_grid_host(int h_s, int w_s, int s_o, double* x, double* y) {
  const double half = s_o / 2.0;  // cvig_fov.py:198 (s_o/2) is a true division in Python 3
  const double two_pi = 2 * 3.141592653589793;  // math.pi
  for (int r = 0; r < h_s; ++r) {
    // (s_o/2) * (h_s - 1 - yy) / h_s : product first, then the division (left to right)
    const double radial = half * (double)(h_s - 1 - r) / (double)h_s;
    for (int c = 0; c < w_s; ++c) {
      const double ang = two_pi * (double)c / (double)w_s;
      y[(size_t)r * w_s + c] = half + radial * std::cos(ang);
      x[(size_t)r * w_s + c] = half - radial * std::sin(ang);
    }
  }
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

static void bilinear_lut_host(const double* x, const double* y, int64_t n, int src_h, int src_w,
                              int32_t* idx4, float* w4) {
  for (int64_t i = 0; i < n; ++i) {
    int x0 = (int)std::floor(x[i]), y0 = (int)std::floor(y[i]);
    int x1 = x0 + 1, y1 = y0 + 1;
    x0 = clampi(x0, 0, src_w - 1);
    x1 = clampi(x1, 0, src_w - 1);
    y0 = clampi(y0, 0, src_h - 1);
    y1 = clampi(y1, 0, src_h - 1);
    idx4[4 * i + 0] = x0;
    idx4[4 * i + 1] = x1;
    idx4[4 * i + 2] = y0;
    idx4[4 * i + 3] = y1;
    // float64 products of the clipped differences, then one rounding to fp32 (cvig_fov.py:178-181)
    w4[4 * i + 0] = (float)(((double)x1 - x[i]) * ((double)y1 - y[i]));
    w4[4 * i + 1] = (float)(((double)x1 - x[i]) * (y[i] - (double)y0));
    w4[4 * i + 2] = (float)((x[i] - (double)x0) * ((double)y1 - y[i]));
    w4[4 * i + 3] = (float)((x[i] - (double)x0) * (y[i] - (double)y0));
  }
}

// ------------------------------------------------------------------------------------------
// generic gather kernel (bit-exact path)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float gather_value(const float* s, int o, const float*, int) { return __ldg(s + o); }
__device__ __forceinline__ float gather_value(const uint8_t* s, int o, const float* lut, int c) { return __ldg(lut + c * 256 + __ldg(s + o)); }

template <typename T>
__global__ void __launch_bounds__(256) bilinear_gather_kernel(const T* __restrict__ src,
                                                              float* __restrict__ dst,
                                                              const int4* __restrict__ idx4,
                                                              const float4* __restrict__ w4,
                                                              int64_t n_img, int src_h, int src_w,
                                                              int64_t n_out, int img_per_block,
                                                              const float* __restrict__ lut, int n_ch) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_out) return;
  const int4 t = idx4[p];  // x0, x1, y0, y1
  const float4 w = w4[p];  // wa, wb, wc, wd
  const int oa = t.z * src_w + t.x, ob = t.w * src_w + t.x, oc = t.z * src_w + t.y, od = t.w * src_w + t.y;
  const int64_t plane = (int64_t)src_h * src_w;
  const int64_t i0 = (int64_t)blockIdx.y * img_per_block;
  const int64_t i1 = min(i0 + img_per_block, n_img);
  for (int64_t i = i0; i < i1; ++i) {
    const T* s = src + i * plane;
    const int ch = n_ch > 0 ? (int)(i % n_ch) : 0;
    const float a = gather_value(s, oa, lut, ch), b = gather_value(s, ob, lut, ch), c = gather_value(s, oc, lut, ch),
                d = gather_value(s, od, lut, ch);
    // ((wa*Ia + wb*Ib) + wc*Ic) + wd*Id, no contraction (cvig_fov.py:183)
    float r = __fadd_rn(__fmul_rn(w.x, a), __fmul_rn(w.y, b));
    r = __fadd_rn(r, __fmul_rn(w.z, c));
    r = __fadd_rn(r, __fmul_rn(w.w, d));
    dst[i * n_out + p] = r;
  }
}

// ------------------------------------------------------------------------------------------
// fast path: plan
// ------------------------------------------------------------------------------------------
constexpr int kPolarThreads = 1024;
constexpr int kPolarPx = 16;  // output pixels per thread per quadrant
constexpr int kBoxW = 132;    // floats per staged source row (multiple of 4: TMA inner box is 16-byte granular)
constexpr int kBoxWU8 = 160;  // bytes per staged row of a uint8 source (box start on a 16-pixel boundary: up to 15 + 130 columns)
constexpr int kBoxH = 129;
constexpr int kPolarStages = 3;
constexpr uint32_t kPlanMagic = 0x57495031u;  // "WIP1"

struct PolarException {  // a pixel whose taps the reference clips: patched with the exact formula
  int32_t pix;           // row * w_s + col
  int32_t x0, x1, y0, y1;
  float w[4];
};

struct PolarPlanHeader {
  uint32_t magic;
  int32_t h_s, w_s, s_o;
  int32_t box_x0[4], box_y0[4];
  int32_t n_exc;
  int32_t patch_w;   // a warp's 32 lanes cover patch_w columns x 32/patch_w rows of the output
  uint32_t lut_off;  // byte offset of the register table: float fx[4][PX][1024], float fy[..], uint32 off2[4][PX/2][1024]
                     // (off2 packs the 16-bit box offsets of pixels 2i and 2i+1)
  uint32_t exc_off;  // byte offset of PolarException[n_exc]
  uint32_t total_bytes;
  int32_t box_w;      // elements per staged source row (kBoxW for fp32 sources, kBoxWU8 for uint8 sources)
  int32_t elem_bytes; // 4 or 1
};

static bool polar_fast_supported(int h_s, int w_s, int s_o) {
  if (w_s % 128 != 0 || s_o < 8 || h_s < 1) return false;
  if ((w_s / 4) % 32 != 0 || (kPolarThreads / 32) % ((w_s / 4) / 8) != 0) return false;
  return (int64_t)h_s * (w_s / 4) == (int64_t)kPolarThreads * kPolarPx;
}

static int polar_patch_w() {
  // Lanes of a warp on a 2-D output patch touch a compact 2-D footprint of the staged source window, which
  // spreads over the shared-memory banks better than 32 samples along one arc (measured: profiles/).
  // measured on B200, 1024x3 planes: 32 -> 5053 GB/s, 16 -> 5515 GB/s, 8 -> 5332 GB/s.  The width is a constant of the
  // shipped library; only builds with -DWITW_DEBUG_HOOKS (tools/ probes) read it from the environment.
#ifdef WITW_DEBUG_HOOKS
  static int v = -1;
  if (v < 0) {
    const char* e = std::getenv("WITW_POLAR_PW");
    v = e ? std::atoi(e) : 16;
    if (v != 8 && v != 16 && v != 32) v = 16;
  }
  return v;
#else
  return 16;
#endif
}

// pixel (row, col-in-quadrant) handled by thread t in iteration i: warps tile the quadrant with patch_w x 32/patch_w patches
static inline void polar_thread_pixel(int i, int t, int qw, int pw, int* row, int* col) {
  const int ph = 32 / pw, npx = qw / pw;
  const int patch = i * (kPolarThreads / 32) + (t >> 5), lane = t & 31;
  *col = (patch % npx) * pw + lane % pw;
  *row = (patch / npx) * ph + lane / pw;
}

// Builds header + tables into `out` (may be null to only count).  Returns bytes, 0 if unsupported.
static size_t polar_plan_build_host(int h_s, int w_s, int s_o, void* out, int elem_bytes = 4) {
  if (!polar_fast_supported(h_s, w_s, s_o)) return 0;
  const int box_w = elem_bytes == 4 ? kBoxW : kBoxWU8;
  const int align = 16 / elem_bytes;  // elements per 16 bytes
  const int64_t n = (int64_t)h_s * w_s;
  std::vector<double> x(n), y(n);
  polar_grid_host(h_s, w_s, s_o, x.data(), y.data());
  std::vector<int32_t> idx(4 * n);
  std::vector<float> w(4 * n);
  bilinear_lut_host(x.data(), y.data(), n, s_o, s_o, idx.data(), w.data());
  const int qw = w_s / 4;
  std::vector<PolarException> exc;
  std::vector<char> is_exc(n, 0);
  int bx0[4], by0[4], bx1[4], by1[4];
  for (int q = 0; q < 4; ++q) { bx0[q] = by0[q] = 1 << 30; bx1[q] = by1[q] = -(1 << 30); }
  for (int64_t p = 0; p < n; ++p) {
    const int fx0 = (int)std::floor(x[p]), fy0 = (int)std::floor(y[p]);
    const bool clipped = idx[4 * p] != fx0 || idx[4 * p + 1] != fx0 + 1 || idx[4 * p + 2] != fy0 || idx[4 * p + 3] != fy0 + 1;
    if (clipped) {
      PolarException e;
      e.pix = (int32_t)p;
      e.x0 = idx[4 * p]; e.x1 = idx[4 * p + 1]; e.y0 = idx[4 * p + 2]; e.y1 = idx[4 * p + 3];
      for (int k = 0; k < 4; ++k) e.w[k] = w[4 * p + k];
      exc.push_back(e);
      is_exc[p] = 1;
      continue;
    }
    const int q = (int)((p % w_s) / qw);
    bx0[q] = std::min(bx0[q], fx0); bx1[q] = std::max(bx1[q], fx0 + 1);
    by0[q] = std::min(by0[q], fy0); by1[q] = std::max(by1[q], fy0 + 1);
  }
  for (int q = 0; q < 4; ++q) {
    // measured on B200 (tools/probe/tma_probe.cu): a tile-mode TMA whose innermost start coordinate is not a
    // multiple of 16 bytes raises an illegal-instruction fault, so the box starts on a 4-float boundary
    bx0[q] &= ~(align - 1);
    if (bx1[q] - bx0[q] + 1 > box_w || by1[q] - by0[q] + 1 > kBoxH) return 0;
  }
  const size_t lut_elems = (size_t)4 * kPolarPx * kPolarThreads;
  PolarPlanHeader h;
  std::memset(&h, 0, sizeof(h));
  h.magic = kPlanMagic; h.h_s = h_s; h.w_s = w_s; h.s_o = s_o;
  for (int q = 0; q < 4; ++q) { h.box_x0[q] = bx0[q]; h.box_y0[q] = by0[q]; }
  h.n_exc = (int32_t)exc.size();
  h.patch_w = polar_patch_w();
  h.box_w = box_w; h.elem_bytes = elem_bytes;
  h.lut_off = 256;
  h.exc_off = (uint32_t)(h.lut_off + 3 * lut_elems * 4);
  h.total_bytes = (uint32_t)(h.exc_off + std::max<size_t>(exc.size(), 1) * sizeof(PolarException));
  if (out == nullptr) return h.total_bytes;
  char* base = (char*)out;
  std::memset(base, 0, h.total_bytes);
  std::memcpy(base, &h, sizeof(h));
  float* fx = (float*)(base + h.lut_off);
  float* fy = fx + lut_elems;
  uint32_t* off = (uint32_t*)(fy + lut_elems);
  for (int q = 0; q < 4; ++q)
    for (int i = 0; i < kPolarPx; ++i)
      for (int t = 0; t < kPolarThreads; ++t) {
        int row, qcol;
        polar_thread_pixel(i, t, qw, h.patch_w, &row, &qcol);
        const int col = q * qw + qcol;
        const int64_t p = (int64_t)row * w_s + col;
        const size_t o = ((size_t)q * kPolarPx + i) * kPolarThreads + t;
        const size_t o2 = ((size_t)q * (kPolarPx / 2) + i / 2) * kPolarThreads + t;
        const int sh = (i & 1) * 16;
        if (is_exc[p]) { fx[o] = 0.f; fy[o] = 0.f; continue; }  // offset 0, patched afterwards
        const int fx0 = (int)std::floor(x[p]), fy0 = (int)std::floor(y[p]);
        fx[o] = (float)(x[p] - (double)fx0);
        fy[o] = (float)(y[p] - (double)fy0);
        off[o2] |= (uint32_t)((fy0 - by0[q]) * box_w + (fx0 - bx0[q])) << sh;
      }
  if (!exc.empty()) std::memcpy(base + h.exc_off, exc.data(), exc.size() * sizeof(PolarException));
  return h.total_bytes;
}

// ------------------------------------------------------------------------------------------
// fast path: kernel
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

struct PolarQuadBoxes { int x0[4], y0[4]; };

// uint8 sources: per-channel affine form of the reference's ImageNormalization (cvig_fov.py:137-149),
// value = a[c] * pixel + b[c]; plane p of the batch is channel p % n_ch.  fp32 sources: unused.
constexpr int kMaxNormCh = 8;
struct PolarNorm { float a[kMaxNormCh], b[kMaxNormCh]; int n_ch; };

template <typename T, int BOXW>
__global__ void __launch_bounds__(kPolarThreads, 1)
polar_quadrant_kernel(const __grid_constant__ CUtensorMap src_map, float* __restrict__ dst, int n_img,
                      const float* __restrict__ lut_fx, const float* __restrict__ lut_fy,
                      const uint32_t* __restrict__ lut_off, PolarQuadBoxes boxes, int h_s, int w_s, int patch_w,
                      const __grid_constant__ PolarNorm norm) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr uint32_t kStageBytes = BOXW * kBoxH * sizeof(T);
  constexpr uint32_t kStageStride = (kStageBytes + 127) & ~127u;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + kPolarStages * kStageStride);

  const int tid = threadIdx.x;
  const int q = blockIdx.x & 3;
  const int worker = blockIdx.x >> 2, n_workers = gridDim.x >> 2;
  const int qw = w_s >> 2;

  // this thread's 16 sample points stay in registers for the whole launch
  float fx[kPolarPx], fy[kPolarPx];
  uint32_t off2[kPolarPx / 2];
#pragma unroll
  for (int i = 0; i < kPolarPx; ++i) {
    const size_t o = ((size_t)q * kPolarPx + i) * kPolarThreads + tid;
    fx[i] = lut_fx[o];
    fy[i] = lut_fy[o];
  }
#pragma unroll
  for (int i = 0; i < kPolarPx / 2; ++i) off2[i] = lut_off[((size_t)q * (kPolarPx / 2) + i) * kPolarThreads + tid];
  // output offset of this thread's pixel in iteration i (same tiling as polar_thread_pixel on the host):
  // the 32 warps cover 32 patches per iteration, which is a whole number of patch rows, so the row advances
  // by a constant kPolarThreads/qw per iteration
  const int patch_h = 32 / patch_w, npx = qw / patch_w, lane = tid & 31, wrp = tid >> 5;
  const size_t out_base = (size_t)((wrp / npx) * patch_h + lane / patch_w) * w_s + (size_t)q * qw + (wrp % npx) * patch_w + lane % patch_w;
  const size_t out_step = (size_t)(kPolarThreads / qw) * w_s;
  const size_t plane_out = (size_t)h_s * w_s;

  if (tid == 0) {
    for (int s = 0; s < kPolarStages; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int bx = q == 0 ? boxes.x0[0] : (q == 1 ? boxes.x0[1] : (q == 2 ? boxes.x0[2] : boxes.x0[3]));
  const int by = q == 0 ? boxes.y0[0] : (q == 1 ? boxes.y0[1] : (q == 2 ? boxes.y0[2] : boxes.y0[3]));
  if (tid == 0) {
    for (int s = 0; s < kPolarStages; ++s) {
      const int it = worker + s * n_workers;
      if (it < n_img) {
        mbar_expect_tx(&full[s], kStageBytes);
        tma_load_3d(smem_raw + s * kStageStride, &src_map, &full[s], bx, by, it);
      }
    }
  }

  int j = 0;
  for (int it = worker; it < n_img; it += n_workers, ++j) {
    const int s = j % kPolarStages;
    mbar_wait(&full[s], (uint32_t)((j / kPolarStages) & 1));
    const T* tile = reinterpret_cast<const T*>(smem_raw + s * kStageStride);
    float* o = dst + (size_t)it * plane_out + out_base;
    float na = 1.f, nb = 0.f;
    if constexpr (sizeof(T) == 1) {
      const int c = it % norm.n_ch;
      na = norm.a[c];
      nb = norm.b[c];
    }
#pragma unroll
    for (int b = 0; b < kPolarPx; b += 4) {  // batches of 4 pixels bound the live registers (1024 threads -> 64 regs each)
      float r[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = b + u;
        const T* p = tile + ((off2[i >> 1] >> ((i & 1) * 16)) & 0xffffu);
        float ia, ib, ic, id;
        if constexpr (sizeof(T) == 1) {
          // uint8 -> fp32 without I2F (a quarter-rate conversion): 0x4B000000 | v is the float 2^23 + v, and subtracting
          // 2^23 is exact
          ia = __uint_as_float(0x4B000000u | p[0]) - 8388608.0f;
          ic = __uint_as_float(0x4B000000u | p[1]) - 8388608.0f;
          ib = __uint_as_float(0x4B000000u | p[BOXW]) - 8388608.0f;
          id = __uint_as_float(0x4B000000u | p[BOXW + 1]) - 8388608.0f;
        } else {
          ia = p[0]; ic = p[1]; ib = p[BOXW]; id = p[BOXW + 1];
        }
        // separable form of the reference's four-weight blend (cvig_fov.py:178-183): lerp along x on both rows, then
        // along y -- 3 subtractions + 3 FMAs and no weight registers; within a few fp32 roundings of the reference's
        // sum (the bit-exact form is witw_bilinear_gather_*)
        const float fxi = fx[i], fyi = fy[i];
        const float top = fmaf(fxi, ic - ia, ia);
        const float bot = fmaf(fxi, id - ib, ib);
        float t = fmaf(fyi, bot - top, top);
        // uint8 source: the blend of the raw pixels, then the channel's normalisation (the four weights sum to one)
        r[u] = sizeof(T) == 1 ? fmaf(na, t, nb) : t;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) __stcs(o + (b + u) * out_step, r[u]);
      asm volatile("" ::: "memory");
    }
    __syncthreads();  // every thread is done with stage s
    if (tid == 0) {
      const int nxt = it + kPolarStages * n_workers;
      if (nxt < n_img) {
        mbar_expect_tx(&full[s], kStageBytes);
        tma_load_3d(smem_raw + s * kStageStride, &src_map, &full[s], bx, by, nxt);
      }
    }
  }
}

// value of a source pixel as the reference's blend sees it: the fp32 pixel, or the exact normalised value of a uint8
// pixel from the per-channel table lut[c][256] (built on the host with the reference's fp32 operations)
__device__ __forceinline__ float polar_src_value(const float* s, int o, const float*, int) { return s[o]; }
__device__ __forceinline__ float polar_src_value(const uint8_t* s, int o, const float* lut, int c) { return __ldg(lut + c * 256 + s[o]); }

template <typename T>
__global__ void polar_exception_kernel(const T* __restrict__ src, float* __restrict__ dst, int64_t n_img,
                                       const PolarException* __restrict__ exc, int n_exc, int s_o, int64_t plane_out,
                                       const float* __restrict__ lut, int n_ch) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_img * n_exc) return;
  const int64_t img = i / n_exc;
  const PolarException e = exc[i % n_exc];
  const T* s = src + img * (int64_t)s_o * s_o;
  const int ch = n_ch > 0 ? (int)(img % n_ch) : 0;
  const float a = polar_src_value(s, e.y0 * s_o + e.x0, lut, ch), b = polar_src_value(s, e.y1 * s_o + e.x0, lut, ch),
              c = polar_src_value(s, e.y0 * s_o + e.x1, lut, ch), d = polar_src_value(s, e.y1 * s_o + e.x1, lut, ch);
  float r = __fadd_rn(__fmul_rn(e.w[0], a), __fmul_rn(e.w[1], b));
  r = __fadd_rn(r, __fmul_rn(e.w[2], c));
  r = __fadd_rn(r, __fmul_rn(e.w[3], d));
  dst[img * plane_out + e.pix] = r;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

void* get_encode_tiled() {
  static void* fn = nullptr;
  if (fn == nullptr) {
    cudaDriverEntryPointQueryResult qres;
    void* f = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = f;
  }
  return fn;
}

}  // namespace witw

using namespace witw;

extern "C" int witw_polar_grid(int h_s, int w_s, int s_o, double* x, double* y) {
  WITW_REQUIRE(h_s > 0 && w_s > 0 && s_o > 0 && x && y, WITW_ERR_INVALID, "witw_polar_grid: bad arguments");
  polar_grid_host(h_s, w_s, s_o, x, y);
  return WITW_OK;
}

extern "C" int witw_bilinear_lut(const double* x, const double* y, int64_t n, int src_h, int src_w, int32_t* idx4,
                                 float* w4) {
  WITW_REQUIRE(x && y && idx4 && w4 && n >= 0 && src_h > 0 && src_w > 0, WITW_ERR_INVALID, "witw_bilinear_lut: bad arguments");
  bilinear_lut_host(x, y, n, src_h, src_w, idx4, w4);
  return WITW_OK;
}

template <typename T>
static int launch_gather(const T* src, float* dst, const int32_t* idx4, const float* w4, int64_t n_img, int src_h, int src_w,
                         int64_t n_out, const float* lut, int n_ch, witw_stream_t stream, const char* who) {
  WITW_REQUIRE(src && dst && idx4 && w4, WITW_ERR_INVALID, "%s: null pointer", who);
  WITW_REQUIRE(n_img >= 0 && n_out >= 0 && src_h > 0 && src_w > 0 && (int64_t)src_h * src_w < (1ll << 31), WITW_ERR_INVALID,
               "%s: bad shape", who);
  if (n_img == 0 || n_out == 0) return WITW_OK;
  const int64_t bx = ceil_div<int64_t>(n_out, 256);
  // enough blocks in y to fill the machine a few times over, each looping over a run of planes
  int64_t by = std::min<int64_t>(n_img, std::max<int64_t>(1, (int64_t)sm_count() * 16 / bx));
  by = std::min<int64_t>(by, 65535);
  const int per = (int)ceil_div<int64_t>(n_img, by);
  by = ceil_div<int64_t>(n_img, per);
  WITW_REQUIRE(bx < (1ll << 31), WITW_ERR_INVALID, "%s: too many sample points", who);
  bilinear_gather_kernel<T><<<dim3((unsigned)bx, (unsigned)by), 256, 0, as_stream(stream)>>>(
      src, dst, reinterpret_cast<const int4*>(idx4), reinterpret_cast<const float4*>(w4), n_img, src_h, src_w, n_out, per, lut, n_ch);
  WITW_LAUNCH_CHECK();
  return WITW_OK;
}

extern "C" int witw_bilinear_gather_f32(const float* src, float* dst, const int32_t* idx4, const float* w4, int64_t n_img,
                                        int src_h, int src_w, int64_t n_out, witw_stream_t stream) {
  return launch_gather<float>(src, dst, idx4, w4, n_img, src_h, src_w, n_out, nullptr, 0, stream, "witw_bilinear_gather_f32");
}

extern "C" int witw_bilinear_gather_u8(const uint8_t* src, float* dst, const int32_t* idx4, const float* w4, int64_t n_planes,
                                       int src_h, int src_w, int64_t n_out, const float* norm_lut, int n_ch, witw_stream_t stream) {
  WITW_REQUIRE(norm_lut && n_ch >= 1, WITW_ERR_INVALID, "witw_bilinear_gather_u8: the normalisation table [n_ch][256] is required");
  return launch_gather<uint8_t>(src, dst, idx4, w4, n_planes, src_h, src_w, n_out, norm_lut, n_ch, stream, "witw_bilinear_gather_u8");
}

extern "C" int witw_norm_lut(const float* divisor, const float* mean, const float* std, int n_ch, float* lut_host) {
  WITW_REQUIRE(divisor && mean && std && lut_host && n_ch >= 1, WITW_ERR_INVALID, "witw_norm_lut: bad arguments");
  for (int c = 0; c < n_ch; ++c)
    for (int v = 0; v < 256; ++v) {
      // cvig_fov.py:147: norm(data / 255.) = ((v / 255) - mean) / std, every step rounded to fp32 as torch does
      volatile float x = (float)v / divisor[c];
      volatile float y = x - mean[c];
      lut_host[c * 256 + v] = y / std[c];
    }
  return WITW_OK;
}

extern "C" size_t witw_polar_plan_bytes(int h_s, int w_s, int s_o) {
  const size_t b = polar_plan_build_host(h_s, w_s, s_o, nullptr);
  if (b == 0) set_error(WITW_ERR_UNSUPPORTED, "witw_polar_plan: geometry %dx%d from %d is outside the staged fast path", h_s, w_s, s_o);
  return b;
}

extern "C" int witw_polar_plan_build(int h_s, int w_s, int s_o, void* plan_host) {
  WITW_REQUIRE(plan_host, WITW_ERR_INVALID, "witw_polar_plan_build: null plan");
  const size_t b = polar_plan_build_host(h_s, w_s, s_o, plan_host);
  WITW_REQUIRE(b != 0, WITW_ERR_UNSUPPORTED, "witw_polar_plan_build: geometry %dx%d from %d is outside the staged fast path", h_s, w_s, s_o);
  return WITW_OK;
}

extern "C" size_t witw_polar_plan_bytes_u8(int h_s, int w_s, int s_o) {
  const size_t b = polar_plan_build_host(h_s, w_s, s_o, nullptr, 1);
  if (b == 0) set_error(WITW_ERR_UNSUPPORTED, "witw_polar_plan_u8: geometry %dx%d from %d is outside the staged fast path", h_s, w_s, s_o);
  return b;
}

extern "C" int witw_polar_plan_build_u8(int h_s, int w_s, int s_o, void* plan_host) {
  WITW_REQUIRE(plan_host, WITW_ERR_INVALID, "witw_polar_plan_build_u8: null plan");
  const size_t b = polar_plan_build_host(h_s, w_s, s_o, plan_host, 1);
  WITW_REQUIRE(b != 0, WITW_ERR_UNSUPPORTED, "witw_polar_plan_build_u8: geometry %dx%d from %d is outside the staged fast path", h_s, w_s, s_o);
  return WITW_OK;
}

template <typename T, int BOXW>
static int launch_polar(const T* src, float* dst, int64_t n_img, const void* plan_host, const void* plan_dev, const PolarNorm& norm,
                        const float* lut_dev, witw_stream_t stream, const char* who) {
  WITW_REQUIRE(src && dst && plan_host && plan_dev, WITW_ERR_INVALID, "%s: null pointer", who);
  PolarPlanHeader h;
  std::memcpy(&h, plan_host, sizeof(h));
  WITW_REQUIRE(h.magic == kPlanMagic, WITW_ERR_INVALID, "%s: plan_host is not a polar plan", who);
  WITW_REQUIRE(h.elem_bytes == (int)sizeof(T) && h.box_w == BOXW, WITW_ERR_INVALID, "%s: the plan was built for %d-byte source pixels", who, h.elem_bytes);
  WITW_REQUIRE(n_img >= 0 && n_img < (1ll << 31), WITW_ERR_INVALID, "%s: bad plane count", who);
  WITW_REQUIRE(((uintptr_t)src & 15) == 0, WITW_ERR_INVALID, "%s: src must be 16-byte aligned", who);
  if (n_img == 0) return WITW_OK;
  int rc = witw_device_check();
  if (rc != WITW_OK) return rc;

  auto encode = reinterpret_cast<EncodeTiledFn>(get_encode_tiled());
  WITW_REQUIRE(encode != nullptr, WITW_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  CUtensorMap map;
  const cuuint64_t dims[3] = {(cuuint64_t)h.s_o, (cuuint64_t)h.s_o, (cuuint64_t)n_img};
  const cuuint64_t strides[2] = {(cuuint64_t)h.s_o * sizeof(T), (cuuint64_t)h.s_o * h.s_o * sizeof(T)};
  const cuuint32_t box[3] = {BOXW, kBoxH, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult cr = encode(&map, sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<T*>(src), dims,
                       strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  WITW_REQUIRE(cr == CUDA_SUCCESS, WITW_ERR_CUDA, "cuTensorMapEncodeTiled(polar source) failed with CUresult %d", (int)cr);

  constexpr uint32_t kStageBytes = BOXW * kBoxH * sizeof(T);
  constexpr uint32_t kStageStride = (kStageBytes + 127) & ~127u;
  const size_t smem = (size_t)kPolarStages * kStageStride + kPolarStages * sizeof(uint64_t);
  WITW_CUDA(cudaFuncSetAttribute(polar_quadrant_kernel<T, BOXW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const char* pd = (const char*)plan_dev;
  const size_t lut_elems = (size_t)4 * kPolarPx * kPolarThreads;
  const float* fx = (const float*)(pd + h.lut_off);
  const float* fy = fx + lut_elems;
  const uint32_t* off = (const uint32_t*)(fy + lut_elems);
  PolarQuadBoxes boxes;
  for (int q = 0; q < 4; ++q) { boxes.x0[q] = h.box_x0[q]; boxes.y0[q] = h.box_y0[q]; }
  int grid = (sm_count() / 4) * 4;
  if ((int64_t)grid > 4 * n_img) grid = (int)(4 * n_img);
  polar_quadrant_kernel<T, BOXW><<<grid, kPolarThreads, smem, as_stream(stream)>>>(map, dst, (int)n_img, fx, fy, off, boxes, h.h_s, h.w_s,
                                                                                  h.patch_w, norm);
  WITW_LAUNCH_CHECK();
  if (h.n_exc > 0) {
    const int64_t n = n_img * h.n_exc;
    polar_exception_kernel<T><<<(unsigned)ceil_div<int64_t>(n, 256), 256, 0, as_stream(stream)>>>(
        src, dst, n_img, (const PolarException*)(pd + h.exc_off), h.n_exc, h.s_o, (int64_t)h.h_s * h.w_s, lut_dev, norm.n_ch);
    WITW_LAUNCH_CHECK();
  }
  return WITW_OK;
}

extern "C" int witw_polar_resample_f32(const float* src, float* dst, int64_t n_img, const void* plan_host,
                                       const void* plan_dev, witw_stream_t stream) {
  PolarNorm norm;
  std::memset(&norm, 0, sizeof(norm));
  return launch_polar<float, kBoxW>(src, dst, n_img, plan_host, plan_dev, norm, nullptr, stream, "witw_polar_resample_f32");
}

extern "C" int witw_polar_resample_u8(const uint8_t* src, float* dst, int64_t n_planes, int n_ch, const float* norm_lut_host,
                                      const float* norm_lut_dev, const void* plan_host, const void* plan_dev, witw_stream_t stream) {
  WITW_REQUIRE(norm_lut_host && norm_lut_dev, WITW_ERR_INVALID, "witw_polar_resample_u8: the normalisation table is required on host and device");
  WITW_REQUIRE(n_ch >= 1 && n_ch <= kMaxNormCh, WITW_ERR_UNSUPPORTED, "witw_polar_resample_u8: 1..%d channels (got %d)", kMaxNormCh, n_ch);
  PolarNorm norm;
  std::memset(&norm, 0, sizeof(norm));
  norm.n_ch = n_ch;
  for (int c = 0; c < n_ch; ++c) {  // the table is affine in the pixel value up to fp32 rounding: value = a * v + b
    norm.b[c] = norm_lut_host[c * 256];
    norm.a[c] = (float)(((double)norm_lut_host[c * 256 + 255] - (double)norm_lut_host[c * 256]) / 255.0);
  }
  return launch_polar<uint8_t, kBoxWU8>(src, dst, n_planes, plan_host, plan_dev, norm, norm_lut_dev, stream, "witw_polar_resample_u8");
}
