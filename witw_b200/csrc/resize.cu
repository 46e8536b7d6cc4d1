// Upstream image preparation: the bilinear resize of `Resize` (model/cvig_fov.py:117-134,
// torchvision.transforms.functional.resize on float images, align_corners=False) fused with `ImageNormalization`
// (model/cvig_fov.py:137-149, norm(data / 255.)) and, for panoramas, with the wrap-around column window of
// cvig_fov.py:120-129.  Raw images (uint8 or fp32, any size) in, model-sized normalised fp32 images out; the overhead
// output is what witw_polar_resample_f32 (polar.cu) consumes.
//
// Both variants of the resize are separable tap tables built on the host with ATen's arithmetic (the resize itself
// lives in torch, not under /root/reference):
//   antialias = 0  ATen upsample_bilinear2d -- the reference's pinned torch 1.8.1 / torchvision 0.9.1
//   antialias = 1  ATen _upsample_bilinear2d_aa -- torchvision >= 0.17's default, the reference as it runs today
// Kernel: one CTA per (64-column x tile_rows) output tile of one plane.  Pass 1 resamples the source rows the tile needs
// along x into shared memory (fp32, rounded once, as ATen's intermediate tensor is); pass 2 blends those rows along y,
// normalises and stores 256-byte row segments.  Each thread owns one output column, so its x taps live in registers.
#include <cmath>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace witw {

struct ResizePlanHeader {
  uint32_t magic;
  int32_t in_h, in_w, out_h, out_w, antialias;
  int32_t kx, ky;            // taps per output column / row (table stride)
  int32_t tile_rows;         // output rows per CTA
  int32_t span_max;          // most source rows any row tile touches
  int32_t off_sx, off_cx, off_wx, off_sy, off_cy, off_wy;   // byte offsets from the start of the blob
  int32_t total_bytes;
};
static constexpr uint32_t kResizeMagic = 0x575a5352u;  // "RSZW"
static constexpr int kTileCols = 64;
static constexpr int kMaxSpanRows = 176;                // 176 x 64 floats = 44 KB of static-limit shared memory

// taps of one axis, mirroring the C++ types of ATen (float scalar_t; the 0.5 literals are doubles)
static int axis_kmax(int in_size, int out_size, int antialias) {
  if (!antialias) return 2;
  const float scale = (float)in_size / (float)out_size;
  const float support = (scale >= 1.0f) ? scale : 1.0f;
  return (int)std::ceil(support) * 2 + 1;
}

static void axis_taps(int in_size, int out_size, int antialias, int k, int32_t* start, int32_t* count, float* w) {
  const float scale = (float)in_size / (float)out_size;
  std::memset(w, 0, sizeof(float) * (size_t)out_size * k);
  if (!antialias) {
    for (int i = 0; i < out_size; ++i) {
      if (in_size == out_size) {           // ATen: scale 1 is a plain copy
        start[i] = i; count[i] = 1; w[(size_t)i * k] = 1.0f;
        continue;
      }
      float src = scale * ((float)i + 0.5f) - 0.5f;
      if (src < 0.0f) src = 0.0f;
      int i0 = (int)std::floor(src);
      if (i0 > in_size - 1) i0 = in_size - 1;
      float l1 = src - (float)i0;
      l1 = l1 < 0.0f ? 0.0f : (l1 > 1.0f ? 1.0f : l1);
      const float l0 = 1.0f - l1;
      start[i] = i0;
      if (i0 < in_size - 1) {
        count[i] = 2; w[(size_t)i * k] = l0; w[(size_t)i * k + 1] = l1;
      } else {                             // both taps are the last sample
        count[i] = 1; w[(size_t)i * k] = l0 + l1;
      }
    }
    return;
  }
  const float support = (scale >= 1.0f) ? scale : 1.0f;
  const float invscale = (scale >= 1.0f) ? (float)(1.0 / (double)scale) : 1.0f;
  for (int i = 0; i < out_size; ++i) {
    const float center = (float)((double)scale * ((double)i + 0.5));
    long lo = (long)((double)(center - support) + 0.5);
    if (lo < 0) lo = 0;
    long hi = (long)((double)(center + support) + 0.5);
    if (hi > in_size) hi = in_size;
    const int n = (int)(hi - lo);
    float total = 0.0f;
    float* wi = w + (size_t)i * k;
    for (int j = 0; j < n && j < k; ++j) {
      float x = (float)(((double)((float)(j + lo) - center) + 0.5) * (double)invscale);
      x = std::fabs(x);
      wi[j] = (x < 1.0f) ? 1.0f - x : 0.0f;
      total += wi[j];
    }
    if (total != 0.0f)
      for (int j = 0; j < n && j < k; ++j) wi[j] /= total;
    start[i] = (int32_t)lo;
    count[i] = n < k ? n : k;
  }
}

static size_t align16(size_t v) { return (v + 15) / 16 * 16; }

static bool plan_layout(int in_h, int in_w, int out_h, int out_w, int antialias, ResizePlanHeader* h) {
  if (in_h < 1 || in_w < 1 || out_h < 1 || out_w < 1) return false;
  std::memset(h, 0, sizeof(*h));
  h->magic = kResizeMagic;
  h->in_h = in_h; h->in_w = in_w; h->out_h = out_h; h->out_w = out_w; h->antialias = antialias ? 1 : 0;
  h->kx = axis_kmax(in_w, out_w, antialias);
  h->ky = axis_kmax(in_h, out_h, antialias);
  size_t off = align16(sizeof(ResizePlanHeader));
  h->off_sx = (int32_t)off; off = align16(off + sizeof(int32_t) * out_w);
  h->off_cx = (int32_t)off; off = align16(off + sizeof(int32_t) * out_w);
  h->off_wx = (int32_t)off; off = align16(off + sizeof(float) * (size_t)out_w * h->kx);
  h->off_sy = (int32_t)off; off = align16(off + sizeof(int32_t) * out_h);
  h->off_cy = (int32_t)off; off = align16(off + sizeof(int32_t) * out_h);
  h->off_wy = (int32_t)off; off = align16(off + sizeof(float) * (size_t)out_h * h->ky);
  h->total_bytes = (int32_t)off;
  return true;
}

// ------------------------------------------------------------------------------------------
// device
// ------------------------------------------------------------------------------------------
struct ResizeArgs {
  const void* src;
  float* dst;
  long long n_planes;
  int n_ch;
  int in_h, in_w, out_h, out_cols, full_w, col_start;
  const int32_t* sx; const int32_t* cx; const float* wx; int kx;
  const int32_t* sy; const int32_t* cy; const float* wy; int ky;
  int tile_rows;
  int normalize;
  float divisor[8], mean[8], stdv[8];
};

__device__ __forceinline__ float px(const float* p) { return __ldg(p); }
__device__ __forceinline__ float px(const uint8_t* p) { return (float)__ldg(p); }

template <typename T, int KX>
__global__ void __launch_bounds__(256) resize_norm_kernel(const ResizeArgs a) {
  extern __shared__ float tmp[];                     // [span][kTileCols]
  const int col = threadIdx.x & (kTileCols - 1);
  const int lane_row = threadIdx.x >> 6;             // 0..3
  const int xo = blockIdx.x * kTileCols + col;       // output column
  const bool col_ok = xo < a.out_cols;
  // the column of the full resized image this output column shows (panorama window with wrap-around)
  int xg = xo + a.col_start;
  if (xg >= a.full_w) xg -= a.full_w;
  int sx = 0, cx = 0;
  float wx[KX];
#pragma unroll
  for (int j = 0; j < KX; ++j) wx[j] = 0.0f;
  if (col_ok) {
    sx = __ldg(a.sx + xg);
    cx = __ldg(a.cx + xg);
#pragma unroll
    for (int j = 0; j < KX; ++j)
      if (j < a.kx) wx[j] = __ldg(a.wx + (size_t)xg * a.kx + j);
  }
  const int y0 = blockIdx.y * a.tile_rows;
  const int y1 = min(y0 + a.tile_rows, a.out_h);
  const int r_lo = __ldg(a.sy + y0);
  const int r_hi = __ldg(a.sy + (y1 - 1)) + __ldg(a.cy + (y1 - 1));
  const size_t plane_in = (size_t)a.in_h * a.in_w;
  const size_t plane_out = (size_t)a.out_h * a.out_cols;

  for (long long p = blockIdx.z; p < a.n_planes; p += gridDim.z) {
    const T* src = reinterpret_cast<const T*>(a.src) + (size_t)p * plane_in;
    // pass 1: along x, one fp32 rounding per intermediate sample
    if (col_ok) {
      for (int r = r_lo + lane_row; r < r_hi; r += 4) {
        const T* row = src + (size_t)r * a.in_w + sx;
        float acc = px(row) * wx[0];
#pragma unroll
        for (int j = 1; j < KX; ++j)
          if (j < cx) acc = __fmaf_rn(px(row + j), wx[j], acc);
        tmp[(r - r_lo) * kTileCols + col] = acc;
      }
    }
    __syncthreads();
    // pass 2: along y, then ImageNormalization: (v / 255 - mean) / std, each step rounded to fp32 as torch does
    if (col_ok) {
      const int c = (int)(p % a.n_ch);
      for (int y = y0 + lane_row; y < y1; y += 4) {
        const int sy = __ldg(a.sy + y), cy = __ldg(a.cy + y);
        const float* wy = a.wy + (size_t)y * a.ky;
        const float* t = tmp + (sy - r_lo) * kTileCols + col;
        float acc = t[0] * __ldg(wy);
        for (int j = 1; j < cy; ++j) acc = __fmaf_rn(t[j * kTileCols], __ldg(wy + j), acc);
        if (a.normalize) acc = __fdiv_rn(__fsub_rn(__fdiv_rn(acc, a.divisor[c]), a.mean[c]), a.stdv[c]);
        __stcs(a.dst + (size_t)p * plane_out + (size_t)y * a.out_cols + xo, acc);
      }
    }
    __syncthreads();
  }
}

template <typename T>
static int launch_resize(const ResizeArgs& a, dim3 grid, size_t smem, cudaStream_t st) {
  if (a.kx <= 2) resize_norm_kernel<T, 2><<<grid, 256, smem, st>>>(a);
  else if (a.kx <= 8) resize_norm_kernel<T, 8><<<grid, 256, smem, st>>>(a);
  else if (a.kx <= 16) resize_norm_kernel<T, 16><<<grid, 256, smem, st>>>(a);
  else resize_norm_kernel<T, 32><<<grid, 256, smem, st>>>(a);
  WITW_LAUNCH_CHECK();
  return WITW_OK;
}

}  // namespace witw

using namespace witw;

extern "C" size_t witw_resize_plan_bytes(int in_h, int in_w, int out_h, int out_w, int antialias) {
  ResizePlanHeader h;
  if (!plan_layout(in_h, in_w, out_h, out_w, antialias, &h)) return 0;
  return (size_t)h.total_bytes;
}

extern "C" int witw_resize_plan_build(int in_h, int in_w, int out_h, int out_w, int antialias, void* plan_host) {
  WITW_REQUIRE(plan_host != nullptr, WITW_ERR_INVALID, "witw_resize_plan_build: null plan");
  ResizePlanHeader h;
  WITW_REQUIRE(plan_layout(in_h, in_w, out_h, out_w, antialias, &h), WITW_ERR_INVALID,
               "witw_resize_plan_build: sizes must be positive (%dx%d -> %dx%d)", in_h, in_w, out_h, out_w);
  WITW_REQUIRE(h.kx <= 32, WITW_ERR_UNSUPPORTED, "witw_resize_plan_build: %d taps per output column (width %d -> %d); at most 32", h.kx,
               in_w, out_w);
  char* base = static_cast<char*>(plan_host);
  std::memset(base, 0, (size_t)h.total_bytes);
  int32_t* sx = reinterpret_cast<int32_t*>(base + h.off_sx);
  int32_t* cx = reinterpret_cast<int32_t*>(base + h.off_cx);
  float* wx = reinterpret_cast<float*>(base + h.off_wx);
  int32_t* sy = reinterpret_cast<int32_t*>(base + h.off_sy);
  int32_t* cy = reinterpret_cast<int32_t*>(base + h.off_cy);
  float* wy = reinterpret_cast<float*>(base + h.off_wy);
  axis_taps(in_w, out_w, antialias, h.kx, sx, cx, wx);
  axis_taps(in_h, out_h, antialias, h.ky, sy, cy, wy);
  // rows per CTA: the largest of 32, 16, ... 1 whose source-row span fits the shared-memory tile
  int tile_rows = 32, span = 0;
  for (;; tile_rows /= 2) {
    span = 0;
    for (int y0 = 0; y0 < out_h; y0 += tile_rows) {
      const int y1 = (y0 + tile_rows < out_h ? y0 + tile_rows : out_h) - 1;
      int lo = sy[y0], hi = sy[y1] + cy[y1];
      for (int y = y0; y <= y1; ++y) {     // the tables are monotone; do not rely on it
        if (sy[y] < lo) lo = sy[y];
        if (sy[y] + cy[y] > hi) hi = sy[y] + cy[y];
      }
      WITW_REQUIRE(lo == sy[y0] && hi == sy[y1] + cy[y1], WITW_ERR_UNSUPPORTED, "witw_resize_plan_build: non-monotone row taps");
      if (hi - lo > span) span = hi - lo;
    }
    if (span <= kMaxSpanRows || tile_rows == 1) break;
  }
  WITW_REQUIRE(span <= kMaxSpanRows, WITW_ERR_UNSUPPORTED, "witw_resize_plan_build: one output row needs %d source rows (height %d -> %d); at most %d",
               span, in_h, out_h, kMaxSpanRows);
  h.tile_rows = tile_rows;
  h.span_max = span;
  std::memcpy(base, &h, sizeof(h));
  return WITW_OK;
}

extern "C" int witw_resize_norm(const void* src_dev, int src_is_u8, float* dst_dev, int64_t n_planes, int n_ch, const void* plan_host,
                                const void* plan_dev, int col_start, int col_count, const float* divisor, const float* mean,
                                const float* stdv, witw_stream_t stream) {
  WITW_REQUIRE(plan_host && plan_dev, WITW_ERR_INVALID, "witw_resize_norm: null plan");
  ResizePlanHeader h;
  std::memcpy(&h, plan_host, sizeof(h));
  WITW_REQUIRE(h.magic == kResizeMagic && h.tile_rows > 0, WITW_ERR_INVALID, "witw_resize_norm: not a resize plan");
  WITW_REQUIRE(n_planes >= 0 && n_ch >= 1, WITW_ERR_INVALID, "witw_resize_norm: bad plane count");
  WITW_REQUIRE(col_start >= 0 && col_start < h.out_w && col_count >= 1 && col_count <= h.out_w, WITW_ERR_INVALID,
               "witw_resize_norm: column window [%d, +%d) outside the %d resized columns", col_start, col_count, h.out_w);
  const bool normalize = mean != nullptr;
  WITW_REQUIRE(!normalize || (divisor && stdv && n_ch <= 8), WITW_ERR_INVALID,
               "witw_resize_norm: normalisation needs divisor, mean and std for at most 8 channels (got %d)", n_ch);
  if (n_planes == 0) return WITW_OK;
  WITW_REQUIRE(src_dev && dst_dev, WITW_ERR_INVALID, "witw_resize_norm: null image pointer");
  const char* base = static_cast<const char*>(plan_dev);
  ResizeArgs a;
  std::memset(&a, 0, sizeof(a));
  a.src = src_dev; a.dst = dst_dev; a.n_planes = n_planes; a.n_ch = n_ch;
  a.in_h = h.in_h; a.in_w = h.in_w; a.out_h = h.out_h; a.out_cols = col_count; a.full_w = h.out_w; a.col_start = col_start;
  a.sx = reinterpret_cast<const int32_t*>(base + h.off_sx);
  a.cx = reinterpret_cast<const int32_t*>(base + h.off_cx);
  a.wx = reinterpret_cast<const float*>(base + h.off_wx);
  a.kx = h.kx;
  a.sy = reinterpret_cast<const int32_t*>(base + h.off_sy);
  a.cy = reinterpret_cast<const int32_t*>(base + h.off_cy);
  a.wy = reinterpret_cast<const float*>(base + h.off_wy);
  a.ky = h.ky;
  a.tile_rows = h.tile_rows;
  a.normalize = normalize ? 1 : 0;
  for (int c = 0; c < n_ch && normalize; ++c) { a.divisor[c] = divisor[c]; a.mean[c] = mean[c]; a.stdv[c] = stdv[c]; }
  dim3 grid((unsigned)ceil_div(col_count, kTileCols), (unsigned)ceil_div(h.out_h, h.tile_rows),
            (unsigned)(n_planes < 32768 ? n_planes : 32768));
  const size_t smem = (size_t)h.span_max * kTileCols * sizeof(float);
  cudaStream_t st = as_stream(stream);
  return src_is_u8 ? launch_resize<uint8_t>(a, grid, smem, st) : launch_resize<float>(a, grid, smem, st);
}
