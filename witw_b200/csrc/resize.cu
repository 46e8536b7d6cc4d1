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
//
// Kernel: one CTA per (64 resized columns x tile_rows) output tile, looping over planes with a two-deep pipeline:
//   A  the source rectangle the tile needs in the NEXT plane is copied into one of two shared-memory landing buffers with
//      16-byte cp.async copies of aligned global chunks, while the current plane is computed.  Rows of an image whose row
//      pitch is not a multiple of 16 bytes start at any byte offset inside their first chunk; the offset is a per-row
//      constant that phase B applies (a funnel shift of two adjacent words on the rows that are not word aligned).
//   B  y pass first (it shrinks the rows by the scale factor before the irregular x pass): 8, 16 or 32 lanes share an output
//      row and a lane owns one to three adjacent four-pixel groups (the host picks the split with the fewest instructions per
//      row for the tile width) -- one 32-bit shared load brings four uint8 pixels (a float4 for aligned fp32 rows) per tap
//      row, PRMT + packed FADD2 turn the bytes into floats, packed FFMA2 blends them -- and writes the fp32 intermediate tile
//   C  x pass: a thread owns one output column (its taps live in registers), blends, normalises, stores 128-byte rows.
// ATen runs the x pass first and rounds its intermediate to fp32; running y first moves results by an ulp or two of the
// pixel scale (measured against the oracle in tests/test_gpu_resize.py), well inside the parity tolerance.
#include <cmath>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace witw {

struct ResizePlanHeader {
  uint32_t magic;
  int32_t in_h, in_w, out_h, out_w, antialias;
  int32_t kx, ky;            // taps per output column / row (table stride)
  int32_t cols_max;          // most source columns any 64-column tile touches
  int32_t tile_rows[2];      // output rows per CTA for uint8 / fp32 sources (0: that source type does not fit)
  int32_t span[2];           // most source rows any row tile of that height touches
  int32_t off_sx, off_cx, off_wx, off_sy, off_cy, off_wy;   // byte offsets from the start of the blob
  int32_t total_bytes;
};
static constexpr uint32_t kResizeMagic = 0x325a5352u;  // "RSZ2"
static constexpr int kTileCols = 64;
static constexpr size_t kSmemPreferred = 72 * 1024;     // three CTAs per SM (the register file allows no more)
static constexpr size_t kSmemLimit = 200 * 1024;

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
      float src = std::fmaf(scale, (float)i + 0.5f, -0.5f);   // ATen's builds contract this into one fused multiply-add
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

// shared memory of one CTA: two landing buffers (span rows x pitch bytes, raw source type; two chunks more than the widest
// rectangle: a row starts anywhere inside its first 16-byte chunk, and the y pass reads one word past a lane's last group)
// + the fp32 intermediate tile
static size_t land_pitch_bytes(int cols_max, int esz) { return align16((size_t)cols_max * esz) + 32; }
static size_t mid_pitch_cols(int cols_max) { return (size_t)(cols_max + 3) / 4 * 4; }
static constexpr int kMidPad = 32;   // floats after the intermediate tile: the x pass may read up to its tap bound past a row's end
static size_t tile_smem_bytes(int span, int tile_rows, int cols_max, int esz) {
  return 2 * (size_t)span * land_pitch_bytes(cols_max, esz) + ((size_t)tile_rows * mid_pitch_cols(cols_max) + kMidPad) * sizeof(float);
}

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
  const unsigned char* src;       // 16-byte aligned
  const unsigned char* src_end;   // one past the last byte of the source tensor
  float* dst;
  long long n_planes;
  int n_ch;
  int in_h, in_w, out_h, out_cols, full_w, col_start;
  const int32_t* sx; const int32_t* cx; const float* wx; int kx;
  const int32_t* sy; const int32_t* cy; const float* wy; int ky;
  int tile_rows;
  int pitch_bytes;                // row pitch of a landing buffer (multiple of 16)
  int land_bytes;                 // span x pitch_bytes: size of one landing buffer
  int pitch_cols;                 // row pitch of the intermediate tile, in floats (multiple of 4)
  int lpr_shift;                  // phase B: log2 of the lanes that share an output row (5, 4 or 3)
  int normalize;                  // 0: resize only, 1: (v / divisor - mean) / std with IEEE divisions, 2: v * scale + bias
  float divisor[8], mean[8], stdv[8], scale[8], bias[8];
};

// phase B: one tap row of GPL adjacent four-column groups, accumulated with packed f32x2 operations (two columns per
// instruction).  `row` points at the 4-byte word holding the row's first needed byte and `bs` is that byte's bit offset
// in the word (uint8); for fp32 rows `row` is the first needed float and `bs` its byte offset from 16-byte alignment.
struct Acc4 { float2 lo, hi; };

template <int GPL>
__device__ __forceinline__ void tap_run(Acc4 (&acc)[GPL], const uint8_t*, const unsigned char* row, uint32_t bs, int g0, float2 w2) {
  const uint32_t* wp = reinterpret_cast<const uint32_t*>(row) + g0;
  uint32_t w[GPL + 1];
#pragma unroll
  for (int i = 0; i <= GPL; ++i) w[i] = wp[i];
  const float2 bias = make_float2(-8388608.0f, -8388608.0f);
#pragma unroll
  for (int i = 0; i < GPL; ++i) {
    const uint32_t v = __funnelshift_r(w[i], w[i + 1], bs);      // bs == 0: the word itself
    // byte k of v becomes the mantissa of 2^23 + pixel (PRMT, no conversion pipe); the packed add removes the 2^23
    const float2 p01 = __fadd2_rn(make_float2(__uint_as_float(__byte_perm(v, 0x4B000000u, 0x7540u)),
                                              __uint_as_float(__byte_perm(v, 0x4B000000u, 0x7541u))), bias);
    const float2 p23 = __fadd2_rn(make_float2(__uint_as_float(__byte_perm(v, 0x4B000000u, 0x7542u)),
                                              __uint_as_float(__byte_perm(v, 0x4B000000u, 0x7543u))), bias);
    acc[i].lo = __ffma2_rn(p01, w2, acc[i].lo);
    acc[i].hi = __ffma2_rn(p23, w2, acc[i].hi);
  }
}

template <int GPL>
__device__ __forceinline__ void tap_run(Acc4 (&acc)[GPL], const float*, const unsigned char* row, uint32_t bs, int g0, float2 w2) {
  const float* fp = reinterpret_cast<const float*>(row) + 4 * g0;
#pragma unroll
  for (int i = 0; i < GPL; ++i) {
    float4 v;
    if (bs == 0) v = *reinterpret_cast<const float4*>(fp + 4 * i);
    else v = make_float4(fp[4 * i], fp[4 * i + 1], fp[4 * i + 2], fp[4 * i + 3]);
    acc[i].lo = __ffma2_rn(make_float2(v.x, v.y), w2, acc[i].lo);
    acc[i].hi = __ffma2_rn(make_float2(v.z, v.w), w2, acc[i].hi);
  }
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <typename T, int KX, int GPL>
__global__ void __launch_bounds__(256) resize_norm_kernel(const ResizeArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  float* mid = reinterpret_cast<float*>(smem + 2 * (size_t)a.land_bytes);   // [tile_rows][pitch_cols] fp32 + kMidPad
  constexpr int ESZ = (int)sizeof(T);
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int col = tid & (kTileCols - 1);
  const int lane_row = tid >> 6;                     // 0..3
  // tiles are cut in the resized image's own columns xg; the panorama window maps xg to output column xo
  const int xg = blockIdx.x * kTileCols + col;
  const bool col_ok = xg < a.full_w;
  int xo = xg - a.col_start;
  if (xo < 0) xo += a.full_w;
  const bool mine = col_ok && xo < a.out_cols;
  if (!__syncthreads_or(mine)) return;               // no column of this tile is inside the window
  int sx = 0, cx = 0;
  float wx[KX];                                      // this column's taps; the table holds zeros past its count
#pragma unroll
  for (int j = 0; j < KX; ++j) wx[j] = 0.0f;
  if (col_ok) {
    sx = __ldg(a.sx + xg);
    cx = __ldg(a.cx + xg);
#pragma unroll
    for (int j = 0; j < KX; ++j)
      if (j < a.kx) wx[j] = __ldg(a.wx + (size_t)xg * a.kx + j);
  }
  const int x_first = blockIdx.x * kTileCols;
  const int x_last = min(x_first + kTileCols, a.full_w) - 1;
  const int c_lo = __ldg(a.sx + x_first);
  const int c_hi = __ldg(a.sx + x_last) + __ldg(a.cx + x_last);
  const int n_chunks = (((c_hi - c_lo) * ESZ + 15) >> 4) + 1;      // aligned 16-byte chunks that can hold a row of the rectangle
  const int n_groups = (c_hi - c_lo + 3) >> 2;                     // four-column groups per row
  const int y0 = blockIdx.y * a.tile_rows;
  const int y1 = min(y0 + a.tile_rows, a.out_h);
  const int rows = y1 - y0;
  const int r_lo = __ldg(a.sy + y0);
  const int r_hi = __ldg(a.sy + (y1 - 1)) + __ldg(a.cy + (y1 - 1));
  const int span = r_hi - r_lo;
  const size_t plane_bytes = (size_t)a.in_h * a.in_w * ESZ;
  const uint32_t row_bytes = (uint32_t)a.in_w * ESZ;
  const size_t plane_out = (size_t)a.out_h * a.out_cols;
  const int step_rr = 256 / n_chunks, step_ch = 256 - step_rr * n_chunks;
  const int rr_first = tid / n_chunks, ch_first = tid - rr_first * n_chunks;
  // phase B: 2^lpr_shift lanes share an output row, so a warp works on 32 >> lpr_shift rows at a time
  const int lpr = 1 << a.lpr_shift;
  const int sub = lane >> a.lpr_shift, l = lane & (lpr - 1);
  const int rows_per_warp = 32 >> a.lpr_shift;
  const int pitch_bytes = a.pitch_bytes, pitch_cols = a.pitch_cols;

  // the x pass reads KX taps per column with zero weights past the column's own count: what it reads there must be finite
  for (int i = tid; i < a.tile_rows * pitch_cols + kMidPad; i += 256) mid[i] = 0.0f;

  // A: asynchronous copy of plane p's source rectangle [r_lo, r_hi) x [c_lo, c_hi) into landing buffer `buf`.
  // Row rr lands from the aligned chunk that holds its first byte; chunks past the tensor's end are skipped.
  auto prefetch = [&](long long p, int buf) {
    const unsigned char* rect = a.src + (size_t)p * plane_bytes + ((size_t)r_lo * a.in_w + c_lo) * ESZ;
    unsigned char* land = smem + (size_t)buf * a.land_bytes;
    int rr = rr_first, ch = ch_first;                 // item tid, then steps of 256 without further divisions
    for (; rr < span; ch += step_ch, rr += step_rr) {
      if (ch >= n_chunks) { ch -= n_chunks; ++rr; if (rr >= span) break; }
      const unsigned char* g = rect + (size_t)rr * row_bytes;
      const unsigned char* chunk = g - (reinterpret_cast<uintptr_t>(g) & 15) + 16 * ch;
      unsigned char* dst = land + rr * pitch_bytes + 16 * ch;
      if (chunk + 16 <= a.src_end) {
        cp_async16(dst, chunk);
      } else if (chunk < a.src_end) {                 // the tensor's last, partial chunk: nothing is read past its end
        for (int k = 0; k < (int)(a.src_end - chunk); ++k) dst[k] = __ldg(chunk + k);
      }
    }
    cp_async_commit();
  };

  int buf = 0;
  prefetch(blockIdx.z, 0);                            // gridDim.z <= n_planes
  for (long long p = blockIdx.z; p < a.n_planes; p += gridDim.z, buf ^= 1) {
    const long long p_next = p + gridDim.z;
    if (p_next < a.n_planes) {
      prefetch(p_next, buf ^ 1);                      // that buffer was last read before the previous iteration's second barrier
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();                                  // everybody's copies of plane p have landed; phase C of the previous plane is over
    const unsigned char* land = smem + (size_t)buf * a.land_bytes;
    // low bits of the address of the rectangle's first byte: they give every row's offset inside its first chunk
    const uint32_t rect_lo = (uint32_t)reinterpret_cast<uintptr_t>(a.src + (size_t)p * plane_bytes + ((size_t)r_lo * a.in_w + c_lo) * ESZ);
    // B: y pass.  lpr lanes share an output row; a lane takes GPL adjacent four-column groups per pass
    for (int yy = warp * rows_per_warp + sub; yy < rows; yy += 8 * rows_per_warp) {
      const int y = y0 + yy;
      const int sy = __ldg(a.sy + y), cy = __ldg(a.cy + y);
      const float* wy = a.wy + (size_t)y * a.ky;
      float* mrow = mid + yy * pitch_cols;
      const int rr0 = sy - r_lo;
      for (int g0 = GPL * l; g0 < n_groups; g0 += GPL * lpr) {
        Acc4 acc[GPL];
#pragma unroll
        for (int i = 0; i < GPL; ++i) acc[i].lo = acc[i].hi = make_float2(0.f, 0.f);
        uint32_t g_lo = rect_lo + (uint32_t)rr0 * row_bytes;
        const unsigned char* lrow = land + rr0 * pitch_bytes;
        for (int j = 0; j < cy; ++j, g_lo += row_bytes, lrow += pitch_bytes) {
          const float wj = __ldg(wy + j);
          const uint32_t off = g_lo & 15u;            // where the row starts in its chunk
          tap_run<GPL>(acc, static_cast<const T*>(nullptr), lrow + (off & 12u), ESZ == 1 ? (off & 3u) * 8u : off, g0, make_float2(wj, wj));
        }
#pragma unroll
        for (int i = 0; i < GPL; ++i)
          if (g0 + i < n_groups)
            *reinterpret_cast<float4*>(mrow + 4 * (g0 + i)) = make_float4(acc[i].lo.x, acc[i].lo.y, acc[i].hi.x, acc[i].hi.y);
      }
    }
    __syncthreads();
    // C: x pass, then ImageNormalization
    if (mine) {
      const int c = (int)(p % a.n_ch);
      const float dv = a.divisor[c], mn = a.mean[c], sd = a.stdv[c], sc = a.scale[c], bi = a.bias[c];
      const float* t = mid + (sx - c_lo) + lane_row * pitch_cols;
      float* o = a.dst + (size_t)p * plane_out + (size_t)(y0 + lane_row) * a.out_cols + xo;
      for (int yy = lane_row; yy < rows; yy += 4, t += 4 * pitch_cols, o += 4 * (size_t)a.out_cols) {
        float acc = t[0] * wx[0];
#pragma unroll
        for (int j = 1; j < KX; ++j) {
          if (ESZ == 1) acc = __fmaf_rn(t[j], wx[j], acc);           // zero weight past the column's count, finite operand
          else if (j < cx) acc = __fmaf_rn(t[j], wx[j], acc);        // fp32 sources: what lies past the rectangle may be NaN
        }
        // 1: (v / divisor - mean) / std, each step rounded to fp32 as torch does (the drop-in for ImageNormalization itself);
        // 2: one multiply-add with the same constants folded (behind a resize, whose result is an ulp or two from ATen's anyway)
        if (a.normalize == 1) acc = __fdiv_rn(__fsub_rn(__fdiv_rn(acc, dv), mn), sd);
        else if (a.normalize == 2) acc = __fmaf_rn(acc, sc, bi);
        __stcs(o, acc);
      }
    }
  }
}

// ImageNormalization alone (identity geometry): a streaming kernel, four pixels per thread and iteration.  uint8 pixels go
// through a per-channel table of the 256 possible results held in shared memory (the same three IEEE operations, evaluated
// once per table entry by each CTA); fp32 pixels are divided, shifted and divided in place.  Bit-identical to torch.
template <typename T>
__global__ void __launch_bounds__(256) normalize_kernel(const ResizeArgs a, long long plane_groups, long long n_groups_total) {
  __shared__ float lut[8 * 256];
  constexpr bool kU8 = sizeof(T) == 1;
  if (kU8) {
    for (int i = threadIdx.x; i < a.n_ch * 256; i += 256) {
      const int c = i >> 8;
      lut[i] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)(i & 255), a.divisor[c]), a.mean[c]), a.stdv[c]);
    }
    __syncthreads();
  }
  const long long stride = (long long)gridDim.x * 256;
  for (long long g = (long long)blockIdx.x * 256 + threadIdx.x; g < n_groups_total; g += stride) {
    const int c = (int)((g / plane_groups) % a.n_ch);       // a group of four pixels never straddles two planes
    float4 o;
    if (kU8) {
      const uint32_t v = __ldcs(reinterpret_cast<const uint32_t*>(a.src) + g);
      const float* t = lut + c * 256;
      o = make_float4(t[v & 255u], t[(v >> 8) & 255u], t[(v >> 16) & 255u], t[v >> 24]);
    } else {
      const float4 v = __ldcs(reinterpret_cast<const float4*>(a.src) + g);
      const float dv = a.divisor[c], mn = a.mean[c], sd = a.stdv[c];
      o = make_float4(__fdiv_rn(__fsub_rn(__fdiv_rn(v.x, dv), mn), sd), __fdiv_rn(__fsub_rn(__fdiv_rn(v.y, dv), mn), sd),
                      __fdiv_rn(__fsub_rn(__fdiv_rn(v.z, dv), mn), sd), __fdiv_rn(__fsub_rn(__fdiv_rn(v.w, dv), mn), sd));
    }
    __stcs(reinterpret_cast<float4*>(a.dst) + g, o);
  }
}

template <typename T, int GPL>
static int launch_resize_g(const ResizeArgs& a, dim3 grid, size_t smem, cudaStream_t st) {
#define WITW_RESIZE_LAUNCH(K)                                                                                             \
  do {                                                                                                                    \
    if (smem > 48 * 1024)                                                                                                 \
      WITW_CUDA(cudaFuncSetAttribute(resize_norm_kernel<T, K, GPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    resize_norm_kernel<T, K, GPL><<<grid, 256, smem, st>>>(a);                                                            \
  } while (0)
  // the x pass runs exactly K taps per column (zero weights past a column's own count)
  if (a.kx <= 2) WITW_RESIZE_LAUNCH(2);
  else if (a.kx <= 3) WITW_RESIZE_LAUNCH(3);
  else if (a.kx <= 5) WITW_RESIZE_LAUNCH(5);
  else if (a.kx <= 7) WITW_RESIZE_LAUNCH(7);
  else if (a.kx <= 9) WITW_RESIZE_LAUNCH(9);
  else if (a.kx <= 16) WITW_RESIZE_LAUNCH(16);
  else WITW_RESIZE_LAUNCH(32);
#undef WITW_RESIZE_LAUNCH
  WITW_LAUNCH_CHECK();
  return WITW_OK;
}

template <typename T>
static int launch_resize(const ResizeArgs& a, int gpl, dim3 grid, size_t smem, cudaStream_t st) {
  if (gpl == 1) return launch_resize_g<T, 1>(a, grid, smem, st);
  if (gpl == 2) return launch_resize_g<T, 2>(a, grid, smem, st);
  return launch_resize_g<T, 3>(a, grid, smem, st);
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
  // the kernel takes a tile's source rectangle from its first and last row / column: the tables must be monotone
  for (int i = 1; i < out_w; ++i)
    WITW_REQUIRE(sx[i] >= sx[i - 1] && sx[i] + cx[i] >= sx[i - 1] + cx[i - 1], WITW_ERR_UNSUPPORTED, "witw_resize_plan_build: non-monotone column taps");
  for (int i = 1; i < out_h; ++i)
    WITW_REQUIRE(sy[i] >= sy[i - 1] && sy[i] + cy[i] >= sy[i - 1] + cy[i - 1], WITW_ERR_UNSUPPORTED, "witw_resize_plan_build: non-monotone row taps");
  int cols_max = 0;
  for (int x0 = 0; x0 < out_w; x0 += kTileCols) {
    const int x1 = (x0 + kTileCols < out_w ? x0 + kTileCols : out_w) - 1;
    const int n = sx[x1] + cx[x1] - sx[x0];
    if (n > cols_max) cols_max = n;
  }
  h.cols_max = cols_max;
  // rows per CTA: the largest of 32, 16, ... 1 whose tile fits the preferred shared-memory size, else the limit
  for (int t = 0; t < 2; ++t) {
    const int esz = t == 0 ? 1 : 4;
    int fallback_rows = 0, fallback_span = 0;
    for (int tile_rows = 32; tile_rows >= 1; tile_rows /= 2) {
      int span = 0;
      for (int y0 = 0; y0 < out_h; y0 += tile_rows) {
        const int y1 = (y0 + tile_rows < out_h ? y0 + tile_rows : out_h) - 1;
        const int n = sy[y1] + cy[y1] - sy[y0];
        if (n > span) span = n;
      }
      const size_t need = tile_smem_bytes(span, tile_rows, cols_max, esz);
      if (need <= kSmemPreferred) { h.tile_rows[t] = tile_rows; h.span[t] = span; break; }
      if (need <= kSmemLimit && fallback_rows == 0) { fallback_rows = tile_rows; fallback_span = span; }
    }
    if (h.tile_rows[t] == 0) { h.tile_rows[t] = fallback_rows; h.span[t] = fallback_span; }
  }
  WITW_REQUIRE(h.tile_rows[0] > 0 || h.tile_rows[1] > 0, WITW_ERR_UNSUPPORTED,
               "witw_resize_plan_build: %dx%d -> %dx%d needs more than %zu KB of shared memory per output row tile", in_h, in_w, out_h, out_w,
               kSmemLimit / 1024);
  std::memcpy(base, &h, sizeof(h));
  return WITW_OK;
}

extern "C" int witw_resize_norm(const void* src_dev, int src_is_u8, float* dst_dev, int64_t n_planes, int n_ch, const void* plan_host,
                                const void* plan_dev, int col_start, int col_count, const float* divisor, const float* mean,
                                const float* stdv, witw_stream_t stream) {
  WITW_REQUIRE(plan_host && plan_dev, WITW_ERR_INVALID, "witw_resize_norm: null plan");
  ResizePlanHeader h;
  std::memcpy(&h, plan_host, sizeof(h));
  WITW_REQUIRE(h.magic == kResizeMagic, WITW_ERR_INVALID, "witw_resize_norm: not a resize plan");
  const int t = src_is_u8 ? 0 : 1;
  const int esz = src_is_u8 ? 1 : 4;
  WITW_REQUIRE(h.tile_rows[t] > 0, WITW_ERR_UNSUPPORTED, "witw_resize_norm: %dx%d -> %dx%d does not fit shared memory for %s sources", h.in_h,
               h.in_w, h.out_h, h.out_w, src_is_u8 ? "uint8" : "fp32");
  WITW_REQUIRE(n_planes >= 0 && n_ch >= 1, WITW_ERR_INVALID, "witw_resize_norm: bad plane count");
  WITW_REQUIRE(col_start >= 0 && col_start < h.out_w && col_count >= 1 && col_count <= h.out_w, WITW_ERR_INVALID,
               "witw_resize_norm: column window [%d, +%d) outside the %d resized columns", col_start, col_count, h.out_w);
  const bool normalize = mean != nullptr;
  WITW_REQUIRE(!normalize || (divisor && stdv && n_ch <= 8), WITW_ERR_INVALID,
               "witw_resize_norm: normalisation needs divisor, mean and std for at most 8 channels (got %d)", n_ch);
  if (n_planes == 0) return WITW_OK;
  WITW_REQUIRE(src_dev && dst_dev, WITW_ERR_INVALID, "witw_resize_norm: null image pointer");
  WITW_REQUIRE(reinterpret_cast<uintptr_t>(src_dev) % 16 == 0, WITW_ERR_INVALID, "witw_resize_norm: source images must be 16-byte aligned");
  const char* base = static_cast<const char*>(plan_dev);
  ResizeArgs a;
  std::memset(&a, 0, sizeof(a));
  a.src = static_cast<const unsigned char*>(src_dev);
  a.src_end = a.src + (size_t)n_planes * h.in_h * h.in_w * esz;
  a.dst = dst_dev; a.n_planes = n_planes; a.n_ch = n_ch;
  a.in_h = h.in_h; a.in_w = h.in_w; a.out_h = h.out_h; a.out_cols = col_count; a.full_w = h.out_w; a.col_start = col_start;
  a.sx = reinterpret_cast<const int32_t*>(base + h.off_sx);
  a.cx = reinterpret_cast<const int32_t*>(base + h.off_cx);
  a.wx = reinterpret_cast<const float*>(base + h.off_wx);
  a.kx = h.kx;
  a.sy = reinterpret_cast<const int32_t*>(base + h.off_sy);
  a.cy = reinterpret_cast<const int32_t*>(base + h.off_cy);
  a.wy = reinterpret_cast<const float*>(base + h.off_wy);
  a.ky = h.ky;
  a.tile_rows = h.tile_rows[t];
  a.pitch_bytes = (int)land_pitch_bytes(h.cols_max, esz);
  a.land_bytes = h.span[t] * a.pitch_bytes;
  a.pitch_cols = (int)mid_pitch_cols(h.cols_max);
  // behind an identity geometry this is ImageNormalization itself: keep torch's two divisions bit for bit
  const bool identity = h.in_h == h.out_h && h.in_w == h.out_w;
  a.normalize = !normalize ? 0 : (identity ? 1 : 2);
  for (int c = 0; c < n_ch && normalize; ++c) {
    a.divisor[c] = divisor[c]; a.mean[c] = mean[c]; a.stdv[c] = stdv[c];
    a.scale[c] = (float)(1.0 / ((double)divisor[c] * (double)stdv[c]));
    a.bias[c] = (float)(-(double)mean[c] / (double)stdv[c]);
  }
  // y pass: 2^lpr_shift lanes share an output row and each takes gpl adjacent four-column groups per pass.  Pick the split
  // of a warp with the fewest instructions per output row on the widest tile: one pass of a lane costs about 15 + 13 gpl
  // issue slots per tap (12 per group to convert and blend, gpl + 1 shared loads, the row bookkeeping)
  int gpl = 1;
  {
    const int groups = (h.cols_max + 3) / 4;
    double best = 1e30;
    a.lpr_shift = 5;
    for (int g = 3; g >= 1; --g)
      for (int shift = 5; shift >= 3; --shift) {
        const double cost = (double)ceil_div(groups, g << shift) * (15.0 + 13.0 * g) / (double)(32 >> shift);
        if (cost < best - 1e-9) { best = cost; a.lpr_shift = shift; gpl = g; }
      }
  }
  cudaStream_t st = as_stream(stream);
  // ImageNormalization on its own: no resampling, stream the pixels
  const long long plane_px = (long long)h.in_h * h.in_w;
  if (identity && normalize && col_start == 0 && col_count == h.out_w && plane_px % 4 == 0) {
    const long long plane_groups = plane_px / 4, total = plane_groups * n_planes;
    long long blocks = ceil_div<long long>(total, 256 * 8);             // about eight groups per thread
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    if (src_is_u8) normalize_kernel<uint8_t><<<(unsigned)blocks, 256, 0, st>>>(a, plane_groups, total);
    else normalize_kernel<float><<<(unsigned)blocks, 256, 0, st>>>(a, plane_groups, total);
    WITW_LAUNCH_CHECK();
    return WITW_OK;
  }
  const unsigned gx = (unsigned)ceil_div(h.out_w, kTileCols), gy = (unsigned)ceil_div(h.out_h, a.tile_rows);
  // a few resident waves of CTAs, each looping over planes: per-thread taps and tile geometry are set up once
  int64_t gz = ceil_div<int64_t>((int64_t)sm_count() * 8, (int64_t)gx * gy);
  if (gz > n_planes) gz = n_planes;
  if (gz > 65535) gz = 65535;
  if (gz < 1) gz = 1;
  dim3 grid(gx, gy, (unsigned)gz);
  const size_t smem = tile_smem_bytes(h.span[t], a.tile_rows, h.cols_max, esz);
  return src_is_u8 ? launch_resize<uint8_t>(a, gpl, grid, smem, st) : launch_resize<float>(a, gpl, grid, smem, st);
}
