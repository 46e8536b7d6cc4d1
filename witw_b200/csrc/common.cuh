// Shared helpers for libwitw_b200: error reporting, launch checks, small PTX wrappers.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/witw_b200.h"

namespace witw {

int set_error(int code, const char* fmt, ...);

#define WITW_REQUIRE(cond, code, ...)                      \
  do {                                                     \
    if (!(cond)) return ::witw::set_error((code), __VA_ARGS__); \
  } while (0)

#define WITW_CUDA(expr)                                                                        \
  do {                                                                                         \
    cudaError_t e__ = (expr);                                                                  \
    if (e__ != cudaSuccess)                                                                    \
      return ::witw::set_error(WITW_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                  \
                               cudaGetErrorString(e__), __FILE__, __LINE__);                   \
  } while (0)

#define WITW_LAUNCH_CHECK() WITW_CUDA(cudaGetLastError())

inline cudaStream_t as_stream(witw_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

int sm_count();  // SMs of the current device (cached per device)

// spectral.cu: exact fp32 distances of explicit pairs from packed azimuth spectra (n_pairs_dev: optional device-side
// count that caps n_pairs, so a list filled on the device needs no host round trip)
int launch_pairs_spec(const float* gal_spec, const float* crop_inv_norm, const float* qry_spec, const float* q_inv_norm,
                      const int64_t* pair_g, const int64_t* pair_q, int64_t n_pairs, const int32_t* n_pairs_dev, int CH,
                      float* dist, int64_t* ori, witw_stream_t stream);

template <typename T>
__host__ __device__ constexpr T ceil_div(T a, T b) { return (a + b - 1) / b; }

}  // namespace witw
