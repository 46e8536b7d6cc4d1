// The exchange of a gallery-sharded evaluation (witw_b200/sharded.py) over peer memory: every rank writes its part
// straight into the other GPUs' exchange buffers with NVLink stores from its own kernels, signals a flag per peer and waits for
// the peers' flags -- one kernel per exchange instead of an NCCL all-reduce and an all-gather (whose launch latencies, not their
// bytes, are what a 5 ms step pays for: 40 KB of thresholds, 840 KB of counts + top-k per rank).
//
//   thresholds   the owner of gallery item true_idx[q] holds the fp32 distance d_true[q] of the rank rule (cvig_fov.py:552);
//                it stores the value into every rank's thr[q] (each q has exactly one owner: no reduction), then all ranks wait
//                until all owners have signalled, and copy the complete vector out for the sweep
//   results      every rank stores its [Q] rank counts, its [Q,k] top-k candidates and its finish flag into block `rank` of every
//                rank's gather area; after the wait, the counts are summed and the candidate lists merged (witw_topk_merge)
//
// Buffers: one cudaMalloc'ed exchange buffer per rank (witw_peer_alloc), opened by the other ranks through CUDA IPC
// (witw_peer_open); `peers` is a device array of the world's buffer addresses as this rank sees them (own buffer included).
// Data areas exist twice and alternate with the sequence number's parity; flags hold the sequence number and only grow.  A rank
// can run at most one exchange ahead of the slowest rank (it waits for everybody's flag of exchange i before it starts i + 1), so
// the area of parity p is never rewritten while a peer still reads it.
#include <cstring>

#include "common.cuh"

namespace witw {

constexpr int kPeerThreads = 256;
constexpr int kPeerMaxWorld = 16;

struct PeerLayout {
  size_t flags;        // uint32 [2 kinds][kPeerMaxWorld]  (0: thresholds, 1: results)
  size_t ticket;       // uint32 [2]: blocks of the running kernel that have finished their stores
  size_t error;        // uint32 [1]: a wait timed out
  size_t thr[2];       // float [Q]
  size_t cnt[2];       // int32 [world][Q]
  size_t flag[2];      // int32 [world]
  size_t td[2];        // float [world][Q][k]
  size_t ti[2];        // int32 [world][Q][k]
  size_t total;
};

__host__ __device__ inline PeerLayout peer_layout(int64_t Q, int k, int world) {
  PeerLayout L;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  L.flags = take(2 * kPeerMaxWorld * 4);
  L.ticket = take(8);
  L.error = take(4);
  for (int p = 0; p < 2; ++p) {
    L.thr[p] = take((size_t)Q * 4);
    L.cnt[p] = take((size_t)world * Q * 4);
    L.flag[p] = take((size_t)world * 4);
    L.td[p] = take((size_t)world * Q * (k > 0 ? k : 1) * 4);
    L.ti[p] = take((size_t)world * Q * (k > 0 ? k : 1) * 4);
  }
  L.total = off;
  return L;
}

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// The block that finishes last (all stores of this kernel are fenced behind it) signals every peer and waits for every peer's
// signal of the same exchange.  Returns true in that block, after the wait.
__device__ bool peer_signal_and_wait(unsigned char* const* peers, unsigned char* mine, const PeerLayout& L, int kind, int world, int rank,
                                     uint32_t seq) {
  __shared__ bool last;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t* ticket = reinterpret_cast<uint32_t*>(mine + L.ticket) + kind;
    last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    if (last) *ticket = 0;
  }
  __syncthreads();
  if (!last) return false;
  __threadfence_system();
  if (threadIdx.x < world) {
    const int r = threadIdx.x;
    st_release_sys(reinterpret_cast<uint32_t*>(peers[r] + L.flags) + kind * kPeerMaxWorld + rank, seq);
    const uint32_t* f = reinterpret_cast<const uint32_t*>(mine + L.flags) + kind * kPeerMaxWorld + r;
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(f) - seq) < 0) {
      __nanosleep(200);
      if (clock64() - t0 > 40000000000ll) {            // ~20 s: a peer never arrived; fail the evaluation instead of hanging the GPU
        *reinterpret_cast<uint32_t*>(mine + L.error) = 1u;
        break;
      }
    }
  }
  __syncthreads();
  __threadfence_system();
  return true;
}

__global__ void __launch_bounds__(kPeerThreads)
peer_thresholds_kernel(const float* __restrict__ d_local, const int64_t* __restrict__ true_idx, int64_t g_offset, int64_t g_local, int64_t Q,
                       unsigned char* const* __restrict__ peers, int world, int rank, int k, uint32_t seq, float* __restrict__ d_true_out) {
  const PeerLayout L = peer_layout(Q, k, world);
  const int p = seq & 1u;
  unsigned char* mine = peers[rank];
  for (int64_t q = (int64_t)blockIdx.x * kPeerThreads + threadIdx.x; q < Q; q += (int64_t)gridDim.x * kPeerThreads) {
    const int64_t t = true_idx[q];
    if (t >= g_offset && t < g_offset + g_local) {
      const float v = d_local[q];
      for (int r = 0; r < world; ++r) reinterpret_cast<float*>(peers[r] + L.thr[p])[q] = v;
    }
  }
  if (!peer_signal_and_wait(peers, mine, L, 0, world, rank, seq)) return;
  const float* thr = reinterpret_cast<const float*>(mine + L.thr[p]);
  for (int64_t q = threadIdx.x; q < Q; q += kPeerThreads) d_true_out[q] = __ldcg(thr + q);
}

__global__ void __launch_bounds__(kPeerThreads)
peer_results_kernel(const int32_t* __restrict__ counts, const float* __restrict__ td, const int32_t* __restrict__ ti,
                    const int32_t* __restrict__ flagged, int64_t Q, int k, unsigned char* const* __restrict__ peers, int world, int rank,
                    uint32_t seq, int64_t* __restrict__ total_out, int32_t* __restrict__ n_flag_out) {
  const PeerLayout L = peer_layout(Q, k, world);
  const int p = seq & 1u;
  unsigned char* mine = peers[rank];
  const int64_t n_td = Q * k;
  const int64_t stride = (int64_t)gridDim.x * kPeerThreads;
  for (int64_t i = (int64_t)blockIdx.x * kPeerThreads + threadIdx.x; i < Q; i += stride) {
    const int32_t c = counts[i];
    for (int r = 0; r < world; ++r) reinterpret_cast<int32_t*>(peers[r] + L.cnt[p])[(int64_t)rank * Q + i] = c;
  }
  for (int64_t i = (int64_t)blockIdx.x * kPeerThreads + threadIdx.x; i < n_td; i += stride) {
    const float d = td[i];
    const int32_t x = ti[i];
    for (int r = 0; r < world; ++r) {
      reinterpret_cast<float*>(peers[r] + L.td[p])[(int64_t)rank * n_td + i] = d;
      reinterpret_cast<int32_t*>(peers[r] + L.ti[p])[(int64_t)rank * n_td + i] = x;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < world)
    reinterpret_cast<int32_t*>(peers[threadIdx.x] + L.flag[p])[rank] = flagged ? flagged[0] : 0;
  if (!peer_signal_and_wait(peers, mine, L, 1, world, rank, seq)) return;
  // the sum of the counts: the reference's ranks over the whole gallery (this block alone: Q x world additions)
  const int32_t* cnt = reinterpret_cast<const int32_t*>(mine + L.cnt[p]);
  for (int64_t q = threadIdx.x; q < Q; q += kPeerThreads) {
    int64_t s = 0;
    for (int r = 0; r < world; ++r) s += __ldcg(cnt + (int64_t)r * Q + q);
    total_out[q] = s;
  }
  if (threadIdx.x == 0) {
    int32_t f = 0;
    for (int r = 0; r < world; ++r) f += __ldcg(reinterpret_cast<const int32_t*>(mine + L.flag[p]) + r);
    if (*reinterpret_cast<volatile uint32_t*>(mine + L.error)) f |= 1 << 30;
    n_flag_out[0] = f;
  }
}

}  // namespace witw

using namespace witw;

extern "C" size_t witw_peer_exchange_bytes(int64_t Q, int k, int world) {
  if (Q < 0 || k < 0 || world < 1 || world > kPeerMaxWorld) return 0;
  return peer_layout(Q, k, world).total;
}

extern "C" int witw_peer_alloc(size_t bytes, void** buf_dev, void* ipc_handle_64) {
  WITW_REQUIRE(bytes > 0 && buf_dev && ipc_handle_64, WITW_ERR_INVALID, "witw_peer_alloc: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the binding passes 64-byte handles");
  void* p = nullptr;
  WITW_CUDA(cudaMalloc(&p, bytes));
  WITW_CUDA(cudaMemset(p, 0, bytes));
  WITW_CUDA(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    cudaGetLastError();
    return set_error(WITW_ERR_CUDA, "witw_peer_alloc: cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
  }
  std::memcpy(ipc_handle_64, &h, 64);
  *buf_dev = p;
  return WITW_OK;
}

extern "C" int witw_peer_open(const void* ipc_handle_64, void** buf_dev) {
  WITW_REQUIRE(ipc_handle_64 && buf_dev, WITW_ERR_INVALID, "witw_peer_open: bad arguments");
  cudaIpcMemHandle_t h;
  std::memcpy(&h, ipc_handle_64, 64);
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    cudaGetLastError();   // not sticky: the caller falls back to the collectives, later launch checks must not see this
    return set_error(WITW_ERR_CUDA, "witw_peer_open: cudaIpcOpenMemHandle failed: %s (the exporting device must be visible to this process)",
                     cudaGetErrorString(e));
  }
  *buf_dev = p;
  return WITW_OK;
}

extern "C" int witw_peer_close(void* buf_dev) {
  if (buf_dev) WITW_CUDA(cudaIpcCloseMemHandle(buf_dev));
  return WITW_OK;
}

extern "C" int witw_peer_free(void* buf_dev) {
  if (buf_dev) WITW_CUDA(cudaFree(buf_dev));
  return WITW_OK;
}

static int peer_check(const char* fn, int64_t Q, int k, void* const* peers_dev, int world, int rank) {
  WITW_REQUIRE(Q >= 0 && k >= 0 && k <= 128 && world >= 2 && world <= kPeerMaxWorld && rank >= 0 && rank < world && peers_dev, WITW_ERR_INVALID,
               "%s: bad arguments (Q %lld, k %d, world %d, rank %d)", fn, (long long)Q, k, world, rank);
  return WITW_OK;
}

extern "C" int witw_peer_thresholds(const float* d_local, const int64_t* true_idx, int64_t g_offset, int64_t g_local, int64_t Q, int k,
                                    void* const* peers_dev, int world, int rank, uint32_t seq, float* d_true_out, witw_stream_t stream) {
  int rc = peer_check("witw_peer_thresholds", Q, k, peers_dev, world, rank);
  if (rc != WITW_OK) return rc;
  WITW_REQUIRE(d_local && true_idx && d_true_out && seq > 0, WITW_ERR_INVALID, "witw_peer_thresholds: null pointer or zero sequence number");
  const int grid = (int)std::min<int64_t>(std::max<int64_t>(ceil_div<int64_t>(Q, kPeerThreads), 1), 32);
  peer_thresholds_kernel<<<grid, kPeerThreads, 0, as_stream(stream)>>>(d_local, true_idx, g_offset, g_local, Q,
                                                                      reinterpret_cast<unsigned char* const*>(peers_dev), world, rank, k, seq, d_true_out);
  WITW_LAUNCH_CHECK();
  return WITW_OK;
}

extern "C" int witw_peer_results(const int32_t* counts, const float* td, const int32_t* ti, const int32_t* flagged, int64_t Q, int k,
                                 void* const* peers_dev, const void* own_buf_dev, int world, int rank, uint32_t seq, int64_t* total_out,
                                 int32_t* n_flag_out, float* merged_dist, int32_t* merged_idx, witw_stream_t stream) {
  int rc = peer_check("witw_peer_results", Q, k, peers_dev, world, rank);
  if (rc != WITW_OK) return rc;
  WITW_REQUIRE(counts && total_out && n_flag_out && own_buf_dev && seq > 0, WITW_ERR_INVALID, "witw_peer_results: null pointer or zero sequence number");
  WITW_REQUIRE(k == 0 || (td && ti && merged_dist && merged_idx), WITW_ERR_INVALID, "witw_peer_results: top-k buffers missing");
  const int grid = (int)std::min<int64_t>(std::max<int64_t>(ceil_div<int64_t>(Q * std::max(k, 1), kPeerThreads * 4), 1), 64);
  peer_results_kernel<<<grid, kPeerThreads, 0, as_stream(stream)>>>(counts, td, ti, flagged, Q, k, reinterpret_cast<unsigned char* const*>(peers_dev),
                                                                   world, rank, seq, total_out, n_flag_out);
  WITW_LAUNCH_CHECK();
  if (k > 0 && Q > 0) {
    const PeerLayout L = peer_layout(Q, k, world);
    const unsigned char* mine = reinterpret_cast<const unsigned char*>(own_buf_dev);
    return witw_topk_merge(reinterpret_cast<const float*>(mine + L.td[seq & 1u]), reinterpret_cast<const int32_t*>(mine + L.ti[seq & 1u]), world, Q, k,
                           merged_dist, merged_idx, stream);
  }
  return WITW_OK;
}
