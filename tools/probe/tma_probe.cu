// Probe: which TMA tile boxes does cp.async.bulk.tensor.{2d,3d} accept on this part?
// usage: tma_probe <rank 2|3> <box_w> <box_h> <start_x> <start_y>   (fp32 256x256xN tensor)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
__device__ __forceinline__ unsigned s2u(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ CUtensorMap map, float* out, int rank, int bytes, int cx, int cy, int cz) {
  extern __shared__ __align__(128) unsigned char sm[];
  unsigned long long* bar = (unsigned long long*)(sm + ((bytes + 127) & ~127));
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s2u(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s2u(bar)), "r"(bytes) : "memory");
    if (rank == 3)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(s2u(sm)), "l"(&map), "r"(s2u(bar)), "r"(cx), "r"(cy), "r"(cz) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(s2u(sm)), "l"(&map), "r"(s2u(bar)), "r"(cx), "r"(cy) : "memory");
  }
  __syncthreads();
  unsigned done = 0;
  while (!done) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0,1,0,p;}" : "=r"(done) : "r"(s2u(bar)) : "memory");
  for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) out[i] = ((float*)sm)[i];
}
typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
  int rank = atoi(argv[1]), bw = atoi(argv[2]), bh = atoi(argv[3]), sx = atoi(argv[4]), sy = atoi(argv[5]);
  const int S = 256, N = 3;
  std::vector<float> h((size_t)S * S * N);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 100003);
  float *d, *o;
  cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  int bytes = bw * bh * 4;
  cudaMalloc(&o, bytes);
  void* f = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
  CUtensorMap map;
  cuuint64_t dims[3] = {S, S, N}; cuuint64_t str[2] = {S * 4, (cuuint64_t)S * S * 4};
  if (rank == 2) { dims[1] = S * N; }
  cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1}; cuuint32_t es[3] = {1, 1, 1};
  CUresult r = ((Enc)f)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("rank%d box %dx%d: encode failed %d\n", rank, bw, bh, (int)r); return 0; }
  int smem = ((bytes + 127) & ~127) + 16;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<<<1, 256, smem>>>(map, o, rank, bytes, sx, sy, 1);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("rank%d box %dx%d (%d B) at (%d,%d): KERNEL ERROR %s\n", rank, bw, bh, bytes, sx, sy, cudaGetErrorString(e)); return 0; }
  std::vector<float> got(bw * bh); cudaMemcpy(got.data(), o, bytes, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int y = 0; y < bh; ++y) for (int x = 0; x < bw; ++x) {
    int gx = sx + x, gy = sy + y; float want = 0.f;
    if (gx < S && gy < S) want = h[(size_t)(rank == 3 ? 1 : 0) * S * S + (size_t)(rank == 3 ? gy : gy) * S + gx];
    if (rank == 2) want = (gx < S && gy < S * N) ? h[(size_t)gy * S + gx] : 0.f;
    if (got[y * bw + x] != want) ++bad;
  }
  printf("rank%d box %dx%d (%d B) at (%d,%d): ok, %d mismatches\n", rank, bw, bh, bytes, sx, sy, bad);
  return 0;
}
