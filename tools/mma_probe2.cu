#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "common.cuh"
using namespace catb;
__device__ __forceinline__ uint32_t elect_one_sync() {
  uint32_t pred = 0;
  asm volatile("{\n.reg .b32 rx;\n.reg .pred px;\nelect.sync rx|px, %1;\n@px mov.s32 %0, 1;\n}\n" : "+r"(pred) : "r"(0xFFFFFFFF));
  return pred;
}
struct Args { int M, N, n_mma, mode; long long* cycles; };
__global__ void __launch_bounds__(128) k(Args a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  for (int i = threadIdx.x; i < (128 + 256) * 128 / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(base)[i] = make_uint4(0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  fence_proxy_async();
  __syncthreads();
  if (threadIdx.x < 32) { tmem_alloc_dyn(&tmem_slot, 512); tmem_relinquish(); }
  tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = make_idesc_bf16(a.M, a.N, 0, 0);
  const uint32_t a_addr = smem_u32(base), b_addr = smem_u32(base + 128 * 128);
  if (a.mode == 0) {
    if (threadIdx.x == 0) {
      const long long t0 = clock64();
      for (int i = 0; i < a.n_mma; i += 4) {
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4)
          umma_bf16(tmem, make_sw128_desc(a_addr + k4 * 32, 0, 1024), make_sw128_desc(b_addr + k4 * 32, 0, 1024), idesc, 1u);
      }
      umma_commit(&bar);
      mbar_wait(&bar, 0);
      a.cycles[blockIdx.x] = clock64() - t0;
    }
  } else {
    if (threadIdx.x < 32) {   // whole warp runs the loop, one elected lane issues
      const long long t0 = clock64();
      for (int i = 0; i < a.n_mma; i += 4) {
        if (elect_one_sync()) {
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            umma_bf16(tmem, make_sw128_desc(a_addr + k4 * 32, 0, 1024), make_sw128_desc(b_addr + k4 * 32, 0, 1024), idesc, 1u);
        }
        __syncwarp();
      }
      if (elect_one_sync()) umma_commit(&bar);
      __syncwarp();
      mbar_wait(&bar, 0);
      if (threadIdx.x == 0) a.cycles[blockIdx.x] = clock64() - t0;
    }
  }
  tcgen05_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc_dyn(tmem, 512);
}
int main() {
  long long* d; cudaMalloc(&d, 8 * 148);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int mode = 0; mode < 2; ++mode)
    for (int grid : {1, 148})
      for (int N : {16, 64, 128, 256}) {
        Args a{128, N, 4096, mode, d};
        k<<<grid, 128, 100 * 1024>>>(a);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
        std::vector<long long> h(grid);
        cudaMemcpy(h.data(), d, 8 * grid, cudaMemcpyDeviceToHost);
        double s = 0; for (auto v : h) s += v;
        printf("mode %d (%s) grid %3d M 128 N %3d : %.1f cycles/mma\n", mode, mode ? "elect_one" : "lane0", grid, N, s / grid / a.n_mma);
      }
  return 0;
}
