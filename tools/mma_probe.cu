// tcgen05.mma issue-cost probe (round-2 design input).
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I include -I cat_b200/csrc -o tools/mma_probe tools/mma_probe.cu
//   gpurun -- './tools/mma_probe > gpurun_out/mma_probe.txt'     (the binary is git-ignored but travels with the snapshot)
//
// Section 9 of DESIGN.md measured ~135-180 cycles per M = 128 MMA with both operands in shared memory, independent of N
// up to 64, which makes the thin convolutions of the student / SPADE generators MMA-count bound.  Before writing the next
// kernel this probe measures, per SM and free of any load traffic (operands are static shared-memory / TMEM tiles), the
// sustained cycles per tcgen05.mma.kind::f16 (K = 16) for
//   * M = 128 and M = 64, N = 16 ... 256, A and B from shared memory (K-major, SWIZZLE_128B)   -> "pixels on N" pays off
//     only if an M = 64 x N = 256 instruction costs clearly less than two M = 128 x N = 64 ones;
//   * the same with A read from TENSOR MEMORY (tcgen05.mma [d], [a], b-desc), which removes the A stream from shared memory;
//   * 1, 2 and 4 resident CTAs per SM (do concurrent CTAs share the MMA issue slots or add up?);
//   * one accumulator vs. two alternating accumulators.
// Results are printed as a table: cycles per instruction and the implied dense TFLOP/s per SM x 148.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#include "common.cuh"

using namespace catb;

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      std::fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); \
      std::exit(1);                                                                    \
    }                                                                                  \
  } while (0)

// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

struct ProbeArgs {
  int M, N, a_in_tmem, n_acc, n_mma, k_slabs;   // k_slabs: distinct 16-wide K slices cycled through (<= 4 per 128-byte row)
  int tmem_cols;                                // power of two >= 32; second accumulator / TMEM A operand at tmem_cols / 2
  long long* cycles;                            // [gridDim.x]
};

// One warp issues; the other three only exist so that the TMEM allocation / barrier pattern matches the real kernels.
__global__ void __launch_bounds__(128) mma_probe_kernel(ProbeArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  // A tile: 128 rows x 128 bytes (64 bf16 of K) ; B tile: 256 rows x 128 bytes; both K-major SW128, 1024-byte aligned
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sA = base;
  unsigned char* sB = base + 128 * 128;
  for (int i = threadIdx.x; i < (128 + 256) * 128 / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(base)[i] = make_uint4(0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);   // finite bf16 values
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  fence_proxy_async();
  __syncthreads();
  if (threadIdx.x < 32) {
    tmem_alloc_dyn(&tmem_slot, a.tmem_cols);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_bf16(a.M, a.N, 0, 0);
    const uint32_t a_addr = smem_u32(sA), b_addr = smem_u32(sB);
    // accumulators at columns 0 and tmem_cols / 2; the TMEM A operand (M x 16 bf16 = 8 columns per K slab) takes the
    // place of the second accumulator in the single-accumulator runs
    const uint32_t half = static_cast<uint32_t>(a.tmem_cols / 2);
    const uint32_t a_tmem = tmem + half;
    // warm-up
    for (int i = 0; i < 8; ++i)
      umma_bf16(tmem, make_sw128_desc(a_addr, 0, 1024), make_sw128_desc(b_addr, 0, 1024), idesc, i > 0);
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    tcgen05_fence_after();
    const long long t0 = clock64();
    for (int i = 0; i < a.n_mma; ++i) {
      const int ks = i % a.k_slabs;                 // 32 bytes per K = 16 slice inside the 128-byte swizzled row
      const uint32_t d = tmem + ((a.n_acc == 2 && (i & 1)) ? half : 0u);
      const uint64_t bd = make_sw128_desc(b_addr + ks * 32, 0, 1024);
      if (a.a_in_tmem)
        umma_bf16_ts(d, a_tmem + ks * 8, bd, idesc, i >= a.n_acc);
      else
        umma_bf16(d, make_sw128_desc(a_addr + ks * 32, 0, 1024), bd, idesc, i >= a.n_acc);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 1);
    const long long t1 = clock64();
    a.cycles[blockIdx.x] = t1 - t0;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc_dyn(tmem, a.tmem_cols);
}

static double run(int M, int N, int a_in_tmem, int n_acc, int ctas_per_sm, int n_sm, long long* d_cycles) {
  ProbeArgs a;
  a.M = M;
  a.N = N;
  a.a_in_tmem = a_in_tmem;
  a.n_acc = n_acc;
  a.n_mma = 2048;
  a.k_slabs = 4;
  a.cycles = d_cycles;
  // TMEM columns: accumulator(s) of N fp32 columns (+ 32 for a TMEM A operand), rounded up to a power of two; co-resident
  // CTAs must fit into the 512 columns of the SM together
  int need = (n_acc == 2 || a_in_tmem) ? 2 * (N > 32 ? N : 32) : N;
  a.tmem_cols = 32;
  while (a.tmem_cols < need) a.tmem_cols *= 2;
  if (a.tmem_cols * ctas_per_sm > 512) {
    std::fprintf(stderr, "config does not fit into tensor memory\n");
    std::exit(1);
  }
  const int grid = n_sm * ctas_per_sm;
  // shared memory sized so that exactly ctas_per_sm CTAs fit (227 KB per SM)
  const int smem = (ctas_per_sm == 1 ? 160 : ctas_per_sm == 2 ? 100 : 50) * 1024;
  CK(cudaFuncSetAttribute(mma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  mma_probe_kernel<<<grid, 128, smem>>>(a);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<long long> h(grid);
  CK(cudaMemcpy(h.data(), d_cycles, grid * sizeof(long long), cudaMemcpyDeviceToHost));
  double s = 0;
  for (long long v : h) s += static_cast<double>(v);
  return s / grid / a.n_mma;
}

int main() {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  int clk_khz = 0;
  CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  std::printf("# %s, %d SMs, %d MHz (attribute)\n", p.name, p.multiProcessorCount, clk_khz / 1000);
  long long* d_cycles;
  CK(cudaMalloc(&d_cycles, sizeof(long long) * p.multiProcessorCount * 4));
  std::printf("# cycles per tcgen05.mma.kind::f16 (K=16), 2048 back-to-back instructions per CTA, no operand loads\n");
  std::printf("%4s %4s %7s %5s %5s %10s %14s\n", "M", "N", "A from", "accs", "CTAs", "cyc/mma", "TFLOP/s @148SM");
  const int Ms[] = {128, 64};
  const int Ns[] = {16, 32, 64, 128, 256};
  for (int a_tmem = 0; a_tmem < 2; ++a_tmem)
    for (int M : Ms)
      for (int N : Ns)
        for (int n_acc = 1; n_acc <= 2; ++n_acc) {
          if (n_acc == 2 && (N > 256 || a_tmem)) continue;      // the TMEM A operand occupies the second accumulator's columns
          const double c = run(M, N, a_tmem, n_acc, 1, p.multiProcessorCount, d_cycles);
          const double tf = 2.0 * M * N * 16 / c * (clk_khz * 1e3) * 148 / 1e12;
          std::printf("%4d %4d %7s %5d %5d %10.1f %14.1f\n", M, N, a_tmem ? "tmem" : "smem", n_acc, 1, c, tf);
        }
  std::printf("# co-resident CTAs issuing concurrently (each with its own accumulator): do the instruction costs add up?\n");
  for (int ctas : {2, 4})
    for (int N : {64, 128}) {
      const double c = run(128, N, 0, 1, ctas, p.multiProcessorCount, d_cycles);
      std::printf("%4d %4d %7s %5d %5d %10.1f %14s\n", 128, N, "smem", 1, ctas, c, "-");
    }
  CK(cudaFree(d_cycles));
  return 0;
}
