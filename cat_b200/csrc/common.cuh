// Shared device helpers for libcatb200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "catb200.h"

namespace catb {

// ------------------------------------------------------------------------------------------
// error plumbing (host)
// ------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define CATB_REQUIRE(cond, ...)            \
  do {                                     \
    if (!(cond)) {                         \
      ::catb::set_error(__VA_ARGS__);      \
      return CATB_ERR_INVALID;             \
    }                                      \
  } while (0)

// ------------------------------------------------------------------------------------------
// bf16 x 8 vectors (one 16-byte K-unit)
// ------------------------------------------------------------------------------------------
struct f8 {
  float v[8];
};

__device__ __forceinline__ f8 unpack8(const uint4& q) {
  f8 r;
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    r.v[2 * i] = __uint_as_float(w[i] << 16);
    r.v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
  return r;
}

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ uint4 pack8(const f8& r) {
  uint4 q;
  q.x = pack2(r.v[0], r.v[1]);
  q.y = pack2(r.v[2], r.v[3]);
  q.z = pack2(r.v[4], r.v[5]);
  q.w = pack2(r.v[6], r.v[7]);
  return q;
}

__device__ __forceinline__ uint4 ldg16(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ uint4 ld16(const void* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void st16(void* p, const uint4& v) { *reinterpret_cast<uint4*>(p) = v; }

// Activations.  The piecewise-linear ones are max(v, slope * v) with slope 1 (none) / 0 (ReLU) / 0.2 / 0.01: the slope
// is a loop-invariant select that the compiler hoists, leaving two instructions per element.  (A `switch` over the
// activation inside the element loop cost ~45 instructions per element -- tanhf is expanded inline in every arm of
// every unrolled iteration -- and made the GEMM epilogues and the norm kernels instruction bound: ~1 us per 16
// accumulator columns, profiles/r02_epilogue_ablation.txt.)
__device__ __forceinline__ float act_slope(int act) {
  return act == CATB_ACT_NONE ? 1.f : (act == CATB_ACT_LEAKY02 ? 0.2f : (act == CATB_ACT_LEAKY001 ? 0.01f : 0.f));
}
__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == CATB_ACT_TANH) return tanhf(v);
  return fmaxf(v, act_slope(act) * v);
}

// derivative of the activation expressed through its *output*
__device__ __forceinline__ float act_grad_from_out(float out, int act) {
  if (act == CATB_ACT_TANH) return 1.f - out * out;
  return out > 0.f ? 1.f : act_slope(act);
}

// nn.ReflectionPad2d index map; requires -L < i < 2L-1
__device__ __forceinline__ int reflect_idx(int i, int L) {
  if (i < 0) i = -i;
  if (i >= L) i = 2 * (L - 1) - i;
  return i;
}

// Adjoint of reflect_idx: the padded-frame coordinates j (in [-p, L-1+p]) that mirror onto i (at most three)
__device__ __forceinline__ int reflect_sources(int i, int L, int p, int* src) {
  int n = 0;
  src[n++] = i;
  if (i >= 1 && i <= p) src[n++] = -i;
  if (i <= L - 2 && i >= L - 1 - p) src[n++] = 2 * (L - 1) - i;
  return n;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------
// PTX wrappers: mbarrier, bulk async copy (TMA engine), tcgen05 / TMEM
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(addr),
      "r"(parity)
      : "memory");
}
// make generic-proxy shared-memory writes visible to the async proxy (tcgen05.mma / bulk copies)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 16-byte LDGSTS (cp.async) with zero fill: src-size 0 writes 16 zero bytes without touching global memory
__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gmem_src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(sz)
               : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// 1-D bulk copy global -> shared through the TMA engine, completion on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_in_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)),
               "n"(kCols)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc_dyn(uint32_t* slot_in_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)),
               "r"(cols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_dyn(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs, fp32 accumulate; issued by ONE thread
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// The same instruction with the two matrix descriptors passed as (low, high) 32-bit halves.  Inside an issue loop
// only the 14-bit start-address field of the low word changes (tile base + row / K offset, all multiples of 16
// bytes), so a descriptor update is ONE 32-bit add.  Measured on B200 (tools/mma_probe2.cu): a single thread
// sustains one M=128 instruction per 39 / 48 / 64 / 128 cycles at N = 16 / 64 / 128 / 256 when nothing but the
// instruction itself sits in the loop; every extra dependent scalar instruction of the issuing thread costs
// 4-6 cycles, which is what bounded the round-1 kernels (~25 instructions = ~175 cycles per MMA whatever N).
__device__ __forceinline__ void umma_bf16_lh(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                             uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      "setp.ne.b32 p, %6, 0;\n"
      "mov.b64 da, {%1, %2};\n"
      "mov.b64 db, {%3, %4};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Leader election inside a converged warp.  The issue loops are walked by the WHOLE MMA warp (barrier waits, ring
// bookkeeping) and only the tcgen05 instructions sit under this predicate: ptxas then emits uniform-predicated
// UTCHMMA instead of the ELECT / BRA.U.ANY loop it wraps around an MMA issued under `if (lane == 0)`.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .b32 rx;\n"
      ".reg .pred px;\n"
      "elect.sync rx|px, %1;\n"
      "@px mov.s32 %0, 1;\n"
      "}\n"
      : "+r"(pred)
      : "r"(0xffffffffu));
  return pred != 0;
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// 16 consecutive fp32 columns of this warp's 32 TMEM lanes (thread i <-> lane base+i)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// The same load without the wait, so that several can be in flight (tmem_wait_ld() before the registers are read).
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// The wait with the 16 destination registers of an earlier tmem_ld16_nowait as in/out operands: no use of them can be
// scheduled before the wait.
__device__ __forceinline__ void tmem_wait_ld16(uint32_t* r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}

// Epilogue of the 32 accumulator rows a warp owns (TMEM lanes 32w .. 32w+31, thread <-> row), bf16 output, no
// accumulation: bias, activation, then a COALESCED store.  With one thread per output pixel a direct store makes every
// warp instruction touch 32 different pixels (16 bytes each, a pixel pitch apart): 32 partial-sector transactions per
// instruction, measured at 7-25 us per CTA (1/4 - 1/2 of its lifetime, profiles/r02_timeline_*.txt).  Here every
// 64-column slab is staged in this warp's private 4 KB of shared memory (rows of 128 bytes, 16-byte chunks XOR-swizzled
// with the row so that both the row-wise writes and the chunk-wise reads are conflict free) and written out with
// consecutive lanes on consecutive 16-byte chunks of a pixel, i.e. whole 128-byte lines whenever ldy == the slab.
//   trow      TMEM address of the warp's lanes, first column of the tile
//   col_base  global output column of tile column 0 (tile_n * n_tile)
//   rvalid / ypix  of THIS thread's row (ypix = pixel index in Y, < 2^31)
//   stg       4 KB, 128-byte aligned, private to the warp, free of any async-proxy traffic
__device__ __forceinline__ void sts16(uint32_t saddr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds16(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t saddr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(saddr) : "memory");
  return v;
}

__device__ __forceinline__ void epilogue_rows_bf16(uint32_t trow, int n_tile, int col_base, int n_store, int n_rows,
                                                   const float* __restrict__ bias, int act, bool rvalid, uint32_t ypix,
                                                   __nv_bfloat16* __restrict__ y, int ldy, int y_coff, uint8_t* stg,
                                                   uint32_t* pixtab, int lane, int dbg_mode = 0,
                                                   long long* dbg_cyc = nullptr, float* stat_s = nullptr) {
  // stat_s (optional, shared by the CTA's epilogue warps): [2][n_tile] floats, column sums and sums of squares of the
  // STORED (bf16-rounded, post-activation) values of the rows that are written -- the statistics pass of the
  // normalisation layer behind this conv, taken from the staging tile while it is in shared memory anyway.
  // dbg_cyc (development aid, one thread): cycles spent in [0] waiting for TMEM loads, [1] bias / activation / pack /
  // staging stores, [2] the coalesced copy-out.
  // pixtab: 32 words private to the warp: the output pixel of every row (0xffffffff: row not stored), so that the
  // copy-out needs one shared-memory load per chunk instead of two shuffles.
  const uint32_t stg_s = smem_u32(stg), pix_s = smem_u32(pixtab);
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(pix_s + lane * 4), "r"(rvalid ? ypix : 0xffffffffu) : "memory");
  const float slope = act_slope(act);
  const uint32_t row_s = stg_s + lane * 128;
  const int swz = lane & 7;
  uint32_t acc[2][16];
  tmem_ld16_nowait(trow, acc[0]);                              // 16-column pieces, the next one in flight while this
  for (int c64 = 0; c64 < n_tile; c64 += 64) {                 // one is converted
    const int ncol = n_tile - c64 < 64 ? n_tile - c64 : 64;   // multiple of 16
    long long tc0 = dbg_cyc != nullptr ? clock64() : 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (q * 16 < ncol) {
        tmem_wait_ld16(acc[q & 1]);
        if (c64 + q * 16 + 16 < n_tile && !(dbg_mode & 2)) tmem_ld16_nowait(trow + c64 + q * 16 + 16, acc[(q + 1) & 1]);
        if (dbg_cyc != nullptr) {
          const long long t = clock64();
          dbg_cyc[0] += t - tc0;
          tc0 = t;
        }
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int col = col_base + c64 + q * 16 + g * 8;
          f8 o;
#pragma unroll
          for (int e = 0; e < 8; ++e) o.v[e] = __uint_as_float(acc[q & 1][g * 8 + e]);
          if (bias != nullptr) {
#pragma unroll
            for (int e = 0; e < 8; ++e)
              if (col + e < n_rows) o.v[e] += __ldg(bias + col + e);
          }
          if (act == CATB_ACT_TANH) {
#pragma unroll
            for (int e = 0; e < 8; ++e) o.v[e] = tanhf(o.v[e]);
          } else if (act != CATB_ACT_NONE) {
#pragma unroll
            for (int e = 0; e < 8; ++e) o.v[e] = fmaxf(o.v[e], slope * o.v[e]);
          }
          sts16(row_s + (((q * 2 + g) ^ swz) << 4), pack8(o));
        }
        if (dbg_cyc != nullptr) {
          const long long t = clock64();
          dbg_cyc[1] += t - tc0;
          tc0 = t;
        }
      }
    }
    __syncwarp();
    if (stat_s != nullptr && 2 * lane < ncol) {
      // lane <-> the two channels in word `lane` of every staged row (a row's 32 words sit in 32 different banks)
      const uint32_t word_s = stg_s + ((lane & 3) << 2);
      const int chunk = lane >> 2;
      float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll 8
      for (int row = 0; row < 32; ++row) {
        const uint32_t pix = lds32(pix_s + row * 4);
        const uint32_t v = lds32(word_s + row * 128 + ((chunk ^ (row & 7)) << 4));
        if (pix != 0xffffffffu) {
          const float a = __uint_as_float(v << 16), b = __uint_as_float(v & 0xffff0000u);
          s0 += a;
          q0 += a * a;
          s1 += b;
          q1 += b * b;
        }
      }
      atomicAdd(stat_s + c64 + 2 * lane, s0);
      atomicAdd(stat_s + c64 + 2 * lane + 1, s1);
      atomicAdd(stat_s + n_tile + c64 + 2 * lane, q0);
      atomicAdd(stat_s + n_tile + c64 + 2 * lane + 1, q1);
    }
    const int nch = ncol >> 3;                                  // 16-byte chunks per row: 2, 4, 6 or 8
    const uint32_t inv = (65536u + nch - 1) / nch;              // idx / nch == (idx * inv) >> 16 for idx < 256
    const int colb = col_base + c64;
    for (int it = 0; it < nch; ++it) {
      const uint32_t idx = it * 32 + lane;
      const uint32_t row = (idx * inv) >> 16, ch = idx - row * nch;
      const uint4 v = lds16(stg_s + row * 128 + ((ch ^ (row & 7)) << 4));
      const uint32_t pix = lds32(pix_s + row * 4);
      const int col = colb + ch * 8;
      if (pix != 0xffffffffu && col < n_store && !(dbg_mode & 1)) st16(y + static_cast<size_t>(pix) * ldy + y_coff + col, v);
    }
    __syncwarp();
    if (dbg_cyc != nullptr) dbg_cyc[2] += clock64() - tc0;
  }
  tmem_wait_ld();
}

// Named barrier among `count` threads (the epilogue warps of a warp-specialised kernel).
__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// Flush a CTA's [2][n_tile] statistics tile (see epilogue_rows_bf16) into the layer's sums [G][2][C] and clear it.
// Called by `nthreads` threads (tid = 0 .. nthreads-1) after a barrier among them.
__device__ __forceinline__ void flush_epilogue_stats(float* stat_s, int n_tile, int col_base, int n_store,
                                                     const catb_epilogue_stats& st, int n_img, int tid, int nthreads) {
  float* dst = st.sums + static_cast<size_t>(st.per_sample ? n_img : 0) * 2 * st.C + st.coff;
  for (int i = tid; i < 2 * n_tile; i += nthreads) {
    const int which = i >= n_tile ? 1 : 0, c = i - which * n_tile;
    const float v = stat_s[i];
    stat_s[i] = 0.f;
    if (col_base + c < n_store && v != 0.f) atomicAdd(dst + which * st.C + col_base + c, v);
  }
}

// Shared-memory matrix descriptor, 128-byte swizzle (cute::UMMA::SmemDescriptor, version 1).
// Tiles are stored as rows of 128 bytes (64 bf16), 8-row atoms of 1024 bytes, 16-byte units XOR-swizzled
// with (row & 7).  For a K-major operand the rows are M/N indices (SBO = 1024 between 8-row groups,
// LBO unused); for an MN-major operand the rows are K indices (SBO = 1024 between 8-k groups, LBO = byte
// distance between consecutive 64-element MN chunks).
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3ffffu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;  // LayoutType::SWIZZLE_128B
  return d;
}

// The two halves of the same descriptor: hi is constant for a kernel, lo = lbo field | (address >> 4).
__host__ __device__ constexpr uint32_t sw128_desc_hi(uint32_t sbo_bytes) {
  return ((sbo_bytes >> 4) & 0x3fffu) | (1u << 14) | (2u << 29);
}
__device__ __forceinline__ uint32_t sw128_desc_lo(uint32_t smem_addr, uint32_t lbo_bytes) {
  return ((smem_addr & 0x3ffffu) >> 4) | (((lbo_bytes >> 4) & 0x3fffu) << 16);
}

// cute::UMMA::InstrDescriptor for kind::f16, BF16 x BF16 -> F32
__host__ __device__ inline uint32_t make_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;   // c_format = F32
  d |= 1u << 7;   // a_format = BF16
  d |= 1u << 10;  // b_format = BF16
  d |= static_cast<uint32_t>(a_mn_major & 1) << 15;
  d |= static_cast<uint32_t>(b_mn_major & 1) << 16;
  d |= static_cast<uint32_t>(N >> 3) << 17;
  d |= static_cast<uint32_t>(M >> 4) << 24;
  return d;
}

// ------------------------------------------------------------------------------------------
// halo fill shared by the v2 forward and weight-gradient kernels
// ------------------------------------------------------------------------------------------
// One parity plane of a halo chunk: frame rows [m0, m0 + Lh) in pitch space (row pitch Wf), 128 bytes (one <= 64
// channel chunk) per row, 16-byte units XOR-swizzled with the absolute row inside the 1024-byte aligned buffer.
// 8 * STEP threads: thread -> (row residue rsub = tid / 8 of STEP, unit ul = tid % 8); every thread issues its 16-byte
// cp.async copies (zero fill outside the image / frame) back to back.  All geometry arrives in registers and the
// addresses advance incrementally: ~30 instructions per copy instead of the ~85 of the first version, whose loop
// re-read the descriptor from constant memory and rebuilt every address from scratch (the fill of a 25 KB tile took
// 3-6 us, as long as its MMAs: profiles/r02_timeline_*.txt).
// Pins a kernel parameter in a register (otherwise ptxas re-reads it from the constant bank inside the loop).
__device__ __forceinline__ int in_reg(int v) {
  int r;
  asm volatile("mov.b32 %0, %1;" : "=r"(r) : "r"(v));
  return r;
}

template <bool REFLECT, int STEP = 16>
__device__ __forceinline__ void halo_fill_plane(uint32_t plane_smem, int row_phase, const __nv_bfloat16* xc,
                                                const __nv_bfloat16* xsafe, long long img_base, int m0, int rsub, int ul,
                                                int Lh_, int Wf_, int Hf_, int mul_, int y0, int x0, int pa, int pb,
                                                int H_, int W_, int ldx, bool uvalid) {
  const int Lh = in_reg(Lh_), Wf = in_reg(Wf_), Hf = in_reg(Hf_), mul = in_reg(mul_), H = in_reg(H_), W = in_reg(W_);
  const int pitch_bytes = in_reg(ldx * 2);                  // one image stays far below 4 GB: 32-bit byte offsets
  const char* img = reinterpret_cast<const char*>(xc + img_base * ldx);
  int fy = (m0 + rsub) / Wf;
  int fx = (m0 + rsub) - fy * Wf;
  int iy0 = mul * (fy + y0) + pa;                           // advanced together with (fy, fx)
  int ix0 = mul * (fx + x0) + pb;
  const int dix16 = mul * STEP, dixw = mul * Wf;
  uint32_t dst = plane_smem + rsub * 128 + ((ul ^ ((row_phase + rsub) & 7)) << 4);   // hr += STEP (a multiple of 8) keeps the swizzle phase
  static_assert(STEP % 8 == 0, "row step must keep the swizzle phase");
  for (int hr = rsub; hr < Lh; hr += STEP) {
    int iy = iy0, ix = ix0;
    bool ok = uvalid & (fy < Hf);
    if (REFLECT) {
      ok &= (iy > -H) & (iy < 2 * H - 1) & (ix > -W) & (ix < 2 * W - 1);
      iy = reflect_idx(iy, H);
      ix = reflect_idx(ix, W);
    } else {
      ok &= (static_cast<unsigned>(iy) < static_cast<unsigned>(H)) & (static_cast<unsigned>(ix) < static_cast<unsigned>(W));
    }
    const uint32_t off = static_cast<uint32_t>(iy * W + ix) * static_cast<uint32_t>(pitch_bytes);
    const void* src = ok ? static_cast<const void*>(img + off) : static_cast<const void*>(xsafe);
    const int sz = ok ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
    dst += STEP * 128;
    fx += STEP;
    ix0 += dix16;
    while (fx >= Wf) {
      fx -= Wf;
      ix0 -= dixw;
      ++fy;
      iy0 += mul;
    }
  }
}

// Reflection padding on top of a TMA-staged plane: the tensor-map box leaves zeros where the frame reaches outside the
// image (out-of-bounds fill); this pass overwrites exactly those halo rows with the mirrored pixels (nn.ReflectionPad2d),
// 16 bytes per thread and row, after the box has landed.  Same row walk as halo_fill_plane; `row0` = frame pixel index of
// the box origin (fy0 * Wf), `n_rows` = box rows * Wf, plane base 1024-byte aligned (swizzle phase = row & 7).
template <int STEP>
__device__ __forceinline__ void halo_patch_reflect(uint32_t plane_smem, const __nv_bfloat16* xc, const __nv_bfloat16* xsafe,
                                                   long long img_base, int row0, int rsub, int ul, int n_rows_, int Wf_, int mul_,
                                                   int y0, int x0, int pa, int pb, int H_, int W_, int ldx, bool uvalid) {
  const int n_rows = in_reg(n_rows_), Wf = in_reg(Wf_), mul = in_reg(mul_), H = in_reg(H_), W = in_reg(W_);
  const int pitch_bytes = in_reg(ldx * 2);
  const char* img = reinterpret_cast<const char*>(xc + img_base * ldx);
  int fy = (row0 + rsub) / Wf;
  int fx = (row0 + rsub) - fy * Wf;
  int iy0 = mul * (fy + y0) + pa;
  int ix0 = mul * (fx + x0) + pb;
  const int dixs = mul * STEP, dixw = mul * Wf;
  uint32_t dst = plane_smem + rsub * 128 + ((ul ^ (rsub & 7)) << 4);
  static_assert(STEP % 8 == 0, "row step must keep the swizzle phase");
  for (int hr = rsub; hr < n_rows; hr += STEP) {
    int iy = iy0, ix = ix0;
    const bool inside = (static_cast<unsigned>(iy) < static_cast<unsigned>(H)) & (static_cast<unsigned>(ix) < static_cast<unsigned>(W));
    if (!inside) {
      const bool ok = uvalid & (iy > -H) & (iy < 2 * H - 1) & (ix > -W) & (ix < 2 * W - 1);
      iy = reflect_idx(iy, H);
      ix = reflect_idx(ix, W);
      const uint32_t off = static_cast<uint32_t>(iy * W + ix) * static_cast<uint32_t>(pitch_bytes);
      const void* src = ok ? static_cast<const void*>(img + off) : static_cast<const void*>(xsafe);
      const int sz = ok ? 16 : 0;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
    }
    dst += STEP * 128;
    fx += STEP;
    ix0 += dixs;
    while (fx >= Wf) {
      fx -= Wf;
      ix0 -= dixw;
      ++fy;
      iy0 += mul;
    }
  }
}

// ------------------------------------------------------------------------------------------
// the gather shared by every implicit-GEMM kernel (and its SIMT restatement)
// ------------------------------------------------------------------------------------------
// Returns true and the input coordinate when lattice base (bh,bw) = (oh*sn, ow*sn) plus the tap offset
// lands on a real (or mirrored) input pixel.
__device__ __forceinline__ bool gather_coord(int bh, int bw, int dr, int ds, int sd, int pad_mode, int H, int W,
                                             int& ih, int& iw) {
  int h = bh + dr, w = bw + ds;
  if (sd == 2) {
    if ((h | w) & 1) return false;
    h >>= 1;  // arithmetic shift: negative even values stay exact
    w >>= 1;
  }
  if (pad_mode == CATB_PAD_REFLECT) {
    h = reflect_idx(h, H);
    w = reflect_idx(w, W);
  }
  ih = h;
  iw = w;
  return (h >= 0) & (h < H) & (w >= 0) & (w < W);
}

}  // namespace catb
