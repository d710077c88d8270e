// Implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM).
//
// One kernel family covers every dense convolution of the CAT distillation path and its gradients:
//   fprop  : Y[row, n]  = sum_k gather(X)[row, k] * Wp[n, k]        (conv, conv-transpose, dgrad)
//   wgrad  : dW[c, k]  += sum_row Y[row, c] * gather(X)[row, k]      (weight gradients)
// The K dimension is enumerated by a table of 16-byte units (8 channels of one tap), so kernel size,
// stride, reflect/zero padding, fractional stride (conv-transpose / strided dgrad, optionally phase
// decomposed) and multi-branch K-concatenation are all just different tables.
//
// Warp roles (192 threads): warps 0-3 gather the activation tile into 128B-swizzled shared memory and
// later run the epilogue (warp w owns TMEM lanes 32w..32w+31), warp 4 owns TMEM and issues the MMAs,
// warp 5 streams the pre-packed, pre-swizzled weight tiles with 1-D bulk copies (TMA engine).
#include <algorithm>

#include "common.cuh"

namespace catb {

constexpr int kThreads = 192;
constexpr int kTileM = 128;        // lattice rows per CTA (UMMA M)
constexpr int kChunkK = 64;        // bf16 elements per 128-byte smem row
constexpr int kATileBytes = kTileM * 128;
constexpr int kHeaderBytes = 1024;  // barriers + TMEM slot
constexpr int kMaxStages = 6;

struct FpropParams {
  catb_igemm_desc d;
  const catb_gather_unit* units;
  const __nv_bfloat16* x;
  const uint8_t* wpk;
  const float* bias;
  void* y;
  int M_total, n_chunks, stages, tmem_cols, n_store;
  uint32_t idesc;
};

__device__ __forceinline__ uint8_t* align1024(uint8_t* p) {
  return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~static_cast<uintptr_t>(1023));
}

__global__ void __launch_bounds__(kThreads, 2) igemm_fprop_kernel(const FpropParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + kMaxStages;
  uint64_t* accum = full + 2 * kMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(full + 2 * kMaxStages + 1);
  uint8_t* tiles = smem + kHeaderBytes;

  const catb_igemm_desc& d = p.d;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile_m = blockIdx.x, tile_n = blockIdx.y;
  const int b_bytes = d.n_tile * 128;
  const int stage_bytes = kATileBytes + b_bytes;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 128 + 1);  // 128 gather threads + the weight loader's expect_tx arrive
      mbar_init(&empty[s], 1);       // one tcgen05.commit
    }
    mbar_init(accum, 1);
    fence_barrier_init();
  }
  if (warp == 4) {
    tmem_alloc_dyn(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ------------------------------------------------------------------ gather producer
    const int lat = d.OHs * d.OWs;
    int rn[8], rh[8], rw[8];
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int r = warp * 32 + it * 4 + (lane >> 3);
      const int m = tile_m * kTileM + r;
      if (m < p.M_total) {
        const int n = m / lat;
        const int rem = m - n * lat;
        const int i = rem / d.OWs;
        const int j = rem - i * d.OWs;
        rn[it] = n;
        rh[it] = (d.o_ph + i * d.o_step) * d.sn;
        rw[it] = (d.o_pw + j * d.o_step) * d.sn;
      } else {
        rn[it] = -1;
        rh[it] = 0;
        rw[it] = 0;
      }
    }
    const int ul = lane & 7;
    const size_t ldx = static_cast<size_t>(d.ldx);
    for (int c = 0; c < p.n_chunks; ++c) {
      const int s = c % p.stages;
      const uint32_t ph = (c / p.stages) & 1;
      const int u = c * 8 + ul;
      int dr = 0, ds = 0, cu = 0;
      const bool uvalid = u < d.n_units;
      if (uvalid) {
        const catb_gather_unit g = p.units[u];
        dr = g.dr;
        ds = g.ds;
        cu = g.cu;
      }
      uint4 v[8];
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        v[it] = make_uint4(0, 0, 0, 0);
        int ih, iw;
        if (uvalid && rn[it] >= 0 && gather_coord(rh[it], rw[it], dr, ds, d.sd, d.pad_mode, d.H, d.W, ih, iw)) {
          const size_t pix = (static_cast<size_t>(rn[it]) * d.H + ih) * d.W + iw;
          v[it] = ldg16(p.x + pix * ldx + d.x_coff + cu * 8);
        }
      }
      mbar_wait(&empty[s], ph ^ 1);
      uint8_t* a_tile = tiles + static_cast<size_t>(s) * stage_bytes;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int r = warp * 32 + it * 4 + (lane >> 3);
        st16(a_tile + r * 128 + ((ul ^ (r & 7)) << 4), v[it]);
      }
      fence_proxy_async();
      mbar_arrive(&full[s]);
    }

    // ------------------------------------------------------------------ epilogue
    mbar_wait(accum, 0);
    tcgen05_fence_after();
    const int r = warp * 32 + lane;
    const int m = tile_m * kTileM + r;
    const bool rvalid = m < p.M_total;
    size_t ypix = 0;
    if (rvalid) {
      const int n = m / lat;
      const int rem = m - n * lat;
      const int i = rem / d.OWs;
      const int j = rem - i * d.OWs;
      ypix = (static_cast<size_t>(n) * d.OH + (d.o_ph + i * d.o_step)) * d.OW + (d.o_pw + j * d.o_step);
    }
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    if (!d.y_is_f32 && !d.accumulate) {
      // every stage has been consumed (accum barrier): stage 0's activation tile is the staging area (4 KB per warp)
      epilogue_rows_bf16(trow, d.n_tile, tile_n * d.n_tile, p.n_store, d.n_rows, p.bias, d.act, rvalid,
                         static_cast<uint32_t>(ypix), reinterpret_cast<__nv_bfloat16*>(p.y), d.ldy, d.y_coff,
                         tiles + warp * 4096, reinterpret_cast<uint32_t*>(smem + 512) + warp * 32, lane);
    } else
    for (int cc = 0; cc < d.n_tile / 16; ++cc) {
      float acc[16];
      tmem_ld16(trow + cc * 16, acc);
      const int col0 = tile_n * d.n_tile + cc * 16;
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const int col = col0 + g * 8;
        if (!rvalid || col >= p.n_store) continue;
        f8 o;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float t = acc[g * 8 + q];
          if (p.bias != nullptr && col + q < d.n_rows) t += __ldg(p.bias + col + q);
          o.v[q] = t;
        }
        if (d.y_is_f32) {
          float* yp = reinterpret_cast<float*>(p.y) + ypix * d.ldy + d.y_coff + col;
          if (d.accumulate) {
#pragma unroll
            for (int q = 0; q < 8; ++q) o.v[q] += yp[q];
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) o.v[q] = apply_act(o.v[q], d.act);
          *reinterpret_cast<float4*>(yp) = make_float4(o.v[0], o.v[1], o.v[2], o.v[3]);
          *reinterpret_cast<float4*>(yp + 4) = make_float4(o.v[4], o.v[5], o.v[6], o.v[7]);
        } else {
          __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(p.y) + ypix * d.ldy + d.y_coff + col;
          if (d.accumulate) {
            const f8 old = unpack8(ld16(yp));
#pragma unroll
            for (int q = 0; q < 8; ++q) o.v[q] += old.v[q];
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) o.v[q] = apply_act(o.v[q], d.act);
          st16(yp, pack8(o));
        }
      }
    }
  } else if (warp == 4) {
    // ------------------------------------------------------------------ MMA issuer
    // whole warp walks the loop, one elected lane issues; per instruction one add per descriptor (common.cuh)
    const uint32_t hi = sw128_desc_hi(1024);
    const uint32_t lo0 = sw128_desc_lo(smem_u32(tiles), 16);
    const uint32_t stage16 = static_cast<uint32_t>(stage_bytes) >> 4;
    const uint32_t idesc = p.idesc;
    uint32_t s = 0, ph = 0, a_lo = lo0, acc = 0;
    for (int c = 0; c < p.n_chunks; ++c) {
      mbar_wait(&full[s], ph);
      tcgen05_fence_after();
      if (elect_one()) {
        const uint32_t b_lo = a_lo + (kATileBytes >> 4);
        umma_bf16_lh(tmem_base, a_lo, hi, b_lo, hi, idesc, acc);
        umma_bf16_lh(tmem_base, a_lo + 2, hi, b_lo + 2, hi, idesc, 1u);
        umma_bf16_lh(tmem_base, a_lo + 4, hi, b_lo + 4, hi, idesc, 1u);
        umma_bf16_lh(tmem_base, a_lo + 6, hi, b_lo + 6, hi, idesc, 1u);
        umma_commit(&empty[s]);  // frees the stage once these MMAs have read it
      }
      __syncwarp();
      acc = 1u;
      if (++s == static_cast<uint32_t>(p.stages)) {
        s = 0;
        ph ^= 1u;
        a_lo = lo0;
      } else {
        a_lo += stage16;
      }
    }
    if (elect_one()) umma_commit(accum);
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ weight loader (one thread)
    if (lane == 0) {
      const uint8_t* src = p.wpk + static_cast<size_t>(tile_n) * p.n_chunks * b_bytes;
      for (int c = 0; c < p.n_chunks; ++c) {
        const int s = c % p.stages;
        const uint32_t ph = (c / p.stages) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&full[s], b_bytes);
        bulk_g2s(tiles + static_cast<size_t>(s) * stage_bytes + kATileBytes, src + static_cast<size_t>(c) * b_bytes,
                 b_bytes, &full[s]);
      }
    }
    __syncwarp();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc_dyn(tmem_base, p.tmem_cols);
}

// ------------------------------------------------------------------------------------------------
// wgrad: D[c, k] = sum over lattice rows of Y[row, c] * gather(X)[row, k]; both operands MN-major.
// grid = (row splits, 128-channel tiles of Y, 256-k tiles); partial sums are added atomically.
// ------------------------------------------------------------------------------------------------
struct WgradParams {
  catb_igemm_desc d;
  const catb_gather_unit* units;
  const catb_weight_unit* wunits;
  const __nv_bfloat16* x;
  const __nv_bfloat16* y;
  float* grad;
  float* ws;   // two-stage mode: partial tiles [split][n_rows][ws_k] written with plain stores (no atomics), see below
  int ws_k;
  int M_total, steps_total, steps_per_cta, stages, nb_chunks, cy_p;
  uint32_t idesc;
};

constexpr int kWStepRows = 64;             // lattice rows (GEMM K) per pipeline stage
constexpr int kWChunkBytes = kWStepRows * 128;  // one [64 rows][64 channels] block

__global__ void __launch_bounds__(kThreads, 1) igemm_wgrad_kernel(const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + kMaxStages;
  uint64_t* accum = full + 2 * kMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(full + 2 * kMaxStages + 1);
  uint8_t* tiles = smem + kHeaderBytes;

  const catb_igemm_desc& d = p.d;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tile = blockIdx.y, n_tile = blockIdx.z;
  const int step0 = blockIdx.x * p.steps_per_cta;
  const int step1 = min(p.steps_total, step0 + p.steps_per_cta);
  const int nsteps = step1 - step0;
  const int stage_bytes = (2 + p.nb_chunks) * kWChunkBytes;
  const uint32_t tmem_cols = p.nb_chunks * 64 <= 64 ? 64 : (p.nb_chunks * 64 <= 128 ? 128 : 256);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 128);
      mbar_init(&empty[s], 1);
    }
    mbar_init(accum, 1);
    fence_barrier_init();
  }
  if (warp == 4) {
    tmem_alloc_dyn(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    const int lat = d.OHs * d.OWs;
    const int grp = threadIdx.x >> 3;  // 0..15: handles rows grp, grp+16, grp+32, grp+48 of each step
    const int ul = threadIdx.x & 7;
    // table entries of this thread's B units (fixed for the whole kernel)
    int bdr[4], bds[4], bcu[4];
    bool bval[4];
#pragma unroll
    for (int cb = 0; cb < 4; ++cb) {
      const int u = n_tile * 32 + cb * 8 + ul;
      bval[cb] = cb < p.nb_chunks && u < d.n_units;
      bdr[cb] = bds[cb] = bcu[cb] = 0;
      if (bval[cb]) {
        const catb_gather_unit g = p.units[u];
        bdr[cb] = g.dr;
        bds[cb] = g.ds;
        bcu[cb] = g.cu;
      }
    }
    const size_t ldx = static_cast<size_t>(d.ldx), ldy = static_cast<size_t>(d.ldy);
    for (int t = 0; t < nsteps; ++t) {
      const int s = t % p.stages;
      const uint32_t ph = (t / p.stages) & 1;
      uint4 va[2][4], vb[4][4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int m = (step0 + t) * kWStepRows + grp + 16 * k;
        int n = -1, bh = 0, bw = 0;
        size_t ypix = 0;
        if (m < p.M_total) {
          n = m / lat;
          const int rem = m - n * lat;
          const int i = rem / d.OWs;
          const int j = rem - i * d.OWs;
          const int oh = d.o_ph + i * d.o_step, ow = d.o_pw + j * d.o_step;
          bh = oh * d.sn;
          bw = ow * d.sn;
          ypix = (static_cast<size_t>(n) * d.OH + oh) * d.OW + ow;
        }
#pragma unroll
        for (int ca = 0; ca < 2; ++ca) {
          const int c = m_tile * 128 + ca * 64 + ul * 8;
          va[ca][k] = (n >= 0 && c < p.cy_p) ? ldg16(p.y + ypix * ldy + d.y_coff + c) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) {
          vb[cb][k] = make_uint4(0, 0, 0, 0);
          int ih, iw;
          if (bval[cb] && n >= 0 && gather_coord(bh, bw, bdr[cb], bds[cb], d.sd, d.pad_mode, d.H, d.W, ih, iw)) {
            const size_t pix = (static_cast<size_t>(n) * d.H + ih) * d.W + iw;
            vb[cb][k] = ldg16(p.x + pix * ldx + d.x_coff + bcu[cb] * 8);
          }
        }
      }
      mbar_wait(&empty[s], ph ^ 1);
      uint8_t* st = tiles + static_cast<size_t>(s) * stage_bytes;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int row = grp + 16 * k;
        const int off = row * 128 + ((ul ^ (row & 7)) << 4);
#pragma unroll
        for (int ca = 0; ca < 2; ++ca) st16(st + ca * kWChunkBytes + off, va[ca][k]);
#pragma unroll
        for (int cb = 0; cb < 4; ++cb)
          if (cb < p.nb_chunks) st16(st + (2 + cb) * kWChunkBytes + off, vb[cb][k]);
      }
      fence_proxy_async();
      mbar_arrive(&full[s]);
    }

    // ---- epilogue: scatter the accumulator straight into the fp32 gradient arena
    if (nsteps > 0) {
      mbar_wait(accum, 0);
      tcgen05_fence_after();
      const int row = m_tile * 128 + warp * 32 + lane;  // channel of the lattice tensor
      const bool rvalid = row < d.n_rows;
      const uint32_t trow = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
      if (p.ws != nullptr) {
        // two-stage mode: the split's partial tile goes to the workspace in GEMM order (row = channel of the lattice
        // tensor, column = unit * 8 + element) with 64-byte runs per thread; catb_wgrad_unpack sums the splits in a
        // fixed order and adds the result to the arena layout.  No atomics here: every element is written once.
        float* wrow = p.ws + (static_cast<size_t>(blockIdx.x) * d.n_rows + row) * p.ws_k + n_tile * 256;
        for (int cc = 0; cc < p.nb_chunks * 4; ++cc) {
          float acc[16];
          tmem_ld16(trow + cc * 16, acc);
          if (rvalid) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              *reinterpret_cast<float4*>(wrow + cc * 16 + q * 4) = make_float4(acc[q * 4], acc[q * 4 + 1], acc[q * 4 + 2], acc[q * 4 + 3]);
          }
        }
      } else
      for (int cc = 0; cc < p.nb_chunks * 4; ++cc) {
        float acc[16];
        tmem_ld16(trow + cc * 16, acc);
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int u = n_tile * 32 + cc * 2 + g;
          if (!rvalid || u >= d.n_units) continue;
          const catb_weight_unit wu = p.wunits[u];
          float* base = p.grad + wu.w_off + static_cast<long long>(row) * wu.sn_w;
          for (int q = 0; q < wu.nvalid; ++q) atomicAdd(base + q * wu.sc_w, acc[g * 8 + q]);
        }
      }
    }
  } else if (warp == 4) {
    if (nsteps > 0) {
      // both operands MN-major: 16 rows (GEMM K) per instruction = two 8-row swizzle atoms = 2048 bytes (128 units)
      const uint32_t hi = sw128_desc_hi(1024);
      const uint32_t lo0 = sw128_desc_lo(smem_u32(tiles), kWChunkBytes);
      const uint32_t stage16 = static_cast<uint32_t>(stage_bytes) >> 4;
      const uint32_t idesc = p.idesc;
      uint32_t s = 0, ph = 0, a_lo = lo0, acc = 0;
      for (int t = 0; t < nsteps; ++t) {
        mbar_wait(&full[s], ph);
        tcgen05_fence_after();
        if (elect_one()) {
          const uint32_t b_lo = a_lo + (2 * kWChunkBytes >> 4);
          umma_bf16_lh(tmem_base, a_lo, hi, b_lo, hi, idesc, acc);
          umma_bf16_lh(tmem_base, a_lo + 128, hi, b_lo + 128, hi, idesc, 1u);
          umma_bf16_lh(tmem_base, a_lo + 256, hi, b_lo + 256, hi, idesc, 1u);
          umma_bf16_lh(tmem_base, a_lo + 384, hi, b_lo + 384, hi, idesc, 1u);
          umma_commit(&empty[s]);
        }
        __syncwarp();
        acc = 1u;
        if (++s == static_cast<uint32_t>(p.stages)) {
          s = 0;
          ph ^= 1u;
          a_lo = lo0;
        } else {
          a_lo += stage16;
        }
      }
      if (elect_one()) umma_commit(accum);
      __syncwarp();
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc_dyn(tmem_base, tmem_cols);
}

// ------------------------------------------------------------------------------------------------
// weight packing: fp32 arena (reference layout) -> bf16 smem-image tiles [tile_n][chunk][row][128B sw]
// ------------------------------------------------------------------------------------------------
// Writes image rows [row0, row0 + span) of the packed weights: the first `nreal` of them from the arena
// (row r <-> arena row r of this segment's tensor), the rest as zero padding.  A plain conv is one segment
// covering the whole image; an N-concatenation of several convs (the six branch convs of a residual block
// sharing one input) is packed segment by segment, each with its own weight-unit table.
__global__ void pack_weights_kernel(catb_igemm_desc d, const catb_weight_unit* __restrict__ wunits,
                                    const float* __restrict__ arena, uint8_t* __restrict__ packed, int n_chunks,
                                    int row0, int span, int nreal) {
  const long long total = static_cast<long long>(span) * n_chunks * 8;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ul = static_cast<int>(idx & 7);
    long long t = idx >> 3;
    const int r = static_cast<int>(t % span);
    const int chunk = static_cast<int>(t / span);
    const int n = row0 + r;
    const int tile = n / d.n_tile, row = n - tile * d.n_tile;
    const int u = chunk * 8 + ul;
    f8 o;
#pragma unroll
    for (int q = 0; q < 8; ++q) o.v[q] = 0.f;
    if (r < nreal && u < d.n_units) {
      const catb_weight_unit wu = wunits[u];
      const float* base = arena + wu.w_off + static_cast<long long>(r) * wu.sn_w;
      for (int q = 0; q < wu.nvalid; ++q) o.v[q] = base[q * wu.sc_w];
    }
    uint8_t* dst = packed + (static_cast<size_t>(tile) * n_chunks + chunk) * d.n_tile * 128 + row * 128 +
                   ((ul ^ (row & 7)) << 4);
    st16(dst, pack8(o));
  }
}

// Batched form: blockIdx.y selects a job of a device-resident table, so that all GEMM images of a network are
// re-packed by ONE launch after an optimiser step (instead of one launch per GEMM).
__global__ void pack_weights_batch_kernel(const catb_pack_job* __restrict__ jobs, const float* __restrict__ arena) {
  const catb_pack_job j = jobs[blockIdx.y];
  const long long total = static_cast<long long>(j.span) * j.n_chunks * 8;
  const catb_weight_unit* wunits = reinterpret_cast<const catb_weight_unit*>(j.wunits);
  uint8_t* packed = reinterpret_cast<uint8_t*>(j.packed);
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ul = static_cast<int>(idx & 7);
    long long t = idx >> 3;
    const int r = static_cast<int>(t % j.span);
    const int chunk = static_cast<int>(t / j.span);
    const int n = j.row0 + r;
    const int tile = n / j.n_tile, row = n - tile * j.n_tile;
    const int u = chunk * 8 + ul;
    f8 o;
#pragma unroll
    for (int q = 0; q < 8; ++q) o.v[q] = 0.f;
    if (r < j.nreal && u < j.n_units) {
      const catb_weight_unit wu = wunits[u];
      const float* base = arena + wu.w_off + static_cast<long long>(r) * wu.sn_w;
      for (int q = 0; q < wu.nvalid; ++q) o.v[q] = base[q * wu.sc_w];
    }
    uint8_t* dst = packed + (static_cast<size_t>(tile) * j.n_chunks + chunk) * j.n_tile * 128 + row * 128 +
                   ((ul ^ (row & 7)) << 4);
    st16(dst, pack8(o));
  }
}

// ------------------------------------------------------------------------------------------------
// SIMT restatements (tests only)
// ------------------------------------------------------------------------------------------------
__global__ void ref_fprop_kernel(catb_igemm_desc d, const catb_gather_unit* units, const catb_weight_unit* wunits,
                                 const float* arena, const __nv_bfloat16* x, const float* bias, void* y, int M_total,
                                 int n_store) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(M_total) * n_store) return;
  const int m = static_cast<int>(idx / n_store), col = static_cast<int>(idx % n_store);
  const int lat = d.OHs * d.OWs;
  const int n = m / lat, rem = m - n * lat, i = rem / d.OWs, j = rem - i * d.OWs;
  const int oh = d.o_ph + i * d.o_step, ow = d.o_pw + j * d.o_step;
  float acc = 0.f;
  if (col < d.n_rows) {
    for (int u = 0; u < d.n_units; ++u) {
      const catb_gather_unit g = units[u];
      int ih, iw;
      if (!gather_coord(oh * d.sn, ow * d.sn, g.dr, g.ds, d.sd, d.pad_mode, d.H, d.W, ih, iw)) continue;
      const catb_weight_unit wu = wunits[u];
      const __nv_bfloat16* xp = x + ((static_cast<size_t>(n) * d.H + ih) * d.W + iw) * d.ldx + d.x_coff + g.cu * 8;
      for (int q = 0; q < wu.nvalid; ++q) {
        const float w = __bfloat162float(__float2bfloat16(arena[wu.w_off + static_cast<long long>(col) * wu.sn_w + q * wu.sc_w]));
        acc += __bfloat162float(xp[q]) * w;
      }
    }
    if (bias) acc += bias[col];
  }
  const size_t ypix = (static_cast<size_t>(n) * d.OH + oh) * d.OW + ow;
  if (d.y_is_f32) {
    float* yp = reinterpret_cast<float*>(y) + ypix * d.ldy + d.y_coff + col;
    if (d.accumulate) acc += *yp;
    *yp = apply_act(acc, d.act);
  } else {
    __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(y) + ypix * d.ldy + d.y_coff + col;
    if (d.accumulate) acc += __bfloat162float(*yp);
    *yp = __float2bfloat16(apply_act(acc, d.act));
  }
}

__global__ void ref_wgrad_kernel(catb_igemm_desc d, const catb_gather_unit* units, const catb_weight_unit* wunits,
                                 const __nv_bfloat16* x, const __nv_bfloat16* y, float* grad, int M_total) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(d.n_rows) * d.n_units * 8) return;
  const int q = static_cast<int>(idx & 7);
  const int u = static_cast<int>((idx >> 3) % d.n_units);
  const int c = static_cast<int>((idx >> 3) / d.n_units);
  const catb_weight_unit wu = wunits[u];
  if (q >= wu.nvalid) return;
  const catb_gather_unit g = units[u];
  const int lat = d.OHs * d.OWs;
  float acc = 0.f;
  for (int m = 0; m < M_total; ++m) {
    const int n = m / lat, rem = m - n * lat, i = rem / d.OWs, j = rem - i * d.OWs;
    const int oh = d.o_ph + i * d.o_step, ow = d.o_pw + j * d.o_step;
    int ih, iw;
    if (!gather_coord(oh * d.sn, ow * d.sn, g.dr, g.ds, d.sd, d.pad_mode, d.H, d.W, ih, iw)) continue;
    const float xv = __bfloat162float(x[((static_cast<size_t>(n) * d.H + ih) * d.W + iw) * d.ldx + d.x_coff + g.cu * 8 + q]);
    const float yv = __bfloat162float(y[((static_cast<size_t>(n) * d.OH + oh) * d.OW + ow) * d.ldy + d.y_coff + c]);
    acc += xv * yv;
  }
  atomicAdd(grad + wu.w_off + static_cast<long long>(c) * wu.sn_w + q * wu.sc_w, acc);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int validate_desc(const catb_igemm_desc* d) {
  CATB_REQUIRE(d != nullptr, "null descriptor");
  CATB_REQUIRE(d->N > 0 && d->H > 0 && d->W > 0 && d->OH > 0 && d->OW > 0, "bad tensor extents");
  CATB_REQUIRE(d->OHs > 0 && d->OWs > 0 && d->o_step >= 1, "bad lattice");
  CATB_REQUIRE(d->o_ph + (d->OHs - 1) * d->o_step < d->OH && d->o_pw + (d->OWs - 1) * d->o_step < d->OW,
               "lattice exceeds the output tensor");
  CATB_REQUIRE((d->sn == 1 || d->sn == 2) && (d->sd == 1 || d->sd == 2), "sn/sd must be 1 or 2");
  CATB_REQUIRE(d->ldx % 8 == 0 && d->x_coff % 8 == 0 && d->y_coff % 8 == 0, "pitches/offsets must be multiples of 8");
  CATB_REQUIRE(d->n_units > 0 && d->n_rows > 0, "empty GEMM");
  CATB_REQUIRE(static_cast<long long>(d->N) * d->OHs * d->OWs < (1ll << 31), "too many lattice rows");
  CATB_REQUIRE(static_cast<long long>(d->N) * d->OH * d->OW < (1ll << 31), "too many output pixels");
  return CATB_OK;
}

static int pick_stages(int stage_bytes) {
  const int budget = 110 * 1024 - kHeaderBytes - 1024;
  int s = budget / stage_bytes;
  if (s < 2) s = 2;
  if (s > 4) s = 4;
  return s;
}

static uint32_t pow2_cols(int n) {
  uint32_t c = 32;
  while (static_cast<int>(c) < n) c <<= 1;
  return c;
}

}  // namespace catb

using namespace catb;

extern "C" size_t catb_packed_weight_bytes(int n_rows, int n_units, int n_tile) {
  if (n_rows <= 0 || n_units <= 0 || n_tile <= 0) return 0;
  const int n_tiles = (n_rows + n_tile - 1) / n_tile;
  const int n_chunks = (n_units + 7) / 8;
  return static_cast<size_t>(n_tiles) * n_tile * n_chunks * 128;
}

extern "C" int catb_pack_weights_rows(const catb_igemm_desc* d, const catb_weight_unit* wunits, const float* arena,
                                      void* packed, int row0, int span, int nreal, catb_stream_t s) {
  if (int e = validate_desc(d)) return e;
  CATB_REQUIRE(d->n_tile % 16 == 0 && d->n_tile >= 16 && d->n_tile <= 256, "n_tile must be a multiple of 16 in [16,256]");
  const int n_tiles = (d->n_rows + d->n_tile - 1) / d->n_tile;
  CATB_REQUIRE(row0 >= 0 && span > 0 && nreal >= 0 && nreal <= span && row0 + span <= n_tiles * d->n_tile,
               "row segment [%d,+%d) outside the packed image", row0, span);
  const int n_chunks = (d->n_units + 7) / 8;
  const long long total = static_cast<long long>(span) * n_chunks * 8;
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 8));
  pack_weights_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(s)>>>(*d, wunits, arena, static_cast<uint8_t*>(packed),
                                                                         n_chunks, row0, span, nreal);
  return check_launch("pack_weights");
}

extern "C" int catb_pack_weights_batch(const catb_pack_job* jobs, int n_jobs, int blocks_per_job, const float* arena,
                                       catb_stream_t s) {
  CATB_REQUIRE(jobs != nullptr && n_jobs > 0 && n_jobs <= 65535 && blocks_per_job > 0, "bad pack job table (%d jobs)", n_jobs);
  pack_weights_batch_kernel<<<dim3(blocks_per_job, n_jobs), 256, 0, static_cast<cudaStream_t>(s)>>>(jobs, arena);
  return check_launch("pack_weights_batch");
}

extern "C" int catb_pack_weights(const catb_igemm_desc* d, const catb_weight_unit* wunits, const float* arena,
                                 void* packed, catb_stream_t s) {
  if (d == nullptr) {
    set_error("null descriptor");
    return CATB_ERR_INVALID;
  }
  const int n_tiles = (d->n_rows + d->n_tile - 1) / (d->n_tile > 0 ? d->n_tile : 1);
  return catb_pack_weights_rows(d, wunits, arena, packed, 0, n_tiles * d->n_tile, d->n_rows, s);
}

extern "C" int catb_igemm_fprop(const catb_igemm_desc* d, const catb_gather_unit* units, const void* x,
                                const void* packed_w, const float* bias, void* y, catb_stream_t s) {
  if (int e = validate_desc(d)) return e;
  CATB_REQUIRE(d->n_tile % 16 == 0 && d->n_tile >= 16 && d->n_tile <= 256, "n_tile must be a multiple of 16 in [16,256]");
  CATB_REQUIRE(d->ldy % 8 == 0, "ldy must be a multiple of 8");
  FpropParams p;
  p.d = *d;
  p.units = units;
  p.x = static_cast<const __nv_bfloat16*>(x);
  p.wpk = static_cast<const uint8_t*>(packed_w);
  p.bias = bias;
  p.y = y;
  p.M_total = d->N * d->OHs * d->OWs;
  p.n_chunks = (d->n_units + 7) / 8;
  const int stage_bytes = kATileBytes + d->n_tile * 128;
  p.stages = pick_stages(stage_bytes);
  p.tmem_cols = pow2_cols(d->n_tile);
  p.n_store = (d->n_rows + 7) / 8 * 8;
  p.idesc = make_idesc_bf16(kTileM, d->n_tile, 0, 0);
  const int n_tiles = (d->n_rows + d->n_tile - 1) / d->n_tile;
  dim3 grid((p.M_total + kTileM - 1) / kTileM, n_tiles, 1);
  const size_t smem = 1024 + kHeaderBytes + static_cast<size_t>(p.stages) * stage_bytes;
  igemm_fprop_kernel<<<grid, kThreads, smem, static_cast<cudaStream_t>(s)>>>(p);
  return check_launch("igemm_fprop");
}

// Split plan of the v1 weight gradient: row splits, and the column pitch of the two-stage workspace.
static void wgrad_split_plan(const catb_igemm_desc* d, int* splits, int* steps_per_cta, int* ws_k) {
  const int M_total = d->N * d->OHs * d->OWs;
  const int steps_total = (M_total + kWStepRows - 1) / kWStepRows;
  const int m_tiles = (d->n_rows + 127) / 128;
  const int n_tiles = (d->n_units + 31) / 32;
  // aim at ~4 CTAs per SM in total, at least 4 pipeline steps per CTA
  int sp = (148 * 4 + m_tiles * n_tiles - 1) / (m_tiles * n_tiles);
  if (sp > (steps_total + 3) / 4) sp = (steps_total + 3) / 4;
  if (sp < 1) sp = 1;
  *steps_per_cta = (steps_total + sp - 1) / sp;
  *splits = (steps_total + *steps_per_cta - 1) / *steps_per_cta;
  *ws_k = n_tiles * 256;
}

static int launch_wgrad(const catb_igemm_desc* d, const catb_gather_unit* units, const catb_weight_unit* wunits,
                        const void* x, const void* y, float* arena_grad, float* ws, catb_stream_t s) {
  if (int e = validate_desc(d)) return e;
  CATB_REQUIRE(d->ldy % 8 == 0, "ldy must be a multiple of 8");
  WgradParams p;
  p.d = *d;
  p.units = units;
  p.wunits = wunits;
  p.x = static_cast<const __nv_bfloat16*>(x);
  p.y = static_cast<const __nv_bfloat16*>(y);
  p.grad = arena_grad;
  p.ws = ws;
  p.M_total = d->N * d->OHs * d->OWs;
  p.steps_total = (p.M_total + kWStepRows - 1) / kWStepRows;
  const int n_chunks = (d->n_units + 7) / 8;
  p.nb_chunks = n_chunks < 4 ? n_chunks : 4;
  p.cy_p = (d->n_rows + 7) / 8 * 8;
  const int m_tiles = (d->n_rows + 127) / 128;
  const int n_tiles = (d->n_units + 31) / 32;
  int splits;
  wgrad_split_plan(d, &splits, &p.steps_per_cta, &p.ws_k);
  const int stage_bytes = (2 + p.nb_chunks) * kWChunkBytes;
  p.stages = pick_stages(stage_bytes);
  p.idesc = make_idesc_bf16(128, p.nb_chunks * 64, 1, 1);
  dim3 grid(splits, m_tiles, n_tiles);
  const size_t smem = 1024 + kHeaderBytes + static_cast<size_t>(p.stages) * stage_bytes;
  igemm_wgrad_kernel<<<grid, kThreads, smem, static_cast<cudaStream_t>(s)>>>(p);
  return check_launch("igemm_wgrad");
}

extern "C" int catb_igemm_wgrad(const catb_igemm_desc* d, const catb_gather_unit* units, const catb_weight_unit* wunits,
                                const void* x, const void* y, float* arena_grad, catb_stream_t s) {
  return launch_wgrad(d, units, wunits, x, y, arena_grad, nullptr, s);
}

extern "C" int catb_igemm_wgrad_ws_shape(const catb_igemm_desc* d, int* splits, int* ws_k) {
  if (int e = validate_desc(d)) return e;
  int spc;
  wgrad_split_plan(d, splits, &spc, ws_k);
  return CATB_OK;
}

extern "C" int catb_igemm_wgrad_ws(const catb_igemm_desc* d, const catb_gather_unit* units, const void* x, const void* y,
                                   float* ws, catb_stream_t s) {
  CATB_REQUIRE(ws != nullptr, "null workspace");
  return launch_wgrad(d, units, nullptr, x, y, nullptr, ws, s);
}

// Second stage of the two-stage weight gradient: arena_grad[w(row, unit, q)] += sum over splits (in split order) of
// ws[split][row0 + row][unit * 8 + q] for the rows [row0, row0 + n_rows) of the workspace (one call per row segment when
// the GEMM is an N-concatenation of several convolutions: each segment has its own weight-unit table).  One thread per workspace column: coalesced reads, one atomic add per real element (the
// add itself is ordered by the stream, so the result is deterministic).
__global__ void wgrad_unpack_kernel(const float* __restrict__ ws, int n_splits, int ws_rows, int ws_k, int row0, int n_rows,
                                    int n_units, const catb_weight_unit* __restrict__ wunits, float* __restrict__ grad) {
  const long long total = static_cast<long long>(n_rows) * n_units * 8;
  const size_t split_stride = static_cast<size_t>(ws_rows) * ws_k;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int q = static_cast<int>(idx & 7);
    const int u = static_cast<int>((idx >> 3) % n_units);
    const int row = static_cast<int>((idx >> 3) / n_units);
    const catb_weight_unit wu = wunits[u];
    if (q >= wu.nvalid) continue;
    const float* src = ws + static_cast<size_t>(row0 + row) * ws_k + u * 8 + q;
    float acc = 0.f;
    for (int sp = 0; sp < n_splits; ++sp) acc += src[sp * split_stride];
    atomicAdd(grad + wu.w_off + static_cast<long long>(row) * wu.sn_w + q * wu.sc_w, acc);
  }
}

extern "C" int catb_wgrad_unpack(const float* ws, int n_splits, int ws_rows, int ws_k, int row0, int n_rows, int n_units,
                                 const catb_weight_unit* wunits, float* arena_grad, catb_stream_t s) {
  CATB_REQUIRE(ws != nullptr && wunits != nullptr && arena_grad != nullptr, "null pointer");
  CATB_REQUIRE(n_splits > 0 && n_rows > 0 && n_units > 0 && ws_k >= n_units * 8 && row0 >= 0 && row0 + n_rows <= ws_rows,
               "bad workspace shape");
  const long long total = static_cast<long long>(n_rows) * n_units * 8;
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 16));
  wgrad_unpack_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(s)>>>(ws, n_splits, ws_rows, ws_k, row0, n_rows, n_units,
                                                                        wunits, arena_grad);
  return check_launch("wgrad_unpack");
}

// Batched form: blockIdx.y selects a job of a device-resident table, so that the second stage of every weight gradient
// of a backward pass is ONE launch (small jobs run side by side instead of one ~5 us launch each).
__global__ void wgrad_unpack_batch_kernel(const catb_unpack_job* __restrict__ jobs) {
  const catb_unpack_job j = jobs[blockIdx.y];
  const long long total = static_cast<long long>(j.n_rows) * j.n_units * 8;
  const float* ws = reinterpret_cast<const float*>(j.ws);
  const catb_weight_unit* wunits = reinterpret_cast<const catb_weight_unit*>(j.wunits);
  float* grad = reinterpret_cast<float*>(j.grad);
  const size_t split_stride = static_cast<size_t>(j.ws_rows) * j.ws_k;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int q = static_cast<int>(idx & 7);
    const int u = static_cast<int>((idx >> 3) % j.n_units);
    const int row = static_cast<int>((idx >> 3) / j.n_units);
    const catb_weight_unit wu = wunits[u];
    if (q >= wu.nvalid) continue;
    const float* src = ws + static_cast<size_t>(j.row0 + row) * j.ws_k + u * 8 + q;
    float acc = 0.f;
    for (int sp = 0; sp < j.n_splits; ++sp) acc += src[sp * split_stride];
    atomicAdd(grad + wu.w_off + static_cast<long long>(row) * wu.sn_w + q * wu.sc_w, acc);
  }
}

extern "C" int catb_wgrad_unpack_batch(const catb_unpack_job* jobs, int n_jobs, int blocks_per_job, catb_stream_t s) {
  CATB_REQUIRE(jobs != nullptr && n_jobs > 0 && n_jobs <= 65535 && blocks_per_job > 0, "bad unpack job table (%d jobs)", n_jobs);
  wgrad_unpack_batch_kernel<<<dim3(blocks_per_job, n_jobs), 256, 0, static_cast<cudaStream_t>(s)>>>(jobs);
  return check_launch("wgrad_unpack_batch");
}

extern "C" int catb_ref_fprop(const catb_igemm_desc* d, const catb_gather_unit* units, const catb_weight_unit* wunits,
                              const float* arena, const void* x, const float* bias, void* y, catb_stream_t s) {
  if (int e = validate_desc(d)) return e;
  const int M_total = d->N * d->OHs * d->OWs;
  const int n_store = (d->n_rows + 7) / 8 * 8;
  const long long total = static_cast<long long>(M_total) * n_store;
  ref_fprop_kernel<<<static_cast<unsigned>((total + 127) / 128), 128, 0, static_cast<cudaStream_t>(s)>>>(
      *d, units, wunits, arena, static_cast<const __nv_bfloat16*>(x), bias, y, M_total, n_store);
  return check_launch("ref_fprop");
}

extern "C" int catb_ref_wgrad(const catb_igemm_desc* d, const catb_gather_unit* units, const catb_weight_unit* wunits,
                              const void* x, const void* y, float* arena_grad, catb_stream_t s) {
  if (int e = validate_desc(d)) return e;
  const int M_total = d->N * d->OHs * d->OWs;
  const long long total = static_cast<long long>(d->n_rows) * d->n_units * 8;
  ref_wgrad_kernel<<<static_cast<unsigned>((total + 127) / 128), 128, 0, static_cast<cudaStream_t>(s)>>>(
      *d, units, wunits, static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(y), arena_grad,
      M_total);
  return check_launch("ref_wgrad");
}

namespace catb {
int init_igemm_attributes() {
  cudaError_t e = cudaFuncSetAttribute(igemm_fprop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(fprop): %s", cudaGetErrorString(e));
    return CATB_ERR_CUDA;
  }
  e = cudaFuncSetAttribute(igemm_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(wgrad): %s", cudaGetErrorString(e));
    return CATB_ERR_CUDA;
  }
  return CATB_OK;
}
}  // namespace catb
