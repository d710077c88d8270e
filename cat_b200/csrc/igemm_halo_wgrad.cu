// Weight gradient v2: shared-memory halo + shifted windows (see igemm_halo.cu for the forward kernel).
//
//   dW[c, (tap, k)] = sum over lattice positions m of  dY[m, c] * X[m + off(tap), k]
//
// v1 (igemm.cu) gathers X once per tap from L2.  Here a CTA owns one 128-channel tile of dY, one <=64-channel
// chunk of X and a group of up to 8 taps; it walks a range of 128-position tiles (pitch space, one image /
// strip at a time).  Per tile it stages dY[128 pos][128 ch] and the X halo of the chunk ONCE (cp.async,
// zero rows for garbage positions), and issues, for every tap of the group, eight tcgen05.mma
// (M = 128 dY channels, N = 64 X channels, K = 16 positions) whose B descriptor starts at the tap's row offset
// inside the halo.  Both operands are MN-major (rows = positions).  Each tap accumulates in its own 64 TMEM
// columns (8 taps = 512 columns); the epilogue adds them to the fp32 gradient arena with atomics.
#include "common.cuh"

namespace catb {

// One CTA per SM (two ~60 KB operand buffers), so the producers' issue rate is what feeds the tensor pipe: 16 producer
// warps (4 per scheduler) instead of 4 -- a lone warp per scheduler issues one dependent instruction every 4-6 cycles.
constexpr int kWProdWarps = 16;
constexpr int kWProdThreads = kWProdWarps * 32;
constexpr int kWRowStep = kWProdThreads / 8;   // halo / dY rows advanced per producer iteration
constexpr int kWThreads = kWProdThreads + 32;  // + the MMA-issuing warp
constexpr int kWHeader = 1024;
constexpr int kWPos = 128;          // lattice positions (GEMM K) per tile
constexpr int kDyBytes = 2 * kWPos * 128;   // [2 chunks of 64 channels][128 positions][128 B]

struct WHaloParams {
  catb_igemm_desc d;
  catb_halo_desc h;
  const catb_halo_step* steps;
  const catb_halo_chunk* chunks;
  const catb_halo_wgroup* groups;
  const catb_weight_unit* wunits;   // 8 per step
  const __nv_bfloat16* x;
  const __nv_bfloat16* y;
  float* grad;
  float* ws;   // two-stage mode (see igemm.cu): partial tiles [split][n_rows][n_steps * 64]
  int tiles_per_strip, tiles_total, tiles_per_cta, bufs, halo_bytes, cy_p;
};

__device__ __forceinline__ void cp16_zfill(void* smem_dst, const void* gmem_src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(sz)
               : "memory");
}

__global__ void __launch_bounds__(kWThreads, 1) igemm_halo_wgrad_kernel(const WHaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);  // [2]
  uint64_t* empty = full + 2;                            // [2]
  uint64_t* accum = full + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(full + 5);
  uint8_t* bufs = smem + kWHeader;
  const int buf_bytes = kDyBytes + p.halo_bytes;

  const catb_igemm_desc& d = p.d;
  const catb_halo_desc& h = p.h;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t0 = blockIdx.x * p.tiles_per_cta;
  const int t1 = min(p.tiles_total, t0 + p.tiles_per_cta);
  const int m_tile = blockIdx.y;
  const catb_halo_wgroup grp = p.groups[blockIdx.z];
  const catb_halo_chunk ch = p.chunks[grp.chunk];
  const uint32_t tmem_cols = grp.n_steps * 64 <= 64 ? 64 : (grp.n_steps * 64 <= 128 ? 128 : (grp.n_steps * 64 <= 256 ? 256 : 512));

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&full[i], kWProdThreads);
      mbar_init(&empty[i], 1);
    }
    mbar_init(accum, 1);
    fence_barrier_init();
  }
  if (warp == kWProdWarps) {
    tmem_alloc_dyn(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kWProdWarps) {
    const int ul = threadIdx.x & 7, rsub = threadIdx.x >> 3;
    const int Hf = d.OHs + h.Ymax;
    const bool uvalid = ul < ch.n_units;
    const __nv_bfloat16* xc = p.x + d.x_coff + (ch.cu0 + ul) * 8;
    const int per_img = p.tiles_per_strip * h.n_strips;
    for (int t = t0; t < t1; ++t) {
      const int it = t - t0;
      const int buf = it % p.bufs;
      const uint32_t ph = (it / p.bufs) & 1;
      const int n_img = t / per_img;
      const int strip = (t - n_img * per_img) / p.tiles_per_strip;
      const int m0 = (t - n_img * per_img - strip * p.tiles_per_strip) * kWPos;
      const int strip_x = strip * h.TW;
      uint8_t* dyb = bufs + static_cast<size_t>(buf) * buf_bytes;
      uint8_t* hal = dyb + kDyBytes;
      mbar_wait(&empty[buf], ph ^ 1);
      // ---- dY tile: rows = positions (zero rows for garbage positions), 2 chunks of 64 channels
      {
        int i = (m0 + rsub) / h.Wf, j = (m0 + rsub) - i * h.Wf;
        for (int r = rsub; r < kWPos; r += kWRowStep) {
          const bool pv = (i < d.OHs) & (j < h.TW) & (strip_x + j < d.OWs);
          const size_t ypix = (static_cast<size_t>(n_img) * d.OH + (d.o_ph + i * d.o_step)) * d.OW + (d.o_pw + (strip_x + j) * d.o_step);
#pragma unroll
          for (int ca = 0; ca < 2; ++ca) {
            const int c = m_tile * 128 + ca * 64 + ul * 8;
            const bool ok = pv & (c < p.cy_p);
            const __nv_bfloat16* src = ok ? p.y + ypix * d.ldy + d.y_coff + c : p.y;
            cp16_zfill(dyb + ca * (kWPos * 128) + r * 128 + ((ul ^ (r & 7)) << 4), src, ok);
          }
          j += kWRowStep;
          while (j >= h.Wf) {
            j -= h.Wf;
            ++i;
          }
        }
      }
      // ---- X halo of the chunk (same fill as the forward kernel)
      const size_t img_base = static_cast<size_t>(n_img) * d.H * d.W;
      {
        const uint32_t hal_s = smem_u32(hal);
        for (int plane = 0; plane < h.n_planes; ++plane) {
          const uint32_t plane_smem = hal_s + static_cast<uint32_t>(plane) * h.Lh * 128u;
          if (d.pad_mode == CATB_PAD_REFLECT)
            halo_fill_plane<true, kWRowStep>(plane_smem, plane * h.Lh, xc, p.x, static_cast<long long>(img_base), m0, rsub, ul, h.Lh,
                                  h.Wf, Hf, h.mul, h.plane_y0[plane], h.plane_x0[plane] + strip_x, h.plane_pa[plane],
                                  h.plane_pb[plane], d.H, d.W, d.ldx, uvalid);
          else
            halo_fill_plane<false, kWRowStep>(plane_smem, plane * h.Lh, xc, p.x, static_cast<long long>(img_base), m0, rsub, ul, h.Lh,
                                   h.Wf, Hf, h.mul, h.plane_y0[plane], h.plane_x0[plane] + strip_x, h.plane_pa[plane],
                                   h.plane_pb[plane], d.H, d.W, d.ldx, uvalid);
        }
      }
      asm volatile("cp.async.wait_all;" ::: "memory");
      fence_proxy_async();
      mbar_arrive(&full[buf]);
    }

    // ---- epilogue: accumulators -> gradient arena
    if (t1 > t0) {
      mbar_wait(accum, 0);
      tcgen05_fence_after();
      // warp w reads TMEM lanes 32 * (w % 4) ..; the four warps of a lane quarter take the taps round robin
      const int quarter = warp & 3, part = warp >> 2;
      const int row = m_tile * 128 + quarter * 32 + lane;   // channel of dY
      const bool rvalid = row < d.n_rows;
      const uint32_t trow = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
      if (p.ws != nullptr) {
        // two-stage mode: plain 64-byte runs into the split's workspace tile (column = step * 64 + channel), no atomics
        float* wrow = p.ws + (static_cast<size_t>(blockIdx.x) * d.n_rows + row) * (static_cast<size_t>(h.n_steps) * 64) +
                      static_cast<size_t>(grp.first_step) * 64;
        for (int s = part; s < grp.n_steps; s += kWProdWarps / 4) {
          for (int cc = 0; cc < 4; ++cc) {
            float acc[16];
            tmem_ld16(trow + s * 64 + cc * 16, acc);
            if (rvalid) {
#pragma unroll
              for (int q = 0; q < 4; ++q)
                *reinterpret_cast<float4*>(wrow + s * 64 + cc * 16 + q * 4) =
                    make_float4(acc[q * 4], acc[q * 4 + 1], acc[q * 4 + 2], acc[q * 4 + 3]);
            }
          }
        }
      } else
      for (int s = part; s < grp.n_steps; s += kWProdWarps / 4) {
        for (int cc = 0; cc < 4; ++cc) {
          float acc[16];
          tmem_ld16(trow + s * 64 + cc * 16, acc);
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            const catb_weight_unit wu = p.wunits[(grp.first_step + s) * 8 + cc * 2 + g];
            if (!rvalid || wu.nvalid == 0) continue;
            float* base = p.grad + wu.w_off + static_cast<long long>(row) * wu.sn_w;
            for (int q = 0; q < wu.nvalid; ++q) atomicAdd(base + q * wu.sc_w, acc[g * 8 + q]);
          }
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- MMA issuer
    // whole warp walks the loop, one elected lane issues; the taps' window offsets sit in registers, per instruction
    // one add per descriptor (common.cuh: umma_bf16_lh)
    if (t1 > t0) {
      const uint32_t idesc = make_idesc_bf16(128, 64, 1, 1);
      const uint32_t hi = sw128_desc_hi(1024);
      uint32_t boff[8];   // (a_row * 128) >> 4 of the group's taps
#pragma unroll
      for (int s = 0; s < 8; ++s) boff[s] = s < grp.n_steps ? static_cast<uint32_t>(p.steps[grp.first_step + s].a_row) * 8u : 0u;
      const uint32_t dy_lo0 = sw128_desc_lo(smem_u32(bufs), kWPos * 128);
      const uint32_t hal_lo0 = sw128_desc_lo(smem_u32(bufs) + kDyBytes, 1024);
      const uint32_t buf16 = static_cast<uint32_t>(buf_bytes) >> 4;
      uint32_t buf = 0, ph = 0, dy_lo = dy_lo0, hal_lo = hal_lo0, acc = 0;
      for (int t = t0; t < t1; ++t) {
        mbar_wait(&full[buf], ph);
        tcgen05_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int s = 0; s < 8; ++s) {
            if (s < grp.n_steps) {
              const uint32_t b_lo = hal_lo + boff[s];
#pragma unroll
              for (int k = 0; k < kWPos / 16; ++k)
                umma_bf16_lh(tmem_base + s * 64, dy_lo + k * 128, hi, b_lo + k * 128, hi, idesc, k == 0 ? acc : 1u);
            }
          }
          umma_commit(&empty[buf]);
        }
        __syncwarp();
        acc = 1u;
        if (++buf == static_cast<uint32_t>(p.bufs)) {
          buf = 0;
          ph ^= 1u;
          dy_lo = dy_lo0;
          hal_lo = hal_lo0;
        } else {
          dy_lo += buf16;
          hal_lo += buf16;
        }
      }
      if (elect_one()) umma_commit(accum);
      __syncwarp();
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == kWProdWarps) tmem_dealloc_dyn(tmem_base, tmem_cols);
}

int init_halo_wgrad_attributes() {
  const cudaError_t e =
      cudaFuncSetAttribute(igemm_halo_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(halo wgrad): %s", cudaGetErrorString(e));
    return CATB_ERR_CUDA;
  }
  return CATB_OK;
}

}  // namespace catb

using namespace catb;

extern "C" int catb_igemm_halo_wgrad_fits(int n_planes, int Lh) {
  const int halo_bytes = (n_planes * Lh * 128 + 1023) / 1024 * 1024;
  return kWHeader + 1024 + kDyBytes + halo_bytes <= 227 * 1024 ? 1 : 0;
}

static void halo_wgrad_split_plan(const catb_igemm_desc* d, const catb_halo_desc* h, int n_groups, int* tiles_total,
                                  int* tiles_per_strip, int* tiles_per_cta, int* splits) {
  *tiles_per_strip = (d->OHs * h->Wf + kWPos - 1) / kWPos;
  *tiles_total = *tiles_per_strip * h->n_strips * d->N;
  const int m_tiles = (d->n_rows + 127) / 128;
  // split the position tiles so that the grid has ~3 CTAs per SM, at least 2 tiles per CTA
  int sp = (148 * 3 + m_tiles * n_groups - 1) / (m_tiles * n_groups);
  if (sp > (*tiles_total + 1) / 2) sp = (*tiles_total + 1) / 2;
  if (sp < 1) sp = 1;
  *tiles_per_cta = (*tiles_total + sp - 1) / sp;
  *splits = (*tiles_total + *tiles_per_cta - 1) / *tiles_per_cta;
}

static int launch_halo_wgrad(const catb_igemm_desc* d, const catb_halo_desc* h, const catb_halo_step* steps,
                             const catb_halo_chunk* chunks, const catb_halo_wgroup* groups, int n_groups,
                             const catb_weight_unit* wunits, const void* x, const void* y, float* arena_grad, float* ws,
                             catb_stream_t s) {
  CATB_REQUIRE(d != nullptr && h != nullptr && n_groups > 0, "null descriptor");
  CATB_REQUIRE(h->m_sub == 1, "the weight-gradient halo kernel uses single 128-position tiles");
  CATB_REQUIRE(h->n_planes >= 1 && h->n_planes <= 4 && h->n_steps > 0 && h->n_chunks > 0, "bad halo plan");
  CATB_REQUIRE(h->TW > 0 && h->n_strips == (d->OWs + h->TW - 1) / h->TW && h->Wf == h->TW + h->Xmax &&
                   h->Lh == 128 + h->Ymax * h->Wf + h->Xmax,
               "inconsistent halo geometry");
  CATB_REQUIRE(d->ldx % 8 == 0 && d->x_coff % 8 == 0 && d->ldy % 8 == 0 && d->y_coff % 8 == 0, "pitches must be multiples of 8");
  WHaloParams p;
  p.d = *d;
  p.h = *h;
  p.steps = steps;
  p.chunks = chunks;
  p.groups = groups;
  p.wunits = wunits;
  p.x = static_cast<const __nv_bfloat16*>(x);
  p.y = static_cast<const __nv_bfloat16*>(y);
  p.grad = arena_grad;
  p.ws = ws;
  p.halo_bytes = (h->n_planes * h->Lh * 128 + 1023) / 1024 * 1024;
  const int buf_bytes = kDyBytes + p.halo_bytes;
  CATB_REQUIRE(kWHeader + 1024 + buf_bytes <= 227 * 1024, "halo tile (%d bytes) does not fit in shared memory", p.halo_bytes);
  p.bufs = (kWHeader + 1024 + 2 * buf_bytes <= 227 * 1024) ? 2 : 1;
  p.cy_p = (d->n_rows + 7) / 8 * 8;
  int splits;
  halo_wgrad_split_plan(d, h, n_groups, &p.tiles_total, &p.tiles_per_strip, &p.tiles_per_cta, &splits);
  const int m_tiles = (d->n_rows + 127) / 128;
  if (p.tiles_per_cta == 1) p.bufs = 1;
  dim3 grid(splits, m_tiles, n_groups);
  const size_t smem = 1024 + kWHeader + static_cast<size_t>(p.bufs) * buf_bytes;
  igemm_halo_wgrad_kernel<<<grid, kWThreads, smem, static_cast<cudaStream_t>(s)>>>(p);
  return check_launch("igemm_halo_wgrad");
}

extern "C" int catb_igemm_halo_wgrad(const catb_igemm_desc* d, const catb_halo_desc* h, const catb_halo_step* steps,
                                     const catb_halo_chunk* chunks, const catb_halo_wgroup* groups, int n_groups,
                                     const catb_weight_unit* wunits, const void* x, const void* y, float* arena_grad,
                                     catb_stream_t s) {
  return launch_halo_wgrad(d, h, steps, chunks, groups, n_groups, wunits, x, y, arena_grad, nullptr, s);
}

extern "C" int catb_igemm_halo_wgrad_ws_shape(const catb_igemm_desc* d, const catb_halo_desc* h, int n_groups, int* splits,
                                              int* ws_k) {
  CATB_REQUIRE(d != nullptr && h != nullptr && n_groups > 0 && splits != nullptr && ws_k != nullptr, "null pointer");
  int tt, tps, tpc;
  halo_wgrad_split_plan(d, h, n_groups, &tt, &tps, &tpc, splits);
  *ws_k = h->n_steps * 64;
  return CATB_OK;
}

extern "C" int catb_igemm_halo_wgrad_ws(const catb_igemm_desc* d, const catb_halo_desc* h, const catb_halo_step* steps,
                                        const catb_halo_chunk* chunks, const catb_halo_wgroup* groups, int n_groups,
                                        const void* x, const void* y, float* ws, catb_stream_t s) {
  CATB_REQUIRE(ws != nullptr, "null workspace");
  return launch_halo_wgrad(d, h, steps, chunks, groups, n_groups, nullptr, x, y, nullptr, ws, s);
}
