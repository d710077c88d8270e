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
#include <cuda.h>   // CUtensorMap (types only)
#include <cstring>

#include "common.cuh"

namespace catb {

// One CTA per SM (two ~60 KB operand buffers), so the producers' issue rate is what feeds the tensor pipe: 16 producer
// warps (4 per scheduler) instead of 4 -- a lone warp per scheduler issues one dependent instruction every 4-6 cycles.
constexpr int kWProdWarps = 16;
constexpr int kWProdThreads = kWProdWarps * 32;
constexpr int kWRowStep = kWProdThreads / 8;   // halo / dY rows advanced per producer iteration
constexpr int kWThreads = kWProdThreads + 32;  // + the MMA-issuing warp
constexpr int kWHeader = 1024;
constexpr int kWMaxBufs = 4;        // ring depth of the operand buffers (TMA mode: 2-4; cp.async mode: 1-2)
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
  // TMA mode (zero padding, one strip): both operands arrive as tensor-map boxes of whole frame rows; pos = lattice
  // positions (GEMM K) per tile, 128 or 64 (the smaller tile buys a deeper ring when four parity planes fill the SM)
  int use_tma, pos, dy_chunk_bytes, dy_rows, plane_rows, plane_bytes;
};

__device__ __forceinline__ void tma_load_4d_w(uint32_t dst_smem, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1, int c2,
                                              int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__device__ __forceinline__ void cp16_zfill(void* smem_dst, const void* gmem_src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(sz)
               : "memory");
}

__global__ void __launch_bounds__(kWThreads, 1)
igemm_halo_wgrad_kernel(const __grid_constant__ WHaloParams p, const __grid_constant__ CUtensorMap tmap_x,
                        const __grid_constant__ CUtensorMap tmap_y) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);  // [kWMaxBufs]
  uint64_t* empty = full + kWMaxBufs;                    // [kWMaxBufs]
  uint64_t* accum = full + 2 * kWMaxBufs;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(full + 2 * kWMaxBufs + 1);
  uint8_t* bufs = smem + kWHeader;
  const int dy_bytes = 2 * p.dy_chunk_bytes;
  const int buf_bytes = dy_bytes + p.halo_bytes;

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
    for (int i = 0; i < kWMaxBufs; ++i) {
      mbar_init(&full[i], p.use_tma ? 1 : kWProdThreads);
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
    if (p.use_tma) {
      if (threadIdx.x == 0) {
        const uint32_t tx_bytes = (2u * p.dy_rows + static_cast<uint32_t>(h.n_planes) * p.plane_rows) * h.Wf * 128u;
        for (int t = t0; t < t1; ++t) {
          const int it = t - t0;
          const int buf = it % p.bufs;
          const uint32_t ph = (it / p.bufs) & 1;
          const int n_img = t / per_img;
          const int m0 = (t - n_img * per_img) * p.pos;      // one strip per image in this mode
          const int fy0 = m0 / h.Wf;
          const uint32_t dst = smem_u32(bufs + static_cast<size_t>(buf) * buf_bytes);
          mbar_wait(&empty[buf], ph ^ 1);
          mbar_arrive_expect_tx(&full[buf], tx_bytes);
          for (int ca = 0; ca < 2; ++ca)   // dY: positions x 64 channels, rows (i >= OHs) / columns (j >= OWs) out of bounds = 0
            tma_load_4d_w(dst + ca * p.dy_chunk_bytes, &tmap_y, &full[buf], m_tile * 128 + ca * 64, 0, fy0, n_img);
          for (int pl = 0; pl < h.n_planes; ++pl)
            tma_load_4d_w(dst + dy_bytes + pl * p.plane_bytes, &tmap_x, &full[buf], ch.cu0 * 8,
                          h.mul * h.plane_x0[pl] + h.plane_pb[pl], h.mul * (fy0 + h.plane_y0[pl]) + h.plane_pa[pl], n_img);
        }
      }
      __syncwarp();
    } else
    for (int t = t0; t < t1; ++t) {
      const int it = t - t0;
      const int buf = it % p.bufs;
      const uint32_t ph = (it / p.bufs) & 1;
      const int n_img = t / per_img;
      const int strip = (t - n_img * per_img) / p.tiles_per_strip;
      const int m0 = (t - n_img * per_img - strip * p.tiles_per_strip) * kWPos;
      const int strip_x = strip * h.TW;
      uint8_t* dyb = bufs + static_cast<size_t>(buf) * buf_bytes;
      uint8_t* hal = dyb + dy_bytes;
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
      for (int s = 0; s < 8; ++s) {
        uint32_t a_row = s < grp.n_steps ? static_cast<uint32_t>(p.steps[grp.first_step + s].a_row) : 0u;
        if (p.use_tma) {   // a_row = plane * Lh + (dy * Wf + dx): planes are plane_bytes apart in this mode
          const uint32_t plane = a_row / static_cast<uint32_t>(h.Lh), rest = a_row - plane * static_cast<uint32_t>(h.Lh);
          boff[s] = plane * (static_cast<uint32_t>(p.plane_bytes) >> 4) + rest * 8u;
        } else {
          boff[s] = a_row * 8u;
        }
      }
      const uint32_t dy_lo0 = sw128_desc_lo(smem_u32(bufs), p.dy_chunk_bytes);
      const uint32_t hal_lo0 = sw128_desc_lo(smem_u32(bufs) + dy_bytes, 1024);
      const uint32_t buf16 = static_cast<uint32_t>(buf_bytes) >> 4;
      const int ksteps = p.pos / 16;
      const int per_img_m = p.tiles_per_strip * h.n_strips;
      uint32_t buf = 0, ph = 0, dy_lo = dy_lo0, hal_lo = hal_lo0, acc = 0;
      for (int t = t0; t < t1; ++t) {
        uint32_t toff = 0;   // TMA mode: the tile starts (m0 mod Wf) rows into its boxes
        if (p.use_tma) {
          const int m0 = (t - (t / per_img_m) * per_img_m) * p.pos;
          toff = static_cast<uint32_t>(m0 - (m0 / h.Wf) * h.Wf) * 8u;
        }
        mbar_wait(&full[buf], ph);
        tcgen05_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int s = 0; s < 8; ++s) {
            if (s < grp.n_steps) {
              const uint32_t a_lo = dy_lo + toff, b_lo = hal_lo + toff + boff[s];
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_lh(tmem_base + s * 64, a_lo + k * 128, hi, b_lo + k * 128, hi, idesc, k == 0 ? acc : 1u);
              if (ksteps == 8) {
#pragma unroll
                for (int k = 4; k < 8; ++k) umma_bf16_lh(tmem_base + s * 64, a_lo + k * 128, hi, b_lo + k * 128, hi, idesc, 1u);
              }
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

int encode_nhwc_tile_map(CUtensorMap* out, const void* base, int c_visible, int W, int H, int N, int ld, int box_w, int box_h,
                         int stride);   // igemm_halo_persist.cu

}  // namespace catb

using namespace catb;

extern "C" int catb_igemm_halo_wgrad_fits(int n_planes, int Lh) {
  const int halo_bytes = (n_planes * Lh * 128 + 1023) / 1024 * 1024;
  return kWHeader + 1024 + kDyBytes + halo_bytes <= 227 * 1024 ? 1 : 0;
}

// Tile plan shared by the launch and the workspace-shape query.  TMA mode: pos = 128 if at least two ring buffers fit,
// else 64 (halves both boxes); returns false when not even that fits / the boxes exceed the 256-element box limit.
struct WTilePlan {
  int pos, bufs, dy_rows, dy_chunk_bytes, plane_rows, plane_bytes, halo_bytes;
};
static bool wgrad_tile_plan(const catb_halo_desc* h, int use_tma, WTilePlan* tp) {
  if (!use_tma) {
    tp->pos = kWPos;
    tp->dy_rows = kWPos;
    tp->dy_chunk_bytes = kWPos * 128;
    tp->plane_rows = h->Lh;
    tp->plane_bytes = h->Lh * 128;
    tp->halo_bytes = (h->n_planes * h->Lh * 128 + 1023) / 1024 * 1024;
    const int buf_bytes = 2 * tp->dy_chunk_bytes + tp->halo_bytes;
    if (kWHeader + 1024 + buf_bytes > 227 * 1024) return false;
    tp->bufs = (kWHeader + 1024 + 2 * buf_bytes <= 227 * 1024) ? 2 : 1;
    return true;
  }
  for (int pos = 128; pos >= 64; pos /= 2) {
    const int Wf = h->Wf, Lh = h->Lh - 128 + pos;
    tp->pos = pos;
    tp->dy_rows = (Wf - 1 + pos + Wf - 1) / Wf;
    tp->dy_chunk_bytes = (tp->dy_rows * Wf * 128 + 1023) / 1024 * 1024;
    tp->plane_rows = (Wf - 1 + Lh + Wf - 1) / Wf;
    tp->plane_bytes = (tp->plane_rows * Wf * 128 + 1023) / 1024 * 1024;
    tp->halo_bytes = h->n_planes * tp->plane_bytes;
    if (Wf * h->mul > 256 || tp->plane_rows * h->mul > 256 || tp->dy_rows > 256) return false;
    const int buf_bytes = 2 * tp->dy_chunk_bytes + tp->halo_bytes;
    int bufs = (227 * 1024 - kWHeader - 1024) / buf_bytes;
    if (bufs > kWMaxBufs) bufs = kWMaxBufs;
    tp->bufs = bufs;
    if (bufs >= 2 || (pos == 64 && bufs >= 1)) return true;
  }
  return false;
}

extern "C" int catb_igemm_halo_wgrad_tma_fits(const catb_halo_desc* h) {
  WTilePlan tp;
  return h != nullptr && h->n_strips == 1 && wgrad_tile_plan(h, 1, &tp) ? 1 : 0;
}

static void halo_wgrad_split_plan(const catb_igemm_desc* d, const catb_halo_desc* h, int n_groups, int pos, int* tiles_total,
                                  int* tiles_per_strip, int* tiles_per_cta, int* splits) {
  *tiles_per_strip = (d->OHs * h->Wf + pos - 1) / pos;
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
                             int use_tma, int c_visible, catb_stream_t s) {
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
  p.use_tma = use_tma ? 1 : 0;
  WTilePlan tp;
  CATB_REQUIRE(wgrad_tile_plan(h, p.use_tma, &tp), "halo tile does not fit in shared memory (Lh=%d planes=%d)", h->Lh, h->n_planes);
  p.pos = tp.pos;
  p.bufs = tp.bufs;
  p.dy_rows = tp.dy_rows;
  p.dy_chunk_bytes = tp.dy_chunk_bytes;
  p.plane_rows = tp.plane_rows;
  p.plane_bytes = tp.plane_bytes;
  p.halo_bytes = tp.halo_bytes;
  const int buf_bytes = 2 * p.dy_chunk_bytes + p.halo_bytes;
  p.cy_p = (d->n_rows + 7) / 8 * 8;
  CUtensorMap tmx, tmy;
  memset(&tmx, 0, sizeof(tmx));
  memset(&tmy, 0, sizeof(tmy));
  if (p.use_tma) {
    CATB_REQUIRE(d->pad_mode != CATB_PAD_REFLECT || (h->Ymax == 0 && h->Xmax == 0),
                 "TMA-staged weight gradient needs zero padding (out-of-bounds fill)");
    CATB_REQUIRE(h->n_strips == 1 && d->o_step == 1 && d->o_ph == 0 && d->o_pw == 0 && d->OHs == d->OH && d->OWs == d->OW,
                 "TMA-staged weight gradient: one strip, dense output lattice");
    CATB_REQUIRE(c_visible > 0 && c_visible % 8 == 0 && d->x_coff + c_visible <= d->ldx, "bad visible channel count %d", c_visible);
    if (int e = encode_nhwc_tile_map(&tmx, p.x + d->x_coff, c_visible, d->W, d->H, d->N, d->ldx, h->Wf, p.plane_rows, h->mul)) return e;
    if (int e = encode_nhwc_tile_map(&tmy, p.y + d->y_coff, p.cy_p, d->OW, d->OH, d->N, d->ldy, h->Wf, p.dy_rows, 1)) return e;
  }
  int splits;
  halo_wgrad_split_plan(d, h, n_groups, p.pos, &p.tiles_total, &p.tiles_per_strip, &p.tiles_per_cta, &splits);
  const int m_tiles = (d->n_rows + 127) / 128;
  if (p.bufs > p.tiles_per_cta) p.bufs = p.tiles_per_cta;
  dim3 grid(splits, m_tiles, n_groups);
  const size_t smem = 1024 + kWHeader + static_cast<size_t>(p.bufs) * buf_bytes;
  igemm_halo_wgrad_kernel<<<grid, kWThreads, smem, static_cast<cudaStream_t>(s)>>>(p, tmx, tmy);
  return check_launch("igemm_halo_wgrad");
}

extern "C" int catb_igemm_halo_wgrad(const catb_igemm_desc* d, const catb_halo_desc* h, const catb_halo_step* steps,
                                     const catb_halo_chunk* chunks, const catb_halo_wgroup* groups, int n_groups,
                                     const catb_weight_unit* wunits, const void* x, const void* y, float* arena_grad,
                                     catb_stream_t s) {
  return launch_halo_wgrad(d, h, steps, chunks, groups, n_groups, wunits, x, y, arena_grad, nullptr, 0, 0, s);
}

extern "C" int catb_igemm_halo_wgrad_ws_shape(const catb_igemm_desc* d, const catb_halo_desc* h, int n_groups, int use_tma,
                                              int* splits, int* ws_k) {
  CATB_REQUIRE(d != nullptr && h != nullptr && n_groups > 0 && splits != nullptr && ws_k != nullptr, "null pointer");
  WTilePlan tp;
  CATB_REQUIRE(wgrad_tile_plan(h, use_tma ? 1 : 0, &tp), "halo tile does not fit in shared memory");
  int tt, tps, tpc;
  halo_wgrad_split_plan(d, h, n_groups, tp.pos, &tt, &tps, &tpc, splits);
  *ws_k = h->n_steps * 64;
  return CATB_OK;
}

extern "C" int catb_igemm_halo_wgrad_ws(const catb_igemm_desc* d, const catb_halo_desc* h, const catb_halo_step* steps,
                                        const catb_halo_chunk* chunks, const catb_halo_wgroup* groups, int n_groups,
                                        const void* x, const void* y, float* ws, int use_tma, int c_visible, catb_stream_t s) {
  CATB_REQUIRE(ws != nullptr, "null workspace");
  return launch_halo_wgrad(d, h, steps, chunks, groups, n_groups, nullptr, x, y, nullptr, ws, use_tma, c_visible, s);
}
