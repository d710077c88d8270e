// Implicit-GEMM convolution v3: the halo kernel of igemm_halo.cu as a PERSISTENT, fully warp-specialised pipeline.
//
// v2 runs one output tile per CTA: halo fill -> MMAs -> epilogue, one after the other, and relies on several CTAs per
// SM to overlap the phases (profiles/r02_timeline_*.txt: for the short-K GEMMs of the student the MMA phase is < 10 % of
// a CTA's lifetime).  Here a CTA walks tiles  blockIdx.x, blockIdx.x + gridDim.x, ...  with three pipelines that all
// keep running ACROSS tile boundaries:
//   * activation halo ring (a_full / a_empty):  filled either by ONE thread with cp.async.bulk.tensor (TMA, 4-D tiled
//     tensor map over the NHWC activation, SWIZZLE_128B, out-of-bounds = zero padding, traversal stride 2 for the
//     parity planes of stride-2 convs)  or, for reflection-padded convs, by the 128 cp.async producer threads of v2;
//     (use_tma = 2, opt-in: the boxes of a reflection-padded conv land with zeros outside the image and warps 2-3 mirror
//     exactly those halo rows before the chunk is released -- parity-green, but slower than the cp.async producers);
//   * weight ring (b_full / b_empty): 1-D bulk copies of the pre-swizzled weight tiles, as in v2 -- or, when the whole
//     weight matrix of a one-N-tile GEMM fits, one stage per tile loaded ONCE per CTA (b_stationary);
//   * TWO accumulator stages in tensor memory (acc_full / acc_empty): the four epilogue warps drain tile i
//     (tcgen05.ld -> bias / activation -> staged coalesced bf16 stores, optionally the per-channel sum / sum of squares
//     of the stored values for the norm layer behind the conv: catb_epilogue_stats) while the MMA warp already works on
//     tile i+1.
// Shared-memory layout of a TMA-filled plane: the box is R = ceil((Wf - 1 + Lh) / Wf) full frame rows of Wf pixels
// (128 bytes each), i.e. the same "pitch space" as v2 with the tile starting (m0 mod Wf) pixels into it, so a filter
// tap is still a shifted shared-memory descriptor.  Each plane starts on a 1024-byte boundary, which makes the TMA
// swizzle (a function of the shared-memory address bits) identical to what the UMMA descriptors expect.
#include <cuda.h>   // CUtensorMap and its enums only: the encoder is looked up at run time (no link to libcuda)
#include <cstring>

#include "common.cuh"

namespace catb {

constexpr int kPThreads = 320;      // warps 0-3 halo producers, 4 MMA issuer, 5 weight loader, 6-9 epilogue
constexpr int kPHeader = 1024;      // barriers [0, 512), per-warp pixel tables of the epilogue [512, 1024)
constexpr int kPMaxA = 6;
constexpr int kPMaxBStages = 18;     // (3 * 6 + 2 * 18 + 4 barriers + the TMEM slot fit the first 512 header bytes)
constexpr int kPStageBytes = 4 * 4096;   // epilogue staging: 4 KB per epilogue warp
constexpr int kPStatBytes = 2 * 2 * 256 * 4;   // fused statistics: two (tile parity) x [2][n_tile <= 256] floats

struct PersistParams {
  catb_igemm_desc d;
  catb_halo_desc h;
  const catb_halo_step* steps;
  const catb_halo_chunk* chunks;
  const __nv_bfloat16* x;
  const uint8_t* wpk;
  const float* bias;
  void* y;
  int tiles_per_image, tiles_x, tiles_total, a_bufs, b_stages, tmem_cols, acc_cols, n_store, halo_bytes, tab_bytes;
  int use_tma, plane_rows, plane_bytes;   // TMA mode: frame rows per plane box, bytes per plane (1024-aligned)
  int patch;                              // TMA mode on a reflection-padded conv: warps 2-3 mirror the out-of-image fringe
  int b_stationary;                       // every weight tile of the GEMM has its own stage and is loaded ONCE per CTA
  uint32_t idesc;
  catb_epilogue_stats st;                 // st.sums == nullptr: no fused statistics
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst_smem, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

struct TileCoord {
  int tile_n, n_img, strip_x, m0;
};
__device__ __forceinline__ TileCoord decode_tile(const PersistParams& p, int tile) {
  TileCoord t;
  t.tile_n = tile / p.tiles_x;
  const int xi = tile - t.tile_n * p.tiles_x;
  const int per_img = p.tiles_per_image * p.h.n_strips;
  t.n_img = xi / per_img;
  const int r = xi - t.n_img * per_img;
  const int strip = r / p.tiles_per_image;
  t.m0 = (r - strip * p.tiles_per_image) * (128 * p.h.m_sub);
  t.strip_x = strip * p.h.TW;
  return t;
}

__global__ void __launch_bounds__(kPThreads, 2)
igemm_halo_persist_kernel(const __grid_constant__ PersistParams p, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem);   // [kPMaxA]
  uint64_t* a_empty = a_full + kPMaxA;                     // [kPMaxA]
  uint64_t* a_tma = a_empty + kPMaxA;                      // [kPMaxA] patch mode: the box has landed (before the fringe pass)
  uint64_t* b_full = a_tma + kPMaxA;                       // [kPMaxBStages]
  uint64_t* b_empty = b_full + kPMaxBStages;               // [kPMaxBStages]
  uint64_t* acc_full = b_empty + kPMaxBStages;             // [2]
  uint64_t* acc_empty = acc_full + 2;                      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  uint32_t* s_aoff = reinterpret_cast<uint32_t*>(smem + kPHeader);                               // [n_steps], 16-byte units
  int4* s_chunks = reinterpret_cast<int4*>(smem + kPHeader + ((p.h.n_steps * 4 + 15) & ~15));   // [n_chunks]
  uint8_t* a_base = smem + kPHeader + p.tab_bytes;
  uint8_t* b_base = a_base + static_cast<size_t>(p.a_bufs) * p.halo_bytes;
  uint8_t* stg_base = b_base + static_cast<size_t>(p.b_stages) * (p.d.n_tile * 128);
  float* stat_base = reinterpret_cast<float*>(stg_base + kPStageBytes);   // [2][2][n_tile]

  const catb_igemm_desc& d = p.d;
  const catb_halo_desc& h = p.h;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b_bytes = d.n_tile * 128;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.a_bufs; ++i) {
      mbar_init(&a_full[i], p.use_tma ? (p.patch ? 64 : 1) : 128);
      mbar_init(&a_empty[i], 1);
      mbar_init(&a_tma[i], 1);
    }
    for (int i = 0; i < p.b_stages; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 4);   // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 4) {
    tmem_alloc_dyn(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < h.n_steps; i += kPThreads) {
    const uint32_t a_row = static_cast<uint32_t>(p.steps[i].a_row);
    if (p.use_tma) {
      // a_row = plane * Lh + (dy * Wf + dx): planes are plane_bytes apart here, the tile offset (m0 mod Wf) is added per tile
      const uint32_t plane = a_row / static_cast<uint32_t>(h.Lh), rest = a_row - plane * static_cast<uint32_t>(h.Lh);
      s_aoff[i] = plane * (static_cast<uint32_t>(p.plane_bytes) >> 4) + rest * 8u;
    } else {
      s_aoff[i] = a_row * 8u;
    }
  }
  for (int i = threadIdx.x; i < h.n_chunks; i += kPThreads) s_chunks[i] = reinterpret_cast<const int4*>(p.chunks)[i];
  if (p.st.sums != nullptr)
    for (int i = threadIdx.x; i < 4 * d.n_tile; i += kPThreads) stat_base[i] = 0.f;
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ---------------------------------------------------------------- halo producers
    uint32_t g = 0;   // chunks staged so far by this CTA (ring position across tiles)
    if (p.use_tma && p.patch && warp >= 2) {
      // ---- reflection padding: wait for the chunk's boxes, mirror the out-of-image halo rows, then release the chunk
      const int t64 = threadIdx.x - 64, ul = t64 & 7, rsub = t64 >> 3;   // 8 rows per pass, 8 lanes per row
      for (int tile = blockIdx.x; tile < p.tiles_total; tile += gridDim.x) {
        const TileCoord t = decode_tile(p, tile);
        const size_t img_base = static_cast<size_t>(t.n_img) * d.H * d.W;
        const int row0 = (t.m0 / h.Wf) * h.Wf;
        for (int c = 0; c < h.n_chunks; ++c, ++g) {
          const uint32_t buf = g % p.a_bufs, ph = (g / p.a_bufs) & 1;
          const int4 chv = s_chunks[c];
          const __nv_bfloat16* xc = p.x + d.x_coff + (chv.x + ul) * 8;
          mbar_wait(&a_tma[buf], ph);
          const uint32_t abuf_s = smem_u32(a_base + static_cast<size_t>(buf) * p.halo_bytes);
          for (int pl = 0; pl < h.n_planes; ++pl)
            halo_patch_reflect<8>(abuf_s + pl * p.plane_bytes, xc, p.x, static_cast<long long>(img_base), row0, rsub, ul,
                                  p.plane_rows * h.Wf, h.Wf, h.mul, h.plane_y0[pl], h.plane_x0[pl] + t.strip_x, h.plane_pa[pl],
                                  h.plane_pb[pl], d.H, d.W, d.ldx, ul < chv.y);
          cp_async_wait_all();
          fence_proxy_async();
          mbar_arrive(&a_full[buf]);
        }
      }
    } else if (p.use_tma) {
      if (threadIdx.x == 0) {
        uint64_t* a_land = p.patch ? a_tma : a_full;     // patch mode: the fringe pass stands between the box and the MMAs
        const uint32_t tx_bytes = static_cast<uint32_t>(h.n_planes) * p.plane_rows * h.Wf * 128u;
        for (int tile = blockIdx.x; tile < p.tiles_total; tile += gridDim.x) {
          const TileCoord t = decode_tile(p, tile);
          const int fy0 = t.m0 / h.Wf;
          for (int c = 0; c < h.n_chunks; ++c, ++g) {
            const uint32_t buf = g % p.a_bufs, ph = (g / p.a_bufs) & 1;
            const int cu0 = s_chunks[c].x;
            mbar_wait(&a_empty[buf], ph ^ 1);
            mbar_arrive_expect_tx(&a_land[buf], tx_bytes);
            const uint32_t dst = smem_u32(a_base + static_cast<size_t>(buf) * p.halo_bytes);
            for (int pl = 0; pl < h.n_planes; ++pl)
              tma_load_4d(dst + pl * p.plane_bytes, &tmap, &a_land[buf], cu0 * 8,
                          h.mul * (h.plane_x0[pl] + t.strip_x) + h.plane_pb[pl], h.mul * (fy0 + h.plane_y0[pl]) + h.plane_pa[pl],
                          t.n_img);
          }
        }
      }
      __syncwarp();   // warp 0 reconverges before the block-wide barrier at the end
    } else {
      const int ul = threadIdx.x & 7, rsub = threadIdx.x >> 3;   // 16 rows per pass, 8 lanes per row
      const int Hf = d.OHs + h.Ymax;
      for (int tile = blockIdx.x; tile < p.tiles_total; tile += gridDim.x) {
        const TileCoord t = decode_tile(p, tile);
        const size_t img_base = static_cast<size_t>(t.n_img) * d.H * d.W;
        for (int c = 0; c < h.n_chunks; ++c, ++g) {
          const uint32_t buf = g % p.a_bufs, ph = (g / p.a_bufs) & 1;
          const int4 chv = s_chunks[c];
          const bool uvalid = ul < chv.y;
          const bool ufill = ul < ((chv.y + 1) & ~1);   // columns the MMAs of this chunk read (K = 16 granularity)
          const __nv_bfloat16* xc = p.x + d.x_coff + (chv.x + ul) * 8;
          mbar_wait(&a_empty[buf], ph ^ 1);
          if (ufill) {
            const uint32_t abuf_s = smem_u32(a_base + static_cast<size_t>(buf) * p.halo_bytes);
            for (int plane = 0; plane < h.n_planes; ++plane) {
              const uint32_t plane_smem = abuf_s + static_cast<uint32_t>(plane) * h.Lh * 128u;
              if (d.pad_mode == CATB_PAD_REFLECT)
                halo_fill_plane<true>(plane_smem, plane * h.Lh, xc, p.x, static_cast<long long>(img_base), t.m0, rsub, ul, h.Lh,
                                      h.Wf, Hf, h.mul, h.plane_y0[plane], h.plane_x0[plane] + t.strip_x, h.plane_pa[plane],
                                      h.plane_pb[plane], d.H, d.W, d.ldx, uvalid);
              else
                halo_fill_plane<false>(plane_smem, plane * h.Lh, xc, p.x, static_cast<long long>(img_base), t.m0, rsub, ul, h.Lh,
                                       h.Wf, Hf, h.mul, h.plane_y0[plane], h.plane_x0[plane] + t.strip_x, h.plane_pa[plane],
                                       h.plane_pb[plane], d.H, d.W, d.ldx, uvalid);
            }
          }
          cp_async_wait_all();
          fence_proxy_async();
          mbar_arrive(&a_full[buf]);
        }
      }
    }
  } else if (warp == 4) {
    // ---------------------------------------------------------------- MMA issuer (whole warp walks, one lane issues)
    const uint32_t hi = sw128_desc_hi(1024);
    const uint32_t b_lo0 = sw128_desc_lo(smem_u32(b_base), 16);
    const uint32_t b_step = static_cast<uint32_t>(b_bytes) >> 4;
    const uint32_t a_step = static_cast<uint32_t>(p.halo_bytes) >> 4;
    const uint32_t a_lo0 = sw128_desc_lo(smem_u32(a_base), 16);
    const uint32_t idesc = p.idesc;
    const uint32_t n_tile = d.n_tile, m_sub = h.m_sub;
    uint32_t st = 0, phb = 0, b_lo = b_lo0;     // weight ring position
    uint32_t buf = 0, pha = 0, a_lo = a_lo0;    // halo ring position
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.tiles_total; tile += gridDim.x, ++it) {
      const uint32_t as = it & 1u, aph = (it >> 1) & 1u;
      uint32_t tile_aoff = 0;
      if (p.use_tma) {
        const TileCoord t = decode_tile(p, tile);
        tile_aoff = static_cast<uint32_t>(t.m0 - (t.m0 / h.Wf) * h.Wf) * 8u;
      }
      mbar_wait(&acc_empty[as], aph ^ 1u);       // the epilogue has drained this accumulator stage
      tcgen05_fence_after();
      const uint32_t t_acc = tmem_base + as * static_cast<uint32_t>(p.acc_cols);
      uint32_t acc = 0;                          // 0 only for the first instruction into each accumulator
      for (int c = 0; c < h.n_chunks; ++c) {
        const int4 ch = s_chunks[c];             // cu0, n_units, first_step, n_steps
        const int kmax = (ch.y + 1) >> 1;        // K = 16 slices of the chunk that hold real channels
        mbar_wait(&a_full[buf], pha);
        tcgen05_fence_after();
        const int s_end = ch.z + ch.w;
        uint32_t aoff = s_aoff[ch.z];
        for (int s = ch.z; s < s_end; ++s) {
          const uint32_t a_cur = a_lo + tile_aoff + aoff;
          if (s + 1 < s_end) aoff = s_aoff[s + 1];
          mbar_wait(&b_full[st], phb);
          tcgen05_fence_after();
          if (elect_one()) {
            for (uint32_t sub = 0; sub < m_sub; ++sub) {
              const uint32_t a_s = a_cur + sub * 1024u, t_s = t_acc + sub * n_tile;
              umma_bf16_lh(t_s, a_s, hi, b_lo, hi, idesc, acc);
              if (kmax > 1) umma_bf16_lh(t_s, a_s + 2, hi, b_lo + 2, hi, idesc, 1u);
              if (kmax > 2) umma_bf16_lh(t_s, a_s + 4, hi, b_lo + 4, hi, idesc, 1u);
              if (kmax > 3) umma_bf16_lh(t_s, a_s + 6, hi, b_lo + 6, hi, idesc, 1u);
            }
            if (!p.b_stationary) umma_commit(&b_empty[st]);
          }
          __syncwarp();
          acc = 1u;
          if (++st == static_cast<uint32_t>(p.b_stages)) {
            st = 0;
            if (!p.b_stationary) phb ^= 1u;   // stationary tiles: every later wait on the completed phase 0 passes at once
            b_lo = b_lo0;
          } else {
            b_lo += b_step;
          }
        }
        if (elect_one()) umma_commit(&a_empty[buf]);
        __syncwarp();
        if (++buf == static_cast<uint32_t>(p.a_bufs)) {
          buf = 0;
          pha ^= 1u;
          a_lo = a_lo0;
        } else {
          a_lo += a_step;
        }
      }
      if (elect_one()) umma_commit(&acc_full[as]);
      __syncwarp();
    }
  } else if (warp == 5) {
    // ---------------------------------------------------------------- weight loader
    if (lane == 0 && p.b_stationary) {
      // the whole weight matrix (one N tile, <= kPMaxBStages steps) fits: it is fetched once and stays for every tile of
      // this CTA -- a short-K GEMM otherwise re-streams more weight bytes than activation bytes per tile
      for (int s = 0; s < h.n_steps; ++s) {
        mbar_arrive_expect_tx(&b_full[s], b_bytes);
        bulk_g2s(b_base + static_cast<size_t>(s) * b_bytes, p.wpk + static_cast<size_t>(s) * b_bytes, b_bytes, &b_full[s]);
      }
    } else if (lane == 0) {
      uint32_t sg = 0;
      for (int tile = blockIdx.x; tile < p.tiles_total; tile += gridDim.x) {
        const int tile_n = tile / p.tiles_x;
        const uint8_t* src = p.wpk + static_cast<size_t>(tile_n) * h.n_steps * b_bytes;
        for (int s = 0; s < h.n_steps; ++s, ++sg) {
          const uint32_t st = sg % p.b_stages, phb = (sg / p.b_stages) & 1;
          mbar_wait(&b_empty[st], phb ^ 1);
          mbar_arrive_expect_tx(&b_full[st], b_bytes);
          bulk_g2s(b_base + static_cast<size_t>(st) * b_bytes, src + static_cast<size_t>(s) * b_bytes, b_bytes, &b_full[st]);
        }
      }
    }
    __syncwarp();
  } else {
    // ---------------------------------------------------------------- epilogue (warps 6-9 = TMEM lane quarters 2,3,0,1)
    const int q = warp & 3;
    uint8_t* stg = stg_base + q * 4096;
    uint32_t* pixtab = reinterpret_cast<uint32_t*>(smem + 512) + q * 32;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.tiles_total; tile += gridDim.x, ++it) {
      const uint32_t as = it & 1u, aph = (it >> 1) & 1u;
      const TileCoord t = decode_tile(p, tile);
      float* stat_s = p.st.sums != nullptr ? stat_base + as * 2 * d.n_tile : nullptr;
      mbar_wait(&acc_full[as], aph);
      tcgen05_fence_after();
      for (int sub = 0; sub < h.m_sub; ++sub) {
        const int m = t.m0 + sub * 128 + q * 32 + lane;
        const int i = m / h.Wf, j = m - i * h.Wf;
        const int jg = t.strip_x + j;
        const bool rvalid = (i < d.OHs) & (j < h.TW) & (jg < d.OWs);
        const size_t ypix = (static_cast<size_t>(t.n_img) * d.OH + (d.o_ph + i * d.o_step)) * d.OW + (d.o_pw + jg * d.o_step);
        const uint32_t trow = tmem_base + as * static_cast<uint32_t>(p.acc_cols) + (static_cast<uint32_t>(q * 32) << 16) + sub * d.n_tile;
        if (!d.y_is_f32 && !d.accumulate) {
          epilogue_rows_bf16(trow, d.n_tile, t.tile_n * d.n_tile, p.n_store, d.n_rows, p.bias, d.act, rvalid,
                             static_cast<uint32_t>(ypix), reinterpret_cast<__nv_bfloat16*>(p.y), d.ldy, d.y_coff, stg, pixtab, lane,
                             0, nullptr, stat_s);
          continue;
        }
        for (int cc = 0; cc < d.n_tile / 16; ++cc) {
          float acc[16];
          tmem_ld16(trow + cc * 16, acc);
          const int col0 = t.tile_n * d.n_tile + cc * 16;
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            const int col = col0 + g * 8;
            if (!rvalid || col >= p.n_store) continue;
            f8 o;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              float v = acc[g * 8 + e];
              if (p.bias != nullptr && col + e < d.n_rows) v += __ldg(p.bias + col + e);
              o.v[e] = v;
            }
            if (d.y_is_f32) {
              float* yp = reinterpret_cast<float*>(p.y) + ypix * d.ldy + d.y_coff + col;
              if (d.accumulate) {
#pragma unroll
                for (int e = 0; e < 8; ++e) o.v[e] += yp[e];
              }
#pragma unroll
              for (int e = 0; e < 8; ++e) o.v[e] = apply_act(o.v[e], d.act);
              *reinterpret_cast<float4*>(yp) = make_float4(o.v[0], o.v[1], o.v[2], o.v[3]);
              *reinterpret_cast<float4*>(yp + 4) = make_float4(o.v[4], o.v[5], o.v[6], o.v[7]);
            } else {
              __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(p.y) + ypix * d.ldy + d.y_coff + col;
              if (d.accumulate) {
                const f8 old = unpack8(ld16(yp));
#pragma unroll
                for (int e = 0; e < 8; ++e) o.v[e] += old.v[e];
              }
#pragma unroll
              for (int e = 0; e < 8; ++e) o.v[e] = apply_act(o.v[e], d.act);
              st16(yp, pack8(o));
            }
          }
        }
      }
      // every tcgen05.ld of this warp has completed (wait::ld inside the loaders): hand the stage back to the MMA warp
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[as]);
      if (stat_s != nullptr) {
        // the four warps have added their rows of this tile: one of the two tiles (by parity, so the next tile's adds need
        // no second barrier) goes out with one atomic per column, image and CTA
        named_bar_sync(1, 128);
        flush_epilogue_stats(stat_s, d.n_tile, t.tile_n * d.n_tile, p.n_store, p.st, t.n_img, threadIdx.x - 192, 128);
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc_dyn(tmem_base, p.tmem_cols);
}

int init_halo_persist_attributes() {
  const cudaError_t e =
      cudaFuncSetAttribute(igemm_halo_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(halo persist): %s", cudaGetErrorString(e));
    return CATB_ERR_CUDA;
  }
  return CATB_OK;
}

typedef CUresult (*TensorMapEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                           const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                           CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static TensorMapEncodeTiledFn tensor_map_encoder() {
  static TensorMapEncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<TensorMapEncodeTiledFn>(ptr);
  }
  return fn;
}

// 4-D tiled tensor map over an NHWC bf16 activation slice: dims (C = c_visible, W, H, N), box 64 channels x box_w x box_h
// pixels (x 1 image) traversed with `stride` along W and H, SWIZZLE_128B, out-of-bounds elements read as zero.
int encode_nhwc_tile_map(CUtensorMap* out, const void* base, int c_visible, int W, int H, int N, int ld, int box_w, int box_h,
                         int stride) {
  TensorMapEncodeTiledFn enc = tensor_map_encoder();
  CATB_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t dims[4] = {static_cast<cuuint64_t>(c_visible), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                              static_cast<cuuint64_t>(N)};
  const cuuint64_t strides[3] = {static_cast<cuuint64_t>(ld) * 2, static_cast<cuuint64_t>(W) * ld * 2,
                                 static_cast<cuuint64_t>(H) * W * ld * 2};
  const cuuint32_t box[4] = {64, static_cast<cuuint32_t>(box_w * stride), static_cast<cuuint32_t>(box_h * stride), 1};
  const cuuint32_t estr[4] = {1, static_cast<cuuint32_t>(stride), static_cast<cuuint32_t>(stride), 1};
  const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CATB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r));
  return CATB_OK;
}

}  // namespace catb

using namespace catb;

static int persist_table_bytes(int n_steps, int n_chunks) {
  return (((n_steps * 4 + 15) & ~15) + n_chunks * 16 + 1023) / 1024 * 1024;
}

// Frame rows a plane box must hold: the tile starts (m0 mod Wf) pixels into its first row, m0 a multiple of the tile size,
// so the largest offset is Wf - gcd(tile, Wf) (a strip as wide as a power of two needs no slack row at all).
static int gcd_int(int a, int b) { return b == 0 ? a : gcd_int(b, a % b); }
static int persist_plane_rows(int Lh, int Wf, int tile_positions) {
  return (Wf - gcd_int(tile_positions, Wf) + Lh + Wf - 1) / Wf;
}

static int persist_halo_bytes(int n_planes, int Lh, int Wf, int m_sub, int use_tma, int* plane_bytes) {
  if (use_tma) {
    *plane_bytes = (persist_plane_rows(Lh, Wf, 128 * m_sub) * Wf * 128 + 1023) / 1024 * 1024;
    return n_planes * *plane_bytes;
  }
  *plane_bytes = Lh * 128;
  return (n_planes * Lh * 128 + 1023) / 1024 * 1024;
}

// Shared-memory plan: 0 and a_bufs / b_stages / total bytes, or -1 when it does not fit.
static int persist_smem_plan(int halo_bytes, int b_bytes, int tab_bytes, int budget, int n_steps, int* a_bufs, int* b_stages,
                             size_t* total) {
  const int limit = 227 * 1024 - 1024 /*alignment slack*/ - kPHeader - tab_bytes - kPStageBytes - kPStatBytes;
  // weight ring: >= ~64 KB in flight (bulk-copy latency), or the caller's smaller budget (thin GEMMs: more CTAs per SM)
  const int bud = budget > 0 ? budget : 64 * 1024;
  int want_b = (bud + b_bytes - 1) / b_bytes;
  const int min_b = budget > 0 ? 2 : 4;
  if (want_b < min_b) want_b = min_b;
  if (want_b > kPMaxBStages) want_b = kPMaxBStages;
  // Halo ring.  Two buffers let the fill of the next chunk run under this one's MMAs; a GEMM that streams its activations
  // (short K: 1x1 convs, the HBM-bound student / teacher layers) needs ~96 KB of loads in flight per SM to cover the
  // DRAM latency (ncu on the teacher's fused 1x1: two 20 KB boxes in flight = 1.7 TB/s), so small halos get a deeper
  // ring -- unless the caller asked for the small-footprint variant (b_budget > 0: several CTAs per SM do the same job).
  int ab = 2;
  if (2 * halo_bytes + 2 * b_bytes > limit) ab = 1;
  if (ab * halo_bytes + 2 * b_bytes > limit) return -1;
  if (budget == 0 && ab == 2) {
    int want_a = (96 * 1024 + halo_bytes - 1) / halo_bytes;
    if (want_a > kPMaxA) want_a = kPMaxA;
    const int keep_b = (want_b < 3 ? want_b : 3) * b_bytes;   // weight tiles come from L2: three stages suffice if space is short
    while (want_a > 2 && want_a * halo_bytes + keep_b > limit) --want_a;
    ab = want_a;
  }
  int bs = (limit - ab * halo_bytes) / b_bytes;
  // a stage per weight tile when the whole matrix fits (the launch then keeps it resident if the GEMM has one N tile)
  if (budget == 0 && n_steps <= kPMaxBStages && bs >= n_steps && n_steps >= 2 && n_steps * b_bytes <= 112 * 1024) want_b = n_steps;
  if (bs > want_b) bs = want_b;
  if (bs < 2) return -1;
  *a_bufs = ab;
  *b_stages = bs;
  *total = 1024 + kPHeader + tab_bytes + static_cast<size_t>(ab) * halo_bytes + static_cast<size_t>(bs) * b_bytes + kPStageBytes +
           kPStatBytes;
  return 0;
}

extern "C" int catb_igemm_halo_persist_fits(int n_planes, int Lh, int Wf, int mul, int n_tile, int m_sub, int n_steps, int n_chunks,
                                            int b_budget, int use_tma) {
  if (2 * m_sub * n_tile > 512) return 0;
  if (use_tma && (Wf * mul > 256 || persist_plane_rows(Lh, Wf, 128 * m_sub) * mul > 256)) return 0;
  int plane_bytes, ab, bs;
  size_t total;
  const int halo_bytes = persist_halo_bytes(n_planes, Lh, Wf, m_sub, use_tma, &plane_bytes);
  return persist_smem_plan(halo_bytes, n_tile * 128, persist_table_bytes(n_steps, n_chunks), b_budget, n_steps, &ab, &bs, &total) == 0
             ? 1 : 0;
}

extern "C" int catb_igemm_halo_fprop_persist(const catb_igemm_desc* d, const catb_halo_desc* h, const catb_halo_step* steps,
                                             const catb_halo_chunk* chunks, const void* x, const void* packed_w,
                                             const float* bias, void* y, int use_tma, int c_visible,
                                             const catb_epilogue_stats* stats, catb_stream_t s) {
  CATB_REQUIRE(d != nullptr && h != nullptr, "null descriptor");
  CATB_REQUIRE(d->n_tile % 16 == 0 && d->n_tile >= 16 && d->n_tile <= 256, "n_tile must be a multiple of 16 in [16,256]");
  CATB_REQUIRE(h->m_sub >= 1 && h->m_sub <= 4 && 2 * h->m_sub * d->n_tile <= 512,
               "two accumulator stages of m_sub * n_tile columns must fit 512 TMEM columns");
  CATB_REQUIRE(h->n_planes >= 1 && h->n_planes <= 4 && h->n_steps > 0 && h->n_chunks > 0, "bad halo plan");
  CATB_REQUIRE(h->TW > 0 && h->n_strips == (d->OWs + h->TW - 1) / h->TW && h->Wf == h->TW + h->Xmax &&
                   h->Lh == 128 * h->m_sub + h->Ymax * h->Wf + h->Xmax,
               "inconsistent halo geometry");
  CATB_REQUIRE(d->n_units == h->n_steps * 8, "unit table must hold 8 units per step");
  CATB_REQUIRE(d->ldx % 8 == 0 && d->x_coff % 8 == 0 && d->ldy % 8 == 0 && d->y_coff % 8 == 0, "pitches must be multiples of 8");
  PersistParams p;
  p.d = *d;
  p.h = *h;
  p.steps = steps;
  p.chunks = chunks;
  p.x = static_cast<const __nv_bfloat16*>(x);
  p.wpk = static_cast<const uint8_t*>(packed_w);
  p.bias = bias;
  p.y = y;
  memset(&p.st, 0, sizeof(p.st));
  if (stats != nullptr && stats->sums != nullptr) {
    CATB_REQUIRE(!d->y_is_f32 && !d->accumulate, "fused statistics need the plain bf16 store");
    CATB_REQUIRE(stats->C > 0 && stats->coff >= 0 && stats->coff + (d->n_rows + 7) / 8 * 8 <= stats->C, "bad statistics slice");
    p.st = *stats;
  }
  p.use_tma = use_tma ? 1 : 0;
  p.patch = use_tma == 2 ? 1 : 0;
  p.plane_rows = persist_plane_rows(h->Lh, h->Wf, 128 * h->m_sub);
  p.halo_bytes = persist_halo_bytes(h->n_planes, h->Lh, h->Wf, h->m_sub, p.use_tma, &p.plane_bytes);
  p.tab_bytes = persist_table_bytes(h->n_steps, h->n_chunks);
  size_t smem = 0;
  CATB_REQUIRE(persist_smem_plan(p.halo_bytes, d->n_tile * 128, p.tab_bytes, h->b_budget, h->n_steps, &p.a_bufs, &p.b_stages,
                                 &smem) == 0,
               "halo tile (%d bytes) does not fit in shared memory", p.halo_bytes);
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  if (p.use_tma) {
    CATB_REQUIRE(p.patch || d->pad_mode != CATB_PAD_REFLECT || (h->Ymax == 0 && h->Xmax == 0),
                 "TMA-staged tiles of a reflection-padded conv need the fringe pass (use_tma = 2)");
    CATB_REQUIRE(h->Wf * h->mul <= 256 && p.plane_rows * h->mul <= 256, "TMA box exceeds 256 elements per dimension");
    CATB_REQUIRE(c_visible > 0 && c_visible % 8 == 0 && d->x_coff + c_visible <= d->ldx, "bad visible channel count %d", c_visible);
    // NHWC activation as a 4-D tensor (C, W, H, N), channels beyond the GEMM's own slice out of bounds (-> zero)
    if (int e = encode_nhwc_tile_map(&tmap, p.x + d->x_coff, c_visible, d->W, d->H, d->N, d->ldx, h->Wf, p.plane_rows, h->mul)) return e;
  }
  p.b_stationary = (p.b_stages == h->n_steps && (d->n_rows + d->n_tile - 1) / d->n_tile == 1) ? 1 : 0;
  const int positions = d->OHs * h->Wf;
  p.tiles_per_image = (positions + 128 * h->m_sub - 1) / (128 * h->m_sub);
  p.tiles_x = p.tiles_per_image * h->n_strips * d->N;
  const int n_tiles = (d->n_rows + d->n_tile - 1) / d->n_tile;
  p.tiles_total = p.tiles_x * n_tiles;
  p.acc_cols = h->m_sub * d->n_tile;
  uint32_t cols = 32;
  while (static_cast<int>(cols) < 2 * p.acc_cols) cols <<= 1;
  p.tmem_cols = cols;
  p.n_store = (d->n_rows + 7) / 8 * 8;
  p.idesc = make_idesc_bf16(128, d->n_tile, 0, 0);
  // persistent grid: as many CTAs as stay resident (registers / shared memory / tensor memory), each walking its tiles
  int occ = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, igemm_halo_persist_kernel, kPThreads, smem) != cudaSuccess || occ < 1) {
    cudaGetLastError();
    occ = 1;
  }
  if (occ > 512 / static_cast<int>(cols)) occ = 512 / static_cast<int>(cols);
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm <= 0) n_sm = 148;
  }
  int grid = n_sm * occ;
  if (grid > p.tiles_total) grid = p.tiles_total;
  igemm_halo_persist_kernel<<<grid, kPThreads, smem, static_cast<cudaStream_t>(s)>>>(p, tmap);
  return check_launch("igemm_halo_fprop_persist");
}
