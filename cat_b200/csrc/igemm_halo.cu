// Implicit-GEMM convolution v2: shared-memory halo tile + shifted-window UMMA descriptors.
//
// v1 (igemm.cu) re-gathers the activation tile from L2 once per filter tap; profiling showed every shape
// bound by that gather (~2.5 TB/s L2->SM irrespective of N).  Here a CTA owns 128*m_sub consecutive
// positions of ONE image in "pitch space" (m = i*Wf + j, Wf = lattice width + tap extent) and stages, once
// per 64-channel chunk, the contiguous range of frame pixels those positions touch (the halo, one copy per
// input parity plane for stride-2 convs).  A filter tap is then just a row offset into that buffer: the A
// descriptor of tap (dy,dx) starts at halo_row = plane*Lh + dy*Wf + dx, all taps of the chunk reuse the
// same shared-memory bytes, and the gather traffic drops by the tap count (25x for 5x5, 16x for 4x4, ...).
// A window may start at any 128-byte row of the (1024-byte aligned) tile: the SW128 XOR is taken from the
// absolute shared-memory address bits, so the descriptor's base-offset field stays 0 (measured on B200; the
// other reading of the PTX text, base_offset = (start >> 7) & 7, gives wrong results -- CATB_HALO_BO=1 keeps
// it selectable for that experiment).
//
// Warp roles as in v1: warps 0-3 fill the halo then run the epilogue, warp 4 issues tcgen05.mma for every
// (chunk, tap, sub-tile), warp 5 streams the pre-swizzled weight tile of each (chunk, tap) step with one
// bulk copy.  Up to four 128-row sub-tiles share every weight tile: a thin GEMM (N <= 64) issues an MMA every ~45
// cycles, i.e. with one sub-tile it would pull a fresh weight tile from L2 every ~180 cycles per SM (~10 TB/s over the
// chip); sharing the tile between sub-tiles divides that traffic.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace catb {

constexpr int kHThreads = 192;
constexpr int kHHeader = 1024;
constexpr int kHStatBytes = 2048;   // fused statistics tile [2][n_tile <= 256] floats, between the tables and the halo
constexpr int kHMaxBStages = 24;  // small weight tiles need many copies in flight (a 6 KB tile per ~1.5 us round trip otherwise)

struct HaloParams {
  catb_igemm_desc d;
  catb_halo_desc h;
  const catb_halo_step* steps;
  const catb_halo_chunk* chunks;
  const __nv_bfloat16* x;
  const uint8_t* wpk;
  const float* bias;
  void* y;
  int tiles_per_image, a_bufs, b_stages, tmem_cols, n_store, halo_bytes, tab_bytes, dbg_mode;
  uint32_t idesc;
  unsigned long long* dbg;  // optional per-CTA phase timestamps (catb_debug_timeline), null in production
  catb_epilogue_stats st;   // st.sums == nullptr: no fused statistics
};

static unsigned long long* g_halo_dbg = nullptr;
static int g_halo_dbg_mode = 0;   // experiment switches of the epilogue (catb_debug_mode): 1 no global stores, 2 no TMEM loads, 4 direct stores
constexpr int kDbgSlots = 16, kDbgCtas = 4096;
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define DBG_STAMP(slot)                                                                              \
  do {                                                                                               \
    if (p.dbg != nullptr && blockIdx.y == 0 && blockIdx.x < kDbgCtas) p.dbg[blockIdx.x * kDbgSlots + (slot)] = gtimer(); \
  } while (0)

__global__ void __maxnreg__(112) igemm_halo_fprop_kernel(const HaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem);  // [2]
  uint64_t* a_empty = a_full + 2;                         // [2]
  uint64_t* b_full = a_full + 4;                          // [kHMaxBStages]
  uint64_t* b_empty = b_full + kHMaxBStages;              // [kHMaxBStages]
  uint64_t* accum = b_empty + kHMaxBStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum + 1);
  // step / chunk tables in shared memory: the MMA thread must not wait for global loads between instructions
  uint32_t* s_aoff = reinterpret_cast<uint32_t*>(smem + kHHeader);                        // [n_steps] a_row * 8
  int4* s_chunks = reinterpret_cast<int4*>(smem + kHHeader + ((p.h.n_steps * 4 + 15) & ~15));  // [n_chunks]
  float* stat_s = reinterpret_cast<float*>(smem + kHHeader + p.tab_bytes);
  uint8_t* a_base = smem + kHHeader + p.tab_bytes + kHStatBytes;
  uint8_t* b_base = a_base + static_cast<size_t>(p.a_bufs) * p.halo_bytes;

  const catb_igemm_desc& d = p.d;
  const catb_halo_desc& h = p.h;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) DBG_STAMP(7);   // kernel entry
  const int per_img = p.tiles_per_image * h.n_strips;   // tiles_per_image = tiles per (image, strip)
  const int n_img = blockIdx.x / per_img;
  const int strip = (blockIdx.x - n_img * per_img) / p.tiles_per_image;
  const int m0 = (blockIdx.x - n_img * per_img - strip * p.tiles_per_image) * (128 * h.m_sub);
  const int strip_x = strip * h.TW;
  const int tile_n = blockIdx.y;
  const int b_bytes = d.n_tile * 128;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], 128);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < p.b_stages; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    mbar_init(accum, 1);
    fence_barrier_init();
  }
  if (warp == 4) {
    tmem_alloc_dyn(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < h.n_steps; i += kHThreads) s_aoff[i] = static_cast<uint32_t>(p.steps[i].a_row) * 8u;
  for (int i = threadIdx.x; i < h.n_chunks; i += kHThreads) s_chunks[i] = reinterpret_cast<const int4*>(p.chunks)[i];
  if (p.st.sums != nullptr)
    for (int i = threadIdx.x; i < 2 * d.n_tile; i += kHThreads) stat_s[i] = 0.f;
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) DBG_STAMP(0);   // prologue done (barriers, TMEM, tables)

  if (warp < 4) {
    // ---------------------------------------------------------------- halo producer
    const int ul = threadIdx.x & 7, rsub = threadIdx.x >> 3;  // 16 rows per pass, 8 lanes per row
    const int Hf = d.OHs + h.Ymax;                              // frame height
    const size_t img_base = static_cast<size_t>(n_img) * d.H * d.W;
    for (int c = 0; c < h.n_chunks; ++c) {
      const int buf = c % p.a_bufs;
      const uint32_t ph = (c / p.a_bufs) & 1;
      const int4 chv = s_chunks[c];
      catb_halo_chunk ch;
      ch.cu0 = chv.x;
      ch.n_units = chv.y;
      const bool uvalid = ul < ch.n_units;
      const bool ufill = ul < ((ch.n_units + 1) & ~1);   // columns the MMAs of this chunk read (K = 16 granularity)
      const __nv_bfloat16* xc = p.x + d.x_coff + (ch.cu0 + ul) * 8;
      uint8_t* abuf = a_base + static_cast<size_t>(buf) * p.halo_bytes;
      mbar_wait(&a_empty[buf], ph ^ 1);
      // Every thread issues all of its 16-byte copies back to back with cp.async (LDGSTS, zero-fill for
      // padding / out-of-frame pixels): hundreds of loads in flight per SM, no register staging.
      if (ufill) {
        const uint32_t abuf_s = smem_u32(abuf);
        for (int plane = 0; plane < h.n_planes; ++plane) {
          const uint32_t plane_smem = abuf_s + static_cast<uint32_t>(plane) * h.Lh * 128u;
          if (d.pad_mode == CATB_PAD_REFLECT)
            halo_fill_plane<true>(plane_smem, plane * h.Lh, xc, p.x, static_cast<long long>(img_base), m0, rsub, ul, h.Lh,
                                  h.Wf, Hf, h.mul, h.plane_y0[plane], h.plane_x0[plane] + strip_x, h.plane_pa[plane],
                                  h.plane_pb[plane], d.H, d.W, d.ldx, uvalid);
          else
            halo_fill_plane<false>(plane_smem, plane * h.Lh, xc, p.x, static_cast<long long>(img_base), m0, rsub, ul, h.Lh,
                                   h.Wf, Hf, h.mul, h.plane_y0[plane], h.plane_x0[plane] + strip_x, h.plane_pa[plane],
                                   h.plane_pb[plane], d.H, d.W, d.ldx, uvalid);
        }
      }
      cp_async_wait_all();
      fence_proxy_async();
      mbar_arrive(&a_full[buf]);
      if (threadIdx.x == 0 && c == 0) DBG_STAMP(1);   // first halo chunk filled
    }

    // ---------------------------------------------------------------- epilogue
    mbar_wait(accum, 0);
    tcgen05_fence_after();
    if (threadIdx.x == 0) DBG_STAMP(4);   // accumulators complete
    for (int sub = 0; sub < h.m_sub; ++sub) {
      const int m = m0 + sub * 128 + warp * 32 + lane;
      const int i = m / h.Wf, j = m - i * h.Wf;
      const int jg = strip_x + j;
      const bool rvalid = (i < d.OHs) & (j < h.TW) & (jg < d.OWs);
      const size_t ypix = (static_cast<size_t>(n_img) * d.OH + (d.o_ph + i * d.o_step)) * d.OW + (d.o_pw + jg * d.o_step);
      const uint32_t trow = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + sub * d.n_tile;
      if (!d.y_is_f32 && !d.accumulate && !(p.dbg_mode & 4)) {
        // all MMAs have completed (accum barrier), every fill of this CTA is consumed: the halo buffers are free and
        // serve as the staging area of the coalesced store (4 KB per warp)
        long long cyc[3] = {0, 0, 0};
        const bool rec = p.dbg != nullptr && threadIdx.x == 0 && blockIdx.y == 0 && blockIdx.x < kDbgCtas;
        epilogue_rows_bf16(trow, d.n_tile, tile_n * d.n_tile, p.n_store, d.n_rows, p.bias, d.act, rvalid,
                           static_cast<uint32_t>(ypix), reinterpret_cast<__nv_bfloat16*>(p.y), d.ldy, d.y_coff,
                           a_base + warp * 4096, reinterpret_cast<uint32_t*>(smem + 512) + warp * 32, lane, p.dbg_mode,
                           rec ? cyc : nullptr, p.st.sums != nullptr ? stat_s : nullptr);
        if (rec)
          for (int q = 0; q < 3; ++q) p.dbg[blockIdx.x * kDbgSlots + 8 + q] += static_cast<unsigned long long>(cyc[q]);
        continue;
      }
      for (int cc = 0; cc < d.n_tile / 16; ++cc) {
        float acc[16];
        if (p.dbg_mode & 2) {
#pragma unroll
          for (int e = 0; e < 16; ++e) acc[e] = 0.f;
        } else {
          tmem_ld16(trow + cc * 16, acc);
        }
        const int col0 = tile_n * d.n_tile + cc * 16;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int col = col0 + g * 8;
          if (!rvalid || col >= p.n_store || (p.dbg_mode & 1)) continue;
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
    }
    if (p.st.sums != nullptr) {   // fused statistics of the conv + norm blocks: one atomic per column and CTA
      named_bar_sync(1, 128);
      flush_epilogue_stats(stat_s, d.n_tile, tile_n * d.n_tile, p.n_store, p.st, n_img, threadIdx.x, 128);
    }
  } else if (warp == 4) {
    // ---------------------------------------------------------------- MMA issuer
    // The whole warp walks the loop (barrier waits, ring bookkeeping in registers), one elected lane issues.  Per
    // instruction only the low descriptor words change, by one add each (common.cuh: umma_bf16_lh).
    const uint32_t hi = sw128_desc_hi(1024);
    const uint32_t b_lo0 = sw128_desc_lo(smem_u32(b_base), 16);
    const uint32_t b_step = static_cast<uint32_t>(b_bytes) >> 4;
    const uint32_t a_step = static_cast<uint32_t>(p.halo_bytes) >> 4;
    const uint32_t a_lo0 = sw128_desc_lo(smem_u32(a_base), 16);
    const uint32_t idesc = p.idesc;
    const uint32_t n_tile = d.n_tile;
    const uint32_t m_sub = h.m_sub;
    uint32_t st = 0, phb = 0, b_lo = b_lo0;     // weight ring position
    uint32_t buf = 0, pha = 0, a_lo = a_lo0;    // halo buffer position
    uint32_t acc = 0;                           // 0 only for the first instruction into each accumulator
    for (int c = 0; c < h.n_chunks; ++c) {
      const int4 ch = s_chunks[c];              // cu0, n_units, first_step, n_steps
      // a chunk that uses only n_units of its 8 channel units needs only ceil(n_units / 2) of the four K=16
      // MMAs (the remaining columns of the halo rows and of the weight tile are zero)
      const int kmax = (ch.y + 1) >> 1;
      mbar_wait(&a_full[buf], pha);
      tcgen05_fence_after();
      if (c == 0 && lane == 0) DBG_STAMP(2);    // MMA warp sees the first chunk
      const int s_end = ch.z + ch.w;
      uint32_t aoff = s_aoff[ch.z];
      for (int s = ch.z; s < s_end; ++s) {
        const uint32_t a_cur = a_lo + aoff;
        if (s + 1 < s_end) aoff = s_aoff[s + 1];   // next step's window offset is in flight during this step's wait
        mbar_wait(&b_full[st], phb);
        tcgen05_fence_after();
        if (elect_one()) {
          // every 128-row sub-tile (halo rows + 128 * sub = 16 KB further) has its own accumulator columns and shares
          // the step's weight tile
          for (uint32_t sub = 0; sub < m_sub; ++sub) {
            const uint32_t a_s = a_cur + sub * 1024u, t_s = tmem_base + sub * n_tile;
            umma_bf16_lh(t_s, a_s, hi, b_lo, hi, idesc, acc);
            if (kmax > 1) umma_bf16_lh(t_s, a_s + 2, hi, b_lo + 2, hi, idesc, 1u);
            if (kmax > 2) umma_bf16_lh(t_s, a_s + 4, hi, b_lo + 4, hi, idesc, 1u);
            if (kmax > 3) umma_bf16_lh(t_s, a_s + 6, hi, b_lo + 6, hi, idesc, 1u);
          }
          umma_commit(&b_empty[st]);
        }
        __syncwarp();
        acc = 1u;
        if (++st == static_cast<uint32_t>(p.b_stages)) {
          st = 0;
          phb ^= 1u;
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
    if (elect_one()) umma_commit(accum);
    __syncwarp();
    if (lane == 0) DBG_STAMP(3);   // last MMA issued
  } else {
    // ---------------------------------------------------------------- weight loader
    if (lane == 0) {
      const uint8_t* src = p.wpk + static_cast<size_t>(tile_n) * h.n_steps * b_bytes;
      for (int sg = 0; sg < h.n_steps; ++sg) {
        const int st = sg % p.b_stages;
        const uint32_t phb = (sg / p.b_stages) & 1;
        mbar_wait(&b_empty[st], phb ^ 1);
        mbar_arrive_expect_tx(&b_full[st], b_bytes);
        bulk_g2s(b_base + static_cast<size_t>(st) * b_bytes, src + static_cast<size_t>(sg) * b_bytes, b_bytes, &b_full[st]);
      }
    }
    __syncwarp();
  }

  if (threadIdx.x == 0) DBG_STAMP(5);   // epilogue of warp 0 done
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc_dyn(tmem_base, p.tmem_cols);
  if (threadIdx.x == 0) DBG_STAMP(6);
}

int init_halo_attributes() {
  const cudaError_t e =
      cudaFuncSetAttribute(igemm_halo_fprop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(halo): %s", cudaGetErrorString(e));
    return CATB_ERR_CUDA;
  }
  return CATB_OK;
}

}  // namespace catb

using namespace catb;

// Shared-memory plan: returns 0 and fills a_bufs / b_stages, or -1 when the halo does not fit.
// step offsets (4 B) + chunk descriptors (16 B), rounded so that the tiles behind them stay 1024-byte aligned
static int halo_table_bytes(int n_steps, int n_chunks) {
  return (((n_steps * 4 + 15) & ~15) + n_chunks * 16 + 1023) / 1024 * 1024;
}

static int halo_smem_plan(int halo_bytes, int b_bytes, int tab_bytes, int* a_bufs, int* b_stages, size_t* total) {
  const int limit = 227 * 1024 - 1024 /*alignment slack*/ - kHHeader - tab_bytes - kHStatBytes;
  int ab = 2;
  if (2 * halo_bytes + 3 * b_bytes > limit) ab = 1;
  int bs = (limit - ab * halo_bytes) / b_bytes;
  if (bs > kHMaxBStages) bs = kHMaxBStages;
  if (bs < 2) return -1;
  *a_bufs = ab;
  *b_stages = bs;
  *total = 1024 + kHHeader + tab_bytes + kHStatBytes + static_cast<size_t>(ab) * halo_bytes + static_cast<size_t>(bs) * b_bytes;
  return 0;
}

extern "C" int catb_igemm_halo_fits(int n_planes, int Lh, int n_tile, int m_sub, int n_steps, int n_chunks) {
  const int halo_bytes = (n_planes * Lh * 128 + 1023) / 1024 * 1024;
  int ab, bs;
  size_t total;
  if (m_sub * n_tile > 512) return 0;
  return halo_smem_plan(halo_bytes, n_tile * 128, halo_table_bytes(n_steps, n_chunks), &ab, &bs, &total) == 0 ? 1 : 0;
}

extern "C" int catb_igemm_halo_fprop(const catb_igemm_desc* d, const catb_halo_desc* h, const catb_halo_step* steps,
                                     const catb_halo_chunk* chunks, const void* x, const void* packed_w,
                                     const float* bias, void* y, const catb_epilogue_stats* stats, catb_stream_t s) {
  CATB_REQUIRE(d != nullptr && h != nullptr, "null descriptor");
  CATB_REQUIRE(d->n_tile % 16 == 0 && d->n_tile >= 16 && d->n_tile <= 256, "n_tile must be a multiple of 16 in [16,256]");
  CATB_REQUIRE(h->m_sub >= 1 && h->m_sub <= 4 && h->m_sub * d->n_tile <= 512, "m_sub * n_tile must fit 512 TMEM columns");
  CATB_REQUIRE(h->n_planes >= 1 && h->n_planes <= 4 && h->n_steps > 0 && h->n_chunks > 0, "bad halo plan");
  CATB_REQUIRE(h->TW > 0 && h->n_strips == (d->OWs + h->TW - 1) / h->TW && h->Wf == h->TW + h->Xmax &&
                   h->Lh == 128 * h->m_sub + h->Ymax * h->Wf + h->Xmax,
               "inconsistent halo geometry");
  CATB_REQUIRE(d->n_units == h->n_steps * 8, "unit table must hold 8 units per step");
  CATB_REQUIRE(d->ldx % 8 == 0 && d->x_coff % 8 == 0 && d->ldy % 8 == 0 && d->y_coff % 8 == 0, "pitches must be multiples of 8");
  HaloParams p;
  p.d = *d;
  p.h = *h;
  p.steps = steps;
  p.chunks = chunks;
  p.x = static_cast<const __nv_bfloat16*>(x);
  p.wpk = static_cast<const uint8_t*>(packed_w);
  p.bias = bias;
  p.y = y;
  p.dbg = g_halo_dbg;
  p.dbg_mode = g_halo_dbg_mode;
  memset(&p.st, 0, sizeof(p.st));
  if (stats != nullptr && stats->sums != nullptr) {
    CATB_REQUIRE(!d->y_is_f32 && !d->accumulate && !(g_halo_dbg_mode & 4), "fused statistics need the plain bf16 store");
    CATB_REQUIRE(stats->C > 0 && stats->coff >= 0 && stats->coff + (d->n_rows + 7) / 8 * 8 <= stats->C, "bad statistics slice");
    p.st = *stats;
  }
  p.halo_bytes = (h->n_planes * h->Lh * 128 + 1023) / 1024 * 1024;
  p.tab_bytes = halo_table_bytes(h->n_steps, h->n_chunks);
  size_t smem = 0;
  CATB_REQUIRE(halo_smem_plan(p.halo_bytes, d->n_tile * 128, p.tab_bytes, &p.a_bufs, &p.b_stages, &smem) == 0,
               "halo tile (%d bytes) does not fit in shared memory", p.halo_bytes);
  // Ask only for what this launch can use, so that small problems co-schedule several CTAs per SM
  // (the kernel is not persistent: prologue / epilogue of one CTA overlap the MMAs of its neighbours).
  if (p.a_bufs > h->n_chunks) p.a_bufs = h->n_chunks;
  if (p.b_stages > h->n_steps) p.b_stages = h->n_steps;
  {
    // keep >= ~64 KB of weight tiles in flight (bulk-copy latency ~1.5 us), but not more stages than that needs
    const int b_bytes = d->n_tile * 128;
    const int budget = h->b_budget > 0 ? h->b_budget : 64 * 1024;
    int want = (budget + b_bytes - 1) / b_bytes;
    if (want < (h->b_budget > 0 ? 2 : 4)) want = h->b_budget > 0 ? 2 : 4;
    if (p.b_stages > want) p.b_stages = want;
  }
  smem = 1024 + kHHeader + p.tab_bytes + kHStatBytes + static_cast<size_t>(p.a_bufs) * p.halo_bytes +
         static_cast<size_t>(p.b_stages) * d->n_tile * 128;
  const int positions = d->OHs * h->Wf;
  p.tiles_per_image = (positions + 128 * h->m_sub - 1) / (128 * h->m_sub);
  uint32_t cols = 32;
  while (static_cast<int>(cols) < h->m_sub * d->n_tile) cols <<= 1;
  p.tmem_cols = cols;
  p.n_store = (d->n_rows + 7) / 8 * 8;
  p.idesc = make_idesc_bf16(128, d->n_tile, 0, 0);
  const int n_tiles = (d->n_rows + d->n_tile - 1) / d->n_tile;
  dim3 grid(p.tiles_per_image * h->n_strips * d->N, n_tiles, 1);
  igemm_halo_fprop_kernel<<<grid, kHThreads, smem, static_cast<cudaStream_t>(s)>>>(p);
  return check_launch("igemm_halo_fprop");
}

// Development aid (tools/profile_gemm.py --timeline): when a device buffer of 4096 x 8 uint64 is registered, every
// halo-fprop CTA with blockIdx.x < 4096 records %globaltimer at its phase boundaries.  Pass NULL to switch it off.
extern "C" int catb_debug_timeline(void* device_buffer) {
  g_halo_dbg = static_cast<unsigned long long*>(device_buffer);
  return CATB_OK;
}

extern "C" int catb_debug_mode(int mode) {
  g_halo_dbg_mode = mode;
  return CATB_OK;
}
