// Depthwise convolution (per 8-channel unit kernel size k in {1,3,5,7}, padding (k-1)/2, reflect or zero), forward,
// input gradient and weight gradient: models/modules/inception_modules.py:165-173 (ConvBNReLU(groups=midp)) and the
// zero-padded ConvSyncBNReLU(groups=midp) of the SPADE blocks (:441-452, :700-712).
//
// Round-1 version: one thread per (pixel, unit) with the kernel size read per thread.  Neighbouring lanes hold units of
// different k (the dw branches of a block sit side by side in the mid buffer), so a warp ran the 5x5 loop with 2/3 of its
// lanes idle, every tap rebuilt its address from scratch, and the weight gradient re-read both tensors once per tap:
// 300-380 GB/s forward, 125 GB/s weight gradient (profiles/r02_bench_v2_1gpu.json, hbm_kernels).
//
// Here the units are grouped by kernel size (contiguous runs, found once per block), the kernels are instantiated per K
// (taps fully unrolled, K uniform per warp), a thread produces FOUR horizontally adjacent pixels of one unit (a filter
// row of K+3 input vectors serves 4 x K products, the K x K x 8 filter is read once per four pixels), and the weight
// gradient stages a 16 x 16 pixel tile (+ halo) of both tensors in shared memory and gives every (tap, channel) pair its
// own thread, so each tensor is read once.
#include "common.cuh"

namespace catb {

constexpr int kDwMaxTaps = 49;
constexpr int kDwMaxGroups = 8;

struct DwGroups {
  int n;
  int u0[kDwMaxGroups], nu[kDwMaxGroups], k[kDwMaxGroups];
};

// contiguous runs of units with the same kernel size (thread 0), then the filters as w_s[unit][tap][8] (zero padded)
__device__ __forceinline__ void dw_prepare(DwGroups* grp, float* w_s, int C, const int32_t* __restrict__ ksize,
                                           const int32_t* __restrict__ w_off, const float* __restrict__ arena) {
  const int U = C / 8;
  if (threadIdx.x == 0) {
    int n = 0;
    for (int u = 0; u < U; ++u) {
      const int k = ksize[u * 8];
      if (n > 0 && grp->k[n - 1] == k && n <= kDwMaxGroups) {
        ++grp->nu[n - 1];
      } else if (n < kDwMaxGroups) {
        grp->u0[n] = u;
        grp->nu[n] = 1;
        grp->k[n] = k;
        ++n;
      } else {   // more runs than slots: fold into the last group only if k matches (checked on the host)
        ++grp->nu[n - 1];
      }
    }
    grp->n = n;
  }
  if (w_s != nullptr) {
    // one thread per channel: its kernel size and arena offset are loaded ONCE, then the (up to 49) taps are independent
    // loads (the element-wise form -- three dependent global loads for each of the C * 49 table entries -- made this
    // prologue the whole kernel: ~40 us per launch whatever the tensor size, profiles/r02_bench_v2_1gpu.json hbm_kernels)
    for (int c = threadIdx.x; c < U * 8; c += blockDim.x) {
      const int k = ksize[c], wo = w_off[c];
      const int taps = wo >= 0 ? k * k : 0;
      float* dst = w_s + (static_cast<size_t>(c >> 3) * kDwMaxTaps) * 8 + (c & 7);
#pragma unroll 7
      for (int tap = 0; tap < kDwMaxTaps; ++tap) dst[tap * 8] = tap < taps ? __ldg(arena + wo + tap) : 0.f;
    }
  }
  __syncthreads();
}

__device__ __forceinline__ int dw_coord(int i, int L, bool zero_pad, bool& ok) {
  if (zero_pad) {
    ok = (i >= 0) & (i < L);
    return ok ? i : 0;
  }
  ok = true;
  return reflect_idx(i, L);
}

// ---- forward: y[h, w0 .. w0+PX-1] of one unit ---------------------------------------------------
template <int K, int PX>
__device__ __forceinline__ void dw_fwd_quad(const __nv_bfloat16* __restrict__ x, int ldx, __nv_bfloat16* __restrict__ y, int ldy,
                                            int n, int h, int w0, int H, int W, const float* __restrict__ wu, bool zero_pad) {
  constexpr int P = (K - 1) / 2;
  f8 acc[PX];
#pragma unroll
  for (int j = 0; j < PX; ++j)
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[j].v[q] = 0.f;
  const size_t img = static_cast<size_t>(n) * H * W;
#pragma unroll
  for (int r = 0; r < K; ++r) {
    bool rok;
    const int ih = dw_coord(h - P + r, H, zero_pad, rok);
    uint4 xv[K + PX - 1];
#pragma unroll
    for (int c = 0; c < K + PX - 1; ++c) {
      bool cok;
      const int iw = dw_coord(w0 - P + c, W, zero_pad, cok);
      xv[c] = (rok & cok) ? ldg16(x + (img + static_cast<size_t>(ih) * W + iw) * ldx) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int s = 0; s < K; ++s) {
      const float4 wa = *reinterpret_cast<const float4*>(wu + (r * K + s) * 8);
      const float4 wb = *reinterpret_cast<const float4*>(wu + (r * K + s) * 8 + 4);
#pragma unroll
      for (int j = 0; j < PX; ++j) {
        const f8 v = unpack8(xv[j + s]);
        acc[j].v[0] += v.v[0] * wa.x; acc[j].v[1] += v.v[1] * wa.y; acc[j].v[2] += v.v[2] * wa.z; acc[j].v[3] += v.v[3] * wa.w;
        acc[j].v[4] += v.v[4] * wb.x; acc[j].v[5] += v.v[5] * wb.y; acc[j].v[6] += v.v[6] * wb.z; acc[j].v[7] += v.v[7] * wb.w;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < PX; ++j)
    if (w0 + j < W) st16(y + (img + static_cast<size_t>(h) * W + w0 + j) * ldy, pack8(acc[j]));
}

__device__ __noinline__ void dw_fwd_pair(int k, const __nv_bfloat16* __restrict__ x, int ldx, __nv_bfloat16* __restrict__ y, int ldy,
                                         int n, int h, int w0, int H, int W, const float* __restrict__ wu, bool zero_pad) {
  if (k == 5)
    dw_fwd_quad<5, 2>(x, ldx, y, ldy, n, h, w0, H, W, wu, zero_pad);
  else
    dw_fwd_quad<7, 2>(x, ldx, y, ldy, n, h, w0, H, W, wu, zero_pad);
}

// ---- input gradient: dx[h, w] of one unit --------------------------------------------------------
// dx[ih, iw] = sum over (oh, r), (ow, s) with pad(oh - p + r) = ih, pad(ow - p + s) = iw of dy[oh, ow] * w[r, s]
template <int K>
__device__ __forceinline__ void dw_bwd_data_px(const __nv_bfloat16* __restrict__ dy, int ldy, __nv_bfloat16* __restrict__ dx, int ldx,
                                               int n, int h, int w, int H, int W, const float* __restrict__ wu, bool zero_pad) {
  constexpr int P = (K - 1) / 2;
  f8 acc;
#pragma unroll
  for (int q = 0; q < 8; ++q) acc.v[q] = 0.f;
  const size_t img = static_cast<size_t>(n) * H * W;
  const bool interior = zero_pad || (h > P && h < H - 1 - P && w > P && w < W - 1 - P);
  if (interior) {   // only the pixel itself maps onto (h, w): a plain correlation with the flipped filter
#pragma unroll
    for (int r = 0; r < K; ++r) {
      const int oh = h + P - r;
      if (oh < 0 || oh >= H) continue;
#pragma unroll
      for (int s = 0; s < K; ++s) {
        const int ow = w + P - s;
        if (ow < 0 || ow >= W) continue;
        const f8 g = unpack8(ldg16(dy + (img + static_cast<size_t>(oh) * W + ow) * ldy));
        const float* wp = wu + (r * K + s) * 8;
#pragma unroll
        for (int q = 0; q < 8; ++q) acc.v[q] += g.v[q] * wp[q];
      }
    }
  } else {          // border of a reflect-padded conv: mirrored frame positions contribute too
    int hs[3], ws[3];
    const int nh = reflect_sources(h, H, P, hs), nw = reflect_sources(w, W, P, ws);
    for (int a = 0; a < nh; ++a)
      for (int r = 0; r < K; ++r) {
        const int oh = hs[a] + P - r;
        if (oh < 0 || oh >= H) continue;
        for (int b = 0; b < nw; ++b)
          for (int s = 0; s < K; ++s) {
            const int ow = ws[b] + P - s;
            if (ow < 0 || ow >= W) continue;
            const f8 g = unpack8(ldg16(dy + (img + static_cast<size_t>(oh) * W + ow) * ldy));
            const float* wp = wu + (r * K + s) * 8;
#pragma unroll
            for (int q = 0; q < 8; ++q) acc.v[q] += g.v[q] * wp[q];
          }
      }
  }
  st16(dx + (img + static_cast<size_t>(h) * W + w) * ldx, pack8(acc));
}

// MODE 0: forward (four pixels per thread), MODE 1: input gradient (one pixel per thread)
template <int MODE>
__global__ void __launch_bounds__(256, 2) dwconv_grouped_kernel(const __nv_bfloat16* __restrict__ src, int lds, int s_coff,
                                                              __nv_bfloat16* __restrict__ dst, int ldd, int d_coff, int N, int H,
                                                              int W, int C, const int32_t* __restrict__ ksize,
                                                              const int32_t* __restrict__ w_off, const float* __restrict__ arena,
                                                              int zero_pad) {
  extern __shared__ float w_s[];   // [U][49][8]
  __shared__ DwGroups grp;
  dw_prepare(&grp, w_s, C, ksize, w_off, arena);
  const int PX = MODE == 0 ? 4 : 1;
  const int Wq = (W + PX - 1) / PX;
  const long long cells = static_cast<long long>(N) * H * Wq;   // (n, h, quad) cells
  for (int g = 0; g < grp.n; ++g) {
    const int u0 = grp.u0[g], nu = grp.nu[g], k = grp.k[g];
    const long long total = cells * nu;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
      const int ul = static_cast<int>(idx % nu);
      const long long cell = idx / nu;
      const int wq = static_cast<int>(cell % Wq);
      const int h = static_cast<int>((cell / Wq) % H);
      const int n = static_cast<int>(cell / (static_cast<long long>(Wq) * H));
      const int u = u0 + ul;
      const float* wu = w_s + static_cast<size_t>(u) * kDwMaxTaps * 8;
      const __nv_bfloat16* sp = src + s_coff + u * 8;
      __nv_bfloat16* dp = dst + d_coff + u * 8;
      if (MODE == 0) {
        switch (k) {
          case 1: dw_fwd_quad<1, 4>(sp, lds, dp, ldd, n, h, wq * 4, H, W, wu, zero_pad); break;
          case 3: dw_fwd_quad<3, 4>(sp, lds, dp, ldd, n, h, wq * 4, H, W, wu, zero_pad); break;
          default:     // 5x5 / 7x7: two pixels at a time (register footprint)
            dw_fwd_pair(k, sp, lds, dp, ldd, n, h, wq * 4, H, W, wu, zero_pad);
            if (wq * 4 + 2 < W) dw_fwd_pair(k, sp, lds, dp, ldd, n, h, wq * 4 + 2, H, W, wu, zero_pad);
            break;
        }
      } else {
        switch (k) {
          case 1: dw_bwd_data_px<1>(sp, lds, dp, ldd, n, h, wq, H, W, wu, zero_pad); break;
          case 3: dw_bwd_data_px<3>(sp, lds, dp, ldd, n, h, wq, H, W, wu, zero_pad); break;
          case 5: dw_bwd_data_px<5>(sp, lds, dp, ldd, n, h, wq, H, W, wu, zero_pad); break;
          default: dw_bwd_data_px<7>(sp, lds, dp, ldd, n, h, wq, H, W, wu, zero_pad); break;
        }
      }
    }
  }
}

// ---- weight gradient -----------------------------------------------------------------------------
// grid = (tiles of 16 x 16 output pixels over all images, units).  The block stages dy[16x16][8] and
// x[(16+2P) x (16+2P)][8] (padding resolved while staging) as fp32 in shared memory; thread t < k*k*8 owns
// (tap = t / 8, channel = t % 8) and walks the 256 pixels; one atomic add per (tap, channel) and block.
constexpr int kDwTile = 16;
__global__ void __launch_bounds__(256) dwconv_bwd_weight_tiled_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int x_coff,
                                                                       const __nv_bfloat16* __restrict__ dy, int ldy, int y_coff,
                                                                       int N, int H, int W, const int32_t* __restrict__ ksize,
                                                                       const int32_t* __restrict__ w_off, float* __restrict__ grad,
                                                                       int zero_pad) {
  constexpr int XT = kDwTile + 6;                 // widest halo (k = 7)
  __shared__ float dy_s[kDwTile * kDwTile][8];
  __shared__ float x_s[XT * XT][8];
  const int u = blockIdx.y;
  const int k = ksize[u * 8];
  const int p = (k - 1) / 2;
  const int tiles_w = (W + kDwTile - 1) / kDwTile, tiles_h = (H + kDwTile - 1) / kDwTile;
  const int tile = blockIdx.x;
  const int n = tile / (tiles_w * tiles_h);
  const int th = (tile / tiles_w) % tiles_h, tw = tile % tiles_w;
  const int h0 = th * kDwTile, w0 = tw * kDwTile;
  const size_t img = static_cast<size_t>(n) * H * W;
  const int xt = kDwTile + 2 * p;
  // stage dy (zero outside the image)
  {
    const int py = threadIdx.x / kDwTile, px = threadIdx.x % kDwTile;
    const int h = h0 + py, w = w0 + px;
    f8 v;
#pragma unroll
    for (int q = 0; q < 8; ++q) v.v[q] = 0.f;
    if (h < H && w < W) v = unpack8(ldg16(dy + (img + static_cast<size_t>(h) * W + w) * ldy + y_coff + u * 8));
    *reinterpret_cast<float4*>(&dy_s[threadIdx.x][0]) = make_float4(v.v[0], v.v[1], v.v[2], v.v[3]);
    *reinterpret_cast<float4*>(&dy_s[threadIdx.x][4]) = make_float4(v.v[4], v.v[5], v.v[6], v.v[7]);
  }
  // stage x with its halo (padding rule applied here)
  for (int i = threadIdx.x; i < xt * xt; i += blockDim.x) {
    const int yy = i / xt, xx = i - yy * xt;
    bool rok, cok;
    const int ih = dw_coord(h0 - p + yy, H, zero_pad != 0, rok);
    const int iw = dw_coord(w0 - p + xx, W, zero_pad != 0, cok);
    // rows / columns beyond the image on the far side only meet zero dy rows; keep them finite
    const bool inside = rok & cok & (h0 - p + yy < H + p) & (w0 - p + xx < W + p) & (ih < H) & (iw < W);
    f8 v;
#pragma unroll
    for (int q = 0; q < 8; ++q) v.v[q] = 0.f;
    if (inside) v = unpack8(ldg16(x + (img + static_cast<size_t>(ih) * W + iw) * ldx + x_coff + u * 8));
    *reinterpret_cast<float4*>(&x_s[yy * XT + xx][0]) = make_float4(v.v[0], v.v[1], v.v[2], v.v[3]);
    *reinterpret_cast<float4*>(&x_s[yy * XT + xx][4]) = make_float4(v.v[4], v.v[5], v.v[6], v.v[7]);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < k * k * 8; t += blockDim.x) {
    const int q = t & 7, tap = t >> 3;
    const int r = tap / k, s = tap - r * k;
    float acc = 0.f;
#pragma unroll 4
    for (int py = 0; py < kDwTile; ++py) {
      const float* xr = &x_s[(py + r) * XT + s][q];
      const float* dr = &dy_s[py * kDwTile][q];
#pragma unroll
      for (int px = 0; px < kDwTile; ++px) acc += dr[px * 8] * xr[px * 8];
    }
    const int wo = w_off[u * 8 + q];
    if (wo >= 0) atomicAdd(grad + wo + tap, acc);
  }
}

int init_dwconv_attributes() {
  cudaError_t e = cudaFuncSetAttribute(dwconv_grouped_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(dwconv_grouped_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(dwconv): %s", cudaGetErrorString(e));
    return CATB_ERR_CUDA;
  }
  return CATB_OK;
}

}  // namespace catb

using namespace catb;

#define CHK_SLICE(ld, coff, C)                                                                     \
  CATB_REQUIRE((ld) % 8 == 0 && (coff) % 8 == 0 && (C) % 8 == 0 && (C) > 0 && (coff) + (C) <= (ld), \
               "bad channel slice (ld=%d coff=%d C=%d)", (int)(ld), (int)(coff), (int)(C))

static int dw_grid(long long work, int cap) {
  long long g = (work + 255) / 256;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return static_cast<int>(g);
}

extern "C" int catb_dwconv_fwd(const void* x, int ldx, int x_coff, void* y, int ldy, int y_coff, int N, int H, int W,
                               int C, const int32_t* ksize, const int32_t* w_off, const float* arena, int pad_mode,
                               catb_stream_t s) {
  CHK_SLICE(ldx, x_coff, C);
  CHK_SLICE(ldy, y_coff, C);
  const size_t smem = static_cast<size_t>(C) * kDwMaxTaps * sizeof(float);
  CATB_REQUIRE(smem <= 200 * 1024, "too many depthwise channels (%d) for the shared-memory filter stage", C);
  const long long work = static_cast<long long>(N) * H * ((W + 3) / 4) * (C / 8);
  dwconv_grouped_kernel<0><<<dw_grid(work, 148 * 4), 256, smem, static_cast<cudaStream_t>(s)>>>(
      static_cast<const __nv_bfloat16*>(x), ldx, x_coff, static_cast<__nv_bfloat16*>(y), ldy, y_coff, N, H, W, C, ksize, w_off,
      arena, pad_mode == CATB_PAD_ZERO);
  return check_launch("dwconv_fwd");
}

extern "C" int catb_dwconv_bwd_data(const void* dy, int ldy, int y_coff, void* dx, int ldx, int x_coff, int N, int H,
                                    int W, int C, const int32_t* ksize, const int32_t* w_off, const float* arena,
                                    int pad_mode, catb_stream_t s) {
  CHK_SLICE(ldx, x_coff, C);
  CHK_SLICE(ldy, y_coff, C);
  const size_t smem = static_cast<size_t>(C) * kDwMaxTaps * sizeof(float);
  CATB_REQUIRE(smem <= 200 * 1024, "too many depthwise channels (%d) for the shared-memory filter stage", C);
  const long long work = static_cast<long long>(N) * H * W * (C / 8);
  dwconv_grouped_kernel<1><<<dw_grid(work, 148 * 4), 256, smem, static_cast<cudaStream_t>(s)>>>(
      static_cast<const __nv_bfloat16*>(dy), ldy, y_coff, static_cast<__nv_bfloat16*>(dx), ldx, x_coff, N, H, W, C, ksize,
      w_off, arena, pad_mode == CATB_PAD_ZERO);
  return check_launch("dwconv_bwd_data");
}

extern "C" int catb_dwconv_bwd_weight(const void* x, int ldx, int x_coff, const void* dy, int ldy, int y_coff, int N,
                                      int H, int W, int C, const int32_t* ksize, const int32_t* w_off,
                                      float* arena_grad, int pad_mode, catb_stream_t s) {
  CHK_SLICE(ldx, x_coff, C);
  CHK_SLICE(ldy, y_coff, C);
  const long long tiles = static_cast<long long>(N) * ((H + kDwTile - 1) / kDwTile) * ((W + kDwTile - 1) / kDwTile);
  CATB_REQUIRE(tiles < (1ll << 31) && C / 8 <= 65535, "depthwise weight gradient grid too large");
  dim3 grid(static_cast<unsigned>(tiles), C / 8, 1);
  dwconv_bwd_weight_tiled_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(s)>>>(
      static_cast<const __nv_bfloat16*>(x), ldx, x_coff, static_cast<const __nv_bfloat16*>(dy), ldy, y_coff, N, H, W, ksize,
      w_off, arena_grad, pad_mode == CATB_PAD_ZERO);
  return check_launch("dwconv_bwd_weight");
}
