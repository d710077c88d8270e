// CUDA-core kernels of the CAT distillation path: depthwise convolution, InstanceNorm/BatchNorm
// (statistics, apply, backward), layout conversion, reflect-pad adjoint, GAN / L1 / KA losses, Adam.
// All activations are NHWC bf16 slices (pixel pitch `ld`, channel offset `coff`, 8 channels = 16 bytes
// per access); reductions use warp shuffles + shared-memory partials + one global atomic per block.
#include <algorithm>
#include <cstring>

#include "common.cuh"

namespace catb {

static inline cudaStream_t S(catb_stream_t s) { return static_cast<cudaStream_t>(s); }
static inline int grid_for(long long work, int block, int max_blocks = 148 * 16) {
  long long g = (work + block - 1) / block;
  if (g < 1) g = 1;
  if (g > max_blocks) g = max_blocks;
  return static_cast<int>(g);
}

// ------------------------------------------------------------------------------------------------
// normalisation
// ------------------------------------------------------------------------------------------------
// Generic per-channel reduction skeleton: thread <-> (pixel lane, unit); block partials in smem.
// MODE 0: sum x, sum x^2            (forward statistics)
// MODE 1: sum dz, sum dz*xhat        (backward reduction)
template <int MODE>
__global__ void __launch_bounds__(512, 1) channel_reduce_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int x_coff,
                                      const __nv_bfloat16* __restrict__ dout, int ldd, int d_coff,
                                      const __nv_bfloat16* __restrict__ out, int ldo, int o_coff, int HW, int C,
                                      int per_sample, const float* __restrict__ mean_rstd, int act,
                                      float* __restrict__ sums, long long pixels_total) {
  extern __shared__ float sh[];  // [2*C]
  const int U = C / 8;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  const int g = per_sample ? blockIdx.y : 0;
  const long long pix0 = per_sample ? static_cast<long long>(g) * HW : 0;
  const long long npix = per_sample ? HW : pixels_total;
  const int lanes = U >= static_cast<int>(blockDim.x) ? 1 : blockDim.x / U;  // pixel lanes per block
  const int u_first = threadIdx.x % U;
  const int pl = threadIdx.x / U;
  const bool active = (U >= static_cast<int>(blockDim.x)) || pl < lanes;
  if (active) {
    for (int u = u_first; u < U; u += blockDim.x) {  // only loops when U > blockDim.x
      f8 a, b;
#pragma unroll
      for (int q = 0; q < 8; ++q) a.v[q] = b.v[q] = 0.f;
      f8 mu, rs;
      if (MODE == 1) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          mu.v[q] = mean_rstd[(static_cast<size_t>(g) * 2 + 0) * C + u * 8 + q];
          rs.v[q] = mean_rstd[(static_cast<size_t>(g) * 2 + 1) * C + u * 8 + q];
        }
      }
      // four pixels in flight per thread: all loads of an iteration are issued before the first use (at ~40 % occupancy
      // one outstanding 16-byte load per thread left the memory system two thirds idle)
      constexpr int kIn = MODE == 0 ? 4 : 2;      // MODE 1 streams three tensors per pixel
      const long long pstep = static_cast<long long>(gridDim.x) * lanes;
      for (long long pp = static_cast<long long>(blockIdx.x) * lanes + (U >= static_cast<int>(blockDim.x) ? 0 : pl);
           pp < npix; pp += kIn * pstep) {
        uint4 xr[kIn], dr[kIn], orr[kIn];
        bool ok[kIn];
#pragma unroll
        for (int j = 0; j < kIn; ++j) {
          const long long pj = pp + j * pstep;
          ok[j] = pj < npix;
          const size_t pix = static_cast<size_t>(pix0 + (ok[j] ? pj : pp));
          xr[j] = ldg16(x + pix * ldx + x_coff + u * 8);
          if (MODE == 1) {
            dr[j] = ldg16(dout + pix * ldd + d_coff + u * 8);
            if (act != CATB_ACT_NONE) orr[j] = ldg16(out + pix * ldo + o_coff + u * 8);
          }
        }
#pragma unroll
        for (int j = 0; j < kIn; ++j) {
          if (!ok[j]) continue;
          const f8 xv = unpack8(xr[j]);
          if (MODE == 0) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              a.v[q] += xv.v[q];
              b.v[q] += xv.v[q] * xv.v[q];
            }
          } else {
            const f8 dv = unpack8(dr[j]);
            f8 ov;
            if (act != CATB_ACT_NONE) ov = unpack8(orr[j]);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float dz = act != CATB_ACT_NONE ? dv.v[q] * act_grad_from_out(ov.v[q], act) : dv.v[q];
              a.v[q] += dz;
              b.v[q] += dz * (xv.v[q] - mu.v[q]) * rs.v[q];
            }
          }
        }
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        atomicAdd(&sh[u * 8 + q], a.v[q]);
        atomicAdd(&sh[C + u * 8 + q], b.v[q]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(sums + static_cast<size_t>(g) * 2 * C + i, sh[i]);
}

__global__ void norm_finalize_kernel(const float* __restrict__ sums, int G, int C, float count, float eps,
                                     float momentum, const float* gamma, const float* beta, float* running_mean,
                                     float* running_var, float* scale, float* shift, float* mean_rstd) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= G * C) return;
  const int g = idx / C, c = idx % C;
  float mean, var;
  if (sums != nullptr) {
    const double sm = sums[static_cast<size_t>(g) * 2 * C + c], sq = sums[static_cast<size_t>(g) * 2 * C + C + c];
    const double m = sm / count;
    double v = sq / count - m * m;
    if (v < 0) v = 0;
    mean = static_cast<float>(m);
    var = static_cast<float>(v);
    if (running_mean != nullptr && G == 1) {
      const float unbiased = count > 1.f ? var * count / (count - 1.f) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
    }
  } else {
    mean = running_mean[c];
    var = running_var[c];
  }
  const float rstd = rsqrtf(var + eps);
  const float ga = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f;
  scale[idx] = ga * rstd;
  shift[idx] = be - mean * ga * rstd;
  if (mean_rstd) {
    mean_rstd[(static_cast<size_t>(g) * 2 + 0) * C + c] = mean;
    mean_rstd[(static_cast<size_t>(g) * 2 + 1) * C + c] = rstd;
  }
}

// Element-wise normalisation kernels: a thread owns ONE 8-channel unit (its per-channel constants live in
// registers for the whole kernel) and walks over pixels, `lanes` = blockDim / U pixels per block iteration, two
// pixels in flight per thread; blockIdx.y = sample when the statistics are per sample.  No per-element
// divisions, no per-element constant loads: the loop is 16-byte loads, FMAs and one 16-byte store.
struct UnitMap {
  int u, pl, lanes;
  bool active;
};
__device__ __forceinline__ UnitMap unit_map(int U) {
  UnitMap m;
  m.lanes = U >= static_cast<int>(blockDim.x) ? 1 : static_cast<int>(blockDim.x) / U;
  m.u = threadIdx.x % U;
  m.pl = threadIdx.x / U;
  m.active = (U >= static_cast<int>(blockDim.x)) || m.pl < m.lanes;
  if (U >= static_cast<int>(blockDim.x)) m.pl = 0;
  return m;
}

// Optional "finalize" folded into the apply kernel (fin.sums != nullptr): every thread derives scale / shift of its own
// 8 channels from the accumulated sums (the same arithmetic as norm_finalize_kernel), and the first pixel block of each
// group also writes scale / shift / mean_rstd for the backward pass and moves the running statistics -- one launch less
// per normalisation layer.
struct NormFin {
  const float* sums;   // [G][2][C] or nullptr (scale / shift are read from memory)
  const float* gamma;
  const float* beta;
  float* running_mean;
  float* running_var;
  float* scale_out;
  float* shift_out;
  float* mean_rstd;
  float count, eps, momentum;
};

// Three blocks per SM (<= 80 registers; 93 unconstrained = two blocks): the kernel is a pure stream, and with four pixels in
// flight per thread two blocks keep only ~32-64 KB of loads outstanding per SM, about half of what the HBM latency needs.
__global__ void __launch_bounds__(256, 3) norm_apply_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int x_coff, __nv_bfloat16* __restrict__ y,
                                  int ldy, int y_coff, const __nv_bfloat16* __restrict__ res, int ldr, int r_coff, int HW,
                                  int C, int per_sample, const float* __restrict__ scale, const float* __restrict__ shift,
                                  int act, long long pixels, const NormFin fin) {
  extern __shared__ float fin_s[];   // fused finalize: [2][C] scale / shift of this block's group
  const int U = C / 8;
  const UnitMap m = unit_map(U);
  const int g = per_sample ? blockIdx.y : 0;
  if (fin.sums != nullptr) {
    // one channel per thread, ONCE per block (fp64 only here: E[x^2] - E[x]^2 cancels; per-thread copies of this for all
    // 8 channels of every thread cost more than the whole element-wise pass -- fp64 runs at 1/64 rate)
    const double inv = 1.0 / static_cast<double>(fin.count);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const double mu = fin.sums[static_cast<size_t>(g) * 2 * C + c] * inv;
      double v = fin.sums[static_cast<size_t>(g) * 2 * C + C + c] * inv - mu * mu;
      if (v < 0) v = 0;
      const float mean = static_cast<float>(mu), var = static_cast<float>(v);
      const float rstd = rsqrtf(var + fin.eps);
      const float ga = fin.gamma ? fin.gamma[c] : 1.f, be = fin.beta ? fin.beta[c] : 0.f;
      const float sc1 = ga * rstd, sh1 = be - mean * ga * rstd;
      fin_s[c] = sc1;
      fin_s[C + c] = sh1;
      if (blockIdx.x == 0) {
        if (fin.running_mean != nullptr && !per_sample) {
          const float unbiased = fin.count > 1.f ? var * fin.count / (fin.count - 1.f) : var;
          fin.running_mean[c] = (1.f - fin.momentum) * fin.running_mean[c] + fin.momentum * mean;
          fin.running_var[c] = (1.f - fin.momentum) * fin.running_var[c] + fin.momentum * unbiased;
        }
        fin.scale_out[static_cast<size_t>(g) * C + c] = sc1;
        fin.shift_out[static_cast<size_t>(g) * C + c] = sh1;
        if (fin.mean_rstd != nullptr) {
          fin.mean_rstd[(static_cast<size_t>(g) * 2 + 0) * C + c] = mean;
          fin.mean_rstd[(static_cast<size_t>(g) * 2 + 1) * C + c] = rstd;
        }
      }
    }
    __syncthreads();
  }
  if (!m.active) return;
  const long long pix0 = per_sample ? static_cast<long long>(g) * HW : 0;
  const long long npix = per_sample ? HW : pixels;
  const long long stride = static_cast<long long>(gridDim.x) * m.lanes;
  for (int u = m.u; u < U; u += blockDim.x) {  // only loops when U > blockDim.x
    float sc[8], sh[8];
    if (fin.sums != nullptr) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        sc[q] = fin_s[u * 8 + q];
        sh[q] = fin_s[C + u * 8 + q];
      }
    } else {
    *reinterpret_cast<float4*>(sc) = __ldg(reinterpret_cast<const float4*>(scale + static_cast<size_t>(g) * C + u * 8));
    *reinterpret_cast<float4*>(sc + 4) = __ldg(reinterpret_cast<const float4*>(scale + static_cast<size_t>(g) * C + u * 8 + 4));
    *reinterpret_cast<float4*>(sh) = __ldg(reinterpret_cast<const float4*>(shift + static_cast<size_t>(g) * C + u * 8));
    *reinterpret_cast<float4*>(sh + 4) = __ldg(reinterpret_cast<const float4*>(shift + static_cast<size_t>(g) * C + u * 8 + 4));
    }
    constexpr int kIn = 4;      // pixels in flight per thread (loads issued before the first use)
    for (long long pp = static_cast<long long>(blockIdx.x) * m.lanes + m.pl; pp < npix; pp += kIn * stride) {
      uint4 xr[kIn], rr[kIn];
      bool ok[kIn];
#pragma unroll
      for (int j = 0; j < kIn; ++j) {
        const long long pj = pp + j * stride;
        ok[j] = pj < npix;
        const size_t pj_ = static_cast<size_t>(pix0 + (ok[j] ? pj : pp));
        xr[j] = ldg16(x + pj_ * ldx + x_coff + u * 8);
        if (res != nullptr) rr[j] = ldg16(res + pj_ * ldr + r_coff + u * 8);
      }
#pragma unroll
      for (int j = 0; j < kIn; ++j) {
        if (!ok[j]) continue;
        f8 v = unpack8(xr[j]);
#pragma unroll
        for (int q = 0; q < 8; ++q) v.v[q] = apply_act(v.v[q] * sc[q] + sh[q], act);
        if (res != nullptr) {
          const f8 r = unpack8(rr[j]);
#pragma unroll
          for (int q = 0; q < 8; ++q) v.v[q] += r.v[q];
        }
        st16(y + static_cast<size_t>(pix0 + pp + j * stride) * ldy + y_coff + u * 8, pack8(v));
      }
    }
  }
}

__global__ void norm_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dout, int ldd, int d_coff,
                                      const __nv_bfloat16* __restrict__ out, int ldo, int o_coff,
                                      const __nv_bfloat16* __restrict__ x, int ldx, int x_coff,
                                      __nv_bfloat16* __restrict__ dx, int ldg, int g_coff, int HW, int C, int per_sample,
                                      const float* __restrict__ mean_rstd, const float* __restrict__ gamma,
                                      const float* __restrict__ red, float count, int act, float* dgamma, float* dbeta,
                                      int G, long long pixels) {
  const int U = C / 8;
  if (blockIdx.x == 0 && blockIdx.y == 0 && (dgamma != nullptr || dbeta != nullptr)) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float sb = 0.f, sg = 0.f;
      for (int g = 0; g < G; ++g) {
        sb += red[(static_cast<size_t>(g) * 2 + 0) * C + c];
        sg += red[(static_cast<size_t>(g) * 2 + 1) * C + c];
      }
      if (dbeta) atomicAdd(dbeta + c, sb);
      if (dgamma) atomicAdd(dgamma + c, sg);
    }
  }
  const UnitMap m = unit_map(U);
  if (!m.active) return;
  const int g = per_sample ? blockIdx.y : 0;
  const long long pix0 = per_sample ? static_cast<long long>(g) * HW : 0;
  const long long npix = per_sample ? HW : pixels;
  const long long stride = static_cast<long long>(gridDim.x) * m.lanes;
  const float inv_count = 1.f / count;
  for (int u = m.u; u < U; u += blockDim.x) {
    // dx = k1 * dz + k2 * x + k3 with per-channel constants (xhat = (x - mu) * rs):
    //   k1 = ga*rs,  k2 = -ga*rs*rs*s2/count,  k3 = -ga*rs*s1/count + ga*rs*rs*mu*s2/count
    float k1[8], k2[8], k3[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int c = u * 8 + q;
      const float mu = __ldg(mean_rstd + (static_cast<size_t>(g) * 2 + 0) * C + c);
      const float rs = __ldg(mean_rstd + (static_cast<size_t>(g) * 2 + 1) * C + c);
      const float ga = gamma ? __ldg(gamma + c) : 1.f;
      const float s1 = __ldg(red + (static_cast<size_t>(g) * 2 + 0) * C + c) * inv_count;
      const float s2 = __ldg(red + (static_cast<size_t>(g) * 2 + 1) * C + c) * inv_count;
      k1[q] = ga * rs;
      k2[q] = -ga * rs * rs * s2;
      k3[q] = -ga * rs * s1 + ga * rs * rs * mu * s2;
    }
    constexpr int kIn = 3;      // pixels in flight per thread (up to nine 16-byte loads issued before the first use)
    for (long long pp = static_cast<long long>(blockIdx.x) * m.lanes + m.pl; pp < npix; pp += kIn * stride) {
      uint4 dr[kIn], xr[kIn], orr[kIn];
      bool ok[kIn];
#pragma unroll
      for (int j = 0; j < kIn; ++j) {
        const long long pj = pp + j * stride;
        ok[j] = pj < npix;
        const size_t pix = static_cast<size_t>(pix0 + (ok[j] ? pj : pp));
        dr[j] = ldg16(dout + pix * ldd + d_coff + u * 8);
        xr[j] = ldg16(x + pix * ldx + x_coff + u * 8);
        if (act != CATB_ACT_NONE) orr[j] = ldg16(out + pix * ldo + o_coff + u * 8);
      }
#pragma unroll
      for (int j = 0; j < kIn; ++j) {
        if (!ok[j]) continue;
        const f8 dv = unpack8(dr[j]), xv = unpack8(xr[j]);
        f8 ov;
        if (act != CATB_ACT_NONE) ov = unpack8(orr[j]);
        f8 r;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float dz = act != CATB_ACT_NONE ? dv.v[q] * act_grad_from_out(ov.v[q], act) : dv.v[q];
          r.v[q] = k1[q] * dz + k2[q] * xv.v[q] + k3[q];
        }
        st16(dx + static_cast<size_t>(pix0 + pp + j * stride) * ldg + g_coff + u * 8, pack8(r));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// element-wise helpers
// ------------------------------------------------------------------------------------------------
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, int C, int H, int W, __nv_bfloat16* __restrict__ dst,
                                    int ldd, int d_coff, int Cp, long long pixels) {
  const long long total = pixels * Cp;
  const long long hw = static_cast<long long>(H) * W;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % Cp);
    const long long pix = idx / Cp;
    const long long n = pix / hw, rem = pix % hw;
    const float v = c < C ? src[(n * C + c) * hw + rem] : 0.f;
    dst[static_cast<size_t>(pix) * ldd + d_coff + c] = __float2bfloat16(v);
  }
}

__global__ void nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ src, int lds, int s_coff, int C, int H, int W,
                                    float* __restrict__ dst, long long pixels) {
  const long long hw = static_cast<long long>(H) * W;
  const long long total = pixels * C;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long rem = idx % hw;
    const int c = static_cast<int>((idx / hw) % C);
    const long long n = idx / (hw * C);
    dst[idx] = __bfloat162float(src[static_cast<size_t>(n * hw + rem) * lds + s_coff + c]);
  }
}

__global__ void copy_channels_kernel(const __nv_bfloat16* __restrict__ src, int lds, int s_coff,
                                     __nv_bfloat16* __restrict__ dst, int ldd, int d_coff, long long pixels, int C) {
  const long long total = pixels * C;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % C);
    const size_t pix = static_cast<size_t>(idx / C);
    dst[pix * ldd + d_coff + c] = src[pix * lds + s_coff + c];
  }
}

__global__ void act_bwd_kernel(const __nv_bfloat16* __restrict__ dout, int ldd, int d_coff,
                               const __nv_bfloat16* __restrict__ out, int ldo, int o_coff, __nv_bfloat16* __restrict__ dz,
                               int ldz, int z_coff, long long pixels, int C, int act) {
  const int U = C / 8;
  const long long total = pixels * U;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int u = static_cast<int>(idx % U);
    const size_t pix = static_cast<size_t>(idx / U);
    f8 d = unpack8(ldg16(dout + pix * ldd + d_coff + u * 8));
    const f8 o = unpack8(ldg16(out + pix * ldo + o_coff + u * 8));
#pragma unroll
    for (int q = 0; q < 8; ++q) d.v[q] *= act_grad_from_out(o.v[q], act);
    st16(dz + pix * ldz + z_coff + u * 8, pack8(d));
  }
}

__global__ void channel_sum_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int x_coff, long long pixels, int C,
                                   float* __restrict__ out) {
  extern __shared__ float sh[];
  const int U = C / 8;
  for (int i = threadIdx.x; i < C; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  const int lanes = U >= static_cast<int>(blockDim.x) ? 1 : blockDim.x / U;
  const int pl = threadIdx.x / U;
  if (U >= static_cast<int>(blockDim.x) || pl < lanes) {
    for (int u = threadIdx.x % U; u < U; u += blockDim.x) {
      f8 a;
#pragma unroll
      for (int q = 0; q < 8; ++q) a.v[q] = 0.f;
      for (long long pp = static_cast<long long>(blockIdx.x) * lanes + (U >= static_cast<int>(blockDim.x) ? 0 : pl);
           pp < pixels; pp += static_cast<long long>(gridDim.x) * lanes) {
        const f8 v = unpack8(ldg16(x + static_cast<size_t>(pp) * ldx + x_coff + u * 8));
#pragma unroll
        for (int q = 0; q < 8; ++q) a.v[q] += v.v[q];
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) atomicAdd(&sh[u * 8 + q], a.v[q]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(out + i, sh[i]);
}

__global__ void reflect_fold_kernel(const __nv_bfloat16* __restrict__ src, int lds, int s_coff,
                                    __nv_bfloat16* __restrict__ dst, int ldd, int d_coff,
                                    const __nv_bfloat16* __restrict__ add, int lda, int a_coff, int N, int H, int W, int C,
                                    int p) {
  const int U = C / 8;
  const int Hp = H + 2 * p, Wp = W + 2 * p;
  const long long total = static_cast<long long>(N) * H * W * U;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int u = static_cast<int>(idx % U);
    const long long pix = idx / U;
    const int w = static_cast<int>(pix % W);
    const int h = static_cast<int>((pix / W) % H);
    const int n = static_cast<int>(pix / (static_cast<long long>(W) * H));
    int hs[3], ws[3];
    const int nh = reflect_sources(h, H, p, hs), nw = reflect_sources(w, W, p, ws);
    f8 acc;
#pragma unroll
    for (int q = 0; q < 8; ++q) acc.v[q] = 0.f;
    for (int a = 0; a < nh; ++a)
      for (int b = 0; b < nw; ++b) {
        const f8 v = unpack8(
            ldg16(src + ((static_cast<size_t>(n) * Hp + hs[a] + p) * Wp + ws[b] + p) * lds + s_coff + u * 8));
#pragma unroll
        for (int q = 0; q < 8; ++q) acc.v[q] += v.v[q];
      }
    if (add != nullptr) {
      const f8 v = unpack8(ldg16(add + static_cast<size_t>(pix) * lda + a_coff + u * 8));
#pragma unroll
      for (int q = 0; q < 8; ++q) acc.v[q] += v.v[q];
    }
    st16(dst + static_cast<size_t>(pix) * ldd + d_coff + u * 8, pack8(acc));
  }
}

__global__ void add_kernel(const __nv_bfloat16* __restrict__ a, int lda, int a_coff, const __nv_bfloat16* __restrict__ b,
                           int ldb, int b_coff, __nv_bfloat16* __restrict__ dst, int ldd, int d_coff, long long pixels,
                           int C) {
  const int U = C / 8;
  const long long total = pixels * U;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int u = static_cast<int>(idx % U);
    const size_t pix = static_cast<size_t>(idx / U);
    f8 x = unpack8(ldg16(a + pix * lda + a_coff + u * 8));
    const f8 y = unpack8(ldg16(b + pix * ldb + b_coff + u * 8));
#pragma unroll
    for (int q = 0; q < 8; ++q) x.v[q] += y.v[q];
    st16(dst + pix * ldd + d_coff + u * 8, pack8(x));
  }
}

// ------------------------------------------------------------------------------------------------
// losses
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum(float v, float* sh) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float t = 0.f;
  if (warp == 0) {
    t = lane < static_cast<int>(blockDim.x >> 5) ? sh[lane] : 0.f;
    t = warp_sum(t);
  }
  __syncthreads();
  return t;  // valid in warp 0
}

// pred: fp32 with element stride `ld`; dpred: bf16 rows of 8 channels (channel 0 carries the gradient)
__global__ void gan_loss_kernel(const float* __restrict__ pred, long long n, int ld, int mode, int target_is_real,
                                int for_discriminator, float grad_scale, float* loss, __nv_bfloat16* dpred, int ldg,
                                int g_coff) {
  __shared__ float sh[32];
  float acc = 0.f;
  const float inv_n = 1.f / static_cast<float>(n);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float p = pred[i * ld];
    float l = 0.f, g = 0.f;
    if (mode == CATB_GAN_HINGE) {
      if (for_discriminator) {
        const float t = target_is_real ? p - 1.f : -p - 1.f;  // loss = -mean(min(t, 0))
        if (t < 0.f) {
          l = -t;
          g = target_is_real ? -1.f : 1.f;
        }
      } else {
        l = -p;
        g = -1.f;
      }
    } else if (mode == CATB_GAN_LSGAN) {
      const float t = target_is_real ? 1.f : 0.f;
      l = (p - t) * (p - t);
      g = 2.f * (p - t);
    } else {  // vanilla: BCE with logits
      const float t = target_is_real ? 1.f : 0.f;
      l = fmaxf(p, 0.f) - p * t + log1pf(expf(-fabsf(p)));
      g = 1.f / (1.f + expf(-p)) - t;
    }
    acc += l;
    if (dpred != nullptr) {
      f8 o;
#pragma unroll
      for (int q = 0; q < 8; ++q) o.v[q] = 0.f;
      o.v[0] = g * inv_n * grad_scale;
      st16(dpred + static_cast<size_t>(i) * ldg + g_coff, pack8(o));
    }
  }
  const float t = block_sum(acc, sh);
  if (threadIdx.x == 0) atomicAdd(loss, t * inv_n);
}

__global__ void recon_loss_kernel(const __nv_bfloat16* __restrict__ a, int lda, int a_coff, const __nv_bfloat16* __restrict__ b,
                               int ldb, int b_coff, long long pixels, int C, int Creal, int kind, float grad_scale, float* loss,
                               __nv_bfloat16* da, int ldg, int g_coff, const __nv_bfloat16* extra, int lde, int e_coff) {
  __shared__ float sh[32];
  const int U = C / 8;
  const long long total = pixels * U;
  const float inv = 1.f / (static_cast<float>(pixels) * Creal);
  float acc = 0.f;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int u = static_cast<int>(idx % U);
    const size_t pix = static_cast<size_t>(idx / U);
    const f8 x = unpack8(ldg16(a + pix * lda + a_coff + u * 8));
    const f8 y = unpack8(ldg16(b + pix * ldb + b_coff + u * 8));
    f8 g;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float dlt = x.v[q] - y.v[q];
      const bool real = u * 8 + q < Creal;
      float l, dl;
      if (kind == 1) {  // MSE
        l = dlt * dlt;
        dl = 2.f * dlt;
      } else if (kind == 2 && fabsf(dlt) < 1.f) {  // smooth L1, quadratic zone
        l = 0.5f * dlt * dlt;
        dl = dlt;
      } else {  // L1 (and the linear zone of smooth L1)
        l = fabsf(dlt) - (kind == 2 ? 0.5f : 0.f);
        dl = dlt > 0.f ? 1.f : (dlt < 0.f ? -1.f : 0.f);
      }
      if (real) acc += l;
      g.v[q] = real ? grad_scale * inv * dl : 0.f;
    }
    if (da != nullptr) {
      if (extra != nullptr) {
        const f8 e = unpack8(ld16(extra + pix * lde + e_coff + u * 8));  // may alias da: coherent load
#pragma unroll
        for (int q = 0; q < 8; ++q) g.v[q] += e.v[q];
      }
      st16(da + pix * ldg + g_coff + u * 8, pack8(g));
    }
  }
  const float t = block_sum(acc, sh);
  if (threadIdx.x == 0) atomicAdd(loss, t * inv);
}

// Gram matrix of B samples over K = pixels * C elements on the tensor cores (mma.sync m16n8k16, bf16 x bf16 -> fp32): the
// reduction runs over K, so X is both operands -- A = X (16 samples x 16 k), B = X^T -- and the SAME registers serve as
// the A and the B fragments.  A warp takes one (pixel, 32-channel group) per iteration: lane (g, t) loads 16 bytes
// (channels 8t .. 8t+7) of samples g and g + 8 of every 16-sample block, i.e. four lanes cover 64 contiguous bytes of a
// sample; the order of k inside an MMA is irrelevant as long as both operands use the same one, which they do by
// construction.  Two loads + four MMAs per KB of activations: HBM bound.  (The round-1 kernel formed every pair's dot
// product with scalar FMAs out of shared memory: 370 GB/s.)
__device__ __forceinline__ void mma_bf16_16816(float* d, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                               uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int NB>   // 16-sample blocks: 1 (B <= 16) or 2 (B <= 32)
__global__ void __launch_bounds__(256) gram_mma_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int x_coff, int B,
                                                        long long pps, int C, float* __restrict__ G) {
  __shared__ float red[32 * 32];
  for (int i = threadIdx.x; i < B * B; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int cg = (C + 31) / 32;                      // 32-channel groups per pixel
  const long long blocks = pps * cg;
  const long long wstride = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
  float acc[NB][NB][2][4];
#pragma unroll
  for (int i = 0; i < NB * NB * 8; ++i) (&acc[0][0][0][0])[i] = 0.f;
  for (long long blk = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); blk < blocks; blk += wstride) {
    const long long pix = blk / cg;
    const int c0 = static_cast<int>(blk - pix * cg) * 32 + t * 8;
    uint4 r[NB][2];
#pragma unroll
    for (int bi = 0; bi < NB; ++bi)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int smp = bi * 16 + hf * 8 + g;
        r[bi][hf] = (smp < B && c0 < C) ? ldg16(x + (static_cast<size_t>(smp) * pps + pix) * ldx + x_coff + c0) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
    for (int I = 0; I < NB; ++I)
#pragma unroll
      for (int J = 0; J < NB; ++J)
#pragma unroll
        for (int nh = 0; nh < 2; ++nh) {
          mma_bf16_16816(acc[I][J][nh], r[I][0].x, r[I][1].x, r[I][0].y, r[I][1].y, r[J][nh].x, r[J][nh].y);
          mma_bf16_16816(acc[I][J][nh], r[I][0].z, r[I][1].z, r[I][0].w, r[I][1].w, r[J][nh].z, r[J][nh].w);
        }
  }
  // accumulator layout: d0, d1 = (row g, cols 2t, 2t+1), d2, d3 = (row g + 8, same cols)
#pragma unroll
  for (int I = 0; I < NB; ++I)
#pragma unroll
    for (int J = 0; J < NB; ++J)
#pragma unroll
      for (int nh = 0; nh < 2; ++nh)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int i = I * 16 + g + (e >> 1) * 8, j = J * 16 + nh * 8 + 2 * t + (e & 1);
          if (i < B && j < B) atomicAdd(&red[i * B + j], acc[I][J][nh][e]);
        }
  __syncthreads();
  for (int i = threadIdx.x; i < B * B; i += blockDim.x) atomicAdd(G + i, red[i]);
}

// SIMT restatement of the Gram matrix (the round-1 kernel; kept for on-device bisection in tests: catb_gram_ref).
constexpr int kGramKT = 256;
__global__ void gram_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int x_coff, int B, long long pps, int C,
                            float* __restrict__ G) {
  extern __shared__ float tile[];  // [B][kGramKT + 1]
  const int U = C / 8;
  const long long units_per_sample = pps * U;  // 8-element units
  const long long tiles = (units_per_sample * 8 + kGramKT - 1) / kGramKT;
  const int pitch = kGramKT + 1;
  const int npairs = B * B;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};  // up to 4 pairs per thread (B <= 32 with 256 threads)
  for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
    // load: B samples x 32 units
    for (int i = threadIdx.x; i < B * (kGramKT / 8); i += blockDim.x) {
      const int b = i / (kGramKT / 8), uu = i % (kGramKT / 8);
      const long long unit = t * (kGramKT / 8) + uu;
      f8 v;
#pragma unroll
      for (int q = 0; q < 8; ++q) v.v[q] = 0.f;
      if (unit < units_per_sample) {
        const long long pix = static_cast<long long>(b) * pps + unit / U;
        v = unpack8(ldg16(x + static_cast<size_t>(pix) * ldx + x_coff + (unit % U) * 8));
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) tile[b * pitch + uu * 8 + q] = v.v[q];
    }
    __syncthreads();
#pragma unroll
    for (int pi = 0; pi < 4; ++pi) {
      const int pr = threadIdx.x + pi * blockDim.x;
      if (pr < npairs) {
        const float* ri = tile + (pr / B) * pitch;
        const float* rj = tile + (pr % B) * pitch;
        float s = 0.f;
#pragma unroll 8
        for (int k = 0; k < kGramKT; ++k) s += ri[k] * rj[k];
        acc[pi] += s;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int pi = 0; pi < 4; ++pi) {
    const int pr = threadIdx.x + pi * blockDim.x;
    if (pr < npairs) atomicAdd(G + pr, acc[pi]);
  }
}

// single block: KA value and dX coefficient matrix
__global__ void ka_finish_kernel(const float* __restrict__ Gx, const float* __restrict__ Gy, int B, float loss_scale,
                                 float* loss, float* ka_value, float* coef) {
  __shared__ float sh[32];
  __shared__ float res[3];
  float num = 0.f, sx = 0.f, sy = 0.f;
  for (int i = threadIdx.x; i < B * B; i += blockDim.x) {
    num += Gx[i] * Gy[i];
    sx += Gx[i] * Gx[i];
    sy += Gy[i] * Gy[i];
  }
  float t = block_sum(num, sh);
  if (threadIdx.x == 0) res[0] = t;
  t = block_sum(sx, sh);
  if (threadIdx.x == 0) res[1] = t;
  t = block_sum(sy, sh);
  if (threadIdx.x == 0) res[2] = t;
  __syncthreads();
  const float nx = sqrtf(res[1]), ny = sqrtf(res[2]);
  const float ka = res[0] / (nx * ny);
  if (threadIdx.x == 0) {
    if (ka_value) *ka_value = ka;
    if (loss) atomicAdd(loss, loss_scale * ka);
  }
  if (coef != nullptr) {
    // d(KA)/dX = 2 (Ky/(nx ny) - num Kx/(nx^3 ny)) X
    const float c1 = 1.f / (nx * ny), c2 = res[0] / (nx * nx * nx * ny);
    for (int i = threadIdx.x; i < B * B; i += blockDim.x) coef[i] = loss_scale * 2.f * (Gy[i] * c1 - Gx[i] * c2);
  }
}

__global__ void ka_bwd_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int x_coff, int B, long long pps, int C,
                              const float* __restrict__ coef, __nv_bfloat16* __restrict__ dx, int ldg, int g_coff,
                              int accumulate) {
  extern __shared__ float cf[];  // [B*B]
  for (int i = threadIdx.x; i < B * B; i += blockDim.x) cf[i] = coef[i];
  __syncthreads();
  const int U = C / 8;
  const long long total = pps * U;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int u = static_cast<int>(idx % U);
    const long long pp = idx / U;
    for (int b = 0; b < B; ++b) {
      f8 acc;
#pragma unroll
      for (int q = 0; q < 8; ++q) acc.v[q] = 0.f;
      for (int j = 0; j < B; ++j) {
        const f8 v = unpack8(ldg16(x + static_cast<size_t>(j * pps + pp) * ldx + x_coff + u * 8));
        const float c = cf[b * B + j];
#pragma unroll
        for (int q = 0; q < 8; ++q) acc.v[q] += c * v.v[q];
      }
      __nv_bfloat16* dp = dx + static_cast<size_t>(b * pps + pp) * ldg + g_coff + u * 8;
      if (accumulate) {
        const f8 old = unpack8(ld16(dp));
#pragma unroll
        for (int q = 0; q < 8; ++q) acc.v[q] += old.v[q];
      }
      st16(dp, pack8(acc));
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Adam
// ------------------------------------------------------------------------------------------------
__global__ void adam_tick_kernel(int* step) { *step += 1; }

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, const float* __restrict__ lr_ptr, float b1, float b2,
                            float eps, float gscale, const int* __restrict__ step) {
  const float t = static_cast<float>(*step);
  const float bc1 = 1.f - powf(b1, t), bc2 = 1.f - powf(b2, t);
  const float step_size = *lr_ptr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float gi = g[i] * gscale;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
  }
}

int init_dwconv_attributes();
int init_simt_attributes() { return init_dwconv_attributes(); }

}  // namespace catb

using namespace catb;

#define CHK_SLICE(ld, coff, C)                                                                     \
  CATB_REQUIRE((ld) % 8 == 0 && (coff) % 8 == 0 && (C) % 8 == 0 && (C) > 0 && (coff) + (C) <= (ld), \
               "bad channel slice (ld=%d coff=%d C=%d)", (int)(ld), (int)(coff), (int)(C))

// Blocks of 512 threads, at most one per SM in total: every block ends with 2*C atomic adds onto the SAME 2*C addresses
// (per group), and same-address atomics from different SMs serialise in L2 (~20-30 ns each) -- with 592 blocks that tail
// was a fixed ~10 us per launch, more than the streaming part of most of these reductions.
constexpr int kReduceThreads = 512;
static dim3 reduce_grid(long long pixels_per_group, int C, int groups) {
  const int U = C / 8;
  const int lanes = U >= kReduceThreads ? 1 : kReduceThreads / U;
  long long gx = (pixels_per_group + static_cast<long long>(lanes) * 8 - 1) / (static_cast<long long>(lanes) * 8);
  const long long cap = std::max(1, 148 / groups);
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  return dim3(static_cast<unsigned>(gx), groups, 1);
}

// grid of the unit-mapped element-wise kernels: x = pixel blocks (two pixels per thread iteration), y = groups
static dim3 elementwise_grid(long long pixels_per_group, int C, int groups, int blocks_per_sm = 8) {
  const int U = C / 8;
  const int lanes = U >= 256 ? 1 : 256 / U;
  long long gx = (pixels_per_group + static_cast<long long>(lanes) * 2 - 1) / (static_cast<long long>(lanes) * 2);
  const long long cap = std::max(1, 148 * blocks_per_sm / groups);
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  return dim3(static_cast<unsigned>(gx), groups, 1);
}

extern "C" int catb_norm_stats(const void* x, int ldx, int x_coff, int N, int HW, int C, int per_sample, float* sums,
                               catb_stream_t s) {
  CHK_SLICE(ldx, x_coff, C);
  const long long pixels = static_cast<long long>(N) * HW;
  const dim3 grid = reduce_grid(per_sample ? HW : pixels, C, per_sample ? N : 1);
  channel_reduce_kernel<0><<<grid, kReduceThreads, 2 * C * sizeof(float), S(s)>>>(
      static_cast<const __nv_bfloat16*>(x), ldx, x_coff, nullptr, 0, 0, nullptr, 0, 0, HW, C, per_sample, nullptr, 0,
      sums, pixels);
  return check_launch("norm_stats");
}

extern "C" int catb_norm_finalize(const float* sums, int G, int C, float count, float eps, float momentum,
                                  const float* gamma, const float* beta, float* running_mean, float* running_var,
                                  float* scale, float* shift, float* mean_rstd, catb_stream_t s) {
  CATB_REQUIRE(G > 0 && C > 0, "bad norm extents");
  CATB_REQUIRE(sums != nullptr || (running_mean != nullptr && running_var != nullptr),
               "eval-mode normalisation needs running statistics");
  norm_finalize_kernel<<<(G * C + 127) / 128, 128, 0, S(s)>>>(sums, G, C, count, eps, momentum, gamma, beta,
                                                               running_mean, running_var, scale, shift, mean_rstd);
  return check_launch("norm_finalize");
}

extern "C" int catb_norm_apply(const void* x, int ldx, int x_coff, void* y, int ldy, int y_coff, const void* residual,
                               int ldr, int r_coff, int N, int HW, int C, int per_sample, const float* scale,
                               const float* shift, int act, catb_stream_t s) {
  CHK_SLICE(ldx, x_coff, C);
  CHK_SLICE(ldy, y_coff, C);
  const long long pixels = static_cast<long long>(N) * HW;
  NormFin fin;
  memset(&fin, 0, sizeof(fin));
  norm_apply_kernel<<<elementwise_grid(per_sample ? HW : pixels, C, per_sample ? N : 1), 256, 0, S(s)>>>(
      static_cast<const __nv_bfloat16*>(x), ldx, x_coff, static_cast<__nv_bfloat16*>(y), ldy, y_coff,
      static_cast<const __nv_bfloat16*>(residual), ldr, r_coff, HW, C, per_sample, scale, shift, act, pixels, fin);
  return check_launch("norm_apply");
}

extern "C" int catb_norm_apply_fused(const void* x, int ldx, int x_coff, void* y, int ldy, int y_coff, const void* residual,
                                     int ldr, int r_coff, int N, int HW, int C, int per_sample, const float* sums, float count,
                                     float eps, float momentum, const float* gamma, const float* beta, float* running_mean,
                                     float* running_var, float* scale, float* shift, float* mean_rstd, int act,
                                     catb_stream_t s) {
  CHK_SLICE(ldx, x_coff, C);
  CHK_SLICE(ldy, y_coff, C);
  CATB_REQUIRE(sums != nullptr && scale != nullptr && shift != nullptr, "fused finalize needs the sums and the scale / shift outputs");
  const long long pixels = static_cast<long long>(N) * HW;
  NormFin fin;
  fin.sums = sums;
  fin.gamma = gamma;
  fin.beta = beta;
  fin.running_mean = running_mean;
  fin.running_var = running_var;
  fin.scale_out = scale;
  fin.shift_out = shift;
  fin.mean_rstd = mean_rstd;
  fin.count = count;
  fin.eps = eps;
  fin.momentum = momentum;
  CATB_REQUIRE(2 * C * sizeof(float) <= 48 * 1024, "too many channels (%d) for the fused finalize", C);
  // fewer, longer-lived blocks than the plain apply: every block derives the group's scale / shift once
  norm_apply_kernel<<<elementwise_grid(per_sample ? HW : pixels, C, per_sample ? N : 1, 3), 256, 2 * C * sizeof(float), S(s)>>>(
      static_cast<const __nv_bfloat16*>(x), ldx, x_coff, static_cast<__nv_bfloat16*>(y), ldy, y_coff,
      static_cast<const __nv_bfloat16*>(residual), ldr, r_coff, HW, C, per_sample, nullptr, nullptr, act, pixels, fin);
  return check_launch("norm_apply_fused");
}

extern "C" int catb_norm_bwd_reduce(const void* dout, int ldd, int d_coff, const void* out, int ldo, int o_coff,
                                    const void* x, int ldx, int x_coff, int N, int HW, int C, int per_sample,
                                    const float* mean_rstd, int act, float* red, catb_stream_t s) {
  CHK_SLICE(ldx, x_coff, C);
  CHK_SLICE(ldd, d_coff, C);
  const long long pixels = static_cast<long long>(N) * HW;
  const dim3 grid = reduce_grid(per_sample ? HW : pixels, C, per_sample ? N : 1);
  channel_reduce_kernel<1><<<grid, kReduceThreads, 2 * C * sizeof(float), S(s)>>>(
      static_cast<const __nv_bfloat16*>(x), ldx, x_coff, static_cast<const __nv_bfloat16*>(dout), ldd, d_coff,
      static_cast<const __nv_bfloat16*>(out), ldo, o_coff, HW, C, per_sample, mean_rstd, act, red, pixels);
  return check_launch("norm_bwd_reduce");
}

extern "C" int catb_norm_bwd_apply(const void* dout, int ldd, int d_coff, const void* out, int ldo, int o_coff,
                                   const void* x, int ldx, int x_coff, void* dx, int ldg, int g_coff, int N, int HW,
                                   int C, int per_sample, const float* mean_rstd, const float* gamma, const float* red,
                                   float count, int act, float* dgamma, float* dbeta, catb_stream_t s) {
  CHK_SLICE(ldx, x_coff, C);
  CHK_SLICE(ldd, d_coff, C);
  CHK_SLICE(ldg, g_coff, C);
  const long long pixels = static_cast<long long>(N) * HW;
  norm_bwd_apply_kernel<<<elementwise_grid(per_sample ? HW : pixels, C, per_sample ? N : 1), 256, 0, S(s)>>>(
      static_cast<const __nv_bfloat16*>(dout), ldd, d_coff, static_cast<const __nv_bfloat16*>(out), ldo, o_coff,
      static_cast<const __nv_bfloat16*>(x), ldx, x_coff, static_cast<__nv_bfloat16*>(dx), ldg, g_coff, HW, C,
      per_sample, mean_rstd, gamma, red, count, act, dgamma, dbeta, per_sample ? N : 1, pixels);
  return check_launch("norm_bwd_apply");
}

extern "C" int catb_nchw_to_nhwc(const float* src, int N, int C, int H, int W, void* dst, int ldd, int d_coff,
                                 catb_stream_t s) {
  const int Cp = (C + 7) / 8 * 8;
  CHK_SLICE(ldd, d_coff, Cp);
  const long long pixels = static_cast<long long>(N) * H * W;
  nchw_to_nhwc_kernel<<<grid_for(pixels * Cp, 256), 256, 0, S(s)>>>(src, C, H, W, static_cast<__nv_bfloat16*>(dst), ldd,
                                                                     d_coff, Cp, pixels);
  return check_launch("nchw_to_nhwc");
}

extern "C" int catb_nhwc_to_nchw(const void* src, int lds, int s_coff, int N, int C, int H, int W, float* dst,
                                 catb_stream_t s) {
  CATB_REQUIRE(C > 0 && s_coff + C <= lds, "bad slice");
  const long long pixels = static_cast<long long>(N) * H * W;
  nhwc_to_nchw_kernel<<<grid_for(pixels * C, 256), 256, 0, S(s)>>>(static_cast<const __nv_bfloat16*>(src), lds, s_coff,
                                                                    C, H, W, dst, pixels);
  return check_launch("nhwc_to_nchw");
}

extern "C" int catb_copy_channels(const void* src, int lds, int s_coff, void* dst, int ldd, int d_coff,
                                  long long pixels, int C, catb_stream_t s) {
  CATB_REQUIRE(C > 0 && s_coff + C <= lds && d_coff + C <= ldd, "bad slice");
  copy_channels_kernel<<<grid_for(pixels * C, 256), 256, 0, S(s)>>>(static_cast<const __nv_bfloat16*>(src), lds, s_coff,
                                                                     static_cast<__nv_bfloat16*>(dst), ldd, d_coff,
                                                                     pixels, C);
  return check_launch("copy_channels");
}

extern "C" int catb_act_bwd(const void* dout, int ldd, int d_coff, const void* out, int ldo, int o_coff, void* dz,
                            int ldz, int z_coff, long long pixels, int C, int act, catb_stream_t s) {
  CHK_SLICE(ldd, d_coff, C);
  CHK_SLICE(ldo, o_coff, C);
  CHK_SLICE(ldz, z_coff, C);
  act_bwd_kernel<<<grid_for(pixels * (C / 8), 256), 256, 0, S(s)>>>(
      static_cast<const __nv_bfloat16*>(dout), ldd, d_coff, static_cast<const __nv_bfloat16*>(out), ldo, o_coff,
      static_cast<__nv_bfloat16*>(dz), ldz, z_coff, pixels, C, act);
  return check_launch("act_bwd");
}

extern "C" int catb_channel_sum(const void* x, int ldx, int x_coff, long long pixels, int C, float* out,
                                catb_stream_t s) {
  CHK_SLICE(ldx, x_coff, C);
  const dim3 grid = reduce_grid(pixels, C, 1);
  channel_sum_kernel<<<grid, 256, C * sizeof(float), S(s)>>>(static_cast<const __nv_bfloat16*>(x), ldx, x_coff, pixels,
                                                              C, out);
  return check_launch("channel_sum");
}

extern "C" int catb_reflect_fold(const void* src, int lds, int s_coff, void* dst, int ldd, int d_coff, const void* add,
                                 int lda, int a_coff, int N, int H, int W, int C, int p, catb_stream_t s) {
  CHK_SLICE(lds, s_coff, C);
  CHK_SLICE(ldd, d_coff, C);
  CATB_REQUIRE(p >= 0 && p < H && p < W, "reflect padding must be smaller than the image");
  const long long total = static_cast<long long>(N) * H * W * (C / 8);
  reflect_fold_kernel<<<grid_for(total, 256), 256, 0, S(s)>>>(static_cast<const __nv_bfloat16*>(src), lds, s_coff,
                                                               static_cast<__nv_bfloat16*>(dst), ldd, d_coff,
                                                               static_cast<const __nv_bfloat16*>(add), lda, a_coff, N, H,
                                                               W, C, p);
  return check_launch("reflect_fold");
}

extern "C" int catb_add(const void* a, int lda, int a_coff, const void* b, int ldb, int b_coff, void* dst, int ldd,
                        int d_coff, long long pixels, int C, catb_stream_t s) {
  CHK_SLICE(lda, a_coff, C);
  CHK_SLICE(ldb, b_coff, C);
  CHK_SLICE(ldd, d_coff, C);
  add_kernel<<<grid_for(pixels * (C / 8), 256), 256, 0, S(s)>>>(
      static_cast<const __nv_bfloat16*>(a), lda, a_coff, static_cast<const __nv_bfloat16*>(b), ldb, b_coff,
      static_cast<__nv_bfloat16*>(dst), ldd, d_coff, pixels, C);
  return check_launch("add");
}

extern "C" int catb_gan_loss(const float* pred, long long n, int ld, int mode, int target_is_real,
                             int for_discriminator, float grad_scale, float* loss, void* dpred, int ldg, int g_coff,
                             catb_stream_t s) {
  CATB_REQUIRE(n > 0 && ld > 0, "empty prediction");
  CATB_REQUIRE(mode >= CATB_GAN_HINGE && mode <= CATB_GAN_VANILLA, "unknown gan mode %d", mode);
  CATB_REQUIRE(for_discriminator || target_is_real || mode != CATB_GAN_HINGE,
               "hinge generator loss requires target_is_real (loss.py:94)");
  gan_loss_kernel<<<grid_for(n, 256, 148), 256, 0, S(s)>>>(pred, n, ld, mode, target_is_real, for_discriminator,
                                                            grad_scale, loss, static_cast<__nv_bfloat16*>(dpred), ldg,
                                                            g_coff);
  return check_launch("gan_loss");
}

extern "C" int catb_recon_loss(const void* a, int lda, int a_coff, const void* b, int ldb, int b_coff, long long pixels,
                               int C, int Creal, int kind, float grad_scale, float* loss, void* da, int ldg, int g_coff,
                               const void* extra, int lde, int e_coff, catb_stream_t s) {
  CATB_REQUIRE(kind >= 0 && kind <= 2, "unknown reconstruction loss kind %d", kind);
  CHK_SLICE(lda, a_coff, C);
  CHK_SLICE(ldb, b_coff, C);
  CATB_REQUIRE(Creal > 0 && Creal <= C, "bad real channel count");
  recon_loss_kernel<<<grid_for(pixels * (C / 8), 256, 148 * 4), 256, 0, S(s)>>>(
      static_cast<const __nv_bfloat16*>(a), lda, a_coff, static_cast<const __nv_bfloat16*>(b), ldb, b_coff, pixels, C,
      Creal, kind, grad_scale, loss, static_cast<__nv_bfloat16*>(da), ldg, g_coff, static_cast<const __nv_bfloat16*>(extra),
      lde, e_coff);
  return check_launch("recon_loss");
}

extern "C" int catb_gram(const void* x, int ldx, int x_coff, int B, long long pixels_per_sample, int C, float* G,
                         catb_stream_t s) {
  CHK_SLICE(ldx, x_coff, C);
  CATB_REQUIRE(B >= 1 && B <= 32, "KA kernels support 1 <= batch <= 32 per device (got %d)", B);
  const long long blocks = pixels_per_sample * ((C + 31) / 32);      // one per warp iteration
  const int grid = grid_for(blocks, 8, 148 * 4);
  if (B <= 16)
    gram_mma_kernel<1><<<grid, 256, 0, S(s)>>>(static_cast<const __nv_bfloat16*>(x), ldx, x_coff, B, pixels_per_sample, C, G);
  else
    gram_mma_kernel<2><<<grid, 256, 0, S(s)>>>(static_cast<const __nv_bfloat16*>(x), ldx, x_coff, B, pixels_per_sample, C, G);
  return check_launch("gram");
}

extern "C" int catb_gram_ref(const void* x, int ldx, int x_coff, int B, long long pixels_per_sample, int C, float* G,
                             catb_stream_t s) {
  CHK_SLICE(ldx, x_coff, C);
  CATB_REQUIRE(B >= 1 && B <= 32, "KA kernels support 1 <= batch <= 32 per device (got %d)", B);
  const long long tiles = (pixels_per_sample * C + kGramKT - 1) / kGramKT;
  const size_t smem = static_cast<size_t>(B) * (kGramKT + 1) * sizeof(float);
  gram_kernel<<<grid_for(tiles, 1, 148 * 4), 256, smem, S(s)>>>(static_cast<const __nv_bfloat16*>(x), ldx, x_coff, B,
                                                                 pixels_per_sample, C, G);
  return check_launch("gram_ref");
}

extern "C" int catb_ka_finish(const float* Gx, const float* Gy, int B, float loss_scale, float* loss, float* ka_value,
                              float* coef, catb_stream_t s) {
  CATB_REQUIRE(B >= 1 && B <= 32, "KA kernels support 1 <= batch <= 32 per device (got %d)", B);
  ka_finish_kernel<<<1, 256, 0, S(s)>>>(Gx, Gy, B, loss_scale, loss, ka_value, coef);
  return check_launch("ka_finish");
}

extern "C" int catb_ka_bwd(const void* x, int ldx, int x_coff, int B, long long pixels_per_sample, int C,
                           const float* coef, void* dx, int ldg, int g_coff, int accumulate, catb_stream_t s) {
  CHK_SLICE(ldx, x_coff, C);
  CHK_SLICE(ldg, g_coff, C);
  CATB_REQUIRE(B >= 1 && B <= 32, "KA kernels support 1 <= batch <= 32 per device (got %d)", B);
  ka_bwd_kernel<<<grid_for(pixels_per_sample * (C / 8), 128), 128, B * B * sizeof(float), S(s)>>>(
      static_cast<const __nv_bfloat16*>(x), ldx, x_coff, B, pixels_per_sample, C, coef,
      static_cast<__nv_bfloat16*>(dx), ldg, g_coff, accumulate);
  return check_launch("ka_bwd");
}

extern "C" int catb_adam(float* param, const float* grad, float* m, float* v, long long n, const float* lr, float beta1,
                         float beta2, float eps, float grad_scale, int* step_count, catb_stream_t s) {
  CATB_REQUIRE(n > 0, "empty parameter arena");
  adam_tick_kernel<<<1, 1, 0, S(s)>>>(step_count);
  adam_kernel<<<grid_for(n, 256), 256, 0, S(s)>>>(param, grad, m, v, n, lr, beta1, beta2, eps, grad_scale, step_count);
  return check_launch("adam");
}
