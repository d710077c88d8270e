// CUDA-core kernels of the SPADE distillation path: nearest resize / 2x up-sampling
// adjoint, SPADE modulation (forward + backward), 3x3/s2 average pool and 2x2 max pool (forward + backward),
// table-driven spectral normalisation (power iteration, scaling, backward), one-hot + instance-edge
// preprocessing, small fp32 vector gathers for fused biases.  All HBM bound: one thread per (pixel, 8-channel
// unit) with 16-byte accesses, grid-stride loops sized in multiples of the SM count.
#include "common.cuh"

namespace catb {

static inline cudaStream_t S(catb_stream_t s) { return static_cast<cudaStream_t>(s); }
static inline int grid_for(long long work, int block, int max_blocks = 148 * 16) {
  long long g = (work + block - 1) / block;
  if (g < 1) g = 1;
  if (g > max_blocks) g = max_blocks;
  return static_cast<int>(g);
}

#define PIXEL_UNIT_LOOP(total)                                                                             \
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < (total);       \
       idx += static_cast<long long>(gridDim.x) * blockDim.x)

// ------------------------------------------------------------------------------------------------
// nearest-neighbour resize (F.interpolate(mode='nearest') / nn.Upsample(scale_factor=2))
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int nearest_src(int dst, float scale, int in) {
  const int s = static_cast<int>(floorf(dst * scale));
  return s < in - 1 ? s : in - 1;
}

__global__ void resize_nearest_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int x_coff, int H, int W,
                                      __nv_bfloat16* __restrict__ y, int ldy, int y_coff, int N, int OH, int OW, int C) {
  const int U = C / 8;
  const long long total = static_cast<long long>(N) * OH * OW * U;
  const float sh = static_cast<float>(H) / OH, sw = static_cast<float>(W) / OW;
  PIXEL_UNIT_LOOP(total) {
    const int u = static_cast<int>(idx % U);
    const long long pix = idx / U;
    const int ow = static_cast<int>(pix % OW);
    const int oh = static_cast<int>((pix / OW) % OH);
    const int n = static_cast<int>(pix / (static_cast<long long>(OW) * OH));
    const int ih = nearest_src(oh, sh, H), iw = nearest_src(ow, sw, W);
    st16(y + static_cast<size_t>(pix) * ldy + y_coff + u * 8,
         ldg16(x + ((static_cast<size_t>(n) * H + ih) * W + iw) * ldx + x_coff + u * 8));
  }
}

// adjoint of the 2x nearest up-sampling: dx[h,w] = sum of the 2x2 block of dy (+ add)
__global__ void upsample2x_bwd_kernel(const __nv_bfloat16* __restrict__ dy, int ldy, int y_coff,
                                      __nv_bfloat16* __restrict__ dx, int ldx, int x_coff, int N, int H, int W, int C) {
  const int U = C / 8;
  const long long total = static_cast<long long>(N) * H * W * U;
  PIXEL_UNIT_LOOP(total) {
    const int u = static_cast<int>(idx % U);
    const long long pix = idx / U;
    const int w = static_cast<int>(pix % W);
    const int h = static_cast<int>((pix / W) % H);
    const int n = static_cast<int>(pix / (static_cast<long long>(W) * H));
    f8 acc;
#pragma unroll
    for (int q = 0; q < 8; ++q) acc.v[q] = 0.f;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const f8 g = unpack8(ldg16(dy + ((static_cast<size_t>(n) * 2 * H + 2 * h + a) * 2 * W + 2 * w + b) * ldy + y_coff + u * 8));
#pragma unroll
        for (int q = 0; q < 8; ++q) acc.v[q] += g.v[q];
      }
    st16(dx + static_cast<size_t>(pix) * ldx + x_coff + u * 8, pack8(acc));
  }
}

// ------------------------------------------------------------------------------------------------
// SPADE modulation: y = act(xhat * (1 + gamma) + beta), xhat = x * scale[c] + shift[c]
// ------------------------------------------------------------------------------------------------
__global__ void spade_modulate_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int x_coff,
                                      const __nv_bfloat16* __restrict__ gm, int ldg, int g_coff,
                                      const __nv_bfloat16* __restrict__ bt, int ldb, int b_coff,
                                      __nv_bfloat16* __restrict__ y, int ldy, int y_coff, long long pixels, int C,
                                      const float* __restrict__ scale, const float* __restrict__ shift, int act) {
  const int U = C / 8;
  const long long total = pixels * U;
  PIXEL_UNIT_LOOP(total) {
    const int u = static_cast<int>(idx % U);
    const size_t pix = static_cast<size_t>(idx / U);
    const f8 xv = unpack8(ldg16(x + pix * ldx + x_coff + u * 8));
    const f8 g = unpack8(ldg16(gm + pix * ldg + g_coff + u * 8));
    const f8 b = unpack8(ldg16(bt + pix * ldb + b_coff + u * 8));
    f8 o;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int c = u * 8 + q;
      const float xh = xv.v[q] * scale[c] + shift[c];
      o.v[q] = apply_act(xh * (1.f + g.v[q]) + b.v[q], act);
    }
    st16(y + pix * ldy + y_coff + u * 8, pack8(o));
  }
}

// dz = dy * act'(y);  dgamma = dz * xhat;  dbeta = dz;  dn = dz * (1 + gamma)   (dn then goes through the
// backward of the parameter-free norm, catb_norm_bwd_*)
__global__ void spade_modulate_bwd_kernel(const __nv_bfloat16* __restrict__ dy, int ldd, int d_coff,
                                          const __nv_bfloat16* __restrict__ y, int ldy, int y_coff,
                                          const __nv_bfloat16* __restrict__ x, int ldx, int x_coff,
                                          const __nv_bfloat16* __restrict__ gm, int ldg, int g_coff,
                                          __nv_bfloat16* __restrict__ dgm, int ldo, int o_coff,
                                          __nv_bfloat16* __restrict__ dbt, int ldp, int p_coff,
                                          __nv_bfloat16* __restrict__ dn, int ldn, int n_coff, long long pixels, int C,
                                          const float* __restrict__ scale, const float* __restrict__ shift, int act) {
  const int U = C / 8;
  const long long total = pixels * U;
  PIXEL_UNIT_LOOP(total) {
    const int u = static_cast<int>(idx % U);
    const size_t pix = static_cast<size_t>(idx / U);
    const f8 d = unpack8(ldg16(dy + pix * ldd + d_coff + u * 8));
    const f8 o = unpack8(ldg16(y + pix * ldy + y_coff + u * 8));
    const f8 xv = unpack8(ldg16(x + pix * ldx + x_coff + u * 8));
    const f8 g = unpack8(ldg16(gm + pix * ldg + g_coff + u * 8));
    f8 a, b, c3;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int c = u * 8 + q;
      const float dz = d.v[q] * act_grad_from_out(o.v[q], act);
      const float xh = xv.v[q] * scale[c] + shift[c];
      a.v[q] = dz * xh;
      b.v[q] = dz;
      c3.v[q] = dz * (1.f + g.v[q]);
    }
    st16(dgm + pix * ldo + o_coff + u * 8, pack8(a));
    st16(dbt + pix * ldp + p_coff + u * 8, pack8(b));
    st16(dn + pix * ldn + n_coff + u * 8, pack8(c3));
  }
}

__global__ void act_fwd_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int x_coff, __nv_bfloat16* __restrict__ y,
                               int ldy, int y_coff, long long pixels, int C, int act) {
  const int U = C / 8;
  const long long total = pixels * U;
  PIXEL_UNIT_LOOP(total) {
    const int u = static_cast<int>(idx % U);
    const size_t pix = static_cast<size_t>(idx / U);
    f8 v = unpack8(ldg16(x + pix * ldx + x_coff + u * 8));
#pragma unroll
    for (int q = 0; q < 8; ++q) v.v[q] = apply_act(v.v[q], act);
    st16(y + pix * ldy + y_coff + u * 8, pack8(v));
  }
}

// ------------------------------------------------------------------------------------------------
// pooling
// ------------------------------------------------------------------------------------------------
// F.avg_pool2d(kernel 3, stride 2, padding 1, count_include_pad=False): OH = (H+1)/2
__global__ void avgpool3s2_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int x_coff, int H, int W,
                                  __nv_bfloat16* __restrict__ y, int ldy, int y_coff, int N, int OH, int OW, int C) {
  const int U = C / 8;
  const long long total = static_cast<long long>(N) * OH * OW * U;
  PIXEL_UNIT_LOOP(total) {
    const int u = static_cast<int>(idx % U);
    const long long pix = idx / U;
    const int ow = static_cast<int>(pix % OW);
    const int oh = static_cast<int>((pix / OW) % OH);
    const int n = static_cast<int>(pix / (static_cast<long long>(OW) * OH));
    f8 acc;
#pragma unroll
    for (int q = 0; q < 8; ++q) acc.v[q] = 0.f;
    int cnt = 0;
    for (int r = -1; r <= 1; ++r) {
      const int ih = 2 * oh + r;
      if (ih < 0 || ih >= H) continue;
      for (int s = -1; s <= 1; ++s) {
        const int iw = 2 * ow + s;
        if (iw < 0 || iw >= W) continue;
        const f8 v = unpack8(ldg16(x + ((static_cast<size_t>(n) * H + ih) * W + iw) * ldx + x_coff + u * 8));
#pragma unroll
        for (int q = 0; q < 8; ++q) acc.v[q] += v.v[q];
        ++cnt;
      }
    }
    const float inv = 1.f / cnt;
#pragma unroll
    for (int q = 0; q < 8; ++q) acc.v[q] *= inv;
    st16(y + static_cast<size_t>(pix) * ldy + y_coff + u * 8, pack8(acc));
  }
}

__device__ __forceinline__ int avg_count_1d(int o, int L) {  // valid taps of output o along one axis
  int c = 0;
  for (int r = -1; r <= 1; ++r) c += (2 * o + r >= 0 && 2 * o + r < L) ? 1 : 0;
  return c;
}

// dx[ih,iw] = add[ih,iw] + sum over outputs whose window contains (ih,iw) of dy / count
__global__ void avgpool3s2_bwd_kernel(const __nv_bfloat16* __restrict__ dy, int ldy, int y_coff, int OH, int OW,
                                      const __nv_bfloat16* __restrict__ add, int lda, int a_coff,
                                      __nv_bfloat16* __restrict__ dx, int ldx, int x_coff, int N, int H, int W, int C) {
  const int U = C / 8;
  const long long total = static_cast<long long>(N) * H * W * U;
  PIXEL_UNIT_LOOP(total) {
    const int u = static_cast<int>(idx % U);
    const long long pix = idx / U;
    const int w = static_cast<int>(pix % W);
    const int h = static_cast<int>((pix / W) % H);
    const int n = static_cast<int>(pix / (static_cast<long long>(W) * H));
    f8 acc;
    if (add != nullptr) {
      acc = unpack8(ld16(add + static_cast<size_t>(pix) * lda + a_coff + u * 8));  // may alias dx: coherent load
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q) acc.v[q] = 0.f;
    }
    // outputs oh with |2*oh - h| <= 1
    for (int oh = (h > 0 ? (h - 1 + 1) / 2 : 0); oh <= (h + 1) / 2 && oh < OH; ++oh) {
      if (2 * oh - h > 1 || h - 2 * oh > 1) continue;
      const int ch = avg_count_1d(oh, H);
      for (int ow = (w > 0 ? (w - 1 + 1) / 2 : 0); ow <= (w + 1) / 2 && ow < OW; ++ow) {
        if (2 * ow - w > 1 || w - 2 * ow > 1) continue;
        const float inv = 1.f / (ch * avg_count_1d(ow, W));
        const f8 g = unpack8(ldg16(dy + ((static_cast<size_t>(n) * OH + oh) * OW + ow) * ldy + y_coff + u * 8));
#pragma unroll
        for (int q = 0; q < 8; ++q) acc.v[q] += g.v[q] * inv;
      }
    }
    st16(dx + static_cast<size_t>(pix) * ldx + x_coff + u * 8, pack8(acc));
  }
}

// F.max_pool2d(2, 2) (floor mode): OH = H/2
__global__ void maxpool2_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int x_coff, int H, int W,
                                __nv_bfloat16* __restrict__ y, int ldy, int y_coff, int N, int OH, int OW, int C) {
  const int U = C / 8;
  const long long total = static_cast<long long>(N) * OH * OW * U;
  PIXEL_UNIT_LOOP(total) {
    const int u = static_cast<int>(idx % U);
    const long long pix = idx / U;
    const int ow = static_cast<int>(pix % OW);
    const int oh = static_cast<int>((pix / OW) % OH);
    const int n = static_cast<int>(pix / (static_cast<long long>(OW) * OH));
    f8 m;
#pragma unroll
    for (int q = 0; q < 8; ++q) m.v[q] = -INFINITY;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const f8 v = unpack8(ldg16(x + ((static_cast<size_t>(n) * H + 2 * oh + a) * W + 2 * ow + b) * ldx + x_coff + u * 8));
#pragma unroll
        for (int q = 0; q < 8; ++q) m.v[q] = fmaxf(m.v[q], v.v[q]);
      }
    st16(y + static_cast<size_t>(pix) * ldy + y_coff + u * 8, pack8(m));
  }
}

// gradient goes to the first (row-major) element of the window that equals the maximum; input pixels outside
// every window (odd H / W) get zero
__global__ void maxpool2_bwd_kernel(const __nv_bfloat16* __restrict__ dy, int ldy, int y_coff,
                                    const __nv_bfloat16* __restrict__ x, int ldx, int x_coff,
                                    __nv_bfloat16* __restrict__ dx, int ldg, int g_coff, int N, int H, int W, int OH,
                                    int OW, int C) {
  const int U = C / 8;
  const long long total = static_cast<long long>(N) * OH * OW * U;
  PIXEL_UNIT_LOOP(total) {
    const int u = static_cast<int>(idx % U);
    const long long pix = idx / U;
    const int ow = static_cast<int>(pix % OW);
    const int oh = static_cast<int>((pix / OW) % OH);
    const int n = static_cast<int>(pix / (static_cast<long long>(OW) * OH));
    const f8 g = unpack8(ldg16(dy + static_cast<size_t>(pix) * ldy + y_coff + u * 8));
    f8 v[4];
#pragma unroll
    for (int t = 0; t < 4; ++t)
      v[t] = unpack8(ldg16(x + ((static_cast<size_t>(n) * H + 2 * oh + (t >> 1)) * W + 2 * ow + (t & 1)) * ldx + x_coff + u * 8));
    f8 o[4];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      int best = 0;
#pragma unroll
      for (int t = 1; t < 4; ++t)
        if (v[t].v[q] > v[best].v[q]) best = t;
#pragma unroll
      for (int t = 0; t < 4; ++t) o[t].v[q] = t == best ? g.v[q] : 0.f;
    }
#pragma unroll
    for (int t = 0; t < 4; ++t)
      st16(dx + ((static_cast<size_t>(n) * H + 2 * oh + (t >> 1)) * W + 2 * ow + (t & 1)) * ldg + g_coff + u * 8, pack8(o[t]));
  }
}

// ------------------------------------------------------------------------------------------------
// one-hot label map + 4-neighbour instance edges (SPADEModel.preprocess_input / get_edges)
// ------------------------------------------------------------------------------------------------
__global__ void onehot_edges_kernel(const int32_t* __restrict__ label, const int32_t* __restrict__ inst, int N, int H,
                                    int W, int n_label, int with_edge, __nv_bfloat16* __restrict__ y, int ldy,
                                    int y_coff, int C) {
  const int U = C / 8;
  const long long total = static_cast<long long>(N) * H * W * U;
  PIXEL_UNIT_LOOP(total) {
    const int u = static_cast<int>(idx % U);
    const long long pix = idx / U;
    const int w = static_cast<int>(pix % W);
    const int h = static_cast<int>((pix / W) % H);
    const int lab = label[pix];
    f8 o;
#pragma unroll
    for (int q = 0; q < 8; ++q) o.v[q] = (u * 8 + q == lab && lab < n_label) ? 1.f : 0.f;
    if (with_edge && n_label >= u * 8 && n_label < u * 8 + 8) {
      const int t = inst[pix];
      bool e = false;
      if (w > 0) e |= inst[pix - 1] != t;
      if (w < W - 1) e |= inst[pix + 1] != t;
      if (h > 0) e |= inst[pix - W] != t;
      if (h < H - 1) e |= inst[pix + W] != t;
      o.v[n_label - u * 8] = e ? 1.f : 0.f;
    }
    st16(y + static_cast<size_t>(pix) * ldy + y_coff + u * 8, pack8(o));
  }
}

// ------------------------------------------------------------------------------------------------
// fp32 vector helpers: out[i] = sum_k arena[idx[k*n+i]] (idx < 0: skipped); the adjoint scatter-add;
// shift[i] += bias[i] * scale[i] (bias in front of an eval-mode BatchNorm folded into its shift)
// ------------------------------------------------------------------------------------------------
__global__ void gather_sum_kernel(const float* __restrict__ arena, const int32_t* __restrict__ idx, int K, int n,
                                  float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = 0.f;
  for (int k = 0; k < K; ++k) {
    const int j = idx[k * n + i];
    if (j >= 0) v += arena[j];
  }
  out[i] = v;
}

__global__ void scatter_add_kernel(const float* __restrict__ src, const int32_t* __restrict__ idx, int K, int n,
                                   float* __restrict__ arena) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = src[i];
  for (int k = 0; k < K; ++k) {
    const int j = idx[k * n + i];
    if (j >= 0) atomicAdd(arena + j, v);
  }
}

__global__ void fma_vec_kernel(float* __restrict__ shift, const float* __restrict__ bias, const float* __restrict__ scale, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) shift[i] += bias[i] * scale[i];
}

// ------------------------------------------------------------------------------------------------
// spectral normalisation (torch.nn.utils.spectral_norm, one power iteration), table driven:
// blockIdx.y selects the weight
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum_all(float v, float* sh) {  // result broadcast to every thread
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < static_cast<int>(blockDim.x >> 5); ++i) t += sh[i];
  __syncthreads();
  return t;
}

// t[c] = sum_r W[r,c] u[r]   (threads <-> columns: coalesced rows)
__global__ void sn_wtu_kernel(const catb_sn_desc* __restrict__ tab, const float* __restrict__ arena,
                              const float* __restrict__ bufs, float* __restrict__ tmp, int tmp_stride) {
  const catb_sn_desc d = tab[blockIdx.y];
  const float* W = arena + d.w_off;
  const float* u = bufs + d.u_off;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < d.cols; c += gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int r = 0; r < d.rows; ++r) acc += W[static_cast<size_t>(r) * d.cols + c] * u[r];
    tmp[static_cast<size_t>(blockIdx.y) * tmp_stride + c] = acc;
  }
}

// dst = t / max(|t|, eps); optionally sigma = dst . t  (one block per weight)
__global__ void sn_normalize_kernel(const catb_sn_desc* __restrict__ tab, const float* __restrict__ tmp, int tmp_stride,
                                    float* __restrict__ bufs, int which /*0: v, 1: u*/, float* __restrict__ sigma) {
  __shared__ float sh[32];
  const catb_sn_desc d = tab[blockIdx.x];
  const int n = which ? d.rows : d.cols;
  const float* t = tmp + static_cast<size_t>(blockIdx.x) * tmp_stride;
  float* dst = bufs + (which ? d.u_off : d.v_off);
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += t[i] * t[i];
  const float nrm = sqrtf(block_sum_all(acc, sh));
  const float inv = 1.f / fmaxf(nrm, 1e-12f);
  float dot = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = t[i] * inv;
    dst[i] = v;
    dot += v * t[i];
  }
  if (sigma != nullptr) {
    const float sg = block_sum_all(dot, sh);
    if (threadIdx.x == 0) sigma[blockIdx.x] = sg;
  }
}

// s[r] = sum_c W[r,c] v[c]   (one warp per row)
__global__ void sn_wv_kernel(const catb_sn_desc* __restrict__ tab, const float* __restrict__ arena,
                             const float* __restrict__ bufs, float* __restrict__ tmp, int tmp_stride) {
  const catb_sn_desc d = tab[blockIdx.y];
  const float* W = arena + d.w_off;
  const float* v = bufs + d.v_off;
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < d.rows; r += warps) {
    float acc = 0.f;
    for (int c = lane; c < d.cols; c += 32) acc += W[static_cast<size_t>(r) * d.cols + c] * v[c];
    acc = warp_sum(acc);
    if (lane == 0) tmp[static_cast<size_t>(blockIdx.y) * tmp_stride + r] = acc;
  }
}

// eval mode: sigma = u . (W v) with the stored vectors (tmp holds W v)
__global__ void sn_sigma_kernel(const catb_sn_desc* __restrict__ tab, const float* __restrict__ tmp, int tmp_stride,
                                const float* __restrict__ bufs, float* __restrict__ sigma) {
  __shared__ float sh[32];
  const catb_sn_desc d = tab[blockIdx.x];
  float dot = 0.f;
  for (int i = threadIdx.x; i < d.rows; i += blockDim.x) dot += bufs[d.u_off + i] * tmp[static_cast<size_t>(blockIdx.x) * tmp_stride + i];
  const float sg = block_sum_all(dot, sh);
  if (threadIdx.x == 0) sigma[blockIdx.x] = sg;
}

__global__ void sn_scale_kernel(const catb_sn_desc* __restrict__ tab, const float* __restrict__ arena,
                                const float* __restrict__ sigma, float* __restrict__ w_eff) {
  const catb_sn_desc d = tab[blockIdx.y];
  const float inv = 1.f / sigma[blockIdx.y];
  const int n = d.rows * d.cols;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    w_eff[d.w_off + i] = arena[d.w_off + i] * inv;
}

// cdot[d] += <g, w_eff> over the weight (caller zeroes cdot)
__global__ void sn_bwd_dot_kernel(const catb_sn_desc* __restrict__ tab, const float* __restrict__ grad,
                                  const float* __restrict__ w_eff, float* __restrict__ cdot) {
  __shared__ float sh[32];
  const catb_sn_desc d = tab[blockIdx.y];
  const int n = d.rows * d.cols;
  float acc = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    acc += grad[d.w_off + i] * w_eff[d.w_off + i];
  const float t = block_sum_all(acc, sh);
  if (threadIdx.x == 0) atomicAdd(cdot + blockIdx.y, t);
}

// g <- (g - cdot * u v^T) / sigma     (gradient w.r.t. weight_orig from the gradient w.r.t. W / sigma)
__global__ void sn_bwd_apply_kernel(const catb_sn_desc* __restrict__ tab, float* __restrict__ grad,
                                    const float* __restrict__ bufs, const float* __restrict__ sigma,
                                    const float* __restrict__ cdot) {
  const catb_sn_desc d = tab[blockIdx.y];
  const float inv = 1.f / sigma[blockIdx.y], c = cdot[blockIdx.y];
  const int n = d.rows * d.cols;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int r = i / d.cols, col = i - r * d.cols;
    grad[d.w_off + i] = (grad[d.w_off + i] - c * bufs[d.u_off + r] * bufs[d.v_off + col]) * inv;
  }
}

}  // namespace catb

using namespace catb;

#define CHK_SLICE(ld, coff, C)                                                                     \
  CATB_REQUIRE((ld) % 8 == 0 && (coff) % 8 == 0 && (C) % 8 == 0 && (C) > 0 && (coff) + (C) <= (ld), \
               "bad channel slice (ld=%d coff=%d C=%d)", (int)(ld), (int)(coff), (int)(C))
#define BF(p) static_cast<const __nv_bfloat16*>(p)
#define BFM(p) static_cast<__nv_bfloat16*>(p)

extern "C" int catb_resize_nearest(const void* x, int ldx, int x_coff, int H, int W, void* y, int ldy, int y_coff, int N,
                                   int OH, int OW, int C, catb_stream_t s) {
  CHK_SLICE(ldx, x_coff, C);
  CHK_SLICE(ldy, y_coff, C);
  const long long total = static_cast<long long>(N) * OH * OW * (C / 8);
  resize_nearest_kernel<<<grid_for(total, 256), 256, 0, S(s)>>>(BF(x), ldx, x_coff, H, W, BFM(y), ldy, y_coff, N, OH, OW, C);
  return check_launch("resize_nearest");
}

extern "C" int catb_upsample2x_bwd(const void* dy, int ldy, int y_coff, void* dx, int ldx, int x_coff, int N, int H, int W,
                                   int C, catb_stream_t s) {
  CHK_SLICE(ldx, x_coff, C);
  CHK_SLICE(ldy, y_coff, C);
  const long long total = static_cast<long long>(N) * H * W * (C / 8);
  upsample2x_bwd_kernel<<<grid_for(total, 256), 256, 0, S(s)>>>(BF(dy), ldy, y_coff, BFM(dx), ldx, x_coff, N, H, W, C);
  return check_launch("upsample2x_bwd");
}

extern "C" int catb_spade_modulate(const void* x, int ldx, int x_coff, const void* gamma, int ldg, int g_coff,
                                   const void* beta, int ldb, int b_coff, void* y, int ldy, int y_coff, long long pixels,
                                   int C, const float* scale, const float* shift, int act, catb_stream_t s) {
  CHK_SLICE(ldx, x_coff, C);
  CHK_SLICE(ldg, g_coff, C);
  CHK_SLICE(ldb, b_coff, C);
  CHK_SLICE(ldy, y_coff, C);
  spade_modulate_kernel<<<grid_for(pixels * (C / 8), 256), 256, 0, S(s)>>>(BF(x), ldx, x_coff, BF(gamma), ldg, g_coff, BF(beta),
                                                                          ldb, b_coff, BFM(y), ldy, y_coff, pixels, C, scale,
                                                                          shift, act);
  return check_launch("spade_modulate");
}

extern "C" int catb_spade_modulate_bwd(const void* dy, int ldd, int d_coff, const void* y, int ldy, int y_coff, const void* x,
                                       int ldx, int x_coff, const void* gamma, int ldg, int g_coff, void* dgamma, int ldo,
                                       int o_coff, void* dbeta, int ldp, int p_coff, void* dn, int ldn, int n_coff,
                                       long long pixels, int C, const float* scale, const float* shift, int act,
                                       catb_stream_t s) {
  CHK_SLICE(ldd, d_coff, C);
  CHK_SLICE(ldy, y_coff, C);
  CHK_SLICE(ldx, x_coff, C);
  CHK_SLICE(ldg, g_coff, C);
  CHK_SLICE(ldo, o_coff, C);
  CHK_SLICE(ldp, p_coff, C);
  CHK_SLICE(ldn, n_coff, C);
  spade_modulate_bwd_kernel<<<grid_for(pixels * (C / 8), 256), 256, 0, S(s)>>>(
      BF(dy), ldd, d_coff, BF(y), ldy, y_coff, BF(x), ldx, x_coff, BF(gamma), ldg, g_coff, BFM(dgamma), ldo, o_coff,
      BFM(dbeta), ldp, p_coff, BFM(dn), ldn, n_coff, pixels, C, scale, shift, act);
  return check_launch("spade_modulate_bwd");
}

extern "C" int catb_act_fwd(const void* x, int ldx, int x_coff, void* y, int ldy, int y_coff, long long pixels, int C, int act,
                            catb_stream_t s) {
  CHK_SLICE(ldx, x_coff, C);
  CHK_SLICE(ldy, y_coff, C);
  act_fwd_kernel<<<grid_for(pixels * (C / 8), 256), 256, 0, S(s)>>>(BF(x), ldx, x_coff, BFM(y), ldy, y_coff, pixels, C, act);
  return check_launch("act_fwd");
}

extern "C" int catb_avgpool3s2(const void* x, int ldx, int x_coff, int H, int W, void* y, int ldy, int y_coff, int N, int C,
                               catb_stream_t s) {
  CHK_SLICE(ldx, x_coff, C);
  CHK_SLICE(ldy, y_coff, C);
  const int OH = (H + 1) / 2, OW = (W + 1) / 2;
  const long long total = static_cast<long long>(N) * OH * OW * (C / 8);
  avgpool3s2_kernel<<<grid_for(total, 256), 256, 0, S(s)>>>(BF(x), ldx, x_coff, H, W, BFM(y), ldy, y_coff, N, OH, OW, C);
  return check_launch("avgpool3s2");
}

extern "C" int catb_avgpool3s2_bwd(const void* dy, int ldy, int y_coff, const void* add, int lda, int a_coff, void* dx, int ldx,
                                   int x_coff, int N, int H, int W, int C, catb_stream_t s) {
  CHK_SLICE(ldx, x_coff, C);
  CHK_SLICE(ldy, y_coff, C);
  const int OH = (H + 1) / 2, OW = (W + 1) / 2;
  const long long total = static_cast<long long>(N) * H * W * (C / 8);
  avgpool3s2_bwd_kernel<<<grid_for(total, 256), 256, 0, S(s)>>>(BF(dy), ldy, y_coff, OH, OW, BF(add), lda, a_coff, BFM(dx), ldx,
                                                                x_coff, N, H, W, C);
  return check_launch("avgpool3s2_bwd");
}

extern "C" int catb_maxpool2(const void* x, int ldx, int x_coff, int H, int W, void* y, int ldy, int y_coff, int N, int C,
                             catb_stream_t s) {
  CHK_SLICE(ldx, x_coff, C);
  CHK_SLICE(ldy, y_coff, C);
  const int OH = H / 2, OW = W / 2;
  CATB_REQUIRE(OH > 0 && OW > 0, "max pool of a %dx%d map", H, W);
  const long long total = static_cast<long long>(N) * OH * OW * (C / 8);
  maxpool2_kernel<<<grid_for(total, 256), 256, 0, S(s)>>>(BF(x), ldx, x_coff, H, W, BFM(y), ldy, y_coff, N, OH, OW, C);
  return check_launch("maxpool2");
}

extern "C" int catb_maxpool2_bwd(const void* dy, int ldy, int y_coff, const void* x, int ldx, int x_coff, void* dx, int ldg,
                                 int g_coff, int N, int H, int W, int C, catb_stream_t s) {
  CHK_SLICE(ldx, x_coff, C);
  CHK_SLICE(ldy, y_coff, C);
  CHK_SLICE(ldg, g_coff, C);
  CATB_REQUIRE(H % 2 == 0 && W % 2 == 0, "max-pool backward expects even extents (got %dx%d)", H, W);
  const int OH = H / 2, OW = W / 2;
  const long long total = static_cast<long long>(N) * OH * OW * (C / 8);
  maxpool2_bwd_kernel<<<grid_for(total, 256), 256, 0, S(s)>>>(BF(dy), ldy, y_coff, BF(x), ldx, x_coff, BFM(dx), ldg, g_coff, N,
                                                              H, W, OH, OW, C);
  return check_launch("maxpool2_bwd");
}

extern "C" int catb_onehot_edges(const int32_t* label, const int32_t* instance, int N, int H, int W, int n_label, void* y,
                                 int ldy, int y_coff, int C, catb_stream_t s) {
  CHK_SLICE(ldy, y_coff, C);
  CATB_REQUIRE(n_label + (instance != nullptr ? 1 : 0) <= C, "one-hot needs %d channels, slice has %d", n_label + 1, C);
  const long long total = static_cast<long long>(N) * H * W * (C / 8);
  onehot_edges_kernel<<<grid_for(total, 256), 256, 0, S(s)>>>(label, instance, N, H, W, n_label, instance != nullptr ? 1 : 0,
                                                              BFM(y), ldy, y_coff, C);
  return check_launch("onehot_edges");
}

extern "C" int catb_gather_sum_f32(const float* arena, const int32_t* idx, int K, int n, float* out, catb_stream_t s) {
  CATB_REQUIRE(K > 0 && n > 0, "empty gather");
  gather_sum_kernel<<<(n + 127) / 128, 128, 0, S(s)>>>(arena, idx, K, n, out);
  return check_launch("gather_sum_f32");
}

extern "C" int catb_scatter_add_f32(const float* src, const int32_t* idx, int K, int n, float* arena, catb_stream_t s) {
  CATB_REQUIRE(K > 0 && n > 0, "empty scatter");
  scatter_add_kernel<<<(n + 127) / 128, 128, 0, S(s)>>>(src, idx, K, n, arena);
  return check_launch("scatter_add_f32");
}

extern "C" int catb_fma_vec(float* shift, const float* bias, const float* scale, int n, catb_stream_t s) {
  fma_vec_kernel<<<(n + 127) / 128, 128, 0, S(s)>>>(shift, bias, scale, n);
  return check_launch("fma_vec");
}

extern "C" int catb_sn_forward(const catb_sn_desc* table, int n, int max_rows, int max_cols, const float* arena, float* bufs,
                               int training, float* tmp, float* sigma, float* w_eff, catb_stream_t s) {
  CATB_REQUIRE(n > 0 && max_rows > 0 && max_cols > 0, "empty spectral-norm table");
  const int stride = max_rows > max_cols ? max_rows : max_cols;
  if (training) {
    sn_wtu_kernel<<<dim3((max_cols + 127) / 128, n), 128, 0, S(s)>>>(table, arena, bufs, tmp, stride);
    sn_normalize_kernel<<<n, 256, 0, S(s)>>>(table, tmp, stride, bufs, 0, nullptr);
  }
  sn_wv_kernel<<<dim3((max_rows * 32 + 255) / 256, n), 256, 0, S(s)>>>(table, arena, bufs, tmp, stride);
  if (training)
    sn_normalize_kernel<<<n, 256, 0, S(s)>>>(table, tmp, stride, bufs, 1, sigma);
  else
    sn_sigma_kernel<<<n, 256, 0, S(s)>>>(table, tmp, stride, bufs, sigma);
  const long long mx = static_cast<long long>(max_rows) * max_cols;
  sn_scale_kernel<<<dim3(grid_for(mx, 256, 148), n), 256, 0, S(s)>>>(table, arena, sigma, w_eff);
  return check_launch("sn_forward");
}

extern "C" int catb_sn_backward(const catb_sn_desc* table, int n, int max_rows, int max_cols, float* grad, const float* w_eff,
                                const float* bufs, const float* sigma, float* cdot /* [n], zeroed here */, catb_stream_t s) {
  CATB_REQUIRE(n > 0 && max_rows > 0 && max_cols > 0, "empty spectral-norm table");
  if (cudaMemsetAsync(cdot, 0, sizeof(float) * n, S(s)) != cudaSuccess) return check_launch("sn_backward memset");
  const long long mx = static_cast<long long>(max_rows) * max_cols;
  const dim3 grid(grid_for(mx, 256, 148), n);
  sn_bwd_dot_kernel<<<grid, 256, 0, S(s)>>>(table, grad, w_eff, cdot);
  sn_bwd_apply_kernel<<<grid, 256, 0, S(s)>>>(table, grad, bufs, sigma, cdot);
  return check_launch("sn_backward");
}
