// Helper kernels of the "x-packed" 7x7 stem / head convolutions of InceptionGenerator (inception_generator.py:37-56,
// 130-134).  A 7x7 conv with 3 input (stem) or 3 output (head) channels wastes a 128 x N x 16 tensor-core instruction per
// tap; packing the seven horizontal taps into the channel dimension turns it into a 7x1 conv with 21 (-> 24) channels on
// that side, i.e. 7x fewer GEMM steps, at the price of one cheap element-wise pass:
//   stem:  X'[r, c, dx*3 + ci] = X[r, reflect(c + dx - 3), ci]                              (catb_expand_x)
//   head:  P'[r, c', co*8 + dx] = sum_{dy, ci} X[reflect(r + dy - 3), reflect(c' - 3), ci] W[co, ci, dy, dx]   (GEMM, c' in [0, W+6))
//          out[r, c, co] = act(bias[co] + sum_dx P'[r, c + dx, co*8 + dx])                  (catb_shift_sum)
//          dP'[r, c', co*8 + dx] = dz[r, c' - dx, co]  (0 outside)                          (catb_shift_expand, its adjoint)
#include "common.cuh"

namespace catb {

static inline cudaStream_t S(catb_stream_t s) { return static_cast<cudaStream_t>(s); }
static inline int grid_for(long long work, int block, int max_blocks = 148 * 16) {
  long long g = (work + block - 1) / block;
  if (g < 1) g = 1;
  if (g > max_blocks) g = max_blocks;
  return static_cast<int>(g);
}

// x: [N,H,W,ldx] (Cin real channels at x_coff), y: [N,H,W,ldy] with taps*Cin channels (+ zero padding up to Cy)
__global__ void expand_x_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int x_coff, __nv_bfloat16* __restrict__ y, int ldy,
                                int y_coff, int N, int H, int W, int Cin, int taps, int Cy) {
  const int U = Cy / 8;
  const long long total = static_cast<long long>(N) * H * W * U;
  const int p = (taps - 1) / 2;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int u = static_cast<int>(idx % U);
    const long long pix = idx / U;
    const int w = static_cast<int>(pix % W);
    const long long row = pix / W;  // n*H + h
    f8 o;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int c = u * 8 + q;
      float v = 0.f;
      if (c < taps * Cin) {
        const int dx = c / Cin, ci = c - dx * Cin;
        const int iw = reflect_idx(w + dx - p, W);
        v = __bfloat162float(x[(static_cast<size_t>(row) * W + iw) * ldx + x_coff + ci]);
      }
      o.v[q] = v;
    }
    st16(y + static_cast<size_t>(pix) * ldy + y_coff + u * 8, pack8(o));
  }
}

// P: [N,H,W+taps-1,ldp] with channel co*8+dx; out: [N,H,W,ldo] (Cout <= 8 real channels, the rest zero)
__global__ void shift_sum_kernel(const __nv_bfloat16* __restrict__ P, int ldp, int p_coff, __nv_bfloat16* __restrict__ out, int ldo,
                                 int o_coff, int N, int H, int W, int Cout, int taps, const float* __restrict__ bias, int act) {
  const long long total = static_cast<long long>(N) * H * W;
  const int Wp = W + taps - 1;
  for (long long pix = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; pix < total;
       pix += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int w = static_cast<int>(pix % W);
    const long long row = pix / W;
    f8 o;
#pragma unroll
    for (int q = 0; q < 8; ++q) o.v[q] = 0.f;
    for (int co = 0; co < Cout; ++co) {
      float acc = bias != nullptr ? bias[co] : 0.f;
      for (int dx = 0; dx < taps; ++dx) {
        // the 8 channels co*8 .. co*8+7 of pixel (row, w+dx) are one 16-byte unit; take element dx
        const f8 v = unpack8(ldg16(P + (static_cast<size_t>(row) * Wp + w + dx) * ldp + p_coff + co * 8));
        acc += v.v[dx];
      }
      o.v[co] = apply_act(acc, act);
    }
    st16(out + static_cast<size_t>(pix) * ldo + o_coff, pack8(o));
  }
}

// dP[r, c', co*8 + dx] = dz[r, c' - dx, co] for 0 <= c' - dx < W, else 0
__global__ void shift_expand_kernel(const __nv_bfloat16* __restrict__ dz, int ldz, int z_coff, __nv_bfloat16* __restrict__ dP, int ldp,
                                    int p_coff, int N, int H, int W, int Cout, int taps) {
  const int Wp = W + taps - 1;
  const long long total = static_cast<long long>(N) * H * Wp * Cout;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int co = static_cast<int>(idx % Cout);
    const long long ppix = idx / Cout;
    const int cp = static_cast<int>(ppix % Wp);
    const long long row = ppix / Wp;
    f8 o;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int c = cp - q;
      o.v[q] = (q < taps && c >= 0 && c < W) ? __bfloat162float(dz[(static_cast<size_t>(row) * W + c) * ldz + z_coff + co]) : 0.f;
    }
    st16(dP + static_cast<size_t>(ppix) * ldp + p_coff + co * 8, pack8(o));
  }
}

// ---- "tap-split" form of a conv with ONE output channel (PatchGAN head, discriminators.py:72-73) -------------------
// A 4x4 conv Cin -> 1 on a 128 x 16 tensor-core tile wastes 15/16 of every instruction and re-reads each input pixel
// once per tap.  Instead  P[n, iy, ix, t] = sum_c X[n, iy, ix, c] W[0, c, t]  is ONE 1x1 GEMM with the R*S taps as its
// output channels (each input pixel read once), and the conv is the shifted sum of P below; the adjoint spreads dY
// over the taps (tap_expand), after which the weight / input gradients are 1x1 GEMMs as well.
__global__ void tap_sum_kernel(const float* __restrict__ P, int ldp, int p_coff, float* __restrict__ out, int ldo, int o_coff,
                               int N, int H, int W, int OH, int OW, int R, int S, int pad, const float* __restrict__ bias) {
  const long long total = static_cast<long long>(N) * OH * OW;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ox = static_cast<int>(idx % OW);
    const int oy = static_cast<int>((idx / OW) % OH);
    const long long n = idx / (static_cast<long long>(OW) * OH);
    float acc = bias != nullptr ? bias[0] : 0.f;
    for (int r = 0; r < R; ++r) {
      const int iy = oy + r - pad;
      if (iy < 0 || iy >= H) continue;
      for (int s = 0; s < S; ++s) {
        const int ix = ox + s - pad;
        if (ix < 0 || ix >= W) continue;
        acc += P[((n * H + iy) * W + ix) * ldp + p_coff + r * S + s];
      }
    }
    out[idx * ldo + o_coff] = acc;
  }
}

// dP[n, iy, ix, r*S + s] = dY[n, iy - r + pad, ix - s + pad] (0 outside the output), one 8-tap unit per thread
__global__ void tap_expand_kernel(const __nv_bfloat16* __restrict__ dy, int ldy, int y_coff, __nv_bfloat16* __restrict__ dP, int ldp,
                                  int p_coff, int N, int H, int W, int OH, int OW, int R, int S, int pad) {
  const int U = (R * S + 7) / 8;
  const long long total = static_cast<long long>(N) * H * W * U;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int u = static_cast<int>(idx % U);
    const long long pix = idx / U;
    const int ix = static_cast<int>(pix % W);
    const int iy = static_cast<int>((pix / W) % H);
    const long long n = pix / (static_cast<long long>(W) * H);
    f8 o;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int t = u * 8 + q;
      const int r = t / S, s = t - r * S;
      const int oy = iy - r + pad, ox = ix - s + pad;
      o.v[q] = (t < R * S && oy >= 0 && oy < OH && ox >= 0 && ox < OW)
                   ? __bfloat162float(dy[((n * OH + oy) * OW + ox) * ldy + y_coff]) : 0.f;
    }
    st16(dP + pix * ldp + p_coff + u * 8, pack8(o));
  }
}

}  // namespace catb

using namespace catb;

#define CHK_SLICE(ld, coff, C)                                                                     \
  CATB_REQUIRE((ld) % 8 == 0 && (coff) % 8 == 0 && (C) % 8 == 0 && (C) > 0 && (coff) + (C) <= (ld), \
               "bad channel slice (ld=%d coff=%d C=%d)", (int)(ld), (int)(coff), (int)(C))

extern "C" int catb_expand_x(const void* x, int ldx, int x_coff, void* y, int ldy, int y_coff, int N, int H, int W, int Cin,
                             int taps, int Cy, catb_stream_t s) {
  CHK_SLICE(ldy, y_coff, Cy);
  CATB_REQUIRE(ldx % 8 == 0 && x_coff % 8 == 0 && Cin > 0 && taps % 2 == 1 && taps * Cin <= Cy && taps / 2 < W,
               "bad x-expansion (Cin=%d taps=%d Cy=%d W=%d)", Cin, taps, Cy, W);
  const long long total = static_cast<long long>(N) * H * W * (Cy / 8);
  expand_x_kernel<<<grid_for(total, 256), 256, 0, S(s)>>>(static_cast<const __nv_bfloat16*>(x), ldx, x_coff,
                                                          static_cast<__nv_bfloat16*>(y), ldy, y_coff, N, H, W, Cin, taps, Cy);
  return check_launch("expand_x");
}

extern "C" int catb_shift_sum(const void* P, int ldp, int p_coff, void* out, int ldo, int o_coff, int N, int H, int W, int Cout,
                              int taps, const float* bias, int act, catb_stream_t s) {
  CATB_REQUIRE(ldp % 8 == 0 && p_coff % 8 == 0 && ldo % 8 == 0 && o_coff % 8 == 0 && Cout >= 1 && Cout <= 8 && taps >= 1 &&
                   taps <= 8 && p_coff + Cout * 8 <= ldp && o_coff + 8 <= ldo,
               "bad shift-sum (Cout=%d taps=%d)", Cout, taps);
  const long long total = static_cast<long long>(N) * H * W;
  shift_sum_kernel<<<grid_for(total, 256), 256, 0, S(s)>>>(static_cast<const __nv_bfloat16*>(P), ldp, p_coff,
                                                           static_cast<__nv_bfloat16*>(out), ldo, o_coff, N, H, W, Cout, taps, bias, act);
  return check_launch("shift_sum");
}

extern "C" int catb_shift_expand(const void* dz, int ldz, int z_coff, void* dP, int ldp, int p_coff, int N, int H, int W, int Cout,
                                 int taps, catb_stream_t s) {
  CATB_REQUIRE(ldp % 8 == 0 && p_coff % 8 == 0 && ldz % 8 == 0 && z_coff % 8 == 0 && Cout >= 1 && Cout <= 8 && taps >= 1 &&
                   taps <= 8 && p_coff + Cout * 8 <= ldp,
               "bad shift-expand (Cout=%d taps=%d)", Cout, taps);
  const long long total = static_cast<long long>(N) * H * (W + taps - 1) * Cout;
  shift_expand_kernel<<<grid_for(total, 256), 256, 0, S(s)>>>(static_cast<const __nv_bfloat16*>(dz), ldz, z_coff,
                                                              static_cast<__nv_bfloat16*>(dP), ldp, p_coff, N, H, W, Cout, taps);
  return check_launch("shift_expand");
}

extern "C" int catb_tap_sum(const float* P, int ldp, int p_coff, float* out, int ldo, int o_coff, int N, int H, int W, int OH,
                            int OW, int R, int S, int pad, const float* bias, catb_stream_t s) {
  CATB_REQUIRE(R >= 1 && S >= 1 && p_coff + R * S <= ldp && o_coff < ldo && OH == H + 2 * pad - R + 1 && OW == W + 2 * pad - S + 1,
               "bad tap-sum geometry (R=%d S=%d pad=%d)", R, S, pad);
  const long long total = static_cast<long long>(N) * OH * OW;
  tap_sum_kernel<<<grid_for(total, 128), 128, 0, static_cast<cudaStream_t>(s)>>>(P, ldp, p_coff, out, ldo, o_coff, N, H, W, OH, OW, R, S, pad, bias);
  return check_launch("tap_sum");
}

extern "C" int catb_tap_expand(const void* dy, int ldy, int y_coff, void* dP, int ldp, int p_coff, int N, int H, int W, int OH,
                               int OW, int R, int S, int pad, catb_stream_t s) {
  CATB_REQUIRE(R >= 1 && S >= 1 && ldp % 8 == 0 && p_coff % 8 == 0 && p_coff + (R * S + 7) / 8 * 8 <= ldp && y_coff < ldy &&
                   OH == H + 2 * pad - R + 1 && OW == W + 2 * pad - S + 1,
               "bad tap-expand geometry (R=%d S=%d pad=%d)", R, S, pad);
  const long long total = static_cast<long long>(N) * H * W * ((R * S + 7) / 8);
  tap_expand_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(s)>>>(static_cast<const __nv_bfloat16*>(dy), ldy, y_coff,
                                                            static_cast<__nv_bfloat16*>(dP), ldp, p_coff, N, H, W, OH, OW, R, S, pad);
  return check_launch("tap_expand");
}
