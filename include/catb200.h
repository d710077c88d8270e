/*
 * catb200.h -- C ABI of libcatb200.so: hand-written sm_100a kernels for the CAT distillation hot path.
 *
 * The reference (snap-research/CAT) has no FFI/plugin layer of its own: every op on the path is a stock
 * ATen call made from Python.  The entry points below are therefore the functions a
 * reference-side binding (ctypes) would call *instead of* those ATen calls; each one
 * cites the reference call site it replaces (paths relative to the reference root).
 *
 * Contract: the caller owns every buffer (including workspaces), the callee never
 * allocates, never synchronises, only enqueues on the given stream, and returns 0 or a negative
 * catb_status.  No C++ exceptions cross the boundary.  All descriptors are POD.
 *
 * Data layout: activations are NHWC bf16 with the channel count padded to a multiple of 8 ("Cp");
 * padding channels are always zero.  A tensor may be a channel slice of a wider buffer: `ld*` is the
 * pixel pitch in elements and `*_coff` the first channel of the slice (multiple of 8).
 * Parameters, gradients and optimiser state are fp32 in the reference's own (PyTorch) layout.
 */
#ifndef CATB200_H_
#define CATB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* catb_stream_t; /* cudaStream_t */

typedef enum {
  CATB_OK = 0,
  CATB_ERR_INVALID = -1, /* bad argument / unsupported shape */
  CATB_ERR_CUDA = -2,    /* a CUDA runtime call failed; see catb_last_error_string() */
  CATB_ERR_NO_DEVICE = -3
} catb_status;

/* CATB_ACT_LEAKY001 = nn.LeakyReLU() at its default slope: the generator activation of SPADE TEACHER training
 * (models/spade_model.py:92 sets active_fn='nn.LeakyReLU'; the distillers keep nn.ReLU, options/distill_options.py:123). */
enum { CATB_ACT_NONE = 0, CATB_ACT_RELU = 1, CATB_ACT_LEAKY02 = 2, CATB_ACT_TANH = 3, CATB_ACT_LEAKY001 = 4 };
enum { CATB_PAD_ZERO = 0, CATB_PAD_REFLECT = 1 };
enum { CATB_GAN_HINGE = 0, CATB_GAN_LSGAN = 1, CATB_GAN_VANILLA = 2 };

/* One 16-byte K-unit (8 consecutive channels of one tap) of an implicit-GEMM operand. */
typedef struct {
  int8_t dr, ds; /* tap offset added to (oh*sn, ow*sn) before the division by sd           */
  int16_t cu;    /* channel offset inside the gathered pixel, in units of 8 channels        */
} catb_gather_unit;

/* Where the 8 k-elements of a unit live in the fp32 parameter / gradient arena. */
typedef struct {
  int32_t w_off;  /* element offset of (row 0, first channel of the unit, this tap)          */
  int32_t sn_w;   /* element stride between GEMM rows (output channels)                      */
  int32_t sc_w;   /* element stride between the 8 channels of the unit                       */
  int32_t nvalid; /* number of real channels in the unit (0..8); the rest is zero padding    */
} catb_weight_unit;

/* Geometry shared by the implicit-GEMM kernels.  The "gathered" tensor X is read through the unit
 * table; the "lattice" tensor Y is addressed one GEMM row per lattice point:
 *   row m  <->  (n, i, j),  oh = o_ph + i*o_step,  ow = o_pw + j*o_step,
 *   unit u <->  X[n, (oh*sn + dr_u)/sd, (ow*sn + ds_u)/sd, x_coff + 8*cu_u .. +8)
 * with zero fill when the division is inexact or the coordinate is outside [0,H)x[0,W)
 * (CATB_PAD_ZERO) or mirrored back inside (CATB_PAD_REFLECT, nn.ReflectionPad2d semantics). */
typedef struct {
  int32_t N, H, W;        /* gathered tensor X: batch and spatial size                       */
  int32_t ldx, x_coff;    /* X pixel pitch (elements) and slice offset                       */
  int32_t OH, OW;         /* lattice tensor Y: full spatial size                             */
  int32_t ldy, y_coff;    /* Y pixel pitch and slice offset                                  */
  int32_t o_step, o_ph, o_pw, OHs, OWs; /* sub-lattice covered by this launch                */
  int32_t sn, sd;         /* gather numerator / denominator (1 or 2)                         */
  int32_t pad_mode;       /* CATB_PAD_*                                                      */
  int32_t n_units;        /* length of the unit tables; GEMM K = 8*n_units                   */
  int32_t n_rows;         /* GEMM rows on the weight side: real output channels (fprop) /
                             real lattice-tensor channels (wgrad)                            */
  int32_t n_tile;         /* rows of the packed weight image per tile (multiple of 16, <=256)*/
  int32_t act;            /* CATB_ACT_* applied in the fprop epilogue                        */
  int32_t accumulate;     /* fprop: add to the existing contents of Y                        */
  int32_t y_is_f32;       /* fprop: Y is fp32 instead of bf16                                */
  int32_t reserved;
} catb_igemm_desc;

/* ---- library ---------------------------------------------------------------------------- */
const char* catb_version(void);
const char* catb_last_error_string(void);
/* Sets kernel attributes (opt-in shared memory); call once per device before any launch. */
int catb_init(int device);
/* Bytes of the packed bf16 weight image used by catb_igemm_fprop for (n_rows, n_units, n_tile). */
size_t catb_packed_weight_bytes(int n_rows, int n_units, int n_tile);

/* ---- implicit-GEMM convolution (tcgen05 / TMEM) -----------------------------------------------
 * Replaces every dense F.conv2d / F.conv_transpose2d on the path and their input gradients:
 *   models/modules/inception_architecture/inception_generator.py:37-56,116-132 (7x7 reflect stem/head,
 *   3x3 s2 down, ConvTranspose 3x3 s2 up), models/modules/inception_modules.py:129-177 (block convs),
 *   models/modules/discriminators.py:39-74 (4x4 PatchGAN convs); autograd's conv backward-data. */
int catb_pack_weights(const catb_igemm_desc* d, const catb_weight_unit* wunits /*device*/,
                      const float* arena /*device*/, void* packed /*device*/, catb_stream_t s);
/* Row-segment variant for N-concatenated GEMMs (several convs sharing one gathered input, e.g. the six
 * first-stage convs of inception_modules.py:129-163): fills image rows [row0, row0+span) -- the first
 * `nreal` from this segment's weights, the rest with zeros -- using this segment's weight-unit table
 * (nvalid = 0 for taps the segment's kernel does not have).  d->n_rows is the row count of the whole image. */
int catb_pack_weights_rows(const catb_igemm_desc* d, const catb_weight_unit* wunits /*device*/, const float* arena,
                           void* packed, int row0, int span, int nreal, catb_stream_t s);
int catb_igemm_fprop(const catb_igemm_desc* d, const catb_gather_unit* units /*device*/, const void* x,
                     const void* packed_w, const float* bias /*nullable*/, void* y, catb_stream_t s);
/* Weight gradient: arena_grad[w] += sum_rows Y[row, c] * gather(X)[row, k]   (atomic fp32 adds).
 * Replaces autograd's conv backward-weight for the same call sites. */
int catb_igemm_wgrad(const catb_igemm_desc* d, const catb_gather_unit* units, const catb_weight_unit* wunits,
                     const void* x, const void* y, float* arena_grad, catb_stream_t s);
/* v2 forward kernel: the CTA stages each input pixel ONCE per 64-channel chunk in a shared-memory halo
 * tile and every filter tap addresses a shifted window of it (see cat_b200/csrc/igemm_halo.cu).  Same
 * packed weights and unit table as catb_igemm_fprop, with the table ordered chunk-major, 8 units per
 * (chunk, tap) step (cat_b200/igemm_plan.py: make_halo_plan).  Lattice positions are enumerated in pitch
 * space m = i*Wf + j inside vertical strips of TW lattice columns; frame pixel (fy, fx) of plane p is
 * input pixel (mul*(fy+y0[p]) + pa[p], mul*(fx+x0[p]+strip*TW) + pb[p]) under the padding rule. */
typedef struct {
  int32_t a_row; /* halo row of lattice position m0 for this step: plane*Lh + dy*Wf + dx */
  int32_t chunk; /* index into the chunk table                                               */
} catb_halo_step;
typedef struct {
  int32_t cu0, n_units;         /* first 8-channel unit of the chunk in the gathered pixel, units used */
  int32_t first_step, n_steps;  /* the chunk's (tap) steps are contiguous in the step table            */
} catb_halo_chunk;
typedef struct {
  int32_t n_steps, n_chunks, n_planes;
  int32_t plane_pa[4], plane_pb[4]; /* input parity of the plane (stride-2 convs), else 0              */
  int32_t plane_y0[4], plane_x0[4]; /* frame origin of the plane                                        */
  int32_t mul;                      /* frame -> input: (mul*(fy+y0)+pa, mul*(fx+x0+strip*TW)+pb)        */
  int32_t TW, n_strips;             /* lattice columns per vertical strip, strips per image             */
  int32_t Wf, Lh;                   /* frame pitch TW+Xmax; halo rows per plane 128*m_sub+Ymax*Wf+Xmax  */
  int32_t Ymax, Xmax;               /* tap extent in frame rows / columns                               */
  int32_t m_sub;                    /* 128-row sub-tiles per CTA (1 or 2) sharing every weight tile     */
  int32_t b_budget;                 /* bytes of weight tiles kept in flight (0: 64 KB); a smaller ring lets
                                       several CTAs share an SM when the tiles are short (thin GEMMs)    */
} catb_halo_desc;
/* 1 when the halo tile + weight ring + the step / chunk tables (kept in shared memory for the MMA-issuing warp)
 * fit in shared memory / TMEM for these parameters, else 0. */
int catb_igemm_halo_fits(int n_planes, int Lh, int n_tile, int m_sub, int n_steps, int n_chunks);
/* Optional fused statistics of the halo kernels' bf16 epilogue: the per-channel sum and sum of squares of the values the
 * launch stores, added atomically to sums[g][0][coff + column] / sums[g][1][coff + column] (g = image when per_sample,
 * else 0) -- the statistics pass of the InstanceNorm / BatchNorm layer that follows the conv
 * (models/modules/inception_modules.py:22-44 ConvBNReLU: conv + norm + activation), so catb_norm_stats is not launched. */
typedef struct {
  float* sums;        /* [G][2][C] fp32, zeroed by the caller before the producing launches */
  int32_t C, coff;    /* channels per statistics row; column of the GEMM's output channel 0 in it */
  int32_t per_sample; /* 1: one group per image (InstanceNorm); 0: one group (BatchNorm) */
  int32_t reserved;
} catb_epilogue_stats;
int catb_igemm_halo_fprop(const catb_igemm_desc* d, const catb_halo_desc* h, const catb_halo_step* steps /*device*/,
                          const catb_halo_chunk* chunks /*device*/, const void* x, const void* packed_w,
                          const float* bias /*nullable*/, void* y, const catb_epilogue_stats* stats /*nullable*/,
                          catb_stream_t s);

/* v3 forward kernel: the same halo GEMM as a persistent, warp-specialised pipeline (cat_b200/csrc/igemm_halo_persist.cu):
 * every CTA walks output tiles with the halo ring, the weight ring and TWO tensor-memory accumulator stages running
 * across tile boundaries, so the epilogue of one tile overlaps the fill and the MMAs of the next.  use_tma = 1 stages
 * the activation halo with cp.async.bulk.tensor from a 4-D tiled tensor map over the NHWC input (zero padding = the
 * map's out-of-bounds fill, parity planes of stride-2 convs = traversal stride 2); it needs zero padding (or a GEMM
 * without border taps) and c_visible = the channel count of the GEMM's input slice (channels past it read as zero).
 * use_tma = 2 does the same for a reflection-padded conv: the boxes land with zeros outside the image and two warps then
 * mirror exactly those halo rows (nn.ReflectionPad2d) before the chunk is released to the MMAs.  use_tma = 0 keeps the
 * cp.async producers.  Same tables and packed weights as
 * catb_igemm_halo_fprop; replaces the same ATen conv / conv-transpose / conv-backward-input call sites
 * (models/modules/inception_modules.py:22-44, models/modules/discriminators.py:37-75 of the reference). */
int catb_igemm_halo_persist_fits(int n_planes, int Lh, int Wf, int mul, int n_tile, int m_sub, int n_steps, int n_chunks,
                                 int b_budget, int use_tma);
int catb_igemm_halo_fprop_persist(const catb_igemm_desc* d, const catb_halo_desc* h, const catb_halo_step* steps /*device*/,
                                  const catb_halo_chunk* chunks /*device*/, const void* x, const void* packed_w,
                                  const float* bias /*nullable*/, void* y, int use_tma, int c_visible,
                                  const catb_epilogue_stats* stats /*nullable*/, catb_stream_t s);

/* v2 weight gradient on the same halo plan (m_sub = 1): a CTA handles one 128-channel tile of the lattice
 * tensor, one channel chunk of X and one group of <= 8 consecutive steps (taps) of that chunk, each tap
 * accumulating in its own 64 TMEM columns (cat_b200/csrc/igemm_halo_wgrad.cu).  `wunits` holds 8 weight
 * units per step (the chunk-aligned table of make_halo_plan). */
typedef struct {
  int32_t chunk, first_step, n_steps, reserved;
} catb_halo_wgroup;
int catb_igemm_halo_wgrad_fits(int n_planes, int Lh);
int catb_igemm_halo_wgrad(const catb_igemm_desc* d, const catb_halo_desc* h, const catb_halo_step* steps /*device*/,
                          const catb_halo_chunk* chunks /*device*/, const catb_halo_wgroup* groups /*device*/,
                          int n_groups, const catb_weight_unit* wunits /*device*/, const void* x, const void* y,
                          float* arena_grad, catb_stream_t s);

/* Two-stage (deterministic) weight gradient, replacing the atomic accumulation of the two entry points above on the
 * product path: stage 1 writes every row split's partial tile to a caller-owned fp32 workspace
 * [splits][n_rows][ws_k] in GEMM order (column = unit * 8 + element; the halo variant uses its chunk-aligned unit
 * table, 8 units per step) with plain coalesced stores; stage 2 (catb_wgrad_unpack) sums the splits in split order and
 * adds rows [row0, row0 + n_rows) of it to the gradient arena through a weight-unit table (one call per row segment of
 * an N-concatenated GEMM: the first-stage convs of a residual block share one weight-gradient GEMM, with the 1x1 and
 * 3x3 kernels embedded in the 5x5 tap grid).  *_ws_shape return the split count and
 * the column pitch for a descriptor (workspace elements = splits * n_rows * ws_k). */
int catb_igemm_wgrad_ws_shape(const catb_igemm_desc* d, int* splits, int* ws_k);
int catb_igemm_wgrad_ws(const catb_igemm_desc* d, const catb_gather_unit* units, const void* x, const void* y, float* ws,
                        catb_stream_t s);
/* use_tma = 1 (zero padding or no border taps, one strip, dense output lattice; catb_igemm_halo_wgrad_tma_fits): both
 * operands of a tile -- dY[positions x 128 channels] and the X halo planes -- arrive as cp.async.bulk.tensor boxes of whole
 * frame rows issued by one thread (positions past the image read as zero = the garbage rows of pitch space), in a 2-4
 * deep ring of 128- or 64-position tiles; c_visible as for catb_igemm_halo_fprop_persist.  The split count depends on
 * the tile size, so the shape query takes the same flag. */
int catb_igemm_halo_wgrad_tma_fits(const catb_halo_desc* h);
int catb_igemm_halo_wgrad_ws_shape(const catb_igemm_desc* d, const catb_halo_desc* h, int n_groups, int use_tma, int* splits,
                                   int* ws_k);
int catb_igemm_halo_wgrad_ws(const catb_igemm_desc* d, const catb_halo_desc* h, const catb_halo_step* steps /*device*/,
                             const catb_halo_chunk* chunks /*device*/, const catb_halo_wgroup* groups /*device*/,
                             int n_groups, const void* x, const void* y, float* ws, int use_tma, int c_visible,
                             catb_stream_t s);
int catb_wgrad_unpack(const float* ws, int n_splits, int ws_rows /* rows per split */, int ws_k, int row0, int n_rows,
                      int n_units, const catb_weight_unit* wunits /*device*/, float* arena_grad, catb_stream_t s);

/* All second stages of one backward pass in ONE launch: a device-resident job table (fields as the arguments of
 * catb_wgrad_unpack, plus the gradient arena each job adds to). */
typedef struct {
  const void* ws;      /* const float* (device) */
  const void* wunits;  /* const catb_weight_unit* (device) */
  void* grad;          /* float* gradient arena (device) */
  int32_t n_splits, ws_rows, ws_k, row0, n_rows, n_units;
} catb_unpack_job;
int catb_wgrad_unpack_batch(const catb_unpack_job* jobs /*device*/, int n_jobs, int blocks_per_job, catb_stream_t s);

/* Development aid: with a device buffer of 4096 x 16 uint64 registered (zeroed by the caller), every catb_igemm_halo_fprop CTA with
 * blockIdx.x < 4096 records %globaltimer (ns) at its phase boundaries: 0 prologue done, 1 first halo chunk filled,
 * 2 MMA warp released, 3 last MMA issued, 4 accumulators complete, 5 epilogue done, 6 exit, 7 kernel entry; 8-10 cycles of thread 0 in the epilogue's TMEM loads / pack + staging / copy-out.  NULL: off. */
int catb_debug_timeline(void* device_buffer);
/* Experiment switches of the halo-fprop epilogue (results become wrong): 1 no global stores, 2 no TMEM loads, 4 direct
 * per-thread stores instead of the staged coalesced ones.  0: production. */
int catb_debug_mode(int mode);

/* All GEMM images of a network in ONE launch: a device-resident job table (one entry per packed image or row
 * segment, same meaning as the arguments of catb_pack_weights_rows; n_chunks = ceil(n_units / 8)). */
typedef struct {
  const void* wunits;   /* const catb_weight_unit* (device) */
  void* packed;         /* packed image (device) */
  int32_t n_tile, n_units, n_chunks;
  int32_t row0, span, nreal;
} catb_pack_job;
int catb_pack_weights_batch(const catb_pack_job* jobs /*device*/, int n_jobs, int blocks_per_job, const float* arena,
                            catb_stream_t s);

/* Slow SIMT restatements of the two kernels above (same descriptors); kept for on-device bisection
 * in tests.  Not used by the product path. */
int catb_ref_fprop(const catb_igemm_desc* d, const catb_gather_unit* units, const catb_weight_unit* wunits,
                   const float* arena, const void* x, const float* bias, void* y, catb_stream_t s);
int catb_ref_wgrad(const catb_igemm_desc* d, const catb_gather_unit* units, const catb_weight_unit* wunits,
                   const void* x, const void* y, float* arena_grad, catb_stream_t s);

/* ---- depthwise convolution (per-channel kernel size, padding (k-1)/2) ---------------------------
 * models/modules/inception_modules.py:165-173 (ConvBNReLU(groups=midp), reflect padded: pad_mode
 * CATB_PAD_REFLECT) and :441-452 / :700-712 (ConvSyncBNReLU(groups=midp) of the SPADE blocks, zero padded:
 * CATB_PAD_ZERO).  ksize[c] in {1,3,5,7}; w_off[c] = element offset of the channel's k*k filter in the arena
 * (-1: padding channel). */
int catb_dwconv_fwd(const void* x, int ldx, int x_coff, void* y, int ldy, int y_coff, int N, int H, int W, int C,
                    const int32_t* ksize, const int32_t* w_off, const float* arena, int pad_mode, catb_stream_t s);
int catb_dwconv_bwd_data(const void* dy, int ldy, int y_coff, void* dx, int ldx, int x_coff, int N, int H, int W,
                         int C, const int32_t* ksize, const int32_t* w_off, const float* arena, int pad_mode,
                         catb_stream_t s);
int catb_dwconv_bwd_weight(const void* x, int ldx, int x_coff, const void* dy, int ldy, int y_coff, int N, int H,
                           int W, int C, const int32_t* ksize, const int32_t* w_off, float* arena_grad, int pad_mode,
                           catb_stream_t s);

/* ---- normalisation (InstanceNorm2d / BatchNorm2d, models/networks.py:29-64) --------------------
 * stats: sums[g*2*C + c] = sum x, sums[g*2*C + C + c] = sum x^2 with g = n (per_sample) or 0.
 * The buffer must be zeroed by the caller (atomic accumulation). */
int catb_norm_stats(const void* x, int ldx, int x_coff, int N, int HW, int C, int per_sample, float* sums,
                    catb_stream_t s);
/* scale/shift [G,C] from the sums (training) or from running stats (eval: sums == NULL).
 * Training with running buffers also applies the momentum update with the unbiased variance
 * (F.batch_norm semantics).  gamma/beta/running_* are arena offsets (-1: absent). */
int catb_norm_finalize(const float* sums, int G, int C, float count, float eps, float momentum,
                       const float* gamma, const float* beta, float* running_mean, float* running_var,
                       float* scale, float* shift, float* mean_rstd /* [G,2,C] saved for backward */,
                       catb_stream_t s);
/* y = act(x*scale + shift) (+ residual).  Replaces norm + ReLU/LeakyReLU (+ the residual add of
 * inception_modules.py:235-236). */
int catb_norm_apply(const void* x, int ldx, int x_coff, void* y, int ldy, int y_coff, const void* residual,
                    int ldr, int r_coff, int N, int HW, int C, int per_sample, const float* scale,
                    const float* shift, int act, catb_stream_t s);
/* catb_norm_finalize + catb_norm_apply in one launch: every thread derives scale / shift of its channels from the sums;
 * the first pixel block of each group also writes scale / shift / mean_rstd (for the backward pass) and moves the
 * running statistics.  With the sums produced by the conv epilogue (catb_epilogue_stats) a conv + norm + activation
 * block of the reference (inception_modules.py:22-44) is two launches. */
int catb_norm_apply_fused(const void* x, int ldx, int x_coff, void* y, int ldy, int y_coff, const void* residual,
                          int ldr, int r_coff, int N, int HW, int C, int per_sample, const float* sums, float count,
                          float eps, float momentum, const float* gamma, const float* beta, float* running_mean,
                          float* running_var, float* scale, float* shift, float* mean_rstd, int act, catb_stream_t s);
/* Backward of act(norm(x)): pass 1 reduces sum(dz) and sum(dz*xhat) into red[G,2,C] (zeroed by the
 * caller), pass 2 writes dx.  `out` is the saved activation output (act' is taken from it);
 * act == NONE ignores it. */
int catb_norm_bwd_reduce(const void* dout, int ldd, int d_coff, const void* out, int ldo, int o_coff,
                         const void* x, int ldx, int x_coff, int N, int HW, int C, int per_sample,
                         const float* mean_rstd, int act, float* red, catb_stream_t s);
int catb_norm_bwd_apply(const void* dout, int ldd, int d_coff, const void* out, int ldo, int o_coff,
                        const void* x, int ldx, int x_coff, void* dx, int ldg, int g_coff, int N, int HW, int C,
                        int per_sample, const float* mean_rstd, const float* gamma, const float* red,
                        float count, int act, float* dgamma, float* dbeta, catb_stream_t s);

/* ---- element-wise helpers --------------------------------------------------------------------- */
/* NCHW fp32 (reference layout) -> NHWC bf16 channel slice, and back. */
int catb_nchw_to_nhwc(const float* src, int N, int C, int H, int W, void* dst, int ldd, int d_coff, catb_stream_t s);
int catb_nhwc_to_nchw(const void* src, int lds, int s_coff, int N, int C, int H, int W, float* dst, catb_stream_t s);
/* dst[..., d_coff:d_coff+C] = src[..., s_coff:s_coff+C] (torch.cat along channels,
 * base_inception_distiller.py:295-296). */
int catb_copy_channels(const void* src, int lds, int s_coff, void* dst, int ldd, int d_coff, long long pixels,
                       int C, catb_stream_t s);
/* dz = dout * act'(out) for activations without a norm (LeakyReLU after the first D conv, Tanh head). */
int catb_act_bwd(const void* dout, int ldd, int d_coff, const void* out, int ldo, int o_coff, void* dz, int ldz,
                 int z_coff, long long pixels, int C, int act, catb_stream_t s);
/* per-channel sum over pixels of a bf16 tensor, atomically added to out[c] (conv bias gradients). */
int catb_channel_sum(const void* x, int ldx, int x_coff, long long pixels, int C, float* out, catb_stream_t s);
/* Adjoint of nn.ReflectionPad2d(p): dx[n,h,w,:] = sum of the padded-frame gradient entries that
 * mirror onto (h,w) (+ add, optional).  src has spatial size (H+2p, W+2p). */
int catb_reflect_fold(const void* src, int lds, int s_coff, void* dst, int ldd, int d_coff, const void* add,
                      int lda, int a_coff, int N, int H, int W, int C, int p, catb_stream_t s);
int catb_add(const void* a, int lda, int a_coff, const void* b, int ldb, int b_coff, void* dst, int ldd, int d_coff,
             long long pixels, int C, catb_stream_t s);

/* ---- losses ----------------------------------------------------------------------------------- */
/* GANLoss (models/modules/loss.py:52-99) on a fp32 prediction of n elements with element stride `ld`:
 * *loss += value (atomic; caller zeroes); dpred (nullable) is written as bf16 rows of 8 channels with
 * pitch ldg, channel g_coff carrying grad_scale * dloss/dpred and the other 7 channels zero. */
int catb_gan_loss(const float* pred, long long n, int ld, int mode, int target_is_real, int for_discriminator,
                  float grad_scale, float* loss, void* dpred, int ldg, int g_coff, catb_stream_t s);
/* Reconstruction loss (base_inception_distiller.py:171-176): kind 0 = L1Loss, 1 = MSELoss,
 * 2 = SmoothL1Loss (beta 1).  *loss += mean over the Creal real channels; da (bf16, nullable) =
 * grad_scale * dloss/da + extra (optional bf16 tensor added, e.g. the GAN gradient). */
int catb_recon_loss(const void* a, int lda, int a_coff, const void* b, int ldb, int b_coff, long long pixels, int C,
                    int Creal, int kind, float grad_scale, float* loss, void* da, int ldg, int g_coff,
                    const void* extra, int lde, int e_coff, catb_stream_t s);
/* KA (utils/common.py:38-46).  gram: G[B,B] += X X^T over K = pixels*C elements per sample (caller
 * zeroes G).  ka_finish: value and the B x B coefficient matrix of dX = coef * X.  ka_bwd: dX (+)= coef X. */
int catb_gram(const void* x, int ldx, int x_coff, int B, long long pixels_per_sample, int C, float* G, catb_stream_t s);
/* SIMT restatement of catb_gram (the round-1 kernel), kept for on-device bisection in tests. */
int catb_gram_ref(const void* x, int ldx, int x_coff, int B, long long pixels_per_sample, int C, float* G, catb_stream_t s);
int catb_ka_finish(const float* Gx, const float* Gy, int B, float loss_scale, float* loss /* += */,
                   float* ka_value, float* coef, catb_stream_t s);
int catb_ka_bwd(const void* x, int ldx, int x_coff, int B, long long pixels_per_sample, int C, const float* coef,
                void* dx, int ldg, int g_coff, int accumulate, catb_stream_t s);

/* ---- optimiser -------------------------------------------------------------------------------- */
/* torch.optim.Adam (base_inception_distiller.py:205-214) over a flat fp32 arena.  `step_count` is a
 * device counter incremented by the kernel; `lr` is read from device memory so that a captured
 * CUDA graph follows the scheduler (models/networks.py:80-87). */
int catb_adam(float* param, const float* grad, float* m, float* v, long long n, const float* lr, float beta1,
              float beta2, float eps, float grad_scale, int* step_count, catb_stream_t s);

/* ---- SPADE distillation path ----------------------------------------- */
/* F.interpolate(mode='nearest') / nn.Upsample(scale_factor=2) (inception_spade_generator.py:68,79-113;
 * inception_modules.py:750) on an NHWC bf16 slice: y[n,oh,ow] = x[n, floor(oh*H/OH), floor(ow*W/OW)]. */
int catb_resize_nearest(const void* x, int ldx, int x_coff, int H, int W, void* y, int ldy, int y_coff, int N,
                        int OH, int OW, int C, catb_stream_t s);
/* adjoint of the 2x nearest up-sampling: dx[n,h,w] = sum of the 2x2 block of dy ([N,2H,2W]). */
int catb_upsample2x_bwd(const void* dy, int ldy, int y_coff, void* dx, int ldx, int x_coff, int N, int H, int W,
                        int C, catb_stream_t s);
/* InceptionSPADE.forward (inception_modules.py:746-762) fused with the activation that follows it (:555-556):
 * y = act((x*scale[c] + shift[c]) * (1 + gamma) + beta); scale/shift = the parameter-free BatchNorm as
 * produced by catb_norm_finalize without gamma/beta. */
int catb_spade_modulate(const void* x, int ldx, int x_coff, const void* gamma, int ldg, int g_coff,
                        const void* beta, int ldb, int b_coff, void* y, int ldy, int y_coff, long long pixels,
                        int C, const float* scale, const float* shift, int act, catb_stream_t s);
/* its backward: dz = dy*act'(y); dgamma = dz*xhat; dbeta = dz; dn = dz*(1+gamma) (dn then goes through
 * catb_norm_bwd_reduce / catb_norm_bwd_apply of the parameter-free norm). */
int catb_spade_modulate_bwd(const void* dy, int ldd, int d_coff, const void* y, int ldy, int y_coff, const void* x,
                            int ldx, int x_coff, const void* gamma, int ldg, int g_coff, void* dgamma, int ldo,
                            int o_coff, void* dbeta, int ldp, int p_coff, void* dn, int ldn, int n_coff,
                            long long pixels, int C, const float* scale, const float* shift, int act,
                            catb_stream_t s);
/* y = act(x) (F.leaky_relu before conv_img, inception_spade_generator.py:115). */
int catb_act_fwd(const void* x, int ldx, int x_coff, void* y, int ldy, int y_coff, long long pixels, int C, int act,
                 catb_stream_t s);
/* MultiscaleDiscriminator.downsample (discriminators.py:205-210): avg_pool2d(3, stride 2, padding 1,
 * count_include_pad=False), OH = (H+1)/2; backward writes dx = add (nullable) + adjoint(dy). */
int catb_avgpool3s2(const void* x, int ldx, int x_coff, int H, int W, void* y, int ldy, int y_coff, int N, int C,
                    catb_stream_t s);
int catb_avgpool3s2_bwd(const void* dy, int ldy, int y_coff, const void* add, int lda, int a_coff, void* dx, int ldx,
                        int x_coff, int N, int H, int W, int C, catb_stream_t s);
/* VGG19 max_pool2d(2,2) (models/modules/loss.py:151-184) and its backward (gradient to the first maximum of
 * each window in row-major order, like ATen). */
int catb_maxpool2(const void* x, int ldx, int x_coff, int H, int W, void* y, int ldy, int y_coff, int N, int C,
                  catb_stream_t s);
int catb_maxpool2_bwd(const void* dy, int ldy, int y_coff, const void* x, int ldx, int x_coff, void* dx, int ldg,
                      int g_coff, int N, int H, int W, int C, catb_stream_t s);
/* SPADEModel.preprocess_input / get_edges (models/spade_model.py:142-179): y[..., label] = 1 for
 * label < n_label, y[..., n_label] = 4-neighbour instance boundary (instance nullable: no edge channel),
 * every other channel of the slice 0.  label / instance: int32 [N,H,W]. */
int catb_onehot_edges(const int32_t* label, const int32_t* instance, int N, int H, int W, int n_label, void* y,
                      int ldy, int y_coff, int C, catb_stream_t s);
/* out[i] = sum_k arena[idx[k*n+i]] (negative index: skipped) -- the summed biases of K-concatenated convs in
 * the padded channel order of their output slice; scatter_add is its adjoint (bias gradients);
 * fma_vec: shift[i] += bias[i]*scale[i] (a conv bias in front of an eval-mode BatchNorm). */
int catb_gather_sum_f32(const float* arena, const int32_t* idx, int K, int n, float* out, catb_stream_t s);
int catb_scatter_add_f32(const float* src, const int32_t* idx, int K, int n, float* arena, catb_stream_t s);
int catb_fma_vec(float* shift, const float* bias, const float* scale, int n, catb_stream_t s);
/* torch.nn.utils.spectral_norm as applied by get_nonspade_norm_layer (spade_architecture/normalization.py:
 * 17-50): per weight [rows, cols] one power iteration on (u, v) when training (in place in `bufs`),
 * sigma = u^T W v, w_eff = W / sigma.  Backward turns the gradient w.r.t. w_eff (in `grad`, in place) into
 * the gradient w.r.t. weight_orig: (g - <g, w_eff> u v^T) / sigma.  tmp: n*max(max_rows,max_cols) floats. */
typedef struct {
  int32_t w_off;        /* element offset of weight_orig in the parameter arena (same offset in w_eff / grad) */
  int32_t rows, cols;   /* Cout, Cin*kh*kw */
  int32_t u_off, v_off; /* element offsets of weight_u / weight_v in the buffer arena */
  int32_t reserved;
} catb_sn_desc;
int catb_sn_forward(const catb_sn_desc* table /*device*/, int n, int max_rows, int max_cols, const float* arena,
                    float* bufs, int training, float* tmp, float* sigma /*[n]*/, float* w_eff, catb_stream_t s);
int catb_sn_backward(const catb_sn_desc* table /*device*/, int n, int max_rows, int max_cols, float* grad,
                     const float* w_eff, const float* bufs, const float* sigma, float* cdot /*[n]*/, catb_stream_t s);

/* ---- x-packed 7x7 stem / head convolutions (inception_generator.py:37-56,130-134) -------------------
 * Packing the horizontal taps into the channel dimension turns a 7x7 conv with 3 input (stem) or 3 output (head)
 * channels into a 7x1 implicit GEMM with 24 channels on that side (7x fewer GEMM steps):
 *   expand_x:     y[n,h,w, dx*Cin + ci] = x[n,h, reflect(w + dx - taps/2), ci]   (channels >= taps*Cin: 0)
 *   shift_sum:    out[n,h,w,co] = act(bias[co] + sum_dx P[n,h,w+dx, co*8+dx]),  P is [N,H,W+taps-1,.]
 *   shift_expand: dP[n,h,c', co*8+dx] = dz[n,h,c'-dx,co] (0 outside [0,W)): the adjoint of shift_sum. */
int catb_expand_x(const void* x, int ldx, int x_coff, void* y, int ldy, int y_coff, int N, int H, int W, int Cin,
                  int taps, int Cy, catb_stream_t s);
int catb_shift_sum(const void* P, int ldp, int p_coff, void* out, int ldo, int o_coff, int N, int H, int W, int Cout,
                   int taps, const float* bias /*nullable*/, int act, catb_stream_t s);
int catb_shift_expand(const void* dz, int ldz, int z_coff, void* dP, int ldp, int p_coff, int N, int H, int W, int Cout,
                      int taps, catb_stream_t s);

/* "Tap-split" form of a conv with one output channel (the PatchGAN head, models/modules/discriminators.py:72-73, and
 * the last conv of the SPADE sub-discriminators :160-161): P[n,iy,ix,t] = sum_c X[n,iy,ix,c] W[0,c,t] is a 1x1 GEMM with
 * the R*S taps as output channels; catb_tap_sum adds bias + the shifted taps into channel o_coff of the fp32 prediction,
 * catb_tap_expand spreads dY (channel y_coff of a bf16 buffer) back over the taps for the 1x1 weight / input gradients. */
int catb_tap_sum(const float* P, int ldp, int p_coff, float* out, int ldo, int o_coff, int N, int H, int W, int OH, int OW,
                 int R, int S, int pad, const float* bias /*nullable, 1 element*/, catb_stream_t s);
int catb_tap_expand(const void* dy, int ldy, int y_coff, void* dP, int ldp, int p_coff, int N, int H, int W, int OH, int OW,
                    int R, int S, int pad, catb_stream_t s);

#ifdef __cplusplus
}
#endif
#endif /* CATB200_H_ */
