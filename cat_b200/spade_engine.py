"""Execution engine for the CAT *SPADE* distillation step on libcatb200 kernels.

Same execution model as cat_b200/engine.py (static launch sequences with hand-derived backward passes over
pre-allocated NHWC bf16 buffers, parameters in flat fp32 arenas with the reference's state_dict keys).

Networks restated here (same maths as the reference modules, different execution):
  SpadeGenNet   -- InceptionSPADEGenerator (models/modules/inception_architecture/inception_spade_generator.py:15-124)
                   with SPADEInvertedResidualChannels / InceptionSPADE blocks (models/modules/inception_modules.py:
                   345-762) and SynchronizedBatchNorm2d in its single-replica form (sync_batchnorm/batchnorm.py:68-72)
  MultiScaleDis -- MultiscaleDiscriminator over SPADENLayerDiscriminator (models/modules/discriminators.py:129-226) with
                   torch.nn.utils.spectral_norm + InstanceNorm2d(affine=False) (spade_architecture/normalization.py:17-50)
  VggNet        -- VGG19 slices 1-5 of VGGLoss (models/modules/loss.py:151-203), forward + input gradient
"""
import os

import numpy as np
import torch

from . import _C
from . import igemm_plan as P
from . import ops
from .engine import Arena, DisNet, Norm, pool_norm_buffers
from .igemm_plan import cpad
from .ops import ACT, Act, Gemm

MAPPING_LAYERS = ['head_0', 'G_middle_1', 'up_1']      # base_spade_distiller_modules.py:70
PAD_ZERO = _C.PAD_ZERO


# ------------------------------------------------------------------------------------------------
# pooled bias vectors: out[i] = sum_k arena[idx[k, i]]
# ------------------------------------------------------------------------------------------------
class BiasPool:
    """All fused bias vectors of a network (summed biases of K-concatenated convs, biases folded into eval-mode
    BatchNorm shifts) are gathered from the parameter arena by ONE launch per forward pass, and all bias
    gradients are scattered back by one launch per backward pass."""

    def __init__(self):
        self.cols = []        # per pooled element: list of arena offsets
        self.vec = self.dvec = self.idx = None

    def add(self, per_channel_offsets):
        """per_channel_offsets: list (length n, padded channel order) of lists of arena offsets.  Returns the
        (start, n) slot in the pooled vector."""
        start = len(self.cols)
        self.cols.extend([list(o) for o in per_channel_offsets])
        while len(self.cols) % 8:
            self.cols.append([])
        return start, len(per_channel_offsets)

    def finalize(self, dev, with_grad):
        n = max(len(self.cols), 8)
        K = max([len(c) for c in self.cols] + [1])
        idx = -np.ones((K, n), dtype=np.int32)
        for i, c in enumerate(self.cols):
            idx[:len(c), i] = c
        self.idx = torch.from_numpy(idx).to(dev)
        self.vec = torch.zeros(n, dtype=torch.float32, device=dev)
        self.dvec = torch.zeros(n, dtype=torch.float32, device=dev) if with_grad else None
        self.used = len(self.cols) > 0

    def slot(self, s, which='vec'):
        return getattr(self, which)[s[0]:s[0] + s[1]]

    def gather(self, arena_p):
        if self.used:
            ops.gather_sum(arena_p, self.idx, self.vec)

    def scatter(self, arena_g):
        if self.used:
            ops.scatter_add(self.dvec, self.idx, arena_g)


def _vec_offsets(arena, name, C):
    """Arena offsets of a [C] parameter vector in padded channel order (padding channels: none)."""
    off = arena.off(name)
    return [[off + c] if c < C else [] for c in range(cpad(C))]


class BNorm(Norm):
    """BatchNorm with running statistics whose producing conv carries a bias.  In training the bias is inert
    for the output (batch statistics absorb it) and only shifts running_mean; in eval it is folded into the
    shift.  `bias` is a slice of the network's BiasPool."""

    def __init__(self, *a, bias=None, **k):
        super().__init__(*a, **k)
        self.bias = bias
        if bias is not None:
            self.mom_vec = torch.full_like(bias, self.momentum)

    def _stats(self, x):
        if not self.pooled:
            self.sums.zero_()
        ops.norm_stats(x, self.per_sample, self.sums)
        upd = self.training and self.track
        ops.norm_finalize(self.sums, self.G, self.Cp, self.count, self.eps, self.momentum, self.gamma, self.beta,
                          self.rmean if upd else None, self.rvar if upd else None, self.scale, self.shift, self.mean_rstd)
        if upd and self.bias is not None:
            ops.fma_vec(self.rmean, self.bias, self.mom_vec)     # running_mean of (conv + bias)

    def prepare(self, x: Act):
        """scale / shift / saved statistics without applying them (SPADE modulation applies them itself)."""
        if self.batch_stats:
            self._stats(x)
        else:
            ops.norm_finalize(None, self.G, self.Cp, self.count, self.eps, self.momentum, self.gamma, self.beta,
                              self.rmean, self.rvar, self.scale, self.shift, self.mean_rstd)
            if self.bias is not None:
                ops.fma_vec(self.shift.view(-1), self.bias, self.scale.view(-1))

    def forward(self, x: Act, y: Act, act, residual=None):
        self.prepare(x)
        ops.norm_apply(x, y, self.scale, self.shift, self.per_sample, act, residual)

    def backward(self, dout: Act, out, x: Act, dx: Act, act, param_grads=True):
        super().backward(dout, out, x, dx, act, param_grads)
        if not self.batch_stats and param_grads and getattr(self, 'dbias', None) is not None:
            # eval mode (the reference's first step of a run): the conv bias in front of the norm is live,
            # d bias = gamma * rstd * sum(dz) = scale * red[0]
            ops.fma_vec(self.dbias, self.red.view(-1)[:self.Cp], self.scale.view(-1)[:self.Cp])


class _Net:
    """Arena / norm / bias bookkeeping shared by the networks below."""

    def _init_common(self, B, device, training, need_grad, share=None):
        self.B, self.dev, self.training, self.need_grad = B, device, training, need_grad
        if share is not None:      # another compilation (shape / mode) of the same weights: reuse its finalised arenas
            assert share.arena.with_grad or not need_grad
            self.arena, self.bufs = share.arena, share.bufs
        else:
            self.arena, self.bufs = Arena(with_grad=need_grad), Arena(with_grad=False)
        self.biases = BiasPool()
        self.norms = []
        self.fprop_gemms, self.bwd_gemms = [], []

    # ---- weight gradients on a side stream (parallel branches of the captured graph) ------------------------------
    def _side_begin(self):
        """Called at the start of a backward pass.  Weight-gradient launches only feed the optimiser, so they run on a
        side stream while the main stream carries the input-gradient chain; the side stream reads one workspace shared
        by all six-branch bodies (`dmid_raw`), which the main stream re-writes only after waiting for `_ev_raw`."""
        self._ov = getattr(self, 'overlap_wgrad', False) and str(self.dev) != 'cpu'
        self._ev_raw = None
        if self._ov:
            self._main = torch.cuda.current_stream()
            if getattr(self, '_wside', None) is None:
                self._wside = torch.cuda.Stream(device=self.dev)

    def _on_side(self, fn):
        if not self._ov:
            return fn()
        self._wside.wait_stream(self._main)
        with torch.cuda.stream(self._wside):
            fn()

    def _mark_raw(self):
        if self._ov:
            self._ev_raw = torch.cuda.Event()
            self._ev_raw.record(self._wside)

    def _wait_raw(self):
        if self._ov and self._ev_raw is not None:
            self._main.wait_event(self._ev_raw)

    def _side_end(self):
        if self._ov:
            self._main.wait_stream(self._wside)

    def G(self, geo, units, n_rows, bwd=False, **kw):
        g = Gemm(geo, units, n_rows, self.dev, **kw)
        if kw.get('need_pack', True):
            (self.bwd_gemms if bwd else self.fprop_gemms).append(g)
        return g

    def alloc_bn(self, groups, affine):
        """groups: [(prefix, C)] allocated contiguously in padded channel order (like engine._NormSpec)."""
        a, b = self.arena, self.bufs
        if affine:
            for p, C in groups:
                a.alloc(p + '.weight', (C,), cpad(C), 1.0)
            for p, C in groups:
                a.alloc(p + '.bias', (C,), cpad(C), 0.0)
        for p, C in groups:
            b.alloc(p + '.running_mean', (C,), cpad(C), 0.0)
        for p, C in groups:
            b.alloc(p + '.running_var', (C,), cpad(C), 1.0)
        for p, C in groups:
            b.alloc(p + '.num_batches_tracked', (1,))

    def make_bn(self, HW, groups, affine, bias_slot=None):
        a, b = self.arena, self.bufs
        first, last = groups[0][0], groups[-1][0]
        Cp = sum(cpad(C) for _, C in groups)
        kw = {}
        if affine:
            kw['gamma'] = a.span(first + '.weight', last + '.weight')
            kw['beta'] = a.span(first + '.bias', last + '.bias')
            if a.with_grad:
                kw['dgamma'] = a.span(first + '.weight', last + '.weight', 'g')
                kw['dbeta'] = a.span(first + '.bias', last + '.bias', 'g')
        kw['rmean'] = b.span(first + '.running_mean', last + '.running_mean')
        kw['rvar'] = b.span(first + '.running_var', last + '.running_var')
        n = BNorm(self.dev, self.B, HW, Cp, 'batch', 1e-5, getattr(self, 'bn_momentum', 0.1), self.training, True, **kw)
        n._bias_slot = bias_slot
        self.norms.append(n)
        return n

    def set_training(self, training):
        """train() / eval() of the network: only the BatchNorm layers depend on it.  The launch sequence changes, so captured
        CUDA graphs of this network must be re-captured."""
        self.training = training
        for n in self.norms:
            n.set_training(training)

    def _finish_build(self):
        self.biases.finalize(self.dev, self.need_grad)
        for n in self.norms:
            if getattr(n, '_bias_slot', None) is not None:
                n.bias = self.biases.slot(n._bias_slot)
                n.mom_vec = torch.full_like(n.bias, n.momentum)
                n.dbias = self.biases.slot(n._bias_slot, 'dvec') if self.need_grad else None
        self.pool_sums, self.pool_red = pool_norm_buffers(self.norms, self.dev)

    def pack_weights(self):
        if getattr(self, '_packer', None) is None:
            self._packer = ops.PackBatch(self.fprop_gemms + self.bwd_gemms, self.dev)
        self._packer.run(self.arena.p)

    def load_state_dict(self, sd):
        self.arena.load_state_dict(sd)
        self.bufs.load_state_dict(sd)
        for n in self.norms:
            n._frozen = False
        self.pack_weights()

    def state_dict(self):
        sd = self.arena.state_dict()
        sd.update(self.bufs.state_dict())
        return sd


# ------------------------------------------------------------------------------------------------
# the six-branch body shared by SPADEInvertedResidualChannels and InceptionSPADE
# ------------------------------------------------------------------------------------------------
class SixBranch:
    """res branch j: Conv k (bias) -> BN(affine) -> ReLU -> Conv k (bias);  dw branch j: Conv 1x1 (bias) -> BN(affine)
    -> ReLU -> depthwise Conv k (bias) -> BN(dw_affine) -> ReLU -> Conv 1x1 (bias); zero padding (k-1)/2; the branch
    outputs are summed (inception_modules.py:412-470, 672-722).  Execution as in engine.GenNet blocks: the 1x1
    first convs are one N-concatenated GEMM, all first-stage norms one launch, the last convs of every branch ONE
    K-concatenated GEMM (= the branch sum) whose bias is the sum of the branch biases.

    out_segs: [(Act slice, n_real_rows, first weight row)] -- the main body writes one segment (all fout rows), the
    SPADE body two (gamma rows [0,C), beta rows [C,2C)) into 8-aligned halves of one buffer."""

    def __init__(self, net, prefix, res_w, dw_w, ks, Cin, last_key, dw_affine):
        self.net, self.prefix, self.Cin, self.last_key, self.dw_affine = net, prefix, Cin, last_key, dw_affine
        # the modulation body (InceptionSPADE, dw_affine) is always nn.ReLU (inception_modules.py:600); the main body of
        # a block follows the generator's active_fn (:357, 442-474)
        self.act = ACT['relu'] if dw_affine else net.act
        self.res = [(j, m, k) for j, (m, k) in enumerate((mk for mk in zip(res_w, ks) if mk[0] > 0))]
        self.dw = [(j, m, k) for j, (m, k) in enumerate((mk for mk in zip(dw_w, ks) if mk[0] > 0))]
        self.empty = not self.res and not self.dw
        # slice order in the mid buffer: [1x1 first convs (res k=1, dw) | res k>1 | dw second stage]
        self.order = ([('res', j, m, k) for (j, m, k) in self.res if k == 1] + [('dw', j, m, 1) for (j, m, k) in self.dw] +
                      [('res', j, m, k) for (j, m, k) in self.res if k > 1])

    def first_conv(self, kind, j):
        return f'{self.prefix}.res_ops.{j}.0.conv' if kind == 'res' else f'{self.prefix}.dw_ops.{j}.0.conv'

    def first_norm(self, kind, j):
        return f'{self.prefix}.res_ops.{j}.0.norm' if kind == 'res' else f'{self.prefix}.dw_ops.{j}.0.norm'

    def last_conv(self, kind, j):
        return (f'{self.prefix}.res_ops.{j}.1' if kind == 'res' else f'{self.prefix}.dw_ops.{j}.2') + self.last_key

    def alloc(self, Cout_total):
        if self.empty:
            return
        ar, pre, Cin = self.net.arena, self.prefix, self.Cin
        for j, m, k in self.res:
            ar.alloc(self.first_conv('res', j) + '.weight', (m, Cin, k, k))
            ar.alloc(self.first_conv('res', j) + '.bias', (m,))
            ar.alloc(self.last_conv('res', j) + '.weight', (Cout_total, m, k, k))
            ar.alloc(self.last_conv('res', j) + '.bias', (Cout_total,))
        for j, m, k in self.dw:
            ar.alloc(self.first_conv('dw', j) + '.weight', (m, Cin, 1, 1))
            ar.alloc(self.first_conv('dw', j) + '.bias', (m,))
            ar.alloc(f'{pre}.dw_ops.{j}.1.conv.weight', (m, 1, k, k))
            ar.alloc(f'{pre}.dw_ops.{j}.1.conv.bias', (m,))
            ar.alloc(self.last_conv('dw', j) + '.weight', (Cout_total, m, 1, 1))
            ar.alloc(self.last_conv('dw', j) + '.bias', (Cout_total,))
        self.grpA = [(self.first_norm(kind, j), m) for (kind, j, m, _k) in self.order]
        self.grpB = [(f'{pre}.dw_ops.{j}.1.norm', m) for j, m, k in self.dw]
        self.net.alloc_bn(self.grpA, True)
        if self.grpB:
            self.net.alloc_bn(self.grpB, self.dw_affine)

    def build(self, x: Act, out_segs, need_input_grad):
        """x: input Act (pitch / offset taken from it); out_segs as in the class docstring."""
        if self.empty:
            return
        net, ar, dev, pre, Cin = self.net, self.net.arena, self.net.dev, self.prefix, self.Cin
        B, H, W = x.N, x.H, x.W
        ng = net.need_grad
        self.x, self.out_segs = x, out_segs
        off = 0
        self.res_sl, self.dw1_sl, self.dw2_sl = [None] * len(self.res), [None] * len(self.dw), []
        self.D0 = self.D1 = 0
        for (kind, j, m, _k) in self.order:
            if kind == 'res':
                self.res_sl[j] = off
            else:
                if j == 0:
                    self.D0 = off
                self.dw1_sl[j] = off
                self.D1 = off + cpad(m)
            off += cpad(m)
        self.LA = off
        for _, m, _k in self.dw:
            self.dw2_sl.append(off)
            off += cpad(m)
        self.L = L = off
        self.mid_raw = Act.empty(B, H, W, L, dev, zero=True)
        self.mid_act = Act.empty(B, H, W, L, dev, zero=True)
        geo_in = dict(N=B, H=H, W=W, ldx=x.ld, x_coff=x.coff)
        # ---- stage 1
        self.s1, self.s1_fwd = [], []
        ones = [(kind, j, m) for (kind, j, m, k) in self.order if k == 1]
        for (kind, j, m, k) in self.order:
            wn = self.first_conv(kind, j) + '.weight'
            sl = self.res_sl[j] if kind == 'res' else self.dw1_sl[j]
            fused = k == 1 and len(ones) > 1
            g = net.G(P.Geometry(**geo_in, OH=H, OW=W, ldy=L, y_coff=sl), P.conv_fprop_units(ar.off(wn), m, Cin, k, k, (k - 1) // 2),
                      m, need_pack=not fused)
            self.s1.append((g, sl, m, k, wn))
            if not fused:
                self.s1_fwd.append(g)
        if len(ones) > 1:
            rows = sum(cpad(m) for (_, _, m) in ones)
            base = P.conv_fprop_units(0, rows, Cin, 1, 1, 0)
            segs = [(sl, cpad(m), m, P.conv_fprop_units(ar.off(wn), m, Cin, 1, 1, 0)) for (_, sl, m, k, wn) in self.s1 if k == 1]
            self.s1_fwd.insert(0, net.G(P.Geometry(**geo_in, OH=H, OW=W, ldy=L, y_coff=0), base, rows, segments=segs))
        biasA = []
        for (kind, j, m, _k) in self.order:
            biasA += _vec_offsets(ar, self.first_conv(kind, j) + '.bias', m)
        self.nA = net.make_bn(H * W, self.grpA, True, bias_slot=net.biases.add(biasA))
        # ---- depthwise convs
        if self.dw:
            Cdw = L - self.LA
            ksz = torch.ones(Cdw, dtype=torch.int32)
            wof = torch.full((Cdw,), -1, dtype=torch.int32)
            biasB = []
            for (j, m, k), sl in zip(self.dw, self.dw2_sl):
                o = sl - self.LA
                ksz[o:o + cpad(m)] = k
                wof[o:o + m] = ar.off(f'{pre}.dw_ops.{j}.1.conv.weight') + torch.arange(m, dtype=torch.int32) * k * k
                biasB += _vec_offsets(ar, f'{pre}.dw_ops.{j}.1.conv.bias', m)
            self.dw_k, self.dw_w = ksz.to(dev), wof.to(dev)
            self.nB = net.make_bn(H * W, self.grpB, self.dw_affine, bias_slot=net.biases.add(biasB))
        # ---- stage 2: one K-concatenated GEMM per output buffer (N-concatenated over the segments)
        def s2_units(row0, nrows):
            u = P.Units()
            for (j, m, k), sl in zip(self.res, self.res_sl):
                w = ar.off(self.last_conv('res', j) + '.weight') + row0 * m * k * k
                u.extend(P.conv_fprop_units(w, nrows, m, k, k, (k - 1) // 2, cu0=sl // 8))
            for (j, m, k), sl in zip(self.dw, self.dw2_sl):
                w = ar.off(self.last_conv('dw', j) + '.weight') + row0 * m
                u.extend(P.conv_fprop_units(w, nrows, m, 1, 1, 0, cu0=sl // 8))
            return u

        def s2_bias(row0, nrows):
            cols = [[] for _ in range(cpad(nrows))]
            for kind, lst in (('res', self.res), ('dw', self.dw)):
                for (j, m, k) in lst:
                    o = ar.off(self.last_conv(kind, j) + '.bias') + row0
                    for c in range(nrows):
                        cols[c].append(o + c)
            return cols

        geo_mid = dict(N=B, H=H, W=W, ldx=L, x_coff=0)
        buf = out_segs[0][0]
        for seg, _, _ in out_segs:
            assert seg.t is buf.t
        span0 = out_segs[0][0].coff
        self.s2_w = []      # per segment: (Gemm used for the weight gradient, Act slice)
        bias_cols = []
        for (seg, nreal, row0) in out_segs:
            assert seg.coff - span0 == len(bias_cols), 'segments must tile the output buffer in 8-aligned slices'
            bias_cols += s2_bias(row0, nreal)
        self.s2_bias_slot = net.biases.add(bias_cols)
        if len(out_segs) == 1:
            seg, nreal, row0 = out_segs[0]
            g = net.G(P.Geometry(**geo_mid, OH=H, OW=W, ldy=seg.ld, y_coff=seg.coff), s2_units(row0, nreal), nreal)
            self.g2 = g
            self.s2_w.append((g, seg))
        else:
            rows = len(bias_cols)
            segs = [(seg.coff - span0, cpad(nreal), nreal, s2_units(row0, nreal)) for (seg, nreal, row0) in out_segs]
            self.g2 = net.G(P.Geometry(**geo_mid, OH=H, OW=W, ldy=buf.ld, y_coff=span0), s2_units(0, rows), rows, segments=segs)
            for (seg, nreal, row0) in out_segs:
                gw = net.G(P.Geometry(**geo_mid, OH=H, OW=W, ldy=seg.ld, y_coff=seg.coff), s2_units(row0, nreal), nreal,
                           need_pack=False)
                self.s2_w.append((gw, seg))
        self.out_span = Act(buf.t, span0, len(bias_cols))
        if not ng:
            return
        # ---- backward GEMMs
        self.d2 = []        # stage-2 input gradients, one GEMM per branch (zero padding: direct)
        for kind, lst, sls in (('res', self.res, self.res_sl), ('dw', self.dw, self.dw2_sl)):
            for (j, m, k), sl in zip(lst, sls):
                kk = k if kind == 'res' else 1
                un = P.Units()
                for (seg, nreal, row0) in out_segs:
                    w = ar.off(self.last_conv(kind, j) + '.weight') + row0 * m * kk * kk
                    un.extend(P.conv_dgrad_units(w, nreal, m, kk, kk, (kk - 1) // 2, cu0=(seg.coff - span0) // 8))
                g = net.G(P.Geometry(N=B, H=H, W=W, ldx=buf.ld, x_coff=span0, OH=H, OW=W, ldy=L, y_coff=sl), un, m, bwd=True)
                self.d2.append(g)
        self.g1d = None
        if need_input_grad:
            u1 = P.Units()
            for (g, sl, m, k, wn) in self.s1:
                u1.extend(P.conv_dgrad_units(ar.off(wn), m, Cin, k, k, (k - 1) // 2, cu0=sl // 8))
            self.g1d_units = u1

    def build_input_grad(self, dx: Act):
        """The stage-1 input gradient GEMM (one K-concatenation over every first conv) writing into dx."""
        B, H, W = self.x.N, self.x.H, self.x.W
        self.g1d = self.net.G(P.Geometry(N=B, H=H, W=W, ldx=self.L, x_coff=0, OH=H, OW=W, ldy=dx.ld, y_coff=dx.coff),
                              self.g1d_units, self.Cin, bwd=True)
        self.dx = dx

    def forward(self):
        net, relu = self.net, self.act
        for g in self.s1_fwd:
            g.fprop(self.x.t, self.mid_raw.t)
        self.nA.forward(self.mid_raw.slice(0, self.LA), self.mid_act.slice(0, self.LA), relu)
        if self.dw:
            nB = self.L - self.LA
            ops.dwconv_fwd(self.mid_act.slice(self.D0, self.D1 - self.D0), self.mid_raw.slice(self.LA, nB), self.dw_k, self.dw_w,
                           net.arena.p, PAD_ZERO)
            self.nB.forward(self.mid_raw.slice(self.LA, nB), self.mid_act.slice(self.LA, nB), relu)
        self.g2.fprop(self.mid_act.t, self.out_span.t, bias=net.biases.slot(self.s2_bias_slot))

    def backward(self, d_out: Act, dmid_act: Act, dmid_raw: Act):
        """d_out: gradient buffer with the layout of the output buffer (same pitch / offsets as out_span)."""
        net, ar, relu = self.net, self.net.arena, self.act
        assert d_out.ld == self.out_span.ld and d_out.coff == self.out_span.coff
        ops.channel_sum(Act(d_out.t, d_out.coff, self.out_span.C), net.biases.slot(self.s2_bias_slot, 'dvec'))

        def s2_wgrads():
            for (gw, seg) in self.s2_w:
                gw.wgrad(self.mid_act.t, d_out.t, ar.g)
        net._on_side(s2_wgrads)                 # reads d_out (owned by the caller's block) and this body's mid_act
        for g in self.d2:
            g.fprop(d_out.t, dmid_act.t)
        net._wait_raw()                         # the previous body's weight gradients have read the shared dmid_raw
        if self.dw:
            nB = self.L - self.LA
            self.nB.backward(dmid_act.slice(self.LA, nB), self.mid_act.slice(self.LA, nB), self.mid_raw.slice(self.LA, nB),
                             dmid_raw.slice(self.LA, nB), relu)
            net._on_side(lambda: ops.dwconv_bwd_weight(self.mid_act.slice(self.D0, self.D1 - self.D0), dmid_raw.slice(self.LA, nB),
                                                       self.dw_k, self.dw_w, ar.g, PAD_ZERO))
            ops.dwconv_bwd_data(dmid_raw.slice(self.LA, nB), dmid_act.slice(self.D0, self.D1 - self.D0), self.dw_k, self.dw_w,
                                ar.p, PAD_ZERO)
        self.nA.backward(dmid_act.slice(0, self.LA), self.mid_act.slice(0, self.LA), self.mid_raw.slice(0, self.LA),
                         dmid_raw.slice(0, self.LA), relu)

        def s1_wgrads():
            for (g, sl, m, k, wn) in self.s1:
                g.wgrad(self.x.t, dmid_raw.t, ar.g)
        net._on_side(s1_wgrads)
        net._mark_raw()
        if self.g1d is not None:
            self.g1d.fprop(dmid_raw.t, self.dx.t)


# ------------------------------------------------------------------------------------------------
# SPADE generator
# ------------------------------------------------------------------------------------------------
class _SpadeBlock:
    pass


class SpadeGenNet(_Net):
    """InceptionSPADEGenerator compiled for a fixed (B, H, W)."""

    UPSAMPLED = ('G_middle_0', 'up_0', 'up_1', 'up_2', 'up_3', 'up_4')

    def __init__(self, arch, seg: Act, device, training, need_grad, alloc_only=False, share=None):
        """seg: the persistent NHWC bf16 input buffer [B,H,W,cpad(semantic_nc)] (one-hot labels + edge map) every
        forward pass reads; teacher and student are compiled against the same buffer.
        alloc_only: only lay out the parameter / buffer tables (names and shapes of the reference state_dict)."""
        B, H, W = (seg.N, seg.H, seg.W) if seg is not None else (1, 0, 0)
        self._init_common(B, device, training, need_grad, share=share)
        self.bn_momentum = arch.get('momentum', 0.1)
        # generator activation: nn.ReLU on the distillation path (options/distill_options.py:123), nn.LeakyReLU() at its
        # default slope 0.01 when SPADEModel trains the teacher (models/spade_model.py:92)
        self.act = ACT[{'nn.ReLU': 'relu', 'nn.LeakyReLU': 'leaky001'}[arch.get('active_fn', 'nn.ReLU')]]
        self.overlap_wgrad = os.environ.get('CATB_NO_WOVERLAP', '0') != '1'
        self.arch, self.H, self.W = arch, H, W
        self.snc = arch['semantic_nc']
        assert alloc_only or seg.C == cpad(self.snc)
        self.seg_in = seg
        ks = arch['kernel_sizes']
        ar = self.arena
        # ---- parameters
        ar.alloc('fc.weight', (arch['fc_out'], self.snc, 3, 3))
        ar.alloc('fc.bias', (arch['fc_out'],))
        self.alloc_bn([('fc_norm', arch['fc_out'])], True)
        self.blocks = []
        for name in arch['block_names']:
            a = arch['blocks'][name]
            b = _SpadeBlock()
            b.name, b.fin, b.fout, b.learned = name, a['fin'], a['fout'], a['learned_shortcut']
            b.main = SixBranch(self, name, a['res'], a['dw'], ks, b.fin, '.conv', False)
            b.spade = SixBranch(self, name + '.spade', a['spade_res'], a['spade_dw'], ks, self.snc, '', True)
            b.main.alloc(b.fout)
            if b.learned:
                self.alloc_bn([(name + '.shortcut.0', b.fin)], True)
                ar.alloc(name + '.shortcut.1.conv.weight', (b.fout, b.fin, 1, 1))
            if not b.main.empty:
                self.alloc_bn([(name + '.spade.param_free_norm', b.fin)], False)
                b.spade.alloc(2 * b.fin)
            self.blocks.append(b)
        ar.alloc('conv_img.weight', (3, arch['final_nc'], 3, 3))
        ar.alloc('conv_img.bias', (3,))
        if alloc_only:
            return
        if share is None:
            ar.finalize(device)
            self.bufs.finalize(device)
        self._build()
        self._finish_build()

    def _act(self, h, w, C, zero=False):
        return Act.empty(self.B, h, w, C, self.dev, zero=zero)

    def _build(self):
        arch, B, H, W, dev, ar, ng = self.arch, self.B, self.H, self.W, self.dev, self.arena, self.need_grad
        more = arch['num_upsampling_layers'] in ('more', 'most')
        snc = self.snc
        h, w = arch['sh'], arch['sw']
        self.seg_pyr = {}       # (h, w) -> resized segmentation map (the full-resolution one is the input itself)

        def seg_at(h, w):
            if (h, w) != (H, W) and (h, w) not in self.seg_pyr:
                self.seg_pyr[(h, w)] = self._act(h, w, snc, zero=True)
            return (h, w)

        # fc (3x3, zero padding; its bias sits in front of fc_norm) on the latent-size map
        self.fc_res = seg_at(h, w)
        C0 = arch['fc_out']
        self.y_fc, self.a_fc = self._act(h, w, C0), self._act(h, w, C0)
        self.n_fc = self.make_bn(h * w, [('fc_norm', C0)], True, bias_slot=self.biases.add(_vec_offsets(ar, 'fc.bias', C0)))
        x = self.a_fc
        maxL = 8
        for b in self.blocks:
            if b.name in self.UPSAMPLED or (b.name == 'G_middle_1' and more):
                h, w = 2 * h, 2 * w
                b.up_in = x
                x = self._act(h, w, b.fin)
                b.up_out = x
            else:
                b.up_in = None
            b.h, b.w, b.x = h, w, x
            b.seg_res = seg_at(h, w)
            b.empty = b.main.empty
            if b.empty and not b.learned:     # forward returns its input (inception_modules.py:550-553)
                b.out = x
                if ng and b.up_in is not None:
                    b.d_up_in = self._act(h // 2, w // 2, b.fin)
                continue
            b.out = self._act(h, w, b.fout)
            if b.learned:
                b.xs = self._act(h, w, b.fin)
                b.n_sc = self.make_bn(h * w, [(b.name + '.shortcut.0', b.fin)], True)
                b.g_sc = self.G(P.Geometry(B, h, w, cpad(b.fin), 0, h, w, cpad(b.fout), 0),
                                P.conv_fprop_units(ar.off(b.name + '.shortcut.1.conv.weight'), b.fout, b.fin, 1, 1, 0), b.fout)
            if not b.empty:
                b.pfn = self.make_bn(h * w, [(b.name + '.spade.param_free_norm', b.fin)], False)
                Cp = cpad(b.fin)
                b.gb = self._act(h, w, 2 * Cp, zero=True)
                b.t = self._act(h, w, b.fin)
                b.has_gb = not b.spade.empty
                b.main.build(b.t, [(b.out, b.fout, 0)], need_input_grad=True)
                maxL = max(maxL, b.main.L * h * w)
            if ng:
                b.d_x = self._act(h, w, b.fin)          # gradient w.r.t. the block input
                if b.learned:
                    b.d_xs, b.dx_sc = self._act(h, w, b.fin), self._act(h, w, b.fin)
                    b.gb_sc = self.G(P.Geometry(B, h, w, cpad(b.fout), 0, h, w, cpad(b.fin), 0),
                                     P.conv_dgrad_units(ar.off(b.name + '.shortcut.1.conv.weight'), b.fout, b.fin, 1, 1, 0),
                                     b.fin, bwd=True)
                if not b.empty:
                    b.d_t, b.dn, b.dx_n = self._act(h, w, b.fin), self._act(h, w, b.fin), self._act(h, w, b.fin)
                    b.dgb = self._act(h, w, 2 * cpad(b.fin), zero=True)
                    b.main.build_input_grad(b.d_t)
                if b.up_in is not None:
                    b.d_up_in = self._act(h // 2, w // 2, b.fin)
            x = b.out
        self.feat_out = x
        Cf = arch['final_nc']
        self.l_img = self._act(H, W, Cf)
        self.out = self._act(H, W, 3)
        assert (h, w) == (H, W), f'generator output {h}x{w} does not match the compiled size {H}x{W}'
        self.g_img = self.G(P.Geometry(B, H, W, cpad(Cf), 0, H, W, 8, 0), P.conv_fprop_units(ar.off('conv_img.weight'), 3, Cf, 3, 3, 1), 3)
        self.img_bias = ar.view('conv_img.bias')
        self.acts = {b.name: b.out for b in self.blocks if b.name in MAPPING_LAYERS}
        # GEMMs that read the segmentation pyramid
        s0 = self._seg(self.fc_res)
        self.g_fc = self.G(P.Geometry(B, s0.H, s0.W, s0.ld, s0.coff, s0.H, s0.W, cpad(C0), 0),
                           P.conv_fprop_units(ar.off('fc.weight'), C0, snc, 3, 3, 1), C0)
        for b in self.blocks:
            if b.empty or not b.has_gb:
                continue
            Cp = cpad(b.fin)
            b.spade.build(self._seg(b.seg_res), [(b.gb.slice(0, Cp), b.fin, 0), (b.gb.slice(Cp, Cp), b.fin, b.fin)],
                          need_input_grad=False)
            maxL = max(maxL, b.spade.L * b.h * b.w)
        if ng:
            f = dict(dtype=ops.BF16, device=dev)
            self.ws_dmid_act = torch.zeros(B * maxL, **f)
            self.ws_dmid_raw = torch.zeros(B * maxL, **f)
            self.d_img_z = self._act(H, W, 3)
            self.d_l = self._act(H, W, Cf)
            self.d_feat = self._act(H, W, Cf)
            self.gb_img = self.G(P.Geometry(B, H, W, 8, 0, H, W, cpad(Cf), 0), P.conv_dgrad_units(ar.off('conv_img.weight'), 3, Cf, 3, 3, 1),
                                 Cf, bwd=True)
            self.d_a_fc, self.d_y_fc = self._act(arch['sh'], arch['sw'], C0), self._act(arch['sh'], arch['sw'], C0)

    def _seg(self, res):
        return self.seg_in if res == (self.H, self.W) else self.seg_pyr[res]

    def _ws(self, flat, h, w, C):
        return Act(flat[:self.B * h * w * C].view(self.B, h, w, C))

    # ---- forward -------------------------------------------------------------------------------
    def forward(self):
        """Reads the bound segmentation buffer.  Returns the output Act (tanh applied)."""
        seg = self.seg_in
        relu, none = self.act, ACT['none']
        self.pool_sums.zero_()
        self.biases.gather(self.arena.p)
        for (h, w), t in self.seg_pyr.items():
            ops.resize_nearest(seg, t)
        self.g_fc.fprop(self._seg(self.fc_res).t, self.y_fc.t)
        self.n_fc.forward(self.y_fc, self.a_fc, none)
        for b in self.blocks:
            if b.up_in is not None:
                ops.resize_nearest(b.up_in, b.up_out)
            if b.empty and not b.learned:
                continue
            if not b.empty:
                b.pfn.prepare(b.x)
                Cp = cpad(b.fin)
                if b.has_gb:
                    b.spade.forward()
                ops.spade_modulate(b.x, b.gb.slice(0, Cp), b.gb.slice(Cp, Cp), b.t, b.pfn.scale, b.pfn.shift, relu)
                b.main.forward()
            if b.learned:
                b.n_sc.forward(b.x, b.xs, none)
                b.g_sc.fprop(b.xs.t, b.out.t, accumulate=not b.empty)
            else:
                ops.add(b.out, b.x, b.out)
        ops.act_fwd(self.feat_out, self.l_img, ACT['leaky'])
        self.g_img.fprop(self.l_img.t, self.out.t, bias=self.img_bias, act=ACT['tanh'])
        return self.out

    # ---- backward ------------------------------------------------------------------------------
    def backward(self, d_out: Act, act_grads=None):
        """See _backward; the second stages of all weight gradients run as one launch when the pass is left."""
        if getattr(self, '_unpack', None) is None:
            self._unpack = ops.UnpackQueue(self.dev)
        with self._unpack:
            return self._backward(d_out, act_grads)

    def _backward(self, d_out: Act, act_grads=None):
        """d_out: gradient w.r.t. the tanh output; act_grads: {mapping layer: callable(Act)} accumulating the KA
        gradient into d(block output)."""
        assert self.need_grad
        relu, none, ar = self.act, ACT['none'], self.arena
        act_grads = act_grads or {}
        self.pool_red.zero_()
        if self.biases.used:
            self.biases.dvec.zero_()
        self._side_begin()
        ops.act_bwd(d_out, self.out, self.d_img_z, ACT['tanh'])
        self.g_img.wgrad(self.l_img.t, self.d_img_z.t, ar.g)
        ops.channel_sum(self.d_img_z, ar.view('conv_img.bias', 'g'))
        self.gb_img.fprop(self.d_img_z.t, self.d_l.t)
        ops.act_bwd(self.d_l, self.l_img, self.d_feat, ACT['leaky'])
        cur = self.d_feat
        for b in reversed(self.blocks):
            if b.name in act_grads:
                act_grads[b.name](cur)
            if not (b.empty and not b.learned):
                if b.learned:
                    b.g_sc.wgrad(b.xs.t, cur.t, ar.g)
                    b.gb_sc.fprop(cur.t, b.d_xs.t)
                    b.n_sc.backward(b.d_xs, None, b.x, b.dx_sc, none)
                    short = b.dx_sc
                else:
                    short = cur
                if not b.empty:
                    Cp = cpad(b.fin)
                    b.main.backward(cur, self._ws(self.ws_dmid_act, b.h, b.w, b.main.L), self._ws(self.ws_dmid_raw, b.h, b.w, b.main.L))
                    ops.spade_modulate_bwd(b.d_t, b.t, b.x, b.gb.slice(0, Cp), b.dgb.slice(0, Cp), b.dgb.slice(Cp, Cp), b.dn,
                                           b.pfn.scale, b.pfn.shift, relu)
                    b.pfn.backward(b.dn, None, b.x, b.dx_n, none)
                    if b.has_gb:
                        b.spade.backward(b.dgb, self._ws(self.ws_dmid_act, b.h, b.w, b.spade.L), self._ws(self.ws_dmid_raw, b.h, b.w, b.spade.L))
                    ops.add(b.dx_n, short, b.d_x)
                    cur = b.d_x
                else:
                    cur = short
            if b.up_in is not None:
                ops.upsample2x_bwd(cur, b.d_up_in)
                cur = b.d_up_in
        self.n_fc.backward(cur, None, self.y_fc, self.d_y_fc, none)
        self.g_fc.wgrad(self._seg(self.fc_res).t, self.d_y_fc.t, ar.g)
        self._side_end()
        self.biases.scatter(ar.g)


# ------------------------------------------------------------------------------------------------
# multi-scale discriminator with spectral norm
# ------------------------------------------------------------------------------------------------
def spade_D_layers(arch):
    """(conv index, cin, cout, stride, has_norm, has_act) of SPADENLayerDiscriminator (discriminators.py:140-170)."""
    nf, n_layers = arch['ndf'], arch['n_layers']
    layers = [(0, arch['input_nc'], nf, 2, False, True)]
    for n in range(1, n_layers):
        prev, nf = nf, min(nf * 2, 512)
        layers.append((n, prev, nf, 1 if n == n_layers - 1 else 2, True, True))
    layers.append((n_layers, nf, 1, 1, False, False))
    return layers


class MultiScaleDis:
    """MultiscaleDiscriminator compiled for a fixed batch of N = 2B images (the reference discriminates the
    batch-concatenation [fake; real] in one pass, spade_model_modules.py:136-156).  All sub-discriminators share one
    parameter arena (one Adam launch, one all-reduce); spectrally normalised weights are materialised in `w_eff`
    and the GEMM images are packed from there."""

    def __init__(self, arch, N, H, W, device, alloc_only=False):
        assert arch['norm_D'] == 'spectralinstance', arch['norm_D']
        self.arch, self.N, self.H, self.W, self.dev = arch, N, H, W, device
        self.arena, self.bufs = Arena(True), Arena(False)
        norm_arch = dict(norm='instance', affine=False, track_running_stats=False, eps=1e-5, momentum=0.1, use_bias=False)
        norm_arch.update(arch)
        n_layers = arch['n_layers']
        self.nets, self.inputs = [], []
        h, w = H, W
        for d in range(arch['num_D']):
            def names(ci, d=d):
                sn = 0 < ci < n_layers
                base = f'discriminator_{d}.model{ci}.0' + ('.0' if sn else '')
                return base + ('.weight_orig' if sn else '.weight'), base + '.bias', f'discriminator_{d}.model{ci}.0.1'
            net = DisNet(norm_arch, N, h, w, device, layers=spade_D_layers(arch), pad=2, names=names,
                         arenas=(self.arena, self.bufs))
            self.nets.append(net)
            h, w = (h + 1) // 2, (w + 1) // 2
        sn_rows = []
        for d, net in enumerate(self.nets):
            for L in net.layers:
                if L.has_norm:
                    base = L.wn[:-len('.weight_orig')]
                    self.bufs.alloc(base + '.weight_u', (L.cout,))
                    self.bufs.alloc(base + '.weight_v', (L.cin * 16,))
                    sn_rows.append((L.wn, L.cout, L.cin * 16, base + '.weight_u', base + '.weight_v'))
        if alloc_only:
            return
        self.arena.finalize(device)
        self.bufs.finalize(device)
        self.w_eff = torch.zeros_like(self.arena.p)
        tab = np.array([[self.arena.off(wn), r, c, self.bufs.off(u), self.bufs.off(v), 0] for (wn, r, c, u, v) in sn_rows], dtype=np.int32)
        self.sn_table = torch.from_numpy(tab).to(device)
        self.sn_n, self.sn_rows, self.sn_cols = len(sn_rows), int(tab[:, 1].max()), int(tab[:, 2].max())
        f32 = dict(dtype=torch.float32, device=device)
        self.sn_tmp = torch.zeros(self.sn_n * max(self.sn_rows, self.sn_cols), **f32)
        self.sn_sigma = torch.ones(self.sn_n, **f32)
        self.sn_cdot = torch.zeros(self.sn_n, **f32)
        for i, net in enumerate(self.nets):
            net.w_src = self.w_eff
            net.build()
            if i > 0:
                self.inputs.append(Act.empty(N, net.H, net.W, arch['input_nc'], device, zero=True))
        self.d_in = Act.empty(N, H, W, arch['input_nc'], device, zero=True)

    def load_state_dict(self, sd):
        self.arena.load_state_dict(sd)
        self.bufs.load_state_dict(sd)
        self.spectral_forward(training=False)

    def state_dict(self):
        sd = self.arena.state_dict()
        sd.update(self.bufs.state_dict())
        return sd

    def spectral_forward(self, training=True):
        """One power iteration (training) + W / sigma for every spectrally normalised conv, then re-pack."""
        self.w_eff.copy_(self.arena.p)
        ops.sn_forward(self.sn_table, self.sn_n, self.sn_rows, self.sn_cols, self.arena.p, self.bufs.p, training, self.sn_tmp,
                       self.sn_sigma, self.w_eff)
        for net in self.nets:
            net.pack_weights()

    def pack_weights(self):
        """After an optimiser step nothing needs re-packing here: every forward pass starts with spectral_forward."""

    def forward(self, x: Act):
        """x: [N,H,W,cpad(input_nc)].  Returns the list of sub-discriminator nets (activations: net.layers[i].a /
        .yraw, prediction net.pred)."""
        self.spectral_forward(training=True)
        cur = x
        for i, net in enumerate(self.nets):
            if i > 0:
                ops.avgpool3s2(cur, self.inputs[i - 1])
                cur = self.inputs[i - 1]
            net.forward(cur)
        return self.nets

    def backward(self, dpreds, param_grads, input_grad, act_grad_hook=None):
        """See _backward; the second stages of all weight gradients run as one launch when the pass is left."""
        if getattr(self, '_unpack', None) is None:
            self._unpack = ops.UnpackQueue(self.arena.p.device)
        with self._unpack:
            return self._backward(dpreds, param_grads, input_grad, act_grad_hook)

    def _backward(self, dpreds, param_grads, input_grad, act_grad_hook=None):
        """dpreds[i]: gradient w.r.t. the prediction of scale i.  Returns d(input) (scale-0 resolution)."""
        for i in range(len(self.nets) - 1, -1, -1):
            net = self.nets[i]
            hook = (lambda li, d, i=i: act_grad_hook(i, li, d)) if act_grad_hook is not None else None
            net.backward(dpreds[i], param_grads, input_grad, act_grad_hook=hook)
        if input_grad:
            # d(input of scale i) flows to scale i-1 through the average pool
            for i in range(len(self.nets) - 1, 0, -1):
                tgt = self.nets[i - 1].d_in
                ops.avgpool3s2_bwd(self.nets[i].d_in, tgt, add=tgt)
            return self.nets[0].d_in
        return None

    def finish_param_grads(self):
        """Gradient w.r.t. W / sigma (accumulated by the weight-gradient GEMMs) -> gradient w.r.t. weight_orig."""
        ops.sn_backward(self.sn_table, self.sn_n, self.sn_rows, self.sn_cols, self.arena.g, self.w_eff, self.bufs.p, self.sn_sigma,
                        self.sn_cdot)


def dis_feature(net, li):
    """Intermediate output li of a sub-discriminator as returned by SPADENLayerDiscriminator.forward."""
    L = net.layers[li]
    return L.a if L.has_norm else L.yraw


# ------------------------------------------------------------------------------------------------
# VGG19 perceptual features
# ------------------------------------------------------------------------------------------------
VGG_CFG = [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 256, 'M', 512, 512, 512, 512, 'M', 512]
VGG_TAP_CONVS = [0, 5, 10, 19, 28]       # torchvision features indices of conv{1..5}_1
VGG_WEIGHTS = [1.0 / 32, 1.0 / 16, 1.0 / 8, 1.0 / 4, 1.0]


class _VLayer:
    pass


class VggNet(_Net):
    """torchvision vgg19.features[0:30] (VGG19 of models/modules/loss.py:151-184) for a fixed (B, H, W): forward
    with bias + ReLU fused into the GEMM epilogue, 2x2 max pools, and the input gradient (no weight gradients:
    the network is frozen)."""

    def __init__(self, B, H, W, device, need_grad=True, alloc_only=False):
        self._init_common(B, device, False, False)
        self.need_input_grad = need_grad
        self.layers = []
        idx, cin, h, w = 0, 3, H, W
        for c in VGG_CFG:
            L = _VLayer()
            if c == 'M':
                L.kind, L.h, L.w, L.C = 'pool', h, w, cin
                h, w = h // 2, w // 2
                idx += 1
            else:
                L.kind, L.idx, L.cin, L.cout, L.h, L.w = 'conv', idx, cin, c, h, w
                self.arena.alloc(f'{idx}.weight', (c, cin, 3, 3))
                self.arena.alloc(f'{idx}.bias', (c,))
                L.tap = VGG_TAP_CONVS.index(idx) if idx in VGG_TAP_CONVS else None
                cin = c
                idx += 2
            self.layers.append(L)
        if alloc_only:
            return
        self.arena.finalize(device)
        self.bufs.finalize(device)
        ar = self.arena
        for L in self.layers:
            if L.kind == 'pool':
                L.out = Act.empty(B, L.h // 2, L.w // 2, L.C, device)
                if need_grad:
                    L.d_out = Act.empty(B, L.h // 2, L.w // 2, L.C, device)
                continue
            L.out = Act.empty(B, L.h, L.w, L.cout, device)
            L.g = self.G(P.Geometry(B, L.h, L.w, cpad(L.cin), 0, L.h, L.w, cpad(L.cout), 0),
                         P.conv_fprop_units(ar.off(f'{L.idx}.weight'), L.cout, L.cin, 3, 3, 1), L.cout)
            L.bias = ar.view(f'{L.idx}.bias')
            if L.tap is not None:
                L.ref = Act.empty(B, L.h, L.w, L.cout, device)       # features of the real image at this tap
                if need_grad and L.tap < len(VGG_TAP_CONVS) - 1:
                    L.d_above = Act.empty(B, L.h, L.w, L.cout, device)   # gradient arriving from the layers above
            if need_grad:
                L.d_out = Act.empty(B, L.h, L.w, L.cout, device, zero=True)   # gradient w.r.t. the ReLU output
                L.dz = Act.empty(B, L.h, L.w, L.cout, device)
                L.gb = self.G(P.Geometry(B, L.h, L.w, cpad(L.cout), 0, L.h, L.w, cpad(L.cin), 0),
                              P.conv_dgrad_units(ar.off(f'{L.idx}.weight'), L.cout, L.cin, 3, 3, 1), L.cin, bwd=True)
        if need_grad:
            self.d_in = Act.empty(B, H, W, 3, device)
        self._finish_build()

    def forward(self, x: Act, save_ref=False):
        cur = x
        for L in self.layers:
            if L.kind == 'pool':
                ops.maxpool2(cur, L.out)
            else:
                L.x = cur
                L.g.fprop(cur.t, L.out.t, bias=L.bias, act=ACT['relu'])
                if save_ref and L.tap is not None:
                    ops.copy_channels(L.out, L.ref, L.out.C)
            cur = L.out

    def _grad_target(self, i):
        """Where the gradient w.r.t. the output of layer i is written by the layer above it."""
        L = self.layers[i]
        return L.d_above if (L.kind == 'conv' and L.tap is not None and hasattr(L, 'd_above')) else L.d_out

    def loss_and_backward(self, loss_slots, grad_scale):
        """sum_i w_i * L1(features_i(x), ref_i) (VGGLoss.forward, loss.py:195-203): loss_slots[i] += the unweighted
        mean of tap i; back-propagates grad_scale * d(loss)/dx into self.d_in (returned)."""
        last = max(i for i, L in enumerate(self.layers) if L.kind == 'conv' and L.tap is not None)
        for i in range(last, -1, -1):
            L = self.layers[i]
            tgt = self._grad_target(i - 1) if i > 0 else self.d_in
            if L.kind == 'pool':
                ops.maxpool2_bwd(L.d_out, self.layers[i - 1].out, tgt)
                continue
            d = L.d_out
            if L.tap is not None:
                ops.recon_loss(L.out, L.ref, L.cout, 'l1', grad_scale * VGG_WEIGHTS[L.tap], loss_slots[L.tap:L.tap + 1], L.d_out,
                               L.d_above if i < last else None)
            ops.act_bwd(d, L.out, L.dz, ACT['relu'])
            L.gb.fprop(L.dz.t, tgt.t)
        return self.d_in
