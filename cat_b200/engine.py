"""Execution engine for the CAT distillation step on libcatb200 kernels.

The reference runs the step as ~600 eager ATen calls under autograd.  Here each network
is compiled once, for a fixed batch shape, into a static sequence of kernel launches with a hand-derived
backward; every buffer is pre-allocated, nothing synchronises, so the whole step can be replayed from a
CUDA graph.  Parameters live in flat fp32 arenas (reference layout per tensor, so ``state_dict``s
round-trip), gradients / Adam state in parallel arenas (one NCCL all-reduce per optimiser).

Networks restated here (same maths as the reference modules, different execution):
  GenNet  -- InceptionGenerator (models/modules/inception_architecture/inception_generator.py:37-142)
             with InvertedResidualChannels blocks (models/modules/inception_modules.py:124-236)
  DisNet  -- NLayerDiscriminator (models/modules/discriminators.py:14-79)
"""
import math
import os
from collections import OrderedDict

import torch

from . import igemm_plan as P
from . import ops
from .igemm_plan import cpad
from .ops import ACT, Act, Gemm

MAPPING_LAYERS = ['down_sampling.9', 'features.2', 'features.5', 'features.8']


# ------------------------------------------------------------------------------------------------
# parameter arenas
# ------------------------------------------------------------------------------------------------
class Arena:
    """Flat fp32 storage for a set of named tensors.  Entries may be padded (norm vectors are padded
    to the 8-channel unit so that kernels can index the padded channel range directly)."""

    def __init__(self, with_grad):
        self.entries = OrderedDict()  # name -> (offset, shape, padded_len, pad_value)
        self.size = 8
        self.with_grad = with_grad
        self.p = self.g = self.m = self.v = None

    def alloc(self, name, shape, pad_to=None, pad_value=0.0):
        n = int(math.prod(shape))
        L = max(n, pad_to or n)
        if self.p is not None:   # finalised arena shared by a second compilation of the same network: look up
            assert name in self.entries and self.entries[name][1] == tuple(shape), name
            return self.entries[name][0]
        assert name not in self.entries, name
        self.entries[name] = (self.size, tuple(shape), L, pad_value)
        self.size += L
        return self.entries[name][0]

    def finalize(self, device):
        total = self.size + 64
        self.p = torch.zeros(total, dtype=torch.float32, device=device)
        for (off, shape, L, pv) in self.entries.values():
            n = int(math.prod(shape))
            if L > n and pv != 0.0:
                self.p[off + n:off + L] = pv
        if self.with_grad:
            self.g = torch.zeros_like(self.p)
            self.m = torch.zeros_like(self.p)
            self.v = torch.zeros_like(self.p)

    def off(self, name):
        return self.entries[name][0]

    def has(self, name):
        return name in self.entries

    def view(self, name, which='p'):
        off, shape, L, _ = self.entries[name]
        n = int(math.prod(shape))
        return getattr(self, which)[off:off + n].view(shape)

    def padded(self, name, which='p'):
        off, shape, L, _ = self.entries[name]
        return getattr(self, which)[off:off + L]

    def span(self, first, last, which='p'):
        """Contiguous slice covering entries first..last (allocated back to back)."""
        o0 = self.entries[first][0]
        o1, _, L1, _ = self.entries[last]
        return getattr(self, which)[o0:o1 + L1]

    def load_state_dict(self, sd, prefix='', strict=False):
        """Copies the entries of `sd` that this arena holds.  A shape mismatch always raises (copy_ would broadcast, e.g. a
        checkpoint of a differently pruned student); entries of the arena that `sd` lacks keep their current values and
        are returned (strict=True raises instead) -- a network's parameters and buffers live in two arenas that are both
        loaded from one state dict, so unexpected keys are normal and not reported."""
        missing = []
        for name in self.entries:
            key = prefix + name
            if key not in sd:
                missing.append(key)
                continue
            dst, src = self.view(name), sd[key]
            if tuple(dst.shape) != tuple(src.shape) and dst.numel() != src.numel():
                raise ValueError(f'{key}: checkpoint shape {tuple(src.shape)} does not match {tuple(dst.shape)}')
            dst.copy_(src.to(torch.float32).reshape(dst.shape))
        if strict and missing:
            raise KeyError(f'state dict lacks {len(missing)} entries, e.g. {missing[:4]}')
        return missing

    def state_dict(self, prefix=''):
        return OrderedDict((prefix + n, self.view(n).detach().clone().cpu()) for n in self.entries)


# ------------------------------------------------------------------------------------------------
# normalisation layer over a channel range of a (raw, activated) buffer pair
# ------------------------------------------------------------------------------------------------
class Norm:
    def __init__(self, dev, N, HW, Cp, kind, eps, momentum, training, track, gamma=None, beta=None, rmean=None,
                 rvar=None, dgamma=None, dbeta=None):
        self.per_sample = kind == 'instance'
        if self.per_sample and track:
            raise NotImplementedError('InstanceNorm with running statistics is not on the CAT distillation path')
        self.N, self.HW, self.Cp = N, HW, Cp
        self.G = N if self.per_sample else 1
        self.count = HW if self.per_sample else N * HW
        self.eps, self.momentum = eps, momentum
        self.batch_stats = self.per_sample or training or not track
        self.training, self.track = training, track
        self.gamma, self.beta, self.rmean, self.rvar = gamma, beta, rmean, rvar
        self.dgamma, self.dbeta = dgamma, dbeta
        f = dict(dtype=torch.float32, device=dev)
        self.sums = torch.zeros(self.G, 2, Cp, **f)
        self.scale = torch.empty(self.G, Cp, **f)
        self.shift = torch.empty(self.G, Cp, **f)
        self.mean_rstd = torch.empty(self.G, 2, Cp, **f)
        self.red = torch.zeros(self.G, 2, Cp, **f)
        self._frozen = False
        self.pooled = False   # True once sums / red live in a per-network pool that is zeroed once per pass

    def set_training(self, training):
        """nn.Module.train() / eval() on the layer this object stands for."""
        self.training = training
        self.batch_stats = self.per_sample or training or not self.track
        self._frozen = False

    def stats_for(self, x: Act, y_coff):
        """The `stats` argument of Gemm.fprop for a conv that writes channels [y_coff, ...) of the buffer this layer
        normalises as slice `x` (None when the layer uses running statistics: nothing to accumulate)."""
        if not self.batch_stats:
            return None
        return (self.sums, self.Cp, y_coff - x.coff, self.per_sample)

    def forward(self, x: Act, y: Act, act, residual=None, have_stats=False):
        """have_stats: the producing convs accumulated sum / sum of squares in their epilogues (Gemm.fprop(stats=...)
        returned True for every one of them), so the statistics pass over x is skipped."""
        if self.batch_stats:
            if not have_stats:
                if not self.pooled:
                    self.sums.zero_()
                ops.norm_stats(x, self.per_sample, self.sums)
            upd = self.training and self.track
            # finalize folded into the apply kernel: one launch
            ops.norm_apply_fused(x, y, self.sums, self.count, self.eps, self.momentum, self.gamma, self.beta,
                                 self.rmean if upd else None, self.rvar if upd else None, self.scale, self.shift,
                                 self.mean_rstd, self.per_sample, act, residual)
            return
        elif not self._frozen:
            ops.norm_finalize(None, self.G, self.Cp, self.count, self.eps, self.momentum, self.gamma, self.beta,
                              self.rmean, self.rvar, self.scale, self.shift, self.mean_rstd)
            # eval-mode affine of a frozen net is computed once; a net that is being optimised in eval mode (the
            # reference's first distillation step, see backward) recomputes it from the current gamma / beta
            self._frozen = not self.training and self.dgamma is None
        ops.norm_apply(x, y, self.scale, self.shift, self.per_sample, act, residual)

    def backward(self, dout: Act, out, x: Act, dx: Act, act, param_grads=True):
        if not self.pooled:
            self.red.zero_()
        ops.norm_bwd_reduce(dout, out, x, self.per_sample, self.mean_rstd, act, self.red)
        # Eval-mode BatchNorm (the student until the first evaluate_model: model_profiling leaves it in eval(),
        # utils/model_profiling.py:299, inception_distiller.py:280 -- the reference's FIRST step back-propagates through
        # running statistics): mean / rstd are constants, so dx = gamma * rstd * dz and the batch-mean terms of the
        # training-mode formula drop out -- the same kernel with an infinite element count; d gamma / d beta are the
        # same reductions against the running statistics saved by the forward pass.
        count = self.count if self.batch_stats else float('inf')
        ops.norm_bwd_apply(dout, out, x, dx, self.per_sample, self.mean_rstd, self.gamma, self.red, count, act,
                           self.dgamma if param_grads else None, self.dbeta if param_grads else None)


def conv_norm(gemms, x_t, y: Act, norm: Norm, nx: Act, ny: Act, act, residual=None, **fprop_kw):
    """conv (one GEMM, or several writing disjoint pixels / channels of y) + norm + activation: the GEMM epilogues
    accumulate the layer's statistics when every one of them can (halo kernels); otherwise the statistics pass runs."""
    if not isinstance(gemms, (list, tuple)):
        gemms = [gemms]
    fused = []
    for g in gemms:
        fused.append(g.fprop(x_t, y.t, stats=norm.stats_for(nx, g.geo.y_coff), **fprop_kw))
    ok = all(fused)
    if any(fused) and not ok:      # mixed: discard the partial sums, the statistics pass recomputes them
        norm.sums.zero_()
    norm.forward(nx, ny, act, residual=residual, have_stats=ok)


def pool_norm_buffers(norms, dev):
    """Move the atomic-accumulation buffers of all norm layers of a network into two flat tensors, so that
    one memset per forward / backward pass replaces two per layer."""
    n_s = sum(n.sums.numel() for n in norms)
    n_r = sum(n.red.numel() for n in norms)
    pool_s = torch.zeros(max(n_s, 1), dtype=torch.float32, device=dev)
    pool_r = torch.zeros(max(n_r, 1), dtype=torch.float32, device=dev)
    o_s = o_r = 0
    for n in norms:
        n.sums = pool_s[o_s:o_s + n.sums.numel()].view(n.sums.shape)
        n.red = pool_r[o_r:o_r + n.red.numel()].view(n.red.shape)
        o_s += n.sums.numel()
        o_r += n.red.numel()
        n.pooled = True
    return pool_s, pool_r


class _NormSpec:
    """Allocation helper that keeps gamma / beta / running stats of several norm layers contiguous and
    in the same (padded) channel order as the activation slices they normalise."""

    def __init__(self, arena: Arena, bufs: Arena, arch):
        self.arena, self.bufs, self.arch = arena, bufs, arch
        self.created = []

    def alloc_group(self, prefixes_and_C):
        a, b, arch = self.arena, self.bufs, self.arch
        track = arch['norm'] == 'batch' and arch['track_running_stats']
        if arch['affine']:
            for p, C in prefixes_and_C:
                a.alloc(p + '.weight', (C,), cpad(C), 1.0)
            for p, C in prefixes_and_C:
                a.alloc(p + '.bias', (C,), cpad(C), 0.0)
        if track:
            for p, C in prefixes_and_C:
                b.alloc(p + '.running_mean', (C,), cpad(C), 0.0)
            for p, C in prefixes_and_C:
                b.alloc(p + '.running_var', (C,), cpad(C), 1.0)
            for p, C in prefixes_and_C:
                b.alloc(p + '.num_batches_tracked', (1,))

    def make(self, dev, N, HW, prefixes_and_C, training):
        a, b, arch = self.arena, self.bufs, self.arch
        track = arch['norm'] == 'batch' and arch['track_running_stats']
        first, last = prefixes_and_C[0][0], prefixes_and_C[-1][0]
        Cp = sum(cpad(C) for _, C in prefixes_and_C)
        kw = {}
        if arch['affine']:
            kw['gamma'] = a.span(first + '.weight', last + '.weight')
            kw['beta'] = a.span(first + '.bias', last + '.bias')
            if a.with_grad:
                kw['dgamma'] = a.span(first + '.weight', last + '.weight', 'g')
                kw['dbeta'] = a.span(first + '.bias', last + '.bias', 'g')
        if track:
            kw['rmean'] = b.span(first + '.running_mean', last + '.running_mean')
            kw['rvar'] = b.span(first + '.running_var', last + '.running_var')
        for t in kw.values():
            assert t.numel() == Cp
        n = Norm(dev, N, HW, Cp, arch['norm'], arch['eps'], arch['momentum'], training, track, **kw)
        self.created.append(n)
        return n


def _strided_dgrad(geo_kw, units, n_rows, dev, stride):
    """Input-gradient GEMMs of a stride-`stride` zero-padded conv: one launch (stride 1) or one per
    output parity class with only the taps that hit it (4-phase sub-pixel decomposition)."""
    if stride == 1:
        return [Gemm(P.Geometry(**geo_kw), units, n_rows, dev)]
    gemms = []
    for a in range(2):
        for b in range(2):
            ph = units.phase(a, b)
            if len(ph):
                gemms.append(Gemm(P.Geometry(**geo_kw, sn=1, sd=2, o_step=2, o_ph=a, o_pw=b), ph, n_rows, dev))
    return gemms


# ------------------------------------------------------------------------------------------------
# generator
# ------------------------------------------------------------------------------------------------
class _Block:
    pass


def _stage1_order(res, dw):
    """Slice order of the first-stage outputs in the mid buffer: all 1x1 convs first (the k=1 res branch and
    the first conv of every dw branch -- they share one N-concatenated GEMM), then the k>1 res branches.
    Entries: (kind, j, m, k_of_first_conv)."""
    return ([('res', j, m, k) for (j, m, k) in res if k == 1] + [('dw', j, m, 1) for (j, m, k) in dw] +
            [('res', j, m, k) for (j, m, k) in res if k > 1])


class GenNet:
    """InceptionGenerator compiled for a fixed (B, H, W)."""

    def __init__(self, arch, B, H, W, device, training, need_grad, share=None, input_grad=False):
        """share: another GenNet of the same architecture whose parameter / buffer arenas are reused (e.g. an
        eval-mode or different-resolution compilation of the same weights, or a second application of the same
        generator inside one CycleGAN step).
        input_grad: backward() also produces the gradient w.r.t. the input image in self.d_in (CycleGAN's
        rec_A = G_B(G_A(real_A)) back-propagates through the input of G_B, models/cycle_gan_model.py:221-226)."""
        assert H % 4 == 0 and W % 4 == 0, 'the generator down-samples twice'
        self.arch, self.B, self.H, self.W, self.dev = arch, B, H, W, device
        self.training, self.need_grad = training, need_grad
        self.input_grad = input_grad and need_grad
        self.use_bias = arch['use_bias']
        # x-packed 7x7 stem / head (cat_b200/csrc/packx.cu): the seven horizontal taps become channels of a 7x1 conv
        self.packx = os.environ.get('CATB_NO_PACKX', '0') != '1' and 7 * arch['input_nc'] <= 24 and arch['output_nc'] <= 8
        # weight-gradient GEMMs of the residual blocks run on a side stream next to the input-gradient chain
        self.overlap_wgrad = os.environ.get('CATB_NO_WOVERLAP', '0') != '1' and str(device) != 'cpu'
        self._wside = None
        if share is not None:
            assert share.arch == arch and (share.arena.with_grad or not need_grad)
            self.arena, self.bufs, self.aux = share.arena, share.bufs, share.aux
            self.aux_idx = share.aux_idx
            self.ns = _NormSpec(self.arena, self.bufs, arch)
        else:
            self.arena, self.bufs = Arena(with_grad=need_grad), Arena(with_grad=False)
            self.ns = _NormSpec(self.arena, self.bufs, arch)
            self._alloc_params()
            self.arena.finalize(device)
            self.bufs.finalize(device)
            self._alloc_aux(device)
        self._build()
        self.pool_sums, self.pool_red = pool_norm_buffers(self.ns.created, device)

    # ---- parameters --------------------------------------------------------------------------
    def _conv_alloc(self, name, shape, bias, transposed=False):
        """Conv2d weight [Cout,Cin,k,k] (ConvTranspose2d: [Cin,Cout,k,k]) and optional bias [Cout]."""
        self.arena.alloc(name + '.weight', shape)
        if bias:
            self.arena.alloc(name + '.bias', (shape[1] if transposed else shape[0],))

    def _alloc_params(self):
        A, ub = self.arch, self.use_bias
        c0, c1, c2, c3, c4 = A['widths']
        ks = A['kernel_sizes']
        self._conv_alloc('down_sampling.1', (c0, A['input_nc'], 7, 7), ub)
        self.ns.alloc_group([('down_sampling.2', c0)])
        self._conv_alloc('down_sampling.4', (c1, c0, 3, 3), ub)
        self.ns.alloc_group([('down_sampling.5', c1)])
        self._conv_alloc('down_sampling.7', (c2, c1, 3, 3), ub)
        self.ns.alloc_group([('down_sampling.8', c2)])
        for i, blk in enumerate(A['blocks']):
            pre = f'features.{i}'
            res = [(j, m, k) for j, (m, k) in enumerate((mk for mk in zip(blk['res'], ks) if mk[0] > 0))]
            dw = [(j, m, k) for j, (m, k) in enumerate((mk for mk in zip(blk['dw'], ks) if mk[0] > 0))]
            for j, m, k in res:
                self._conv_alloc(f'{pre}.res_ops.{j}.1.0', (m, c2, k, k), ub)
                self._conv_alloc(f'{pre}.res_ops.{j}.4', (c2, m, k, k), ub)
            for j, m, k in dw:
                self._conv_alloc(f'{pre}.dw_ops.{j}.0.0', (m, c2, 1, 1), ub)
                self._conv_alloc(f'{pre}.dw_ops.{j}.2.0', (m, 1, k, k), ub)
                self._conv_alloc(f'{pre}.dw_ops.{j}.4', (c2, m, 1, 1), ub)
            grpA = [(f'{pre}.res_ops.{j}.1.1' if kind == 'res' else f'{pre}.dw_ops.{j}.0.1', m)
                    for (kind, j, m, _k) in _stage1_order(res, dw)]
            grpB = [(f'{pre}.dw_ops.{j}.2.1', m) for j, m, k in dw]
            if grpA:
                self.ns.alloc_group(grpA)
            if grpB:
                self.ns.alloc_group(grpB)
            if res or dw:
                self.ns.alloc_group([(f'{pre}.pw_bn', c2)])
        self._conv_alloc('up_sampling.0', (c2, c3, 3, 3), ub, transposed=True)
        self.ns.alloc_group([('up_sampling.1', c3)])
        self._conv_alloc('up_sampling.3', (c3, c4, 3, 3), ub, transposed=True)
        self.ns.alloc_group([('up_sampling.4', c4)])
        self._conv_alloc('up_sampling.7', (A['output_nc'], c4, 7, 7), True)

    def _alloc_aux(self, device):
        """Derived weight tensors of the x-packed stem / head, refreshed from the parameters before every packing:
        stem.wx[co, dx*cin + ci, dy] = W[co, ci, dy, dx];  head.wx[co*8 + dx, ci, dy] = W[co, ci, dy, dx].  Their
        gradients accumulate in aux.g and are scattered back through the same index table."""
        import numpy as np
        A = self.arch
        self.aux = Arena(with_grad=self.need_grad)
        self.aux_idx = None
        if not self.packx:
            return
        c0, c4, cin, cout = A['widths'][0], A['widths'][4], A['input_nc'], A['output_nc']
        Cx = cpad(7 * cin)
        o_stem = self.aux.alloc('stem.wx', (c0, Cx, 7, 1))
        o_head = self.aux.alloc('head.wx', (cout * 8, c4, 7, 1))
        self.aux.finalize(device)
        idx = -np.ones(self.aux.p.numel(), dtype=np.int32)
        w = self.arena.off('down_sampling.1.weight')
        co, ch, dy = np.meshgrid(np.arange(c0), np.arange(7 * cin), np.arange(7), indexing='ij')
        dx, ci = ch // cin, ch % cin
        idx[o_stem + (co * Cx + ch) * 7 + dy] = w + ((co * cin + ci) * 7 + dy) * 7 + dx
        w = self.arena.off('up_sampling.7.weight')
        co, dx, ci, dy = np.meshgrid(np.arange(cout), np.arange(7), np.arange(c4), np.arange(7), indexing='ij')
        idx[o_head + ((co * 8 + dx) * c4 + ci) * 7 + dy] = w + ((co * c4 + ci) * 7 + dy) * 7 + dx
        self.aux_idx = torch.from_numpy(idx).view(1, -1).to(device)

    def load_state_dict(self, sd):
        self.arena.load_state_dict(sd)
        self.bufs.load_state_dict(sd)
        self.pack_weights()

    def state_dict(self):
        sd = self.arena.state_dict()
        sd.update(self.bufs.state_dict())
        return sd

    def set_training(self, training):
        """train() / eval() of the generator: only the normalisation layers depend on it (dropout_rate is 0).  The launch
        sequence changes (statistics kernels), so a captured CUDA graph of this network must be re-captured."""
        self.training = training
        for n in self.ns.created:
            n.set_training(training)

    # ---- graph construction --------------------------------------------------------------------
    def _act(self, H, W, C, zero=False):
        return Act.empty(self.B, H, W, C, self.dev, zero=zero)

    def _build(self):
        A, B, H, W, dev = self.arch, self.B, self.H, self.W, self.dev
        ar, ng, tr = self.arena, self.need_grad, self.training
        c0, c1, c2, c3, c4 = A['widths']
        cin, cout = A['input_nc'], A['output_nc']
        ks = A['kernel_sizes']
        H2, W2, H4, W4 = H // 2, W // 2, H // 4, W // 4
        self.fprop_gemms, self.bwd_gemms = [], []   # everything that needs packed weights

        def G(geo, units, n_rows, bwd=False):
            g = Gemm(geo, units, n_rows, dev)
            (self.bwd_gemms if bwd else self.fprop_gemms).append(g)
            return g

        # ---- stem / down-sampling
        self.x_in = None  # set per forward (shared NHWC input)
        self.y0, self.a0 = self._act(H, W, c0), self._act(H, W, c0)
        self.aux_gemms = []
        if self.packx:
            Cx = cpad(7 * cin)
            self.x_stem = self._act(H, W, Cx)
            self.g_stem = Gemm(P.Geometry(B, H, W, Cx, 0, H, W, cpad(c0), 0, pad_mode=P.PAD_REFLECT),
                               P.conv_fprop_units(self.aux.off('stem.wx'), c0, Cx, 7, 1, 3, pad_s=0), c0, dev)
            self.aux_gemms.append(self.g_stem)
        else:
            self.g_stem = G(P.Geometry(B, H, W, cpad(cin), 0, H, W, cpad(c0), 0, pad_mode=P.PAD_REFLECT),
                            P.conv_fprop_units(ar.off('down_sampling.1.weight'), c0, cin, 7, 7, 3), c0)
        self.n_stem = self.ns.make(dev, B, H * W, [('down_sampling.2', c0)], tr)
        self.y1, self.a1 = self._act(H2, W2, c1), self._act(H2, W2, c1)
        self.g_d1 = G(P.Geometry(B, H, W, cpad(c0), 0, H2, W2, cpad(c1), 0, sn=2),
                      P.conv_fprop_units(ar.off('down_sampling.4.weight'), c1, c0, 3, 3, 1), c1)
        self.n_d1 = self.ns.make(dev, B, H2 * W2, [('down_sampling.5', c1)], tr)
        self.y2, self.a2 = self._act(H4, W4, c2), self._act(H4, W4, c2)
        self.g_d2 = G(P.Geometry(B, H2, W2, cpad(c1), 0, H4, W4, cpad(c2), 0, sn=2),
                      P.conv_fprop_units(ar.off('down_sampling.7.weight'), c2, c1, 3, 3, 1), c2)
        self.n_d2 = self.ns.make(dev, B, H4 * W4, [('down_sampling.8', c2)], tr)

        # ---- residual blocks
        C, Cp = c2, cpad(c2)
        self.blocks = []
        x = self.a2
        maxL = maxLpad = 0
        for i, blk in enumerate(A['blocks']):
            pre = f'features.{i}'
            b = _Block()
            b.x = x
            b.res = [(j, m, k) for j, (m, k) in enumerate((mk for mk in zip(blk['res'], ks) if mk[0] > 0))]
            b.dw = [(j, m, k) for j, (m, k) in enumerate((mk for mk in zip(blk['dw'], ks) if mk[0] > 0))]
            b.empty = not b.res and not b.dw
            if b.empty:  # forward returns x unchanged (inception_modules.py:231-232)
                b.out = x
                self.blocks.append(b)
                continue
            # channel slices of the mid buffer: [1x1 first-stage convs (res k=1, dw) | res k>1 | dw second-stage]
            off = 0
            b.res_sl, b.dw1_sl, b.dw2_sl = [None] * len(b.res), [None] * len(b.dw), []
            b.order = _stage1_order(b.res, b.dw)
            b.D0 = b.D1 = 0   # the dw first-stage slices are contiguous: [D0, D1)
            for (kind, j, m, _k) in b.order:
                if kind == 'res':
                    b.res_sl[j] = off
                else:
                    if j == 0:
                        b.D0 = off
                    b.dw1_sl[j] = off
                    b.D1 = off + cpad(m)
                off += cpad(m)
            b.LA = off
            for _, m, _k in b.dw:
                b.dw2_sl.append(off)
                off += cpad(m)
            b.L = off
            maxL = max(maxL, b.L)
            b.mid_raw = self._act(H4, W4, b.L, zero=True)
            b.mid_act = self._act(H4, W4, b.L, zero=True)
            b.tmp, b.out = self._act(H4, W4, C), self._act(H4, W4, C)
            # stage 1: the first convs of all branches read x.  Forward: the 1x1 convs (k=1 res branch + the
            # first conv of every dw branch) are N-concatenated into ONE GEMM (rows packed from each weight
            # tensor, catb_pack_weights_rows); each k>1 res branch is its own GEMM.  The per-branch objects
            # are kept for the weight gradients.
            b.s1, b.s1_fwd = [], []
            ones = [(kind, j, m) for (kind, j, m, k) in b.order if k == 1]
            # pruned students (all first-stage slices within one 128-row N tile): every first conv in ONE GEMM on the tap
            # grid of the largest kernel -- the block is launch / latency bound, and an N = 128 MMA costs ~1.5x an N = 24 one
            fuse_all = len(b.order) > 1 and b.LA <= 128 and os.environ.get('CATB_NO_S1FUSE', '0') != '1'
            for (kind, j, m, k) in b.order:
                if kind == 'res':
                    wn, sl = f'{pre}.res_ops.{j}.1.0.weight', b.res_sl[j]
                else:
                    wn, sl = f'{pre}.dw_ops.{j}.0.0.weight', b.dw1_sl[j]
                fused = fuse_all or (k == 1 and len(ones) > 1)
                g = Gemm(P.Geometry(B, H4, W4, Cp, 0, H4, W4, b.L, sl, pad_mode=P.PAD_REFLECT),
                         P.conv_fprop_units(ar.off(wn), m, C, k, k, (k - 1) // 2), m, dev, need_pack=not fused)
                b.s1.append((g, sl, m, k, wn))
                if not fused:
                    b.s1_fwd.append(g)
                    self.fprop_gemms.append(g)
            if fuse_all:
                kmax = max(k for (_g, _sl, _m, k, _wn) in b.s1)
                segs = [(sl, cpad(m), m, P.conv_embedded_units(ar.off(wn), m, C, k, kmax)) for (_g, sl, m, k, wn) in b.s1]
                g = Gemm(P.Geometry(B, H4, W4, Cp, 0, H4, W4, b.L, 0, pad_mode=P.PAD_REFLECT),
                         P.conv_fprop_units(0, b.LA, C, kmax, kmax, (kmax - 1) // 2), b.LA, dev, segments=segs)
                b.s1_fwd = [g]
                self.fprop_gemms.append(g)
            elif len(ones) > 1:
                rows = sum(cpad(m) for (_, _, m) in ones)      # the 1x1 slices start at channel 0 and are contiguous
                base = P.conv_fprop_units(0, rows, C, 1, 1, 0)
                segs = [(sl, cpad(m), m, P.conv_fprop_units(ar.off(wn), m, C, 1, 1, 0)) for (_, sl, m, k, wn) in b.s1 if k == 1]
                g = Gemm(P.Geometry(B, H4, W4, Cp, 0, H4, W4, b.L, 0), base, rows, dev, segments=segs)
                b.s1_fwd.insert(0, g)
                self.fprop_gemms.append(g)
            # weight gradients of stage 1: all first convs read x, so when their output slices fit ONE 128-row MMA tile
            # (pruned students: 6 branches x <= 24 channels) they are a single GEMM -- rows = the slices of the mid-buffer
            # gradient, K = the tap grid of the largest kernel over x -- at the cost of the largest kernel's gradient
            # alone (the M side of the MMA is free up to 128 rows); the 1x1 / 3x3 branches simply ignore the taps outside
            # their kernel when the workspace is added to the arena (catb_wgrad_unpack per row segment).
            b.s1w = None
            if ng and len(b.s1) > 1 and b.LA <= 128 and os.environ.get('CATB_NO_WFUSE', '0') != '1':
                kmax = max(k for (_g, _sl, _m, k, _wn) in b.s1)
                segs = [(sl, cpad(m), m, P.conv_embedded_units(ar.off(wn), m, C, k, kmax)) for (_g, sl, m, k, wn) in b.s1]
                b.s1w = Gemm(P.Geometry(B, H4, W4, Cp, 0, H4, W4, b.L, 0, pad_mode=P.PAD_REFLECT),
                             P.conv_fprop_units(0, b.LA, C, kmax, kmax, (kmax - 1) // 2), b.LA, dev, need_pack=False,
                             segments=segs)
            grpA = [(f'{pre}.res_ops.{j}.1.1' if kind == 'res' else f'{pre}.dw_ops.{j}.0.1', m) for (kind, j, m, _k) in b.order]
            b.nA = self.ns.make(dev, B, H4 * W4, grpA, tr)
            # depthwise convs over the dw slices
            if b.dw:
                Cdw = b.L - b.LA
                ksz = torch.ones(Cdw, dtype=torch.int32)
                wof = torch.full((Cdw,), -1, dtype=torch.int32)
                for (j, m, k), sl in zip(b.dw, b.dw2_sl):
                    o = sl - b.LA
                    ksz[o:o + cpad(m)] = k
                    wof[o:o + m] = ar.off(f'{pre}.dw_ops.{j}.2.0.weight') + torch.arange(m, dtype=torch.int32) * k * k
                b.dw_k, b.dw_w = ksz.to(dev), wof.to(dev)
                b.nB = self.ns.make(dev, B, H4 * W4, [(f'{pre}.dw_ops.{j}.2.1', m) for j, m, k in b.dw], tr)
            # stage 2: ONE GEMM, K-concatenation of every branch's last conv
            u2 = P.Units()
            for (j, m, k), sl in zip(b.res, b.res_sl):
                u2.extend(P.conv_fprop_units(ar.off(f'{pre}.res_ops.{j}.4.weight'), C, m, k, k, (k - 1) // 2, cu0=sl // 8))
            for (j, m, k), sl in zip(b.dw, b.dw2_sl):
                u2.extend(P.conv_fprop_units(ar.off(f'{pre}.dw_ops.{j}.4.weight'), C, m, 1, 1, 0, cu0=sl // 8))
            b.g2 = G(P.Geometry(B, H4, W4, b.L, 0, H4, W4, Cp, 0, pad_mode=P.PAD_REFLECT), u2, C)
            b.npw = self.ns.make(dev, B, H4 * W4, [(f'{pre}.pw_bn', C)], tr)
            b.d2f = None
            if ng and (len(b.res) + len(b.dw)) > 1 and b.L <= 256 and os.environ.get('CATB_NO_D2FUSE', '0') != '1':
                # stage-2 input gradients of all branches as ONE N-concatenated GEMM into the frame of the largest kernel
                # (rows = the branches' slices of d(mid); the dw first-stage slices in between get zero rows and are
                # written by the depthwise input gradient afterwards), folded once
                k2 = max([k for _, _, k in b.res] + [1])
                b.P2 = (k2 - 1) // 2
                segs = [(sl, cpad(m), m, P.conv_dgrad_embedded_units(ar.off(f'{pre}.res_ops.{j}.4.weight'), C, m, k, k2))
                        for (j, m, k), sl in zip(b.res, b.res_sl)]
                segs += [(sl, cpad(m), m, P.conv_dgrad_embedded_units(ar.off(f'{pre}.dw_ops.{j}.4.weight'), C, m, 1, k2))
                         for (j, m, k), sl in zip(b.dw, b.dw2_sl)]
                Hp, Wp = H4 + 2 * b.P2, W4 + 2 * b.P2
                if b.P2 > 0:
                    maxLpad = max(maxLpad, Hp * Wp * b.L)
                b.d2f = Gemm(P.Geometry(B, H4, W4, Cp, 0, Hp, Wp, b.L, 0), P.conv_dgrad_units(0, C, b.L, k2, k2, 0), b.L, dev,
                             segments=segs)
                self.bwd_gemms.append(b.d2f)
                b.d2 = []
            elif ng:
                # stage-2 input gradients: one GEMM per branch (padded frame + fold for k > 1)
                b.d2 = []
                for (j, m, k), sl in zip(b.res, b.res_sl):
                    p = (k - 1) // 2
                    un = P.conv_dgrad_units(ar.off(f'{pre}.res_ops.{j}.4.weight'), C, m, k, k, 0)
                    if p > 0:
                        maxLpad = max(maxLpad, (H4 + 2 * p) * (W4 + 2 * p) * cpad(m))
                        g = G(P.Geometry(B, H4, W4, Cp, 0, H4 + 2 * p, W4 + 2 * p, cpad(m), 0), un, m, bwd=True)
                    else:
                        g = G(P.Geometry(B, H4, W4, Cp, 0, H4, W4, b.L, sl), un, m, bwd=True)
                    b.d2.append((g, sl, m, p))
                for (j, m, k), sl in zip(b.dw, b.dw2_sl):
                    un = P.conv_dgrad_units(ar.off(f'{pre}.dw_ops.{j}.4.weight'), C, m, 1, 1, 0)
                    b.d2.append((G(P.Geometry(B, H4, W4, Cp, 0, H4, W4, b.L, sl), un, m, bwd=True), sl, m, 0))
            if ng:
                # stage-1 input gradient: ONE GEMM over the concatenated first-stage gradients
                b.P1 = max([(k - 1) // 2 for _, _, k in b.res] + [0])
                u1 = P.Units()
                for (j, m, k), sl in zip(b.res, b.res_sl):
                    q = (k - 1) // 2 - b.P1
                    u1.extend(P.conv_dgrad_units(ar.off(f'{pre}.res_ops.{j}.1.0.weight'), m, C, k, k, q, cu0=sl // 8))
                for (j, m, k), sl in zip(b.dw, b.dw1_sl):
                    u1.extend(P.conv_dgrad_units(ar.off(f'{pre}.dw_ops.{j}.0.0.weight'), m, C, 1, 1, -b.P1, cu0=sl // 8))
                Hp, Wp = H4 + 2 * b.P1, W4 + 2 * b.P1
                b.g1d = G(P.Geometry(B, H4, W4, b.L, 0, Hp, Wp, Cp, 0), u1, C, bwd=True)
            self.blocks.append(b)
            x = b.out
        self.feat_out = x

        # ---- up-sampling (ConvTranspose 3x3 s2 p1 op1 as four sub-pixel phases) and head
        def up(prefix_conv, prefix_norm, xin_C, out_C, h, w):
            y, a = self._act(2 * h, 2 * w, out_C), self._act(2 * h, 2 * w, out_C)
            fu = P.convT_fprop_units(ar.off(prefix_conv + '.weight'), xin_C, out_C, 3, 3, 1)
            gs = []
            for pa in range(2):
                for pb in range(2):
                    gs.append(G(P.Geometry(B, h, w, cpad(xin_C), 0, 2 * h, 2 * w, cpad(out_C), 0, sn=1, sd=2, o_step=2,
                                           o_ph=pa, o_pw=pb), fu.phase(pa, pb), out_C))
            n = self.ns.make(dev, B, 4 * h * w, [(prefix_norm, out_C)], tr)
            bu = P.convT_dgrad_units(ar.off(prefix_conv + '.weight'), xin_C, out_C, 3, 3, 1)
            gb = G(P.Geometry(B, 2 * h, 2 * w, cpad(out_C), 0, h, w, cpad(xin_C), 0, sn=2, sd=1), bu, xin_C, bwd=True) if ng else None
            return y, a, gs, n, gb

        self.yu1, self.au1, self.g_u1, self.n_u1, self.gb_u1 = up('up_sampling.0', 'up_sampling.1', c2, c3, H4, W4)
        self.yu2, self.au2, self.g_u2, self.n_u2, self.gb_u2 = up('up_sampling.3', 'up_sampling.4', c3, c4, H2, W2)
        self.out = self._act(H, W, cout)
        if self.packx:
            self.head_P = Act.empty(B, H, W + 6, cout * 8, dev)
            self.g_head = Gemm(P.Geometry(B, H, W, cpad(c4), 0, H, W + 6, cout * 8, 0, pad_mode=P.PAD_REFLECT),
                               P.conv_fprop_units(self.aux.off('head.wx'), cout * 8, c4, 7, 1, 3, pad_s=3), cout * 8, dev)
            self.aux_gemms.append(self.g_head)
        else:
            self.g_head = G(P.Geometry(B, H, W, cpad(c4), 0, H, W, cpad(cout), 0, pad_mode=P.PAD_REFLECT),
                            P.conv_fprop_units(ar.off('up_sampling.7.weight'), cout, c4, 7, 7, 3), cout)
        self.head_bias = ar.view('up_sampling.7.bias')

        self.acts = {'down_sampling.9': self.a2}
        for i in (2, 5, 8):
            self.acts[f'features.{i}'] = self.blocks[i].out

        if ng:
            f = dict(dtype=ops.BF16, device=dev)
            # gradient workspaces shared by all blocks
            self.ws_dmid_act = torch.zeros(B * H4 * W4 * max(maxL, 8), **f)
            self.ws_dmid_raw = torch.zeros(B * H4 * W4 * max(maxL, 8), **f)
            self.ws_frame = torch.zeros(B * max(maxLpad, 8), **f)
            self.ws_dtmp = self._act(H4, W4, C)
            self.ws_dxp = torch.zeros(B * (H4 + 6) * (W4 + 6) * Cp, **f)
            self.dfeat = [self._act(H4, W4, C), self._act(H4, W4, C)]   # ping-pong d(block output)
            # head / up / down gradient buffers
            self.d_head_z = self._act(H, W, cout)
            self.d_head_frame = Act.empty(B, H + 6, W + 6, c4, dev)
            if self.packx:
                self.d_head_P = Act.empty(B, H, W + 6, cout * 8, dev, zero=True)
                self.gb_head = Gemm(P.Geometry(B, H, W + 6, cout * 8, 0, H + 6, W + 6, cpad(c4), 0),
                                    P.conv_dgrad_units(self.aux.off('head.wx'), cout * 8, c4, 7, 1, 0, q_s=0), c4, dev)
                self.aux_gemms.append(self.gb_head)
            else:
                self.gb_head = G(P.Geometry(B, H, W, cpad(cout), 0, H + 6, W + 6, cpad(c4), 0),
                                 P.conv_dgrad_units(ar.off('up_sampling.7.weight'), cout, c4, 7, 7, 0), c4, bwd=True)
            self.d_au2, self.d_yu2 = self._act(H, W, c4), self._act(H, W, c4)
            self.d_au1, self.d_yu1 = self._act(H2, W2, c3), self._act(H2, W2, c3)
            self.d_y2, self.d_a1, self.d_y1, self.d_a0, self.d_y0 = (self._act(H4, W4, c2), self._act(H2, W2, c1),
                                                                       self._act(H2, W2, c1), self._act(H, W, c0),
                                                                       self._act(H, W, c0))
            self.gb_d2 = _strided_dgrad(dict(N=B, H=H4, W=W4, ldx=cpad(c2), x_coff=0, OH=H2, OW=W2, ldy=cpad(c1), y_coff=0),
                                        P.conv_dgrad_units(ar.off('down_sampling.7.weight'), c2, c1, 3, 3, 1), c1, dev, 2)
            self.gb_d1 = _strided_dgrad(dict(N=B, H=H2, W=W2, ldx=cpad(c1), x_coff=0, OH=H, OW=W, ldy=cpad(c0), y_coff=0),
                                        P.conv_dgrad_units(ar.off('down_sampling.4.weight'), c1, c0, 3, 3, 1), c0, dev, 2)
            self.bwd_gemms += self.gb_d2 + self.gb_d1
            if self.input_grad:
                # gradient w.r.t. the reflection-padded input frame of the 7x7 stem (always from the 7x7 parameter tensor,
                # also when the forward pass uses the x-packed image), folded by catb_reflect_fold
                self.d_in_frame = Act.empty(B, H + 6, W + 6, cin, dev)
                self.d_in = self._act(H, W, cin)
                # (gather-per-tap kernel only: a 49-tap frame gradient with 3 output channels has not been timed or run on
                # the halo kernel yet; this GEMM is a small part of the CycleGAN step)
                self.gb_stem = Gemm(P.Geometry(B, H, W, cpad(c0), 0, H + 6, W + 6, cpad(cin), 0),
                                    P.conv_dgrad_units(ar.off('down_sampling.1.weight'), c0, cin, 7, 7, 0), cin, dev, halo=False)
                self.bwd_gemms.append(self.gb_stem)

    def _ws(self, flat, H, W, C):
        return Act(flat[:self.B * H * W * C].view(self.B, H, W, C))

    def pack_weights(self):
        """Re-pack every GEMM weight image from the parameters: one launch (+ one for the x-packed stem / head)."""
        if getattr(self, '_packer', None) is None:
            self._packer = ops.PackBatch(self.fprop_gemms + self.bwd_gemms, self.dev)
            self._packer_aux = ops.PackBatch(self.aux_gemms, self.dev)
        self._packer.run(self.arena.p)
        if self.aux_gemms:
            ops.gather_sum(self.arena.p, self.aux_idx, self.aux.p)
            self._packer_aux.run(self.aux.p)

    # ---- forward -------------------------------------------------------------------------------
    def forward(self, x_in: Act):
        """x_in: NHWC bf16 input image [B,H,W,cpad(input_nc)].  Returns the output Act (tanh applied)."""
        self.x_in = x_in
        relu, none = ACT['relu'], ACT['none']
        self.pool_sums.zero_()
        if self.packx:
            ops.expand_x(x_in, self.x_stem, self.arch['input_nc'], 7)
            conv_norm(self.g_stem, self.x_stem.t, self.y0, self.n_stem, self.y0, self.a0, relu)
        else:
            conv_norm(self.g_stem, x_in.t, self.y0, self.n_stem, self.y0, self.a0, relu)
        conv_norm(self.g_d1, self.a0.t, self.y1, self.n_d1, self.y1, self.a1, relu)
        conv_norm(self.g_d2, self.a1.t, self.y2, self.n_d2, self.y2, self.a2, relu)
        for b in self.blocks:
            if b.empty:
                continue
            conv_norm(b.s1_fwd, b.x.t, b.mid_raw, b.nA, b.mid_raw.slice(0, b.LA), b.mid_act.slice(0, b.LA), relu)
            if b.dw:
                ops.dwconv_fwd(b.mid_act.slice(b.D0, b.D1 - b.D0), b.mid_raw.slice(b.LA, b.L - b.LA), b.dw_k, b.dw_w,
                               self.arena.p)
                b.nB.forward(b.mid_raw.slice(b.LA, b.L - b.LA), b.mid_act.slice(b.LA, b.L - b.LA), relu)
            conv_norm(b.g2, b.mid_act.t, b.tmp, b.npw, b.tmp, b.out, none, residual=b.x)
        conv_norm(self.g_u1, self.feat_out.t, self.yu1, self.n_u1, self.yu1, self.au1, relu)
        conv_norm(self.g_u2, self.au1.t, self.yu2, self.n_u2, self.yu2, self.au2, relu)
        if self.packx:
            self.g_head.fprop(self.au2.t, self.head_P.t)
            ops.shift_sum(self.head_P, self.out, self.arch['output_nc'], 7, self.head_bias, ACT['tanh'])
        else:
            self.g_head.fprop(self.au2.t, self.out.t, bias=self.head_bias, act=ACT['tanh'])
        return self.out

    # ---- backward ------------------------------------------------------------------------------
    def backward(self, d_out: Act, act_grads=None):
        """See _backward; the second stages of all weight gradients run as one launch when the pass is left."""
        if getattr(self, '_unpack', None) is None:
            self._unpack = ops.UnpackQueue(self.dev)
        with self._unpack:
            return self._backward(d_out, act_grads)

    def _backward(self, d_out: Act, act_grads=None):
        """d_out: gradient w.r.t. the tanh output.  act_grads: {mapping layer: callable(Act)} invoked to
        accumulate extra gradient (the KA loss) into d(activation) at the four mapping layers."""
        assert self.need_grad
        relu, none, ar = ACT['relu'], ACT['none'], self.arena
        self.pool_red.zero_()
        act_grads = act_grads or {}
        B, H, W = self.B, self.H, self.W
        c0, c1, c2, c3, c4 = self.arch['widths']
        H4, W4 = H // 4, W // 4
        Cp = cpad(c2)
        # head: tanh, 7x7 reflect conv
        ops.act_bwd(d_out, self.out, self.d_head_z, ACT['tanh'])
        ops.channel_sum(self.d_head_z, ar.view('up_sampling.7.bias', 'g'))
        if self.packx:
            self.aux.g.zero_()
            ops.shift_expand(self.d_head_z, self.d_head_P, self.arch['output_nc'], 7)
            self.g_head.wgrad(self.au2.t, self.d_head_P.t, self.aux.g)
            self.gb_head.fprop(self.d_head_P.t, self.d_head_frame.t)
        else:
            self.g_head.wgrad(self.au2.t, self.d_head_z.t, ar.g)
            self.gb_head.fprop(self.d_head_z.t, self.d_head_frame.t)
        ops.reflect_fold(self.d_head_frame, self.d_au2, 3)
        # up 2
        self.n_u2.backward(self.d_au2, self.au2, self.yu2, self.d_yu2, relu)
        self.gb_u2.wgrad(self.d_yu2.t, self.au1.t, ar.g)
        self.gb_u2.fprop(self.d_yu2.t, self.d_au1.t)
        # up 1
        self.n_u1.backward(self.d_au1, self.au1, self.yu1, self.d_yu1, relu)
        self.gb_u1.wgrad(self.d_yu1.t, self.feat_out.t, ar.g)
        cur = self.dfeat[0]
        self.gb_u1.fprop(self.d_yu1.t, cur.t)
        nxt_i = 1
        # Weight gradients only feed the optimiser, so inside the residual blocks they are issued on a side stream
        # (parallel branches of the captured graph) while the main stream carries the input-gradient chain.  The side
        # stream reads two workspaces shared by all blocks: `ws_dtmp` (stage-2 weight gradient) and `dmid_raw`
        # (depthwise / stage-1 weight gradients); the main stream waits for those readers right before the next block
        # overwrites the respective buffer (ev_tmp / ev_raw), and joins the side stream after the last block.
        ov = self.overlap_wgrad
        main = torch.cuda.current_stream() if ov else None
        if ov and self._wside is None:
            self._wside = torch.cuda.Stream(device=self.dev)
        side = self._wside
        ev_tmp = ev_raw = None

        def on_side(fn):
            if not ov:
                return fn()
            side.wait_stream(main)
            with torch.cuda.stream(side):
                fn()

        def mark():
            if not ov:
                return None
            ev = torch.cuda.Event()
            ev.record(side)
            return ev

        for i in range(len(self.blocks) - 1, -1, -1):
            b = self.blocks[i]
            name = f'features.{i}'
            if name in act_grads:
                act_grads[name](cur)
            if b.empty:
                continue
            dmid_act = self._ws(self.ws_dmid_act, H4, W4, b.L)
            dmid_raw = self._ws(self.ws_dmid_raw, H4, W4, b.L)
            # x + pw_bn(tmp): norm backward without activation
            if ev_tmp is not None:
                main.wait_event(ev_tmp)          # the previous block's stage-2 weight gradient has read ws_dtmp
            b.npw.backward(cur, None, b.tmp, self.ws_dtmp, none)
            on_side(lambda: b.g2.wgrad(b.mid_act.t, self.ws_dtmp.t, ar.g))
            ev_tmp = mark()
            if b.d2f is not None:
                if b.P2 > 0:
                    fr = self._ws(self.ws_frame, H4 + 2 * b.P2, W4 + 2 * b.P2, b.L)
                    b.d2f.fprop(self.ws_dtmp.t, fr.t)
                    ops.reflect_fold(fr, dmid_act, b.P2)
                else:
                    b.d2f.fprop(self.ws_dtmp.t, dmid_act.t)
            for (g, sl, m, p) in b.d2:
                if p > 0:
                    fr = self._ws(self.ws_frame, H4 + 2 * p, W4 + 2 * p, cpad(m))
                    g.fprop(self.ws_dtmp.t, fr.t)
                    ops.reflect_fold(fr, dmid_act.slice(sl, cpad(m)), p)
                else:
                    g.fprop(self.ws_dtmp.t, dmid_act.t)
            if ev_raw is not None:
                main.wait_event(ev_raw)          # the previous block's depthwise / stage-1 weight gradients have read dmid_raw
            if b.dw:
                nB = b.L - b.LA
                b.nB.backward(dmid_act.slice(b.LA, nB), b.mid_act.slice(b.LA, nB), b.mid_raw.slice(b.LA, nB),
                              dmid_raw.slice(b.LA, nB), relu)
                on_side(lambda: ops.dwconv_bwd_weight(b.mid_act.slice(b.D0, b.D1 - b.D0), dmid_raw.slice(b.LA, nB), b.dw_k,
                                                      b.dw_w, ar.g))
                ops.dwconv_bwd_data(dmid_raw.slice(b.LA, nB), dmid_act.slice(b.D0, b.D1 - b.D0), b.dw_k, b.dw_w, ar.p)
            b.nA.backward(dmid_act.slice(0, b.LA), b.mid_act.slice(0, b.LA), b.mid_raw.slice(0, b.LA),
                          dmid_raw.slice(0, b.LA), relu)

            def s1_wgrads(b=b, dmid_raw=dmid_raw):
                if b.s1w is not None:
                    b.s1w.wgrad(b.x.t, dmid_raw.t, ar.g)
                    return
                for (g, sl, m, k, wn) in b.s1:
                    g.wgrad(b.x.t, dmid_raw.t, ar.g)
            on_side(s1_wgrads)
            ev_raw = mark()
            nxt = self.dfeat[nxt_i]
            if b.P1 > 0:
                fr = self._ws(self.ws_dxp, H4 + 2 * b.P1, W4 + 2 * b.P1, Cp)
                b.g1d.fprop(dmid_raw.t, fr.t)
                ops.reflect_fold(fr, nxt, b.P1, add=cur)
            else:
                b.g1d.fprop(dmid_raw.t, nxt.t)
                ops.add(nxt, cur, nxt)
            cur, nxt_i = nxt, 1 - nxt_i
        if ov:
            main.wait_stream(side)               # join: every block weight gradient is in the arena
        if 'down_sampling.9' in act_grads:
            act_grads['down_sampling.9'](cur)
        # down 2, down 1, stem (no input gradient needed for the image)
        self.n_d2.backward(cur, self.a2, self.y2, self.d_y2, relu)
        self.g_d2.wgrad(self.a1.t, self.d_y2.t, ar.g)
        for g in self.gb_d2:
            g.fprop(self.d_y2.t, self.d_a1.t)
        self.n_d1.backward(self.d_a1, self.a1, self.y1, self.d_y1, relu)
        self.g_d1.wgrad(self.a0.t, self.d_y1.t, ar.g)
        for g in self.gb_d1:
            g.fprop(self.d_y1.t, self.d_a0.t)
        self.n_stem.backward(self.d_a0, self.a0, self.y0, self.d_y0, relu)
        if self.packx:
            self.g_stem.wgrad(self.x_stem.t, self.d_y0.t, self.aux.g)
            ops.flush_unpack()                                  # the two aux weight gradients must have landed in aux.g
            ops.scatter_add(self.aux.g, self.aux_idx, ar.g)     # d(stem.wx), d(head.wx) -> the 7x7 parameter gradients
        else:
            self.g_stem.wgrad(self.x_in.t, self.d_y0.t, ar.g)
        if self.input_grad:
            self.gb_stem.fprop(self.d_y0.t, self.d_in_frame.t)
            ops.reflect_fold(self.d_in_frame, self.d_in, 3)
            return self.d_in


# ------------------------------------------------------------------------------------------------
# discriminator
# ------------------------------------------------------------------------------------------------
def discriminator_layers(arch):
    """(seq index of the conv, cin, cout, stride, has_norm, has_act) -- discriminators.py:37-75."""
    ndf, n_layers = arch['ndf'], arch['n_layers']
    layers = [(0, arch['input_nc'], ndf, 2, False, True)]
    idx, mult = 2, 1
    for n in range(1, n_layers):
        prev, mult = mult, min(2 ** n, 8)
        layers.append((idx, ndf * prev, ndf * mult, 2, True, True))
        idx += 3
    prev, mult = mult, min(2 ** n_layers, 8)
    layers.append((idx, ndf * prev, ndf * mult, 1, True, True))
    idx += 3
    layers.append((idx, ndf * mult, 1, 1, False, False))
    return layers


class _DLayer:
    pass


class DisNet:
    """NLayerDiscriminator (70x70 PatchGAN) compiled for a fixed (B, H, W); always in train mode on the
    distillation path (netD is never put in eval(), base_inception_distiller.py:144-169)."""

    def __init__(self, arch, B, H, W, device, layers=None, pad=1, names=None, arenas=None):
        """layers / pad / names / arenas generalise the class to the sub-discriminators of the SPADE
        MultiscaleDiscriminator (discriminators.py:129-180; cat_b200/spade_engine.py): `names(ci)` returns the
        (weight, bias, norm-prefix) keys of conv ci, `arenas=(arena, bufs)` are shared, not yet finalised arenas
        owned by the caller, who then calls build() after finalising them."""
        self.arch, self.B, self.H, self.W, self.dev = arch, B, H, W, device
        self.pad = pad
        self.names = names or (lambda ci: (f'model.{ci}.weight', f'model.{ci}.bias', f'model.{ci + 1}'))
        own = arenas is None
        self.arena, self.bufs = (Arena(True), Arena(False)) if own else arenas
        self.ns = _NormSpec(self.arena, self.bufs, arch)
        self.layers = []
        h, w = H, W
        for (ci, cin, cout, stride, has_norm, has_act) in (layers or discriminator_layers(arch)):
            L = _DLayer()
            L.ci, L.cin, L.cout, L.stride, L.has_norm, L.has_act = ci, cin, cout, stride, has_norm, has_act
            L.h, L.w = h, w
            L.oh = (h + 2 * pad - 4) // stride + 1
            L.ow = (w + 2 * pad - 4) // stride + 1
            L.has_bias = not has_norm  # biases in front of a norm layer are mathematically inert
            L.wn, L.bn, L.nn = self.names(ci)
            self.arena.alloc(L.wn, (cout, cin, 4, 4))
            if not has_norm or arch['use_bias']:
                self.arena.alloc(L.bn, (cout,))
            if has_norm:
                self.ns.alloc_group([(L.nn, cout)])
            h, w = L.oh, L.ow
            self.layers.append(L)
        self.w_src = None       # effective weights the GEMM images are packed from (default: the parameters)
        if own:
            self.arena.finalize(device)
            self.bufs.finalize(device)
            self.build()

    def build(self):
        arch, B, H, W, device, pad = self.arch, self.B, self.H, self.W, self.dev, self.pad
        ar, dev = self.arena, device
        self.fprop_gemms, self.bwd_gemms = [], []
        self.d_in = Act.empty(B, H, W, arch['input_nc'], dev)  # gradient w.r.t. the input image
        prev_C = arch['input_nc']
        for li, L in enumerate(self.layers):
            last = li == len(self.layers) - 1
            wn = L.wn
            L.units = P.conv_fprop_units(ar.off(wn), L.cout, L.cin, 4, 4, pad)
            L.y_f32 = last
            L.tap = bool(last and L.cout == 1 and L.stride == 1 and ops.TAP_HEAD)
            if L.tap:
                # one output channel: the conv in "tap-split" form (csrc/packx.cu) -- P = X . W[taps] is a 1x1 GEMM with the
                # 16 taps as output channels (each input pixel read once), the prediction is the shifted tap sum; the
                # adjoint spreads d(prediction) over the taps and both gradients are 1x1 GEMMs too
                L.y = torch.zeros(B, L.oh, L.ow, 8, dtype=torch.float32, device=dev)
                L.P = torch.zeros(B, L.h, L.w, 16, dtype=torch.float32, device=dev)
                L.dP = Act.empty(B, L.h, L.w, 16, dev, zero=True)
                L.units = P.tap_split_units(ar.off(wn), L.cin, 4, 4)
                L.g = Gemm(P.Geometry(B, L.h, L.w, cpad(L.cin), 0, L.h, L.w, 16, 0), L.units, 16, dev)
                self.fprop_gemms.append(L.g)
                L.bias = ar.view(L.bn) if L.has_bias else None
                L.dbias = ar.view(L.bn, 'g') if L.has_bias else None
                L.dy, L.da = None, None
                L.gb = [Gemm(P.Geometry(B, L.h, L.w, 16, 0, L.h, L.w, cpad(L.cin), 0),
                             P.tap_split_dgrad_units(ar.off(wn), L.cin, 4, 4), L.cin, dev)]
                self.bwd_gemms += L.gb
                L.gw = Gemm(P.Geometry(B, L.h, L.w, cpad(L.cin), 0, L.h, L.w, 16, 0), L.units, 16, dev, need_pack=False)
                continue
            if last:
                L.y = torch.zeros(B, L.oh, L.ow, 8, dtype=torch.float32, device=dev)
                ldy = 8
            else:
                L.yraw = Act.empty(B, L.oh, L.ow, L.cout, dev)
                L.a = Act.empty(B, L.oh, L.ow, L.cout, dev) if L.has_norm else L.yraw
                ldy = cpad(L.cout)
            L.g = Gemm(P.Geometry(B, L.h, L.w, cpad(L.cin), 0, L.oh, L.ow, ldy, 0, sn=L.stride), L.units, L.cout, dev)
            self.fprop_gemms.append(L.g)
            L.bias = ar.view(L.bn) if L.has_bias else None
            L.dbias = ar.view(L.bn, 'g') if L.has_bias else None
            if L.has_norm:
                L.norm = self.ns.make(dev, B, L.oh * L.ow, [(L.nn, L.cout)], True)
            # backward buffers
            L.dy = Act.empty(B, L.oh, L.ow, L.cout, dev)       # gradient w.r.t. the conv output
            L.da = Act.empty(B, L.oh, L.ow, L.cout, dev) if not last else None  # w.r.t. the activation
            geo_kw = dict(N=B, H=L.oh, W=L.ow, ldx=cpad(L.cout), x_coff=0, OH=L.h, OW=L.w, ldy=cpad(L.cin), y_coff=0)
            L.gb = _strided_dgrad(geo_kw, P.conv_dgrad_units(ar.off(wn), L.cout, L.cin, 4, 4, pad), L.cin, dev, L.stride)
            self.bwd_gemms += L.gb
            # wgrad geometry: lattice tensor = dy (bf16, pitch cpad(cout))
            L.gw = Gemm(P.Geometry(B, L.h, L.w, cpad(L.cin), 0, L.oh, L.ow, cpad(L.cout), 0, sn=L.stride), L.units,
                        L.cout, dev, need_pack=False)
        self.pool_sums, self.pool_red = pool_norm_buffers(self.ns.created, dev)
        self.pred = self.layers[-1].y
        self.pred_n = B * self.layers[-1].oh * self.layers[-1].ow

    def load_state_dict(self, sd):
        self.arena.load_state_dict(sd)
        self.bufs.load_state_dict(sd)
        self.pack_weights()

    def state_dict(self):
        sd = self.arena.state_dict()
        sd.update(self.bufs.state_dict())
        return sd

    def pack_weights(self):
        src = self.arena.p if self.w_src is None else self.w_src
        if getattr(self, '_packer', None) is None:
            self._packer = ops.PackBatch(self.fprop_gemms + self.bwd_gemms, self.dev)
        self._packer.run(src)

    def forward(self, x: Act):
        self.x = x
        cur = x
        self.pool_sums.zero_()
        for L in self.layers:
            if L.tap:
                L.g.fprop(cur.t, L.P, y_is_f32=True)
                ops.tap_sum(L.P, L.y, L.h, L.w, L.oh, L.ow, 4, 4, self.pad, L.bias)
            elif L.y_f32:
                L.g.fprop(cur.t, L.y, bias=L.bias, y_is_f32=True)
            elif L.has_norm:
                conv_norm(L.g, cur.t, L.yraw, L.norm, L.yraw, L.a, ACT['leaky'])
                cur = L.a
            else:
                L.g.fprop(cur.t, L.yraw.t, bias=L.bias, act=ACT['leaky'])
                cur = L.yraw
        return self.pred

    def backward(self, dpred: Act, param_grads, input_grad, act_grad_hook=None, grads_final_hook=None):
        """See _backward; the second stages of all weight gradients run as one launch when the pass is left.
        grads_final_hook(grad_slice): called right after a layer's parameter gradients are complete (its weight-gradient
        second stage flushed) with that layer's contiguous slice of the gradient arena -- the data-parallel step starts
        the slice's all-reduce there (cat_b200/parallel.py: LayerwiseReducer)."""
        if getattr(self, '_unpack', None) is None:
            self._unpack = ops.UnpackQueue(self.dev)
        with self._unpack:
            return self._backward(dpred, param_grads, input_grad, act_grad_hook, grads_final_hook)

    def layer_grad_slice(self, li):
        """Contiguous slice of the gradient arena holding every parameter of layer li (conv weight, bias, norm scale /
        shift: allocated back to back per layer)."""
        a = self.arena
        start = a.off(self.layers[li].wn)
        end = a.off(self.layers[li + 1].wn) if li + 1 < len(self.layers) else a.size
        return a.g[start:end]

    def _backward(self, dpred: Act, param_grads, input_grad, act_grad_hook=None, grads_final_hook=None):
        """dpred: bf16 [B,oh,ow,8] gradient of the loss w.r.t. the prediction (channel 0).
        act_grad_hook(li, d): called with the gradient w.r.t. the activation output of layer li before it is
        back-propagated (the feature-matching loss of the SPADE path adds its term there)."""
        ar = self.arena
        d = dpred
        self.pool_red.zero_()
        for li in range(len(self.layers) - 1, -1, -1):
            L = self.layers[li]
            x_in = self.layers[li - 1].a if li > 0 else self.x
            if act_grad_hook is not None and not L.y_f32:
                act_grad_hook(li, d)
            if L.y_f32:
                dy = d
            elif L.has_norm:
                L.norm.backward(d, L.a, L.yraw, L.dy, ACT['leaky'], param_grads=param_grads)
                dy = L.dy
            else:
                ops.act_bwd(d, L.yraw, L.dy, ACT['leaky'])
                dy = L.dy
            if L.tap:
                ops.tap_expand(d, L.dP, 4, 4, self.pad)
                dy_w = L.dP        # lattice tensor of the 1x1 weight / input gradients
            else:
                dy_w = dy
            if param_grads:
                L.gw.wgrad(x_in.t, dy_w.t, ar.g)
                if L.has_bias:
                    ops.channel_sum(dy, L.dbias)
                if grads_final_hook is not None:
                    ops.flush_unpack()
                    grads_final_hook(self.layer_grad_slice(li))
            if li > 0:
                tgt = self.layers[li - 1].da
                for g in L.gb:
                    g.fprop(dy_w.t, tgt.t)
                d = tgt
            elif input_grad:
                for g in L.gb:
                    g.fprop(dy_w.t, self.d_in.t)
        return self.d_in
