"""GPU parity tests of the individual libcatb200 kernels against CPU torch (fp64) on the same
bf16-rounded inputs.  Tolerances: the implicit GEMMs accumulate in fp32 from exactly representable
bf16 products, so they must match a fp64 evaluation to ~1e-5 relative before the final bf16 rounding
of the output (2^-9 relative, the only loss on a bf16 output)."""
import math

import pytest
import torch
import torch.nn.functional as F

from cat_b200 import igemm_plan as P

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]

DEV = 'cuda:0'


@pytest.fixture(scope='module', autouse=True)
def _init():
    from cat_b200 import ops
    ops.require_cuda()


def bf(x):
    return x.to(torch.bfloat16).to(torch.float64)


def to_dev_nhwc(x, ld=None, coff=0):
    """NCHW float (cpu) -> NHWC bf16 cuda tensor with pitch ld, slice offset coff."""
    N, C, H, W = x.shape
    ld = ld or P.cpad(C)
    out = torch.zeros(N, H, W, ld, dtype=torch.bfloat16)
    out[..., coff:coff + C] = x.permute(0, 2, 3, 1).to(torch.bfloat16)
    return out.to(DEV)


def from_dev_nhwc(t, C, coff=0):
    return t[..., coff:coff + C].permute(0, 3, 1, 2).to(torch.float64).cpu()


def arena_of(*tensors):
    flat = torch.cat([torch.zeros(5)] + [t.reshape(-1).float() for t in tensors])
    offs, o = [], 5
    for t in tensors:
        offs.append(o)
        o += t.numel()
    return flat.to(DEV), offs


def rel_err(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


CONV_CASES = [
    # k, stride, pad, mode, Cin, Cout, N, H, W
    (3, 1, 1, 'zero', 5, 7, 2, 10, 12),
    (5, 1, 2, 'reflect', 9, 4, 1, 16, 16),
    (7, 1, 3, 'reflect', 3, 17, 2, 20, 24),
    (1, 1, 0, 'zero', 24, 40, 2, 8, 8),
    (3, 2, 1, 'zero', 17, 31, 2, 16, 16),
    (4, 2, 1, 'zero', 64, 128, 2, 16, 16),      # K = 16 taps * 8 units = 16 chunks: ring wraps
    (4, 1, 1, 'zero', 128, 1, 2, 9, 9),          # PatchGAN head: one output channel
    (3, 1, 1, 'zero', 16, 300, 1, 12, 12),       # two N tiles
    (5, 1, 2, 'reflect', 256, 48, 1, 16, 16),    # teacher-like wide K
]


@pytest.mark.parametrize('k,stride,pad,mode,Cin,Cout,N,H,W', CONV_CASES)
def test_igemm_conv_fprop_dgrad_wgrad(k, stride, pad, mode, Cin, Cout, N, H, W):
    from cat_b200 import ops
    torch.manual_seed(k * 100 + Cin)
    x = torch.randn(N, Cin, H, W)
    w = torch.randn(Cout, Cin, k, k) / math.sqrt(Cin * k * k)
    b = torch.randn(Cout)
    xb, wb = bf(x), bf(w)
    xin = F.pad(xb, (pad,) * 4, mode='reflect') if mode == 'reflect' else xb
    y_ref = F.conv2d(xin, wb, b.double(), stride=stride, padding=0 if mode == 'reflect' else pad)
    OH, OW = y_ref.shape[2:]
    arena, (w_off,) = arena_of(w)
    pm = P.PAD_REFLECT if mode == 'reflect' else P.PAD_ZERO
    units = P.conv_fprop_units(w_off, Cout, Cin, k, k, pad)
    ldy = P.cpad(Cout) + 8
    geo = P.Geometry(N, H, W, P.cpad(Cin), 0, OH, OW, ldy, 8, sn=stride, pad_mode=pm)
    gm = ops.Gemm(geo, units, Cout, DEV)
    gm.pack(arena)
    xd = to_dev_nhwc(x)
    bias = b.to(DEV)
    # tcgen05 path, bf16 output
    y = torch.full((N, OH, OW, ldy), 7.0, dtype=torch.bfloat16, device=DEV)
    gm.fprop(xd, y, bias=bias)
    # SIMT restatement
    y2 = torch.full((N, OH, OW, ldy), 7.0, dtype=torch.bfloat16, device=DEV)
    gm.ref_fprop(arena, xd, y2, bias=bias)
    torch.cuda.synchronize()
    got, got2 = from_dev_nhwc(y, Cout, 8), from_dev_nhwc(y2, Cout, 8)
    assert rel_err(got2, y_ref) < 6e-3, 'SIMT restatement vs torch'
    assert rel_err(got, y_ref) < 6e-3, 'tcgen05 fprop vs torch'
    assert float(y[..., :8].float().min()) == 7.0, 'slice below y_coff must be untouched'
    if P.cpad(Cout) > Cout:
        assert float(y[..., 8 + Cout:8 + P.cpad(Cout)].float().abs().max()) == 0.0, 'padding channels must be zero'
    # fp32 output + activation + accumulate
    yf = torch.ones(N, OH, OW, ldy, dtype=torch.float32, device=DEV)
    gm.fprop(xd, yf, bias=bias, act=ops.ACT['leaky'], accumulate=True, y_is_f32=True)
    torch.cuda.synchronize()
    gotf = yf[..., 8:8 + Cout].permute(0, 3, 1, 2).double().cpu()
    assert rel_err(gotf, F.leaky_relu(y_ref + 1.0, 0.2)) < 2e-5, 'fp32 epilogue'

    # ---- weight gradient
    dy = torch.randn(N, Cout, OH, OW)
    dyd = to_dev_nhwc(dy)
    geo_w = P.Geometry(N, H, W, P.cpad(Cin), 0, OH, OW, P.cpad(Cout), 0, sn=stride, pad_mode=pm)
    gw = ops.Gemm(geo_w, units, Cout, DEV, need_pack=False)
    g1 = torch.zeros_like(arena)
    gw.wgrad(xd, dyd, g1)
    torch.cuda.synchronize()
    xr = xb.clone().requires_grad_(True)
    wr = wb.clone().requires_grad_(True)
    xin = F.pad(xr, (pad,) * 4, mode='reflect') if mode == 'reflect' else xr
    F.conv2d(xin, wr, None, stride=stride, padding=0 if mode == 'reflect' else pad).backward(bf(dy))
    gw_got = g1[w_off:w_off + w.numel()].view_as(w).double().cpu()
    assert rel_err(gw_got, wr.grad) < 2e-4, 'tcgen05 wgrad vs torch'
    assert float(g1[:w_off].abs().max()) == 0.0

    # ---- data gradient
    if mode == 'reflect':
        du = P.conv_dgrad_units(w_off, Cout, Cin, k, k, 0)
        Hp, Wp = H + 2 * pad, W + 2 * pad
        geo_d = P.Geometry(N, OH, OW, P.cpad(Cout), 0, Hp, Wp, P.cpad(Cin), 0)
        gd = ops.Gemm(geo_d, du, Cin, DEV)
        gd.pack(arena)
        dxp = torch.empty(N, Hp, Wp, P.cpad(Cin), dtype=torch.bfloat16, device=DEV)
        gd.fprop(dyd, dxp)
        dx = torch.empty(N, H, W, P.cpad(Cin), dtype=torch.bfloat16, device=DEV)
        ops.reflect_fold(ops.Act(dxp), ops.Act(dx), pad)
        torch.cuda.synchronize()
        assert rel_err(from_dev_nhwc(dx, Cin), xr.grad) < 1.5e-2, 'dgrad (padded frame + fold)'
    else:
        du = P.conv_dgrad_units(w_off, Cout, Cin, k, k, pad)
        dx = torch.zeros(N, H, W, P.cpad(Cin), dtype=torch.bfloat16, device=DEV)
        if stride == 1:
            gd = ops.Gemm(P.Geometry(N, OH, OW, P.cpad(Cout), 0, H, W, P.cpad(Cin), 0), du, Cin, DEV)
            gd.pack(arena)
            gd.fprop(dyd, dx)
        else:
            for a in range(2):
                for c in range(2):
                    ph = du.phase(a, c)
                    if len(ph) == 0:
                        continue
                    g = P.Geometry(N, OH, OW, P.cpad(Cout), 0, H, W, P.cpad(Cin), 0, sn=1, sd=2, o_step=2, o_ph=a, o_pw=c)
                    gd = ops.Gemm(g, ph, Cin, DEV)
                    gd.pack(arena)
                    gd.fprop(dyd, dx)
        torch.cuda.synchronize()
        assert rel_err(from_dev_nhwc(dx, Cin), xr.grad) < 6e-3, 'dgrad'


@pytest.mark.parametrize('Cin,Cout,H,W', [(9, 5, 6, 7), (256, 128, 16, 16)])
def test_igemm_conv_transpose(Cin, Cout, H, W):
    from cat_b200 import ops
    torch.manual_seed(3)
    N, k, pad = 2, 3, 1
    x = torch.randn(N, Cin, H, W)
    w = torch.randn(Cin, Cout, k, k) / math.sqrt(Cin * k * k / 4)
    xr, wr = bf(x).requires_grad_(True), bf(w).requires_grad_(True)
    y_ref = F.conv_transpose2d(xr, wr, None, stride=2, padding=pad, output_padding=1)
    OH, OW = y_ref.shape[2:]
    dy = torch.randn(N, Cout, OH, OW)
    y_ref.backward(bf(dy))
    arena, (w_off,) = arena_of(w)
    xd, dyd = to_dev_nhwc(x), to_dev_nhwc(dy)
    fu = P.convT_fprop_units(w_off, Cin, Cout, k, k, pad)
    y = torch.zeros(N, OH, OW, P.cpad(Cout), dtype=torch.bfloat16, device=DEV)
    for a in range(2):
        for b in range(2):
            g = P.Geometry(N, H, W, P.cpad(Cin), 0, OH, OW, P.cpad(Cout), 0, sn=1, sd=2, o_step=2, o_ph=a, o_pw=b)
            gm = ops.Gemm(g, fu.phase(a, b), Cout, DEV)
            gm.pack(arena)
            gm.fprop(xd, y)
    # un-decomposed variant must agree as well (exercises the inexact-division zero fill)
    y2 = torch.zeros_like(y)
    gm = ops.Gemm(P.Geometry(N, H, W, P.cpad(Cin), 0, OH, OW, P.cpad(Cout), 0, sn=1, sd=2), fu, Cout, DEV)
    gm.pack(arena)
    gm.fprop(xd, y2)
    torch.cuda.synchronize()
    assert rel_err(from_dev_nhwc(y, Cout), y_ref.detach()) < 6e-3
    assert rel_err(from_dev_nhwc(y2, Cout), y_ref.detach()) < 6e-3
    bu = P.convT_dgrad_units(w_off, Cin, Cout, k, k, pad)
    gd = ops.Gemm(P.Geometry(N, OH, OW, P.cpad(Cout), 0, H, W, P.cpad(Cin), 0, sn=2, sd=1), bu, Cin, DEV)
    gd.pack(arena)
    dx = torch.zeros(N, H, W, P.cpad(Cin), dtype=torch.bfloat16, device=DEV)
    gd.fprop(dyd, dx)
    g1 = torch.zeros_like(arena)
    gd.wgrad(dyd, xd, g1)
    torch.cuda.synchronize()
    assert rel_err(from_dev_nhwc(dx, Cin), xr.grad) < 6e-3
    assert rel_err(g1[w_off:w_off + w.numel()].view_as(w).double().cpu(), wr.grad) < 2e-4


def test_layout_and_elementwise_kernels():
    from cat_b200 import ops
    torch.manual_seed(0)
    N, C, H, W = 2, 5, 6, 7
    x = torch.randn(N, C, H, W)
    a = ops.Act.empty(N, H, W, 16, DEV, zero=True)
    ops.nchw_to_nhwc(x.to(DEV), a.slice(8, 8))
    torch.cuda.synchronize()
    assert torch.equal(from_dev_nhwc(a.t, C, 8), bf(x))
    assert float(a.t[..., :8].float().abs().max()) == 0 and float(a.t[..., 13:].float().abs().max()) == 0
    back = ops.nhwc_to_nchw(a.slice(8, 8), C)
    torch.cuda.synchronize()
    assert torch.equal(back.cpu().double(), bf(x))
    # copy_channels (torch.cat along channels)
    dst = ops.Act.empty(N, H, W, 8, DEV, zero=True)
    ops.copy_channels(a.slice(8, 8), dst.slice(0, 8), 3)
    ops.copy_channels(a.slice(8, 8), Act_off(dst, 3), 3)
    torch.cuda.synchronize()
    ref = torch.cat([bf(x)[:, :3], bf(x)[:, :3]], 1)
    assert torch.equal(from_dev_nhwc(dst.t, 6), ref)
    # reflect fold == adjoint of ReflectionPad2d
    for p in (1, 2, 3):
        g = torch.randn(N, 8, H + 2 * p, W + 2 * p)
        xx = torch.zeros(N, 8, H, W, dtype=torch.float64, requires_grad=True)
        F.pad(xx, (p,) * 4, mode='reflect').backward(bf(g))
        out = ops.Act.empty(N, H, W, 8, DEV)
        ops.reflect_fold(ops.Act(to_dev_nhwc(g)), out, p)
        torch.cuda.synchronize()
        assert rel_err(from_dev_nhwc(out.t, 8), xx.grad) < 6e-3
    # add / act_bwd / channel_sum
    u, v = torch.randn(N, 8, H, W), torch.randn(N, 8, H, W)
    o = ops.Act.empty(N, H, W, 8, DEV)
    ops.add(ops.Act(to_dev_nhwc(u)), ops.Act(to_dev_nhwc(v)), o)
    torch.cuda.synchronize()
    assert rel_err(from_dev_nhwc(o.t, 8), bf(u) + bf(v)) < 5e-3
    for name, fn in (('relu', torch.relu), ('leaky', lambda t: F.leaky_relu(t, 0.2)), ('tanh', torch.tanh)):
        z = bf(u).requires_grad_(True)
        outv = fn(z)
        outv.backward(bf(v))
        ops.act_bwd(ops.Act(to_dev_nhwc(v)), ops.Act(to_dev_nhwc(outv.detach().float())), o, ops.ACT[name])
        torch.cuda.synchronize()
        assert rel_err(from_dev_nhwc(o.t, 8), z.grad) < 1.2e-2, name
    s = torch.zeros(8, device=DEV)
    ops.channel_sum(ops.Act(to_dev_nhwc(u)), s)
    torch.cuda.synchronize()
    assert rel_err(s.double().cpu(), bf(u).sum((0, 2, 3))) < 1e-5


def Act_off(act, c):
    """helper: a view whose slice starts at an arbitrary (unaligned) channel -- only for copy_channels"""
    from cat_b200 import ops
    v = ops.Act(act.t)
    v.coff, v.C = c, act.ld - c
    return v


@pytest.mark.parametrize('per_sample,C,N,H,W', [(False, 24, 3, 9, 7), (True, 24, 3, 9, 7), (False, 512, 2, 4, 4),
                                                (True, 8, 2, 31, 31), (False, 2048, 1, 4, 8)])
def test_norm_forward_backward(per_sample, C, N, H, W):
    from cat_b200 import ops
    torch.manual_seed(1)
    x = torch.randn(N, C, H, W) * 2 + 0.5
    gamma, beta = torch.rand(C) + 0.5, torch.randn(C) * 0.1
    dout = torch.randn(N, C, H, W)
    xr, gr, br = bf(x).requires_grad_(True), gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    rm, rv = torch.zeros(C, dtype=torch.float64), torch.ones(C, dtype=torch.float64)
    if per_sample:
        z = F.instance_norm(xr, None, None, gr, br, True, 0.1, 1e-5)
    else:
        z = F.batch_norm(xr, rm, rv, gr, br, True, 0.1, 1e-5)
    out_ref = F.leaky_relu(z, 0.2)
    out_ref.backward(bf(dout))
    G = N if per_sample else 1
    xa = ops.Act(to_dev_nhwc(x))
    sums = torch.zeros(G, 2, C, device=DEV)
    ops.norm_stats(xa, per_sample, sums)
    scale, shift = torch.empty(G, C, device=DEV), torch.empty(G, C, device=DEV)
    mr = torch.empty(G, 2, C, device=DEV)
    gd, bd = gamma.to(DEV), beta.to(DEV)
    rmd, rvd = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
    count = H * W if per_sample else N * H * W
    ops.norm_finalize(sums, G, C, count, 1e-5, 0.1, gd, bd, None if per_sample else rmd, None if per_sample else rvd,
                      scale, shift, mr)
    ya = ops.Act.empty(N, H, W, C, DEV)
    ops.norm_apply(xa, ya, scale, shift, per_sample, ops.ACT['leaky'])
    torch.cuda.synchronize()
    assert rel_err(from_dev_nhwc(ya.t, C), out_ref.detach()) < 6e-3
    if not per_sample:
        assert rel_err(rmd.double().cpu(), rm) < 1e-4 and rel_err(rvd.double().cpu(), rv) < 1e-4
    # backward
    red = torch.zeros(G, 2, C, device=DEV)
    da = ops.Act(to_dev_nhwc(dout))
    out_saved = ops.Act(to_dev_nhwc(out_ref.detach().float()))
    ops.norm_bwd_reduce(da, out_saved, xa, per_sample, mr, ops.ACT['leaky'], red)
    dxa = ops.Act.empty(N, H, W, C, DEV)
    dg, db = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
    ops.norm_bwd_apply(da, out_saved, xa, dxa, per_sample, mr, gd, red, count, ops.ACT['leaky'], dg, db)
    torch.cuda.synchronize()
    assert rel_err(from_dev_nhwc(dxa.t, C), xr.grad) < 1.5e-2
    assert rel_err(dg.double().cpu(), gr.grad) < 2e-3 and rel_err(db.double().cpu(), br.grad) < 2e-3
    # eval mode (running statistics) + residual add
    if not per_sample:
        rme, rve = torch.randn(C).to(DEV), (torch.rand(C) + 0.5).to(DEV)
        ops.norm_finalize(None, 1, C, count, 1e-5, 0.1, gd, bd, rme, rve, scale, shift, None)
        ops.norm_apply(xa, ya, scale, shift, False, ops.ACT['none'], residual=xa)
        torch.cuda.synchronize()
        ref = F.batch_norm(bf(x), rme.double().cpu(), rve.double().cpu(), gamma.double(), beta.double(), False) + bf(x)
        assert rel_err(from_dev_nhwc(ya.t, C), ref) < 6e-3


@pytest.mark.parametrize('ks', [(1, 3, 5), (5,), (3, 7)])
def test_depthwise(ks):
    from cat_b200 import ops
    torch.manual_seed(4)
    N, H, W = 2, 9, 11
    mids = [3 + 2 * i for i in range(len(ks))]
    ws = [torch.randn(m, 1, k, k) for m, k in zip(mids, ks)]
    arena, offs = arena_of(*ws)
    C = sum(P.cpad(m) for m in mids)
    ksize = torch.ones(C, dtype=torch.int32)
    w_off = torch.full((C,), -1, dtype=torch.int32)
    xs, c0 = [], 0
    xbuf = torch.zeros(N, H, W, C, dtype=torch.bfloat16)
    dybuf = torch.zeros(N, H, W, C, dtype=torch.bfloat16)
    refs = []
    for m, k, w, off in zip(mids, ks, ws, offs):
        x = torch.randn(N, m, H, W)
        dy = torch.randn(N, m, H, W)
        xbuf[..., c0:c0 + m] = x.permute(0, 2, 3, 1).to(torch.bfloat16)
        dybuf[..., c0:c0 + m] = dy.permute(0, 2, 3, 1).to(torch.bfloat16)
        ksize[c0:c0 + P.cpad(m)] = k
        w_off[c0:c0 + m] = off + torch.arange(m, dtype=torch.int32) * k * k
        xr, wr = bf(x).requires_grad_(True), w.double().requires_grad_(True)
        y = F.conv2d(F.pad(xr, ((k - 1) // 2,) * 4, mode='reflect') if k > 1 else xr, wr, groups=m)
        y.backward(bf(dy))
        refs.append((c0, m, y.detach(), xr.grad, wr.grad, off))
        c0 += P.cpad(m)
    xa, dya = ops.Act(xbuf.to(DEV)), ops.Act(dybuf.to(DEV))
    ya, dxa = ops.Act.empty(N, H, W, C, DEV), ops.Act.empty(N, H, W, C, DEV)
    kd, wd = ksize.to(DEV), w_off.to(DEV)
    garena = torch.zeros_like(arena)
    ops.dwconv_fwd(xa, ya, kd, wd, arena)
    ops.dwconv_bwd_data(dya, dxa, kd, wd, arena)
    ops.dwconv_bwd_weight(xa, dya, kd, wd, garena)
    torch.cuda.synchronize()
    for (c0, m, y, dx, dw, off) in refs:
        assert rel_err(from_dev_nhwc(ya.t, m, c0), y) < 6e-3
        assert rel_err(from_dev_nhwc(dxa.t, m, c0), dx) < 6e-3
        assert rel_err(garena[off:off + dw.numel()].view_as(dw).double().cpu(), dw) < 1e-4


@pytest.mark.parametrize('mode', ['hinge', 'lsgan', 'vanilla'])
def test_gan_and_l1_losses(mode):
    from cat_b200 import ops
    from oracle import cat_oracle as O
    torch.manual_seed(5)
    n = 2 * 30 * 30
    pred = torch.randn(n)
    pd = torch.zeros(n, 8, device=DEV)
    pd[:, 0] = pred.to(DEV)
    for real, for_d in ((True, True), (False, True), (True, False)):
        pr = pred.double().requires_grad_(True)
        ref = O.gan_loss(mode, pr, real, for_d)
        ref.backward()
        loss = torch.zeros(1, device=DEV)
        dp = ops.Act.empty(1, 1, n, 8, DEV)
        ops.gan_loss(pd, n, 8, mode, real, for_d, 0.5, loss, dp)
        torch.cuda.synchronize()
        assert abs(float(loss) - float(ref)) < 1e-5 * max(1, abs(float(ref)))
        got = dp.t.view(n, 8)[:, 0].double().cpu()
        assert rel_err(got, 0.5 * pr.grad) < 5e-3
        assert float(dp.t.view(n, 8)[:, 1:].float().abs().max()) == 0
    a, b = torch.rand(2, 3, 8, 9) * 2 - 1, torch.rand(2, 3, 8, 9) * 2 - 1
    ar = bf(a).requires_grad_(True)
    ref = F.l1_loss(ar, bf(b)) * 100.0
    ref.backward()
    extra = torch.randn(2, 3, 8, 9) * 1e-3
    loss = torch.zeros(1, device=DEV)
    da = ops.Act.empty(2, 8, 9, 8, DEV)
    ops.recon_loss(ops.Act(to_dev_nhwc(a)), ops.Act(to_dev_nhwc(b)), 3, 'l1', 100.0, loss, da, ops.Act(to_dev_nhwc(extra)))
    torch.cuda.synchronize()
    assert abs(float(loss) * 100.0 - float(ref)) < 1e-4 * float(ref)
    assert rel_err(from_dev_nhwc(da.t, 3), ar.grad + bf(extra)) < 6e-3
    for kind, fn in (('l2', F.mse_loss), ('smooth_l1', F.smooth_l1_loss)):
        ar = bf(a * 1.5).requires_grad_(True)
        ref = fn(ar, bf(b)) * 10.0
        ref.backward()
        loss.zero_()
        ops.recon_loss(ops.Act(to_dev_nhwc(a * 1.5)), ops.Act(to_dev_nhwc(b)), 3, kind, 10.0, loss, da)
        torch.cuda.synchronize()
        assert abs(float(loss) * 10.0 - float(ref)) < 1e-4 * float(ref), kind
        assert rel_err(from_dev_nhwc(da.t, 3), ar.grad) < 6e-3, kind


@pytest.mark.parametrize('B,Cs,Ct,H,W', [(2, 5, 16, 8, 8), (16, 62, 256, 16, 16), (1, 8, 8, 4, 4), (32, 24, 48, 8, 8)])
def test_ka_loss_forward_backward(B, Cs, Ct, H, W):
    from cat_b200 import ops
    from oracle import cat_oracle as O
    torch.manual_seed(6)
    X, Y = torch.randn(B, Cs, H, W) + 0.3, torch.randn(B, Ct, H, W) + 0.3
    Xr = bf(X).requires_grad_(True)
    ref = -O.ka(Xr, bf(Y)) * 0.5
    ref.backward()
    xa, ya = ops.Act(to_dev_nhwc(X)), ops.Act(to_dev_nhwc(Y))
    Gx, Gy = torch.zeros(B, B, device=DEV), torch.zeros(B, B, device=DEV)
    ops.gram(xa, Gx)
    ops.gram(ya, Gy)
    loss, kav, coef = torch.zeros(1, device=DEV), torch.zeros(1, device=DEV), torch.zeros(B, B, device=DEV)
    ops.ka_finish(Gx, Gy, B, -0.5, loss, kav, coef)
    dx = ops.Act(torch.ones(B, H, W, P.cpad(Cs), dtype=torch.bfloat16, device=DEV))
    ops.ka_bwd(xa, coef, dx, True)
    torch.cuda.synchronize()
    assert rel_err(Gx.double().cpu(), bf(X).flatten(1) @ bf(X).flatten(1).T) < 1e-5
    assert abs(float(loss) - float(ref)) < 2e-6
    assert abs(float(kav) * -0.5 - float(ref)) < 2e-6
    got = from_dev_nhwc(dx.t, Cs) - 1.0
    scale = float(Xr.grad.abs().max())
    if B == 1:
        assert float(got.abs().max()) < 1e-6
    else:
        # dX is added in bf16 to a buffer holding 1.0: absolute resolution 2^-8 of the sum
        assert float((got - Xr.grad).abs().max()) < 4e-3 + 1e-2 * scale


def test_adam_matches_torch():
    from cat_b200 import ops
    torch.manual_seed(7)
    n = 10007
    p0, g1, g2 = torch.randn(n), torch.randn(n), torch.randn(n) * 0.1
    pr = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([pr], lr=2e-4, betas=(0.5, 0.999))
    p, m, v = p0.to(DEV), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    lr, step = torch.tensor([2e-4], device=DEV), torch.zeros(1, dtype=torch.int32, device=DEV)
    for g in (g1, g2, g1):
        pr.grad = g.clone()
        opt.step()
        ops.adam(p, g.to(DEV), m, v, lr, 0.5, 0.999, 1e-8, 1.0, step)
    torch.cuda.synchronize()
    assert int(step) == 3
    assert float((p.cpu() - pr.detach()).abs().max()) < 2e-6
