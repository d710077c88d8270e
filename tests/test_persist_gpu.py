"""GPU parity of the v3 forward kernel (persistent halo pipeline, cat_b200/csrc/igemm_halo_persist.cu): both halo
producers -- cp.async.bulk.tensor tiles from a 4-D tensor map (mode 2) and the cp.async threads (mode 1) -- against torch
(fp64 on bf16-rounded operands) and against the v1 gather-per-tap kernel, on shapes with several tiles per CTA (the
rings and the two accumulator stages wrap), several N tiles, parity planes, channel slices and partial chunks."""
import math

import pytest
import torch
import torch.nn.functional as F

from cat_b200 import igemm_plan as P

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(120)]
DEV = 'cuda:0'


@pytest.fixture(scope='module', autouse=True)
def _init():
    from cat_b200 import ops
    ops.require_cuda()


def bf(x):
    return x.to(torch.bfloat16).to(torch.float64)


def to_dev_nhwc(x, ld=None, coff=0, fill=0.0):
    N, C, H, W = x.shape
    ld = ld or P.cpad(C)
    out = torch.full((N, H, W, ld), fill, dtype=torch.bfloat16)
    out[..., coff:coff + P.cpad(C)] = 0
    out[..., coff:coff + C] = x.permute(0, 2, 3, 1).to(torch.bfloat16)
    return out.to(DEV)


def from_dev_nhwc(t, C, coff=0):
    return t[..., coff:coff + C].permute(0, 3, 1, 2).to(torch.float64).cpu()


def rel_err(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


CASES = [
    # k, stride, pad, mode, Cin, Cout, N, H, W
    (3, 1, 1, 'zero', 5, 7, 2, 10, 12),
    (5, 1, 2, 'reflect', 72, 42, 2, 16, 16),     # two channel chunks, second one partial
    (7, 1, 3, 'reflect', 64, 3, 1, 24, 40),      # generator head
    (1, 1, 0, 'reflect', 72, 40, 2, 8, 8),       # 1x1: no border taps, TMA although the padding rule is reflection
    (1, 1, 0, 'zero', 24, 40, 2, 8, 8),
    (3, 2, 1, 'zero', 17, 31, 2, 16, 16),        # 4 parity planes (traversal stride 2)
    (4, 2, 1, 'zero', 64, 128, 2, 16, 16),
    (4, 1, 1, 'zero', 128, 1, 2, 9, 9),
    (4, 1, 1, 'zero', 128, 300, 1, 12, 12),      # two N tiles
    (3, 1, 1, 'zero', 64, 64, 4, 96, 96),        # 288+ tiles: every CTA walks several (rings / accumulator stages wrap)
    (1, 1, 0, 'zero', 128, 24, 8, 128, 128),     # short K, 1024 tiles
    (4, 2, 1, 'zero', 24, 64, 4, 128, 128),      # stride 2 with many tiles
    (5, 1, 2, 'reflect', 40, 24, 4, 64, 64),     # reflection: cp.async producers / TMA + fringe pass in the persistent pipeline
    (3, 1, 1, 'reflect', 136, 48, 3, 40, 56),    # three channel chunks, strips touch both borders
    (7, 1, 3, 'reflect', 24, 40, 2, 12, 12),     # halo wider than a third of the image
]


def _tile_forms(OW, Cout):
    forms = [('auto', None), ('msub2', (OW, 2)), ('strips', (max(4, OW // 3), 1))]
    return forms


@pytest.mark.parametrize('k,stride,pad,mode,Cin,Cout,N,H,W', CASES)
@pytest.mark.parametrize('pmode', [1, 2, 3])
@pytest.mark.parametrize('tile', ['auto', 'msub2', 'strips'])
def test_persistent_fprop_and_dgrad(k, stride, pad, mode, Cin, Cout, N, H, W, pmode, tile):
    from cat_b200 import ops
    torch.manual_seed(k * 100 + Cin)
    x = torch.randn(N, Cin, H, W)
    w = torch.randn(Cout, Cin, k, k) / math.sqrt(Cin * k * k)
    b = torch.randn(Cout)
    xb, wb = bf(x).requires_grad_(True), bf(w)
    xin = F.pad(xb, (pad,) * 4, mode='reflect') if mode == 'reflect' else xb
    y_ref = F.conv2d(xin, wb, b.double(), stride=stride, padding=0 if mode == 'reflect' else pad)
    OH, OW = y_ref.shape[2:]
    arena = torch.cat([torch.zeros(5), w.flatten()]).to(DEV)
    pm = P.PAD_REFLECT if mode == 'reflect' else P.PAD_ZERO
    units = P.conv_fprop_units(5, Cout, Cin, k, k, pad)
    ldy = P.cpad(Cout) + 8
    # the input is a channel slice [16, 16 + Cin) of a wider buffer whose other channels hold a large value: nothing of
    # it may leak into the result (tensor-map base / visible channel count, zero fill of partial chunks)
    ldx, xc = P.cpad(Cin) + 24, 16
    geo = P.Geometry(N, H, W, ldx, xc, OH, OW, ldy, 8, sn=stride, pad_mode=pm)
    force = dict(_tile_forms(OW, Cout))[tile]
    if force is not None and force[1] * 2 * P.choose_n_tile(Cout) > 512:
        pytest.skip('two accumulator stages of this tiling exceed 512 TMEM columns')
    gm = ops.Gemm(geo, units, Cout, DEV, force_tile=force, force_mode=pmode)
    if gm.halo is None or not gm.tilings:
        pytest.skip('mode %d does not apply to this shape (reflection padding with border taps / does not fit)' % pmode)
    assert all(t[4] == pmode for t in gm.tilings)
    gm.pack(arena)
    gm.choice = 'v2'
    xd, bias = to_dev_nhwc(x, ldx, xc, fill=1000.0), b.to(DEV)
    ref = F.leaky_relu(y_ref.detach(), 0.2)
    y1 = torch.full((N, OH, OW, ldy), 7.0, dtype=torch.bfloat16, device=DEV)
    gm.fprop(xd, y1, bias=bias, act=ops.ACT['leaky'], force_v1=True)
    for cand in gm.tilings:
        gm._use_tiling(cand)
        y3 = torch.full((N, OH, OW, ldy), 7.0, dtype=torch.bfloat16, device=DEV)
        gm.fprop(xd, y3, bias=bias, act=ops.ACT['leaky'])
        torch.cuda.synchronize()
        assert rel_err(from_dev_nhwc(y3, Cout, 8), ref) < 6e-3, 'v3 persistent kernel vs torch'
        assert float((y1.float() - y3.float()).abs().max()) <= 2 ** -7 * float(ref.abs().max()), 'v3 vs v1'
        assert float(y3[..., :8].float().min()) == 7.0 and float(y3[..., :8].float().max()) == 7.0
        # run it again into the same buffer: a persistent kernel must leave no state behind
        gm.fprop(xd, y3, bias=bias, act=ops.ACT['leaky'])
        torch.cuda.synchronize()
        assert rel_err(from_dev_nhwc(y3, Cout, 8), ref) < 6e-3
    # fp32 output with accumulate (the epilogue's second path)
    yf = torch.ones(N, OH, OW, ldy, dtype=torch.float32, device=DEV)
    gm.fprop(xd, yf, bias=bias, accumulate=True, y_is_f32=True)
    torch.cuda.synchronize()
    assert rel_err(yf[..., 8:8 + Cout].permute(0, 3, 1, 2).double().cpu(), y_ref.detach() + 1.0) < 2e-5
    # input gradient (zero-padded convs: direct / 4 sub-pixel phases)
    if mode == 'zero' and N * H * W <= 40000:
        dy = torch.randn(N, Cout, OH, OW)
        y_ref.backward(bf(dy))
        dyd = to_dev_nhwc(dy)
        du = P.conv_dgrad_units(5, Cout, Cin, k, k, pad)
        dx = torch.zeros(N, H, W, P.cpad(Cin), dtype=torch.bfloat16, device=DEV)
        if stride == 1:
            gd = ops.Gemm(P.Geometry(N, OH, OW, P.cpad(Cout), 0, H, W, P.cpad(Cin), 0), du, Cin, DEV, force_mode=pmode)
            assert gd.halo is not None
            gd.choice = 'v2'
            gd.pack(arena)
            gd.fprop(dyd, dx)
        else:
            for a in range(2):
                for c in range(2):
                    ph = du.phase(a, c)
                    if len(ph) == 0:
                        continue
                    g = P.Geometry(N, OH, OW, P.cpad(Cout), 0, H, W, P.cpad(Cin), 0, sn=1, sd=2, o_step=2, o_ph=a, o_pw=c)
                    gd = ops.Gemm(g, ph, Cin, DEV, force_mode=pmode)
                    assert gd.halo is not None
                    gd.choice = 'v2'
                    gd.pack(arena)
                    gd.fprop(dyd, dx)
        torch.cuda.synchronize()
        assert rel_err(from_dev_nhwc(dx, Cin), xb.grad) < 6e-3, 'v3 dgrad'


def test_autotune_offers_the_persistent_variants():
    """Without force_mode every Gemm of a zero-padded conv lists v2, and v3 with TMA; a reflection-padded 3x3 lists v3 with
    the cp.async producers."""
    from cat_b200 import ops
    units = P.conv_fprop_units(0, 64, 64, 3, 3, 1)
    g0 = ops.Gemm(P.Geometry(2, 32, 32, 64, 0, 32, 32, 64, 0, pad_mode=P.PAD_ZERO), units, 64, DEV)
    g1 = ops.Gemm(P.Geometry(2, 32, 32, 64, 0, 32, 32, 64, 0, pad_mode=P.PAD_REFLECT), units, 64, DEV)
    assert {t[4] for t in g0.tilings} == {0, 2}
    assert {t[4] for t in g1.tilings} == {0, 1}      # (+ 3, TMA with the reflection fringe pass, under CATB_TMA_REFLECT=1)


@pytest.mark.parametrize('k,stride,pad,Cin,Cout,N,H,W', [
    (3, 1, 1, 24, 40, 3, 20, 28),       # several tiles per image, garbage positions in pitch space
    (1, 1, 0, 64, 300, 2, 16, 16),      # two N tiles
    (4, 2, 1, 16, 72, 4, 64, 64),       # stride 2, many tiles per CTA in the persistent kernel
    (3, 1, 1, 64, 64, 4, 96, 96),
])
@pytest.mark.parametrize('pmode', [0, 1, 2])
@pytest.mark.parametrize('per_sample', [False, True])
def test_fused_epilogue_statistics(k, stride, pad, Cin, Cout, N, H, W, pmode, per_sample):
    """Sum / sum of squares accumulated by the conv epilogue (v2 and both v3 producers) == catb_norm_stats over the
    stored output (same bf16 values; fp32 summation order differs), into a channel slice of a wider statistics row."""
    from cat_b200 import ops
    torch.manual_seed(k + Cout)
    x = torch.randn(N, Cin, H, W)
    w = torch.randn(Cout, Cin, k, k) / math.sqrt(Cin * k * k)
    OH, OW = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    arena = w.flatten().to(DEV)
    units = P.conv_fprop_units(0, Cout, Cin, k, k, pad)
    Cp = P.cpad(Cout)
    ldy, yc = Cp + 16, 8
    geo = P.Geometry(N, H, W, P.cpad(Cin), 0, OH, OW, ldy, yc, sn=stride, pad_mode=P.PAD_ZERO)
    gm = ops.Gemm(geo, units, Cout, DEV, force_mode=pmode)
    assert gm.halo is not None and gm.tilings
    gm.pack(arena)
    gm.choice = 'v2'
    xd = to_dev_nhwc(x)
    G = N if per_sample else 1
    Cs, coff = Cp + 24, 16          # statistics row wider than the GEMM's channels
    for cand in gm.tilings:
        gm._use_tiling(cand)
        y = torch.zeros(N, OH, OW, ldy, dtype=torch.bfloat16, device=DEV)
        sums = torch.zeros(G, 2, Cs, device=DEV)
        assert gm.fprop(xd, y, act=ops.ACT['leaky'], stats=(sums, Cs, coff, per_sample)) is True
        ref = torch.zeros(G, 2, Cp, device=DEV)
        ops.norm_stats(ops.Act(y, yc, Cp), per_sample, ref)
        torch.cuda.synchronize()
        got = sums[:, :, coff:coff + Cp]
        assert float((got - ref).abs().max()) <= 2e-4 * float(ref.abs().max()), (cand[0], cand[1], cand[4])
        assert float(sums[:, :, :coff].abs().max()) == 0 and float(sums[:, :, coff + Cp:].abs().max()) == 0


@pytest.mark.parametrize('per_sample,track,residual', [(False, True, False), (True, False, True), (False, False, True)])
def test_norm_apply_fused_matches_finalize_plus_apply(per_sample, track, residual):
    from cat_b200 import ops
    torch.manual_seed(3)
    N, H, W, C = 3, 12, 20, 40
    x = ops.Act((torch.randn(N, H, W, C, device=DEV) * 2 + 0.5).to(torch.bfloat16))
    r = ops.Act(torch.randn(N, H, W, C, device=DEV).to(torch.bfloat16)) if residual else None
    G = N if per_sample else 1
    count = H * W if per_sample else N * H * W
    gamma, beta = torch.rand(C, device=DEV) + 0.5, torch.randn(C, device=DEV)
    sums = torch.zeros(G, 2, C, device=DEV)
    ops.norm_stats(x, per_sample, sums)
    out = []
    for fused in (False, True):
        rm, rv = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
        scale, shift, mr = (torch.zeros(G, C, device=DEV), torch.zeros(G, C, device=DEV), torch.zeros(G, 2, C, device=DEV))
        y = ops.Act(torch.zeros(N, H, W, C, dtype=torch.bfloat16, device=DEV))
        a = (rm if track else None, rv if track else None)
        if fused:
            ops.norm_apply_fused(x, y, sums, count, 1e-5, 0.1, gamma, beta, a[0], a[1], scale, shift, mr, per_sample,
                                 ops.ACT['relu'], r)
        else:
            ops.norm_finalize(sums, G, C, count, 1e-5, 0.1, gamma, beta, a[0], a[1], scale, shift, mr)
            ops.norm_apply(x, y, scale, shift, per_sample, ops.ACT['relu'], r)
        torch.cuda.synchronize()
        out.append((y.t.float(), scale, shift, mr, rm, rv))
    assert float((out[0][0] - out[1][0]).abs().max()) <= 2 ** -7 * float(out[0][0].abs().max())   # bf16 outputs: one ulp
    for a, b in zip(out[0][1:], out[1][1:]):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('k,stride,pad,Cin,Cout,N,H,W', [
    (3, 1, 1, 5, 7, 2, 10, 12),
    (1, 1, 0, 72, 40, 2, 8, 8),
    (3, 2, 1, 17, 31, 2, 16, 16),        # 4 parity planes
    (4, 2, 1, 64, 200, 2, 32, 32),       # two channel tiles of dY, PatchGAN-like
    (4, 1, 1, 128, 72, 2, 17, 17),
    (4, 2, 1, 128, 256, 4, 64, 64),      # four planes fill the SM: the 64-position tile form
    (3, 1, 1, 64, 64, 4, 96, 96),        # many tiles per CTA: the ring wraps
    (1, 1, 0, 128, 24, 8, 64, 64),
    (1, 1, 0, 1024, 16, 16, 31, 31),     # 16 channel chunks: few splits, 5 tiles per CTA through a 3-deep ring
])
def test_tma_weight_gradient(k, stride, pad, Cin, Cout, N, H, W):
    """Weight gradient with both operands staged by cp.async.bulk.tensor (dY tile + X halo planes as tensor-map boxes)
    against torch and, bit for bit in structure, against the cp.async form (same MMAs; the row splits may differ)."""
    from cat_b200 import ops
    torch.manual_seed(k * 10 + Cout)
    x = torch.randn(N, Cin, H, W)
    w = torch.randn(Cout, Cin, k, k)
    xb, wb = bf(x), bf(w).requires_grad_(True)
    y_ref = F.conv2d(xb, wb, None, stride=stride, padding=pad)
    OH, OW = y_ref.shape[2:]
    dy = torch.randn(N, Cout, OH, OW)
    y_ref.backward(bf(dy))
    units = P.conv_fprop_units(5, Cout, Cin, k, k, pad)
    ldx, xc = P.cpad(Cin) + 24, 16           # channel slices of wider buffers on both sides
    geo = P.Geometry(N, H, W, ldx, xc, OH, OW, P.cpad(Cout) + 8, 8, sn=stride, pad_mode=P.PAD_ZERO)
    gw = ops.Gemm(geo, units, Cout, DEV, need_pack=False)
    xd, dyd = to_dev_nhwc(x, ldx, xc, fill=1000.0), to_dev_nhwc(dy, P.cpad(Cout) + 8, 8, fill=1000.0)
    gw._wgrad_plan()
    assert gw.w_halo is not None and gw.w_tma_ok, 'the TMA form must apply to every zero-padded single-strip conv'
    gw.w_choice = 'v2'
    out = []
    for tma in (False, True):
        gw.w_tma = tma
        g = torch.zeros(5 + w.numel() + 64, device=DEV)
        gw.wgrad(xd, dyd, g)
        torch.cuda.synchronize()
        got = g[5:5 + w.numel()].view_as(w).double().cpu()
        assert rel_err(got, wb.grad) < 2e-4, 'TMA weight gradient' if tma else 'cp.async weight gradient'
        assert float(g[:5].abs().max()) == 0 and float(g[5 + w.numel():].abs().max()) == 0
        g_again = torch.zeros_like(g)
        gw.wgrad(xd, dyd, g_again)
        torch.cuda.synchronize()
        assert torch.equal(g, g_again), 'two-stage weight gradient must be deterministic'
        out.append(got)
    assert rel_err(out[1], out[0]) < 1e-5
