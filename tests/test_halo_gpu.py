"""GPU parity of the v2 (halo / shifted-window) forward kernel: against torch (fp64 on bf16-rounded
operands) and bit-for-bit against the v1 gather-per-tap kernel, which reads the same packed weights and
accumulates the same products in fp32 (only the summation order of the K steps differs)."""
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
    N, C, H, W = x.shape
    ld = ld or P.cpad(C)
    out = torch.zeros(N, H, W, ld, dtype=torch.bfloat16)
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
    (7, 1, 3, 'reflect', 3, 17, 2, 20, 24),      # generator stem
    (1, 1, 0, 'zero', 24, 40, 2, 8, 8),
    (3, 2, 1, 'zero', 17, 31, 2, 16, 16),        # 4 parity planes
    (4, 2, 1, 'zero', 64, 128, 2, 16, 16),
    (4, 1, 1, 'zero', 128, 1, 2, 9, 9),
    (4, 1, 1, 'zero', 128, 300, 1, 12, 12),      # two N tiles
    (5, 1, 2, 'reflect', 256, 48, 1, 16, 16),
]


@pytest.mark.parametrize('k,stride,pad,mode,Cin,Cout,N,H,W', CASES)
@pytest.mark.parametrize('tile', ['auto', 'msub2', 'msub4', 'strips'])
def test_halo_fprop_and_dgrad(k, stride, pad, mode, Cin, Cout, N, H, W, tile):
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
    geo = P.Geometry(N, H, W, P.cpad(Cin), 0, OH, OW, ldy, 8, sn=stride, pad_mode=pm)
    force = {'auto': None, 'msub2': (OW, 2), 'msub4': (OW, 4), 'strips': (max(4, OW // 3), 1)}[tile]
    if tile == 'msub4' and 4 * P.choose_n_tile(Cout) > 512:
        pytest.skip('four sub-tiles need 4 * n_tile <= 512 TMEM columns')
    gm = ops.Gemm(geo, units, Cout, DEV, force_tile=force)
    if tile == 'msub4' and gm.halo is None:
        pytest.skip('four sub-tiles of this shape do not fit in shared memory')
    assert gm.halo is not None, 'every conv of the path must qualify for the halo kernel'
    gm.pack(arena)     # packs both weight images (v1: compact table, v2: chunk-aligned table)
    gm.choice = 'v2'   # then pin the kernel under test (the engine autotunes v1 / v2 per GEMM)
    xd, bias = to_dev_nhwc(x), b.to(DEV)
    y2 = torch.full((N, OH, OW, ldy), 7.0, dtype=torch.bfloat16, device=DEV)
    gm.fprop(xd, y2, bias=bias, act=ops.ACT['leaky'])
    y1 = torch.full((N, OH, OW, ldy), 7.0, dtype=torch.bfloat16, device=DEV)
    gm.fprop(xd, y1, bias=bias, act=ops.ACT['leaky'], force_v1=True)
    torch.cuda.synchronize()
    ref = F.leaky_relu(y_ref.detach(), 0.2)
    assert rel_err(from_dev_nhwc(y1, Cout, 8), ref) < 6e-3, 'v1 on the chunk-aligned table'
    assert rel_err(from_dev_nhwc(y2, Cout, 8), ref) < 6e-3, 'v2 halo kernel vs torch'
    assert float((y1.float() - y2.float()).abs().max()) <= 2 ** -7 * float(ref.abs().max()), 'v2 vs v1'
    assert float(y2[..., :8].float().min()) == 7.0 and float(y2[..., :8].float().max()) == 7.0
    # fp32 output with accumulate
    yf = torch.ones(N, OH, OW, ldy, dtype=torch.float32, device=DEV)
    gm.fprop(xd, yf, bias=bias, accumulate=True, y_is_f32=True)
    torch.cuda.synchronize()
    assert rel_err(yf[..., 8:8 + Cout].permute(0, 3, 1, 2).double().cpu(), y_ref.detach() + 1.0) < 2e-5
    # input gradient through the halo kernel (zero-padded convs: direct / 4 phases)
    if mode == 'zero':
        dy = torch.randn(N, Cout, OH, OW)
        y_ref.backward(bf(dy))
        dyd = to_dev_nhwc(dy)
        du = P.conv_dgrad_units(5, Cout, Cin, k, k, pad)
        dx = torch.zeros(N, H, W, P.cpad(Cin), dtype=torch.bfloat16, device=DEV)
        if stride == 1:
            gd = ops.Gemm(P.Geometry(N, OH, OW, P.cpad(Cout), 0, H, W, P.cpad(Cin), 0), du, Cin, DEV)
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
                    gd = ops.Gemm(g, ph, Cin, DEV)
                    assert gd.halo is not None
                    gd.choice = 'v2'
                    gd.pack(arena)
                    gd.fprop(dyd, dx)
        torch.cuda.synchronize()
        assert rel_err(from_dev_nhwc(dx, Cin), xb.grad) < 6e-3, 'halo dgrad'


@pytest.mark.parametrize('k,stride,pad,mode,Cin,Cout,N,H,W', CASES + [(3, 1, 1, 'zero', 40, 200, 2, 12, 12)])
def test_halo_wgrad(k, stride, pad, mode, Cin, Cout, N, H, W):
    """v2 weight gradient (MN-major shifted windows, one TMEM accumulator per tap) vs torch and vs v1."""
    from cat_b200 import ops
    torch.manual_seed(k * 10 + Cout)
    x = torch.randn(N, Cin, H, W)
    w = torch.randn(Cout, Cin, k, k)
    xb, wb = bf(x), bf(w).requires_grad_(True)
    xin = F.pad(xb, (pad,) * 4, mode='reflect') if mode == 'reflect' else xb
    y_ref = F.conv2d(xin, wb, None, stride=stride, padding=0 if mode == 'reflect' else pad)
    OH, OW = y_ref.shape[2:]
    dy = torch.randn(N, Cout, OH, OW)
    y_ref.backward(bf(dy))
    pm = P.PAD_REFLECT if mode == 'reflect' else P.PAD_ZERO
    units = P.conv_fprop_units(5, Cout, Cin, k, k, pad)
    geo = P.Geometry(N, H, W, P.cpad(Cin), 0, OH, OW, P.cpad(Cout) + 8, 8, sn=stride, pad_mode=pm)
    gw = ops.Gemm(geo, units, Cout, DEV, need_pack=False)
    xd, dyd = to_dev_nhwc(x), to_dev_nhwc(dy, P.cpad(Cout) + 8, 8)
    g2 = torch.zeros(5 + w.numel() + 64, device=DEV)
    g1 = torch.zeros_like(g2)
    gw._wgrad_plan()
    assert gw.w_halo is not None
    gw.w_choice = 'v2'
    gw.wgrad(xd, dyd, g2)
    gw.wgrad(xd, dyd, g1, force_v1=True)
    torch.cuda.synchronize()
    got2 = g2[5:5 + w.numel()].view_as(w).double().cpu()
    got1 = g1[5:5 + w.numel()].view_as(w).double().cpu()
    assert rel_err(got1, wb.grad) < 2e-4, 'v1 wgrad'
    assert rel_err(got2, wb.grad) < 2e-4, 'v2 (halo) wgrad'
    assert float(g2[:5].abs().max()) == 0 and float(g2[5 + w.numel():].abs().max()) == 0
    # the default path is the two-stage form (workspace + catb_wgrad_unpack): bit-identical from run to run, and equal to
    # the one-launch atomic form up to the summation order of the row splits
    g3, g4, g5 = torch.zeros_like(g2), torch.zeros_like(g2), torch.zeros_like(g2)
    gw.wgrad(xd, dyd, g3)
    gw.wgrad(xd, dyd, g4, atomic=True)
    gw.wgrad(xd, dyd, g5, force_v1=True, atomic=True)
    torch.cuda.synchronize()
    assert torch.equal(g3, g2), 'two-stage weight gradient must be deterministic'
    for ga in (g4, g5):
        assert rel_err(ga[5:5 + w.numel()].view_as(w).double().cpu(), wb.grad) < 2e-4, 'atomic form'
    # accumulation semantics: a second call adds
    gw.wgrad(xd, dyd, g3)
    torch.cuda.synchronize()
    assert rel_err(g3[5:5 + w.numel()].view_as(w).double().cpu(), 2 * wb.grad) < 2e-4


def test_halo_k_concat_block_stage2():
    """Stage-2 GEMM of a residual block: K-concatenation of 1x1 / 3x3 / 5x5 convs over channel slices."""
    from cat_b200 import ops
    torch.manual_seed(2)
    N, H, W, C = 2, 16, 16, 62
    mids, ks = [17, 9, 70, 5], [1, 3, 5, 1]
    xs = [torch.randn(N, m, H, W) for m in mids]
    ws = [torch.randn(C, m, k, k) / math.sqrt(m * k * k) for m, k in zip(mids, ks)]
    ref = sum(F.conv2d(F.pad(bf(x), ((k - 1) // 2,) * 4, mode='reflect') if k > 1 else bf(x), bf(w)) for x, w, k in zip(xs, ws, ks))
    arena = torch.cat([torch.zeros(3)] + [w.flatten() for w in ws]).to(DEV)
    ld = sum(P.cpad(m) for m in mids)
    buf = torch.zeros(N, H, W, ld, dtype=torch.bfloat16)
    units, cu0, off = P.Units(), 0, 3
    for x, w, k, m in zip(xs, ws, ks, mids):
        buf[..., cu0 * 8:cu0 * 8 + m] = x.permute(0, 2, 3, 1).to(torch.bfloat16)
        units.extend(P.conv_fprop_units(off, C, m, k, k, (k - 1) // 2, cu0=cu0))
        cu0 += P.cpad(m) // 8
        off += w.numel()
    g = P.Geometry(N, H, W, ld, 0, H, W, P.cpad(C), 0, pad_mode=P.PAD_REFLECT)
    gm = ops.Gemm(g, units, C, DEV)
    assert gm.halo is not None
    gm.choice = 'v2'
    gm.pack(arena)
    y = torch.zeros(N, H, W, P.cpad(C), dtype=torch.bfloat16, device=DEV)
    gm.fprop(buf.to(DEV), y)
    torch.cuda.synchronize()
    assert rel_err(from_dev_nhwc(y, C), ref) < 6e-3


@pytest.mark.parametrize('force_v1', [False, True])
def test_fused_stage1_wgrad(force_v1):
    """Weight gradients of the first-stage convs of a residual block (kernel sizes 1 / 3 / 5 reading the same input) as
    ONE N-concatenated GEMM on the 5x5 tap grid, unpacked per row segment (cat_b200/engine.py: b.s1w)."""
    from cat_b200 import ops
    torch.manual_seed(5)
    N, H, W, C = 2, 16, 16, 62
    mids, ks = [17, 9, 12, 5], [1, 3, 5, 1]
    x = torch.randn(N, C, H, W)
    ws = [(torch.randn(m, C, k, k) / math.sqrt(C * k * k)) for m, k in zip(mids, ks)]
    dys = [torch.randn(N, m, H, W) for m in mids]
    xb = bf(x)
    refs = []
    for w, k, dy in zip(ws, ks, dys):
        wb = bf(w).requires_grad_(True)
        xin = F.pad(xb, ((k - 1) // 2,) * 4, mode='reflect') if k > 1 else xb
        F.conv2d(xin, wb).backward(bf(dy))
        refs.append(wb.grad)
    L = sum(P.cpad(m) for m in mids)
    dyb = torch.zeros(N, H, W, L, dtype=torch.bfloat16)
    segs, sl, off = [], 0, 7
    for w, k, m, dy in zip(ws, ks, mids, dys):
        dyb[..., sl:sl + m] = dy.permute(0, 2, 3, 1).to(torch.bfloat16)
        segs.append((sl, P.cpad(m), m, P.conv_embedded_units(off, m, C, k, 5)))
        sl += P.cpad(m)
        off += w.numel()
    geo = P.Geometry(N, H, W, P.cpad(C), 0, H, W, L, 0, pad_mode=P.PAD_REFLECT)
    gw = ops.Gemm(geo, P.conv_fprop_units(0, L, C, 5, 5, 2), L, DEV, need_pack=False, segments=segs)
    gw._wgrad_plan()
    assert gw.w_halo is not None
    gw.w_choice = 'v1' if force_v1 else 'v2'
    g = torch.zeros(off + 9, device=DEV)
    gw.wgrad(to_dev_nhwc(x), dyb.to(DEV), g, force_v1=force_v1)
    torch.cuda.synchronize()
    off = 7
    assert float(g[:7].abs().max()) == 0
    for w, ref in zip(ws, refs):
        got = g[off:off + w.numel()].view_as(w).double().cpu()
        assert rel_err(got, ref) < 2e-4
        off += w.numel()
    assert float(g[off:].abs().max()) == 0
