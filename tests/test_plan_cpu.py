"""Validate the K-unit tables (cat_b200/igemm_plan.py) on CPU: the torch restatement of the device
gather, driven by the tables, must reproduce F.conv2d / F.conv_transpose2d and their gradients."""
import pytest
import torch
import torch.nn.functional as F

from cat_b200 import igemm_plan as P


def to_nhwc(x, ld=None, coff=0):
    N, C, H, W = x.shape
    ld = ld or P.cpad(C)
    out = torch.zeros(N, H, W, ld, dtype=x.dtype)
    out[..., coff:coff + C] = x.permute(0, 2, 3, 1)
    return out


def from_nhwc(y, C, coff=0):
    return y[..., coff:coff + C].permute(0, 3, 1, 2).contiguous()


def make_arena(*tensors):
    offs, flat, o = [], [], 3  # odd start offset on purpose
    for t in tensors:
        offs.append(o)
        flat.append(t.reshape(-1))
        o += t.numel()
    arena = torch.zeros(o, dtype=tensors[0].dtype)
    for off, f in zip(offs, flat):
        arena[off:off + f.numel()] = f
    return arena, offs


@pytest.mark.parametrize('k,stride,pad,mode,Cin,Cout', [
    (3, 1, 1, 'zero', 5, 7), (5, 1, 2, 'reflect', 9, 4), (7, 1, 3, 'reflect', 3, 17), (1, 1, 0, 'zero', 12, 10),
    (3, 2, 1, 'zero', 6, 11), (4, 2, 1, 'zero', 6, 8), (4, 1, 1, 'zero', 8, 1)])
def test_conv_fprop_dgrad_wgrad(k, stride, pad, mode, Cin, Cout):
    torch.manual_seed(0)
    N, H, W = 2, 10, 12
    x = torch.randn(N, Cin, H, W, dtype=torch.float64, requires_grad=True)
    w = torch.randn(Cout, Cin, k, k, dtype=torch.float64, requires_grad=True)
    b = torch.randn(Cout, dtype=torch.float64)
    xin = F.pad(x, (pad,) * 4, mode='reflect') if mode == 'reflect' else x
    y_ref = F.conv2d(xin, w, b, stride=stride, padding=0 if mode == 'reflect' else pad)
    OH, OW = y_ref.shape[2:]
    dy = torch.randn_like(y_ref)
    y_ref.backward(dy)
    arena, (w_off,) = make_arena(w.detach())
    # fprop
    units = P.conv_fprop_units(w_off, Cout, Cin, k, k, pad)
    geo = P.Geometry(N, H, W, P.cpad(Cin), 0, OH, OW, P.cpad(Cout) + 8, 8, sn=stride,
                     pad_mode=P.PAD_REFLECT if mode == 'reflect' else P.PAD_ZERO)
    y = torch.zeros(N, OH, OW, geo.ldy, dtype=torch.float64)
    P.emulate_fprop(geo, units, Cout, to_nhwc(x.detach()), arena, y, bias=b)
    assert torch.allclose(from_nhwc(y, Cout, 8), y_ref.detach(), atol=1e-10)
    # wgrad (same tables, lattice tensor = dY)
    garena = torch.zeros_like(arena)
    geo_w = P.Geometry(N, H, W, P.cpad(Cin), 0, OH, OW, P.cpad(Cout), 0, sn=stride, pad_mode=geo.pad_mode)
    P.emulate_wgrad(geo_w, units, Cout, to_nhwc(x.detach()), to_nhwc(dy), garena)
    assert torch.allclose(garena[w_off:w_off + w.numel()].view_as(w), w.grad, atol=1e-9)
    assert garena[:w_off].abs().max() == 0
    # dgrad
    if mode == 'reflect':
        # gradient w.r.t. the padded frame, then the adjoint of ReflectionPad2d (checked via autograd)
        du = P.conv_dgrad_units(w_off, Cout, Cin, k, k, 0)
        Hp, Wp = H + 2 * pad, W + 2 * pad
        geo_d = P.Geometry(N, OH, OW, P.cpad(Cout), 0, Hp, Wp, P.cpad(Cin), 0)
        dxp = torch.zeros(N, Hp, Wp, P.cpad(Cin), dtype=torch.float64)
        P.emulate_fprop(geo_d, du, Cin, to_nhwc(dy), arena, dxp)
        xx = torch.zeros(N, Cin, H, W, dtype=torch.float64, requires_grad=True)
        F.pad(xx, (pad,) * 4, mode='reflect').backward(from_nhwc(dxp, Cin))
        assert torch.allclose(xx.grad, x.grad, atol=1e-9)
    else:
        du = P.conv_dgrad_units(w_off, Cout, Cin, k, k, pad)
        geo_d = P.Geometry(N, OH, OW, P.cpad(Cout), 0, H, W, P.cpad(Cin), 0, sn=1, sd=stride)
        dx = torch.zeros(N, H, W, P.cpad(Cin), dtype=torch.float64)
        P.emulate_fprop(geo_d, du, Cin, to_nhwc(dy), arena, dx)
        assert torch.allclose(from_nhwc(dx, Cin), x.grad, atol=1e-9)
        if stride == 2:  # phase-decomposed variant: 4 launches with disjoint sub-lattices and tap subsets
            dx2 = torch.zeros_like(dx)
            total_units = 0
            for a in range(2):
                for bb in range(2):
                    ph = du.phase(a, bb)
                    total_units += len(ph)
                    if len(ph) == 0:
                        continue
                    g = P.Geometry(N, OH, OW, P.cpad(Cout), 0, H, W, P.cpad(Cin), 0, sn=1, sd=2, o_step=2, o_ph=a, o_pw=bb)
                    P.emulate_fprop(g, ph, Cin, to_nhwc(dy), arena, dx2)
            assert total_units == len(du)
            assert torch.allclose(from_nhwc(dx2, Cin), x.grad, atol=1e-9)


@pytest.mark.parametrize('Cin,Cout', [(9, 5), (16, 8)])
def test_conv_transpose(Cin, Cout):
    torch.manual_seed(1)
    N, H, W, k, pad = 2, 6, 7, 3, 1
    x = torch.randn(N, Cin, H, W, dtype=torch.float64, requires_grad=True)
    w = torch.randn(Cin, Cout, k, k, dtype=torch.float64, requires_grad=True)
    y_ref = F.conv_transpose2d(x, w, None, stride=2, padding=pad, output_padding=1)
    OH, OW = y_ref.shape[2:]
    assert (OH, OW) == (2 * H, 2 * W)
    dy = torch.randn_like(y_ref)
    y_ref.backward(dy)
    arena, (w_off,) = make_arena(w.detach())
    fu = P.convT_fprop_units(w_off, Cin, Cout, k, k, pad)
    for decomposed in (False, True):
        y = torch.zeros(N, OH, OW, P.cpad(Cout), dtype=torch.float64)
        if decomposed:
            for a in range(2):
                for b in range(2):
                    g = P.Geometry(N, H, W, P.cpad(Cin), 0, OH, OW, P.cpad(Cout), 0, sn=1, sd=2, o_step=2, o_ph=a, o_pw=b)
                    P.emulate_fprop(g, fu.phase(a, b), Cout, to_nhwc(x.detach()), arena, y)
        else:
            g = P.Geometry(N, H, W, P.cpad(Cin), 0, OH, OW, P.cpad(Cout), 0, sn=1, sd=2)
            P.emulate_fprop(g, fu, Cout, to_nhwc(x.detach()), arena, y)
        assert torch.allclose(from_nhwc(y, Cout), y_ref.detach(), atol=1e-10)
    # dgrad + wgrad share the strided-correlation tables
    bu = P.convT_dgrad_units(w_off, Cin, Cout, k, k, pad)
    gd = P.Geometry(N, OH, OW, P.cpad(Cout), 0, H, W, P.cpad(Cin), 0, sn=2, sd=1)
    dx = torch.zeros(N, H, W, P.cpad(Cin), dtype=torch.float64)
    P.emulate_fprop(gd, bu, Cin, to_nhwc(dy), arena, dx)
    assert torch.allclose(from_nhwc(dx, Cin), x.grad, atol=1e-9)
    garena = torch.zeros_like(arena)
    P.emulate_wgrad(gd, bu, Cin, to_nhwc(dy), to_nhwc(x.detach()), garena)
    assert torch.allclose(garena[w_off:w_off + w.numel()].view_as(w), w.grad, atol=1e-9)


def test_k_concatenated_branches():
    """Second-stage convs of an InvertedResidualChannels block as ONE GEMM: K-concatenation of a 1x1,
    a 3x3 and a 5x5 reflect-padded conv reading different channel slices of one buffer."""
    torch.manual_seed(2)
    N, H, W, C = 1, 9, 8, 6
    mids, ks = [3, 9, 5], [1, 3, 5]
    xs = [torch.randn(N, m, H, W, dtype=torch.float64) for m in mids]
    ws = [torch.randn(C, m, k, k, dtype=torch.float64) for m, k in zip(mids, ks)]
    ref = sum(F.conv2d(F.pad(x, ((k - 1) // 2,) * 4, mode='reflect') if k > 1 else x, w) for x, w, k in zip(xs, ws, ks))
    arena, offs = make_arena(*ws)
    ld = sum(P.cpad(m) for m in mids)
    buf = torch.zeros(N, H, W, ld, dtype=torch.float64)
    units, cu0 = P.Units(), 0
    for x, w, k, m, off in zip(xs, ws, ks, mids, offs):
        buf[..., cu0 * 8:cu0 * 8 + m] = x.permute(0, 2, 3, 1)
        units.extend(P.conv_fprop_units(off, C, m, k, k, (k - 1) // 2, cu0=cu0))
        cu0 += P.cpad(m) // 8
    g = P.Geometry(N, H, W, ld, 0, H, W, P.cpad(C), 0, pad_mode=P.PAD_REFLECT)
    y = torch.zeros(N, H, W, P.cpad(C), dtype=torch.float64)
    P.emulate_fprop(g, units, C, buf, arena, y)
    assert torch.allclose(from_nhwc(y, C), ref, atol=1e-10)


def test_choose_n_tile():
    for n in (1, 3, 16, 17, 62, 126, 256, 257, 512, 1024):
        t = P.choose_n_tile(n)
        assert t % 16 == 0 and 16 <= t <= 256
        tiles = (n + t - 1) // t
        assert tiles * t >= n and (tiles - 1) * t < n


@pytest.mark.parametrize('k,stride,pad,mode,Cin,Cout,H,W', [
    (3, 1, 1, 'zero', 5, 7, 10, 12), (5, 1, 2, 'reflect', 70, 4, 12, 9), (7, 1, 3, 'reflect', 3, 17, 9, 11),
    (1, 1, 0, 'zero', 130, 10, 8, 8), (3, 2, 1, 'zero', 6, 11, 10, 12), (4, 2, 1, 'zero', 72, 8, 12, 10),
    (4, 1, 1, 'zero', 8, 1, 7, 9)])
@pytest.mark.parametrize('m_sub', [1, 2, 4])
def test_halo_plan_conv_and_dgrad(k, stride, pad, mode, Cin, Cout, H, W, m_sub):
    torch.manual_seed(0)
    N = 2
    x = torch.randn(N, Cin, H, W, dtype=torch.float64, requires_grad=True)
    w = torch.randn(Cout, Cin, k, k, dtype=torch.float64)
    b = torch.randn(Cout, dtype=torch.float64)
    xin = F.pad(x, (pad,) * 4, mode='reflect') if mode == 'reflect' else x
    y_ref = F.conv2d(xin, w, b, stride=stride, padding=0 if mode == 'reflect' else pad)
    OH, OW = y_ref.shape[2:]
    arena, (w_off,) = make_arena(w)
    units = P.conv_fprop_units(w_off, Cout, Cin, k, k, pad)
    geo = P.Geometry(N, H, W, P.cpad(Cin), 0, OH, OW, P.cpad(Cout), 0, sn=stride,
                     pad_mode=P.PAD_REFLECT if mode == 'reflect' else P.PAD_ZERO)
    plan = P.make_halo_plan(geo, units)
    assert plan is not None
    plan.m_sub = m_sub
    y = torch.zeros(N, OH, OW, geo.ldy, dtype=torch.float64)
    P.emulate_halo_fprop(geo, plan, Cout, to_nhwc(x.detach()), arena, y, bias=b)
    assert torch.allclose(from_nhwc(y, Cout), y_ref.detach(), atol=1e-10)
    for tw in (4, 5):  # vertical strips (wide images whose full-width halo would not fit in shared memory)
        plan.TW = tw
        y = torch.zeros(N, OH, OW, geo.ldy, dtype=torch.float64)
        P.emulate_halo_fprop(geo, plan, Cout, to_nhwc(x.detach()), arena, y, bias=b)
        assert torch.allclose(from_nhwc(y, Cout), y_ref.detach(), atol=1e-10), tw
    plan.TW = plan.OWs
    # the re-ordered table must still drive the v1 gather kernel identically
    y1 = torch.zeros_like(y)
    P.emulate_fprop(geo, plan.units, Cout, to_nhwc(x.detach()), arena, y1, bias=b)
    assert torch.allclose(y1, y, atol=1e-10)
    if mode == 'zero':
        dy = torch.randn_like(y_ref)
        y_ref.backward(dy)
        du = P.conv_dgrad_units(w_off, Cout, Cin, k, k, pad)
        dx = torch.zeros(N, H, W, P.cpad(Cin), dtype=torch.float64)
        if stride == 1:
            g = P.Geometry(N, OH, OW, P.cpad(Cout), 0, H, W, P.cpad(Cin), 0)
            pl = P.make_halo_plan(g, du)
            pl.m_sub = m_sub
            P.emulate_halo_fprop(g, pl, Cin, to_nhwc(dy), arena, dx)
        else:
            for a in range(2):
                for c in range(2):
                    ph = du.phase(a, c)
                    if len(ph) == 0:
                        continue
                    g = P.Geometry(N, OH, OW, P.cpad(Cout), 0, H, W, P.cpad(Cin), 0, sn=1, sd=2, o_step=2, o_ph=a, o_pw=c)
                    pl = P.make_halo_plan(g, ph)
                    assert pl is not None
                    pl.m_sub = m_sub
                    P.emulate_halo_fprop(g, pl, Cin, to_nhwc(dy), arena, dx)
        assert torch.allclose(from_nhwc(dx, Cin), x.grad, atol=1e-9)


def test_halo_plan_k_concat():
    torch.manual_seed(2)
    N, H, W, C = 1, 9, 8, 6
    mids, ks = [3, 9, 70], [1, 3, 5]
    xs = [torch.randn(N, m, H, W, dtype=torch.float64) for m in mids]
    ws = [torch.randn(C, m, k, k, dtype=torch.float64) for m, k in zip(mids, ks)]
    ref = sum(F.conv2d(F.pad(x, ((k - 1) // 2,) * 4, mode='reflect') if k > 1 else x, w) for x, w, k in zip(xs, ws, ks))
    arena, offs = make_arena(*ws)
    ld = sum(P.cpad(m) for m in mids)
    buf = torch.zeros(N, H, W, ld, dtype=torch.float64)
    units, cu0 = P.Units(), 0
    for x, w, k, m, off in zip(xs, ws, ks, mids, offs):
        buf[..., cu0 * 8:cu0 * 8 + m] = x.permute(0, 2, 3, 1)
        units.extend(P.conv_fprop_units(off, C, m, k, k, (k - 1) // 2, cu0=cu0))
        cu0 += P.cpad(m) // 8
    g = P.Geometry(N, H, W, ld, 0, H, W, P.cpad(C), 0, pad_mode=P.PAD_REFLECT)
    plan = P.make_halo_plan(g, units)
    y = torch.zeros(N, H, W, P.cpad(C), dtype=torch.float64)
    P.emulate_halo_fprop(g, plan, C, buf, arena, y)
    assert torch.allclose(from_nhwc(y, C), ref, atol=1e-10)


@pytest.mark.parametrize('Cin,Cout,H,W', [(3, 17, 9, 11), (3, 64, 12, 8)])
def test_stem_input_gradient_frame(Cin, Cout, H, W):
    """GenNet(input_grad=True) (CycleGAN's cycle term back-propagates through a generator's input): gradient of the 7x7
    reflect-padded stem w.r.t. its PADDED input frame, as one GEMM with q = 0 over a (H+6) x (W+6) output lattice and
    n_rows = 3; both the gather-per-tap table and the halo plan, then the real ops.Gemm object (halo tilings must fit)."""
    torch.manual_seed(3)
    N, k, p = 2, 7, 3
    x = torch.randn(N, Cin, H, W, dtype=torch.float64)
    xp = F.pad(x, (p,) * 4, mode='reflect').requires_grad_(True)
    w = torch.randn(Cout, Cin, k, k, dtype=torch.float64)
    y = F.conv2d(xp, w)
    dy = torch.randn_like(y)
    y.backward(dy)
    arena, (w_off,) = make_arena(w)
    units = P.conv_dgrad_units(w_off, Cout, Cin, k, k, 0)
    geo = P.Geometry(N, H, W, P.cpad(Cout), 0, H + 2 * p, W + 2 * p, P.cpad(Cin), 0)
    fr = torch.zeros(N, H + 2 * p, W + 2 * p, P.cpad(Cin), dtype=torch.float64)
    P.emulate_fprop(geo, units, Cin, to_nhwc(dy), arena, fr)
    assert torch.allclose(from_nhwc(fr, Cin), xp.grad, atol=1e-9)
    assert float(fr[..., Cin:].abs().max()) == 0.0           # padding channels stay exactly zero
    plan = P.make_halo_plan(geo, units)
    assert plan is not None
    for m_sub in (1, 2):
        plan.m_sub = m_sub
        fr2 = torch.zeros_like(fr)
        P.emulate_halo_fprop(geo, plan, Cin, to_nhwc(dy), arena, fr2)
        assert torch.allclose(from_nhwc(fr2, Cin), xp.grad, atol=1e-9), m_sub
    from cat_b200 import ops
    g = ops.Gemm(geo, units, Cin, 'cpu')
    assert g.halo is not None and len(g.tilings) > 0


def test_adaptor_gemms():
    """cat_b200/adaptors.py ('mse' distillation loss): netA_i as a biased 1x1 GEMM C_S -> C_T over the mapped student
    activation, and its input gradient C_T -> C_S accumulated into an existing d(activation) buffer."""
    torch.manual_seed(4)
    N, H, W, Cs, Ct = 2, 6, 5, 18, 48
    x = torch.randn(N, Cs, H, W, dtype=torch.float64, requires_grad=True)
    w = torch.randn(Ct, Cs, 1, 1, dtype=torch.float64)
    b = torch.randn(Ct, dtype=torch.float64)
    y_ref = F.conv2d(x, w, b)
    dy = torch.randn_like(y_ref)
    y_ref.backward(dy)
    arena, (w_off,) = make_arena(w)
    fu = P.conv_fprop_units(w_off, Ct, Cs, 1, 1, 0)
    gf = P.Geometry(N, H, W, P.cpad(Cs), 0, H, W, P.cpad(Ct), 0)
    y = torch.zeros(N, H, W, P.cpad(Ct), dtype=torch.float64)
    P.emulate_fprop(gf, fu, Ct, to_nhwc(x.detach()), arena, y, bias=b)
    assert torch.allclose(from_nhwc(y, Ct), y_ref.detach(), atol=1e-10)
    du = P.conv_dgrad_units(w_off, Ct, Cs, 1, 1, 0)
    gb = P.Geometry(N, H, W, P.cpad(Ct), 0, H, W, P.cpad(Cs), 0)
    prev = torch.randn(N, Cs, H, W, dtype=torch.float64)
    for plan in (None, P.make_halo_plan(gb, du)):
        dx = to_nhwc(prev).clone()
        if plan is None:
            P.emulate_fprop(gb, du, Cs, to_nhwc(dy), arena, dx, None, True)
        else:       # the halo emulator has no accumulate flag (the device epilogue has): add the product to the buffer
            tmp = torch.zeros_like(dx)
            P.emulate_halo_fprop(gb, plan, Cs, to_nhwc(dy), arena, tmp)
            dx = dx + tmp
        assert torch.allclose(from_nhwc(dx, Cs), prev + x.grad, atol=1e-9)
    from cat_b200 import ops
    for geo, units, rows in ((gf, fu, Ct), (gb, du, Cs)):
        assert ops.Gemm(geo, units, rows, 'cpu').n_units == len(units)
