"""GPU parity tests of the SPADE-path kernels (cat_b200/csrc/spade.cu, zero-padded depthwise convs) against their
torch restatements (oracle/kernel_emu.py, test infrastructure) on the same bf16 inputs.  Element-wise kernels must
agree to one bf16 rounding of the output; integer / index work (one-hot, edges, resize, max-pool routing) bit-exact."""
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]
DEV = 'cuda:0'


@pytest.fixture(scope='module', autouse=True)
def _init():
    from cat_b200 import ops
    ops.require_cuda()


def mk(N, H, W, C, ld=None, coff=0, seed=0, scale=1.0):
    """A random NHWC bf16 activation slice on the GPU and its CPU twin."""
    from cat_b200.ops import Act
    g = torch.Generator().manual_seed(seed)
    ld = ld or C
    t = (torch.randn(N, H, W, ld, generator=g) * scale).to(torch.bfloat16)
    return Act(t.to(DEV), coff, C), Act(t.clone(), coff, C)


def same(a, b, tol=1e-2, exact=False):
    x, y = a.t.float().cpu(), b.t.float()
    if exact:
        assert torch.equal(x, y)
    else:
        err = float((x - y).abs().max())
        assert err <= tol * max(1.0, float(y.abs().max())), err


@pytest.mark.parametrize('H,W,OH,OW', [(8, 16, 1, 2), (8, 16, 4, 8), (4, 6, 8, 12), (8, 16, 8, 16), (9, 15, 3, 5)])
def test_resize_nearest_and_upsample_adjoint(H, W, OH, OW):
    from cat_b200 import ops
    from oracle import kernel_emu as E
    xg, xc = mk(3, H, W, 16, ld=24, coff=8, seed=1)
    yg, yc = mk(3, OH, OW, 16, ld=16, seed=2)
    ops.resize_nearest(xg, yg)
    E.resize_nearest(xc, yc)
    same(yg, yc, exact=True)
    if (OH, OW) == (2 * H, 2 * W):
        dg, dc = mk(3, H, W, 16, seed=3)
        ops.upsample2x_bwd(yg, dg)
        E.upsample2x_bwd(yc, dc)
        same(dg, dc)


def test_spade_modulate_forward_backward():
    from cat_b200 import ops, _C
    from oracle import kernel_emu as E
    N, H, W, C = 2, 7, 9, 24
    xg, xc = mk(N, H, W, C, seed=1)
    gbg, gbc = mk(N, H, W, 2 * C, seed=2, scale=0.5)
    yg, yc = mk(N, H, W, C, seed=3)
    g = torch.Generator().manual_seed(4)
    scale, shift = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)
    for act in (_C.ACT_RELU, _C.ACT_NONE):
        ops.spade_modulate(xg, gbg.slice(0, C), gbg.slice(C, C), yg, scale.to(DEV), shift.to(DEV), act)
        E.spade_modulate(xc, gbc.slice(0, C), gbc.slice(C, C), yc, scale, shift, act)
        same(yg, yc)
        yg.t.copy_(yc.t)      # identical activation masks for the backward comparison
        dyg, dyc = mk(N, H, W, C, seed=5)
        dgbg, dgbc = mk(N, H, W, 2 * C, seed=6)
        dng, dnc = mk(N, H, W, C, seed=7)
        ops.spade_modulate_bwd(dyg, yg, xg, gbg.slice(0, C), dgbg.slice(0, C), dgbg.slice(C, C), dng, scale.to(DEV), shift.to(DEV), act)
        E.spade_modulate_bwd(dyc, yc, xc, gbc.slice(0, C), dgbc.slice(0, C), dgbc.slice(C, C), dnc, scale, shift, act)
        same(dgbg, dgbc)
        same(dng, dnc)


def test_act_fwd_and_pools():
    from cat_b200 import ops, _C
    from oracle import kernel_emu as E
    for (H, W) in ((8, 12), (9, 13)):
        xg, xc = mk(2, H, W, 40, seed=H)
        yg, yc = mk(2, H, W, 40, seed=1)
        ops.act_fwd(xg, yg, _C.ACT_LEAKY02)
        E.act_fwd(xc, yc, _C.ACT_LEAKY02)
        same(yg, yc)
        OH, OW = (H + 1) // 2, (W + 1) // 2
        pg, pc = mk(2, OH, OW, 40, seed=2)
        ops.avgpool3s2(xg, pg)
        E.avgpool3s2(xc, pc)
        same(pg, pc)
        dg, dc = mk(2, H, W, 40, seed=3)
        ops.avgpool3s2_bwd(pg, dg, add=dg)
        E.avgpool3s2_bwd(pc, dc, add=dc)
        same(dg, dc)
    xg, xc = mk(2, 8, 12, 16, seed=5)
    xg.t[0, :2, :2, 0] = 0          # a tie: the gradient goes to the first element of the window
    xc.t[0, :2, :2, 0] = 0
    mg, mc = mk(2, 4, 6, 16, seed=6)
    ops.maxpool2(xg, mg)
    E.maxpool2(xc, mc)
    same(mg, mc, exact=True)
    dyg, dyc = mk(2, 4, 6, 16, seed=7)
    dxg, dxc = mk(2, 8, 12, 16, seed=8)
    ops.maxpool2_bwd(dyg, xg, dxg)
    E.maxpool2_bwd(dyc, xc, dxc)
    same(dxg, dxc, exact=True)


def test_onehot_edges_bit_exact():
    from cat_b200 import ops
    from cat_b200.ops import Act
    from oracle import kernel_emu as E
    from oracle import spade_oracle as SO
    g = torch.Generator().manual_seed(0)
    N, H, W, nl = 2, 16, 24, 35
    lab = torch.randint(0, nl, (N, 1, H // 4, W // 4), generator=g).repeat_interleave(4, 2).repeat_interleave(4, 3)
    inst = torch.randint(0, 5, (N, 1, H // 2, W // 2), generator=g).repeat_interleave(2, 2).repeat_interleave(2, 3)
    ref = SO.preprocess_input(lab.float(), inst, nl)                    # [N, nl+1, H, W]
    yg = Act.empty(N, H, W, nl + 1, DEV)
    yg.t.fill_(7.0)                                                      # stale data must be overwritten
    ops.onehot_edges(lab.reshape(N, H, W).int().to(DEV), inst.reshape(N, H, W).int().to(DEV), nl, yg)
    got = yg.t.float().cpu()
    assert torch.equal(got[..., :nl + 1].permute(0, 3, 1, 2), ref)
    assert float(got[..., nl + 1:].abs().max()) == 0.0


def test_bias_vector_helpers():
    from cat_b200 import ops
    from oracle import kernel_emu as E
    g = torch.Generator().manual_seed(0)
    arena = torch.randn(1000, generator=g)
    idx = torch.randint(-1, 1000, (3, 50), generator=g).int()
    out_c, out_g = torch.zeros(56), torch.zeros(56, device=DEV)
    E.gather_sum(arena, idx, out_c)
    ops.gather_sum(arena.to(DEV), idx.to(DEV), out_g)
    assert float((out_g.cpu() - out_c).abs().max()) < 1e-6
    src = torch.randn(50, generator=g)
    ga_c, ga_g = torch.zeros(1000), torch.zeros(1000, device=DEV)
    E.scatter_add(src, idx, ga_c)
    ops.scatter_add(src.to(DEV), idx.to(DEV), ga_g)
    assert float((ga_g.cpu() - ga_c).abs().max()) < 1e-5
    sh = torch.randn(40, generator=g)
    b, sc = torch.randn(40, generator=g), torch.randn(40, generator=g)
    sh_g = sh.to(DEV)
    ops.fma_vec(sh_g, b.to(DEV), sc.to(DEV))
    assert float((sh_g.cpu() - (sh + b * sc)).abs().max()) < 1e-6


@pytest.mark.parametrize('training', [True, False])
def test_spectral_norm_forward_backward(training):
    import numpy as np
    from cat_b200 import ops
    from oracle import kernel_emu as E
    g = torch.Generator().manual_seed(1)
    shapes = [(16, 128), (32, 256), (64, 512), (512, 4096)]
    arena, bufs, rows = [torch.zeros(8)], [torch.zeros(8)], []
    wo, bo = 8, 8
    for (r, c) in shapes:
        arena.append(torch.randn(r * c, generator=g) * 0.05)
        u = torch.nn.functional.normalize(torch.randn(r, generator=g), dim=0)
        v = torch.nn.functional.normalize(torch.randn(c, generator=g), dim=0)
        bufs += [u, v]
        rows.append([wo, r, c, bo, bo + r, 0])
        wo += r * c
        bo += r + c
    arena, bufs = torch.cat(arena), torch.cat(bufs)
    tab = torch.from_numpy(np.array(rows, dtype=np.int32))
    n, mr, mc = len(rows), 512, 4096
    res = []
    for dev, mod in ((DEV, ops), ('cpu', E)):
        a, b = arena.clone().to(dev), bufs.clone().to(dev)
        w_eff = a.clone()
        tmp = torch.zeros(n * max(mr, mc), device=dev)
        sigma, cdot = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
        mod.sn_forward(tab.to(dev), n, mr, mc, a, b, training, tmp, sigma, w_eff)
        grad = torch.randn(arena.numel(), generator=torch.Generator().manual_seed(2)).to(dev)
        mod.sn_backward(tab.to(dev), n, mr, mc, grad, w_eff, b, sigma, cdot)
        res.append([t.cpu() for t in (b, sigma, w_eff, grad)])
    for name, x, y in zip(('u/v', 'sigma', 'w_eff', 'grad'), res[0], res[1]):
        err = float((x - y).abs().max() / y.abs().max())
        assert err < 1e-4, (name, err)


@pytest.mark.parametrize('k', [1, 3, 5])
def test_dwconv_zero_padding(k):
    from cat_b200 import ops, _C
    from oracle import kernel_emu as E
    N, H, W, C = 2, 9, 11, 16
    g = torch.Generator().manual_seed(k)
    arena = torch.randn(8 + C * k * k, generator=g) * 0.3
    ksz = torch.full((C,), k, dtype=torch.int32)
    wof = (8 + torch.arange(C) * k * k).int()
    wof[13:] = -1                                    # padding channels
    xg, xc = mk(N, H, W, C, seed=1)
    yg, yc = mk(N, H, W, C, seed=2)
    ops.dwconv_fwd(xg, yg, ksz.to(DEV), wof.to(DEV), arena.to(DEV), _C.PAD_ZERO)
    E.dwconv_fwd(xc, yc, ksz, wof, arena, _C.PAD_ZERO)
    same(yg, yc)
    dxg, dxc = mk(N, H, W, C, seed=3)
    ops.dwconv_bwd_data(yg, dxg, ksz.to(DEV), wof.to(DEV), arena.to(DEV), _C.PAD_ZERO)
    E.dwconv_bwd_data(yc, dxc, ksz, wof, arena, _C.PAD_ZERO)
    same(dxg, dxc)
    ga_g, ga_c = torch.zeros(arena.numel(), device=DEV), torch.zeros(arena.numel())
    ops.dwconv_bwd_weight(xg, yg, ksz.to(DEV), wof.to(DEV), ga_g, _C.PAD_ZERO)
    E.dwconv_bwd_weight(xc, yc, ksz, wof, ga_c, _C.PAD_ZERO)
    assert float((ga_g.cpu() - ga_c).abs().max()) <= 2e-3 * max(1.0, float(ga_c.abs().max()))


def test_xpack_helpers():
    """catb_expand_x / catb_shift_sum / catb_shift_expand (x-packed 7x7 stem / head) against their restatements."""
    from cat_b200 import ops, _C
    from oracle import kernel_emu as E
    N, H, W = 2, 9, 13
    xg, xc = mk(N, H, W, 8, seed=1)
    yg, yc = mk(N, H, W, 24, seed=2)
    ops.expand_x(xg, yg, 3, 7)
    E.expand_x(xc, yc, 3, 7)
    same(yg, yc, exact=True)
    Pg, Pc = mk(N, H, W + 6, 24, seed=3)
    og, oc = mk(N, H, W, 8, seed=4)
    bias = torch.tensor([0.1, -0.2, 0.3])
    ops.shift_sum(Pg, og, 3, 7, bias.to(DEV), _C.ACT_TANH)
    E.shift_sum(Pc, oc, 3, 7, bias, _C.ACT_TANH)
    same(og, oc)
    assert float(og.t[..., 3:].float().abs().max()) == 0.0
    dPg, dPc = mk(N, H, W + 6, 24, seed=5)
    ops.shift_expand(og, dPg, 3, 7)
    E.shift_expand(oc, dPc, 3, 7)
    same(dPg, dPc)
