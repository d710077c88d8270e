"""GPU parity of the tap-split form of a one-output-channel conv (PatchGAN head, reference
models/modules/discriminators.py:72-73): 1x1 tap-product GEMM + catb_tap_sum forward, catb_tap_expand + 1x1 weight / input
gradient GEMMs backward, against F.conv2d and autograd in fp64 on bf16-rounded operands."""
import pytest
import torch
import torch.nn.functional as F

from cat_b200 import igemm_plan as P

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
DEV = 'cuda:0'


@pytest.fixture(scope='module', autouse=True)
def _init():
    from cat_b200 import ops
    ops.require_cuda()


def bf(x):
    return x.to(torch.bfloat16).to(torch.float64)


@pytest.mark.parametrize('N,Cin,H,W,pad', [(2, 40, 9, 11, 1), (3, 128, 31, 31, 1), (2, 64, 10, 7, 2)])
def test_tap_split_head_matches_conv2d(N, Cin, H, W, pad):
    from cat_b200 import ops
    from cat_b200.ops import Act, Gemm
    torch.manual_seed(Cin + H)
    R = S = 4
    x = torch.randn(N, Cin, H, W)
    w = torch.randn(1, Cin, R, S) / (Cin * R * S) ** 0.5
    b = torch.randn(1)
    xb, wb = bf(x).requires_grad_(True), bf(w).requires_grad_(True)
    y_ref = F.conv2d(xb, wb, b.double(), padding=pad)
    OH, OW = y_ref.shape[2:]
    dy = torch.randn(N, 1, OH, OW)
    y_ref.backward(bf(dy))
    w_off = 7
    arena = torch.cat([torch.zeros(w_off), w.flatten()]).to(DEV)
    grad = torch.zeros_like(arena)
    Cp = P.cpad(Cin)
    xd = torch.zeros(N, H, W, Cp, dtype=torch.bfloat16, device=DEV)
    xd[..., :Cin] = x.permute(0, 2, 3, 1).to(torch.bfloat16).to(DEV)
    units = P.tap_split_units(w_off, Cin, R, S)
    g = Gemm(P.Geometry(N, H, W, Cp, 0, H, W, 16, 0), units, 16, DEV)
    gb = Gemm(P.Geometry(N, H, W, 16, 0, H, W, Cp, 0), P.tap_split_dgrad_units(w_off, Cin, R, S), Cin, DEV)
    gw = Gemm(P.Geometry(N, H, W, Cp, 0, H, W, 16, 0), units, 16, DEV, need_pack=False)
    g.pack(arena)
    gb.pack(arena)
    Pbuf = torch.zeros(N, H, W, 16, dtype=torch.float32, device=DEV)
    y = torch.zeros(N, OH, OW, 8, dtype=torch.float32, device=DEV)
    g.fprop(xd, Pbuf, y_is_f32=True)
    ops.tap_sum(Pbuf, y, H, W, OH, OW, R, S, pad, b.to(DEV))
    torch.cuda.synchronize()
    got = y[..., 0].double().cpu()
    assert float((got - y_ref.detach()[:, 0]).abs().max()) <= 2e-5 * float(y_ref.abs().max()) + 1e-6
    assert float(y[..., 1:].abs().max()) == 0
    # backward
    dyd = torch.zeros(N, OH, OW, 8, dtype=torch.bfloat16, device=DEV)
    dyd[..., 0] = dy[:, 0].to(torch.bfloat16).to(DEV)
    dP = Act.empty(N, H, W, 16, DEV, zero=True)
    ops.tap_expand(Act(dyd), dP, R, S, pad)
    gw.wgrad(xd, dP.t, grad)
    dx = torch.zeros(N, H, W, Cp, dtype=torch.bfloat16, device=DEV)
    gb.fprop(dP.t, dx)
    torch.cuda.synchronize()
    gwt = grad[w_off:].view(1, Cin, R, S).double().cpu()
    assert float((gwt - wb.grad).abs().max()) <= 2e-4 * float(wb.grad.abs().max())
    gx = dx[..., :Cin].permute(0, 3, 1, 2).double().cpu()
    assert float((gx - xb.grad).abs().max()) <= 6e-3 * float(xb.grad.abs().max())
    assert float(grad[:w_off].abs().max()) == 0
