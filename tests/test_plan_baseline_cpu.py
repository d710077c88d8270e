"""Kernel plans of every GEMM of the published benchmark networks at BASELINE.json's resolution (host side only: unit tables,
halo plans, shared-memory / tensor-memory / tensor-map-box fits are all decided on the host by cat_b200.ops.Gemm and the
`*_fits` entry points of libcatb200.so, so they can be checked without a GPU).  What must hold for configs[1] / [2]:
every conv, transposed conv and input gradient has a halo plan with at least one tiling; every zero-padded or 1x1 GEMM offers
the persistent kernel with TMA-staged tiles (v3 mode 2) and every reflection-padded one the persistent kernel with cp.async
producers (mode 1); the wide PatchGAN weight gradients qualify for the TMA producers."""
import pytest
import torch

from cat_b200 import _C


@pytest.fixture()
def no_cuda_check(monkeypatch):
    from cat_b200 import ops
    monkeypatch.setattr(ops, 'require_cuda', lambda: None)


@pytest.mark.timeout(600)
@pytest.mark.parametrize('name,H,W', [('pix2pix_5p6B', 256, 256), ('cyclegan_2p6B', 256, 256), ('pix2pix_5p6B', 256, 512)])
def test_every_benchmark_gemm_has_its_kernels(no_cuda_check, name, H, W):
    from cat_b200 import workload as WL
    from cat_b200.engine import DisNet, GenNet
    arch = WL.load_arch(name)
    B = 2          # plans depend on the batch only through the tile count
    nets = {'student': GenNet(arch['student_arch'], B, H, W, 'cpu', training=True, need_grad=True),
            'teacher': GenNet(arch['teacher_arch'], B, H, W, 'cpu', training=False, need_grad=False),
            'D': DisNet(arch['D_arch'], B, H, W, 'cpu')}
    n_tma = n_all = 0
    for tag, net in nets.items():
        for g in net.fprop_gemms + net.bwd_gemms:
            n_all += 1
            assert g.halo is not None and g.tilings, (tag, g.n_rows, g.n_units)
            modes = {t[4] for t in g.tilings}
            plan = g.halo
            border = plan.Ymax or plan.Xmax or any(pl[2] or pl[3] for pl in plan.planes)
            if g.geo.pad_mode == _C.PAD_REFLECT and border:
                assert 1 in modes and 2 not in modes, (tag, g.n_rows, g.n_units, modes)
            else:
                assert 2 in modes, (tag, g.n_rows, g.n_units, modes)      # TMA-staged tiles
                n_tma += 1
                assert g.c_visible % 8 == 0 and g.geo.x_coff + g.c_visible <= g.geo.ldx
            for (tw, ms, hd, _st, mode) in g.tilings:
                if mode:      # two accumulator stages in tensor memory
                    assert 2 * ms * g.n_tile <= 512
    assert n_tma >= n_all // 2
    # PatchGAN weight gradients: layers 2-4 (zero padding, one strip at this width) take the TMA producers
    D = nets['D']
    tma_w = []
    for L in D.layers:
        L.gw._wgrad_plan()
        assert L.gw.w_halo is not None, L.wn
        tma_w.append(bool(L.gw.w_tma_ok))
    assert tma_w[-1] and sum(tma_w) >= len(tma_w) - 2, tma_w
    assert D.layers[-1].tap            # one output channel: tap-split form


@pytest.mark.timeout(600)
def test_every_spade_benchmark_gemm_has_its_kernels(no_cuda_check):
    """configs[3]: the GauGAN / SPADE student and teacher at 512 x 512 (every conv of the SPADE path is zero padded, so every
    halo GEMM must offer the TMA-staged persistent kernel)."""
    from cat_b200 import workload as WL
    from cat_b200.ops import Act
    from cat_b200.spade_engine import SpadeGenNet
    sarch = WL.spade_arch_for(WL.load_arch('gaugan_5p6B'), 512, 512)
    seg = Act.empty(1, 512, 512, sarch['student_arch']['semantic_nc'], 'cpu', zero=True)
    for tag, training in (('student_arch', True), ('teacher_arch', False)):
        net = SpadeGenNet(sarch[tag], seg, 'cpu', training=training, need_grad=training)
        gemms = net.fprop_gemms + net.bwd_gemms
        assert len(gemms) > 50
        n_halo = 0
        for g in gemms:
            if g.halo is None:
                continue
            n_halo += 1
            modes = {t[4] for t in g.tilings}
            assert 2 in modes or 1 in modes, (tag, g.n_rows, g.n_units, modes)
            assert g.geo.pad_mode != _C.PAD_REFLECT
        assert n_halo >= 0.9 * len(gemms), (tag, n_halo, len(gemms))
