"""Parity at BASELINE.json's FULL sizes: the published benchmark networks (teacher ngf 64, the student pruned to 5.6 GMAC by
the reference's own shrink(), PatchGAN ndf 128; committed as tests/golden/arch_*.json) at the benchmark resolution, directly
against the fp32 CPU oracle on a two-image batch (the oracle needs seconds there), plus size-independent properties of the
bandwidth-bound kernels at the benchmark's tensor sizes (K_T = 256 x 64 x 64 = 1 048 576 per sample):

  * KA(X, X) = 1, KA(cX, Y) = KA(X, Y), 0 <= KA <= 1, and <dKA/dX, X> = 0 (Euler's identity for a scale-invariant function);
  * InstanceNorm / BatchNorm output statistics (zero mean, unit variance per channel) of the fused statistics + apply kernels;
  * linearity of the implicit-GEMM convolution, conv(x1 + x2) = conv(x1) + conv(x2), on the widest PatchGAN layer.

Written after the round-1 GPU budget was spent (not yet run on a B200); CATB_FULLSIZE_HW=<pixels> shrinks the resolution for
the CPU run of the same bodies under the kernel emulation (tests/test_train_bf16_emulated_cpu.py)."""
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(1800)]

DEV = ['cuda:0']


def _hw():
    return int(os.environ.get('CATB_FULLSIZE_HW', '256'))


def _sync():
    if DEV[0] != 'cpu':
        torch.cuda.synchronize()


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def test_pix2pix_5p6B_step_at_benchmark_resolution():
    """BASELINE configs[1] networks at 256x256, two images: one distillation step against the fp32 oracle."""
    from cat_b200 import ops
    from cat_b200 import workload as WL
    from cat_b200.distill_engine import DistillStep
    from oracle import cat_oracle as O
    arch = WL.load_arch('pix2pix_5p6B')
    # lr / 20: the first Adam step moves EVERY discriminator weight by +-lr, a coherent shift of the prediction that G_gan
    # (evaluated after optimizer_D.step) sees; the sign of the many rounding-level gradient elements is arbitrary, so at
    # the reference's lr the bf16 and fp32 steps differ by ~10 % there.  This test is about the kernels at full size.
    hp = dict(arch['hp'], lr=arch['hp']['lr'] / 20)
    B, H, W = 2, _hw(), _hw()
    a, b = WL.synthetic_batch(B, H, W, 233)
    # conv gain 0.1 instead of the reference's N(0, 0.02): random (untrained) weights at 0.02 shrink the activations layer
    # by layer until the losses no longer depend on them
    t_sd = WL.init_generator(arch['teacher_arch'], 0, 'uniform', gain=0.1)
    with torch.no_grad():       # synthetic "trained" teacher: running statistics calibrated on the batch (momentum 1)
        O.generator_forward(t_sd, dict(arch['teacher_arch'], momentum=1.0), a, training=True)
    s_sd = WL.init_generator(arch['student_arch'], 1, gain=0.1)
    d_sd = WL.init_discriminator(arch['D_arch'], 2, gain=0.1)
    st = dict(teacher_sd=O.clone_sd(t_sd), student_sd=O.clone_sd(s_sd), D_sd=O.clone_sd(d_sd), teacher_arch=arch['teacher_arch'],
              student_arch=arch['student_arch'], D_arch=arch['D_arch'], adam_G={}, adam_D={})
    ref = O.distill_step(st, a, b, hp)
    eng = DistillStep(arch['teacher_arch'], arch['student_arch'], arch['D_arch'], hp, B, H, W, device=DEV[0],
                      use_cuda_graph=DEV[0] != 'cpu')
    eng.load(t_sd, s_sd, d_sd)
    eng.set_input(a, b)
    eng.step()
    _sync()
    assert rel_l2(ops.nhwc_to_nchw(eng.T.out, 3).cpu(), ref['Tfake_B']) <= 5e-2
    assert rel_l2(ops.nhwc_to_nchw(eng.S.out, 3).cpu(), ref['Sfake_B']) <= 5e-2
    for n in O.MAPPING_LAYERS:
        Ct, Cs = ref['Tacts'][n].shape[1], ref['Sacts'][n].shape[1]
        assert rel_l2(ops.nhwc_to_nchw(eng.T.acts[n], Ct).cpu(), ref['Tacts'][n]) <= 5e-2, n
        assert rel_l2(ops.nhwc_to_nchw(eng.S.acts[n], Cs).cpu(), ref['Sacts'][n]) <= 5e-2, n
    L = eng.get_losses()
    for k_ref, k in (('loss_D_fake', 'D_fake'), ('loss_D_real', 'D_real'), ('loss_G_gan', 'G_gan'), ('loss_G_recon', 'G_recon'),
                     ('loss_G_distill', 'G_distill')):
        r = float(ref[k_ref])
        assert abs(L[k] - r) <= 2e-2 * max(1.0, abs(r)), (k, L[k], r)
    for i in range(4):
        assert abs(L['G_distill%d' % i] - float(ref['loss_G_distill_terms'][i])) <= 5e-3, i
    for tag, net, grads in (('S', eng.S, ref['S_grads']), ('D', eng.D, ref['D_grads'])):
        ks = [k for k in grads if net.arena.has(k)]
        mine = torch.cat([net.arena.view(k, 'g').flatten().cpu() for k in ks])
        assert rel_l2(mine, torch.cat([grads[k].flatten() for k in ks])) <= 0.5, tag


def test_ka_invariances_at_benchmark_size():
    from cat_b200 import ops
    from cat_b200.ops import Act
    dev = DEV[0]
    B, C, hw = 16, 256, _hw() // 4
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, C, hw, hw, generator=g)
    y = torch.randn(B, 62, hw, hw, generator=g) + 0.3 * x[:, :62]

    def act_of(t):
        a = Act.empty(t.shape[0], hw, hw, t.shape[1], dev, zero=True)
        ops.nchw_to_nhwc(t.to(dev), a)
        return a

    def ka(xa, ya, want_grad=False):
        f32 = dict(dtype=torch.float32, device=dev)
        Gx, Gy, coef = torch.zeros(B, B, **f32), torch.zeros(B, B, **f32), torch.zeros(B, B, **f32)
        loss, val = torch.zeros(1, **f32), torch.zeros(1, **f32)
        ops.gram(xa, Gx)
        ops.gram(ya, Gy)
        ops.ka_finish(Gx, Gy, B, 1.0, loss, val, coef)
        if not want_grad:
            return float(val)
        dx = Act.empty(B, hw, hw, xa.C, dev, zero=True)
        ops.ka_bwd(xa, coef, dx, False)
        return float(val), dx

    xa, ya = act_of(x), act_of(y)
    assert abs(ka(xa, xa) - 1.0) <= 1e-4                                   # KA(X, X) = 1
    v = ka(ya, xa)
    assert 0.0 <= v <= 1.0 + 1e-6
    assert abs(ka(act_of(4.0 * y), xa) - v) <= 2e-3                        # scale invariance (bf16 storage of 4 y is exact)
    assert abs(ka(xa, ya) - v) <= 1e-5                                     # symmetry
    v2, dx = ka(ya, xa, want_grad=True)
    inner = float((dx.t.float() * ya.t.float()).sum())
    norm = float(dx.t.float().norm() * ya.t.float().norm())
    assert abs(inner) <= 2e-2 * norm, (inner, norm)                        # <dKA/dX, X> = 0


@pytest.mark.parametrize('per_sample', [True, False])
def test_norm_statistics_at_benchmark_size(per_sample):
    from cat_b200 import ops
    from cat_b200.ops import ACT, Act
    dev = DEV[0]
    B, C, hw = 16, 256, _hw() // 4
    g = torch.Generator().manual_seed(6)
    x = 3.0 * torch.randn(B, C, hw, hw, generator=g) + torch.linspace(-2, 2, C).view(1, C, 1, 1)
    xa, ya = Act.empty(B, hw, hw, C, dev, zero=True), Act.empty(B, hw, hw, C, dev, zero=True)
    ops.nchw_to_nhwc(x.to(dev), xa)
    G = B if per_sample else 1
    f32 = dict(dtype=torch.float32, device=dev)
    sums, scale, shift, mr = torch.zeros(G, 2, C, **f32), torch.empty(G, C, **f32), torch.empty(G, C, **f32), torch.empty(G, 2, C, **f32)
    ops.norm_stats(xa, per_sample, sums)
    ops.norm_finalize(sums, G, C, hw * hw if per_sample else B * hw * hw, 1e-5, 0.1, None, None, None, None, scale, shift, mr)
    ops.norm_apply(xa, ya, scale, shift, per_sample, ACT['none'])
    _sync()
    out = ops.nhwc_to_nchw(ya, C).cpu().double()
    dims = (2, 3) if per_sample else (0, 2, 3)
    assert float(out.mean(dims).abs().max()) <= 2e-2
    assert float((out.var(dims, unbiased=False) - 1.0).abs().max()) <= 2e-2


def test_conv_linearity_on_the_widest_patchgan_layer():
    """conv(x1 + x2) = conv(x1) + conv(x2) for the 4x4 512 -> 1024 PatchGAN layer at its benchmark size (no bias)."""
    from cat_b200 import igemm_plan as P
    from cat_b200 import ops
    from cat_b200.ops import Act
    dev = DEV[0]
    B, Cin, Cout, h = 16, 512, 1024, _hw() // 8
    g = torch.Generator().manual_seed(7)
    w = (torch.randn(Cout * Cin * 16, generator=g) * 0.02).to(dev)
    arena = torch.zeros(8 + w.numel() + 64, device=dev)
    arena[8:8 + w.numel()] = w
    geo = P.Geometry(B, h, h, Cin, 0, h - 1, h - 1, Cout, 0)
    gemm = ops.Gemm(geo, P.conv_fprop_units(8, Cout, Cin, 4, 4, 1), Cout, dev)
    gemm.pack(arena)
    # inputs on a coarse grid (multiples of 1/8 in [-2, 2]) so that x1, x2 and x1 + x2 are all exact in bf16
    x1 = torch.randint(-16, 17, (B, h, h, Cin), generator=g).float() / 8
    x2 = torch.randint(-16, 17, (B, h, h, Cin), generator=g).float() / 8
    outs = []
    for x in (x1, x2, x1 + x2):
        xa, ya = Act(x.to(dev).to(ops.BF16).contiguous()), Act.empty(B, h - 1, h - 1, Cout, dev, zero=True)
        gemm.fprop(xa.t, ya.t)
        _sync()
        outs.append(ya.t.float().cpu())
    assert rel_l2(outs[0] + outs[1], outs[2]) <= 1e-2


def test_gaugan_5p6B_step_at_benchmark_resolution():
    """BASELINE configs[3] networks (SPADE teacher ngf 64, the student pruned by the reference's shrink_spade_model,
    multi-scale spectral D, VGG19) at 256x512 -- the resolution the architecture was profiled on -- two images: one SPADE
    distillation step against the fp32 oracle."""
    from cat_b200 import ops
    from cat_b200 import workload as WL
    from cat_b200.spade_distill_engine import SpadeDistillStep
    from oracle import cat_oracle as O
    from oracle import spade_oracle as SO
    H = _hw()
    W = 2 * H
    B = 2
    arch = WL.spade_arch_for(WL.load_arch('gaugan_5p6B'), H, W)
    hp = dict(arch['hp'], lr_G=arch['hp']['lr_G'] / 20, lr_D=arch['hp']['lr_D'] / 20, ka_scale=1.0)   # see the pix2pix test
    t_sd = WL.init_spade_reference_sd(arch['teacher_arch'], 0, 'uniform')
    s_sd = WL.init_spade_reference_sd(arch['student_arch'], 1)
    d_sd = WL.init_multiscale_D_sd(arch['D_arch'], 2)
    vgg = WL.init_vgg(3)
    lab, inst, img = WL.synthetic_spade_batch(B, H, W, hp['n_label'], 233)
    st = dict(teacher_sd=O.clone_sd(t_sd), student_sd=O.clone_sd(s_sd), D_sd=O.clone_sd(d_sd), vgg_sd=vgg,
              teacher_arch=arch['teacher_arch'], student_arch=arch['student_arch'], D_arch=arch['D_arch'], adam_G={}, adam_D={})
    seg = SO.preprocess_input(lab, inst, hp['n_label'])
    ref = SO.spade_distill_step(st, seg, img, hp)
    eng = SpadeDistillStep(arch['teacher_arch'], arch['student_arch'], arch['D_arch'], hp, B, H, W, device=DEV[0],
                           use_cuda_graph=DEV[0] != 'cpu')
    eng.load(t_sd, s_sd, d_sd, vgg)
    eng.set_input(lab, inst, img)
    eng.step()
    _sync()
    assert torch.equal(ops.nhwc_to_nchw(eng.seg, eng.snc).cpu(), seg)
    # Image-level bounds are wide here: these are RANDOM full-width networks (1024-channel SPADE blocks, He-initialised),
    # which amplify bf16 rounding -- the fp32 oracle with bf16 storage emulated at the same points is itself 1.6 % (teacher)
    # / 8.8 % (student after its Adam step) away from the plain fp32 oracle (measured at 64x128), while every loss agrees
    # to 1e-3.  The losses below carry the parity claim; a wrong kernel moves them by O(1).
    assert rel_l2(ops.nhwc_to_nchw(eng.T.out, 3).cpu(), ref['Tfake_B']) <= 8e-2
    for n, t in ref['Tacts'].items():
        assert rel_l2(ops.nhwc_to_nchw(eng.T.acts[n], t.shape[1]).cpu(), t) <= 8e-2, n
    err_S = rel_l2(ops.nhwc_to_nchw(eng.S.out, 3).cpu(), ref['Sfake_B_D'])
    assert err_S <= 0.25, err_S
    L = eng.get_losses()
    for k in ('G_gan', 'G_feat', 'G_vgg', 'G_distill', 'D_fake', 'D_real'):
        r = float(ref['loss_' + k])
        assert abs(L[k] - r) <= 3e-2 * max(1.0, abs(r)), (k, L[k], r)
    for i in range(3):
        assert abs(L['G_distill%d' % i] - float(ref['loss_G_distill_terms'][i])) <= 5e-3, i
    # (no gradient bound here: with random full-width networks and two images the student gradient of the bf16-emulating
    # oracle is itself O(1) away from the fp32 oracle's; gradients are pinned at fixture scale, tests/test_spade_distill_gpu.py)
    assert float(eng.S.arena.g.abs().max()) > 0 and bool(torch.isfinite(eng.S.arena.g).all())
