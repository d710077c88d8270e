"""Host logic of the SPADE distillation engine on CPU: cat_b200.spade_distill_engine.SpadeDistillStep is executed
with every kernel wrapper swapped for its torch restatement (oracle/kernel_emu.py, test infrastructure) and
compared with the oracle (which is pinned to the real reference by tests/test_spade_oracle_golden.py).

exact mode keeps the emulated buffers in fp32, so launch order, hand-derived backward passes (SPADE modulation,
six-branch bodies, learned shortcuts, up-sampling, spectral norm, feature matching, VGG, KA) and buffer plumbing
must reproduce the fp32 oracle to rounding; bf16 mode mirrors the device storage and is compared with the
bf16-emulating oracle at the tolerances of the GPU suite."""
import os

import pytest
import torch


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def run_case(golden_dir, exact):
    from oracle import cat_oracle as O
    from oracle import spade_oracle as SO
    from oracle.kernel_emu import emulated_kernels
    from cat_b200 import ops
    fix = torch.load(os.path.join(golden_dir, 'spade_more.pt'), weights_only=False)
    vgg = SO.make_vgg_sd(fix['vgg_seed'])
    s, hp = fix['steps'][0], fix['hp']
    st = dict(teacher_sd=O.clone_sd(fix['teacher_sd']), student_sd=O.clone_sd(fix['student_sd0']), D_sd=O.clone_sd(fix['D_sd0']),
              vgg_sd=vgg, teacher_arch=fix['teacher_arch'], student_arch=fix['student_arch'], D_arch=fix['D_arch'],
              adam_G={}, adam_D={})
    seg = SO.preprocess_input(s['label'], s['instance'], hp['n_label'])
    if exact:
        ref = SO.spade_distill_step(st, seg, s['image'], hp)
    else:
        with O.emulate_bf16():
            ref = SO.spade_distill_step(st, seg, s['image'], hp)
    B, _, H, W = s['image'].shape
    out = {}
    with emulated_kernels(exact=exact):
        from cat_b200.spade_distill_engine import SpadeDistillStep
        eng = SpadeDistillStep(fix['teacher_arch'], fix['student_arch'], fix['D_arch'], hp, B, H, W, device='cpu')
        eng.load(fix['teacher_sd'], fix['student_sd0'], fix['D_sd0'], vgg)
        eng.set_input(s['label'], s['instance'], s['image'])
        eng.step()
        out['seg'] = ops.nhwc_to_nchw(eng.seg, eng.snc)
        out['T'] = ops.nhwc_to_nchw(eng.T.out, 3)
        out['S_D'] = ops.nhwc_to_nchw(eng.S.out, 3)
        out['Tacts'] = {n: ops.nhwc_to_nchw(eng.T.acts[n], ref['Tacts'][n].shape[1]) for n in ref['Tacts']}
        out['losses'] = eng.get_losses()
        out['grads'] = {}
        for tag, net, grads in (('S', eng.S, ref['S_grads']), ('D', eng.D, ref['D_grads'])):
            assert all(net.arena.has(k) for k in grads), [k for k in grads if not net.arena.has(k)]
            out['grads'][tag] = {k: net.arena.view(k, 'g').clone() for k in grads}
        out['S_sd'], out['D_sd'] = eng.S.state_dict(), eng.D.state_dict()
    return fix, st, seg, ref, out


LOSSES = (('loss_D_fake', 'D_fake'), ('loss_D_real', 'D_real'), ('loss_G_gan', 'G_gan'), ('loss_G_feat', 'G_feat'),
          ('loss_G_vgg', 'G_vgg'), ('loss_G_distill', 'G_distill'))


@pytest.mark.timeout(900)
def test_spade_step_host_logic_exact(golden_dir):
    fix, st, seg, ref, out = run_case(golden_dir, exact=True)
    assert torch.equal(out['seg'], seg)                              # one-hot + edges: bit exact
    assert rel_l2(out['T'], ref['Tfake_B']) < 1e-5
    assert rel_l2(out['S_D'], ref['Sfake_B_D']) < 1e-4
    for n, t in ref['Tacts'].items():
        assert rel_l2(out['Tacts'][n], t) < 1e-5, n
    for k_ref, k in LOSSES:
        r = float(ref[k_ref])
        assert abs(out['losses'][k] - r) <= 1e-5 * max(1.0, abs(r)), (k, out['losses'][k], r)
    for i in range(3):
        assert abs(out['losses']['G_distill%d' % i] - float(ref['loss_G_distill_terms'][i])) < 1e-5
    for tag, key in (('S', 'S_grads'), ('D', 'D_grads')):
        scale = max(float(g.abs().max()) for g in ref[key].values())
        for k, g in ref[key].items():
            err = float((out['grads'][tag][k] - g).abs().max())
            # biases in front of a BatchNorm: analytically zero, noise in the oracle, exactly zero here
            assert err <= 2e-3 * float(g.abs().max()) + 2e-5 * scale, (tag, k, err, float(g.abs().max()))
    # running statistics after the two student forwards of the step, spectral-norm vectors after two iterations
    for k, v in st['student_sd'].items():
        if 'running_' in k:
            assert float((out['S_sd'][k] - v).abs().max()) < 1e-4, k
    for k, v in st['D_sd'].items():
        if k.endswith('weight_u') or k.endswith('weight_v'):
            assert float((out['D_sd'][k] - v).abs().max()) < 1e-5, k


@pytest.mark.slow      # second numerical mode of the host logic the exact test pins
@pytest.mark.timeout(900)
def test_spade_step_host_logic_bf16(golden_dir):
    fix, st, seg, ref, out = run_case(golden_dir, exact=False)
    assert torch.equal(out['seg'], seg)
    assert rel_l2(out['T'], ref['Tfake_B']) < 3e-2
    assert rel_l2(out['S_D'], ref['Sfake_B_D']) < 5e-2
    for k_ref, k in LOSSES:
        r = float(ref[k_ref])
        assert abs(out['losses'][k] - r) <= 2e-2 * max(1.0, abs(r)), (k, out['losses'][k], r)
    for tag, key in (('S', 'S_grads'), ('D', 'D_grads')):
        ks = list(ref[key])
        mine = torch.cat([out['grads'][tag][k].flatten() for k in ks])
        theirs = torch.cat([ref[key][k].flatten() for k in ks])
        assert rel_l2(mine, theirs) < 0.35, tag


@pytest.mark.timeout(900)
def test_spade_generator_inference_through_the_module_mirror(golden_dir):
    """InceptionSPADEGenerator.forward (evaluate_model's generator inference) in exact emulation: reference
    checkpoints loaded into the module tree, eval-mode forward for two batch shapes (the second compilation shares the
    first one's arenas), mapped activations returned like the reference forward."""
    import argparse
    from oracle import spade_oracle as SO
    from oracle.cat_oracle import clone_sd
    from oracle.kernel_emu import emulated_kernels
    from cat_b200.models.spade_networks import InceptionSPADEGenerator
    fix = torch.load(os.path.join(golden_dir, 'spade_more.pt'), weights_only=False)
    Sa, hp, s = fix['student_arch'], fix['hp'], fix['steps'][0]
    opt = argparse.Namespace(ngf=Sa['fc_out'] // 16, norm_G='spadesyncbatch3x3', semantic_nc=Sa['semantic_nc'],
                             num_upsampling_layers=Sa['num_upsampling_layers'], crop_size=128, aspect_ratio=2.0, channels=None,
                             channels_reduction_factor=6, kernel_sizes=[1, 3, 5], active_fn='nn.ReLU')
    seg = SO.preprocess_input(s['label'], s['instance'], hp['n_label'])
    caps = {}
    ref = SO.spade_generator_forward(clone_sd(fix['student_sd0']), Sa, seg, training=False, capture=caps)
    with emulated_kernels(exact=True):
        net = InceptionSPADEGenerator.from_arch(Sa, opt)
        net.load_state_dict(fix['student_sd0'])
        net.eval()
        out, acts = net(seg, mapping_layers=['head_0', 'up_1'])
        assert rel_l2(out, ref) < 1e-5
        assert rel_l2(acts['up_1'], caps['up_1']) < 1e-5 and rel_l2(acts['head_0'], caps['head_0']) < 1e-5
        out1 = net(seg[:1])                      # second shape: compiled against the same arenas
        assert rel_l2(out1, ref[:1]) < 1e-5
        assert len(net.__dict__['_engines']) == 2
        e = list(net.__dict__['_engines'].values())
        assert e[0].arena is e[1].arena
        # parameters of the module tree alias the engine arena
        assert net.conv_img.weight.data_ptr() == e[0].arena.view('conv_img.weight').data_ptr()


def _toy_spade_arch(mode, blocks):
    names = ['head_0', 'G_middle_0', 'G_middle_1', 'up_0', 'up_1', 'up_2', 'up_3'] + (['up_4'] if mode == 'most' else [])
    n_up = {'normal': 5, 'more': 6, 'most': 7}[mode]
    return {'semantic_nc': 5, 'fc_out': blocks['head_0']['fin'], 'sh': 1, 'sw': 2, 'num_upsampling_layers': mode,
            'kernel_sizes': [1, 3, 5], 'final_nc': blocks[names[-1]]['fout'], 'block_names': names,
            'blocks': {n: blocks[n] for n in names}, 'eps': 1e-5, 'momentum': 0.1}, (1 << n_up, 2 << n_up)


def _blk(fin, fout, res, dw, sres, sdw):
    return {'fin': fin, 'fout': fout, 'res': res, 'dw': dw, 'spade_res': sres, 'spade_dw': sdw, 'learned_shortcut': fin != fout}


@pytest.mark.timeout(900)
@pytest.mark.parametrize('mode,seed', [('normal', 3), ('most', 5)])
def test_spade_generator_edge_architectures_exact(mode, seed):
    """Generator-only forward + backward in exact emulation on hand-made pruned architectures that exercise the corner
    cases of shrink_spade_model's output: 'normal' / 'most' up-sampling, a block without any branch (identity and learned
    shortcut), a block whose SPADE body lost every branch (gamma = beta = 0), zero-width branches in the middle of the
    kernel-size list, widths that are not multiples of 8.  (The seeds are ones for which no LeakyReLU / ReLU input sits
    within rounding distance of zero at a pixel with a large gradient: one such sign flip is a 1e-2 gradient difference
    between ANY two fp32 evaluations of these tiny, badly conditioned networks.)"""
    from oracle import spade_oracle as SO
    from oracle.kernel_emu import emulated_kernels
    from cat_b200 import ops
    from cat_b200 import workload as WL
    from cat_b200.ops import Act
    from cat_b200.spade_engine import SpadeGenNet
    blocks = {
        'head_0': _blk(24, 24, [3, 0, 2], [0, 4, 0], [2, 0, 3], [0, 0, 2]),
        'G_middle_0': _blk(24, 24, [0, 0, 0], [0, 0, 0], [2, 2, 2], [2, 2, 2]),      # no branch, identity shortcut
        'G_middle_1': _blk(24, 24, [2, 2, 0], [3, 0, 0], [0, 0, 0], [0, 0, 0]),      # SPADE body empty
        'up_0': _blk(24, 12, [0, 0, 0], [0, 0, 0], [1, 1, 1], [1, 1, 1]),            # no branch, learned shortcut
        'up_1': _blk(12, 9, [1, 2, 3], [2, 1, 1], [3, 3, 3], [2, 2, 2]),
        'up_2': _blk(9, 7, [0, 3, 0], [0, 0, 5], [4, 0, 0], [0, 3, 0]),
        'up_3': _blk(7, 5, [2, 0, 0], [0, 2, 0], [0, 2, 0], [2, 0, 0]),
        'up_4': _blk(5, 3, [1, 1, 1], [1, 1, 1], [2, 2, 2], [1, 1, 1]),
    }
    arch, (H, W) = _toy_spade_arch(mode, blocks)
    B = 2
    sd = WL.init_spade_reference_sd(arch, 11, 'uniform')
    g = torch.Generator().manual_seed(seed)
    for k in sd:
        if k.endswith('.bias'):
            sd[k] = 0.1 * torch.randn(sd[k].shape, generator=g)
    lab = torch.randint(0, 4, (B, 1, H // 4, W // 4), generator=g).repeat_interleave(4, 2).repeat_interleave(4, 3)
    inst = torch.randint(0, 3, (B, 1, H // 2, W // 2), generator=g).repeat_interleave(2, 2).repeat_interleave(2, 3)
    seg = SO.preprocess_input(lab, inst, 4)
    R = torch.randn(B, 3, H, W, generator=g)
    # oracle: training-mode forward, gradient of <out, R>
    osd = {k: v.clone() for k, v in sd.items()}
    params = {k: v.requires_grad_(True) for k, v in osd.items() if SO._is_param(k)}
    out_ref = SO.spade_generator_forward(osd, arch, seg, training=True)
    (out_ref * R).sum().backward()
    with emulated_kernels(exact=True):
        seg_act = Act.empty(B, H, W, arch['semantic_nc'], 'cpu', zero=True)
        ops.nchw_to_nhwc(seg, seg_act)
        net = SpadeGenNet(arch, seg_act, 'cpu', training=True, need_grad=True)
        net.load_state_dict(sd)
        out = ops.nhwc_to_nchw(net.forward(), 3)
        assert rel_l2(out, out_ref.detach()) < 1e-5
        dS = Act.empty(B, H, W, 3, 'cpu', zero=True)
        ops.nchw_to_nhwc(R, dS)
        net.arena.g.zero_()
        net.backward(dS)
        scale = max(float(p.grad.abs().max()) for p in params.values() if p.grad is not None)
        for k, p in params.items():
            if p.grad is None:
                continue
            err = float((net.arena.view(k, 'g') - p.grad).abs().max())
            assert err <= 2e-3 * float(p.grad.abs().max()) + 2e-5 * scale, (k, err, float(p.grad.abs().max()))
        for k, v in osd.items():
            if 'running_' in k:
                assert float((net.state_dict()[k] - v.detach()).abs().max()) < 1e-4, k


@pytest.mark.timeout(900)
def test_spade_mse_distill_step_exact(golden_dir):
    """--distill_G_loss_type mse through cat_b200.adaptors.Adaptors in exact emulation: the three MSE terms, the gradients of
    the adaptor convs and of the student (the 1x1 input-gradient GEMM accumulated into d(activation) at the mapping layers)
    and the adaptors after the Adam step of optimizer_G."""
    from oracle import cat_oracle as O
    from oracle import spade_oracle as SO
    from oracle.kernel_emu import emulated_kernels
    fix = torch.load(os.path.join(golden_dir, 'spade_more.pt'), weights_only=False)
    add = torch.load(os.path.join(golden_dir, 'spade_more_mse.pt'), weights_only=False)
    vgg = SO.make_vgg_sd(fix['vgg_seed'])
    s = fix['steps'][0]
    hp = dict(fix['hp'], distill_loss_type='mse', lambda_distill=add['lambda_distill'])
    st = dict(teacher_sd=O.clone_sd(fix['teacher_sd']), student_sd=O.clone_sd(fix['student_sd0']), D_sd=O.clone_sd(fix['D_sd0']),
              vgg_sd=vgg, teacher_arch=fix['teacher_arch'], student_arch=fix['student_arch'], D_arch=fix['D_arch'],
              adam_G={}, adam_D={}, netA_sds=[O.clone_sd(sd) for sd in add['netA_sd0']])
    seg = SO.preprocess_input(s['label'], s['instance'], hp['n_label'])
    ref = SO.spade_distill_step(st, seg, s['image'], hp)
    B, _, H, W = s['image'].shape
    with emulated_kernels(exact=True):
        from cat_b200.spade_distill_engine import SpadeDistillStep
        eng = SpadeDistillStep(fix['teacher_arch'], fix['student_arch'], fix['D_arch'], hp, B, H, W, device='cpu')
        eng.load(fix['teacher_sd'], fix['student_sd0'], fix['D_sd0'], vgg, add['netA_sd0'])
        eng.set_input(s['label'], s['instance'], s['image'])
        eng.step()
        L = eng.get_losses()
        for k_ref, k in LOSSES:
            r = float(ref[k_ref])
            assert abs(L[k] - r) <= 1e-5 * max(1.0, abs(r)), (k, L[k], r)
            assert abs(L[k] - add['losses'][('D_loss/' if k.startswith('D_') else 'G_loss/') + k]) <= 1e-4 * max(1.0, abs(r)), k
        for i in range(3):
            assert abs(L['G_distill%d' % i] - float(ref['loss_G_distill_terms'][i])) < 1e-5
        for k, g in ref['A_grads'].items():
            assert float((eng.A.arena.view(k[1:], 'g') - g).abs().max()) <= 2e-3 * float(g.abs().max()) + 1e-8, k
        scale = max(float(g.abs().max()) for g in ref['S_grads'].values())
        for k, g in ref['S_grads'].items():
            err = float((eng.S.arena.view(k, 'g') - g).abs().max())
            assert err <= 2e-3 * float(g.abs().max()) + 2e-5 * scale, (k, err, float(g.abs().max()))
        for i, sd in enumerate(eng.A.state_dicts()):
            for k, v in sd.items():
                assert float((v - st['netA_sds'][i][k]).abs().max()) <= 1e-5, (i, k)


@pytest.mark.timeout(900)
def test_spade_first_step_with_the_student_in_eval_mode_exact(golden_dir):
    """hp['student_training'] = False on the SPADE engine: every BatchNorm of the student on its running statistics (both
    student forwards of the step), Norm.backward with an infinite element count, and the conv biases in front of a
    BatchNorm -- inert in training mode -- receiving d bias = scale * sum(dz) through the bias pool."""
    from oracle import spade_oracle as SO
    from oracle.kernel_emu import emulated_kernels
    from test_spade_oracle_golden import first_step_state
    fix, add, st, hp = first_step_state(golden_dir)
    student0 = {k: v.clone() for k, v in st['student_sd'].items()}
    vgg = st['vgg_sd']
    s = fix['steps'][0]
    seg = SO.preprocess_input(s['label'], s['instance'], hp['n_label'])
    ref = SO.spade_distill_step(st, seg, s['image'], hp)
    B, _, H, W = s['image'].shape
    with emulated_kernels(exact=True):
        from cat_b200.spade_distill_engine import SpadeDistillStep
        eng = SpadeDistillStep(fix['teacher_arch'], fix['student_arch'], fix['D_arch'], hp, B, H, W, device='cpu')
        eng.load(fix['teacher_sd'], student0, fix['D_sd0'], vgg)
        eng.set_input(s['label'], s['instance'], s['image'])
        eng.step()
        L = eng.get_losses()
        for k_ref, k in LOSSES:
            r = float(ref[k_ref])
            assert abs(L[k] - r) <= 1e-5 * max(1.0, abs(r)), (k, L[k], r)
        assert rel_l2(eng_out(eng), ref['Sfake_B_D']) < 1e-4
        scale = max(float(g.abs().max()) for g in ref['S_grads'].values())
        live_mine, live_ref = [], []
        for k, g in ref['S_grads'].items():
            mine = eng.S.arena.view(k, 'g')
            err = float((mine - g).abs().max())
            assert err <= 2e-3 * float(g.abs().max()) + 2e-5 * scale, (k, err, float(g.abs().max()))
            if k.endswith('.0.conv.bias') and float(g.abs().max()) > 0:      # a conv bias in front of a BatchNorm: now live
                live_mine.append(mine.flatten().clone())
                live_ref.append(g.flatten())
        # these gradients are tiny on this fixture (1e-9 ... 1e-7, far below the floor above), so they get their own
        # relative comparison, over all of them together (the smallest are cancellation noise in fp32 on both sides)
        assert len(live_ref) > 10
        assert rel_l2(torch.cat(live_mine), torch.cat(live_ref)) < 2e-2
        sd = eng.S.state_dict()
        for k, v in add['running_stats'].items():
            assert torch.equal(sd[k], v), k
        if os.environ.get('CATB_SLOW_TESTS', '0') != '1':
            return
        # netG_student.train() after the first evaluate_model: the same engine continues in training mode (also exercised
        # by the Inception flow test and by tests/test_zzz_train_gpu.py)
        ref2 = SO.spade_distill_step(st, seg, s['image'], dict(hp, student_training=True))
        eng.set_student_training(True)
        eng.step()
        L = eng.get_losses()
        for k_ref, k in LOSSES:
            r = float(ref2[k_ref])
            assert abs(L[k] - r) <= 2e-3 * max(1.0, abs(r)), (k, L[k], r)


def eng_out(eng):
    from cat_b200 import ops
    return ops.nhwc_to_nchw(eng.S.out, 3)
