"""Host logic of the teacher-training engines (cat_b200/train_engine.py) on CPU: every kernel
wrapper is swapped for its torch restatement (oracle/kernel_emu.py, test infrastructure) in exact mode (fp32 emulated
buffers), so launch order, the hand-derived backward passes (incl. the gradient through a generator's INPUT for the
CycleGAN cycle terms), shared arenas of the three applications of each CycleGAN generator, the device image pools and
the buffer plumbing must reproduce the fp32 oracle -- which tests/test_train_oracle_golden.py pins to the real reference
models -- to rounding.  The GPU suite (tests/test_zzz_train_gpu.py) runs the same steps through libcatb200.so."""
import os
import random

import pytest
import torch


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name + '.pt'), weights_only=False)


def _check_grads(net, grads, tag):
    scale = max(float(g.abs().max()) for g in grads.values())
    n = 0
    for k, g in grads.items():
        if not net.arena.has(k):
            continue      # conv biases in front of a norm layer: analytically zero gradient, not stored
        n += 1
        err = float((net.arena.view(k, 'g') - g).abs().max())
        assert err <= 2e-3 * float(g.abs().max()) + 2e-5 * scale, (tag, k, err, float(g.abs().max()))
    assert n > 10, tag


@pytest.mark.timeout(600)
@pytest.mark.parametrize('name', ['train_pix2pix_bn_hinge', 'train_pix2pix_in_lsgan_l2'])
def test_pix2pix_train_step_exact(golden_dir, name):
    from oracle import train_oracle as TO
    from oracle.cat_oracle import clone_sd
    from oracle.kernel_emu import emulated_kernels
    from cat_b200 import ops
    fix = _load(golden_dir, name)
    s = fix['steps'][0]
    B, _, H, W = s['real_A'].shape
    st = dict(G_sd=clone_sd(fix['G_sd0']), D_sd=clone_sd(fix['D_sd0']), G_arch=fix['G_arch'], D_arch=fix['D_arch'],
              adam_G={}, adam_D={})
    ref = TO.pix2pix_train_step(st, s['real_A'], s['real_B'], fix['hp'])
    with emulated_kernels(exact=True):
        from cat_b200.train_engine import Pix2PixTrainStep
        eng = Pix2PixTrainStep(fix['G_arch'], fix['D_arch'], fix['hp'], B, H, W, device='cpu')
        eng.load(fix['G_sd0'], fix['D_sd0'])
        eng.set_input(s['real_A'], s['real_B'])
        eng.step()
        assert rel_l2(ops.nhwc_to_nchw(eng.G.out, 3), ref['fake_B']) < 1e-5
        L = eng.get_losses()
        assert set(L) == set(s['losses'])
        for k, v in L.items():
            # G_gan is evaluated AFTER optimizer_D.step: in the hinge fixture the gradient of the last conv's bias is
            # +0.5 - 0.5 = 0 up to rounding (every hinge mask is active), and Adam turns that rounding into a +-lr step
            # of arbitrary sign, i.e. a 2 lr = 4e-4 band on the prediction
            tol = 1e-3 if k == 'G_gan' else 1e-5
            r = float(ref['loss_' + k])
            assert abs(v - r) <= tol * max(1.0, abs(r)), (k, v, r)
            assert abs(v - s['losses'][k]) <= max(tol, 1e-4) * max(1.0, abs(s['losses'][k])), (k, v)   # the real reference's value
        _check_grads(eng.G, ref['G_grads'], 'G')
        _check_grads(eng.D, ref['D_grads'], 'D')
        sd = eng.G.state_dict()
        for k, v in st['G_sd'].items():
            if 'running_' in k:
                assert float((sd[k] - v).abs().max()) < 1e-5, k


@pytest.mark.timeout(900)
@pytest.mark.parametrize('name', ['train_cyclegan_in_lsgan', 'train_cyclegan_bn_lsgan'])
def test_cyclegan_train_steps_exact(golden_dir, name):
    from oracle import train_oracle as TO
    from oracle.cat_oracle import clone_sd
    from oracle.kernel_emu import emulated_kernels
    from cat_b200 import ops
    fix = _load(golden_dir, name)
    hp = fix['hp']
    B, _, H, W = fix['steps'][0]['real_A'].shape
    st = dict(G_A_sd=clone_sd(fix['G_A_sd0']), G_B_sd=clone_sd(fix['G_B_sd0']), D_A_sd=clone_sd(fix['D_A_sd0']),
              D_B_sd=clone_sd(fix['D_B_sd0']), G_arch=fix['G_arch'], D_arch=fix['D_arch'], adam_G={}, adam_D={},
              pool_A=TO.ImagePool(hp['pool_size']), pool_B=TO.ImagePool(hp['pool_size']))
    random.seed(fix['python_random_seed'])
    # pool of 3, batch 2: filled during steps 0-1, history decisions from step 1 on; the BatchNorm fixture has no pool and
    # only needs the first step (running statistics after the three applications of each generator)
    steps = fix['steps'][:3 if hp['pool_size'] else 1]
    refs = [TO.cyclegan_train_step(st, s['real_A'], s['real_B'], hp) for s in steps]
    with emulated_kernels(exact=True):
        from cat_b200.train_engine import CycleGANTrainStep
        eng = CycleGANTrainStep(fix['G_arch'], fix['D_arch'], hp, B, H, W, device='cpu')
        eng.load(fix['G_A_sd0'], fix['G_B_sd0'], fix['D_A_sd0'], fix['D_B_sd0'])
        random.seed(fix['python_random_seed'])      # the pools draw from Python's global generator, like the reference
        for it, (s, ref) in enumerate(zip(steps, refs)):
            eng.set_input(s['real_A'], s['real_B'])
            eng.step()
            L = eng.get_losses()
            assert set(L) == set(s['losses'])
            # the D losses from step 2 on depend on the pools' history decisions (pool of 3 images)
            tol = 1e-5 if it == 0 else 3e-3
            for k, v in L.items():
                r = float(ref['loss_' + k])
                assert abs(v - r) <= tol * max(1.0, abs(r)), (it, k, v, r)
            assert rel_l2(ops.nhwc_to_nchw(eng.d_in_fake_B, 3), ref['pooled_B']) < (1e-5 if it == 0 else 2e-3), it
            assert rel_l2(ops.nhwc_to_nchw(eng.d_in_fake_A, 3), ref['pooled_A']) < (1e-5 if it == 0 else 2e-3), it
            if it:
                continue
            for mine, theirs in ((eng.GA_real.out, 'fake_B'), (eng.GB_real.out, 'fake_A'), (eng.GB_cyc.out, 'rec_A'),
                                 (eng.GA_cyc.out, 'rec_B')):
                assert rel_l2(ops.nhwc_to_nchw(mine, 3), ref[theirs]) < 1e-4, theirs   # rec_*: two chained generators
            assert rel_l2(ops.nhwc_to_nchw(eng.d_fake_B, 3), ref['fake_B_grad']) < 1e-4
            assert rel_l2(ops.nhwc_to_nchw(eng.d_fake_A, 3), ref['fake_A_grad']) < 1e-4
            for tag, net in (('G_A', eng.G_A), ('G_B', eng.G_B), ('D_A', eng.D_A), ('D_B', eng.D_B)):
                _check_grads(net, ref[tag + '_grads'], tag)
            # running statistics of G_A after its three training forwards, in the reference's order
            sd = eng.G_A.state_dict()
            for k, v in fix['steps'][0].get('G_A_buffers_after', {}).items():
                if 'running_' in k:
                    assert float((sd[k] - v).abs().max()) < 1e-5 * max(1.0, float(v.abs().max())), k


@pytest.mark.timeout(900)
def test_spade_train_step_exact(golden_dir):
    from oracle import spade_oracle as SO
    from oracle import train_oracle as TO
    from oracle.cat_oracle import clone_sd
    from oracle.kernel_emu import emulated_kernels
    from cat_b200 import ops
    fix = _load(golden_dir, 'train_spade_more')
    vgg = SO.make_vgg_sd(fix['vgg_seed'])
    s, hp = fix['steps'][0], fix['hp']
    st = dict(G_sd=clone_sd(fix['G_sd0']), D_sd=clone_sd(fix['D_sd0']), vgg_sd=vgg, G_arch=fix['G_arch'], D_arch=fix['D_arch'],
              adam_G={}, adam_D={})
    seg = SO.preprocess_input(s['label'], s['instance'], hp['n_label'])
    ref = TO.spade_train_step(st, seg, s['image'], hp)
    B, _, H, W = s['image'].shape
    with emulated_kernels(exact=True):
        from cat_b200.train_engine import SpadeTrainStep
        eng = SpadeTrainStep(fix['G_arch'], fix['D_arch'], hp, B, H, W, device='cpu')
        eng.load(fix['G_sd0'], fix['D_sd0'], vgg)
        eng.set_input(s['label'], s['instance'], s['image'])
        eng.step()
        assert torch.equal(ops.nhwc_to_nchw(eng.seg, eng.snc), seg)
        L = eng.get_losses()
        assert set(L) == set(s['losses'])
        for k, v in L.items():
            r = float(ref['loss_' + k])
            assert abs(v - r) <= 1e-5 * max(1.0, abs(r)), (k, v, r)
            assert abs(v - s['losses'][k]) <= 1e-4 * max(1.0, abs(s['losses'][k])), (k, v)
        _check_grads(eng.G, ref['G_grads'], 'G')
        # D gradients of the two full-resolution layers flip a few LeakyReLU kinks in fp32 (tests/test_train_oracle_golden.py):
        # compared as a whole
        ks = [k for k in ref['D_grads'] if eng.D.arena.has(k)]
        mine = torch.cat([eng.D.arena.view(k, 'g').flatten() for k in ks])
        theirs = torch.cat([ref['D_grads'][k].flatten() for k in ks])
        assert rel_l2(mine, theirs) < 2e-2
        sd = eng.G.state_dict()
        for k, v in st['G_sd'].items():
            if 'running_' in k:
                assert float((sd[k] - v).abs().max()) < 1e-4, k


@pytest.mark.timeout(600)
def test_cyclegan_without_identity_term_and_without_pool_exact(golden_dir):
    """--lambda_identity 0 (no third application of the generators, cycle_gan_model.py:262-273) and --pool_size 0 (the
    discriminators read the current fakes, utils/image_pool.py:31-32): one step against the oracle."""
    from oracle import train_oracle as TO
    from oracle.cat_oracle import clone_sd
    from oracle.kernel_emu import emulated_kernels
    fix = _load(golden_dir, 'train_cyclegan_in_lsgan')
    hp = dict(fix['hp'], lambda_identity=0.0, pool_size=0)
    s = fix['steps'][0]
    B, _, H, W = s['real_A'].shape
    st = dict(G_A_sd=clone_sd(fix['G_A_sd0']), G_B_sd=clone_sd(fix['G_B_sd0']), D_A_sd=clone_sd(fix['D_A_sd0']),
              D_B_sd=clone_sd(fix['D_B_sd0']), G_arch=fix['G_arch'], D_arch=fix['D_arch'], adam_G={}, adam_D={},
              pool_A=TO.ImagePool(0), pool_B=TO.ImagePool(0))
    ref = TO.cyclegan_train_step(st, s['real_A'], s['real_B'], hp)
    with emulated_kernels(exact=True):
        from cat_b200.train_engine import CycleGANTrainStep
        eng = CycleGANTrainStep(fix['G_arch'], fix['D_arch'], hp, B, H, W, device='cpu')
        assert eng.GA_idt is None and eng.GB_idt is None
        eng.load(fix['G_A_sd0'], fix['G_B_sd0'], fix['D_A_sd0'], fix['D_B_sd0'])
        eng.set_input(s['real_A'], s['real_B'])
        eng.step()
        assert eng.d_in_fake_B is eng.GA_real.out and eng.d_in_fake_A is eng.GB_real.out
        L = eng.get_losses()
        assert L['G_idt_A'] == 0.0 and L['G_idt_B'] == 0.0
        for k, v in L.items():
            r = float(ref['loss_' + k])
            assert abs(v - r) <= 1e-5 * max(1.0, abs(r)), (k, v, r)
        for tag, net in (('G_A', eng.G_A), ('G_B', eng.G_B), ('D_A', eng.D_A), ('D_B', eng.D_B)):
            _check_grads(net, ref[tag + '_grads'], tag)


@pytest.mark.timeout(600)
def test_cyclegan_gan_terms_alone_exact(golden_dir):
    """--lambda_A 0 --lambda_B 0: only the two GAN terms of backward_G remain (cycle and identity terms carry the weights
    lambda_A / lambda_B, cycle_gan_model.py:262-289).  With the default weights (10) the cycle terms dominate the generators'
    gradients; here the GAN path through the frozen discriminators must reproduce the oracle on its own."""
    from oracle import train_oracle as TO
    from oracle.cat_oracle import clone_sd
    from oracle.kernel_emu import emulated_kernels
    fix = _load(golden_dir, 'train_cyclegan_in_lsgan')
    hp = dict(fix['hp'], lambda_A=0.0, lambda_B=0.0, pool_size=0)
    s = fix['steps'][0]
    B, _, H, W = s['real_A'].shape
    st = dict(G_A_sd=clone_sd(fix['G_A_sd0']), G_B_sd=clone_sd(fix['G_B_sd0']), D_A_sd=clone_sd(fix['D_A_sd0']),
              D_B_sd=clone_sd(fix['D_B_sd0']), G_arch=fix['G_arch'], D_arch=fix['D_arch'], adam_G={}, adam_D={},
              pool_A=TO.ImagePool(0), pool_B=TO.ImagePool(0))
    ref = TO.cyclegan_train_step(st, s['real_A'], s['real_B'], hp)
    with emulated_kernels(exact=True):
        from cat_b200.train_engine import CycleGANTrainStep
        eng = CycleGANTrainStep(fix['G_arch'], fix['D_arch'], hp, B, H, W, device='cpu')
        eng.load(fix['G_A_sd0'], fix['G_B_sd0'], fix['D_A_sd0'], fix['D_B_sd0'])
        eng.set_input(s['real_A'], s['real_B'])
        eng.step()
        L = eng.get_losses()
        for k in ('G_cycle_A', 'G_cycle_B', 'G_idt_A', 'G_idt_B'):
            assert L[k] == 0.0
        for k, v in L.items():
            r = float(ref['loss_' + k])
            assert abs(v - r) <= 1e-5 * max(1.0, abs(r)), (k, v, r)
        for tag, net in (('G_A', eng.G_A), ('G_B', eng.G_B), ('D_A', eng.D_A), ('D_B', eng.D_B)):
            _check_grads(net, ref[tag + '_grads'], tag)
