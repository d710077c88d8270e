"""The reference-facing teacher-training protocol (cat_b200.models.create_model -> Pix2PixModel / CycleGANModel /
SPADEModel) in exact kernel emulation on CPU: a trainer-style loop (setup -> set_input -> optimize_parameters ->
get_current_losses -> save_networks, trainer.py:79-175) must reproduce the pinned oracle's losses, keep the module
parameters aliased to the engine arenas and write checkpoints with the reference's file names and state_dict keys.
The GPU suite repeats the loop through libcatb200.so (tests/test_zzz_train_gpu.py)."""
import argparse
import os
import random

import pytest
import torch


def _common(fix_arch, D_arch, log_dir, **kw):
    base = dict(isTrain=True, gpu_ids=[0], log_dir=log_dir, input_nc=3, output_nc=3, netG='inception_9blocks', dropout_rate=0,
                norm=fix_arch['norm'], norm_affine=fix_arch['affine'], norm_affine_D=D_arch['affine'],
                norm_track_running_stats=fix_arch['track_running_stats'], norm_momentum=0.1, norm_epsilon=1e-5, channels=None,
                channels_reduction_factor=6, kernel_sizes=[1, 3, 5], active_fn='nn.ReLU', active_fn_D='nn.LeakyReLU',
                init_type='normal', init_gain=0.02, netD='n_layers', ngf=fix_arch['widths'][0], ndf=D_arch['ndf'], n_layers_D=3,
                direction='AtoB', nepochs=5, nepochs_decay=15, lr_policy='linear', cuda_graph=False)
    base.update(kw)
    return argparse.Namespace(**base)


def _check_losses(L, ref, prefix_of, tol):
    for key, v in L.items():
        name = key.split('/')[-1]
        assert key == prefix_of(name) + name
        r = float(ref['loss_' + name])
        assert abs(v - r) <= tol(name) * max(1.0, abs(r)), (key, v, r)


def _prefix(name):
    return 'D_loss/' if name.startswith('D_') else 'G_loss/'


@pytest.mark.timeout(900)
def test_pix2pix_model_protocol(golden_dir, tmp_path):
    from oracle import train_oracle as TO
    from oracle.cat_oracle import clone_sd
    from oracle.kernel_emu import emulated_kernels
    fix = torch.load(os.path.join(golden_dir, 'train_pix2pix_in_lsgan_l2.pt'), weights_only=False)
    hp = fix['hp']
    opt = _common(fix['G_arch'], fix['D_arch'], str(tmp_path), model='pix2pix', dataset_mode='aligned', gan_mode=hp['gan_mode'],
                  recon_loss_type=hp['recon_loss_type'], lambda_recon=hp['lambda_recon'], lambda_gan=hp['lambda_gan'], lr=hp['lr'],
                  beta1=hp['beta1'], restore_G_path=None, restore_D_path=None)
    st = dict(G_sd=clone_sd(fix['G_sd0']), D_sd=clone_sd(fix['D_sd0']), G_arch=fix['G_arch'], D_arch=fix['D_arch'],
              adam_G={}, adam_D={})
    with emulated_kernels(exact=True):
        from cat_b200.models import create_model
        model = create_model(opt, verbose=False)
        model.setup(opt, verbose=False)
        assert list(model.netG.state_dict().keys()) == list(fix['G_sd0'].keys())
        model.netG.load_state_dict(fix['G_sd0'])          # reference checkpoints load straight into the module trees
        model.netD.load_state_dict(fix['D_sd0'])
        w0 = model.netG.state_dict()['up_sampling.7.weight'].clone()
        for it, s in enumerate(fix['steps']):
            ref = TO.pix2pix_train_step(st, s['real_A'], s['real_B'], hp)
            B = s['real_A'].shape[0]
            model.set_input({'A': s['real_A'], 'B': s['real_B'], 'A_paths': ['x'] * B, 'B_paths': ['x'] * B})
            model.optimize_parameters(it)
            L = model.get_current_losses()
            assert list(L.keys()) == ['G_loss/G_gan', 'G_loss/G_recon', 'D_loss/D_real', 'D_loss/D_fake']
            _check_losses(L, ref, _prefix, lambda n: 1e-4 if it == 0 else 2e-3)
            assert model.loss_G_recon == L['G_loss/G_recon']
        # the module parameters ARE the engine arena: the optimiser steps are visible through the modules
        sd = model.netG.state_dict()
        assert not torch.equal(sd['up_sampling.7.weight'], w0)
        eng_sd = model.engine.G.state_dict()
        for k, v in sd.items():
            if v.is_floating_point():
                assert torch.equal(v.reshape(-1), eng_sd[k].reshape(-1)), k
        lr = hp['lr']
        worst = max(float((sd[k] - v).abs().max()) for k, v in st['G_sd'].items()
                    if v.is_floating_point() and not k.endswith(('running_mean', 'running_var')))
        assert worst <= 2.1 * lr * len(fix['steps'])
        model.test()                                        # inference through the same arenas
        assert model.fake_B.shape == s['real_A'].shape and torch.isfinite(model.fake_B).all()
        # evaluate_model (pix2pix_model.py:214-281): eval-mode inference over the evaluation set + metric bookkeeping
        model.eval_dataloader = [{'A': s['real_A'][:2], 'B': s['real_B'][:2], 'A_paths': ['d/a.png', 'd/b.png']}]
        got = {}
        model.metric_fns = {'fid': lambda fakes: got.setdefault('n', len(fakes)) * 7.0,
                            'mIoU': lambda fakes, names: got.setdefault('names', names) and 0.25}
        ret = model.evaluate_model(1)
        assert ret['metric/fid'] == 7.0 and ret['metric/mIoU-best'] == 0.25 and got['names'] == ['a', 'b'] and model.is_best
        assert model.netG.training
        want = TO.O.generator_forward(TO.O.clone_sd({k: v.detach().clone() for k, v in model.netG.state_dict().items()}),
                                      fix['G_arch'], s['real_A'][:2], training=False)
        assert float((model.fake_B - want).norm() / want.norm()) < 1e-5
        model.save_networks('latest')
        ck = os.path.join(str(tmp_path), 'checkpoints')
        g = torch.load(os.path.join(ck, 'latest_net_G.pth'), weights_only=False)
        assert list(g.keys()) == list(fix['G_sd0'].keys())
        d = torch.load(os.path.join(ck, 'latest_net_D.pth'), weights_only=False)
        assert list(d.keys()) == list(fix['D_sd0'].keys())
        model.update_learning_rate()
        assert abs(model.optimizers[0].param_groups[0]['lr'] - lr) < 1e-12   # epoch 1 of 5: still the base rate


@pytest.mark.timeout(900)
def test_cycle_gan_model_protocol(golden_dir, tmp_path):
    from oracle import train_oracle as TO
    from oracle.cat_oracle import clone_sd
    from oracle.kernel_emu import emulated_kernels
    fix = torch.load(os.path.join(golden_dir, 'train_cyclegan_in_lsgan.pt'), weights_only=False)
    hp = fix['hp']
    opt = _common(fix['G_arch'], fix['D_arch'], str(tmp_path), model='cycle_gan', dataset_mode='unaligned', gan_mode=hp['gan_mode'],
                  lambda_A=hp['lambda_A'], lambda_B=hp['lambda_B'], lambda_identity=hp['lambda_identity'], lr=hp['lr'],
                  beta1=hp['beta1'], pool_size=hp['pool_size'])
    st = dict(G_A_sd=clone_sd(fix['G_A_sd0']), G_B_sd=clone_sd(fix['G_B_sd0']), D_A_sd=clone_sd(fix['D_A_sd0']),
              D_B_sd=clone_sd(fix['D_B_sd0']), G_arch=fix['G_arch'], D_arch=fix['D_arch'], adam_G={}, adam_D={},
              pool_A=TO.ImagePool(hp['pool_size']), pool_B=TO.ImagePool(hp['pool_size']))
    steps = fix['steps'][:1]
    random.seed(fix['python_random_seed'])
    refs = [TO.cyclegan_train_step(st, s['real_A'], s['real_B'], hp) for s in steps]
    with emulated_kernels(exact=True):
        from cat_b200.models import create_model
        model = create_model(opt, verbose=False)
        model.setup(opt, verbose=False)
        for n in ('G_A', 'G_B', 'D_A', 'D_B'):
            getattr(model, 'net' + n).load_state_dict(fix[n + '_sd0'])
        random.seed(fix['python_random_seed'])
        for it, (s, ref) in enumerate(zip(steps, refs)):
            B = s['real_A'].shape[0]
            model.set_input({'A': s['real_A'], 'B': s['real_B'], 'A_paths': ['x'] * B, 'B_paths': ['x'] * B})
            model.optimize_parameters(it)
            L = model.get_current_losses()
            assert list(L.keys()) == ['D_loss/D_A', 'G_loss/G_A', 'G_loss/G_cycle_A', 'G_loss/G_idt_A', 'D_loss/D_B', 'G_loss/G_B',
                                      'G_loss/G_cycle_B', 'G_loss/G_idt_B']
            _check_losses(L, ref, _prefix, lambda n: 1e-4 if it == 0 else 3e-3)
        for n in ('G_A', 'G_B'):       # both generators moved, and the modules see it
            sd = getattr(model, 'net' + n).state_dict()
            assert not torch.equal(sd['up_sampling.7.weight'], fix[n + '_sd0']['up_sampling.7.weight'])
            eng_sd = getattr(model.engine, n).state_dict()
            for k, v in sd.items():
                if v.is_floating_point():
                    assert torch.equal(v.reshape(-1), eng_sd[k].reshape(-1)), (n, k)
        model.test()
        assert model.rec_A.shape == s['real_A'].shape and torch.isfinite(model.rec_B).all()
        model.eval_dataloader_AtoB = [{'A': s['real_A'], 'A_paths': ['x/1.jpg', 'x/2.jpg']}]
        model.eval_dataloader_BtoA = [{'A': s['real_B'], 'A_paths': ['y/3.jpg', 'y/4.jpg']}]
        model.metric_fns_A = {'fid': lambda fakes: 3.0}
        model.metric_fns_B = {'fid': lambda fakes: 4.0}
        ret = model.evaluate_model(2)
        assert ret['metric/fid_A'] == 3.0 and ret['metric/fid_B-best'] == 4.0 and model.is_best_A and model.is_best_B
        assert model.netG_A.training and model.netG_B.training
        model.save_networks(3)     # epoch-numbered checkpoints (trainer.py:170)
        ck = os.path.join(str(tmp_path), 'checkpoints')
        for n in ('G_A', 'G_B', 'D_A', 'D_B'):
            sd = torch.load(os.path.join(ck, '3_net_%s.pth' % n), weights_only=False)
            assert list(sd.keys()) == list(fix[n + '_sd0'].keys())


@pytest.mark.timeout(900)
def test_spade_model_protocol(golden_dir, tmp_path):
    from oracle import spade_oracle as SO
    from oracle import train_oracle as TO
    from oracle.cat_oracle import clone_sd
    from oracle.kernel_emu import emulated_kernels
    fix = torch.load(os.path.join(golden_dir, 'train_spade_more.pt'), weights_only=False)
    hp, Ga, Da = fix['hp'], fix['G_arch'], fix['D_arch']
    vgg = SO.make_vgg_sd(fix['vgg_seed'])
    opt = argparse.Namespace(
        isTrain=True, gpu_ids=[0], log_dir=str(tmp_path), model='spade', input_nc=hp['n_label'], output_nc=3,
        semantic_nc=Ga['semantic_nc'], ngf=Ga['fc_out'] // 16, netG='inception_spade', norm_G='spadesyncbatch3x3', norm='instance',
        dropout_rate=0, num_upsampling_layers=Ga['num_upsampling_layers'], crop_size=128, aspect_ratio=2.0, channels=None,
        channels_reduction_factor=6, kernel_sizes=[1, 3, 5], active_fn='nn.LeakyReLU', init_type='xavier', init_gain=0.02,
        netD='multi_scale', ndf=Da['ndf'], n_layers_D=Da['n_layers'], num_D=Da['num_D'], norm_D=Da['norm_D'], gan_mode='hinge',
        lambda_gan=hp['lambda_gan'], lambda_feat=hp['lambda_feat'], lambda_vgg=hp['lambda_vgg'], lr=hp['lr_G'] * 2, beta1=0.5,
        beta2=0.999, no_TTUR=False, nepochs=100, nepochs_decay=100, lr_policy='linear', restore_G_path=None, restore_D_path=None,
        vgg_state_dict=vgg, cuda_graph=False)
    st = dict(G_sd=clone_sd(fix['G_sd0']), D_sd=clone_sd(fix['D_sd0']), vgg_sd=vgg, G_arch=Ga, D_arch=Da, adam_G={}, adam_D={})
    s = fix['steps'][0]
    seg = SO.preprocess_input(s['label'], s['instance'], hp['n_label'])
    ref = TO.spade_train_step(st, seg, s['image'], hp)
    with emulated_kernels(exact=True):
        from cat_b200.models import create_model
        model = create_model(opt, verbose=False)
        model.setup(opt, verbose=False)
        mm = model.modules_on_one_gpu
        assert list(mm.netG.state_dict().keys()) == list(fix['G_sd0'].keys())
        assert list(mm.netD.state_dict().keys()) == list(fix['D_sd0'].keys())
        mm.netG.load_state_dict(fix['G_sd0'])
        mm.netD.load_state_dict(fix['D_sd0'])
        assert mm.netG.arch() == Ga
        B = s['image'].shape[0]
        model.set_input({'label': s['label'], 'instance': s['instance'], 'image': s['image'], 'path': ['x'] * B})
        model.optimize_parameters(0)
        L = model.get_current_losses()
        assert list(L.keys()) == ['G_loss/G_gan', 'G_loss/G_feat', 'G_loss/G_vgg', 'D_loss/D_real', 'D_loss/D_fake']
        _check_losses(L, ref, _prefix, lambda n: 1e-4)
        model.test()
        assert model.fake_B.shape == s['image'].shape and torch.isfinite(model.fake_B).all()
        model.eval_dataloader = [{'label': s['label'][:1], 'instance': s['instance'][:1], 'image': s['image'][:1], 'path': ['c/frankfurt_0.png']}]
        model.metric_fns = {'fid': lambda fakes: 9.0}
        ret = model.evaluate_model(1)
        assert ret == {'metric/fid': 9.0, 'metric/fid-mean': 9.0, 'metric/fid-best': 9.0} and mm.netG.training
        want = SO.spade_generator_forward({k: v.detach().clone() for k, v in mm.netG.state_dict().items()}, Ga, seg[:1], training=False)
        assert float((model.fake_B - want).norm() / want.norm()) < 1e-5
        model.save_networks('latest')
        ck = os.path.join(str(tmp_path), 'checkpoints')
        g = torch.load(os.path.join(ck, 'latest_net_G.pth'), weights_only=False)
        assert list(g.keys()) == list(fix['G_sd0'].keys())
        model.update_learning_rate()
        assert abs(model.optimizer_G.param_groups[0]['lr'] - hp['lr_G']) < 1e-12
        assert abs(model.optimizer_D.param_groups[0]['lr'] - hp['lr_D']) < 1e-12
