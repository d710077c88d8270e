"""The reference-facing protocol (cat_b200.distillers / cat_b200.models.networks) on the GPU: a trainer-style
loop (create_distiller -> setup -> set_input -> optimize_parameters -> get_current_losses -> save_networks)
must produce the oracle's losses, keep the module parameters aliased to the engine arenas, and write
checkpoints with the reference's file names and state_dict keys."""
import argparse
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]


def _opt(fix, log_dir):
    hp, Ta, Da = fix['hp'], fix['teacher_arch'], fix['D_arch']
    return argparse.Namespace(
        isTrain=True, gpu_ids=[0], log_dir=log_dir, distiller='inception', input_nc=3, output_nc=3,
        teacher_ngf=Ta['widths'][0], student_ngf=8, teacher_netG='inception_9blocks', student_netG='inception_9blocks',
        norm=Ta['norm'], norm_affine=Ta['affine'], norm_affine_D=Da['affine'],
        norm_track_running_stats=Ta['track_running_stats'], norm_momentum=0.1, norm_epsilon=1e-5,
        channels=None, channels_reduction_factor=6, kernel_sizes=[1, 3, 5], active_fn='nn.ReLU', active_fn_D='nn.LeakyReLU',
        init_type='normal', init_gain=0.02, netD='n_layers', ndf=Da['ndf'], n_layers_D=3,
        dataset_mode='aligned' if hp['aligned'] else 'unaligned', direction='AtoB', gan_mode=hp['gan_mode'],
        recon_loss_type=hp.get('recon_loss_type', 'l1'), distill_G_loss_type='ka', lambda_distill=hp['lambda_distill'],
        lambda_recon=hp['lambda_recon'], lambda_gan=hp['lambda_gan'], lr=hp['lr'], beta1=hp['beta1'], nepochs=5,
        nepochs_decay=15, student_arch=fix['student_arch'], restore_teacher_G_path=None, restore_student_G_path=None,
        restore_D_path=None, cuda_graph=True)


@pytest.mark.parametrize('name', ['pix2pix_bn_hinge', 'cyclegan_in_lsgan'])
def test_trainer_style_loop(golden_dir, tmp_path, name):
    from cat_b200.distillers import create_distiller
    from oracle import cat_oracle as O
    fix = torch.load(os.path.join(golden_dir, name + '.pt'), weights_only=False)
    opt = _opt(fix, str(tmp_path))
    model = create_distiller(opt, verbose=False)
    model.setup(opt, verbose=False)
    # reference checkpoints load straight into the module trees
    model.netG_teacher.load_state_dict(fix['teacher_sd'])
    model.netG_student.load_state_dict(fix['student_sd0'])
    model.netD.load_state_dict(fix['D_sd0'])
    model.netG_student.train()
    state = dict(teacher_sd=O.clone_sd(fix['teacher_sd']), student_sd=O.clone_sd(fix['student_sd0']),
                 D_sd=O.clone_sd(fix['D_sd0']), teacher_arch=fix['teacher_arch'], student_arch=fix['student_arch'],
                 D_arch=fix['D_arch'], adam_G={}, adam_D={})
    w_before = model.netG_student.state_dict()['up_sampling.7.weight'].clone()
    for it, step in enumerate(fix['steps']):
        ref = O.distill_step(state, step['real_A'], step['real_B'], fix['hp'])
        B = step['real_A'].shape[0]
        model.set_input({'A': step['real_A'], 'B': step['real_B'], 'A_paths': ['x'] * B, 'B_paths': ['x'] * B})
        model.optimize_parameters(it)
        L = model.get_current_losses()
        assert list(L.keys()) == ['G_loss/G_gan', 'G_loss/G_distill', 'G_loss/G_recon', 'D_loss/D_fake', 'D_loss/D_real',
                                  'Specific_loss/G_distill0', 'Specific_loss/G_distill1', 'Specific_loss/G_distill2',
                                  'Specific_loss/G_distill3']
        for mine, theirs in (('G_loss/G_gan', 'loss_G_gan'), ('G_loss/G_recon', 'loss_G_recon'),
                             ('G_loss/G_distill', 'loss_G_distill'), ('D_loss/D_fake', 'loss_D_fake'),
                             ('D_loss/D_real', 'loss_D_real')):
            r = float(ref[theirs])
            assert abs(L[mine] - r) <= 5e-2 * max(1.0, abs(r)), (it, mine, L[mine], r)
        assert float(model.loss_G_recon) == L['G_loss/G_recon']
    # parameters of the module tree ARE the engine arena: the optimiser step is visible through the modules
    sd = model.netG_student.state_dict()
    assert not torch.equal(sd['up_sampling.7.weight'].cpu(), w_before.cpu())
    eng_sd = model.engine.S.state_dict()
    for k, v in sd.items():
        if v.is_floating_point():
            assert torch.equal(v.detach().cpu().reshape(-1), eng_sd[k].reshape(-1)), k
    lr = fix['hp']['lr']
    worst = max(float((sd[k].detach().cpu().double() - v.double()).abs().max()) for k, v in state['student_sd'].items()
                if v.is_floating_point() and not k.endswith(('running_mean', 'running_var')))
    assert worst <= 2.1 * lr * len(fix['steps'])
    # module forward (inference through the same arenas) reproduces the engine's student output
    model.test()
    from cat_b200 import ops
    assert model.Sfake_B.shape == step['real_A'].shape
    assert torch.isfinite(model.Sfake_B).all() and torch.isfinite(model.Tfake_B).all()
    assert float((model.Tfake_B.cpu() - ref['Tfake_B']).norm() / ref['Tfake_B'].norm()) < 3e-2
    # checkpoints: reference file names and key layout, and they load back
    model.save_networks('latest')
    ck = os.path.join(str(tmp_path), 'checkpoints')
    for f in ('latest_net_G.pth', 'latest_net_D.pth', 'latest_net_A-0.pth', 'latest_optim-0.pth', 'latest_optim-1.pth'):
        assert os.path.exists(os.path.join(ck, f)), f
    g = torch.load(os.path.join(ck, 'latest_net_G.pth'), weights_only=False)
    assert list(g.keys()) == list(fix['student_sd0'].keys())
    from cat_b200.models import networks
    fresh = networks.InceptionGenerator.from_arch(fix['student_arch'])
    fresh.load_state_dict(g)
    model.update_learning_rate()
    assert abs(model.optimizers[0].param_groups[0]['lr'] - lr) < 1e-12   # epoch 1 of 5: still the base rate
