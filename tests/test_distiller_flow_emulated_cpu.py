"""The Inception distiller mirror driven like trainer.py:79-175 in exact kernel emulation on CPU, including the reference's
mode sequence: `setup` leaves the student in eval() (model_profiling's side effect), the FIRST optimize_parameters therefore
runs on the BatchNorm running statistics, the first `evaluate_model` (generator inference over the evaluation set, metric
bookkeeping) returns the student to train(), and the following steps use batch statistics -- every step compared with the
oracle run in the same modes."""
import argparse
import os

import pytest
import torch


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
        restore_D_path=None, cuda_graph=False)


@pytest.mark.timeout(900)
def test_trainer_flow_with_the_first_step_in_eval_mode(golden_dir, tmp_path):
    from oracle import cat_oracle as O
    from oracle.kernel_emu import emulated_kernels
    fix = torch.load(os.path.join(golden_dir, 'pix2pix_bn_lsgan_l2.pt'), weights_only=False)
    add = torch.load(os.path.join(golden_dir, 'pix2pix_bn_lsgan_l2_first_step.pt'), weights_only=False)
    student0 = O.clone_sd(fix['student_sd0'])
    student0.update({k: v.clone() for k, v in add['running_stats'].items()})
    st = dict(teacher_sd=O.clone_sd(fix['teacher_sd']), student_sd=O.clone_sd(student0), D_sd=O.clone_sd(fix['D_sd0']),
              teacher_arch=fix['teacher_arch'], student_arch=fix['student_arch'], D_arch=fix['D_arch'], adam_G={}, adam_D={})
    s0, s1 = fix['steps']
    B = s0['real_A'].shape[0]
    batch = lambda s: {'A': s['real_A'], 'B': s['real_B'], 'A_paths': ['a/%d.png' % i for i in range(B)], 'B_paths': ['x'] * B}
    keys = (('G_loss/G_gan', 'loss_G_gan'), ('G_loss/G_recon', 'loss_G_recon'), ('G_loss/G_distill', 'loss_G_distill'),
            ('D_loss/D_fake', 'loss_D_fake'), ('D_loss/D_real', 'loss_D_real'))
    with emulated_kernels(exact=True):
        from cat_b200.distillers import create_distiller
        opt = _opt(fix, str(tmp_path))
        model = create_distiller(opt, verbose=False)
        model.setup(opt, verbose=False)
        assert not model.netG_student.training            # left in eval() like the reference's setup
        model.netG_teacher.load_state_dict(fix['teacher_sd'])
        model.netG_student.load_state_dict(student0)
        model.netD.load_state_dict(fix['D_sd0'])
        # ---- first step: eval-mode student (pinned to the real reference's first step by the add-on fixture)
        ref = O.distill_step(st, s0['real_A'], s0['real_B'], dict(fix['hp'], student_training=False))
        model.set_input(batch(s0))
        model.optimize_parameters(0)
        L = model.get_current_losses()
        for mine, theirs in keys:
            r = float(ref[theirs])
            assert abs(L[mine] - r) <= 1e-5 * max(1.0, abs(r)), (mine, L[mine], r)
            assert abs(L[mine] - add['losses'][mine]) <= 1e-4 * max(1.0, abs(r)), mine
        # ---- first evaluate_model: inference in eval mode through the module mirrors, then train()
        model.eval_dataloader = [batch(s1)]
        seen = {}

        def fid(fakes):
            seen['fakes'] = fakes
            return 12.5
        model.metric_fns = {'fid': fid}
        m_before = model.engine.S.arena.m.clone()
        ret = model.evaluate_model(0)
        assert ret == {'metric/fid': 12.5, 'metric/fid-mean': 12.5, 'metric/fid-best': 12.5} and model.is_best
        assert model.netG_student.training
        with torch.no_grad():
            want = O.generator_forward(O.clone_sd(st['student_sd']), fix['student_arch'], s1['real_A'], training=False)
        assert float((seen['fakes'][0] - want).norm() / want.norm()) < 1e-5
        assert torch.equal(model.engine.S.arena.m, m_before)                  # the training engine was not disturbed
        # ---- second step: training-mode student on the same engine and optimiser state
        ref = O.distill_step(st, s1['real_A'], s1['real_B'], dict(fix['hp'], student_training=True))
        model.set_input(batch(s1))
        model.optimize_parameters(1)
        assert model.engine.S.training
        L = model.get_current_losses()
        for mine, theirs in keys:
            r = float(ref[theirs])
            assert abs(L[mine] - r) <= 2e-3 * max(1.0, abs(r)), (mine, L[mine], r)
        ret = model.evaluate_model(1)
        assert ret['metric/fid-mean'] == 12.5 and not model.is_best


@pytest.mark.timeout(900)
def test_spade_distiller_flow(golden_dir, tmp_path):
    """SPADEDistiller mirror in exact emulation: one optimize_parameters against the oracle, then evaluate_model (student
    inference in eval mode through the module mirror on a different batch shape, metric bookkeeping, back to train())."""
    import importlib
    from oracle import spade_oracle as SO
    from oracle.cat_oracle import clone_sd
    from oracle.kernel_emu import emulated_kernels
    _spade_opt = importlib.import_module('test_spade_distiller_gpu')._opt
    fix = torch.load(os.path.join(golden_dir, 'spade_more.pt'), weights_only=False)
    vgg = SO.make_vgg_sd(fix['vgg_seed'])
    s = fix['steps'][0]
    state = dict(teacher_sd=clone_sd(fix['teacher_sd']), student_sd=clone_sd(fix['student_sd0']), D_sd=clone_sd(fix['D_sd0']),
                 vgg_sd=vgg, teacher_arch=fix['teacher_arch'], student_arch=fix['student_arch'], D_arch=fix['D_arch'],
                 adam_G={}, adam_D={})
    seg = SO.preprocess_input(s['label'], s['instance'], fix['hp']['n_label'])
    ref = SO.spade_distill_step(state, seg, s['image'], fix['hp'])
    with emulated_kernels(exact=True):
        from cat_b200.distillers import create_distiller
        opt = _spade_opt(fix, str(tmp_path), vgg)
        opt.cuda_graph = False
        model = create_distiller(opt, verbose=False)
        model.setup(opt, verbose=False)
        mm = model.modules_on_one_gpu
        mm.netG_teacher.load_state_dict(fix['teacher_sd'])
        mm.netG_student.load_state_dict(fix['student_sd0'])
        mm.netD.load_state_dict(fix['D_sd0'])
        mm.netG_student.train()
        B = s['image'].shape[0]
        model.set_input({'label': s['label'], 'instance': s['instance'], 'image': s['image'], 'path': ['x'] * B})
        model.optimize_parameters(0)
        L = model.get_current_losses()
        for mine, theirs in (('G_loss/G_gan', 'loss_G_gan'), ('G_loss/G_feat', 'loss_G_feat'), ('G_loss/G_vgg', 'loss_G_vgg'),
                             ('G_loss/G_distill', 'loss_G_distill'), ('D_loss/D_fake', 'loss_D_fake'), ('D_loss/D_real', 'loss_D_real')):
            r = float(ref[theirs])
            assert abs(L[mine] - r) <= 1e-5 * max(1.0, abs(r)), (mine, L[mine], r)
        model.eval_dataloader = [{'label': s['label'][:1], 'instance': s['instance'][:1], 'image': s['image'][:1], 'path': ['v/munster_1.png']}]
        got = {}
        model.metric_fns = {'fid': lambda fakes: 5.0, 'mIoU': lambda fakes, names: got.setdefault('names', names) and 0.5}
        ret = model.evaluate_model(0)
        assert ret['metric/fid-best'] == 5.0 and ret['metric/mIoU'] == 0.5 and got['names'] == ['munster_1'] and model.is_best
        assert mm.netG_student.training
        want = SO.spade_generator_forward({k: v.detach().clone() for k, v in mm.netG_student.state_dict().items()},
                                          fix['student_arch'], seg[:1], training=False)
        assert float((model.Sfake_B - want).norm() / want.norm()) < 1e-5
