"""The reference-facing SPADE protocol (cat_b200.distillers.SPADEDistiller / cat_b200.models.spade_networks) on the GPU:
a trainer-style loop (create_distiller -> setup -> set_input -> optimize_parameters -> get_current_losses ->
save_networks) must produce the oracle's losses, keep the module parameters aliased to the engine arenas and write
checkpoints with the reference's file names and state_dict keys."""
import argparse
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]


def _opt(fix, log_dir, vgg):
    hp, Ta, Da = fix['hp'], fix['teacher_arch'], fix['D_arch']
    ngf = Ta['fc_out'] // 16
    return argparse.Namespace(
        isTrain=True, gpu_ids=[0], log_dir=log_dir, distiller='spade', input_nc=hp['n_label'], output_nc=3,
        semantic_nc=Ta['semantic_nc'], teacher_ngf=ngf, student_ngf=ngf, ngf=ngf, teacher_netG='inception_spade',
        student_netG='inception_spade', teacher_norm_G='spadesyncbatch3x3', student_norm_G='spadesyncbatch3x3', norm_G='spadesyncbatch3x3',
        norm='instance', num_upsampling_layers=Ta['num_upsampling_layers'], crop_size=128, aspect_ratio=2.0, channels=None,
        channels_reduction_factor=6, kernel_sizes=[1, 3, 5], active_fn='nn.ReLU', init_type='xavier', init_gain=0.02,
        netD='multi_scale', ndf=Da['ndf'], n_layers_D=Da['n_layers'], num_D=Da['num_D'], norm_D=Da['norm_D'], gan_mode='hinge',
        distill_G_loss_type='ka', lambda_gan=hp['lambda_gan'], lambda_feat=hp['lambda_feat'], lambda_vgg=hp['lambda_vgg'],
        lambda_distill=hp['lambda_distill'], lr=hp['lr_G'] * 2, beta1=0.5, beta2=0.999, no_TTUR=False, nepochs=5, nepochs_decay=15,
        student_arch=fix['student_arch'], restore_teacher_G_path=None, restore_student_G_path=None, restore_D_path=None,
        cuda_graph=True, vgg_state_dict=vgg)


def test_spade_trainer_style_loop(golden_dir, tmp_path):
    from cat_b200.distillers import create_distiller
    from oracle import spade_oracle as SO
    from oracle.cat_oracle import clone_sd
    fix = torch.load(os.path.join(golden_dir, 'spade_more.pt'), weights_only=False)
    vgg = SO.make_vgg_sd(fix['vgg_seed'])
    opt = _opt(fix, str(tmp_path), vgg)
    model = create_distiller(opt, verbose=False)
    model.setup(opt, verbose=False)
    mm = model.modules_on_one_gpu
    mm.netG_teacher.load_state_dict(fix['teacher_sd'])       # reference checkpoints load straight into the module trees
    mm.netG_student.load_state_dict(fix['student_sd0'])
    mm.netD.load_state_dict(fix['D_sd0'])
    mm.netG_student.train()
    state = dict(teacher_sd=clone_sd(fix['teacher_sd']), student_sd=clone_sd(fix['student_sd0']), D_sd=clone_sd(fix['D_sd0']),
                 vgg_sd=vgg, teacher_arch=fix['teacher_arch'], student_arch=fix['student_arch'], D_arch=fix['D_arch'],
                 adam_G={}, adam_D={})
    w_before = mm.netG_student.state_dict()['conv_img.weight'].clone()
    for it, s in enumerate(fix['steps']):
        seg = SO.preprocess_input(s['label'], s['instance'], fix['hp']['n_label'])
        ref = SO.spade_distill_step(state, seg, s['image'], fix['hp'])
        B = s['image'].shape[0]
        model.set_input({'label': s['label'], 'instance': s['instance'], 'image': s['image'], 'path': ['x'] * B})
        model.optimize_parameters(it)
        L = model.get_current_losses()
        assert list(L.keys()) == ['G_loss/G_gan', 'G_loss/G_feat', 'G_loss/G_vgg', 'G_loss/G_distill', 'D_loss/D_real',
                                  'D_loss/D_fake', 'Specific_loss/G_distill0', 'Specific_loss/G_distill1', 'Specific_loss/G_distill2']
        for mine, theirs in (('G_loss/G_gan', 'loss_G_gan'), ('G_loss/G_feat', 'loss_G_feat'), ('G_loss/G_vgg', 'loss_G_vgg'),
                             ('G_loss/G_distill', 'loss_G_distill'), ('D_loss/D_fake', 'loss_D_fake'), ('D_loss/D_real', 'loss_D_real')):
            r = float(ref[theirs])
            assert abs(L[mine] - r) <= 6e-2 * max(1.0, abs(r)), (it, mine, L[mine], r)
    # the module tree sees the trained weights (aliased storage) ...
    w_after = mm.netG_student.state_dict()['conv_img.weight']
    assert float((w_after - w_before).abs().max()) > 0
    assert w_after.data_ptr() == model.engine.S.arena.view('conv_img.weight').data_ptr()
    # ... and checkpoints carry the reference's names and keys
    model.save_networks('latest')
    ck = torch.load(os.path.join(str(tmp_path), 'checkpoints', 'latest_net_G.pth'), map_location='cpu')
    assert list(ck.keys()) == list(fix['student_sd0'].keys())
    ckd = torch.load(os.path.join(str(tmp_path), 'checkpoints', 'latest_net_D.pth'), map_location='cpu')
    assert list(ckd.keys()) == list(fix['D_sd0'].keys())
    u0 = fix['D_sd0']['discriminator_0.model1.0.0.weight_u']
    assert float((ckd['discriminator_0.model1.0.0.weight_u'] - u0).abs().max()) > 0      # power iterations ran
    for f in ('latest_net_A-0.pth', 'latest_optim-0.pth', 'latest_optim-1.pth'):
        assert os.path.exists(os.path.join(str(tmp_path), 'checkpoints', f))
    model.update_learning_rate()
    assert abs(float(model.engine.lr_D) - model.optimizer_D.param_groups[0]['lr']) < 1e-9      # device copy is fp32


def test_spade_inference_paths(golden_dir, tmp_path):
    """SPADEDistiller.test() (generate_fake) and InceptionSPADEGenerator.forward (the generator inference of
    evaluate_model) on the GPU against the oracle's eval-mode teacher."""
    from cat_b200.distillers import create_distiller
    from oracle import spade_oracle as SO
    from oracle.cat_oracle import clone_sd
    fix = torch.load(os.path.join(golden_dir, 'spade_more.pt'), weights_only=False)
    vgg = SO.make_vgg_sd(fix['vgg_seed'])
    model = create_distiller(_opt(fix, str(tmp_path), vgg), verbose=False)
    mm = model.modules_on_one_gpu
    mm.netG_teacher.load_state_dict(fix['teacher_sd'])
    mm.netG_student.load_state_dict(fix['student_sd0'])
    mm.netD.load_state_dict(fix['D_sd0'])
    s = fix['steps'][0]
    B = s['image'].shape[0]
    model.set_input({'label': s['label'], 'instance': s['instance'], 'image': s['image'], 'path': ['x'] * B})
    model.test()
    seg = SO.preprocess_input(s['label'], s['instance'], fix['hp']['n_label'])
    ref = SO.spade_generator_forward(clone_sd(fix['teacher_sd']), fix['teacher_arch'], seg, training=False)
    assert torch.equal(model.input_semantics.cpu(), seg)
    err = float((model.Tfake_B.cpu() - ref).norm() / ref.norm())
    assert err < 3e-2, err
    assert model.Sfake_B.shape == ref.shape and bool(torch.isfinite(model.Sfake_B).all())
    # the module mirror runs the same compiled network (bound to the distiller's arenas) ...
    mm.netG_teacher.eval()
    out = mm.netG_teacher(seg.cuda())
    assert float((out - model.Tfake_B).abs().max()) < 1e-6
    # ... and compiles a second shape against the same weights
    out1 = mm.netG_teacher(seg[:1].cuda())
    assert float((out1.cpu() - ref[:1]).norm() / ref[:1].norm()) < 3e-2
