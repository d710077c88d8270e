"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.pt from the real reference.

Run in the build container (needs /root/reference):   python -m oracle.make_golden
The vectors pin oracle/cat_oracle.py (tests/test_oracle_golden.py) and, through it, the CUDA path.
They are produced by the unmodified reference classes (InceptionDistiller, shrink, init_net) driven
through set_input + optimize_parameters, exactly as trainer.py:128-133 does.
"""
import copy
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_harness import build_reference_distiller, discriminator_arch, generator_arch  # noqa: E402

OUT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')

CASES = {
    # pix2pix-style (scripts/pix2pix/cityscapes/train_inception_student_5p6B.sh): BatchNorm with
    # running stats, aligned pairs, hinge GAN, 6-channel D input.
    'pix2pix_bn_hinge': dict(norm='batch', gan_mode='hinge', dataset_mode='aligned',
                             lambda_distill=0.5, lambda_recon=100.0, batch_size=3, height=32,
                             width=32, frac=0.2),
    # CycleGAN-style (scripts/cycle_gan/horse2zebra/train_inception_student_2p6B.sh):
    # InstanceNorm (affine), unaligned, lsgan, recon against the teacher output, lambda_recon 5.
    'cyclegan_in_lsgan': dict(norm='instance', gan_mode='lsgan', dataset_mode='unaligned',
                              lambda_distill=1.0, lambda_recon=5.0, batch_size=2, height=32,
                              width=48, frac=0.25),
    # smooth losses everywhere (--recon_loss_type l2, lsgan): gradients are Lipschitz in the forward
    # values, so this case pins the whole backward pass tightly (the sign() of L1 and the hinge mask
    # amplify rounding differences of the forward pass).
    'pix2pix_bn_lsgan_l2': dict(norm='batch', gan_mode='lsgan', dataset_mode='aligned', lambda_distill=0.5,
                                lambda_recon=100.0, batch_size=4, height=32, width=32, frac=0.2,
                                recon_loss_type='l2'),
    # --distill_G_loss_type mse (inception_distiller.py:111-133): MSE(netA_i(Sact_i), Tact_i) through the 1x1 adaptor convs
    # netAs, which are parameters of optimizer_G; smooth losses so that the backward pass is pinned tightly.
    'pix2pix_bn_mse': dict(norm='batch', gan_mode='lsgan', dataset_mode='aligned', lambda_distill=2.0, lambda_recon=100.0,
                           batch_size=3, height=32, width=32, frac=0.2, recon_loss_type='l2', distill_G_loss_type='mse'),
}


def snap(sd):
    return {k: v.detach().clone() for k, v in sd.items()}


def make_case(name, cfg):
    probe, _ = build_reference_distiller(norm=cfg['norm'], teacher_ngf=12, student_ngf=8, ndf=8,
                                         height=cfg['height'], width=cfg['width'],
                                         batch_size=cfg['batch_size'], do_shrink=False)
    target = probe.netG_teacher.n_macs * cfg['frac']
    model, opt = build_reference_distiller(norm=cfg['norm'], teacher_ngf=12, student_ngf=8, ndf=8,
                                           height=cfg['height'], width=cfg['width'],
                                           batch_size=cfg['batch_size'], target_flops=target,
                                           gan_mode=cfg['gan_mode'],
                                           dataset_mode=cfg['dataset_mode'],
                                           lambda_distill=cfg['lambda_distill'],
                                           lambda_recon=cfg['lambda_recon'],
                                           recon_loss_type=cfg.get('recon_loss_type', 'l1'),
                                           distill_G_loss_type=cfg.get('distill_G_loss_type', 'ka'))
    # after the first evaluate_model the reference puts the student back in train mode
    # (inception_distiller.py:280); the golden steps are recorded in that steady state.
    model.netG_student.train()
    # D/student weights: larger than N(0,0.02) so that activations/gradients are well scaled.
    g = torch.Generator().manual_seed(7)
    for net in (model.netG_student, model.netD):
        for m in net.modules():
            if isinstance(m, (torch.nn.Conv2d, torch.nn.ConvTranspose2d)):
                m.weight.data = m.weight.data * 5.0
                if m.bias is not None:
                    m.bias.data = 0.05 * torch.randn(m.bias.shape, generator=g)
    B, H, W = cfg['batch_size'], cfg['height'], cfg['width']
    d_in = 6 if cfg['dataset_mode'] == 'aligned' else 3
    fix = {
        'name': name,
        'teacher_arch': generator_arch(model.netG_teacher, opt),
        'student_arch': generator_arch(model.netG_student, opt),
        'D_arch': discriminator_arch(model.netD, opt, d_in),
        'hp': dict(gan_mode=opt.gan_mode, aligned=opt.dataset_mode == 'aligned',
                   lambda_recon=float(opt.lambda_recon), lambda_gan=float(opt.lambda_gan),
                   lambda_distill=float(opt.lambda_distill), lr=float(opt.lr),
                   beta1=float(opt.beta1), student_training=True, recon_loss_type=opt.recon_loss_type,
                   distill_loss_type=opt.distill_G_loss_type),
        'teacher_sd': snap(model.netG_teacher.state_dict()),
        'student_sd0': snap(model.netG_student.state_dict()),
        'D_sd0': snap(model.netD.state_dict()),
        'steps': [],
    }
    if opt.distill_G_loss_type == 'mse':
        for net in model.netAs:      # default Conv2d init is U(+-1/sqrt(fan_in)): keep, add a spread to the biases
            net.bias.data = 0.05 * torch.randn(net.bias.shape, generator=g)
        fix['netA_sd0'] = [snap(net.state_dict()) for net in model.netAs]
    gen = torch.Generator().manual_seed(233)
    for it in range(2):
        A = torch.rand(B, 3, H, W, generator=gen) * 2 - 1
        Bt = torch.rand(B, 3, H, W, generator=gen) * 2 - 1
        data = {'A': A, 'B': Bt, 'A_paths': ['x'] * B, 'B_paths': ['x'] * B}
        model.set_input(data)
        # optimize_parameters, split open only to snapshot gradients between the phases
        model.forward()
        model.set_requires_grad(model.netD, True)
        model.optimizer_D.zero_grad()
        model.backward_D()
        D_grads = {k: p.grad.detach().clone() for k, p in model.netD.named_parameters()}
        model.optimizer_D.step()
        model.set_requires_grad(model.netD, False)
        model.optimizer_G.zero_grad()
        Sacts = {k.replace('cpu', ''): v for k, v in model.Sacts.items()}
        for v in Sacts.values():
            v.retain_grad()
        model.backward_G(it)
        S_grads = {k: p.grad.detach().clone() for k, p in model.netG_student.named_parameters()}
        A_grads = [{k: p.grad.detach().clone() for k, p in net.named_parameters() if p.grad is not None} for net in model.netAs]
        Sact_grads = {k: v.grad.detach().clone() for k, v in Sacts.items()}
        model.optimizer_G.step()
        step = {
            'real_A': A, 'real_B': Bt,
            'losses': {k: float(v) for k, v in model.get_current_losses().items()},
            'student_sd_after': snap(model.netG_student.state_dict()),
            'D_sd_after': snap(model.netD.state_dict()),
        }
        if it == 0:
            step.update({
                'Tfake_B': model.Tfake_B.detach().clone(), 'Sfake_B': model.Sfake_B.detach().clone(),
                'Tacts': {k.replace('cpu', ''): v.detach().clone() for k, v in model.Tacts.items()},
                'Sacts': {k: v.detach().clone() for k, v in Sacts.items()},
                'Sact_grads': Sact_grads, 'S_grads': S_grads, 'D_grads': D_grads,
            })
            if opt.distill_G_loss_type == 'mse':
                step.update(netA_grads=A_grads, netA_sd_after=[snap(net.state_dict()) for net in model.netAs])
        fix['steps'].append(step)
    return fix


def make_first_step_addon(base_name):
    """The reference's FIRST optimize_parameters of a run: model_profiling leaves the pruned student in eval()
    (utils/model_profiling.py:299) and it only returns to train() at the end of the first evaluate_model
    (inception_distiller.py:280, trainer.py:141), so that one step normalises with the BatchNorm running statistics and
    back-propagates through them.  Same seeded networks as `base_name` (asserted), without the .train() call of make_case:
    a small add-on fixture with the losses and gradients of that step."""
    cfg = copy.deepcopy(CASES[base_name])
    probe, _ = build_reference_distiller(norm=cfg['norm'], teacher_ngf=12, student_ngf=8, ndf=8, height=cfg['height'],
                                         width=cfg['width'], batch_size=cfg['batch_size'], do_shrink=False)
    target = probe.netG_teacher.n_macs * cfg['frac']
    model, opt = build_reference_distiller(norm=cfg['norm'], teacher_ngf=12, student_ngf=8, ndf=8, height=cfg['height'],
                                           width=cfg['width'], batch_size=cfg['batch_size'], target_flops=target,
                                           gan_mode=cfg['gan_mode'], dataset_mode=cfg['dataset_mode'],
                                           lambda_distill=cfg['lambda_distill'], lambda_recon=cfg['lambda_recon'],
                                           recon_loss_type=cfg.get('recon_loss_type', 'l1'))
    assert not model.netG_student.training, 'the reference is expected to leave the pruned student in eval mode'
    g = torch.Generator().manual_seed(7)
    for net in (model.netG_student, model.netD):
        for m in net.modules():
            if isinstance(m, (torch.nn.Conv2d, torch.nn.ConvTranspose2d)):
                m.weight.data = m.weight.data * 5.0
                if m.bias is not None:
                    m.bias.data = 0.05 * torch.randn(m.bias.shape, generator=g)
    base = torch.load(os.path.join(OUT_DIR, base_name + '.pt'), weights_only=False)
    for k, v in model.netG_student.state_dict().items():
        assert torch.equal(v, base['student_sd0'][k]), ('student differs from the base fixture', k)
    # non-trivial running statistics (a freshly initialised student has mean 0 / var 1)
    gs = torch.Generator().manual_seed(11)
    stats = {}
    for k, v in model.netG_student.state_dict().items():
        if k.endswith('running_mean'):
            stats[k] = 0.1 * torch.randn(v.shape, generator=gs)
        elif k.endswith('running_var'):
            stats[k] = 0.5 + torch.rand(v.shape, generator=gs)
    model.netG_student.load_state_dict(stats, strict=False)
    s0 = base['steps'][0]
    B = cfg['batch_size']
    model.set_input({'A': s0['real_A'].clone(), 'B': s0['real_B'].clone(), 'A_paths': ['x'] * B, 'B_paths': ['x'] * B})
    model.forward()
    model.set_requires_grad(model.netD, True)
    model.optimizer_D.zero_grad()
    model.backward_D()
    model.optimizer_D.step()
    model.set_requires_grad(model.netD, False)
    model.optimizer_G.zero_grad()
    model.backward_G(0)
    fix = {'name': base_name + '_first_step', 'base': base_name, 'running_stats': stats,
           'Sfake_B': model.Sfake_B.detach().clone(),
           'S_grads': {k: p.grad.detach().clone() for k, p in model.netG_student.named_parameters()},
           'losses': None}
    model.optimizer_G.step()
    fix['losses'] = {k: float(v) for k, v in model.get_current_losses().items()}
    fix['running_stats_after'] = {k: v.detach().clone() for k, v in model.netG_student.state_dict().items() if 'running_' in k}
    return fix


def main():
    os.makedirs(OUT_DIR, exist_ok=True)
    only = sys.argv[1:]
    if only and only[0].endswith('_first_step'):
        fix = make_first_step_addon(only[0][:-len('_first_step')])
        path = os.path.join(OUT_DIR, only[0] + '.pt')
        torch.save(fix, path)
        print(only[0], fix['losses'], '-> %.2f MB' % (os.path.getsize(path) / 1e6))
        return
    for name, cfg in CASES.items():
        if only and name not in only:
            continue
        fix = make_case(name, copy.deepcopy(cfg))
        path = os.path.join(OUT_DIR, name + '.pt')
        torch.save(fix, path)
        print(name, 'student widths', fix['student_arch']['widths'], 'blocks',
              fix['student_arch']['blocks'][:2], 'losses', fix['steps'][0]['losses'],
              '-> %.2f MB' % (os.path.getsize(path) / 1e6))


if __name__ == '__main__':
    main()
