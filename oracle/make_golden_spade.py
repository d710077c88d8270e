"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/spade_*.pt from the real reference SPADEDistiller.

Run in the build container (needs /root/reference):   python -m oracle.make_golden_spade
Produced by the unmodified reference classes (SPADEDistiller, shrink, init_net) driven through
set_input + optimize_parameters (distillers/base_spade_distiller.py:226-234), split open only to snapshot
gradients between the two phases.  The fixtures keep the full first-step state in fp32 and are kept small
(tiny widths, 64x128 images) so that they can be committed.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_harness_spade import (VGG_SEED, build_reference_spade_distiller, multiscale_D_arch,  # noqa: E402
                                      spade_generator_arch)

OUT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')

CASES = {
    # GauGAN-cityscapes-style (scripts/gaugan/cityscapes/train_inception_student_5p6B.sh): spadesyncbatch3x3
    # generators, 'more' up-sampling (latent 1x2 at 64x128), multi-scale spectral-instance D, hinge + feature
    # matching + VGG + KA, TTUR Adam.
    'spade_more': dict(batch_size=2, crop_size=128, aspect_ratio=2.0, teacher_ngf=6, student_ngf=6, ndf=8,
                       input_nc=6, frac=0.4, prune_cin_lb=2),
}


def snap(sd):
    return {k: v.detach().clone() for k, v in sd.items()}


def make_case(name, cfg):
    kw = {k: v for k, v in cfg.items() if k != 'frac'}
    probe, _ = build_reference_spade_distiller(do_shrink=False, **kw)
    target = probe.modules_on_one_gpu.netG_teacher.n_macs * cfg['frac']
    model, opt = build_reference_spade_distiller(target_flops=target, **kw)
    from models import networks
    mm = model.modules_on_one_gpu
    mm.netG_student = networks.init_net(mm.netG_student, opt.init_type, opt.init_gain, []).to(model.device)
    # init_net re-initialises in place, so the optimiser shrink_spade_model re-created (utils/common.py:845-858)
    # still owns the student's parameters
    # the pruned student is a copy of the eval-mode teacher; the reference puts it in train mode at the end of the
    # first evaluate_model (spade_distiller.py:170).  The golden steps are recorded in that steady state.
    mm.netG_student.train()
    g = torch.Generator().manual_seed(7)
    for net in (mm.netG_student, mm.netD):
        for k, p in net.named_parameters():
            if p.dim() == 4:
                p.data = p.data * 2.0
            elif k.endswith('bias'):
                p.data = 0.05 * torch.randn(p.shape, generator=g)
    for m in mm.netG_student.modules():
        if hasattr(m, 'running_mean') and getattr(m, 'weight', None) is not None:
            m.weight.data = 0.5 + torch.rand(m.weight.shape, generator=g)
    B = cfg['batch_size']
    W = cfg['crop_size']
    H = int(round(W / cfg['aspect_ratio']))
    fix = {
        'name': name,
        'teacher_arch': spade_generator_arch(mm.netG_teacher),
        'student_arch': spade_generator_arch(mm.netG_student),
        'D_arch': multiscale_D_arch(mm.netD, opt),
        'hp': dict(lambda_gan=float(opt.lambda_gan), lambda_feat=float(opt.lambda_feat), lambda_vgg=float(opt.lambda_vgg),
                   lambda_distill=float(opt.lambda_distill), lr_G=float(opt.lr) / 2, lr_D=float(opt.lr) * 2,
                   beta1=0.0, beta2=0.9, n_label=int(opt.input_nc)),
        'teacher_sd': snap(mm.netG_teacher.state_dict()),
        'student_sd0': snap(mm.netG_student.state_dict()),
        'D_sd0': snap(mm.netD.state_dict()),
        # VGG19 weights are ~26 MB: not stored.  The reference was built with torchvision vgg19(weights=None) under
        # VGG_SEED (oracle/ref_harness_spade.py); oracle.spade_oracle.make_vgg_sd(seed) regenerates the same tensors
        # and 'vgg_check' pins them.
        'vgg_seed': VGG_SEED,
        'vgg_check': float(sum(v.double().abs().sum() for v in mm.criterionVGG.vgg.state_dict().values())),
        'steps': [],
    }
    gen = torch.Generator().manual_seed(233)
    for it in range(2):
        blk = 8
        lab = torch.randint(0, cfg['input_nc'], (B, 1, H // blk, W // blk), generator=gen)
        lab = lab.repeat_interleave(blk, 2).repeat_interleave(blk, 3).float()
        inst = torch.randint(0, 8, (B, 1, H // blk, W // blk), generator=gen)
        inst = inst.repeat_interleave(blk, 2).repeat_interleave(blk, 3)
        img = torch.rand(B, 3, H, W, generator=gen) * 2 - 1
        data = {'label': lab.clone(), 'instance': inst.clone(), 'image': img.clone(), 'path': ['x'] * B}
        model.set_input(data)
        seg = model.input_semantics.detach().clone()
        # optimize_parameters (base_spade_distiller.py:226-234)
        model.set_requires_grad(mm.netD, False)
        model.optimizer_G.zero_grad()
        model.backward_G()
        S_grads = {k: p.grad.detach().clone() for k, p in mm.netG_student.named_parameters() if p.grad is not None}
        model.optimizer_G.step()
        model.set_requires_grad(mm.netD, True)
        model.optimizer_D.zero_grad()
        model.backward_D()
        D_grads = {k: p.grad.detach().clone() for k, p in mm.netD.named_parameters()}
        model.optimizer_D.step()
        step = {
            'label': lab, 'instance': inst, 'image': img, 'seg': seg.to(torch.uint8),
            'losses': {k: float(v) for k, v in model.get_current_losses().items()},
        }
        if it == 0:
            step.update({'S_grads': S_grads, 'D_grads': D_grads,
                         'student_buffers_after': snap(dict(mm.netG_student.named_buffers())),
                         'D_buffers_after': snap(dict(mm.netD.named_buffers())),
                         'student_param_checksum_after': float(sum(p.double().abs().sum() for p in mm.netG_student.parameters())),
                         'D_param_checksum_after': float(sum(p.double().abs().sum() for p in mm.netD.parameters()))})
        fix['steps'].append(step)
    return fix


def make_mse_addon(base_name, cfg):
    """--distill_G_loss_type mse on the SAME seeded networks as `base_name` (the adaptor convs are created in both modes,
    so the random streams agree): a small add-on fixture with the adaptors, the first-step losses and the gradients the
    'mse' terms reach (adaptors + student); the networks themselves are taken from the base fixture."""
    kw = {k: v for k, v in cfg.items() if k != 'frac'}
    probe, _ = build_reference_spade_distiller(do_shrink=False, **kw)
    target = probe.modules_on_one_gpu.netG_teacher.n_macs * cfg['frac']
    model, opt = build_reference_spade_distiller(target_flops=target, distill_G_loss_type='mse', lambda_distill=2.0, **kw)
    from models import networks
    mm = model.modules_on_one_gpu
    mm.netG_student = networks.init_net(mm.netG_student, opt.init_type, opt.init_gain, []).to(model.device)
    mm.netG_student.train()
    g = torch.Generator().manual_seed(7)
    for net in (mm.netG_student, mm.netD):
        for k, p in net.named_parameters():
            if p.dim() == 4:
                p.data = p.data * 2.0
            elif k.endswith('bias'):
                p.data = 0.05 * torch.randn(p.shape, generator=g)
    for m in mm.netG_student.modules():
        if hasattr(m, 'running_mean') and getattr(m, 'weight', None) is not None:
            m.weight.data = 0.5 + torch.rand(m.weight.shape, generator=g)
    base = torch.load(os.path.join(OUT_DIR, base_name + '.pt'), weights_only=False)
    for k, v in mm.netG_student.state_dict().items():
        assert torch.equal(v, base['student_sd0'][k]), ('student differs from the base fixture', k)
    for k, v in mm.netD.state_dict().items():
        assert torch.equal(v, base['D_sd0'][k]), ('D differs from the base fixture', k)
    s0 = base['steps'][0]
    B = cfg['batch_size']
    fix = {'name': base_name + '_mse', 'base': base_name, 'lambda_distill': float(opt.lambda_distill),
           'netA_sd0': [snap(net.state_dict()) for net in mm.netAs]}
    model.set_input({'label': s0['label'].clone(), 'instance': s0['instance'].clone(), 'image': s0['image'].clone(), 'path': ['x'] * B})
    model.set_requires_grad(mm.netD, False)
    model.optimizer_G.zero_grad()
    model.backward_G()
    fix['netA_grads'] = [{k: p.grad.detach().clone() for k, p in net.named_parameters()} for net in mm.netAs]
    fix['S_grads'] = {k: p.grad.detach().clone() for k, p in mm.netG_student.named_parameters() if p.grad is not None}
    model.optimizer_G.step()
    fix['netA_sd_after'] = [snap(net.state_dict()) for net in mm.netAs]
    model.set_requires_grad(mm.netD, True)
    model.optimizer_D.zero_grad()
    model.backward_D()
    model.optimizer_D.step()
    fix['losses'] = {k: float(v) for k, v in model.get_current_losses().items()}
    return fix


def make_first_step_addon(base_name, cfg):
    """The reference's FIRST optimize_parameters of a run: BaseSPADEDistiller.setup profiles the generators
    (base_spade_distiller.py:178-190) and model_profiling leaves them in eval() (utils/model_profiling.py:299); the student
    returns to train() only at the end of the first evaluate_model (spade_distiller.py:170).  Same seeded networks as
    `base_name` (asserted) without make_case's .train() call: losses, student gradients (the conv biases in front of a
    BatchNorm are live in eval mode) and the D-phase image of that step."""
    kw = {k: v for k, v in cfg.items() if k != 'frac'}
    probe, _ = build_reference_spade_distiller(do_shrink=False, **kw)
    target = probe.modules_on_one_gpu.netG_teacher.n_macs * cfg['frac']
    model, opt = build_reference_spade_distiller(target_flops=target, **kw)
    from models import networks
    mm = model.modules_on_one_gpu
    mm.netG_student = networks.init_net(mm.netG_student, opt.init_type, opt.init_gain, []).to(model.device)
    assert not mm.netG_student.training, 'the reference is expected to leave the pruned student in eval mode'
    g = torch.Generator().manual_seed(7)
    for net in (mm.netG_student, mm.netD):
        for k, p in net.named_parameters():
            if p.dim() == 4:
                p.data = p.data * 2.0
            elif k.endswith('bias'):
                p.data = 0.05 * torch.randn(p.shape, generator=g)
    for m in mm.netG_student.modules():
        if hasattr(m, 'running_mean') and getattr(m, 'weight', None) is not None:
            m.weight.data = 0.5 + torch.rand(m.weight.shape, generator=g)
    base = torch.load(os.path.join(OUT_DIR, base_name + '.pt'), weights_only=False)
    for k, v in mm.netG_student.state_dict().items():
        assert torch.equal(v, base['student_sd0'][k]), ('student differs from the base fixture', k)
    gs = torch.Generator().manual_seed(11)
    stats = {}
    for k, v in mm.netG_student.state_dict().items():
        if k.endswith('running_mean'):
            stats[k] = 0.1 * torch.randn(v.shape, generator=gs)
        elif k.endswith('running_var'):
            stats[k] = 0.5 + torch.rand(v.shape, generator=gs)
    mm.netG_student.load_state_dict(stats, strict=False)
    s0 = base['steps'][0]
    B = cfg['batch_size']
    model.set_input({'label': s0['label'].clone(), 'instance': s0['instance'].clone(), 'image': s0['image'].clone(), 'path': ['x'] * B})
    model.set_requires_grad(mm.netD, False)
    model.optimizer_G.zero_grad()
    model.backward_G()
    fix = {'name': base_name + '_first_step', 'base': base_name, 'running_stats': stats,
           'S_grads': {k: p.grad.detach().clone() for k, p in mm.netG_student.named_parameters() if p.grad is not None}}
    model.optimizer_G.step()
    model.set_requires_grad(mm.netD, True)
    model.optimizer_D.zero_grad()
    model.backward_D()
    model.optimizer_D.step()
    fix['losses'] = {k: float(v) for k, v in model.get_current_losses().items()}
    fix['running_stats_after'] = {k: v.detach().clone() for k, v in mm.netG_student.state_dict().items() if 'running_' in k}
    return fix


def main():
    os.makedirs(OUT_DIR, exist_ok=True)
    only = sys.argv[1:]
    if only == ['spade_more_first_step']:
        fix = make_first_step_addon('spade_more', dict(CASES['spade_more']))
        path = os.path.join(OUT_DIR, 'spade_more_first_step.pt')
        torch.save(fix, path)
        print('spade_more_first_step', fix['losses'], '-> %.2f MB' % (os.path.getsize(path) / 1e6))
        return
    if only == ['spade_more_mse']:
        fix = make_mse_addon('spade_more', dict(CASES['spade_more']))
        path = os.path.join(OUT_DIR, 'spade_more_mse.pt')
        torch.save(fix, path)
        print('spade_more_mse', fix['losses'], '-> %.2f MB' % (os.path.getsize(path) / 1e6))
        return
    for name, cfg in CASES.items():
        if only and name not in only:
            continue
        fix = make_case(name, dict(cfg))
        path = os.path.join(OUT_DIR, name + '.pt')
        torch.save(fix, path)
        print(name, {n: (b['res'], b['dw'], b['spade_res'], b['spade_dw']) for n, b in fix['student_arch']['blocks'].items()},
              'losses', fix['steps'][0]['losses'], '-> %.2f MB' % (os.path.getsize(path) / 1e6))


if __name__ == '__main__':
    main()
