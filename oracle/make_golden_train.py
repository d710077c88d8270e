"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/train_*.pt from the real reference training models.

Run in the build container (needs /root/reference):   python -m oracle.make_golden_train
Produced by the unmodified reference classes (Pix2PixModel, CycleGANModel, SPADEModel) driven through
set_input + optimize_parameters exactly as trainer.py:128-133 does, split open only to snapshot gradients between
the phases.  Fixtures keep the loaded state, the inputs, the losses, the first-step gradients and a checksum of the
updated parameters; they are tiny (ngf 8 / 6) so that they can be committed.
"""
import os
import random
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_harness import discriminator_arch, generator_arch  # noqa: E402
from oracle.ref_harness_spade import VGG_SEED, multiscale_D_arch, spade_generator_arch  # noqa: E402
from oracle.ref_harness_train import (build_reference_cyclegan, build_reference_pix2pix,  # noqa: E402
                                      build_reference_spade)

OUT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def snap(sd):
    return {k: v.detach().clone() for k, v in sd.items()}


def grads_of(net):
    return {k: p.grad.detach().clone() for k, p in net.named_parameters() if p.grad is not None}


def checksum(net):
    return float(sum(p.double().abs().sum() for p in net.parameters()))


def rescale(nets, seed, gain):
    """Larger than N(0, 0.02) weights + non-zero biases so that activations and gradients are well scaled."""
    g = torch.Generator().manual_seed(seed)
    for net in nets:
        for m in net.modules():
            if isinstance(m, (torch.nn.Conv2d, torch.nn.ConvTranspose2d)):
                m.weight.data = m.weight.data * gain
                if m.bias is not None:
                    m.bias.data = 0.05 * torch.randn(m.bias.shape, generator=g)
            elif isinstance(m, (torch.nn.BatchNorm2d, torch.nn.InstanceNorm2d)) and m.weight is not None:
                m.weight.data = 0.5 + torch.rand(m.weight.shape, generator=g)
                m.bias.data = 0.1 * torch.randn(m.bias.shape, generator=g)


def make_pix2pix(name, norm, gan_mode, recon, B, H, W):
    model, opt = build_reference_pix2pix(norm=norm, batch_size=B, gan_mode=gan_mode, recon_loss_type=recon)
    rescale((model.netG, model.netD), 7, 5.0)
    fix = {'name': name, 'G_arch': generator_arch(model.netG, opt), 'D_arch': discriminator_arch(model.netD, opt, 6),
           'hp': dict(gan_mode=opt.gan_mode, aligned=True, lambda_recon=float(opt.lambda_recon), lambda_gan=float(opt.lambda_gan),
                      lambda_distill=0.0, lr=float(opt.lr), beta1=float(opt.beta1), recon_loss_type=opt.recon_loss_type),
           'G_sd0': snap(model.netG.state_dict()), 'D_sd0': snap(model.netD.state_dict()), 'steps': []}
    gen = torch.Generator().manual_seed(233)
    for it in range(2):
        A = torch.rand(B, 3, H, W, generator=gen) * 2 - 1
        Bt = torch.rand(B, 3, H, W, generator=gen) * 2 - 1
        model.set_input({'A': A, 'B': Bt, 'A_paths': ['x'] * B, 'B_paths': ['x'] * B})
        model.forward()
        model.set_requires_grad(model.netD, True)
        model.optimizer_D.zero_grad()
        model.backward_D()
        D_grads = grads_of(model.netD)
        model.optimizer_D.step()
        model.set_requires_grad(model.netD, False)
        model.optimizer_G.zero_grad()
        model.backward_G()
        G_grads = grads_of(model.netG)
        model.optimizer_G.step()
        step = {'real_A': A, 'real_B': Bt, 'losses': {k.split('/')[-1]: float(v) for k, v in model.get_current_losses().items()},
                'G_checksum_after': checksum(model.netG), 'D_checksum_after': checksum(model.netD)}
        if it == 0:
            step.update(fake_B=model.fake_B.detach().clone(), G_grads=G_grads, D_grads=D_grads,
                        G_sd_after=snap(model.netG.state_dict()), D_sd_after=snap(model.netD.state_dict()))
        fix['steps'].append(step)
    return fix


def make_cyclegan(name, norm, gan_mode, B, H, W, pool_size, n_steps, data_seed=233):
    model, opt = build_reference_cyclegan(norm=norm, batch_size=B, gan_mode=gan_mode, pool_size=pool_size)
    nets = (model.netG_A, model.netG_B, model.netD_A, model.netD_B)
    rescale(nets, 7, 5.0)
    fix = {'name': name, 'G_arch': generator_arch(model.netG_A, opt), 'D_arch': discriminator_arch(model.netD_A, opt, 3),
           'hp': dict(gan_mode=opt.gan_mode, lambda_A=float(opt.lambda_A), lambda_B=float(opt.lambda_B),
                      lambda_identity=float(opt.lambda_identity), lr=float(opt.lr), beta1=float(opt.beta1),
                      pool_size=int(opt.pool_size)),
           'python_random_seed': 4321,
           'G_A_sd0': snap(model.netG_A.state_dict()), 'G_B_sd0': snap(model.netG_B.state_dict()),
           'D_A_sd0': snap(model.netD_A.state_dict()), 'D_B_sd0': snap(model.netD_B.state_dict()), 'steps': []}
    random.seed(fix['python_random_seed'])       # ImagePool draws from Python's global generator (utils/image_pool.py:41-44)
    gen = torch.Generator().manual_seed(data_seed)
    for it in range(n_steps):
        A = torch.rand(B, 3, H, W, generator=gen) * 2 - 1
        Bt = torch.rand(B, 3, H, W, generator=gen) * 2 - 1
        model.set_input({'A': A, 'B': Bt, 'A_paths': ['x'] * B, 'B_paths': ['x'] * B})
        model.forward()
        model.set_requires_grad([model.netD_A, model.netD_B], False)
        model.optimizer_G.zero_grad()
        model.backward_G()
        G_A_grads, G_B_grads = grads_of(model.netG_A), grads_of(model.netG_B)
        model.optimizer_G.step()
        model.set_requires_grad([model.netD_A, model.netD_B], True)
        model.optimizer_D.zero_grad()
        model.backward_D_A()
        model.backward_D_B()
        D_A_grads, D_B_grads = grads_of(model.netD_A), grads_of(model.netD_B)
        model.optimizer_D.step()
        step = {'real_A': A, 'real_B': Bt, 'losses': {k.split('/')[-1]: float(v) for k, v in model.get_current_losses().items()},
                'checksums_after': {n: checksum(getattr(model, 'net' + n)) for n in ('G_A', 'G_B', 'D_A', 'D_B')}}
        if it == 0:
            step.update(fake_B=model.fake_B.detach().clone(), fake_A=model.fake_A.detach().clone(),
                        rec_A=model.rec_A.detach().clone(), rec_B=model.rec_B.detach().clone(),
                        G_A_grads=G_A_grads, G_B_grads=G_B_grads, D_A_grads=D_A_grads, D_B_grads=D_B_grads,
                        G_A_buffers_after=snap(dict(model.netG_A.named_buffers())))
        fix['steps'].append(step)
    return fix


def make_spade(name, B, crop, aspect, input_nc):
    model, opt = build_reference_spade(batch_size=B, crop_size=crop, aspect_ratio=aspect, input_nc=input_nc)
    mm = model.modules_on_one_gpu
    g = torch.Generator().manual_seed(7)
    for net in (mm.netG, mm.netD):
        for k, p in net.named_parameters():
            if p.dim() == 4:
                p.data = p.data * 2.0
            elif k.endswith('bias'):
                p.data = 0.05 * torch.randn(p.shape, generator=g)
    for m in mm.netG.modules():
        if hasattr(m, 'running_mean') and getattr(m, 'weight', None) is not None:
            m.weight.data = 0.5 + torch.rand(m.weight.shape, generator=g)
    W = crop
    H = int(round(W / aspect))
    fix = {'name': name, 'G_arch': spade_generator_arch(mm.netG), 'D_arch': multiscale_D_arch(mm.netD, opt),
           'hp': dict(lambda_gan=float(opt.lambda_gan), lambda_feat=float(opt.lambda_feat), lambda_vgg=float(opt.lambda_vgg),
                      lambda_distill=0.0, lr_G=float(opt.lr) / 2, lr_D=float(opt.lr) * 2, beta1=0.0, beta2=0.9,
                      n_label=int(opt.input_nc)),
           'G_sd0': snap(mm.netG.state_dict()), 'D_sd0': snap(mm.netD.state_dict()),
           'vgg_seed': VGG_SEED,
           'vgg_check': float(sum(v.double().abs().sum() for v in mm.criterionVGG.vgg.state_dict().values())),
           'steps': []}
    gen = torch.Generator().manual_seed(233)
    for it in range(2):
        blk = 8
        lab = torch.randint(0, input_nc, (B, 1, H // blk, W // blk), generator=gen)
        lab = lab.repeat_interleave(blk, 2).repeat_interleave(blk, 3).float()
        inst = torch.randint(0, 8, (B, 1, H // blk, W // blk), generator=gen)
        inst = inst.repeat_interleave(blk, 2).repeat_interleave(blk, 3)
        img = torch.rand(B, 3, H, W, generator=gen) * 2 - 1
        model.set_input({'label': lab.clone(), 'instance': inst.clone(), 'image': img.clone(), 'path': ['x'] * B})
        seg = model.input_semantics.detach().clone()
        model.set_requires_grad(mm.netD, False)
        model.optimizer_G.zero_grad()
        model.backward_G()
        G_grads = grads_of(mm.netG)
        model.optimizer_G.step()
        model.set_requires_grad(mm.netD, True)
        model.optimizer_D.zero_grad()
        model.backward_D()
        D_grads = grads_of(mm.netD)
        model.optimizer_D.step()
        step = {'label': lab, 'instance': inst, 'image': img, 'seg': seg.to(torch.uint8),
                'losses': {k.split('/')[-1]: float(v) for k, v in model.get_current_losses().items()},
                'G_checksum_after': checksum(mm.netG), 'D_checksum_after': checksum(mm.netD)}
        if it == 0:
            step.update(G_grads=G_grads, D_grads=D_grads, G_buffers_after=snap(dict(mm.netG.named_buffers())),
                        D_buffers_after=snap(dict(mm.netD.named_buffers())))
        fix['steps'].append(step)
    return fix


CASES = {
    # scripts/pix2pix/cityscapes/train_inception_teacher.sh: BatchNorm with running statistics, hinge
    'train_pix2pix_bn_hinge': lambda n: make_pix2pix(n, 'batch', 'hinge', 'l1', 3, 32, 32),
    # smooth losses (lsgan + l2): pins the whole backward pass tightly
    'train_pix2pix_in_lsgan_l2': lambda n: make_pix2pix(n, 'instance', 'lsgan', 'l2', 2, 32, 48),
    # scripts/cycle_gan/horse2zebra/train_inception_teacher.sh: InstanceNorm, lsgan, identity 0.5; a pool of 3 images
    # so that the history branch (random replacement) is taken within the recorded steps
    # (data seed 235: with 233 one pre-activation of D_A's 3x3 layer sits 5e-6 from the LeakyReLU kink at step 0, which
    # makes d loss / d fake_B flip by 4.5 % under 1e-6 input differences -- useless as a parity vector)
    'train_cyclegan_in_lsgan': lambda n: make_cyclegan(n, 'instance', 'lsgan', 2, 32, 32, 3, 5, data_seed=235),
    # BatchNorm generators: running statistics move three times per generator per step, in the reference's call order
    'train_cyclegan_bn_lsgan': lambda n: make_cyclegan(n, 'batch', 'lsgan', 2, 32, 32, 0, 2),
    # scripts/gaugan/cityscapes/train_inception_teacher.sh
    'train_spade_more': lambda n: make_spade(n, 2, 128, 2.0, 6),
}


def main():
    os.makedirs(OUT_DIR, exist_ok=True)
    only = sys.argv[1:]
    for name, fn in CASES.items():
        if only and name not in only:
            continue
        fix = fn(name)
        path = os.path.join(OUT_DIR, name + '.pt')
        torch.save(fix, path)
        print(name, 'losses', fix['steps'][0]['losses'], '-> %.2f MB' % (os.path.getsize(path) / 1e6))


if __name__ == '__main__':
    main()
