"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's TEACHER-TRAINING steps (SURVEY.md section 8(f)
row 3: the step on the other side of the teacher checkpoint).

Plain PyTorch (CPU, fp32 or fp64), functional over reference-format ``state_dict``s, built from the network and
loss restatements of oracle/cat_oracle.py and oracle/spade_oracle.py.  Only ``tests/`` may import it, as the
checker -- never the product path (cat_b200/).

Pinning: against vectors written by the real reference models run in the build container
(oracle/make_golden_train.py -> tests/golden/train_*.pt, checked by tests/test_train_oracle_golden.py).

Every function cites the reference file:line (relative to /root/reference) that it restates.
"""
import random

import torch
import torch.nn.functional as F

from oracle import cat_oracle as O
from oracle import spade_oracle as SO


def _params(sd):
    return {k: v for k, v in sd.items() if O._is_param(k)}


def _arm(params):
    for p in params.values():
        p.requires_grad_(True)
        p.grad = None


def _disarm(params):
    for p in params.values():
        p.requires_grad_(False)
        p.grad = None


def _grads(params):
    return {k: p.grad.detach().clone() for k, p in params.items() if p.grad is not None}


# --------------------------------------------------------------------------------------------
# pix2pix
# --------------------------------------------------------------------------------------------
def pix2pix_train_step(state, real_A, real_B, hp, grad_hook=None):
    """One Pix2PixModel.optimize_parameters (models/pix2pix_model.py:203-212): forward (:153-155), backward_D
    (:157-172) with D on cat(real_A, .) pairs, optimizer_D.step, backward_G (:174-201: GAN * lambda_gan + recon *
    lambda_recon; the comp-cost term is off at its default weight 0), optimizer_G.step.

    state: 'G_sd','D_sd' (reference-format, updated in place), 'G_arch','D_arch', 'adam_G','adam_D'.
    hp: gan_mode, lambda_recon, lambda_gan, lr, beta1, recon_loss_type."""
    G_sd, D_sd, G_arch, D_arch = state['G_sd'], state['D_sd'], state['G_arch'], state['D_arch']
    out = {}
    G_params, D_params = _params(G_sd), _params(D_sd)
    _arm(G_params)
    fake_B = O.generator_forward(G_sd, G_arch, real_A, training=True)
    fake_B.retain_grad()
    out['fake_B'] = fake_B.detach().clone()
    # ---- backward_D
    _arm(D_params)
    fake_AB = torch.cat((real_A, fake_B), 1).detach()
    real_AB = torch.cat((real_A, real_B), 1).detach()
    loss_D_fake = O.gan_loss(hp['gan_mode'], O.discriminator_forward(D_sd, D_arch, fake_AB, True), False, True)
    loss_D_real = O.gan_loss(hp['gan_mode'], O.discriminator_forward(D_sd, D_arch, real_AB, True), True, True)
    ((loss_D_fake + loss_D_real) * 0.5).backward()
    out['loss_D_fake'], out['loss_D_real'] = loss_D_fake.detach(), loss_D_real.detach()
    out['D_grads'] = _grads(D_params)
    with torch.no_grad():
        if grad_hook is not None:
            out['D_grads'] = grad_hook('D', out['D_grads'])
        O.adam_update(D_params, out['D_grads'], state['adam_D'], hp['lr'], hp['beta1'])
    _disarm(D_params)
    # ---- backward_G
    pred_fake = O.discriminator_forward(D_sd, D_arch, torch.cat((real_A, fake_B), 1), True)
    loss_G_gan = O.gan_loss(hp['gan_mode'], pred_fake, True, False) * hp['lambda_gan']
    loss_G_recon = O.recon_loss(hp.get('recon_loss_type', 'l1'), fake_B, real_B) * hp['lambda_recon']
    (loss_G_gan + loss_G_recon).backward()
    out['loss_G_gan'], out['loss_G_recon'] = loss_G_gan.detach(), loss_G_recon.detach()
    out['fake_B_grad'] = fake_B.grad.detach().clone()
    out['G_grads'] = _grads(G_params)
    with torch.no_grad():
        if grad_hook is not None:
            out['G_grads'] = grad_hook('G', out['G_grads'])
        O.adam_update(G_params, out['G_grads'], state['adam_G'], hp['lr'], hp['beta1'])
    _disarm(G_params)
    return out


# --------------------------------------------------------------------------------------------
# CycleGAN
# --------------------------------------------------------------------------------------------
class ImagePool:
    """utils/image_pool.py:5-53: history buffer of generated images; decisions drawn from Python's global
    ``random`` exactly like the reference (uniform(0,1) > 0.5, then randint(0, pool_size-1))."""

    def __init__(self, pool_size):
        self.pool_size, self.images = pool_size, []

    def query(self, images):
        if self.pool_size == 0:
            return images
        ret = []
        for image in images:
            image = image.detach().unsqueeze(0)
            if len(self.images) < self.pool_size:
                self.images.append(image)
                ret.append(image)
            elif random.uniform(0, 1) > 0.5:
                i = random.randint(0, self.pool_size - 1)
                ret.append(self.images[i].clone())
                self.images[i] = image
            else:
                ret.append(image)
        return torch.cat(ret, 0)


def cyclegan_train_step(state, real_A, real_B, hp, grad_hook=None):
    """One CycleGANModel.optimize_parameters (models/cycle_gan_model.py:292-303): forward (:221-226), generators
    FIRST (backward_G :260-290: identity, GAN, cycle terms; Ds frozen), optimizer_G.step over both generators, then
    backward_D_A / backward_D_B (:228-258: real first, pooled fake second, (real + fake) * 0.5), optimizer_D.step.

    state: 'G_A_sd','G_B_sd','D_A_sd','D_B_sd', 'G_arch','D_arch', 'adam_G','adam_D' (keys prefixed 'A.' / 'B.'),
           'pool_A','pool_B' (ImagePool of fake_A / fake_B).
    hp: gan_mode, lambda_A, lambda_B, lambda_identity, lr, beta1."""
    GA, GB, DA, DB = state['G_A_sd'], state['G_B_sd'], state['D_A_sd'], state['D_B_sd']
    G_arch, D_arch = state['G_arch'], state['D_arch']
    mode, lA, lB, lI = hp['gan_mode'], hp['lambda_A'], hp['lambda_B'], hp['lambda_identity']
    out = {}
    GA_p, GB_p, DA_p, DB_p = _params(GA), _params(GB), _params(DA), _params(DB)
    _arm(GA_p)
    _arm(GB_p)
    gen = lambda sd, x: O.generator_forward(sd, G_arch, x, training=True)
    dis = lambda sd, x: O.discriminator_forward(sd, D_arch, x, True)
    fake_B = gen(GA, real_A)
    rec_A = gen(GB, fake_B)
    fake_A = gen(GB, real_B)
    rec_B = gen(GA, fake_A)
    for t in (fake_A, fake_B):
        t.retain_grad()
    # ---- backward_G
    zero = torch.zeros(())
    if lI > 0:
        idt_A = gen(GA, real_B)
        loss_idt_A = F.l1_loss(idt_A, real_B) * lB * lI
        idt_B = gen(GB, real_A)
        loss_idt_B = F.l1_loss(idt_B, real_A) * lA * lI
    else:
        loss_idt_A = loss_idt_B = zero
    loss_G_A = O.gan_loss(mode, dis(DA, fake_B), True, True)     # criterionGAN(pred, True): for_discriminator defaults to True
    loss_G_B = O.gan_loss(mode, dis(DB, fake_A), True, True)
    loss_cycle_A = F.l1_loss(rec_A, real_A) * lA
    loss_cycle_B = F.l1_loss(rec_B, real_B) * lB
    (loss_G_A + loss_G_B + loss_cycle_A + loss_cycle_B + loss_idt_A + loss_idt_B).backward()
    out.update(fake_A=fake_A.detach().clone(), fake_B=fake_B.detach().clone(), rec_A=rec_A.detach().clone(),
               rec_B=rec_B.detach().clone(), fake_A_grad=fake_A.grad.detach().clone(),
               fake_B_grad=fake_B.grad.detach().clone(),
               loss_G_A=loss_G_A.detach(), loss_G_B=loss_G_B.detach(), loss_G_cycle_A=loss_cycle_A.detach(),
               loss_G_cycle_B=loss_cycle_B.detach(), loss_G_idt_A=loss_idt_A.detach(), loss_G_idt_B=loss_idt_B.detach())
    out['G_A_grads'], out['G_B_grads'] = _grads(GA_p), _grads(GB_p)
    with torch.no_grad():
        if grad_hook is not None:
            out['G_A_grads'] = grad_hook('G_A', out['G_A_grads'])
            out['G_B_grads'] = grad_hook('G_B', out['G_B_grads'])
        both = {**{'A.' + k: v for k, v in GA_p.items()}, **{'B.' + k: v for k, v in GB_p.items()}}
        grads = {**{'A.' + k: v for k, v in out['G_A_grads'].items()}, **{'B.' + k: v for k, v in out['G_B_grads'].items()}}
        O.adam_update(both, grads, state['adam_G'], hp['lr'], hp['beta1'])
    _disarm(GA_p)
    _disarm(GB_p)
    # ---- backward_D_A, backward_D_B
    _arm(DA_p)
    _arm(DB_p)
    out['pooled_B'] = state['pool_B'].query(fake_B.detach())
    loss_real = O.gan_loss(mode, dis(DA, real_B), True, True)
    loss_fake = O.gan_loss(mode, dis(DA, out['pooled_B']), False, True)
    loss_D_A = (loss_real + loss_fake) * 0.5
    loss_D_A.backward()
    out['pooled_A'] = state['pool_A'].query(fake_A.detach())
    loss_real = O.gan_loss(mode, dis(DB, real_A), True, True)
    loss_fake = O.gan_loss(mode, dis(DB, out['pooled_A']), False, True)
    loss_D_B = (loss_real + loss_fake) * 0.5
    loss_D_B.backward()
    out['loss_D_A'], out['loss_D_B'] = loss_D_A.detach(), loss_D_B.detach()
    out['D_A_grads'], out['D_B_grads'] = _grads(DA_p), _grads(DB_p)
    with torch.no_grad():
        if grad_hook is not None:
            out['D_A_grads'] = grad_hook('D_A', out['D_A_grads'])
            out['D_B_grads'] = grad_hook('D_B', out['D_B_grads'])
        both = {**{'A.' + k: v for k, v in DA_p.items()}, **{'B.' + k: v for k, v in DB_p.items()}}
        grads = {**{'A.' + k: v for k, v in out['D_A_grads'].items()}, **{'B.' + k: v for k, v in out['D_B_grads'].items()}}
        O.adam_update(both, grads, state['adam_D'], hp['lr'], hp['beta1'])
    _disarm(DA_p)
    _disarm(DB_p)
    return out


# --------------------------------------------------------------------------------------------
# SPADE / GauGAN
# --------------------------------------------------------------------------------------------
def spade_train_step(state, seg, real_B, hp, grad_hook=None):
    """One SPADEModel.optimize_parameters (models/spade_model.py:207-215): backward_G (:189-196 -> compute_G_loss,
    models/modules/spade_modules/spade_model_modules.py:97-120: hinge GAN on the fake half + feature matching + VGG),
    optimizer_G.step, backward_D (-> compute_D_loss :122-139: a second, no-grad generator forward with the updated
    weights), optimizer_D.step.  TTUR Adam (create_optimizers :53-66).

    state: 'G_sd','D_sd','vgg_sd', 'G_arch','D_arch', 'adam_G','adam_D'.
    hp: lambda_gan, lambda_feat, lambda_vgg, lr_G, lr_D, beta1, beta2."""
    G_sd, D_sd, V_sd, G_arch, D_arch = state['G_sd'], state['D_sd'], state['vgg_sd'], state['G_arch'], state['D_arch']
    out = {}
    G_params = {k: v for k, v in G_sd.items() if SO._is_param(k)}
    _arm(G_params)
    fake = SO.spade_generator_forward(G_sd, G_arch, seg, training=True)
    fake.retain_grad()
    pred_fake, pred_real = SO._discriminate(D_sd, D_arch, seg, fake, real_B)
    loss_gan = SO.hinge_multiscale(pred_fake, True, False) * hp['lambda_gan']
    loss_feat = 0
    num_D = len(pred_fake)
    for i in range(num_D):
        for j in range(len(pred_fake[i]) - 1):
            loss_feat = loss_feat + F.l1_loss(pred_fake[i][j], pred_real[i][j].detach()) * hp['lambda_feat'] / num_D
    loss_vgg = SO.vgg_loss(V_sd, fake, real_B) * hp['lambda_vgg']
    (loss_gan + loss_feat + loss_vgg).backward()
    out.update(fake_B=fake.detach().clone(), fake_B_grad=fake.grad.detach().clone(), loss_G_gan=loss_gan.detach(),
               loss_G_feat=loss_feat.detach(), loss_G_vgg=loss_vgg.detach(), G_grads=_grads(G_params))
    with torch.no_grad():
        if grad_hook is not None:
            out['G_grads'] = grad_hook('G', out['G_grads'])
        SO.adam_update(G_params, out['G_grads'], state['adam_G'], hp['lr_G'], hp['beta1'], hp['beta2'])
    _disarm(G_params)
    D_params = {k: v for k, v in D_sd.items() if SO._is_param(k)}
    _arm(D_params)
    with torch.no_grad():
        fake = SO.spade_generator_forward(G_sd, G_arch, seg, training=True)
    pred_fake, pred_real = SO._discriminate(D_sd, D_arch, seg, fake, real_B)
    loss_D_fake = SO.hinge_multiscale(pred_fake, False, True)
    loss_D_real = SO.hinge_multiscale(pred_real, True, True)
    (loss_D_fake + loss_D_real).backward()
    out.update(loss_D_fake=loss_D_fake.detach(), loss_D_real=loss_D_real.detach(), D_grads=_grads(D_params))
    with torch.no_grad():
        if grad_hook is not None:
            out['D_grads'] = grad_hook('D', out['D_grads'])
        SO.adam_update(D_params, out['D_grads'], state['adam_D'], hp['lr_D'], hp['beta1'], hp['beta2'])
    _disarm(D_params)
    return out
