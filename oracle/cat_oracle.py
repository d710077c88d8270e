"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the CAT distillation hot path.

This is the parity oracle for cat_b200: a plain PyTorch (CPU, fp32 or fp64) restatement of the
reference algorithm, written functionally over reference-format ``state_dict``s.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` leg may import it,
and only as the checker or the timed CPU baseline -- never from the product path (cat_b200/).

Pinning: the reference ships no tests or golden vectors of its own, so the
oracle is pinned against outputs of the real reference run in the build container
(oracle/make_golden.py -> tests/golden/*.pt, checked by tests/test_oracle_golden.py).

Every function cites the reference file:line (relative to /root/reference) that it restates.
"""
import math

import torch
import torch.nn.functional as F

MAPPING_LAYERS = ['down_sampling.9', 'features.2', 'features.5', 'features.8']


# --------------------------------------------------------------------------------------------
# normalisation layers
# --------------------------------------------------------------------------------------------
def _norm(x, sd, prefix, arch, training):
    """nn.BatchNorm2d / nn.InstanceNorm2d as built by get_norm_layer (models/networks.py:29-64)
    with eps/momentum from distill_options.py:112-119.  BatchNorm in training mode normalises with
    biased batch statistics and updates the running buffers in ``sd`` in place (unbiased variance,
    momentum 0.1); in eval mode it uses the running buffers.  InstanceNorm without running stats
    always uses per-(n,c) statistics."""
    w = sd.get(prefix + '.weight') if arch['affine'] else None
    b = sd.get(prefix + '.bias') if arch['affine'] else None
    eps, mom = arch['eps'], arch['momentum']
    if arch['norm'] == 'batch':
        if arch['track_running_stats']:
            rm, rv = sd[prefix + '.running_mean'], sd[prefix + '.running_var']
            out = F.batch_norm(x, rm, rv, w, b, training, mom, eps)
            if training and (prefix + '.num_batches_tracked') in sd:
                sd[prefix + '.num_batches_tracked'] += 1
            return out
        return F.batch_norm(x, None, None, w, b, True, mom, eps)
    if arch['norm'] == 'instance':
        if arch['track_running_stats']:
            rm, rv = sd[prefix + '.running_mean'], sd[prefix + '.running_var']
            return F.instance_norm(x, rm, rv, w, b, training, mom, eps)
        return F.instance_norm(x, None, None, w, b, True, mom, eps)
    raise NotImplementedError(arch['norm'])


_EMULATE_BF16 = [False]


class emulate_bf16:
    """Context manager: round weights of the dense convs and every activation that cat_b200 stores in
    HBM to bf16 (straight-through in backward), at the same points as the CUDA engine.  Used by the GPU
    parity tests to separate kernel correctness from the conditioning of ReLU/hinge/L1 gradients, whose
    masks and signs flip under ~1% forward rounding differences.  Off by default: the plain oracle is the
    fp32 reference algorithm."""

    def __enter__(self):
        self.prev = _EMULATE_BF16[0]
        _EMULATE_BF16[0] = True

    def __exit__(self, *a):
        _EMULATE_BF16[0] = self.prev


def qa(x):
    if not _EMULATE_BF16[0]:
        return x
    return x + (x.to(torch.bfloat16).to(x.dtype) - x).detach()


qw = qa


def _reflect(x, p):
    return F.pad(x, (p, p, p, p), mode='reflect') if p > 0 else x


# --------------------------------------------------------------------------------------------
# generator
# --------------------------------------------------------------------------------------------
def block_forward(x, sd, prefix, blk, arch, training):
    """InvertedResidualChannels.forward (models/modules/inception_modules.py:230-236) with the
    branches built by _build (:124-180): res branch = pad, Conv k, norm, ReLU, Dropout(0), pad,
    Conv k; dw branch = Conv 1x1, norm, ReLU, pad, depthwise Conv k, norm, ReLU, Dropout(0),
    Conv 1x1.  Branches whose width is 0 are skipped and do not consume a ModuleList index
    (:132-133,153-154)."""
    ks = arch['kernel_sizes']
    outs = []
    j = 0
    for mid, k in zip(blk['res'], ks):
        if mid == 0:
            continue
        p = f'{prefix}.res_ops.{j}'
        h = qa(F.conv2d(_reflect(x, (k - 1) // 2), qw(sd[p + '.1.0.weight']), sd.get(p + '.1.0.bias')))
        h = qa(F.relu(_norm(h, sd, p + '.1.1', arch, training)))
        h = F.conv2d(_reflect(h, (k - 1) // 2), qw(sd[p + '.4.weight']), sd.get(p + '.4.bias'))
        outs.append(h)
        j += 1
    j = 0
    for mid, k in zip(blk['dw'], ks):
        if mid == 0:
            continue
        p = f'{prefix}.dw_ops.{j}'
        h = qa(F.conv2d(x, qw(sd[p + '.0.0.weight']), sd.get(p + '.0.0.bias')))
        h = qa(F.relu(_norm(h, sd, p + '.0.1', arch, training)))
        h = qa(F.conv2d(_reflect(h, (k - 1) // 2), sd[p + '.2.0.weight'], sd.get(p + '.2.0.bias'),
                        groups=mid))
        h = qa(F.relu(_norm(h, sd, p + '.2.1', arch, training)))
        h = F.conv2d(h, qw(sd[p + '.4.weight']), sd.get(p + '.4.bias'))
        outs.append(h)
        j += 1
    if not outs:
        return x
    tmp = outs[0]
    for o in outs[1:]:
        tmp = tmp + o
    tmp = _norm(qa(tmp), sd, prefix + '.pw_bn', arch, training)
    return qa(x + tmp)


def generator_forward(sd, arch, x, training=False, capture=None):
    """InceptionGenerator.forward (models/modules/inception_architecture/inception_generator.py:
    137-142; layers built at :37-135).  ``capture`` (dict) receives the activations of the four
    distillation mapping layers (base_inception_distiller.py:183-190): 'down_sampling.9' is the
    last in-place ReLU of the down-sampling stack, 'features.{2,5,8}' are block outputs."""
    h = qa(F.conv2d(_reflect(qa(x), 3), qw(sd['down_sampling.1.weight']), sd.get('down_sampling.1.bias')))
    h = qa(F.relu(_norm(h, sd, 'down_sampling.2', arch, training)))
    for ci, ni in ((4, 5), (7, 8)):
        h = qa(F.conv2d(h, qw(sd[f'down_sampling.{ci}.weight']), sd.get(f'down_sampling.{ci}.bias'),
                        stride=2, padding=1))
        h = qa(F.relu(_norm(h, sd, f'down_sampling.{ni}', arch, training)))
    if capture is not None:
        capture['down_sampling.9'] = h
    for i, blk in enumerate(arch['blocks']):
        h = block_forward(h, sd, f'features.{i}', blk, arch, training)
        if capture is not None and f'features.{i}' in MAPPING_LAYERS:
            capture[f'features.{i}'] = h
    for ci, ni in ((0, 1), (3, 4)):
        h = qa(F.conv_transpose2d(h, qw(sd[f'up_sampling.{ci}.weight']), sd.get(f'up_sampling.{ci}.bias'),
                                  stride=2, padding=1, output_padding=1))
        h = qa(F.relu(_norm(h, sd, f'up_sampling.{ni}', arch, training)))
    h = F.conv2d(_reflect(h, 3), qw(sd['up_sampling.7.weight']), sd.get('up_sampling.7.bias'))
    return qa(torch.tanh(h))


# --------------------------------------------------------------------------------------------
# discriminator
# --------------------------------------------------------------------------------------------
def discriminator_layers(arch):
    """Layer list of NLayerDiscriminator (models/modules/discriminators.py:37-75) as tuples
    (seq_index_of_conv, cin, cout, stride, has_norm, has_act)."""
    ndf, n_layers = arch['ndf'], arch['n_layers']
    layers = [(0, arch['input_nc'], ndf, 2, False, True)]
    idx, mult = 2, 1
    for n in range(1, n_layers):
        prev, mult = mult, min(2 ** n, 8)
        layers.append((idx, ndf * prev, ndf * mult, 2, True, True))
        idx += 3
    prev, mult = mult, min(2 ** n_layers, 8)
    layers.append((idx, ndf * prev, ndf * mult, 1, True, True))
    idx += 3
    layers.append((idx, ndf * mult, 1, 1, False, False))
    return layers


def discriminator_forward(sd, arch, x, training=True):
    """NLayerDiscriminator.forward (models/modules/discriminators.py:77-79): 4x4 convs, padding 1,
    norm on the middle layers, LeakyReLU(0.2) (active_fn(0.2), :41,56,69)."""
    h = qa(x)
    for (ci, cin, cout, stride, has_norm, has_act) in discriminator_layers(arch):
        h = F.conv2d(h, qw(sd[f'model.{ci}.weight']), sd.get(f'model.{ci}.bias'), stride=stride,
                     padding=1)
        if has_norm:
            h = _norm(qa(h), sd, f'model.{ci + 1}', arch, training)
        if has_act:
            h = qa(F.leaky_relu(h, 0.2))
    return h


# --------------------------------------------------------------------------------------------
# losses
# --------------------------------------------------------------------------------------------
def gan_loss(mode, pred, target_is_real, for_discriminator=True):
    """GANLoss.__call__ for a single tensor prediction (models/modules/loss.py:52-99)."""
    if mode == 'lsgan':
        t = torch.ones_like(pred) if target_is_real else torch.zeros_like(pred)
        return F.mse_loss(pred, t)
    if mode == 'vanilla':
        t = torch.ones_like(pred) if target_is_real else torch.zeros_like(pred)
        return F.binary_cross_entropy_with_logits(pred, t)
    if mode == 'hinge':
        if for_discriminator:
            if target_is_real:
                return -torch.mean(torch.min(pred - 1, torch.zeros_like(pred)))
            return -torch.mean(torch.min(-pred - 1, torch.zeros_like(pred)))
        assert target_is_real
        return -torch.mean(pred)
    raise NotImplementedError(mode)


def recon_loss(kind, a, b):
    """criterionRecon (distillers/base_inception_distiller.py:171-176): L1Loss | MSELoss | SmoothL1Loss."""
    return {'l1': F.l1_loss, 'l2': F.mse_loss, 'smooth_l1': F.smooth_l1_loss}[kind](a, b)


def ka(X, Y):
    """KA(X, Y) (utils/common.py:38-46): <XX^T, YY^T>_F / (||XX^T||_F ||YY^T||_F).  The
    denominator is evaluated as a product of two square roots (identical in exact arithmetic;
    avoids the fp32 overflow of ((Kx**2).sum()*(Ky**2).sum()) noted in SURVEY.md section 7)."""
    X_ = X.reshape(X.size(0), -1)
    Y_ = Y.reshape(Y.size(0), -1)
    Kx = X_ @ X_.T
    Ky = Y_ @ Y_.T
    return (Kx * Ky).sum() / ((Kx ** 2).sum().sqrt() * (Ky ** 2).sum().sqrt())


def ka_grad_x(X, Y):
    """Analytic dKA/dX: G = Ky/(|Kx||Ky|) - <Kx,Ky> Kx/(|Kx|^3 |Ky|),
    dX = 2 G X.  No gradient flows to Y (teacher activations, inception_distiller.py:102-103)."""
    X_ = X.reshape(X.size(0), -1)
    Y_ = Y.reshape(Y.size(0), -1)
    Kx = X_ @ X_.T
    Ky = Y_ @ Y_.T
    nx, ny = (Kx ** 2).sum().sqrt(), (Ky ** 2).sum().sqrt()
    num = (Kx * Ky).sum()
    G = Ky / (nx * ny) - num * Kx / (nx ** 3 * ny)
    return (2.0 * G @ X_).reshape(X.shape)


# --------------------------------------------------------------------------------------------
# optimiser
# --------------------------------------------------------------------------------------------
def adam_update(params, grads, state, lr, beta1, beta2=0.999, eps=1e-8):
    """torch.optim.Adam as configured at base_inception_distiller.py:205-214 (no weight decay,
    no amsgrad).  Parameters whose grad is None are skipped (the netA adaptors under the 'ka'
    loss, inception_distiller.py:135-152).  ``state`` maps name -> dict(step, m, v)."""
    for name, p in params.items():
        g = grads.get(name)
        if g is None:
            continue
        st = state.setdefault(name, {'step': 0, 'm': torch.zeros_like(p), 'v': torch.zeros_like(p)})
        st['step'] += 1
        st['m'].mul_(beta1).add_(g, alpha=1 - beta1)
        st['v'].mul_(beta2).addcmul_(g, g, value=1 - beta2)
        bc1 = 1 - beta1 ** st['step']
        bc2 = 1 - beta2 ** st['step']
        denom = (st['v'].sqrt() / math.sqrt(bc2)).add_(eps)
        p.data.addcdiv_(st['m'], denom, value=-lr / bc1)


# --------------------------------------------------------------------------------------------
# the distillation step
# --------------------------------------------------------------------------------------------
def _is_param(key):
    return key.endswith('.weight') or key.endswith('.bias')


def distill_step(state, real_A, real_B, hp, grad_hook=None):
    """One InceptionDistiller.optimize_parameters (distillers/inception_distiller.py:179-188):
    forward (:100-104), backward_D (base_inception_distiller.py:293-312), optimizer_D.step,
    backward_G (inception_distiller.py:159-177), optimizer_G.step.

    state: dict with 'teacher_sd','student_sd','D_sd' (reference-format tensors, updated in place),
           'teacher_arch','student_arch','D_arch', 'adam_G','adam_D' (optimiser state dicts).
    hp:    dict(gan_mode, aligned, lambda_recon, lambda_gan, lambda_distill, lr, beta1,
                student_training, ka_scale) -- ka_scale is the DataParallel replica-sum factor of
           inception_distiller.py:145-148 (1 on a single device).
           hp['distill_loss_type'] == 'mse' (--distill_G_loss_type mse, inception_distiller.py:111-133): the four terms
           are F.mse_loss(netA_i(Sact_i), Tact_i) through the 1x1 adaptor convs state['netA_sds'][i] ('weight', 'bias';
           base_inception_distiller.py:195-202), which are parameters of optimizer_G (:204-210).
    Returns dict of losses, outputs, activations and gradients."""
    T_sd, S_sd, D_sd = state['teacher_sd'], state['student_sd'], state['D_sd']
    T_arch, S_arch, D_arch = state['teacher_arch'], state['student_arch'], state['D_arch']
    out = {}
    # ---- forward (inception_distiller.py:100-104); teacher is in eval mode (base_..:168)
    Tacts, Sacts = {}, {}
    with torch.no_grad():
        Tfake = generator_forward(T_sd, T_arch, real_A, training=False, capture=Tacts)
    S_params = {k: v for k, v in S_sd.items() if _is_param(k)}
    for p in S_params.values():
        p.requires_grad_(True)
        p.grad = None
    Sfake = generator_forward(S_sd, S_arch, real_A, training=hp.get('student_training', True),
                              capture=Sacts)
    for a in Sacts.values():
        a.retain_grad()
    Sfake.retain_grad()
    out['Tfake_B'], out['Sfake_B'] = Tfake, Sfake.detach().clone()
    out['Tacts'] = {k: v for k, v in Tacts.items()}
    out['Sacts'] = {k: v.detach().clone() for k, v in Sacts.items()}

    # ---- backward_D (base_inception_distiller.py:293-312)
    D_params = {k: v for k, v in D_sd.items() if _is_param(k)}
    for p in D_params.values():
        p.requires_grad_(True)
        p.grad = None
    if hp['aligned']:
        fake = torch.cat((real_A, Sfake), 1).detach()
        real = torch.cat((real_A, real_B), 1).detach()
    else:
        fake, real = Sfake.detach(), real_B.detach()
    pred_fake = discriminator_forward(D_sd, D_arch, fake, training=True)
    pred_fake.retain_grad()
    loss_D_fake = gan_loss(hp['gan_mode'], pred_fake, False, True)
    pred_real = discriminator_forward(D_sd, D_arch, real, training=True)
    pred_real.retain_grad()
    loss_D_real = gan_loss(hp['gan_mode'], pred_real, True, True)
    loss_D = (loss_D_fake + loss_D_real) * 0.5
    loss_D.backward()
    out['loss_D_fake'], out['loss_D_real'] = loss_D_fake.detach(), loss_D_real.detach()
    out['pred_fake_D'] = pred_fake.detach()
    out['dpred_fake'], out['dpred_real'] = pred_fake.grad.detach().clone(), pred_real.grad.detach().clone()
    out['D_grads'] = {k: p.grad.detach().clone() for k, p in D_params.items()}
    with torch.no_grad():
        if grad_hook is not None:   # data-parallel tests: reduce the gradients across ranks before the update
            out['D_grads'] = grad_hook('D', out['D_grads'])
        adam_update(D_params, out['D_grads'], state['adam_D'], hp['lr'], hp['beta1'])

    # ---- backward_G (inception_distiller.py:159-177), D frozen (:185)
    for p in D_params.values():
        p.requires_grad_(False)
        p.grad = None
    if hp['aligned']:
        loss_G_recon = recon_loss(hp.get('recon_loss_type', 'l1'), Sfake, real_B) * hp['lambda_recon']
        fake = torch.cat((real_A, Sfake), 1)
    else:
        loss_G_recon = recon_loss(hp.get('recon_loss_type', 'l1'), Sfake, Tfake) * hp['lambda_recon']
        fake = Sfake
    pred_fake = discriminator_forward(D_sd, D_arch, fake, training=True)
    loss_G_gan = gan_loss(hp['gan_mode'], pred_fake, True, False) * hp['lambda_gan']
    distill_terms = []
    mse = hp.get('distill_loss_type', 'ka') == 'mse'
    A_params = {}
    if mse:
        A_params = {f'A{i}.{k}': v for i, sd in enumerate(state['netA_sds']) for k, v in sd.items()}
        for p in A_params.values():
            p.requires_grad_(True)
            p.grad = None
    for i, n in enumerate(MAPPING_LAYERS):
        if mse:
            mapped = qa(F.conv2d(Sacts[n], qw(A_params[f'A{i}.weight']), A_params[f'A{i}.bias']))
            distill_terms.append(F.mse_loss(mapped, Tacts[n]) * hp.get('ka_scale', 1.0))
        else:
            distill_terms.append(-ka(Sacts[n], Tacts[n]) * hp.get('ka_scale', 1.0))
    loss_G_distill = sum(distill_terms) * hp['lambda_distill']
    loss_G = loss_G_gan + loss_G_recon + loss_G_distill
    loss_G.backward()
    out['loss_G_recon'], out['loss_G_gan'] = loss_G_recon.detach(), loss_G_gan.detach()
    out['loss_G_distill'] = loss_G_distill.detach()
    out['loss_G_distill_terms'] = [t.detach() for t in distill_terms]
    out['Sact_grads'] = {k: v.grad.detach().clone() for k, v in Sacts.items()}
    out['Sfake_grad'] = Sfake.grad.detach().clone()
    out['S_grads'] = {k: p.grad.detach().clone() for k, p in S_params.items()}
    out['A_grads'] = {k: p.grad.detach().clone() for k, p in A_params.items()}
    with torch.no_grad():
        if grad_hook is not None:
            out['S_grads'] = grad_hook('S', out['S_grads'])
            if mse:
                out['A_grads'] = grad_hook('A', out['A_grads'])
        adam_update(S_params, out['S_grads'], state['adam_G'], hp['lr'], hp['beta1'])
        adam_update(A_params, out['A_grads'], state['adam_G'], hp['lr'], hp['beta1'])     # same optimizer_G, second group
    for p in list(S_params.values()) + list(A_params.values()):
        p.requires_grad_(False)
        p.grad = None
    return out


def clone_sd(sd, dtype=None):
    out = {}
    for k, v in sd.items():
        v = v.detach().clone()
        if dtype is not None and v.is_floating_point():
            v = v.to(dtype)
        out[k] = v
    return out
