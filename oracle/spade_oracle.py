"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the CAT *SPADE* distillation step.

Plain PyTorch (CPU) over reference-format ``state_dict``s, same conventions as oracle/cat_oracle.py: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import it, as the checker or the
timed CPU baseline.  Pinned against the real reference run in the build container
(oracle/make_golden_spade.py -> tests/golden/spade_*.pt, checked by tests/test_oracle_golden.py).
VGG19 weights are random (seeded) in the fixtures: the pretrained checkpoint is not available offline, so
parity of the VGG loss is against the same random weights ("pretrained parity unpinned").

Every function cites the reference file:line (relative to /root/reference) that it restates.
"""
import torch
import torch.nn.functional as F

from oracle.cat_oracle import adam_update, ka, qa, qw

MAPPING_LAYERS = ['head_0', 'G_middle_1', 'up_1']   # base_spade_distiller_modules.py:70
VGG_CFG = [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 256, 'M', 512, 512, 512, 512, 'M', 512]  # features[0:30]
VGG_CONV_IDX = [0, 2, 5, 7, 10, 12, 14, 16, 19, 21, 23, 25, 28]
VGG_TAPS = [1, 6, 11, 20, 29]        # relu1_1, relu2_1, relu3_1, relu4_1, relu5_1 (loss.py:160-173)
VGG_WEIGHTS = [1.0 / 32, 1.0 / 16, 1.0 / 8, 1.0 / 4, 1.0]


# --------------------------------------------------------------------------------------------
# normalisation
# --------------------------------------------------------------------------------------------
def _bn(x, sd, prefix, affine, training, eps=1e-5, mom=0.1):
    """SynchronizedBatchNorm2d outside DataParallel == F.batch_norm (sync_batchnorm/batchnorm.py:68-72):
    biased batch statistics in training (running buffers updated in place, unbiased variance), running
    buffers in eval."""
    w = sd[prefix + '.weight'] if affine else None
    b = sd[prefix + '.bias'] if affine else None
    # (the module calls F.batch_norm directly, so num_batches_tracked never advances)
    return F.batch_norm(x, sd[prefix + '.running_mean'], sd[prefix + '.running_var'], w, b, training, mom, eps)


def _conv(x, sd, prefix, k, groups=1):
    w = sd[prefix + '.weight']
    return F.conv2d(x, qw(w) if groups == 1 else w, sd.get(prefix + '.bias'), padding=(k - 1) // 2, groups=groups)


# --------------------------------------------------------------------------------------------
# SPADE generator
# --------------------------------------------------------------------------------------------
# Activation of SPADEInvertedResidualChannels (block activation + its six main branches) for the pass being evaluated, set
# by spade_generator_forward from arch['active_fn'] (inception_modules.py:357,401-402).  The InceptionSPADE modulation
# body always uses nn.ReLU (inception_modules.py:600).
_ACTIVE = [F.relu]


def _active(x):
    return _ACTIVE[0](x)


def _branches(x, sd, prefix, res, dw, ks, training, dw_affine, last_conv_key, act=F.relu):
    """The six-branch body shared by SPADEInvertedResidualChannels._build (inception_modules.py:412-470) and
    InceptionSPADE._build (:672-722): res = ConvSyncBNReLU(k) -> Conv(k); dw = ConvSyncBNReLU(1) ->
    ConvSyncBNReLU(k, depthwise) -> Conv(1).  Zero padding (k-1)//2, every conv has a bias, branches of width 0
    are skipped without consuming a ModuleList index."""
    outs = []
    j = 0
    for mid, k in zip(res, ks):
        if mid == 0:
            continue
        p = f'{prefix}.res_ops.{j}'
        h = qa(act(_bn(qa(_conv(x, sd, p + '.0.conv', k)), sd, p + '.0.norm', True, training)))
        outs.append(_conv(h, sd, p + '.1' + last_conv_key, k))
        j += 1
    j = 0
    for mid, k in zip(dw, ks):
        if mid == 0:
            continue
        p = f'{prefix}.dw_ops.{j}'
        h = qa(act(_bn(qa(_conv(x, sd, p + '.0.conv', 1)), sd, p + '.0.norm', True, training)))
        h = qa(act(_bn(qa(_conv(h, sd, p + '.1.conv', k, groups=mid)), sd, p + '.1.norm', dw_affine, training)))
        outs.append(_conv(h, sd, p + '.2' + last_conv_key, 1))
        j += 1
    if not outs:
        return None
    tmp = outs[0]
    for o in outs[1:]:
        tmp = tmp + o
    return tmp


def spade_norm(x, seg, sd, prefix, blk, ks, training):
    """InceptionSPADE.forward (inception_modules.py:746-762): param-free BN of x, gamma/beta from the
    nearest-resized segmentation map through six branches, out = normalized * (1 + gamma) + beta."""
    normalized = _bn(x, sd, prefix + '.param_free_norm', False, training)
    segmap = F.interpolate(seg, size=x.shape[2:], mode='nearest')
    gb = _branches(segmap, sd, prefix, blk['spade_res'], blk['spade_dw'], ks, training, True, '')
    if gb is None:
        return normalized
    gb = qa(gb)
    C = x.shape[1]
    return normalized * (1 + gb[:, :C]) + gb[:, C:]


def spade_block(x, seg, sd, prefix, blk, ks, training):
    """SPADEInvertedResidualChannels.forward (inception_modules.py:549-562)."""
    def shortcut(v):
        if not blk['learned_shortcut']:
            return v
        h = qa(_bn(v, sd, prefix + '.shortcut.0', True, training))
        return F.conv2d(h, qw(sd[prefix + '.shortcut.1.conv.weight']))
    if not any(blk['res']) and not any(blk['dw']):
        return qa(shortcut(x))
    tmp = qa(_active(spade_norm(x, seg, sd, prefix + '.spade', blk, ks, training)))
    tmp = _branches(tmp, sd, prefix, blk['res'], blk['dw'], ks, training, False, '.conv', _active)
    return qa(tmp + shortcut(x))


def spade_generator_forward(sd, arch, seg, training=False, capture=None):
    """InceptionSPADEGenerator.forward (inception_spade_generator.py:63-124).  arch['active_fn'] is the generator's
    activation (get_active_fn, inception_modules.py:12-19): nn.ReLU on the distillation path (distill_options.py:123),
    nn.LeakyReLU() -- default slope 0.01 -- when SPADEModel trains the teacher (models/spade_model.py:92)."""
    ks = arch['kernel_sizes']
    _ACTIVE[0] = {'nn.ReLU': F.relu, 'nn.LeakyReLU': lambda t: F.leaky_relu(t, 0.01)}[arch.get('active_fn', 'nn.ReLU')]
    seg = qa(seg)
    x = F.interpolate(seg, size=(arch['sh'], arch['sw']))
    x = qa(F.conv2d(x, qw(sd['fc.weight']), sd['fc.bias'], padding=1))
    x = qa(_bn(x, sd, 'fc_norm', True, training))
    up = lambda t: F.interpolate(t, scale_factor=2, mode='nearest')
    more = arch['num_upsampling_layers'] in ('more', 'most')
    for name in arch['block_names']:
        if name in ('G_middle_0', 'up_0', 'up_1', 'up_2', 'up_3', 'up_4') or (name == 'G_middle_1' and more):
            x = up(x)
        x = spade_block(x, seg, sd, name, arch['blocks'][name], ks, training)
        if capture is not None and name in MAPPING_LAYERS:
            capture[name] = x
    x = F.conv2d(qa(F.leaky_relu(x, 0.2)), qw(sd['conv_img.weight']), sd['conv_img.bias'], padding=1)
    return qa(torch.tanh(x))


# --------------------------------------------------------------------------------------------
# multi-scale discriminator with spectral norm
# --------------------------------------------------------------------------------------------
def spectral_weight(sd, prefix, training):
    """torch.nn.utils.spectral_norm (legacy hook, as applied by get_nonspade_norm_layer,
    spade_architecture/normalization.py:17-50): one power iteration per training forward on the buffers
    weight_u / weight_v (in place, no grad), sigma = u^T W v, W_sn = weight_orig / sigma."""
    w = sd[prefix + '.weight_orig']
    u, v = sd[prefix + '.weight_u'], sd[prefix + '.weight_v']
    wm = w.reshape(w.shape[0], -1)
    if training:
        with torch.no_grad():
            v.copy_(F.normalize(torch.mv(wm.t(), u), dim=0, eps=1e-12))
            u.copy_(F.normalize(torch.mv(wm, v), dim=0, eps=1e-12))
    sigma = torch.dot(u.clone(), torch.mv(wm, v.clone()))
    return w / sigma


def multiscale_D_forward(sd, arch, x, training=True):
    """MultiscaleDiscriminator.forward (discriminators.py:212-226) over SPADENLayerDiscriminator (:129-180):
    4x4 convs with padding 2; middle layers spectral-norm conv (no bias) + InstanceNorm2d(affine=False) +
    LeakyReLU(0.2); every sub-discriminator returns all of its intermediate outputs; the input of the next
    scale is avg_pool2d(3, stride 2, padding 1, count_include_pad=False)."""
    results = []
    n_layers = arch['n_layers']
    x = qa(x)
    for d in range(arch['num_D']):
        pre = f'discriminator_{d}'
        outs = []
        h = F.conv2d(x, qw(sd[f'{pre}.model0.0.weight']), sd[f'{pre}.model0.0.bias'], stride=2, padding=2)
        h = qa(F.leaky_relu(h, 0.2))
        outs.append(h)
        for n in range(1, n_layers):
            stride = 1 if n == n_layers - 1 else 2
            w = spectral_weight(sd, f'{pre}.model{n}.0.0', training)
            h = qa(F.conv2d(h, qw(w), None, stride=stride, padding=2))
            h = qa(F.leaky_relu(F.instance_norm(h, eps=1e-5), 0.2))
            outs.append(h)
        h = F.conv2d(h, qw(sd[f'{pre}.model{n_layers}.0.weight']), sd[f'{pre}.model{n_layers}.0.bias'], stride=1, padding=2)
        outs.append(h)
        results.append(outs)
        if d + 1 < arch['num_D']:
            x = qa(F.avg_pool2d(x, kernel_size=3, stride=2, padding=[1, 1], count_include_pad=False))
    return results


def hinge_multiscale(preds, target_is_real, for_discriminator):
    """GANLoss.__call__ in hinge mode on a list of lists (loss.py:71-82): the last output of every scale, mean
    over scales."""
    loss = 0
    for p in preds:
        last = p[-1]
        if for_discriminator:
            mv = torch.min((last - 1) if target_is_real else (-last - 1), torch.zeros_like(last))
            loss = loss - torch.mean(mv)
        else:
            loss = loss - torch.mean(last)
    return loss / len(preds)


# --------------------------------------------------------------------------------------------
# VGG perceptual loss
# --------------------------------------------------------------------------------------------
def make_vgg_sd(seed):
    """The fixtures' VGG19: torchvision vgg19(weights=None) initialised under `seed` (the pretrained checkpoint
    of loss.py:154-155 is not available offline).  Keys '<features index>.weight' / '.bias'."""
    import torchvision
    st = torch.random.get_rng_state()
    torch.manual_seed(seed)
    net = torchvision.models.vgg19(weights=None)
    torch.random.set_rng_state(st)
    sd = net.features.state_dict()
    return {k: v.detach().clone() for k, v in sd.items() if int(k.split('.')[0]) <= 28}


def vgg_features(sd, x):
    """VGG19 slices 1-5 (loss.py:151-184): torchvision vgg19.features[0:30], outputs after relu{1..5}_1."""
    outs = []
    idx = 0
    h = qa(x)
    for c in VGG_CFG:
        if c == 'M':
            h = F.max_pool2d(h, 2, 2)
            idx += 1
        else:
            h = qa(F.relu(F.conv2d(h, qw(sd[f'{idx}.weight']), sd[f'{idx}.bias'], padding=1)))
            idx += 1   # conv
            if idx in VGG_TAPS:
                outs.append(h)
            idx += 1   # relu
    return outs


def vgg_loss(sd, x, y):
    """VGGLoss.forward (loss.py:187-203)."""
    fx = vgg_features(sd, x)
    with torch.no_grad():
        fy = vgg_features(sd, y)
    loss = 0
    for w, a, b in zip(VGG_WEIGHTS, fx, fy):
        loss = loss + w * F.l1_loss(a, b.detach())
    return loss


# --------------------------------------------------------------------------------------------
# preprocessing
# --------------------------------------------------------------------------------------------
def preprocess_input(label, instance, n_label):
    """SPADEModel.preprocess_input / get_edges (models/spade_model.py:142-179): one-hot of the label map
    plus the 4-neighbour instance-boundary map as the last channel."""
    label = label.long()
    bs, _, h, w = label.shape
    onehot = torch.zeros(bs, n_label, h, w).scatter_(1, label, 1.0)
    t = instance
    edge = torch.zeros(t.shape, dtype=torch.bool)
    dx = t[:, :, :, 1:] != t[:, :, :, :-1]
    dy = t[:, :, 1:, :] != t[:, :, :-1, :]
    edge[:, :, :, 1:] |= dx
    edge[:, :, :, :-1] |= dx
    edge[:, :, 1:, :] |= dy
    edge[:, :, :-1, :] |= dy
    return torch.cat((onehot, edge.float()), 1)


# --------------------------------------------------------------------------------------------
# the step
# --------------------------------------------------------------------------------------------
def _is_param(key):
    return key.endswith('.weight') or key.endswith('.bias') or key.endswith('.weight_orig')


def _discriminate(D_sd, D_arch, seg, fake, real):
    """SPADEModelModules.discriminate / divide_pred (spade_model_modules.py:136-156): ONE pass over the
    batch-concatenation [fake; real]."""
    both = torch.cat((torch.cat((seg, fake), 1), torch.cat((seg, real), 1)), 0)
    out = multiscale_D_forward(D_sd, D_arch, both, training=True)
    B = seg.shape[0]
    return [[t[:B] for t in p] for p in out], [[t[B:] for t in p] for p in out]


def spade_distill_step(state, seg, real_B, hp, grad_hook=None):
    """One BaseSPADEDistiller.optimize_parameters (distillers/base_spade_distiller.py:226-234):
    backward_G (models/spade_model.py:189-196 -> compute_G_loss, base_spade_distiller_modules.py:128-158),
    optimizer_G.step, backward_D (-> compute_D_loss, :160-175: a second, no-grad student forward with the
    updated weights), optimizer_D.step.

    state: 'teacher_sd','student_sd','D_sd','vgg_sd' + 'teacher_arch','student_arch','D_arch' + 'adam_G','adam_D'.
    hp: lambda_gan, lambda_feat, lambda_vgg, lambda_distill, lr_G, lr_D, beta1, beta2, ka_scale.
    hp['distill_loss_type'] == 'mse' (spade_distiller_modules.py:23-25): the terms are F.mse_loss(netA_i(Sact_i), Tact_i)
    through the 1x1 adaptor convs state['netA_sds'][i], parameters of optimizer_G (base_spade_distiller_modules.py:91-105)."""
    T_sd, S_sd, D_sd, V_sd = state['teacher_sd'], state['student_sd'], state['D_sd'], state['vgg_sd']
    T_arch, S_arch, D_arch = state['teacher_arch'], state['student_arch'], state['D_arch']
    out = {}
    # ---- G phase
    Tacts, Sacts = {}, {}
    with torch.no_grad():
        Tfake = spade_generator_forward(T_sd, T_arch, seg, training=False, capture=Tacts)
    S_params = {k: v for k, v in S_sd.items() if _is_param(k)}
    for p in S_params.values():
        p.requires_grad_(True)
        p.grad = None
    s_train = hp.get('student_training', True)     # False: the reference's first step of a run (student still in eval())
    Sfake = spade_generator_forward(S_sd, S_arch, seg, training=s_train, capture=Sacts)
    for a in Sacts.values():
        a.retain_grad()
    Sfake.retain_grad()
    mse = hp.get('distill_loss_type', 'ka') == 'mse'
    A_params = {}
    if mse:
        A_params = {f'A{i}.{k}': v for i, sd in enumerate(state['netA_sds']) for k, v in sd.items()}
        for p in A_params.values():
            p.requires_grad_(True)
            p.grad = None
        terms = [F.mse_loss(qa(F.conv2d(Sacts[n], qw(A_params[f'A{i}.weight']), A_params[f'A{i}.bias'])), Tacts[n])
                 * hp.get('ka_scale', 1.0) for i, n in enumerate(MAPPING_LAYERS)]
    else:
        terms = [-ka(Sacts[n], Tacts[n]) * hp.get('ka_scale', 1.0) for n in MAPPING_LAYERS]
    loss_distill = sum(terms) * hp['lambda_distill']
    pred_fake, pred_real = _discriminate(D_sd, D_arch, seg, Sfake, real_B)
    loss_gan = hinge_multiscale(pred_fake, True, False) * hp['lambda_gan']
    loss_feat = 0
    num_D = len(pred_fake)
    for i in range(num_D):
        for j in range(len(pred_fake[i]) - 1):
            loss_feat = loss_feat + F.l1_loss(pred_fake[i][j], pred_real[i][j].detach()) * hp['lambda_feat'] / num_D
    loss_vgg = vgg_loss(V_sd, Sfake, real_B) * hp['lambda_vgg']
    loss_G = loss_gan + loss_distill + loss_feat + loss_vgg
    loss_G.backward()
    out.update(Tfake_B=Tfake, Sfake_B=Sfake.detach().clone(), Tacts=dict(Tacts),
               Sacts={k: v.detach().clone() for k, v in Sacts.items()},
               Sact_grads={k: v.grad.detach().clone() for k, v in Sacts.items()},
               Sfake_grad=Sfake.grad.detach().clone(),
               loss_G_gan=loss_gan.detach(), loss_G_distill=loss_distill.detach(), loss_G_feat=loss_feat.detach(),
               loss_G_vgg=loss_vgg.detach(), loss_G_distill_terms=[t.detach() for t in terms],
               S_grads={k: p.grad.detach().clone() for k, p in S_params.items() if p.grad is not None},
               A_grads={k: p.grad.detach().clone() for k, p in A_params.items()})
    with torch.no_grad():
        if grad_hook is not None:
            out['S_grads'] = grad_hook('S', out['S_grads'])
            if mse:
                out['A_grads'] = grad_hook('A', out['A_grads'])
        adam_update(S_params, out['S_grads'], state['adam_G'], hp['lr_G'], hp['beta1'], hp['beta2'])
        adam_update(A_params, out['A_grads'], state['adam_G'], hp['lr_G'], hp['beta1'], hp['beta2'])
    for p in list(S_params.values()) + list(A_params.values()):
        p.requires_grad_(False)
        p.grad = None
    # ---- D phase
    D_params = {k: v for k, v in D_sd.items() if _is_param(k)}
    for p in D_params.values():
        p.requires_grad_(True)
        p.grad = None
    with torch.no_grad():
        fake = spade_generator_forward(S_sd, S_arch, seg, training=s_train)
    out['Sfake_B_D'] = fake
    pred_fake, pred_real = _discriminate(D_sd, D_arch, seg, fake, real_B)
    loss_D_fake = hinge_multiscale(pred_fake, False, True)
    loss_D_real = hinge_multiscale(pred_real, True, True)
    (loss_D_fake + loss_D_real).backward()
    out.update(loss_D_fake=loss_D_fake.detach(), loss_D_real=loss_D_real.detach(),
               D_grads={k: p.grad.detach().clone() for k, p in D_params.items()})
    with torch.no_grad():
        if grad_hook is not None:
            out['D_grads'] = grad_hook('D', out['D_grads'])
        adam_update(D_params, out['D_grads'], state['adam_D'], hp['lr_D'], hp['beta1'], hp['beta2'])
    for p in D_params.values():
        p.requires_grad_(False)
        p.grad = None
    return out
