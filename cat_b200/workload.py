"""Synthetic distillation workloads: network shapes of the published CAT training scripts (channel
counts produced by the reference's own pruning, committed as JSON by oracle/make_bench_arch.py) plus
reference-style random initialisation and synthetic batches.

Everything here is host-side PyTorch on CPU tensors; nothing is computed on the benchmark path.
"""
import json
import os

import torch

from .engine import discriminator_layers

_GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def load_arch(name):
    with open(os.path.join(_GOLDEN, f'arch_{name}.json')) as f:
        return json.load(f)


def _norm_entries(sd, prefix, C, arch, g, gamma_mode):
    if arch['affine']:
        if gamma_mode == 'uniform':        # synthetic trained teacher: spread scales
            sd[prefix + '.weight'] = torch.rand(C, generator=g)
        elif arch['norm'] == 'batch':       # init_weights touches BatchNorm2d only (models/networks.py:137-141)
            sd[prefix + '.weight'] = 1.0 + 0.02 * torch.randn(C, generator=g)
        else:
            sd[prefix + '.weight'] = torch.ones(C)
        sd[prefix + '.bias'] = torch.zeros(C)
    if arch['norm'] == 'batch' and arch['track_running_stats']:
        sd[prefix + '.running_mean'] = torch.zeros(C)
        sd[prefix + '.running_var'] = torch.ones(C)
        sd[prefix + '.num_batches_tracked'] = torch.zeros((), dtype=torch.long)


def _conv(sd, name, shape, bias_len, g, gain):
    sd[name + '.weight'] = gain * torch.randn(*shape, generator=g)   # init 'normal', gain 0.02 (networks.py:123-124)
    if bias_len:
        sd[name + '.bias'] = torch.zeros(bias_len)


def init_generator(arch, seed, gamma_mode='init', gain=0.02):
    """Reference-format state_dict of an InceptionGenerator with the layer list of
    inception_generator.py:37-135 and the init of models/networks.py:108-144."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    ub = arch['use_bias']
    c0, c1, c2, c3, c4 = arch['widths']
    ks = arch['kernel_sizes']
    _conv(sd, 'down_sampling.1', (c0, arch['input_nc'], 7, 7), c0 if ub else 0, g, gain)
    _norm_entries(sd, 'down_sampling.2', c0, arch, g, gamma_mode)
    _conv(sd, 'down_sampling.4', (c1, c0, 3, 3), c1 if ub else 0, g, gain)
    _norm_entries(sd, 'down_sampling.5', c1, arch, g, gamma_mode)
    _conv(sd, 'down_sampling.7', (c2, c1, 3, 3), c2 if ub else 0, g, gain)
    _norm_entries(sd, 'down_sampling.8', c2, arch, g, gamma_mode)
    for i, blk in enumerate(arch['blocks']):
        pre = f'features.{i}'
        j = 0
        for m, k in zip(blk['res'], ks):
            if m == 0:
                continue
            _conv(sd, f'{pre}.res_ops.{j}.1.0', (m, c2, k, k), m if ub else 0, g, gain)
            _norm_entries(sd, f'{pre}.res_ops.{j}.1.1', m, arch, g, gamma_mode)
            _conv(sd, f'{pre}.res_ops.{j}.4', (c2, m, k, k), c2 if ub else 0, g, gain)
            j += 1
        j = 0
        for m, k in zip(blk['dw'], ks):
            if m == 0:
                continue
            _conv(sd, f'{pre}.dw_ops.{j}.0.0', (m, c2, 1, 1), m if ub else 0, g, gain)
            _norm_entries(sd, f'{pre}.dw_ops.{j}.0.1', m, arch, g, gamma_mode)
            _conv(sd, f'{pre}.dw_ops.{j}.2.0', (m, 1, k, k), m if ub else 0, g, gain)
            _norm_entries(sd, f'{pre}.dw_ops.{j}.2.1', m, arch, g, gamma_mode)
            _conv(sd, f'{pre}.dw_ops.{j}.4', (c2, m, 1, 1), c2 if ub else 0, g, gain)
            j += 1
        if any(blk['res']) or any(blk['dw']):
            _norm_entries(sd, f'{pre}.pw_bn', c2, arch, g, gamma_mode)
    _conv(sd, 'up_sampling.0', (c2, c3, 3, 3), c3 if ub else 0, g, gain)
    _norm_entries(sd, 'up_sampling.1', c3, arch, g, gamma_mode)
    _conv(sd, 'up_sampling.3', (c3, c4, 3, 3), c4 if ub else 0, g, gain)
    _norm_entries(sd, 'up_sampling.4', c4, arch, g, gamma_mode)
    _conv(sd, 'up_sampling.7', (arch['output_nc'], c4, 7, 7), arch['output_nc'], g, gain)
    return sd


def init_discriminator(arch, seed, gain=0.02):
    """Reference-format state_dict of an NLayerDiscriminator (discriminators.py:37-75)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for (ci, cin, cout, stride, has_norm, has_act) in discriminator_layers(arch):
        has_bias = (not has_norm) or arch['use_bias']
        _conv(sd, f'model.{ci}', (cout, cin, 4, 4), cout if has_bias else 0, g, gain)
        if has_norm:
            _norm_entries(sd, f'model.{ci + 1}', cout, arch, g, 'init')
    return sd


def synthetic_batch(B, H, W, seed, pin=False):
    """real_A, real_B ~ U(-1, 1), NCHW fp32 (the dataset range after Normalize(0.5, 0.5),
    data/base_dataset.py:122-128); seed = reference default 233 + rank (base_options.py:33-36)."""
    g = torch.Generator().manual_seed(seed)
    a = torch.rand(B, 3, H, W, generator=g) * 2 - 1
    b = torch.rand(B, 3, H, W, generator=g) * 2 - 1
    if pin:
        a, b = a.pin_memory(), b.pin_memory()
    return a, b


def macs_per_image(arch_json, H, W):
    """Algorithmic MACs per image of one distillation step, SURVEY.md 8(d): T + 3 S + 8 D."""
    s = (H * W) / (256.0 * 256.0)
    T = arch_json['teacher_macs'] * s
    S = arch_json['student_macs'] * s
    d = arch_json['D_arch']
    D = 0
    h, w = H, W
    for (ci, cin, cout, stride, has_norm, has_act) in discriminator_layers(d):
        oh, ow = (h + 2 - 4) // stride + 1, (w + 2 - 4) // stride + 1
        D += cin * cout * 16 * oh * ow
        h, w = oh, ow
    return {'T': T, 'S': S, 'D': D, 'step': T + 3 * S + 8 * D}


# ----------------------------------------------------------------------------------------------------
# SPADE (GauGAN) workload: BASELINE.json configs[3]
# ----------------------------------------------------------------------------------------------------
def spade_arch_for(arch_json, H, W):
    """The committed architecture with the latent size of InceptionSPADEGenerator.compute_latent_vector_size
    (inception_spade_generator.py:47-61) for an H x W crop: sw = W / 2^n_up, sh = H / 2^n_up."""
    n_up = {'normal': 5, 'more': 6, 'most': 7}[arch_json['teacher_arch']['num_upsampling_layers']]
    assert H % (1 << n_up) == 0 and W % (1 << n_up) == 0, (H, W)
    out = dict(arch_json)
    for k in ('teacher_arch', 'student_arch'):
        out[k] = dict(arch_json[k], sh=H >> n_up, sw=W >> n_up)
    return out


def init_from_entries(net, seed, gamma_mode='init'):
    """Random reference-format state_dict for a compiled network, driven by its arena tables: conv weights
    ~ N(0, 2 / fan_in) (activations stay O(1) through the stack; the reference's xavier(0.02) init would only make
    the synthetic activations tiny), conv biases 0, norm weights 1 (+ U(0,1) spread for a synthetic trained
    teacher), running statistics (0, 1), spectral-norm vectors normalised Gaussians."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for arena in (net.arena, net.bufs):
        for name, (off, shape, L, pv) in arena.entries.items():
            if len(shape) == 4:
                fan_in = shape[1] * shape[2] * shape[3]
                sd[name] = torch.randn(*shape, generator=g) * (2.0 / fan_in) ** 0.5
            elif name.endswith('weight_u') or name.endswith('weight_v'):
                sd[name] = torch.nn.functional.normalize(torch.randn(*shape, generator=g), dim=0)
            elif name.endswith('.running_var'):
                sd[name] = torch.ones(*shape)
            elif name.endswith('.weight'):
                sd[name] = torch.rand(*shape, generator=g) if gamma_mode == 'uniform' else torch.ones(*shape)
            else:
                sd[name] = torch.zeros(*shape)
    return sd


def init_spade_reference_sd(arch, seed, gamma_mode='init'):
    """Random state_dict of an InceptionSPADEGenerator of this architecture (no device buffers are created)."""
    from .spade_engine import SpadeGenNet
    return init_from_entries(SpadeGenNet(arch, None, 'cpu', False, False, alloc_only=True), seed, gamma_mode)


def init_multiscale_D_sd(arch, seed):
    from .spade_engine import MultiScaleDis
    return init_from_entries(MultiScaleDis(arch, 2, 64, 64, 'cpu', alloc_only=True), seed)


def init_vgg(seed):
    """Random VGG19 (features[0:30]) weights, He-normal like torchvision's initialisation; the pretrained checkpoint
    of models/modules/loss.py:154 is not available offline."""
    from .spade_engine import VggNet
    return init_from_entries(VggNet(1, 32, 32, 'cpu', alloc_only=True), seed)


def synthetic_spade_batch(B, H, W, n_label, seed, pin=False):
    """label ~ randint(0, n_label) and instance ~ randint(0, 8) in 8x8 constant blocks, image ~ U(-1,1)
    (the dict of data/cityscapes_dataset.py:129-134)."""
    g = torch.Generator().manual_seed(seed)
    lab = torch.randint(0, n_label, (B, 1, H // 8, W // 8), generator=g).repeat_interleave(8, 2).repeat_interleave(8, 3)
    inst = torch.randint(0, 8, (B, 1, H // 8, W // 8), generator=g).repeat_interleave(8, 2).repeat_interleave(8, 3)
    img = torch.rand(B, 3, H, W, generator=g) * 2 - 1
    lab, inst = lab.to(torch.int32).contiguous(), inst.to(torch.int32).contiguous()
    if pin:
        lab, inst, img = lab.pin_memory(), inst.pin_memory(), img.pin_memory()
    return lab, inst, img


def spade_macs_per_image(arch_json, H, W):
    """Algorithmic MACs per image of one SPADE distillation step, SURVEY.md 8(d): T + 4 S + 10 D + 3 V."""
    ph, pw = arch_json.get('profiled_hw', [256, 512])
    s = (H * W) / float(ph * pw)
    T, S = arch_json['teacher_macs'] * s, arch_json['student_macs'] * s
    d = arch_json['D_arch']
    D = 0
    h, w = H, W
    for _ in range(d['num_D']):
        hh, ww, nf, cin = h, w, d['ndf'], d['input_nc']
        for n in range(d['n_layers'] + 1):
            stride = 1 if n >= d['n_layers'] - 1 else 2
            cout = 1 if n == d['n_layers'] else (nf if n == 0 else min(nf * 2, 512))
            oh, ow = (hh + 4 - 4) // stride + 1, (ww + 4 - 4) // stride + 1
            D += cin * cout * 16 * oh * ow
            hh, ww, cin = oh, ow, cout
            nf = cout if n else nf
        h, w = (h + 1) // 2, (w + 1) // 2
    V, cin, hh, ww = 0, 3, H, W
    for c in [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 256, 'M', 512, 512, 512, 512, 'M', 512]:
        if c == 'M':
            hh, ww = hh // 2, ww // 2
        else:
            V += cin * c * 9 * hh * ww
            cin = c
    return {'T': T, 'S': S, 'D': D, 'V': V, 'step': T + 4 * S + 10 * D + 3 * V}
