"""TEST INFRASTRUCTURE ONLY -- drives the *real* reference (snap-research/CAT mounted read-only at
/root/reference) on CPU so that golden vectors for the distillation hot path can be generated.

It needs the reference: /root/reference in the build container, or the copy staged under oracle/_ref/ by
oracle/make_ref.py (the CPU arm of bench.py on the GPU box).  The golden vectors it produces are committed under
tests/golden/ by oracle/make_golden.py; no test on the GPU box imports it.

The shims below are the harness-side stubs listed in SURVEY.md section 8(c); the reference tree is
never modified:
  * import order: torch/torchvision first, /root/reference *appended* to sys.path (its root-level
    profile.py shadows the stdlib module of the same name);
  * model_profiling(..., use_cuda=True) default and torch.cuda.synchronize() inside shrink_* are
    neutralised for CPU;
  * FID network / eval dataloader / real-stat file are replaced by inert stand-ins.
"""
import argparse
import os
import sys
import tempfile

import numpy as np
import torch
import torchvision  # noqa: F401  (must be imported before the reference is put on sys.path)
import torch.nn as nn

# the read-only checkout in the build container; on the GPU box the copy staged by oracle/make_ref.py (oracle/_ref/)
_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')
REF_ROOT = os.environ.get('CATB_REF_ROOT') or ('/root/reference' if os.path.isdir('/root/reference') else _STAGED)


def _install_shims():
    if REF_ROOT not in sys.path:
        sys.path.append(REF_ROOT)
    import distillers.base_inception_distiller as bid
    import distillers.inception_distiller as idm
    import utils.common as uc
    import utils.model_profiling as mp

    class _NoFID(nn.Module):
        BLOCK_INDEX_BY_DIM = {64: 0, 192: 1, 768: 2, 2048: 3}

        def __init__(self, *a, **k):
            super().__init__()
            self.dummy = nn.Parameter(torch.zeros(1))

    bid.InceptionV3 = _NoFID
    bid.create_eval_dataloader = lambda opt, direction=None: []
    real_mp = mp.model_profiling

    def mp_cpu(*a, **k):
        k['use_cuda'] = False
        return real_mp(*a, **k)

    if getattr(idm.model_profiling, '__name__', '') != 'mp_cpu':
        idm.model_profiling = mp_cpu
        uc.model_profiling = mp_cpu
    torch.cuda.synchronize = lambda *a, **k: None


def build_reference_distiller(norm='instance', batch_size=2, height=32, width=32, teacher_ngf=16,
                              student_ngf=8, ndf=8, target_flops=None, gan_mode='hinge',
                              dataset_mode='aligned', lambda_distill=1.0, lambda_recon=100.0,
                              prune_cin_lb=4, seed=0, workdir=None, do_shrink=True, recon_loss_type='l1',
                              distill_G_loss_type='ka'):
    """Build the real InceptionDistiller with a seeded synthetic teacher, run the reference's
    shrink() + init_net() exactly as trainer.py:106-107 does, and return (model, opt)."""
    _install_shims()
    from models import networks
    workdir = workdir or tempfile.mkdtemp(prefix='catref_')
    os.makedirs(os.path.join(workdir, 'logs'), exist_ok=True)
    stat = os.path.join(workdir, 'real_stat.npz')
    np.savez(stat, mu=np.zeros(4), sigma=np.eye(4))
    torch.manual_seed(seed)
    track = norm == 'batch'
    topt = argparse.Namespace(channels=None, channels_reduction_factor=6, kernel_sizes=[1, 3, 5],
                              norm_momentum=0.1, norm_epsilon=1e-5, active_fn='nn.ReLU',
                              norm_affine=True, norm_track_running_stats=track)
    teacher = networks.define_G(3, 3, teacher_ngf, 'inception_9blocks', norm, 0, 'normal', 0.02, [],
                                opt=topt)
    g = torch.Generator().manual_seed(seed + 1)
    for m in teacher.modules():
        if isinstance(m, (nn.InstanceNorm2d, nn.BatchNorm2d)):
            if m.weight is not None:
                # spread the norm scales so that pruning is not degenerate
                m.weight.data = torch.rand(m.weight.shape, generator=g)
                m.bias.data = 0.1 * torch.randn(m.bias.shape, generator=g)
            if getattr(m, 'running_mean', None) is not None:
                m.running_mean.data = 0.05 * torch.randn(m.running_mean.shape, generator=g)
                m.running_var.data = 0.5 + torch.rand(m.running_var.shape, generator=g)
    # the synthetic teacher uses a larger init gain so activations do not collapse to ~0
    for m in teacher.modules():
        if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
            m.weight.data = m.weight.data * 4.0
    tpath = os.path.join(workdir, 'teacher.pth')
    torch.save(teacher.state_dict(), tpath)

    argv = ['distill.py', '--dataroot', os.path.join(workdir, 'none'), '--distiller', 'inception',
            '--log_dir', os.path.join(workdir, 'logs'), '--restore_teacher_G_path', tpath,
            '--real_stat_path', stat, '--teacher_ngf', str(teacher_ngf), '--student_ngf',
            str(student_ngf), '--pretrained_ngf', str(teacher_ngf), '--ndf', str(ndf), '--gpu_ids',
            '-1', '--norm', norm, '--norm_affine', '--norm_affine_D', '--channels_reduction_factor',
            '6', '--kernel_sizes', '1', '3', '5', '--lambda_distill', str(lambda_distill),
            '--lambda_recon', str(lambda_recon), '--prune_cin_lb', str(prune_cin_lb),
            '--gan_mode', gan_mode, '--dataset_mode', dataset_mode,
            '--distill_G_loss_type', distill_G_loss_type, '--batch_size', str(batch_size), '--recon_loss_type', recon_loss_type]
    if target_flops is not None:
        argv += ['--target_flops', str(target_flops)]
    if track:
        argv += ['--norm_track_running_stats']
    old_argv = sys.argv
    sys.argv = argv
    try:
        from options.distill_options import DistillOptions
        opt = DistillOptions().parse(verbose=False)
    finally:
        sys.argv = old_argv
    opt.data_channel, opt.data_height, opt.data_width = 3, height, width
    from distillers import create_distiller
    model = create_distiller(opt)
    model.setup(opt, verbose=False)
    if do_shrink and target_flops is not None:
        from utils.common import shrink
        shrink(model, opt)
        model.netG_student = networks.init_net(model.netG_student, opt.init_type, opt.init_gain,
                                               []).to(model.device)
    return model, opt


def generator_arch(net, opt):
    """Describe an InceptionGenerator instance (teacher or pruned student) as a plain dict."""
    ds, us = net.down_sampling, net.up_sampling
    norm_mod = ds[2]
    arch = {
        'input_nc': ds[1].in_channels,
        'output_nc': us[7].out_channels,
        'widths': [ds[1].out_channels, ds[4].out_channels, ds[7].out_channels,
                   us[0].out_channels, us[3].out_channels],
        'kernel_sizes': list(opt.kernel_sizes),
        'norm': 'batch' if isinstance(norm_mod, nn.BatchNorm2d) else 'instance',
        'affine': bool(norm_mod.affine),
        'track_running_stats': bool(norm_mod.track_running_stats),
        'eps': float(norm_mod.eps),
        'momentum': float(norm_mod.momentum),
        'use_bias': ds[1].bias is not None,
        'blocks': [{'res': [int(c) for c in b.res_channels], 'dw': [int(c) for c in b.dw_channels]}
                   for b in net.features],
    }
    return arch


def discriminator_arch(net, opt, input_nc):
    norm_mod = net.model[3]
    return {
        'input_nc': input_nc, 'ndf': int(opt.ndf), 'n_layers': int(opt.n_layers_D),
        'norm': 'batch' if isinstance(norm_mod, nn.BatchNorm2d) else 'instance',
        'affine': bool(norm_mod.affine),
        'track_running_stats': bool(norm_mod.track_running_stats),
        'eps': float(norm_mod.eps), 'momentum': float(norm_mod.momentum),
        'use_bias': net.model[2].bias is not None,
    }
