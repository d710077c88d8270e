"""TEST INFRASTRUCTURE ONLY -- drives the *real* reference SPADEDistiller (snap-research/CAT mounted
read-only at /root/reference) on CPU so that golden vectors for the SPADE distillation step
 can be generated.  Only works in the build container.

Shims (the reference tree is never modified):
  * everything oracle/ref_harness.py installs (import order, CPU profiling, inert FID network);
  * torch.optim.Adam(betas=(0, 0.9)) -- an int beta is rejected by torch >= 2 -> coerced to float;
  * torchvision.models.vgg19(pretrained=True) needs the network -> vgg19(weights=None) under a fixed
    seed, so the VGG-loss parity is against RANDOM VGG weights ("pretrained parity unpinned");
  * --no_fid --no_mIoU, empty eval dataloader, dummy real-statistics file.
"""
import argparse
import os
import sys
import tempfile

import numpy as np
import torch
import torchvision

from oracle.ref_harness import REF_ROOT, _install_shims

VGG_SEED = 1234


def _install_spade_shims():
    _install_shims()
    if REF_ROOT not in sys.path:
        sys.path.append(REF_ROOT)
    import distillers.base_spade_distiller as bsd
    import models.modules.loss as loss_mod
    import utils.model_profiling as mp

    bsd.create_eval_dataloader = lambda opt, direction=None: []
    if getattr(bsd.model_profiling, '__name__', '') != 'mp_cpu':
        real_mp = mp.model_profiling

        def mp_cpu(*a, **k):
            k['use_cuda'] = False
            return real_mp(*a, **k)
        bsd.model_profiling = mp_cpu
    import utils.common as uc
    if getattr(uc.model_profiling, '__name__', '') != 'mp_cpu':
        uc.model_profiling = bsd.model_profiling

    real_vgg19 = torchvision.models.vgg19

    def vgg19_offline(pretrained=False, **kw):
        st = torch.random.get_rng_state()
        torch.manual_seed(VGG_SEED)
        net = real_vgg19(weights=None)
        torch.random.set_rng_state(st)
        return net

    if getattr(loss_mod.torchvision.models.vgg19, '__name__', '') != 'vgg19_offline':
        loss_mod.torchvision.models.vgg19 = vgg19_offline

    real_adam = torch.optim.Adam
    if getattr(real_adam, '__name__', '') != 'AdamFloatBetas':
        class AdamFloatBetas(real_adam):
            def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), **kw):
                super().__init__(params, lr=lr, betas=(float(betas[0]), float(betas[1])), **kw)
        torch.optim.Adam = AdamFloatBetas
    uc.Adam = torch.optim.Adam      # utils/common.py:14 binds the class at import time (shrink_spade_model :854)


def build_reference_spade_distiller(batch_size=2, crop_size=64, aspect_ratio=2.0, teacher_ngf=12, student_ngf=12,
                                    ndf=8, input_nc=6, target_flops=None, prune_cin_lb=4, lambda_distill=0.5,
                                    seed=0, workdir=None, do_shrink=True, num_upsampling_layers='more',
                                    distill_G_loss_type='ka'):
    """Real SPADEDistiller with a seeded synthetic teacher; shrink() as trainer.py:106-107 when target_flops."""
    _install_spade_shims()
    from models import networks
    workdir = workdir or tempfile.mkdtemp(prefix='catref_spade_')
    os.makedirs(os.path.join(workdir, 'logs'), exist_ok=True)
    stat = os.path.join(workdir, 'real_stat.npz')
    np.savez(stat, mu=np.zeros(4), sigma=np.eye(4))
    torch.manual_seed(seed)
    semantic_nc = input_nc + 1    # + instance edge map; no dont-care label (cityscapes_dataset.py:32-46)
    topt = argparse.Namespace(ngf=teacher_ngf, norm_G='spadesyncbatch3x3', semantic_nc=semantic_nc,
                              num_upsampling_layers=num_upsampling_layers, crop_size=crop_size,
                              aspect_ratio=aspect_ratio, channels=None, channels_reduction_factor=6,
                              kernel_sizes=[1, 3, 5], active_fn='nn.ReLU', norm_momentum=0.1, norm_epsilon=1e-5)
    teacher = networks.define_G(input_nc, 3, teacher_ngf, 'inception_spade', 'instance', 0, 'xavier', 0.02, [], opt=topt)
    g = torch.Generator().manual_seed(seed + 1)
    for m in teacher.modules():
        if hasattr(m, 'running_mean') and getattr(m, 'weight', None) is not None:
            m.weight.data = torch.rand(m.weight.shape, generator=g)
            m.bias.data = 0.1 * torch.randn(m.bias.shape, generator=g)
        if getattr(m, 'running_mean', None) is not None:
            m.running_mean.data = 0.05 * torch.randn(m.running_mean.shape, generator=g)
            m.running_var.data = 0.5 + torch.rand(m.running_var.shape, generator=g)
    tpath = os.path.join(workdir, 'teacher.pth')
    torch.save(teacher.state_dict(), tpath)

    argv = ['distill.py', '--dataroot', os.path.join(workdir, 'none'), '--distiller', 'spade',
            '--log_dir', os.path.join(workdir, 'logs'), '--restore_teacher_G_path', tpath,
            '--restore_pretrained_G_path', tpath, '--pretrained_netG', 'inception_spade',
            '--real_stat_path', stat, '--teacher_ngf', str(teacher_ngf), '--student_ngf', str(student_ngf),
            '--pretrained_ngf', str(teacher_ngf), '--ndf', str(ndf), '--gpu_ids', '-1', '--no_fid', '--no_mIoU',
            '--teacher_norm_G', 'spadesyncbatch3x3', '--student_norm_G', 'spadesyncbatch3x3',
            '--channels_reduction_factor', '6', '--kernel_sizes', '1', '3', '5',
            '--lambda_distill', str(lambda_distill), '--prune_cin_lb', str(prune_cin_lb),
            '--distill_G_loss_type', distill_G_loss_type, '--batch_size', str(batch_size), '--input_nc', str(input_nc),
            '--crop_size', str(crop_size), '--load_size', str(crop_size), '--aspect_ratio', str(aspect_ratio),
            '--num_upsampling_layers', num_upsampling_layers]
    if target_flops is not None:
        argv += ['--target_flops', str(target_flops)]
    old_argv = sys.argv
    sys.argv = argv
    try:
        from options.distill_options import DistillOptions
        opt = DistillOptions().parse(verbose=False)
    finally:
        sys.argv = old_argv
    H = int(round(crop_size / aspect_ratio))
    opt.data_channel, opt.data_height, opt.data_width = semantic_nc, H, crop_size
    from distillers import create_distiller
    model = create_distiller(opt)
    model.setup(opt, verbose=False)
    if do_shrink and target_flops is not None:
        from utils.common import shrink
        shrink(model, opt)
    return model, opt


def spade_generator_arch(net, opt_ngf_unused=None):
    """Describe an InceptionSPADEGenerator (teacher or pruned student) as a plain dict."""
    blocks = {}
    names = ['head_0', 'G_middle_0', 'G_middle_1', 'up_0', 'up_1', 'up_2', 'up_3']
    if net.opt.num_upsampling_layers == 'most':
        names.append('up_4')
    for n in names:
        b = getattr(net, n)
        blocks[n] = {
            'fin': int(b.input_dim), 'fout': int(b.output_dim),
            'res': [int(c) for c in b.res_channels], 'dw': [int(c) for c in b.dw_channels],
            'spade_res': [int(c) for c in b.spade.res_channels], 'spade_dw': [int(c) for c in b.spade.dw_channels],
            'learned_shortcut': b.shortcut is not None,
        }
    return {
        'semantic_nc': int(net.opt.semantic_nc), 'fc_out': int(net.fc.out_channels), 'sh': int(net.sh), 'sw': int(net.sw),
        'num_upsampling_layers': net.opt.num_upsampling_layers, 'kernel_sizes': [int(k) for k in net.opt.kernel_sizes],
        'final_nc': int(net.conv_img.in_channels), 'block_names': names, 'blocks': blocks,
        'eps': 1e-5, 'momentum': 0.1, 'active_fn': getattr(net.opt, 'active_fn', 'nn.ReLU'),
    }


def multiscale_D_arch(net, opt):
    return {'input_nc': int(opt.semantic_nc + opt.output_nc), 'ndf': int(opt.ndf), 'n_layers': int(opt.n_layers_D),
            'num_D': int(opt.num_D), 'norm_D': opt.norm_D}
