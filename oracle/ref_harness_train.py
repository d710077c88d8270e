"""TEST INFRASTRUCTURE ONLY -- drives the *real* reference training models (Pix2PixModel, CycleGANModel,
SPADEModel of snap-research/CAT, mounted read-only at /root/reference) on CPU so that golden vectors for the
teacher-training steps can be generated.  Only works in the build container.

Shims (the reference tree is never modified): everything oracle/ref_harness.py / ref_harness_spade.py install,
plus the same inert FID network / eval dataloader stand-ins on the three model modules.
"""
import os
import sys
import tempfile

import numpy as np
import torch
import torch.nn as nn

from oracle.ref_harness import REF_ROOT, _install_shims
from oracle.ref_harness_spade import _install_spade_shims


def _install_train_shims():
    _install_shims()
    _install_spade_shims()
    if REF_ROOT not in sys.path:
        sys.path.append(REF_ROOT)
    import models.cycle_gan_model as cgm
    import models.pix2pix_model as ppm
    import models.spade_model as spm

    class _NoFID(nn.Module):
        BLOCK_INDEX_BY_DIM = {64: 0, 192: 1, 768: 2, 2048: 3}

        def __init__(self, *a, **k):
            super().__init__()
            self.dummy = nn.Parameter(torch.zeros(1))

    for mod in (cgm, ppm, spm):
        mod.InceptionV3 = _NoFID
        mod.create_eval_dataloader = lambda opt, direction=None: []


def _parse(argv):
    old = sys.argv
    sys.argv = argv
    try:
        from options.train_options import TrainOptions
        return TrainOptions().parse(verbose=False)
    finally:
        sys.argv = old


def _workdir():
    wd = tempfile.mkdtemp(prefix='catref_train_')
    os.makedirs(os.path.join(wd, 'logs'), exist_ok=True)
    stat = os.path.join(wd, 'real_stat.npz')
    np.savez(stat, mu=np.zeros(4), sigma=np.eye(4))
    return wd, stat


def _common(wd, model, ngf, ndf, norm, batch_size, gan_mode):
    argv = ['train.py', '--dataroot', os.path.join(wd, 'none'), '--model', model, '--log_dir', os.path.join(wd, 'logs'),
            '--ngf', str(ngf), '--ndf', str(ndf), '--gpu_ids', '-1', '--norm', norm, '--norm_affine', '--norm_affine_D',
            '--channels_reduction_factor', '6', '--kernel_sizes', '1', '3', '5', '--gan_mode', gan_mode,
            '--batch_size', str(batch_size)]
    if norm == 'batch':
        argv += ['--norm_track_running_stats']
    return argv


def build_reference_pix2pix(norm='batch', batch_size=2, ngf=8, ndf=8, gan_mode='hinge', lambda_recon=100.0,
                            recon_loss_type='l1', seed=0):
    """Real Pix2PixModel (models/pix2pix_model.py) with the flags of scripts/pix2pix/*/train_inception_teacher.sh."""
    _install_train_shims()
    wd, stat = _workdir()
    torch.manual_seed(seed)
    argv = _common(wd, 'pix2pix', ngf, ndf, norm, batch_size, gan_mode) + [
        '--real_stat_path', stat, '--lambda_recon', str(lambda_recon), '--recon_loss_type', recon_loss_type]
    opt = _parse(argv)
    from models import create_model
    model = create_model(opt, verbose=False)
    return model, opt


def build_reference_cyclegan(norm='instance', batch_size=2, ngf=8, ndf=8, gan_mode='lsgan', lambda_identity=0.5,
                             pool_size=50, seed=0):
    """Real CycleGANModel (models/cycle_gan_model.py) with the flags of scripts/cycle_gan/*/train_inception_teacher.sh."""
    _install_train_shims()
    wd, stat = _workdir()
    torch.manual_seed(seed)
    argv = _common(wd, 'cycle_gan', ngf, ndf, norm, batch_size, gan_mode) + [
        '--real_stat_A_path', stat, '--real_stat_B_path', stat, '--lambda_identity', str(lambda_identity),
        '--pool_size', str(pool_size), '--dataset_mode', 'unaligned']
    opt = _parse(argv)
    from models import create_model
    model = create_model(opt, verbose=False)
    return model, opt


def build_reference_spade(batch_size=2, crop_size=128, aspect_ratio=2.0, ngf=6, ndf=8, input_nc=6, seed=0,
                          num_upsampling_layers='more'):
    """Real SPADEModel (models/spade_model.py) with the flags of scripts/gaugan/cityscapes/train_inception_teacher.sh."""
    _install_train_shims()
    wd, stat = _workdir()
    torch.manual_seed(seed)
    argv = ['train.py', '--dataroot', os.path.join(wd, 'none'), '--model', 'spade', '--log_dir', os.path.join(wd, 'logs'),
            '--ngf', str(ngf), '--ndf', str(ndf), '--gpu_ids', '-1', '--no_fid', '--no_mIoU', '--real_stat_path', stat,
            '--norm_G', 'spadesyncbatch3x3', '--channels_reduction_factor', '6', '--kernel_sizes', '1', '3', '5',
            '--batch_size', str(batch_size), '--input_nc', str(input_nc), '--crop_size', str(crop_size),
            '--load_size', str(crop_size), '--aspect_ratio', str(aspect_ratio),
            '--num_upsampling_layers', num_upsampling_layers]
    opt = _parse(argv)
    from models import create_model
    model = create_model(opt, verbose=False)
    return model, opt
