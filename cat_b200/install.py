"""Drop-in installation: run the reference's own ``distill.py`` / ``train.py`` / ``trainer.py`` -- unmodified -- on
cat_b200.

The reference reaches its hot path through four imports (trainer.py:20-25, 41-46):

    from distillers import create_distiller          from models import create_model
    from utils.common import shrink                  from models.networks import init_net

``install()`` registers the mirrors under exactly those module names in ``sys.modules`` (``distillers``,
``distillers.inception_distiller``, ``distillers.spade_distiller``, ``models``, ``models.networks``, ``models.*_model`` and a
``utils.common`` that carries ``shrink`` / ``KA``), so that every later ``import`` of the reference's driver code -- options
(options/distill_options.py:189-206 asks ``distillers.get_option_setter`` for the flags), data loading, logging, the training
loop -- binds to this library.  Everything else of the reference (``options``, ``data``, ``utils.logger``, ``utils.util``,
``common``) is left alone and runs as shipped.  Launcher (two lines, next to the reference checkout):

    import cat_b200.install; cat_b200.install.install()
    import runpy; runpy.run_path('distill.py', run_name='__main__')

Multi-GPU: the reference shards with ``nn.DataParallel`` inside one process; here the job is one process per GPU
(``torchrun --nproc-per-node N launcher.py ...`` with ``--gpu_ids 0``).  ``install()`` reads the torchrun environment,
binds the process to ``LOCAL_RANK``, creates the NCCL process group and re-maps ``--gpu_ids`` to the local device; the
distillers pick ``WORLD_SIZE`` up as ``opt.world_size`` (gradient all-reduce per optimiser).
"""
import os
import sys
import types


def install(init_distributed=True):
    import torch
    from . import distillers, models, prune
    from .distillers import inception_distiller, spade_distiller
    from .models import base_model, cycle_gan_model, networks, pix2pix_model, spade_model
    alias = {
        'distillers': distillers, 'distillers.inception_distiller': inception_distiller,
        'distillers.spade_distiller': spade_distiller, 'models': models, 'models.networks': networks,
        'models.base_model': base_model, 'models.pix2pix_model': pix2pix_model, 'models.cycle_gan_model': cycle_gan_model,
        'models.spade_model': spade_model,
    }
    for name, mod in alias.items():
        sys.modules[name] = mod
    # utils.common: the two names the hot path takes from it; the `utils` package itself stays the reference's (logger, util)
    uc = types.ModuleType('utils.common')
    uc.__doc__ = 'cat_b200 stand-in for utils/common.py: shrink (utils/common.py:872-878) and KA (:38-46)'
    uc.shrink = prune.shrink
    uc.KA = KA
    sys.modules['utils.common'] = uc
    try:
        import utils as _utils            # the reference's package (cwd = its checkout)
        _utils.common = uc
    except ImportError:                   # no reference on the path: a bare namespace so that `utils.common` resolves
        pkg = types.ModuleType('utils')
        pkg.__path__ = []
        pkg.common = uc
        sys.modules['utils'] = pkg
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world > 1 and init_distributed:
        import torch.distributed as dist
        local = int(os.environ.get('LOCAL_RANK', '0'))
        if torch.cuda.is_available():
            torch.cuda.set_device(local)
        if not dist.is_initialized():
            dist.init_process_group('nccl' if torch.cuda.is_available() else 'gloo')
        _remap_gpu_ids(local)
    return alias


def _remap_gpu_ids(local):
    """`--gpu_ids X` on the command line -> the local device of this rank (every rank is launched with the same argv)."""
    argv = sys.argv
    for i, a in enumerate(argv):
        if a == '--gpu_ids' and i + 1 < len(argv):
            argv[i + 1] = str(local)
            return
    argv += ['--gpu_ids', str(local)]


def KA(X, Y):
    """utils/common.py:38-46 as a function of two activation tensors [B, C, H, W] (fp32, NCHW): the kernel-alignment value
    computed by the catb_gram / catb_ka_finish kernels (the distillers call the kernels directly; this entry point serves
    code that imports KA by name)."""
    import torch
    from . import ops
    ops.require_cuda()
    B = X.shape[0]
    dev = X.device
    out = []
    for t in (X, Y):
        t = t.reshape(B, t.shape[1], -1, 1) if t.dim() == 4 else t.reshape(B, -1, 1, 1)
        act = ops.Act.empty(B, t.shape[2], t.shape[3], t.shape[1], dev, zero=True)
        ops.nchw_to_nhwc(t.contiguous().float(), act)
        G = torch.zeros(B, B, dtype=torch.float32, device=dev)
        ops.gram(act, G)
        out.append(G)
    loss = torch.zeros(1, dtype=torch.float32, device=dev)
    val = torch.zeros(1, dtype=torch.float32, device=dev)
    coef = torch.zeros(B, B, dtype=torch.float32, device=dev)
    ops.ka_finish(out[0], out[1], B, 1.0, loss, val, coef)
    return val[0]
