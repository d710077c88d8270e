"""Shared protocol of the teacher-training model mirrors (models/base_model.py:12-232 in the reference): what
``trainer.py:79-175`` calls on a model -- ``setup``, ``set_input``, ``optimize_parameters``, ``get_current_losses``,
``update_learning_rate``, ``save_networks`` / ``load_networks``, ``print_networks``, ``eval`` / ``train`` / ``test`` -- on
top of an engine step (cat_b200/train_engine.py).  The networks are the module-tree mirrors of
cat_b200.models.networks whose parameters alias the engine arenas, so checkpoints written here load in the reference
(and in the distillers as ``restore_teacher_G_path`` / ``restore_D_path``) and vice versa.
``evaluate_model`` runs the generator inference over ``self.eval_dataloader`` (an iterable of dataset dicts set by the caller)
in eval mode and keeps the reference's best / mean-of-three bookkeeping; the metric networks (FID InceptionV3, DRN mIoU;
SURVEY.md section 2: out of scope) come from the caller as ``self.metric_fns = {'fid': f(fakes), 'mIoU': f(fakes, names)}``.
Data loading and logging are out of scope.
"""
import os
from collections import OrderedDict

import torch

from .. import ops
from ..optim import ArenaAdam, EngineOwner


ArenaOptimizer = ArenaAdam     # former name


class MetricBook:
    """Best / mean-of-the-last-three bookkeeping of the reference's evaluate_model (e.g. models/pix2pix_model.py:252-281)."""

    def __init__(self):
        self.best_fid, self.best_mIoU, self.fids, self.mIoUs = 1e9, -1e9, [], []

    def update(self, fns, fakes, names, suffix=''):
        """-> (metrics dict, is_best)"""
        ret, best = {}, False
        if 'fid' in fns:
            fid = float(fns['fid'](fakes))
            if fid < self.best_fid:
                best, self.best_fid = True, fid
            self.fids = (self.fids + [fid])[-3:]
            ret.update({'metric/fid' + suffix: fid, 'metric/fid%s-mean' % suffix: sum(self.fids) / len(self.fids),
                        'metric/fid%s-best' % suffix: self.best_fid})
        if 'mIoU' in fns:
            mIoU = float(fns['mIoU'](fakes, names))
            if mIoU > self.best_mIoU:
                best, self.best_mIoU = True, mIoU
            self.mIoUs = (self.mIoUs + [mIoU])[-3:]
            ret.update({'metric/mIoU' + suffix: mIoU, 'metric/mIoU%s-mean' % suffix: sum(self.mIoUs) / len(self.mIoUs),
                        'metric/mIoU%s-best' % suffix: self.best_mIoU})
        return ret, best


def image_names(paths):
    return [os.path.splitext(os.path.basename(p))[0] for p in paths]


class BaseModel(EngineOwner):
    """Subclasses define ``model_names`` (net<name> attributes), ``loss_names``, ``_make_engine(B, H, W)`` and
    ``_set_engine_input(input)``."""

    def __init__(self, opt):
        assert opt.isTrain
        self.opt = opt
        self.isTrain = True
        self.gpu_ids = list(getattr(opt, 'gpu_ids', [0])) or [0]
        # raises without an sm_100 device and libcatb200.so: there is no CPU path.  (Only the kernel emulation of the
        # test suite patches this check out; it then runs the host logic on CPU tensors.)
        ops.require_cuda()
        self.device = torch.device('cuda:%d' % self.gpu_ids[0]) if torch.cuda.is_available() else torch.device('cpu')
        self._ids = self.gpu_ids[:1] if self.device.type == 'cuda' else []
        self.save_dir = os.path.join(getattr(opt, 'log_dir', '.'), 'checkpoints')
        self.model_names, self.loss_names, self.visual_names, self.image_paths = [], [], [], []
        self.optimizers = []
        self.engine = None
        self.is_best = False
        self.metric = 0
        self._epoch = 0
        self.eval_dataloader, self.metric_fns, self.metrics = [], {}, MetricBook()

    @staticmethod
    def modify_commandline_options(parser, is_train):
        return parser

    def _net(self, name):
        return getattr(self, 'net' + name)

    # ---- protocol -------------------------------------------------------------------------------
    def setup(self, opt, verbose=True):
        self.load_networks(verbose)
        if verbose:
            self.print_networks()

    def optimize_parameters(self, steps):
        self.engine.step()

    def get_current_losses(self):
        """One device synchronisation, like float(loss) in base_model.py:187; same keys and order."""
        L = self.engine.get_losses()
        out = OrderedDict()
        for name in self.loss_names:
            if name not in L:
                continue
            key = ('Specific_loss/' if any(ch.isdigit() for ch in name) else ('D_loss/' if name.startswith('D_') else 'G_loss/')) + name
            out[key] = L[name]
            setattr(self, 'loss_' + name, L[name])
        return out

    def _lr_scale(self):
        """'linear' policy of models/networks.py:80-87, stepped once per epoch (trainer.py:175)."""
        o = self.opt
        if getattr(o, 'lr_policy', 'linear') != 'linear':
            raise NotImplementedError('lr_policy [%s]: the CAT scripts use the linear policy' % o.lr_policy)
        self._epoch += 1
        return 1.0 - max(0, self._epoch + 1 - o.nepochs) / float(o.nepochs_decay + 1)

    def update_learning_rate(self, logger=None):
        scale = self._lr_scale()
        lrs = [base * scale for base in self._base_lrs()]
        for opt_, lr in zip(self.optimizers, lrs):
            for pg in opt_.param_groups:
                pg['lr'] = lr
        if self.engine is not None:
            self.engine.set_lr(*lrs)
        msg = 'learning rate = %.7f' % lrs[0]
        logger.print_info(msg + '\n') if logger is not None else print(msg)

    def _base_lrs(self):
        return [self.opt.lr, self.opt.lr]

    def eval(self):
        for name in self.model_names:
            self._net(name).eval()

    def train(self):
        for name in self.model_names:
            self._net(name).train()

    def test(self):
        with torch.no_grad():
            self.forward()

    def get_image_paths(self):
        return self.image_paths

    def get_current_visuals(self):
        return OrderedDict((n, getattr(self, n)) for n in self.visual_names if hasattr(self, n))

    def set_requires_grad(self, nets, requires_grad=False):
        pass    # the engine step freezes / unfreezes the discriminators by construction

    def evaluate_model(self, step, save_image=False):
        """Generator inference over the evaluation set in eval mode + metric bookkeeping (subclasses: _eval_batch)."""
        self.is_best = False
        self.eval()
        fakes, names = [], []
        for data_i in self.eval_dataloader:
            fake, paths = self._eval_batch(data_i)
            fakes.append(fake.cpu())
            names += image_names(paths)
        ret, self.is_best = self.metrics.update(self.metric_fns, fakes, names)
        self.train()
        return ret

    def print_networks(self):
        for name in self.model_names:
            n = sum(p.numel() for p in self._net(name).parameters())
            print('[Network %s] Total number of parameters : %.3f M' % (name, n / 1e6))

    # ---- checkpoints (file names and key layout of base_model.py:196-219, trainer.py:146-160) -------
    def load_networks(self, verbose=True, teacher_only=False, restore_pretrain=True):
        for name in self.model_names:
            path = getattr(self.opt, 'restore_%s_path' % name, None)
            if path is not None:
                self._net(name).load_state_dict(torch.load(path, map_location='cpu'))
                if verbose:
                    print('Load network at %s' % path)
        if getattr(self.opt, 'restore_O_path', None) is not None:      # models/spade_model.py:316-322 (applied at compile time)
            for i, optimizer in enumerate(self.optimizers):
                optimizer.load_state_dict(torch.load('%s-%d.pth' % (self.opt.restore_O_path, i), map_location='cpu',
                                                     weights_only=False))
                for param_group in optimizer.param_groups:
                    param_group['lr'] = self.opt.lr

    def save_networks(self, epoch):
        os.makedirs(self.save_dir, exist_ok=True)
        for name in self.model_names:
            sd = OrderedDict((k, v.detach().cpu().clone()) for k, v in self._net(name).state_dict().items())
            torch.save(sd, os.path.join(self.save_dir, '%s_net_%s.pth' % (epoch, name)))
        for i, optimizer in enumerate(self.optimizers):
            torch.save(optimizer.state_dict(), os.path.join(self.save_dir, '%s_optim-%d.pth' % (epoch, i)))
