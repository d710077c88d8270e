"""Mirror of the reference SPADE distiller protocol (distillers/spade_distiller.py:23-170,
distillers/base_spade_distiller.py:26-262, models/spade_model.py:18-215,
models/modules/spade_modules/base_spade_distiller_modules.py:13-210) on top of the fused CUDA step.

What ``trainer.py:79-175`` touches is kept name for name: ``SPADEDistiller(opt)``, ``setup``, ``set_input`` (dict with
``label`` / ``instance`` / ``image`` / ``path``), ``optimize_parameters``, ``get_current_losses``, ``save_networks`` /
``load_networks``, ``update_learning_rate``, ``print_networks``; ``modules_on_one_gpu`` with ``netG_student /
netG_teacher / netD / netAs / mapping_layers``, ``optimizer_G / optimizer_D / optimizers``, ``loss_*``.
``evaluate_model`` runs the student inference over ``self.eval_dataloader`` in eval mode with the reference's metric bookkeeping;
the metric networks come from the caller as ``self.metric_fns`` (see cat_b200/models/base_model.py).  Data loading and logging are
out of scope.
The VGG19 weights come from ``opt.vgg_state_dict`` (torchvision ``vgg19().features`` keys) -- the pretrained checkpoint
the reference downloads (models/modules/loss.py:154) has to be supplied by the caller.
"""
import copy
import os
from collections import OrderedDict

import torch
from torch import nn

from .. import ops
from ..models import networks
from ..spade_distill_engine import SpadeDistillStep
from ..spade_engine import MAPPING_LAYERS
from ..optim import ArenaAdam, EngineOwner


class SPADEDistillerModules(nn.Module):
    """modules_on_one_gpu: the container of base_spade_distiller_modules.py:13-90."""

    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        self.gpu_ids = list(opt.gpu_ids[:1]) if torch.cuda.is_available() else []
        t_opt, s_opt = copy.deepcopy(opt), copy.deepcopy(opt)
        t_opt.norm_G, t_opt.ngf = opt.teacher_norm_G, opt.teacher_ngf
        s_opt.norm_G, s_opt.ngf = opt.student_norm_G, opt.student_ngf
        self.netG_teacher = networks.define_G(opt.input_nc, opt.output_nc, opt.teacher_ngf, opt.teacher_netG, opt.norm, 0,
                                              opt.init_type, opt.init_gain, self.gpu_ids, opt=t_opt)
        arch_S = getattr(opt, 'student_arch', None)   # pruned architecture (what shrink_spade_model produces)
        if arch_S is not None:
            from ..models.spade_networks import InceptionSPADEGenerator
            self.netG_student = networks.init_net(InceptionSPADEGenerator.from_arch(arch_S, s_opt), opt.init_type,
                                                  opt.init_gain, self.gpu_ids)
        else:
            self.netG_student = networks.define_G(opt.input_nc, opt.output_nc, opt.student_ngf, opt.student_netG, opt.norm, 0,
                                                  opt.init_type, opt.init_gain, self.gpu_ids, opt=s_opt)
        self.netD = networks.define_D(opt.input_nc + opt.output_nc, opt.ndf, opt.netD, opt.n_layers_D, opt.norm,
                                      opt.init_type, opt.init_gain, self.gpu_ids, opt=opt)
        self.mapping_layers = list(MAPPING_LAYERS)
        self.netAs = nn.ModuleList()     # adaptor convs: parameters of optimizer_G, used by the 'mse' loss only
        for layer in self.mapping_layers:
            fs, ft = (opt.student_ngf * 16, opt.teacher_ngf * 16) if layer != 'up_1' else (opt.student_ngf * 4, opt.teacher_ngf * 4)
            self.netAs.append(nn.Conv2d(fs, ft, kernel_size=1))
        self.netG_teacher.eval()


class SPADEDistiller(EngineOwner):
    @staticmethod
    def modify_commandline_options(parser, is_train):
        """The flags of base_spade_distiller.py:28-130 / spade_distiller.py:25-84 that the step uses."""
        assert is_train
        parser.add_argument('--num_upsampling_layers', choices=('normal', 'more', 'most'), default='more')
        parser.add_argument('--teacher_netG', type=str, default='inception_spade')
        parser.add_argument('--student_netG', type=str, default='inception_spade')
        parser.add_argument('--teacher_ngf', type=int, default=64)
        parser.add_argument('--student_ngf', type=int, default=48)
        parser.add_argument('--teacher_norm_G', type=str, default='spadesyncbatch3x3')
        parser.add_argument('--student_norm_G', type=str, default='spadesyncbatch3x3')
        parser.add_argument('--restore_teacher_G_path', type=str, required=True)
        parser.add_argument('--restore_student_G_path', type=str, default=None)
        parser.add_argument('--restore_D_path', type=str, default=None)
        parser.add_argument('--restore_A_path', type=str, default=None)
        parser.add_argument('--restore_O_path', type=str, default=None)
        parser.add_argument('--lambda_gan', type=float, default=1)
        parser.add_argument('--lambda_feat', type=float, default=10)
        parser.add_argument('--lambda_vgg', type=float, default=10)
        parser.add_argument('--lambda_distill', type=float, default=10)
        parser.add_argument('--distill_G_loss_type', type=str, default='ka', choices=['ka', 'mse'])
        parser.add_argument('--beta2', type=float, default=0.999)
        parser.add_argument('--no_TTUR', action='store_true')
        parser.add_argument('--num_D', type=int, default=2)
        parser.add_argument('--norm_D', type=str, default='spectralinstance')
        parser.add_argument('--target_flops', type=float, default=0)
        parser.set_defaults(netD='multi_scale', ndf=64, dataset_mode='cityscapes', batch_size=16, init_type='xavier', n_layers_D=4)
        return parser

    def __init__(self, opt):
        assert opt.isTrain
        self.opt = opt
        self.gpu_ids = list(getattr(opt, 'gpu_ids', [0])) or [0]
        # raises without an sm_100 device and libcatb200.so: there is no CPU path.  (Only the kernel emulation of the test
        # suite patches this check out; it then runs the host logic on CPU tensors.)
        ops.require_cuda()
        if getattr(opt, 'distill_G_loss_type', 'ka') not in ('ka', 'mse'):
            raise NotImplementedError('--distill_G_loss_type [%s]: ka | mse' % opt.distill_G_loss_type)
        if getattr(opt, 'gan_mode', 'hinge') != 'hinge':
            raise NotImplementedError('the SPADE distiller uses the hinge GAN loss (spade_model.py default)')
        self.device = torch.device('cuda:%d' % self.gpu_ids[0]) if torch.cuda.is_available() else torch.device('cpu')
        self.save_dir = os.path.join(getattr(opt, 'log_dir', '.'), 'checkpoints')
        self.model_names = ['G_student', 'G_teacher', 'D']
        self.visual_names = ['labels', 'Tfake_B', 'Sfake_B', 'real_B']
        self.loss_names = ['G_gan', 'G_feat', 'G_vgg', 'G_distill', 'D_real', 'D_fake']
        self.modules = self.modules_on_one_gpu = SPADEDistillerModules(opt).to(self.device)
        self.loss_names += ['G_distill%d' % i for i in range(len(self.modules_on_one_gpu.mapping_layers))]
        if opt.no_TTUR:
            self.betas, self.lr_G, self.lr_D = (opt.beta1, opt.beta2), opt.lr, opt.lr
        else:   # base_spade_distiller_modules.py:91-105
            self.betas, self.lr_G, self.lr_D = (0.0, 0.9), opt.lr / 2, opt.lr * 2
        self.optimizer_G = ArenaAdam(self.lr_G, self.betas)
        self.optimizer_D = ArenaAdam(self.lr_D, self.betas)
        self.optimizers = [self.optimizer_G, self.optimizer_D]
        self.engine = None
        self.is_best = False
        self._epoch = 0
        self.image_paths = []

    # ---- protocol -------------------------------------------------------------------------------
    def setup(self, opt, verbose=True):
        self.load_networks(verbose)
        # the reference profiles both generators here (base_spade_distiller.py:178-190), which leaves them in eval()
        # (utils/model_profiling.py:299) until the end of the first evaluate_model (spade_distiller.py:170)
        self.modules_on_one_gpu.netG_student.eval()
        if verbose:
            self.print_networks()

    def _hp(self):
        o = self.opt
        return dict(lambda_gan=o.lambda_gan, lambda_feat=o.lambda_feat, lambda_vgg=o.lambda_vgg, lambda_distill=o.lambda_distill,
                    lr_G=self.lr_G, lr_D=self.lr_D, beta1=self.betas[0], beta2=self.betas[1], n_label=int(o.input_nc), ka_scale=1.0,
                    distill_loss_type=getattr(o, 'distill_G_loss_type', 'ka'))

    def _make_engine(self, B, H, W):
        mm = self.modules_on_one_gpu
        return SpadeDistillStep(mm.netG_teacher.arch(), mm.netG_student.arch(), mm.netD.arch(), self._hp(), B, H, W,
                                device=str(self.device), world_size=int(getattr(self.opt, 'world_size', 1)),
                                use_cuda_graph=bool(getattr(self.opt, 'cuda_graph', True)))

    def _bind_engine(self, eng):
        mm = self.modules_on_one_gpu
        mm.netG_teacher.bind(eng.T)        # copies the module's weights in, then re-points them at the arena
        mm.netG_student.bind(eng.S)
        mm.netD._alias_into(eng.D)
        vgg = getattr(self.opt, 'vgg_state_dict', None)
        if vgg is None:
            raise RuntimeError('opt.vgg_state_dict (torchvision vgg19().features state_dict) is required: the pretrained '
                               'VGG19 of models/modules/loss.py:154 cannot be downloaded here')
        eng.V.load_state_dict(vgg)
        if eng.A is not None:              # 'mse': the adaptor modules alias the engine's adaptor arena
            eng.A.load_state_dicts([net.state_dict() for net in mm.netAs])
            for i, net in enumerate(mm.netAs):
                net.weight.data = eng.A.arena.view('%d.weight' % i)
                net.bias.data = eng.A.arena.view('%d.bias' % i)
        # base_spade_distiller_modules.py:91-105: one parameter group, the student followed by the adaptor convs
        a_params = [p for net in mm.netAs for p in net.parameters()]
        self.optimizer_G.bind([[(mm.netG_student.parameters(), eng.S.arena, eng.step_G),
                                (a_params, eng.A.arena if eng.A is not None else None, eng.step_A if eng.A is not None else None)]])
        self.optimizer_D.bind([[(mm.netD.parameters(), eng.D.arena, eng.step_D)]])

    def set_input(self, input):
        """models/spade_model.py:132-136 (the one-hot / edge preprocessing itself runs inside the step)."""
        self.data = input
        self.image_paths = input.get('path', [])
        self.labels = input['label']
        B, _, H, W = input['image'].shape
        self._ensure_engine(B, H, W)
        self.engine.set_input(input['label'], input['instance'], input['image'])

    def optimize_parameters(self, steps):
        self.engine.set_student_training(self.modules_on_one_gpu.netG_student.training)   # follows .train() / .eval()
        self.engine.step()

    def forward(self, on_one_gpu=False):
        """generate_fake (base_spade_distiller_modules.py:107-112): teacher and student images for the current input."""
        from .. import ops
        eng = self.engine
        eng._preprocess()
        self.input_semantics = ops.nhwc_to_nchw(eng.seg, eng.snc)
        self.real_B = eng.image
        self.Tfake_B = ops.nhwc_to_nchw(eng.T.forward(), 3)
        self.Sfake_B = ops.nhwc_to_nchw(eng.S.forward(), 3)

    def test(self):
        self.forward(on_one_gpu=True)

    def get_current_losses(self):
        L = self.engine.get_losses()
        out = OrderedDict()
        for name in self.loss_names:
            key = ('Specific_loss/' if any(ch.isdigit() for ch in name) else ('D_loss/' if name.startswith('D_') else 'G_loss/')) + name
            out[key] = L[name]
            setattr(self, 'loss_' + name, L[name])
        return out

    def update_learning_rate(self, logger=None):
        """'linear' policy of models/networks.py:80-87 on both optimisers, stepped once per epoch (trainer.py:175)."""
        o = self.opt
        self._epoch += 1
        scale = 1.0 - max(0, self._epoch + 1 - o.nepochs) / float(o.nepochs_decay + 1)
        self.optimizer_G.param_groups[0]['lr'] = self.lr_G * scale
        self.optimizer_D.param_groups[0]['lr'] = self.lr_D * scale
        if self.engine is not None:
            self.engine.set_lr(self.lr_G * scale, self.lr_D * scale)
        msg = 'learning rate = %.7f' % (self.lr_G * scale)
        logger.print_info(msg + '\n') if logger is not None else print(msg)

    def evaluate_model(self, step, save_image=False):
        """spade_distiller.py:96-171: student inference over the evaluation set in eval mode (module forward = its own
        inference network per batch shape on the shared arena), metric bookkeeping, student back in train()."""
        from ..models.base_model import MetricBook, image_names
        from ..models.spade_model import input_semantics
        if not hasattr(self, 'metrics'):
            self.metrics = MetricBook()
        mm, o = self.modules_on_one_gpu, self.opt
        self.is_best = False
        mm.netG_student.eval()
        fakes, names = [], []
        for data_i in getattr(self, 'eval_dataloader', []):
            with torch.no_grad():
                self.Sfake_B = mm.netG_student(input_semantics(data_i, int(o.input_nc), int(mm.netG_student.opt.semantic_nc), self.device))
            fakes.append(self.Sfake_B.cpu())
            names += image_names(data_i.get('path', []))
        ret, self.is_best = self.metrics.update(getattr(self, 'metric_fns', {}), fakes, names)
        mm.netG_student.train()
        return ret

    def print_networks(self):
        mm = self.modules_on_one_gpu
        for name in ('netG_student', 'netG_teacher', 'netD'):
            n = sum(p.numel() for p in getattr(mm, name).parameters())
            print('[Network %s] Total number of parameters : %.3f M' % (name, n / 1e6))

    # ---- checkpoints (file names and key layout of base_spade_distiller_modules.py:177-210) -----------
    def load_networks(self, verbose=True, teacher_only=False, restore_pretrain=True):
        mm = self.modules_on_one_gpu

        def load(net, path):
            if path is not None:
                net.load_state_dict(torch.load(path, map_location='cpu'))
                if verbose:
                    print('Load network at %s' % path)
        load(mm.netG_teacher, getattr(self.opt, 'restore_teacher_G_path', None))
        if teacher_only:
            return
        load(mm.netG_student, getattr(self.opt, 'restore_student_G_path', None))
        load(mm.netD, getattr(self.opt, 'restore_D_path', None))
        if getattr(self.opt, 'restore_A_path', None) is not None:      # base_spade_distiller_modules.py:187-190
            for i, netA in enumerate(mm.netAs):
                load(netA, '%s-%d.pth' % (self.opt.restore_A_path, i))
        if getattr(self.opt, 'restore_O_path', None) is not None:      # base_spade_distiller.py:209-214; the reference
            for i, optimizer in enumerate(self.optimizers):            # puts BOTH optimisers at opt.lr afterwards (TTUR or not)
                optimizer.load_state_dict(torch.load('%s-%d.pth' % (self.opt.restore_O_path, i), map_location='cpu',
                                                     weights_only=False))
                for param_group in optimizer.param_groups:
                    param_group['lr'] = self.opt.lr

    def save_networks(self, epoch):
        os.makedirs(self.save_dir, exist_ok=True)
        mm = self.modules_on_one_gpu

        def cpu_sd(net):
            return OrderedDict((k, v.detach().cpu().clone()) for k, v in net.state_dict().items())
        torch.save(cpu_sd(mm.netG_student), os.path.join(self.save_dir, '%s_net_G.pth' % epoch))
        torch.save(cpu_sd(mm.netD), os.path.join(self.save_dir, '%s_net_D.pth' % epoch))
        for i, net in enumerate(mm.netAs):
            torch.save(cpu_sd(net), os.path.join(self.save_dir, '%s_net_A-%d.pth' % (epoch, i)))
        for i, optimizer in enumerate(self.optimizers):
            torch.save(optimizer.state_dict(), os.path.join(self.save_dir, '%s_optim-%d.pth' % (epoch, i)))
