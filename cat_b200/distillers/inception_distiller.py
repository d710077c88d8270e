"""Mirror of the reference distiller protocol (distillers/inception_distiller.py:32-281 and
distillers/base_inception_distiller.py:28-403) on top of the fused CUDA step.

What ``trainer.py:79-175`` calls is kept name for name: ``InceptionDistiller(opt)``, ``setup``, ``set_input``,
``optimize_parameters``, ``get_current_losses``, ``save_networks`` / ``load_networks``, ``update_learning_rate``,
``print_networks``, ``test``; attributes ``netG_teacher / netG_student / netD / netAs``, ``optimizers``,
``Tfake_B / Sfake_B``, ``loss_*``.  The networks are the module-tree mirrors of cat_b200.models.networks whose
parameters alias the engine arenas, so checkpoints written here load in the reference and vice versa.
``evaluate_model`` runs the generator inference over ``self.eval_dataloader`` through the module mirrors and keeps the
reference's best / mean bookkeeping; the metric networks themselves (FID InceptionV3, DRN mIoU; SURVEY.md section 2: out of
scope) are supplied by the caller as ``self.metric_fns = {'fid': f(fakes), 'mIoU': f(fakes, names)}``.  Data loading and
logging are out of scope.
"""
import os
from collections import OrderedDict

import torch
from torch import nn

from .. import ops
from ..distill_engine import DistillStep
from ..engine import MAPPING_LAYERS
from ..models import networks
from ..models.base_model import MetricBook, image_names
from ..optim import ArenaAdam, EngineOwner


_ArenaOptimizer = ArenaAdam    # former name


class InceptionDistiller(EngineOwner):
    @staticmethod
    def modify_commandline_options(parser, is_train):
        """Every flag of base_inception_distiller.py:30-101 and inception_distiller.py:34-76, with the reference's types,
        defaults and choices, so that the published scripts parse unchanged (flags of features outside the hot path --
        weight transfer from a pretrained generator, dropout -- are accepted and reported when used)."""
        assert is_train
        parser.add_argument('--teacher_netG', type=str, default='inception_9blocks', choices=['inception_9blocks'])
        parser.add_argument('--student_netG', type=str, default='inception_9blocks', choices=['inception_9blocks'])
        parser.add_argument('--teacher_ngf', type=int, default=64)
        parser.add_argument('--student_ngf', type=int, default=48)
        parser.add_argument('--restore_teacher_G_path', type=str, required=True)
        parser.add_argument('--restore_student_G_path', type=str, default=None)
        parser.add_argument('--restore_A_path', type=str, default=None)
        parser.add_argument('--restore_D_path', type=str, default=None)
        parser.add_argument('--restore_O_path', type=str, default=None)
        parser.add_argument('--recon_loss_type', type=str, default='l1', choices=['l1', 'l2', 'smooth_l1', 'vgg'])
        parser.add_argument('--distill_G_loss_type', type=str, default='mse', choices=['mse', 'ka'])
        parser.add_argument('--lambda_distill', type=float, default=1)
        parser.add_argument('--lambda_recon', type=float, default=100)
        parser.add_argument('--lambda_gan', type=float, default=1)
        parser.add_argument('--teacher_dropout_rate', type=float, default=0)
        parser.add_argument('--student_dropout_rate', type=float, default=0)
        parser.add_argument('--restore_pretrained_G_path', type=str, default=None)
        parser.add_argument('--pretrained_netG', type=str, default='inception_9blocks', choices=['inception_9blocks'])
        parser.add_argument('--pretrained_ngf', type=int, default=64)
        parser.add_argument('--target_flops', type=float, default=0)
        parser.add_argument('--prune_cin_lb', type=int, default=0)
        parser.add_argument('--pretrained_student_G_path', type=str, default=None)
        parser.add_argument('--prune_only', action='store_true')
        parser.add_argument('--prune_continue', action='store_true')
        parser.add_argument('--prune_logging_verbose', action='store_true')
        parser.set_defaults(norm='instance', dataset_mode='aligned', log_dir='logs/inception', teacher_netG='inception_9blocks',
                            student_netG='inception_9blocks')
        return parser

    def __init__(self, opt):
        assert opt.isTrain
        self.opt = opt
        self.gpu_ids = list(getattr(opt, 'gpu_ids', [0])) or [0]
        # raises without an sm_100 device and libcatb200.so: there is no CPU path.  (Only the kernel emulation of the test
        # suite patches this check out; it then runs the host logic on CPU tensors.)
        ops.require_cuda()
        self.device = torch.device('cuda:%d' % self.gpu_ids[0]) if torch.cuda.is_available() else torch.device('cpu')
        self.save_dir = os.path.join(getattr(opt, 'log_dir', '.'), 'checkpoints')
        if getattr(opt, 'distill_G_loss_type', 'ka') not in ('ka', 'mse'):
            raise NotImplementedError('--distill_G_loss_type [%s]: ka | mse' % opt.distill_G_loss_type)
        if opt.recon_loss_type == 'vgg':
            raise NotImplementedError('VGG reconstruction loss is not on the inception distillation scripts')
        if getattr(opt, 'teacher_dropout_rate', 0) or getattr(opt, 'student_dropout_rate', 0):
            raise NotImplementedError('dropout (the CAT scripts run with rate 0)')
        if getattr(opt, 'restore_pretrained_G_path', None) or getattr(opt, 'pretrained_student_G_path', None):
            raise NotImplementedError('weight transfer from a pretrained generator (utils/weight_transfer.py) is outside the '
                                      'hot path: initialise the student with the reference and pass --restore_student_G_path')
        if getattr(opt, 'world_size', None) is None:      # one process per GPU under torchrun (cat_b200.install)
            opt.world_size = int(os.environ.get('WORLD_SIZE', '1'))
        self.loss_names = ['G_gan', 'G_distill', 'G_recon', 'D_fake', 'D_real'] + ['G_distill%d' % i for i in range(4)]
        self.model_names = ['netG_student', 'netG_teacher', 'netD']
        self.visual_names = ['real_A', 'Sfake_B', 'Tfake_B', 'real_B']
        self.image_paths = []
        ids = self.gpu_ids[:1] if self.device.type == 'cuda' else []
        self.netG_teacher = networks.define_G(opt.input_nc, opt.output_nc, opt.teacher_ngf, opt.teacher_netG, opt.norm,
                                              0, opt.init_type, opt.init_gain, ids, opt=opt)
        arch_S = getattr(opt, 'student_arch', None)  # pruned architecture (what shrink_model produces)
        if arch_S is not None:
            self.netG_student = networks.init_net(networks.InceptionGenerator.from_arch(arch_S), opt.init_type,
                                                  opt.init_gain, ids)
        else:
            self.netG_student = networks.define_G(opt.input_nc, opt.output_nc, opt.student_ngf, opt.student_netG,
                                                  opt.norm, 0, opt.init_type, opt.init_gain, ids, opt=opt)
        d_in = opt.input_nc + opt.output_nc if opt.dataset_mode in ('aligned', 'cityscapes') else opt.output_nc
        self.netD = networks.define_D(d_in, opt.ndf, opt.netD, opt.n_layers_D, opt.norm, opt.init_type, opt.init_gain,
                                      ids, opt=opt)
        self.netG_teacher.eval()
        self.mapping_layers = list(MAPPING_LAYERS)
        # adaptor convs (base_inception_distiller.py:195-202): parameters of optimizer_G, used by the 'mse' loss only; their
        # input width follows the (pruned) student like utils/common.py:154-161
        c_s = self.netG_student.arch()['widths'][2]
        self.netAs = [nn.Conv2d(c_s, opt.teacher_ngf * 4, kernel_size=1).to(self.device) for _ in range(4)]
        # optimizer_G: group 0 = the student, group 1 = the adaptor convs (base_inception_distiller.py:205-214)
        self.optimizer_G = ArenaAdam(opt.lr, (opt.beta1, 0.999), n_groups=2)
        self.optimizer_D = ArenaAdam(opt.lr, (opt.beta1, 0.999))
        self.optimizers = [self.optimizer_G, self.optimizer_D]
        self.Tacts, self.Sacts = {}, {}
        self.engine = None
        self.is_best = False
        self._epoch = 0
        self.eval_dataloader = []          # set by the caller (data/ is outside the hot path)
        self.metric_fns = {}               # {'fid': f(fakes) -> float, 'mIoU': f(fakes, names) -> float}
        self.metrics = MetricBook()

    # ---- protocol -------------------------------------------------------------------------------
    def setup(self, opt, verbose=True):
        self.load_networks(verbose)
        # The reference profiles both generators here (inception_distiller.py:85-97) and model_profiling leaves them in
        # eval() (utils/model_profiling.py:299): the student only returns to train() at the end of the first
        # evaluate_model (:280), so the FIRST optimize_parameters of a run uses the BatchNorm running statistics.
        self.netG_student.eval()
        if verbose:
            self.print_networks()

    def _hp(self):
        o = self.opt
        return dict(gan_mode=o.gan_mode, aligned=o.dataset_mode in ('aligned', 'cityscapes'), lambda_recon=o.lambda_recon,
                    lambda_gan=o.lambda_gan, lambda_distill=o.lambda_distill, lr=o.lr, beta1=o.beta1,
                    student_training=self.netG_student.training, recon_loss_type=o.recon_loss_type,
                    ka_scale=float(getattr(o, 'world_size', 1)), distill_loss_type=getattr(o, 'distill_G_loss_type', 'ka'))

    def _make_engine(self, B, H, W):
        return DistillStep(self.netG_teacher.arch(), self.netG_student.arch(), self.netD.arch(), self._hp(), B, H, W,
                           device=str(self.device), world_size=int(getattr(self.opt, 'world_size', 1)),
                           use_cuda_graph=bool(getattr(self.opt, 'cuda_graph', True)))

    def _bind_engine(self, eng):
        for module, net in ((self.netG_teacher, eng.T), (self.netG_student, eng.S), (self.netD, eng.D)):
            module.bind(net)               # copies the module's weights in, then re-points them at the arena
            net.pack_weights()
        if eng.A is not None:              # 'mse': the adaptor modules alias the engine's adaptor arena
            eng.A.load_state_dicts([net.state_dict() for net in self.netAs])
            for i, net in enumerate(self.netAs):
                net.weight.data = eng.A.arena.view('%d.weight' % i)
                net.bias.data = eng.A.arena.view('%d.bias' % i)
        a_params = [p for net in self.netAs for p in net.parameters()]
        self.optimizer_G.bind([[(self.netG_student.parameters(), eng.S.arena, eng.step_G)],
                               [(a_params, eng.A.arena if eng.A is not None else None, eng.step_A if eng.A is not None else None)]])
        self.optimizer_D.bind([[(self.netD.parameters(), eng.D.arena, eng.step_D)]])

    def set_input(self, input):
        AtoB = getattr(self.opt, 'direction', 'AtoB') == 'AtoB'
        self.real_A = input['A' if AtoB else 'B']
        self.real_B = input['B' if AtoB else 'A']
        self.image_paths = input.get('A_paths' if AtoB else 'B_paths', [])
        B, _, H, W = self.real_A.shape
        self._ensure_engine(B, H, W)
        self.engine.set_input(self.real_A, self.real_B)

    def optimize_parameters(self, steps):
        self.engine.set_student_training(self.netG_student.training)     # follows netG_student.train() / .eval()
        self.engine.step()
        self._losses = None

    def forward(self, teacher_forward=True):
        """Inference of both generators on the current input (test() in the reference)."""
        with torch.no_grad():
            if teacher_forward:
                self.Tfake_B = self.netG_teacher(self.engine.real_A)
            self.Sfake_B = self.netG_student(self.engine.real_A)

    def test(self, teacher_forward=True):
        self.forward(teacher_forward)

    def get_current_losses(self):
        L = self.engine.get_losses()   # one device synchronisation, like float(loss) in base_model.py:187
        out = OrderedDict()
        for name in self.loss_names:
            key = ('Specific_loss/' if any(ch.isdigit() for ch in name) else ('D_loss/' if name.startswith('D_') else 'G_loss/')) + name
            out[key] = L[name]
            setattr(self, 'loss_' + name, L[name])
        return out

    def update_learning_rate(self, logger=None):
        """'linear' policy of models/networks.py:80-87, stepped once per epoch (trainer.py:175)."""
        o = self.opt
        self._epoch += 1
        scale = 1.0 - max(0, self._epoch + 1 - o.nepochs) / float(o.nepochs_decay + 1)
        lr = o.lr * scale
        for opt_ in self.optimizers:
            for pg in opt_.param_groups:
                pg['lr'] = lr
        if self.engine is not None:
            self.engine.set_lr(lr)
        msg = 'learning rate = %.7f' % lr
        logger.print_info(msg + '\n') if logger is not None else print(msg)

    def evaluate_model(self, step, save_image=False):
        """inception_distiller.py:204-281: student (and teacher) inference over the evaluation set in eval mode, metric
        bookkeeping, and the student back in train() -- the module forward compiles its own inference network per batch
        shape against the same parameter arena, so the training engine and its optimiser state are untouched."""
        self.is_best = False
        self.netG_student.eval()
        AtoB = getattr(self.opt, 'direction', 'AtoB') == 'AtoB'
        fakes, names = [], []
        for data_i in self.eval_dataloader:
            real_A = data_i['A' if AtoB else 'B'].to(self.device)
            with torch.no_grad():
                self.Tfake_B = self.netG_teacher(real_A)
                self.Sfake_B = self.netG_student(real_A)
            fakes.append(self.Sfake_B.cpu())
            names += image_names(data_i.get('A_paths' if AtoB else 'B_paths', []))
        ret, self.is_best = self.metrics.update(self.metric_fns, fakes, names)
        self.netG_student.train()
        return ret

    def add_mapping_hook(self):
        pass   # the engine exposes the four mapped activations natively (engine.S.acts / engine.T.acts)

    def remove_mapping_hook(self):
        pass

    def print_networks(self):
        for name in self.model_names:
            net = getattr(self, name)
            n = sum(p.numel() for p in net.parameters())
            print('[Network %s] Total number of parameters : %.3f M' % (name, n / 1e6))

    # ---- checkpoints (same file names and key layout as base_inception_distiller.py:342-396) -------
    def load_networks(self, verbose=True, teacher_only=False, restore_pretrain=True):
        def load(net, path):
            if path is not None:
                net.load_state_dict(torch.load(path, map_location='cpu'))
                if verbose:
                    print('Load network at %s' % path)
        load(self.netG_teacher, getattr(self.opt, 'restore_teacher_G_path', None))
        load(self.netG_student, getattr(self.opt, 'restore_student_G_path', None))
        load(self.netD, getattr(self.opt, 'restore_D_path', None))
        if getattr(self.opt, 'restore_A_path', None) is not None:      # base_inception_distiller.py:356-359
            for i, netA in enumerate(self.netAs):
                load(netA, '%s-%d.pth' % (self.opt.restore_A_path, i))
        if getattr(self.opt, 'restore_O_path', None) is not None:      # :360-365 (applied when the engine is compiled)
            for i, optimizer in enumerate(self.optimizers):
                optimizer.load_state_dict(torch.load('%s-%d.pth' % (self.opt.restore_O_path, i), map_location='cpu',
                                                     weights_only=False))
                for param_group in optimizer.param_groups:
                    param_group['lr'] = self.opt.lr

    def save_networks(self, epoch):
        os.makedirs(self.save_dir, exist_ok=True)
        def cpu_sd(net):
            return OrderedDict((k, v.detach().cpu().clone()) for k, v in net.state_dict().items())
        torch.save(cpu_sd(self.netG_student), os.path.join(self.save_dir, '%s_net_G.pth' % epoch))
        torch.save(cpu_sd(self.netD), os.path.join(self.save_dir, '%s_net_D.pth' % epoch))
        for i, net in enumerate(self.netAs):
            torch.save(cpu_sd(net), os.path.join(self.save_dir, '%s_net_A-%d.pth' % (epoch, i)))
        for i, optimizer in enumerate(self.optimizers):
            torch.save(optimizer.state_dict(), os.path.join(self.save_dir, '%s_optim-%d.pth' % (epoch, i)))
