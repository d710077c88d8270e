"""Mirror of the reference CycleGANModel (models/cycle_gan_model.py:20-303): teacher training on unpaired domains,
one ``optimize_parameters`` = cat_b200.train_engine.CycleGANTrainStep.step().  The history buffers
(``fake_A_pool`` / ``fake_B_pool``, utils/image_pool.py) live on the device inside the engine and draw their decisions
from Python's global ``random`` like the reference."""
import torch

from ..train_engine import CycleGANTrainStep
from . import networks
from .base_model import ArenaOptimizer, BaseModel, MetricBook, image_names


class CycleGANModel(BaseModel):
    @staticmethod
    def modify_commandline_options(parser, is_train=True):
        """The flags of cycle_gan_model.py:32-104 that the step uses."""
        assert is_train
        for n in ('G_A', 'G_B', 'D_A', 'D_B'):
            parser.add_argument('--restore_%s_path' % n, type=str, default=None)
        parser.add_argument('--lambda_A', type=float, default=10.0)
        parser.add_argument('--lambda_B', type=float, default=10.0)
        parser.add_argument('--lambda_identity', type=float, default=0.5)
        parser.set_defaults(norm='instance', dataset_mode='unaligned', batch_size=1, ndf=64, gan_mode='lsgan')
        return parser

    def __init__(self, opt):
        super().__init__(opt)
        assert getattr(opt, 'direction', 'AtoB') == 'AtoB' and opt.dataset_mode == 'unaligned'
        if opt.lambda_identity > 0.0:
            assert opt.input_nc == opt.output_nc
        self.loss_names = ['D_A', 'G_A', 'G_cycle_A', 'G_idt_A', 'D_B', 'G_B', 'G_cycle_B', 'G_idt_B']
        self.visual_names = ['real_A', 'fake_B', 'rec_A', 'real_B', 'fake_A', 'rec_B']
        self.model_names = ['G_A', 'G_B', 'D_A', 'D_B']
        ids = self._ids
        mk_G = lambda cin, cout: networks.define_G(cin, cout, opt.ngf, opt.netG, opt.norm, opt.dropout_rate, opt.init_type,
                                                   opt.init_gain, ids, opt=opt)
        mk_D = lambda cin: networks.define_D(cin, opt.ndf, opt.netD, opt.n_layers_D, opt.norm, opt.init_type, opt.init_gain,
                                             ids, opt=opt)
        self.netG_A, self.netG_B = mk_G(opt.input_nc, opt.output_nc), mk_G(opt.output_nc, opt.input_nc)
        self.netD_A, self.netD_B = mk_D(opt.output_nc), mk_D(opt.input_nc)
        self.optimizer_G = ArenaOptimizer(opt.lr, (opt.beta1, 0.999))
        self.optimizer_D = ArenaOptimizer(opt.lr, (opt.beta1, 0.999))
        self.optimizers = [self.optimizer_G, self.optimizer_D]

    def _make_engine(self, B, H, W):
        o = self.opt
        hp = dict(gan_mode=o.gan_mode, lambda_A=o.lambda_A, lambda_B=o.lambda_B, lambda_identity=o.lambda_identity, lr=o.lr,
                  beta1=o.beta1, pool_size=int(getattr(o, 'pool_size', 50)))
        eng = CycleGANTrainStep(self.netG_A.arch(), self.netD_A.arch(), hp, B, H, W, device=str(self.device),
                                world_size=int(getattr(o, 'world_size', 1)), use_cuda_graph=bool(getattr(o, 'cuda_graph', True)))
        return eng

    def _bind_engine(self, eng):
        for module, net in ((self.netG_A, eng.G_A), (self.netG_B, eng.G_B), (self.netD_A, eng.D_A), (self.netD_B, eng.D_B)):
            module.bind(net)               # copies the module's weights in, then re-points them at the arena
        eng._pack_generators()             # every application of a generator packs its own GEMM images from the arena
        eng.D_A.pack_weights()
        eng.D_B.pack_weights()
        # cycle_gan_model.py:165-174: one Adam over chain(netG_A, netG_B) / chain(netD_A, netD_B)
        self.optimizer_G.bind([[(self.netG_A.parameters(), eng.G_A.arena, eng.step_GA),
                                (self.netG_B.parameters(), eng.G_B.arena, eng.step_GB)]])
        self.optimizer_D.bind([[(self.netD_A.parameters(), eng.D_A.arena, eng.step_DA),
                                (self.netD_B.parameters(), eng.D_B.arena, eng.step_DB)]])

    def set_input(self, input):
        self.real_A, self.real_B = input['A'], input['B']
        self.image_paths = input.get('A_paths', [])
        B, _, H, W = self.real_A.shape
        self._ensure_engine(B, H, W)
        self.engine.set_input(self.real_A, self.real_B)

    def forward(self):
        """cycle_gan_model.py:221-226, as inference through the module mirrors (the training step has its own)."""
        with torch.no_grad():
            self.fake_B = self.netG_A(self.engine.real_A)
            self.rec_A = self.netG_B(self.fake_B)
            self.fake_A = self.netG_B(self.engine.real_B)
            self.rec_B = self.netG_A(self.fake_A)

    def evaluate_model(self, step, save_image=False):
        """models/cycle_gan_model.py:310-365: both directions (`eval_dataloader_AtoB` -> G_A, `eval_dataloader_BtoA` -> G_B,
        each an iterable of {'A', 'A_paths'} dicts set by the caller), FID per direction through `metric_fns_A` / `_B`."""
        if not hasattr(self, 'metrics_B'):
            self.metrics_A, self.metrics_B = MetricBook(), MetricBook()
        self.is_best_A = self.is_best_B = False
        self.eval()
        ret = {}
        for side, net in (('A', self.netG_A), ('B', self.netG_B)):
            loader = getattr(self, 'eval_dataloader_AtoB' if side == 'A' else 'eval_dataloader_BtoA', [])
            fakes, names = [], []
            for data_i in loader:
                with torch.no_grad():
                    fakes.append(net(data_i['A'].to(self.device)).cpu())
                names += image_names(data_i.get('A_paths', []))
            r, best = getattr(self, 'metrics_' + side).update(getattr(self, 'metric_fns_' + side, {}), fakes, names, '_' + side)
            setattr(self, 'is_best_' + side, best)
            ret.update(r)
        self.is_best = self.is_best_A or self.is_best_B
        self.train()
        return ret
