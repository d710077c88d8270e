"""Mirror of the reference Pix2PixModel (models/pix2pix_model.py:19-212): teacher training on aligned pairs, one
``optimize_parameters`` = cat_b200.train_engine.Pix2PixTrainStep.step()."""
import torch

from ..train_engine import Pix2PixTrainStep
from . import networks
from .base_model import ArenaOptimizer, BaseModel


class Pix2PixModel(BaseModel):
    @staticmethod
    def modify_commandline_options(parser, is_train=True):
        """The flags of pix2pix_model.py:21-66 that the step uses."""
        assert is_train
        parser.add_argument('--restore_G_path', type=str, default=None)
        parser.add_argument('--restore_D_path', type=str, default=None)
        parser.add_argument('--recon_loss_type', type=str, default='l1', choices=['l1', 'l2', 'smooth_l1'])
        parser.add_argument('--lambda_recon', type=float, default=100)
        parser.add_argument('--lambda_gan', type=float, default=1)
        parser.add_argument('--lambda_comp_cost', type=float, default=0)
        return parser

    def __init__(self, opt):
        super().__init__(opt)
        if getattr(opt, 'lambda_comp_cost', 0) > 0:
            raise NotImplementedError('--lambda_comp_cost > 0 is not used by the CAT training scripts')
        self.loss_names = ['G_gan', 'G_recon', 'D_real', 'D_fake']
        self.visual_names = ['real_A', 'fake_B', 'real_B']
        self.model_names = ['G', 'D']
        ids = self._ids
        self.netG = networks.define_G(opt.input_nc, opt.output_nc, opt.ngf, opt.netG, opt.norm, opt.dropout_rate,
                                      opt.init_type, opt.init_gain, ids, opt=opt)
        self.netD = networks.define_D(opt.input_nc + opt.output_nc, opt.ndf, opt.netD, opt.n_layers_D, opt.norm,
                                      opt.init_type, opt.init_gain, ids, opt=opt)
        self.optimizer_G = ArenaOptimizer(opt.lr, (opt.beta1, 0.999))
        self.optimizer_D = ArenaOptimizer(opt.lr, (opt.beta1, 0.999))
        self.optimizers = [self.optimizer_G, self.optimizer_D]

    def _make_engine(self, B, H, W):
        o = self.opt
        hp = dict(gan_mode=o.gan_mode, lambda_recon=o.lambda_recon, lambda_gan=o.lambda_gan, lr=o.lr, beta1=o.beta1,
                  recon_loss_type=o.recon_loss_type)
        eng = Pix2PixTrainStep(self.netG.arch(), self.netD.arch(), hp, B, H, W, device=str(self.device),
                               world_size=int(getattr(o, 'world_size', 1)), use_cuda_graph=bool(getattr(o, 'cuda_graph', True)))
        return eng

    def _bind_engine(self, eng):
        for module, net in ((self.netG, eng.G), (self.netD, eng.D)):
            module.bind(net)               # copies the module's weights in, then re-points them at the arena
            net.pack_weights()
        self.optimizer_G.bind([[(self.netG.parameters(), eng.G.arena, eng.step_G)]])
        self.optimizer_D.bind([[(self.netD.parameters(), eng.D.arena, eng.step_D)]])

    def set_input(self, input):
        AtoB = getattr(self.opt, 'direction', 'AtoB') == 'AtoB'
        self.real_A = input['A' if AtoB else 'B']
        self.real_B = input['B' if AtoB else 'A']
        self.image_paths = input.get('A_paths' if AtoB else 'B_paths', [])
        B, _, H, W = self.real_A.shape
        self._ensure_engine(B, H, W)
        self.engine.set_input(self.real_A, self.real_B)

    def forward(self):
        with torch.no_grad():
            self.fake_B = self.netG(self.engine.real_A)

    def _eval_batch(self, data):
        """models/pix2pix_model.py:222-226: fake_B for one evaluation batch (module forward = inference engine)."""
        AtoB = getattr(self.opt, 'direction', 'AtoB') == 'AtoB'
        with torch.no_grad():
            self.fake_B = self.netG(data['A' if AtoB else 'B'].to(self.device))
        return self.fake_B, data.get('A_paths' if AtoB else 'B_paths', [])
