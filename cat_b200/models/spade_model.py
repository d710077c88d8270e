"""Mirror of the reference SPADEModel (models/spade_model.py:20-215, models/modules/spade_modules/
spade_model_modules.py:13-175): GauGAN teacher training, one ``optimize_parameters`` =
cat_b200.train_engine.SpadeTrainStep.step().  ``modify_commandline_options`` sets ``active_fn='nn.LeakyReLU'`` like the
reference (spade_model.py:92), so teachers trained here have the reference's activations.
The VGG19 weights come from ``opt.vgg_state_dict`` (torchvision ``vgg19().features`` keys): the pretrained checkpoint
the reference downloads (models/modules/loss.py:154) has to be supplied by the caller."""
import torch
from torch import nn

from ..train_engine import SpadeTrainStep
from . import networks
from .base_model import ArenaOptimizer, BaseModel


def input_semantics(data, n_label, semantic_nc, device):
    """preprocess_input / get_edges (models/spade_model.py:142-179) for one dataset dict -> NCHW fp32 one-hot + edge map."""
    from .. import ops
    B, _, H, W = data['label'].shape
    label = data['label'].reshape(B, H, W).to(device=device, dtype=torch.int32).contiguous()
    inst = data['instance'].reshape(B, H, W).to(device=device, dtype=torch.int32).contiguous()
    seg = ops.Act.empty(B, H, W, semantic_nc, device, zero=True)
    ops.onehot_edges(label, inst, n_label, seg)
    return ops.nhwc_to_nchw(seg, semantic_nc)


class SPADEModelModules(nn.Module):
    """modules_on_one_gpu: the container of spade_model_modules.py:13-51."""

    def __init__(self, opt, ids):
        super().__init__()
        self.opt = opt
        self.netG = networks.define_G(opt.input_nc, opt.output_nc, opt.ngf, opt.netG, opt.norm, opt.dropout_rate,
                                      opt.init_type, opt.init_gain, ids, opt=opt)
        self.netD = networks.define_D(opt.input_nc + opt.output_nc, opt.ndf, opt.netD, opt.n_layers_D, opt.norm,
                                      opt.init_type, opt.init_gain, ids, opt=opt)


class SPADEModel(BaseModel):
    @staticmethod
    def modify_commandline_options(parser, is_train=True):
        """The flags of spade_model.py:22-95 that the step uses."""
        assert is_train
        parser.set_defaults(netG='inception_spade')
        parser.add_argument('--norm_G', type=str, default='spadesyncbatch3x3')
        parser.add_argument('--num_upsampling_layers', choices=('normal', 'more', 'most'), default='more')
        parser.add_argument('--restore_G_path', type=str, default=None)
        parser.add_argument('--restore_D_path', type=str, default=None)
        parser.add_argument('--lambda_gan', type=float, default=1)
        parser.add_argument('--lambda_feat', type=float, default=10)
        parser.add_argument('--lambda_vgg', type=float, default=10)
        parser.add_argument('--beta2', type=float, default=0.999)
        parser.add_argument('--no_TTUR', action='store_true')
        parser.add_argument('--num_D', type=int, default=2)
        parser.add_argument('--norm_D', type=str, default='spectralinstance')
        parser.set_defaults(netD='multi_scale', ndf=64, dataset_mode='cityscapes', batch_size=16, init_type='xavier',
                            n_layers_D=4, active_fn='nn.LeakyReLU')
        return parser

    def __init__(self, opt):
        super().__init__(opt)
        if getattr(opt, 'gan_mode', 'hinge') != 'hinge':
            raise NotImplementedError('SPADEModel uses the hinge GAN loss (the CAT GauGAN scripts)')
        self.model_names = ['G', 'D']
        self.visual_names = ['labels', 'fake_B', 'real_B']
        self.loss_names = ['G_gan', 'G_feat', 'G_vgg', 'D_real', 'D_fake']
        self.modules = self.modules_on_one_gpu = SPADEModelModules(opt, self._ids).to(self.device)
        if getattr(opt, 'no_TTUR', False):
            self.betas, self.lr_G, self.lr_D = (opt.beta1, opt.beta2), opt.lr, opt.lr
        else:   # spade_model_modules.py:53-66
            self.betas, self.lr_G, self.lr_D = (0.0, 0.9), opt.lr / 2, opt.lr * 2
        self.optimizer_G = ArenaOptimizer(self.lr_G, self.betas)
        self.optimizer_D = ArenaOptimizer(self.lr_D, self.betas)
        self.optimizers = [self.optimizer_G, self.optimizer_D]

    def _net(self, name):
        return getattr(self.modules_on_one_gpu, 'net' + name)

    def _base_lrs(self):
        return [self.lr_G, self.lr_D]

    def _make_engine(self, B, H, W):
        o, mm = self.opt, self.modules_on_one_gpu
        hp = dict(lambda_gan=o.lambda_gan, lambda_feat=o.lambda_feat, lambda_vgg=o.lambda_vgg, lr_G=self.lr_G, lr_D=self.lr_D,
                  beta1=self.betas[0], beta2=self.betas[1], n_label=int(o.input_nc))
        eng = SpadeTrainStep(mm.netG.arch(), mm.netD.arch(), hp, B, H, W, device=str(self.device),
                             world_size=int(getattr(o, 'world_size', 1)), use_cuda_graph=bool(getattr(o, 'cuda_graph', True)))
        return eng

    def _bind_engine(self, eng):
        o, mm = self.opt, self.modules_on_one_gpu
        mm.netG.bind(eng.G)                # copies the module's weights in, then re-points them at the arena
        mm.netD._alias_into(eng.D)
        vgg = getattr(o, 'vgg_state_dict', None)
        if vgg is None:
            raise RuntimeError('opt.vgg_state_dict (torchvision vgg19().features state_dict) is required: the pretrained '
                               'VGG19 of models/modules/loss.py:154 cannot be downloaded here')
        eng.V.load_state_dict(vgg)
        self.optimizer_G.bind([[(mm.netG.parameters(), eng.G.arena, eng.step_G)]])
        self.optimizer_D.bind([[(mm.netD.parameters(), eng.D.arena, eng.step_D)]])

    def set_input(self, input):
        """spade_model.py:132-136 (the one-hot / edge preprocessing itself runs inside the step)."""
        self.data = input
        self.image_paths = input.get('path', [])
        self.labels = input['label']
        B, _, H, W = input['image'].shape
        self._ensure_engine(B, H, W)
        self.engine.set_input(input['label'], input['instance'], input['image'])

    def forward(self, on_one_gpu=False):
        """generate_fake (spade_model_modules.py:68-71) for the current input."""
        from .. import ops
        eng = self.engine
        eng._preprocess()
        self.input_semantics = ops.nhwc_to_nchw(eng.seg, eng.snc)
        self.real_B = eng.image
        self.fake_B = ops.nhwc_to_nchw(eng.G.forward(), 3)

    def test(self):
        self.forward(on_one_gpu=True)

    def eval(self):
        self.modules_on_one_gpu.netG.eval()       # evaluate_model (spade_model.py:221) switches the generator only

    def train(self):
        self.modules_on_one_gpu.netG.train()

    def _eval_batch(self, data):
        o = self.opt
        with torch.no_grad():
            self.fake_B = self.modules_on_one_gpu.netG(input_semantics(data, int(o.input_nc), int(o.semantic_nc), self.device))
        return self.fake_B, data.get('path', [])
