"""Mirror of the reference SPADE network classes: InceptionSPADEGenerator (inception_spade_generator.py:15-140),
SPADEInvertedResidualChannels / InceptionSPADE / ConvSyncBNReLU / Conv (inception_modules.py:280-762),
MultiscaleDiscriminator / SPADENLayerDiscriminator (discriminators.py:129-226).

Like cat_b200.models.networks these are parameter *containers* with the reference's module tree, attribute names
(`res_channels`, `dw_channels`, `input_dim`, `output_dim`, `_build()`, `get_named_block_list()`, ...) and
`state_dict` keys, which pruning, weight transfer and checkpoints index directly.  Computation is done
by the compiled engine networks of cat_b200.spade_engine; `bind()` re-points every parameter / buffer at the
engine's flat arenas.
"""
import collections
import functools
import re

from torch import nn

from .networks import BaseNetwork, _EngineBacked, get_active_fn, _pre


class SynchronizedBatchNorm2d(nn.BatchNorm2d):
    """sync_batchnorm/batchnorm.py in its single-replica form (one process per GPU: statistics are per rank)."""


class ConvSyncBNReLU(nn.Module):
    """inception_modules.py:280-316."""

    def __init__(self, in_planes, out_planes, kernel_size=3, stride=1, groups=1, use_bias=True, norm_layer=None,
                 active_fn=None, spectral_norm=False, spade=False):
        super().__init__()
        if spectral_norm or spade:
            raise NotImplementedError('spectral / SPADE-normalised generator convs are not used by the CAT SPADE scripts '
                                      '(norm_G = spadesyncbatch3x3)')
        self.conv = nn.Conv2d(in_planes, out_planes, kernel_size, stride, (kernel_size - 1) // 2, groups=groups, bias=use_bias)
        self.norm = norm_layer(out_planes)
        self.active = active_fn()


class Conv(nn.Module):
    """inception_modules.py:319-342."""

    def __init__(self, in_planes, out_planes, kernel_size=3, stride=1, groups=1, use_bias=True, spectral_norm=False):
        super().__init__()
        if spectral_norm:
            raise NotImplementedError('spectral-normalised generator convs are not used by the CAT SPADE scripts')
        self.conv = nn.Conv2d(in_planes, out_planes, kernel_size, stride, (kernel_size - 1) // 2, groups=groups, bias=use_bias)


def _widths(channels, base, factor, ks):
    if channels is None:
        return [base // factor for _ in ks]
    if isinstance(channels, int):
        return [channels // factor for _ in ks]
    assert len(channels) == len(ks)
    return [c // factor for c in channels]


class _SixBranchBody(nn.Module):
    """Accessors shared by the two block classes (inception_modules.py:486-546, 726-744)."""

    def get_named_first_res_bn(self, prefix=None):
        return collections.OrderedDict((_pre(f'res_ops.{i}.0.norm', prefix), op[0].norm) for i, op in enumerate(self.res_ops))

    def get_named_first_dw_bn(self, prefix=None):
        return collections.OrderedDict((_pre(f'dw_ops.{i}.0.norm', prefix), op[0].norm) for i, op in enumerate(self.dw_ops))

    def get_named_first_bn(self, prefix=None):
        return collections.OrderedDict(list(self.get_named_first_res_bn().items()) + list(self.get_named_first_dw_bn().items()))

    def get_first_res_bn(self):
        return list(self.get_named_first_res_bn().values())

    def get_first_dw_bn(self):
        return list(self.get_named_first_dw_bn().values())

    def get_first_bn(self):
        return self.get_first_res_bn() + self.get_first_dw_bn()

    def forward(self, *a):
        raise RuntimeError('blocks are executed by the compiled generator (InceptionSPADEGenerator.forward), not one by one')


class InceptionSPADE(_SixBranchBody):
    """inception_modules.py:565-762: gamma / beta from the label map through six branches."""

    def __init__(self, norm, norm_nc, label_nc, nhidden=128, opt=None):
        super().__init__()
        ks = [opt.kernel_sizes] if isinstance(opt.kernel_sizes, int) else list(opt.kernel_sizes)
        self.norm_layer = functools.partial(SynchronizedBatchNorm2d, affine=True)
        self.active_fn = functools.partial(nn.ReLU, inplace=True)
        self.param_free_norm_layer = norm
        self.input_dim, self.output_dim = label_nc, norm_nc
        self.res_channels = _widths(opt.channels, nhidden, opt.channels_reduction_factor, ks)
        self.dw_channels = _widths(opt.channels, nhidden, opt.channels_reduction_factor, ks)
        self.res_kernel_sizes, self.dw_kernel_sizes = ks, ks
        self.param_free_norm, self.res_ops, self.dw_ops = self._build()

    def _build(self):
        param_free_norm = self.param_free_norm_layer(self.output_dim, affine=False)
        res_ops = nn.ModuleList()
        for midp, k in zip(self.res_channels, self.res_kernel_sizes):
            if midp == 0:
                continue
            res_ops.append(nn.Sequential(
                ConvSyncBNReLU(self.input_dim, midp, kernel_size=k, norm_layer=self.norm_layer, active_fn=self.active_fn),
                nn.Conv2d(midp, 2 * self.output_dim, kernel_size=k, padding=(k - 1) // 2)))
        dw_ops = nn.ModuleList()
        for midp, k in zip(self.dw_channels, self.dw_kernel_sizes):
            if midp == 0:
                continue
            dw_ops.append(nn.Sequential(
                ConvSyncBNReLU(self.input_dim, midp, kernel_size=1, norm_layer=self.norm_layer, active_fn=self.active_fn),
                ConvSyncBNReLU(midp, midp, kernel_size=k, groups=midp, norm_layer=self.norm_layer, active_fn=self.active_fn),
                nn.Conv2d(midp, 2 * self.output_dim, kernel_size=1)))
        return param_free_norm, res_ops, dw_ops


class SPADEInvertedResidualChannels(_SixBranchBody):
    """inception_modules.py:345-562."""

    def __init__(self, fin, fout, opt):
        super().__init__()
        self.opt = opt
        self.learned_shortcut = fin != fout
        fmiddle = min(fin, fout)
        ks = [opt.kernel_sizes] if isinstance(opt.kernel_sizes, int) else list(opt.kernel_sizes)
        self.input_dim, self.output_dim = fin, fout
        self.res_channels = _widths(opt.channels, fmiddle, opt.channels_reduction_factor, ks)
        self.dw_channels = _widths(opt.channels, fmiddle, opt.channels_reduction_factor, ks)
        self.res_kernel_sizes, self.dw_kernel_sizes = ks, ks
        self.active_fn = get_active_fn(opt.active_fn)
        self.active = self.active_fn()
        if 'spectral' in opt.norm_G:
            raise NotImplementedError('norm_G with spectral norm: the CAT SPADE scripts use spadesyncbatch3x3')
        parsed = re.search(r'spade(\D+)(\d)x\d', opt.norm_G)
        if parsed is None or parsed.group(1) not in ('syncbatch', 'batch'):
            raise NotImplementedError('cat_b200 implements the (sync)batch parameter-free norm of the CAT scripts (got %s)' % opt.norm_G)
        self.norm_layer = SynchronizedBatchNorm2d if parsed.group(1) == 'syncbatch' else nn.BatchNorm2d
        self.semantic_nc = opt.semantic_nc
        self.res_ops, self.dw_ops, self.shortcut, self.spade = self._build()

    def _build(self, build_only=False):
        aff = functools.partial(self.norm_layer, affine=True)
        res_ops = nn.ModuleList()
        for midp, k in zip(self.res_channels, self.res_kernel_sizes):
            if midp == 0:
                continue
            res_ops.append(nn.Sequential(
                ConvSyncBNReLU(self.input_dim, midp, kernel_size=k, norm_layer=aff, active_fn=self.active_fn),
                Conv(midp, self.output_dim, kernel_size=k)))
        dw_ops = nn.ModuleList()
        for midp, k in zip(self.dw_channels, self.dw_kernel_sizes):
            if midp == 0:
                continue
            dw_ops.append(nn.Sequential(
                ConvSyncBNReLU(self.input_dim, midp, kernel_size=1, norm_layer=aff, active_fn=self.active_fn),
                ConvSyncBNReLU(midp, midp, kernel_size=k, groups=midp, norm_layer=functools.partial(self.norm_layer, affine=False),
                               active_fn=self.active_fn),
                Conv(midp, self.output_dim, kernel_size=1)))
        shortcut = nn.Sequential(self.norm_layer(self.input_dim, affine=True),
                                 Conv(self.input_dim, self.output_dim, kernel_size=1, use_bias=False)) if self.learned_shortcut else None
        if build_only:
            self.spade.param_free_norm, self.spade.res_ops, self.spade.dw_ops = self.spade._build()
            spade = self.spade
        else:
            spade = InceptionSPADE(norm=self.norm_layer, norm_nc=self.input_dim, label_nc=self.semantic_nc, opt=self.opt)
        return res_ops, dw_ops, shortcut, spade


BLOCKS = ['head_0', 'G_middle_0', 'G_middle_1', 'up_0', 'up_1', 'up_2', 'up_3']


class InceptionSPADEGenerator(BaseNetwork, _EngineBacked):
    """inception_spade_generator.py:15-140 (same constructor: everything comes from `opt`)."""

    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        nf = opt.ngf
        # nn.ReLU on the distillation path (distill_options default), nn.LeakyReLU() when SPADEModel trains the
        # teacher (models/spade_model.py:92)
        if getattr(opt, 'active_fn', 'nn.ReLU') not in ('nn.ReLU', 'nn.LeakyReLU'):
            raise NotImplementedError('cat_b200 SPADE generators implement nn.ReLU and nn.LeakyReLU (the CAT scripts)')
        self.fc_norm = SynchronizedBatchNorm2d(16 * nf, affine=True)
        self.sw, self.sh = self.compute_latent_vector_size(opt)
        self.fc = nn.Conv2d(opt.semantic_nc, 16 * nf, 3, padding=1)
        self.head_0 = SPADEInvertedResidualChannels(16 * nf, 16 * nf, opt)
        self.G_middle_0 = SPADEInvertedResidualChannels(16 * nf, 16 * nf, opt)
        self.G_middle_1 = SPADEInvertedResidualChannels(16 * nf, 16 * nf, opt)
        self.up_0 = SPADEInvertedResidualChannels(16 * nf, 8 * nf, opt)
        self.up_1 = SPADEInvertedResidualChannels(8 * nf, 4 * nf, opt)
        self.up_2 = SPADEInvertedResidualChannels(4 * nf, 2 * nf, opt)
        self.up_3 = SPADEInvertedResidualChannels(2 * nf, 1 * nf, opt)
        final_nc = nf
        if opt.num_upsampling_layers == 'most':
            self.up_4 = SPADEInvertedResidualChannels(1 * nf, nf // 2, opt)
            final_nc = nf // 2
        self.conv_img = nn.Conv2d(final_nc, 3, 3, padding=1)
        self.up = nn.Upsample(scale_factor=2)

    @staticmethod
    def compute_latent_vector_size(opt):
        n = {'normal': 5, 'more': 6, 'most': 7}.get(opt.num_upsampling_layers)
        if n is None:
            raise ValueError('opt.num_upsampling_layers [%s] not recognized' % opt.num_upsampling_layers)
        sw = opt.crop_size // (2 ** n)
        return sw, round(sw / opt.aspect_ratio)

    def block_names(self):
        return BLOCKS + (['up_4'] if self.opt.num_upsampling_layers == 'most' else [])

    def arch(self):
        """The engine's description of the *current* module tree (in-place pruning is seen)."""
        blocks = {}
        for n in self.block_names():
            b = getattr(self, n)
            blocks[n] = {'fin': int(b.input_dim), 'fout': int(b.output_dim), 'res': [int(c) for c in b.res_channels],
                         'dw': [int(c) for c in b.dw_channels], 'spade_res': [int(c) for c in b.spade.res_channels],
                         'spade_dw': [int(c) for c in b.spade.dw_channels], 'learned_shortcut': b.shortcut is not None}
        ks = self.opt.kernel_sizes
        arch = {'semantic_nc': int(self.opt.semantic_nc), 'fc_out': int(self.fc.out_channels), 'sh': int(self.sh), 'sw': int(self.sw),
                'num_upsampling_layers': self.opt.num_upsampling_layers, 'kernel_sizes': [int(k) for k in ([ks] if isinstance(ks, int) else ks)],
                'final_nc': int(self.conv_img.in_channels), 'block_names': self.block_names(), 'blocks': blocks,
                'eps': 1e-5, 'momentum': 0.1}
        if getattr(self.opt, 'active_fn', 'nn.ReLU') != 'nn.ReLU':      # the engines default to nn.ReLU
            arch['active_fn'] = self.opt.active_fn
        return arch

    @classmethod
    def from_arch(cls, arch, opt):
        """A generator with the (pruned) channel configuration of `arch` (what shrink_spade_model produces by editing
        the module tree of a teacher copy in place, utils/common.py:710-835)."""
        net = cls(opt)
        net.sh, net.sw = arch['sh'], arch['sw']
        c0 = arch['fc_out']
        net.fc = nn.Conv2d(arch['semantic_nc'], c0, 3, padding=1)
        net.fc_norm = SynchronizedBatchNorm2d(c0, affine=True)
        for n in arch['block_names']:
            a, b = arch['blocks'][n], getattr(net, n)
            b.input_dim, b.output_dim = a['fin'], a['fout']
            b.learned_shortcut = a['learned_shortcut']
            b.res_channels, b.dw_channels = list(a['res']), list(a['dw'])
            b.spade.output_dim = a['fin']
            b.spade.res_channels, b.spade.dw_channels = list(a['spade_res']), list(a['spade_dw'])
            b.res_ops, b.dw_ops, b.shortcut, b.spade = b._build(build_only=True)
            if not a['learned_shortcut']:
                b.shortcut = None
        net.conv_img = nn.Conv2d(arch['final_nc'], 3, 3, padding=1)
        return net

    def get_named_block_list(self):
        return collections.OrderedDict((n, getattr(self, n)) for n in self.block_names())

    def remove_spectral_norm(self):
        pass   # the generators of the CAT scripts carry no spectral norm

    def _compile(self, B, H, W, device, training, need_grad, share):
        from .. import ops
        from ..spade_engine import SpadeGenNet
        seg = ops.Act.empty(B, H, W, int(self.opt.semantic_nc), device, zero=True)
        arch = dict(self.arch(), sh=H >> self._n_up(), sw=W >> self._n_up())
        return SpadeGenNet(arch, seg, device, training=training, need_grad=need_grad, share=share)

    def _n_up(self):
        return {'normal': 5, 'more': 6, 'most': 7}[self.opt.num_upsampling_layers]

    def bind(self, net):
        """Adopt an already compiled engine network (the distiller's) as the owner of this module's storage."""
        self._alias_into(net)
        self.__dict__['_primary'] = net
        self.__dict__['_engines'] = {(net.B, net.H, net.W, str(net.dev), bool(net.training)): net}

    def forward(self, input, mapping_layers=()):
        """Inference on the CUDA engine: `input` = one-hot label map (+ edge channel) NCHW fp32 [B, semantic_nc, H, W]
        (what SPADEModel.preprocess_input returns) -> image NCHW fp32; with `mapping_layers` also the requested block
        outputs, like the reference forward (inception_spade_generator.py:63-124)."""
        from .. import ops
        B, C, H, W = input.shape
        net = self.engine(B, H, W, input.device, self.training, False)
        ops.nchw_to_nhwc(input.float().contiguous(), net.seg_in)
        net.pack_weights()
        out = ops.nhwc_to_nchw(net.forward(), 3)
        if not mapping_layers:
            return out
        acts = {}
        for b in net.blocks:
            if b.name in mapping_layers:
                acts[b.name] = ops.nhwc_to_nchw(b.out, b.fout)
        return out, acts


class SPADENLayerDiscriminator(BaseNetwork):
    """discriminators.py:129-180."""

    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        if opt.norm_D != 'spectralinstance':
            raise NotImplementedError('cat_b200 implements norm_D = spectralinstance (the reference default)')
        nf = opt.ndf
        input_nc = opt.semantic_nc + opt.output_nc
        seq = [[nn.Conv2d(input_nc, nf, kernel_size=4, stride=2, padding=2), nn.LeakyReLU(0.2, False)]]
        for n in range(1, opt.n_layers_D):
            prev, nf = nf, min(nf * 2, 512)
            conv = nn.Conv2d(prev, nf, kernel_size=4, stride=1 if n == opt.n_layers_D - 1 else 2, padding=2)
            conv = nn.utils.spectral_norm(conv)
            delattr(conv, 'bias')
            conv.register_parameter('bias', None)
            seq += [[nn.Sequential(conv, nn.InstanceNorm2d(nf, affine=False)), nn.LeakyReLU(0.2, False)]]
        seq += [[nn.Conv2d(nf, 1, kernel_size=4, stride=1, padding=2)]]
        for n, s in enumerate(seq):
            self.add_module('model' + str(n), nn.Sequential(*s))


class MultiscaleDiscriminator(_EngineBacked):
    """discriminators.py:183-226."""

    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        for i in range(opt.num_D):
            self.add_module('discriminator_%d' % i, SPADENLayerDiscriminator(opt))

    def arch(self):
        o = self.opt
        return {'input_nc': int(o.semantic_nc + o.output_nc), 'ndf': int(o.ndf), 'n_layers': int(o.n_layers_D),
                'num_D': int(o.num_D), 'norm_D': o.norm_D}

    def _compile(self, *a):
        raise NotImplementedError('compiled by SpadeDistillStep for the training batch')


def define_spade_G(opt):
    """networks.define_G(..., netG='inception_spade') (models/networks.py:196-199)."""
    return InceptionSPADEGenerator(opt)


def define_spade_D(opt):
    """networks.define_D(..., netD='multi_scale') (models/networks.py:258-260)."""
    return MultiscaleDiscriminator(opt)
