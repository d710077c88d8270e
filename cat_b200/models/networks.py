"""Mirror of the reference network factories (models/networks.py:167-266) and module classes
(inception_generator.py:11-145, inception_modules.py:22-243, discriminators.py:14-79).

The classes below keep the reference's module tree, attribute names and ``state_dict`` keys (they are
load-bearing for checkpoints, the distillation hooks ``down_sampling.9`` / ``features.{2,5,8}`` and the
pruning code), but they are parameter *containers*: no torch op ever computes with them.
``forward`` compiles a cat_b200 engine network (GenNet / DisNet) for the input shape, re-points every
parameter and buffer at the engine's flat arena (so ``load_state_dict`` / optimiser updates / checkpoints
see one storage) and runs the libcatb200 kernels.  Training does not go through ``forward`` + autograd:
it is ``InceptionDistiller.optimize_parameters`` (cat_b200/distillers), which drives the fused step.
"""
import collections
import functools

import torch
from torch import nn



# --------------------------------------------------------------------------------------------------
# helpers shared with the reference API
# --------------------------------------------------------------------------------------------------
class BaseNetwork(nn.Module):
    @staticmethod
    def modify_commandline_options(parser, is_train):
        return parser


class Identity(nn.Module):
    def forward(self, x):
        return x


def get_norm_layer(norm_type='instance', affine=True, track_running_stats=True):
    """models/networks.py:29-64 (batch | instance | none)."""
    if norm_type == 'batch':
        return functools.partial(nn.BatchNorm2d, affine=affine, track_running_stats=track_running_stats)
    if norm_type == 'instance':
        return functools.partial(nn.InstanceNorm2d, affine=affine, track_running_stats=track_running_stats)
    if norm_type == 'none':
        return lambda x: Identity()
    raise NotImplementedError('normalization layer [%s] is not found' % norm_type)


def get_active_fn(name):
    return {'nn.ReLU6': functools.partial(nn.ReLU6, inplace=True), 'nn.ReLU': functools.partial(nn.ReLU, inplace=True),
            'nn.LeakyReLU': functools.partial(nn.LeakyReLU, inplace=True)}[name]


def init_weights(net, init_type='normal', init_gain=0.02):
    """models/networks.py:108-144."""
    def init_func(m):
        classname = m.__class__.__name__
        if hasattr(m, 'weight') and (classname.find('Conv') != -1 or classname.find('Linear') != -1):
            if init_type == 'normal':
                nn.init.normal_(m.weight.data, 0.0, init_gain)
            elif init_type == 'xavier':
                nn.init.xavier_normal_(m.weight.data, gain=init_gain)
            elif init_type == 'kaiming':
                nn.init.kaiming_normal_(m.weight.data, a=0, mode='fan_in')
            elif init_type == 'orthogonal':
                nn.init.orthogonal_(m.weight.data, gain=init_gain)
            else:
                raise NotImplementedError('initialization method [%s] is not implemented' % init_type)
            if hasattr(m, 'bias') and m.bias is not None:
                nn.init.constant_(m.bias.data, 0.0)
        elif classname.find('BatchNorm2d') != -1:
            if hasattr(m, 'weight') and m.weight is not None:
                nn.init.normal_(m.weight.data, 1.0, init_gain)
            if hasattr(m, 'bias') and m.weight is not None:
                nn.init.constant_(m.bias.data, 0.0)
    net.apply(init_func)


def _norm_flags(norm_layer):
    func = norm_layer.func if isinstance(norm_layer, functools.partial) else norm_layer
    kw = norm_layer.keywords if isinstance(norm_layer, functools.partial) else {}
    kind = 'batch' if func is nn.BatchNorm2d else ('instance' if func is nn.InstanceNorm2d else None)
    if kind is None:
        raise NotImplementedError('cat_b200 supports BatchNorm2d and InstanceNorm2d (got %r)' % (func,))
    return kind, bool(kw.get('affine', func is nn.BatchNorm2d)), bool(kw.get('track_running_stats', func is nn.BatchNorm2d))


# --------------------------------------------------------------------------------------------------
# residual block (parameter container with the reference's build logic)
# --------------------------------------------------------------------------------------------------
class ConvBNReLU(nn.Sequential):
    """inception_modules.py:22-44."""

    def __init__(self, in_planes, out_planes, kernel_size=3, stride=1, groups=1, use_bias=True,
                 norm_layer=nn.InstanceNorm2d, norm_kwargs=None, active_fn=None):
        super().__init__(nn.Conv2d(in_planes, out_planes, kernel_size, stride, 0, groups=groups, bias=use_bias),
                         norm_layer(out_planes, **(norm_kwargs or {})), active_fn())


class InvertedResidualChannels(nn.Module):
    """inception_modules.py:47-243: six-branch residual block; `res_channels` / `dw_channels` are the
    per-kernel-size branch widths (0 = branch absent), `_build()` re-creates the branches from them."""

    def __init__(self, inp, res_channels, dw_channels, channels_reduction_factor, res_kernel_sizes, dw_kernel_sizes,
                 padding_type='reflect', use_bias=True, norm_layer=nn.InstanceNorm2d, norm_kwargs=None,
                 dropout_rate=0.0, active_fn=None):
        super().__init__()
        if padding_type != 'reflect':
            raise NotImplementedError('cat_b200 implements the reflect padding used by every CAT script')
        def widths(ch, ks):
            if ch is None:
                return [inp // channels_reduction_factor for _ in ks]
            if isinstance(ch, int):
                return [ch // channels_reduction_factor for _ in ks]
            assert len(ch) == len(ks)
            return [c // channels_reduction_factor for c in ch]
        res_kernel_sizes = [res_kernel_sizes] if isinstance(res_kernel_sizes, int) else list(res_kernel_sizes)
        dw_kernel_sizes = [dw_kernel_sizes] if isinstance(dw_kernel_sizes, int) else list(dw_kernel_sizes)
        self.input_dim = inp
        self.res_channels = widths(res_channels, res_kernel_sizes)
        self.dw_channels = widths(dw_channels, dw_kernel_sizes)
        self.res_kernel_sizes, self.dw_kernel_sizes = res_kernel_sizes, dw_kernel_sizes
        self.padding_type, self.use_bias = padding_type, use_bias
        self.norm_layer, self.norm_kwargs = norm_layer, norm_kwargs
        self.dropout_rate, self.active_fn = dropout_rate, active_fn
        self.pad = nn.ReflectionPad2d
        self.res_ops, self.dw_ops, self.pw_bn = self._build()

    def _build(self):
        kw = self.norm_kwargs or {}
        res_ops = nn.ModuleList()
        for midp, k in zip(self.res_channels, self.res_kernel_sizes):
            if midp == 0:
                continue
            res_ops.append(nn.Sequential(
                self.pad((k - 1) // 2),
                ConvBNReLU(self.input_dim, midp, kernel_size=k, use_bias=self.use_bias, norm_layer=self.norm_layer,
                           norm_kwargs=kw, active_fn=self.active_fn),
                nn.Dropout(self.dropout_rate), self.pad((k - 1) // 2),
                nn.Conv2d(midp, self.input_dim, k, 1, 0, bias=self.use_bias)))
        dw_ops = nn.ModuleList()
        for midp, k in zip(self.dw_channels, self.dw_kernel_sizes):
            if midp == 0:
                continue
            dw_ops.append(nn.Sequential(
                ConvBNReLU(self.input_dim, midp, kernel_size=1, use_bias=self.use_bias, norm_layer=self.norm_layer,
                           norm_kwargs=kw, active_fn=self.active_fn),
                self.pad((k - 1) // 2),
                ConvBNReLU(midp, midp, kernel_size=k, groups=midp, use_bias=self.use_bias, norm_layer=self.norm_layer,
                           norm_kwargs=kw, active_fn=self.active_fn),
                nn.Dropout(self.dropout_rate), nn.Conv2d(midp, self.input_dim, 1, 1, 0, bias=self.use_bias)))
        return res_ops, dw_ops, self.norm_layer(self.input_dim, **kw)

    # accessors used by the reference pruning code (inception_modules.py:182-228)
    def get_named_first_res_bn(self, prefix=None):
        return collections.OrderedDict((_pre(f'res_ops.{i}.1.1', prefix), op[1][1]) for i, op in enumerate(self.res_ops))

    def get_named_first_dw_bn(self, prefix=None):
        return collections.OrderedDict((_pre(f'dw_ops.{i}.0.1', prefix), op[0][1]) for i, op in enumerate(self.dw_ops))

    def get_named_first_bn(self, prefix=None):
        return collections.OrderedDict(list(self.get_named_first_res_bn(prefix).items()) +
                                       list(self.get_named_first_dw_bn(prefix).items()))

    def get_first_res_bn(self):
        return list(self.get_named_first_res_bn().values())

    def get_first_dw_bn(self):
        return list(self.get_named_first_dw_bn().values())

    def get_first_bn(self):
        return self.get_first_res_bn() + self.get_first_dw_bn()

    def forward(self, x):
        raise RuntimeError('blocks are executed by the compiled generator (InceptionGenerator.forward), not one by one')


def _pre(name, prefix):
    return name if prefix is None else f'{prefix}.{name}'


# --------------------------------------------------------------------------------------------------
# generator
# --------------------------------------------------------------------------------------------------
class _EngineBacked(nn.Module):
    """Mixin: compiles engine networks for (input shape, mode) on demand; the first one owns the flat arenas
    and every parameter / buffer of the module tree is re-pointed at them, later ones share them."""

    def _alias_into(self, net):
        sd = {k: v.detach().clone() for k, v in self.state_dict().items()}
        net.load_state_dict(sd)
        for name, p in list(self.named_parameters()) + list(self.named_buffers()):
            if name.endswith('num_batches_tracked'):
                continue
            src = net.arena if net.arena.has(name) else net.bufs
            p.data = src.view(name)

    def bind(self, net):
        """Adopt an already compiled engine network (the distiller's) as the owner of this module's storage."""
        self._alias_into(net)
        self.__dict__['_primary'] = net
        self.__dict__['_engines'] = {(net.B, net.H, net.W, str(net.dev), bool(getattr(net, 'training', True))): net}

    def engine(self, B, H, W, device, training, need_grad):
        cache = self.__dict__.setdefault('_engines', {})
        key = (B, H, W, str(device), bool(training))
        hit = cache.get(key)
        if hit is not None and bool(getattr(hit, 'training', training)) != bool(training):
            hit = None        # the distiller's network was switched between train() and eval() since it was registered
        if hit is None:
            primary = self.__dict__.get('_primary')
            net = self._compile(B, H, W, device, training, need_grad, primary)
            if primary is None:
                self._alias_into(net)
                self.__dict__['_primary'] = net
            cache[key] = net
        return cache[key]


class InceptionGenerator(BaseNetwork, _EngineBacked):
    """inception_generator.py:11-145 (same constructor, same module tree)."""

    def __init__(self, input_nc, output_nc, ngf, channels, channels_reduction_factor, kernel_sizes,
                 padding_type='reflect', norm_layer=nn.InstanceNorm2d, norm_momentum=0.1, norm_epsilon=1e-5,
                 dropout_rate=0, active_fn='nn.ReLU', n_blocks=9, widths=None, block_channels=None):
        """`widths` = (c0, c1, c2, c3, c4) and `block_channels` = [{'res': [...], 'dw': [...]}] describe a
        pruned student (what the reference's shrink_model produces by editing modules in place)."""
        assert n_blocks >= 0 and len(kernel_sizes) == len(set(kernel_sizes))
        super().__init__()
        if dropout_rate != 0:
            raise NotImplementedError('dropout_rate != 0 is not on the CAT distillation path (all scripts use 0)')
        if active_fn != 'nn.ReLU':
            raise NotImplementedError('cat_b200 generators use nn.ReLU (the only activation the CAT scripts use)')
        kind, affine, track = _norm_flags(norm_layer)
        use_bias = kind == 'instance'
        kw = {'momentum': norm_momentum, 'eps': norm_epsilon}
        act = get_active_fn(active_fn)
        c0, c1, c2, c3, c4 = widths or (ngf, ngf * 2, ngf * 4, ngf * 2, ngf)
        self.down_sampling = nn.Sequential(
            nn.ReflectionPad2d(3), nn.Conv2d(input_nc, c0, kernel_size=7, padding=0, bias=use_bias), norm_layer(c0),
            nn.ReLU(True),
            nn.Conv2d(c0, c1, kernel_size=3, stride=2, padding=1, bias=use_bias), norm_layer(c1), nn.ReLU(True),
            nn.Conv2d(c1, c2, kernel_size=3, stride=2, padding=1, bias=use_bias), norm_layer(c2), nn.ReLU(True))
        feats = []
        for i in range(n_blocks):
            blk = InvertedResidualChannels(c2, res_channels=channels, dw_channels=channels,
                                           channels_reduction_factor=channels_reduction_factor,
                                           res_kernel_sizes=kernel_sizes, dw_kernel_sizes=kernel_sizes,
                                           padding_type=padding_type, use_bias=use_bias, norm_layer=norm_layer,
                                           norm_kwargs=kw, dropout_rate=dropout_rate, active_fn=act)
            if block_channels is not None:
                blk.res_channels, blk.dw_channels = list(block_channels[i]['res']), list(block_channels[i]['dw'])
                blk.res_ops, blk.dw_ops, blk.pw_bn = blk._build()
            feats.append(blk)
        self.features = nn.Sequential(*feats)
        self.up_sampling = nn.Sequential(
            nn.ConvTranspose2d(c2, c3, kernel_size=3, stride=2, padding=1, output_padding=1, bias=use_bias), norm_layer(c3),
            nn.ReLU(True),
            nn.ConvTranspose2d(c3, c4, kernel_size=3, stride=2, padding=1, output_padding=1, bias=use_bias), norm_layer(c4),
            nn.ReLU(True), nn.ReflectionPad2d(3), nn.Conv2d(c4, output_nc, kernel_size=7, padding=0), nn.Tanh())
        self._meta = dict(input_nc=input_nc, output_nc=output_nc, kernel_sizes=list(kernel_sizes), norm=kind, affine=affine,
                          track_running_stats=track, eps=norm_epsilon, momentum=norm_momentum, use_bias=use_bias)

    def arch(self):
        """The engine's description of this module tree (reads the *current* modules, so in-place pruning is seen)."""
        ds, us = self.down_sampling, self.up_sampling
        return dict(self._meta, widths=[ds[1].out_channels, ds[4].out_channels, ds[7].out_channels,
                                        us[0].out_channels, us[3].out_channels],
                    blocks=[{'res': [int(c) for c in b.res_channels], 'dw': [int(c) for c in b.dw_channels]}
                            for b in self.features])

    @classmethod
    def from_arch(cls, arch):
        norm_layer = get_norm_layer(arch['norm'], arch['affine'], arch['track_running_stats'])
        return cls(arch['input_nc'], arch['output_nc'], arch['widths'][0], None, 1, arch['kernel_sizes'],
                   norm_layer=norm_layer, norm_momentum=arch['momentum'], norm_epsilon=arch['eps'],
                   n_blocks=len(arch['blocks']), widths=arch['widths'], block_channels=arch['blocks'])

    def get_named_block_list(self):
        return collections.OrderedDict(('features.{}'.format(n), b) for n, b in self.features.named_children())

    def _compile(self, B, H, W, device, training, need_grad, share):
        from ..engine import GenNet
        return GenNet(self.arch(), B, H, W, device, training=training, need_grad=need_grad, share=share)

    def forward(self, input):
        """NCHW fp32 in -> NCHW fp32 out on the CUDA engine (inference; also fills `self.mapped` with the four
        distillation activations, the role of the reference's forward hooks)."""
        from .. import ops
        B, _, H, W = input.shape
        net = self.engine(B, H, W, input.device, self.training, False)
        x = ops.Act.empty(B, H, W, self._meta['input_nc'], input.device, zero=True)
        ops.nchw_to_nhwc(input.float().contiguous(), x)
        net.pack_weights()
        out = net.forward(x)
        self.mapped = {n: ops.nhwc_to_nchw(a, self.down_sampling[7].out_channels) for n, a in net.acts.items()}
        return ops.nhwc_to_nchw(out, self._meta['output_nc'])


# --------------------------------------------------------------------------------------------------
# discriminator
# --------------------------------------------------------------------------------------------------
class NLayerDiscriminator(BaseNetwork, _EngineBacked):
    """discriminators.py:14-79 (70x70 PatchGAN for n_layers=3)."""

    def __init__(self, input_nc, ndf=64, n_layers=3, norm_layer=nn.BatchNorm2d, active_fn='nn.LeakyReLU'):
        super().__init__()
        if active_fn != 'nn.LeakyReLU':
            raise NotImplementedError('cat_b200 discriminators use nn.LeakyReLU(0.2) (distill_options default)')
        kind, affine, track = _norm_flags(norm_layer)
        use_bias = kind == 'instance'
        act = get_active_fn(active_fn)
        seq = [nn.Conv2d(input_nc, ndf, kernel_size=4, stride=2, padding=1), act(0.2)]
        mult = 1
        for n in range(1, n_layers):
            prev, mult = mult, min(2 ** n, 8)
            seq += [nn.Conv2d(ndf * prev, ndf * mult, kernel_size=4, stride=2, padding=1, bias=use_bias),
                    norm_layer(ndf * mult), act(0.2)]
        prev, mult = mult, min(2 ** n_layers, 8)
        seq += [nn.Conv2d(ndf * prev, ndf * mult, kernel_size=4, stride=1, padding=1, bias=use_bias),
                norm_layer(ndf * mult), act(0.2)]
        seq += [nn.Conv2d(ndf * mult, 1, kernel_size=4, stride=1, padding=1)]
        self.model = nn.Sequential(*seq)
        n0 = self.model[3]
        self._meta = dict(input_nc=input_nc, ndf=ndf, n_layers=n_layers, norm=kind, affine=affine,
                          track_running_stats=track, eps=n0.eps, momentum=n0.momentum, use_bias=use_bias)

    def arch(self):
        return dict(self._meta)

    def _compile(self, B, H, W, device, training, need_grad, share):
        from ..engine import DisNet
        if share is not None:
            raise NotImplementedError('a discriminator is compiled for one input shape (the training batch)')
        return DisNet(self.arch(), B, H, W, device)

    def forward(self, input):
        from .. import ops
        B, C, H, W = input.shape
        net = self.engine(B, H, W, input.device, True, True)
        x = ops.Act.empty(B, H, W, C, input.device, zero=True)
        ops.nchw_to_nhwc(input.float().contiguous(), x)
        net.pack_weights()
        pred = net.forward(x)
        return pred[..., :1].permute(0, 3, 1, 2).contiguous()


# --------------------------------------------------------------------------------------------------
# factories (same signatures as the reference)
# --------------------------------------------------------------------------------------------------
def init_net(net, init_type='normal', init_gain=0.02, gpu_ids=()):
    if len(gpu_ids) > 1:
        raise NotImplementedError('cat_b200 runs one process per GPU (torch.distributed), not nn.DataParallel: '
                                  'launch with torchrun and pass a single gpu id per process')
    if len(gpu_ids) > 0:
        assert torch.cuda.is_available()
        net.to(gpu_ids[0])
    init_weights(net, init_type, init_gain=init_gain)
    return net


def define_G(input_nc, output_nc, ngf, netG, norm='batch', dropout_rate=0, init_type='normal', init_gain=0.02,
             gpu_ids=(), opt=None):
    """models/networks.py:167-202.  `opt.arch_G` (optional) carries a pruned architecture."""
    norm_layer = get_norm_layer(norm_type=norm, affine=getattr(opt, 'norm_affine', False),
                                track_running_stats=getattr(opt, 'norm_track_running_stats', False))
    if netG == 'inception_spade':
        from .spade_networks import define_spade_G
        return init_net(define_spade_G(opt), init_type, init_gain, gpu_ids)
    if netG != 'inception_9blocks':
        raise NotImplementedError('Generator model name [%s] is not on the CAT distillation path' % netG)
    net = InceptionGenerator(input_nc, output_nc, ngf=ngf, channels=opt.channels,
                             channels_reduction_factor=opt.channels_reduction_factor, kernel_sizes=opt.kernel_sizes,
                             norm_layer=norm_layer, norm_momentum=opt.norm_momentum, norm_epsilon=opt.norm_epsilon,
                             dropout_rate=dropout_rate, active_fn=opt.active_fn, n_blocks=9)
    return init_net(net, init_type, init_gain, gpu_ids)


def define_D(input_nc, ndf, netD, n_layers_D=3, norm='batch', init_type='normal', init_gain=0.02, gpu_ids=(), opt=None):
    """models/networks.py:205-266."""
    norm_layer = get_norm_layer(norm_type=norm, affine=getattr(opt, 'norm_affine_D', False),
                                track_running_stats=getattr(opt, 'norm_track_running_stats', False))
    if netD == 'multi_scale':
        from .spade_networks import define_spade_D
        return init_net(define_spade_D(opt), init_type, init_gain, gpu_ids)
    if netD != 'n_layers':
        raise NotImplementedError('Discriminator model name [%s] is not on the CAT distillation path' % netD)
    net = NLayerDiscriminator(input_nc, ndf, n_layers_D, norm_layer=norm_layer, active_fn=opt.active_fn_D)
    return init_net(net, init_type, init_gain, gpu_ids)
