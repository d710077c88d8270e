"""Host-side checks that need no GPU: the module-tree mirrors keep the reference's state_dict layout
(keys, shapes, hook names) as recorded in the golden fixtures, the C ABI exports every declared symbol,
and the engine compiles (table building, arena layout) for every fixture architecture."""
import argparse
import os
import re

import pytest
import torch

from cat_b200 import _C
from cat_b200.models import networks

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = ['pix2pix_bn_hinge', 'cyclegan_in_lsgan', 'pix2pix_bn_lsgan_l2']


def _opt(norm, track):
    return argparse.Namespace(channels=None, channels_reduction_factor=6, kernel_sizes=[1, 3, 5], norm_momentum=0.1,
                              norm_epsilon=1e-5, active_fn='nn.ReLU', active_fn_D='nn.LeakyReLU', norm_affine=True,
                              norm_affine_D=True, norm_track_running_stats=track)


@pytest.mark.parametrize('name', CASES)
def test_module_trees_match_reference_state_dicts(golden_dir, name):
    fix = torch.load(os.path.join(golden_dir, name + '.pt'), weights_only=False)
    norm = fix['teacher_arch']['norm']
    opt = _opt(norm, fix['teacher_arch']['track_running_stats'])
    T = networks.define_G(3, 3, fix['teacher_arch']['widths'][0], 'inception_9blocks', norm, 0, 'normal', 0.02, [], opt=opt)
    S = networks.InceptionGenerator.from_arch(fix['student_arch'])
    D = networks.define_D(fix['D_arch']['input_nc'], fix['D_arch']['ndf'], 'n_layers', 3, norm, 'normal', 0.02, [], opt=opt)
    for net, sd, arch in ((T, fix['teacher_sd'], fix['teacher_arch']), (S, fix['student_sd0'], fix['student_arch']),
                          (D, fix['D_sd0'], fix['D_arch'])):
        mine = net.state_dict()
        assert list(mine.keys()) == list(sd.keys())
        assert all(mine[k].shape == v.shape for k, v in sd.items())
        net.load_state_dict(sd)                       # reference checkpoints load as they are
        assert net.arch() == arch                      # and describe the same architecture to the engine
    names = dict(S.named_modules())
    for hook in ('down_sampling.9', 'features.2', 'features.5', 'features.8'):   # base_inception_distiller.py:183-190
        assert hook in names
    blk = S.features[0]
    assert list(blk.get_named_first_bn().keys())[0].startswith('res_ops.0.1.1') or blk.res_channels[0] == 0
    assert list(S.get_named_block_list().keys())[:2] == ['features.0', 'features.1']


@pytest.mark.parametrize('name', CASES)
def test_engine_compiles_on_cpu_and_round_trips_state(golden_dir, name):
    from cat_b200.engine import DisNet, GenNet
    fix = torch.load(os.path.join(golden_dir, name + '.pt'), weights_only=False)
    B, _, H, W = fix['steps'][0]['real_A'].shape
    for cls, arch, sd, kw in ((GenNet, fix['teacher_arch'], fix['teacher_sd'], dict(training=False, need_grad=False)),
                              (GenNet, fix['student_arch'], fix['student_sd0'], dict(training=True, need_grad=True)),
                              (DisNet, fix['D_arch'], fix['D_sd0'], {})):
        net = cls(arch, B, H, W, 'cpu', **kw)
        net.arena.load_state_dict(sd)
        net.bufs.load_state_dict(sd)
        back = net.state_dict()
        for k, v in sd.items():
            if v.is_floating_point():
                assert torch.equal(back[k].float().reshape(v.shape), v.float()), k
        gs = net.fprop_gemms + net.bwd_gemms
        assert gs and all(g.halo is not None for g in gs), 'every GEMM of the path qualifies for the halo kernel'


def test_c_abi_exports_every_declared_symbol():
    lib = _C.load()
    header = open(os.path.join(ROOT, 'include', 'catb200.h')).read()
    declared = set(re.findall(r'\b(catb_[a-z0-9_]+)\s*\(', header))
    assert declared, 'no declarations parsed'
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, f'symbols declared in include/catb200.h but not exported: {missing}'
    assert set(_C.EXPORTED_SYMBOLS) <= declared
    assert lib.catb_version().decode().startswith('catb200')
    assert lib.catb_packed_weight_bytes(300, 144, 160) == 2 * 160 * 18 * 128


def test_product_path_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, 'cat_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dirpath, f)).read()
                assert 'oracle' not in re.sub(r'"""[\s\S]*?"""', '', src), f'{f} must not reference the test oracle'


def test_spade_module_trees_match_reference_state_dicts(golden_dir):
    """The SPADE mirrors (InceptionSPADEGenerator, MultiscaleDiscriminator) keep the reference's state_dict keys, order
    and shapes -- teacher, pruned student (from_arch) and spectral-norm discriminator -- and describe the same
    architecture to the engine."""
    from cat_b200.models.spade_networks import InceptionSPADEGenerator
    fix = torch.load(os.path.join(golden_dir, 'spade_more.pt'), weights_only=False)
    Ta, Sa, Da = fix['teacher_arch'], fix['student_arch'], fix['D_arch']
    opt = argparse.Namespace(ngf=Ta['fc_out'] // 16, norm_G='spadesyncbatch3x3', semantic_nc=Ta['semantic_nc'],
                             num_upsampling_layers=Ta['num_upsampling_layers'], crop_size=128, aspect_ratio=2.0, channels=None,
                             channels_reduction_factor=6, kernel_sizes=[1, 3, 5], active_fn='nn.ReLU', norm_D=Da['norm_D'],
                             ndf=Da['ndf'], n_layers_D=Da['n_layers'], num_D=Da['num_D'], output_nc=3)
    T = networks.define_G(opt.semantic_nc - 1, 3, opt.ngf, 'inception_spade', 'instance', 0, 'xavier', 0.02, [], opt=opt)
    S = InceptionSPADEGenerator.from_arch(Sa, opt)
    D = networks.define_D(Da['input_nc'], Da['ndf'], 'multi_scale', Da['n_layers'], 'instance', 'xavier', 0.02, [], opt=opt)
    for net, sd, arch in ((T, fix['teacher_sd'], Ta), (S, fix['student_sd0'], Sa), (D, fix['D_sd0'], Da)):
        mine = net.state_dict()
        assert list(mine.keys()) == list(sd.keys())
        assert all(mine[k].shape == v.shape for k, v in sd.items())
        net.load_state_dict(sd)
        assert net.arch() == arch
    assert list(S.get_named_block_list().keys()) == Sa['block_names']
    assert list(S.head_0.get_named_first_bn().keys())[0] == 'res_ops.0.0.norm'


def test_spade_engine_compiles_on_cpu_and_round_trips_state(golden_dir):
    """Table building / arena layout of the SPADE networks for the fixture architecture, and state_dict round trip
    (every reference key is stored, nothing else)."""
    from cat_b200.ops import Act
    from cat_b200.spade_engine import MultiScaleDis, SpadeGenNet, VggNet
    fix = torch.load(os.path.join(golden_dir, 'spade_more.pt'), weights_only=False)
    B, _, H, W = fix['steps'][0]['image'].shape
    seg = Act.empty(B, H, W, fix['teacher_arch']['semantic_nc'], 'cpu', zero=True)
    for arch, sd, need_grad in ((fix['teacher_arch'], fix['teacher_sd'], False), (fix['student_arch'], fix['student_sd0'], True)):
        net = SpadeGenNet(arch, seg, 'cpu', training=need_grad, need_grad=need_grad, alloc_only=True)
        keys = set(net.arena.entries) | set(net.bufs.entries)
        assert keys == set(sd.keys()), (sorted(keys - set(sd))[:3], sorted(set(sd) - keys)[:3])
        for k, v in sd.items():
            ent = (net.arena.entries.get(k) or net.bufs.entries.get(k))
            assert tuple(ent[1]) == (tuple(v.shape) or (1,)), k
    D = MultiScaleDis(fix['D_arch'], 2 * B, H, W, 'cpu', alloc_only=True)
    assert set(D.arena.entries) | set(D.bufs.entries) == set(fix['D_sd0'].keys())
    V = VggNet(B, H, W, 'cpu', alloc_only=True)
    assert len(V.arena.entries) == 26
