"""cat_b200.prune (host-side restatement of the student architecture search of utils/common.shrink_model)
against golden cases written by the REAL reference `shrink` (oracle/make_golden_prune.py): the MAC model must equal
model_profiling's count and the bisection must land on exactly the reference's channel counts (integer work: bit exact)."""
import json
import os

import pytest
import torch

from cat_b200 import prune


@pytest.fixture(scope='module')
def cases(golden_dir):
    return torch.load(os.path.join(golden_dir, 'prune_cases.pt'), weights_only=False)


@pytest.mark.parametrize('name', ['small_bn', 'small_in', 'small_in_half', 'pix2pix_5p6B', 'cyclegan_2p6B'])
def test_mac_model_matches_model_profiling(cases, name):
    c = cases[name]
    assert prune.generator_macs(c['teacher_arch'], c['H'], c['W']) == c['teacher_macs']
    assert prune.generator_macs(c['student_arch'], c['H'], c['W']) == c['student_macs']


@pytest.mark.parametrize('name', ['small_bn', 'small_in', 'small_in_half', 'pix2pix_5p6B', 'cyclegan_2p6B'])
def test_shrink_arch_reproduces_the_reference(cases, name):
    c = cases[name]
    student, info = prune.shrink_arch(c['gammas'], c['teacher_arch'], c['target_flops'], c['H'], c['W'], prune_cin_lb=c['prune_cin_lb'])
    assert student['widths'] == c['student_arch']['widths']
    assert student['blocks'] == c['student_arch']['blocks']
    assert info['macs'] == c['student_macs'] and info['macs'] <= c['target_flops']


def test_published_bench_architectures_come_from_this_search(cases, golden_dir):
    """The committed benchmark architectures (tests/golden/arch_*.json, written by the reference) are reproduced."""
    for name in ('pix2pix_5p6B', 'cyclegan_2p6B'):
        c = cases[name]
        ref = json.load(open(os.path.join(golden_dir, f'arch_{name}.json')))
        student, info = prune.shrink_arch(c['gammas'], c['teacher_arch'], c['target_flops'], c['H'], c['W'], prune_cin_lb=c['prune_cin_lb'])
        assert student == ref['student_arch'] and info['macs'] == ref['student_macs']


def test_unreachable_target_raises(cases):
    c = cases['small_bn']
    with pytest.raises(RuntimeError):
        prune.shrink_arch(c['gammas'], c['teacher_arch'], 10.0, c['H'], c['W'], prune_cin_lb=c['prune_cin_lb'])


def test_shrink_drop_in_on_the_module_mirror(cases):
    """prune.shrink(model, opt) on a distiller-like object: the student mirror is rebuilt with the searched architecture
    (reference state_dict layout), initialised, and the compiled engine is dropped."""
    import argparse
    from types import SimpleNamespace
    from cat_b200.models import networks
    c = cases['small_bn']
    ta = c['teacher_arch']
    opt = argparse.Namespace(channels=None, channels_reduction_factor=6, kernel_sizes=ta['kernel_sizes'], norm_momentum=0.1,
                             norm_epsilon=1e-5, active_fn='nn.ReLU', norm_affine=True, norm_track_running_stats=True,
                             target_flops=c['target_flops'], data_height=c['H'], data_width=c['W'], prune_cin_lb=c['prune_cin_lb'],
                             init_type='normal', init_gain=0.02)
    teacher = networks.define_G(3, 3, ta['widths'][0], 'inception_9blocks', 'batch', 0, 'normal', 0.02, [], opt=opt)
    teacher.load_state_dict(c['gammas'], strict=False)
    model = SimpleNamespace(netG_teacher=teacher, netG_student=None, gpu_ids=[], engine=object())
    info = prune.shrink(model, opt)
    assert model.engine is None
    assert model.netG_student.arch()['widths'] == c['student_arch']['widths']
    assert model.netG_student.arch()['blocks'] == c['student_arch']['blocks']
    assert model.netG_student.n_macs == c['student_macs'] == info['macs'] and teacher.n_macs == c['teacher_macs']


# ---- SPADE generator (shrink_spade_model) ------------------------------------------------------------------------------
@pytest.fixture(scope='module')
def spade_cases(golden_dir):
    return torch.load(os.path.join(golden_dir, 'prune_spade_cases.pt'), weights_only=False)


@pytest.mark.parametrize('name', ['spade_small', 'spade_small_b', 'gaugan_5p6B'])
def test_spade_mac_model_and_search_reproduce_the_reference(spade_cases, name):
    c = spade_cases[name]
    assert prune.spade_generator_macs(c['teacher_arch']) == c['teacher_macs']
    assert prune.spade_generator_macs(c['student_arch']) == c['student_macs']
    student, info = prune.shrink_spade_arch(c['gammas'], c['teacher_arch'], c['target_flops'], prune_cin_lb=c['prune_cin_lb'])
    ref = c['student_arch']
    assert student['fc_out'] == ref['fc_out'] and student['final_nc'] == ref['final_nc']
    assert student['blocks'] == ref['blocks']
    assert info['macs'] == c['student_macs'] and info['macs'] <= c['target_flops']


def test_published_gaugan_bench_architecture_comes_from_this_search(spade_cases, golden_dir):
    c = spade_cases['gaugan_5p6B']
    ref = json.load(open(os.path.join(golden_dir, 'arch_gaugan_5p6B.json')))
    student, info = prune.shrink_spade_arch(c['gammas'], c['teacher_arch'], c['target_flops'], prune_cin_lb=c['prune_cin_lb'])
    assert student == ref['student_arch'] and info['macs'] == ref['student_macs']


def test_shrink_drop_in_on_the_spade_mirror(spade_cases):
    import argparse
    from types import SimpleNamespace
    from torch import nn
    from cat_b200.models import networks
    c = spade_cases['spade_small']
    ta = c['teacher_arch']
    ngf = ta['fc_out'] // 16
    t_opt = argparse.Namespace(ngf=ngf, norm_G='spadesyncbatch3x3', semantic_nc=ta['semantic_nc'],
                               num_upsampling_layers=ta['num_upsampling_layers'], crop_size=ta['sw'] * 64, aspect_ratio=ta['sw'] / ta['sh'],
                               channels=None, channels_reduction_factor=6, kernel_sizes=ta['kernel_sizes'], active_fn='nn.ReLU')
    teacher = networks.define_G(ta['semantic_nc'] - 1, 3, ngf, 'inception_spade', 'instance', 0, 'xavier', 0.02, [], opt=t_opt)
    assert teacher.arch() == ta
    teacher.load_state_dict(c['gammas'], strict=False)
    mm = SimpleNamespace(netG_teacher=teacher, netG_student=None, netAs=None, mapping_layers=['head_0', 'G_middle_1', 'up_1'])
    model = SimpleNamespace(modules_on_one_gpu=mm, gpu_ids=[], engine=object())
    opt = argparse.Namespace(target_flops=c['target_flops'], prune_cin_lb=c['prune_cin_lb'], init_type='xavier', init_gain=0.02,
                             teacher_ngf=ngf)
    info = prune.shrink(model, opt)
    assert model.engine is None and info['macs'] == c['student_macs']
    assert mm.netG_student.arch() == c['student_arch']
    assert isinstance(mm.netAs, nn.ModuleList) and mm.netAs[2].in_channels == c['student_arch']['fc_out'] // 4
