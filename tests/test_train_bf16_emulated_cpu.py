"""The bodies of the GPU parity tests of the teacher-training steps (tests/test_zzz_train_gpu.py) re-run on CPU with the
kernel wrappers swapped for their torch restatements in bf16-storage mode (oracle/kernel_emu.py): the stated GPU
tolerances must hold for the same launch sequences when only the storage precision of the device path is modelled."""
import importlib
import os

import pytest

G = importlib.import_module('test_zzz_train_gpu')
# the bodies below repeat what the exact-mode tests establish and only calibrate the GPU tolerances: they run with
# CATB_SLOW_TESTS=1 (as they were when the tolerances were set); the default CPU suite keeps the quick ones
slow = pytest.mark.slow


@pytest.fixture
def on_cpu():
    from oracle.kernel_emu import emulated_kernels
    G.DEV[0] = 'cpu'
    try:
        with emulated_kernels():
            yield
    finally:
        G.DEV[0] = 'cuda:0'


@pytest.mark.timeout(900)
@pytest.mark.parametrize('name', [pytest.param('train_pix2pix_in_lsgan_l2', marks=slow), 'train_pix2pix_bn_hinge'])
def test_pix2pix_bodies(golden_dir, on_cpu, name):
    G.test_pix2pix_train_step(golden_dir, name, False)


@slow
@pytest.mark.timeout(1200)
@pytest.mark.parametrize('name', ['train_cyclegan_in_lsgan'])
def test_cyclegan_bodies(golden_dir, on_cpu, name):
    G.test_cyclegan_train_steps(golden_dir, name, False)


@slow
@pytest.mark.timeout(1200)
def test_spade_body(golden_dir, on_cpu):
    G.test_spade_train_step(golden_dir, False)


@pytest.mark.timeout(300)
def test_leaky001_body(on_cpu):
    G.test_leaky001_activation_kernels()


@slow
@pytest.mark.timeout(900)
def test_mse_distill_body(golden_dir, on_cpu, monkeypatch):
    from cat_b200 import distill_engine, spade_distill_engine
    # the GPU body asks for CUDA graphs; on CPU the same launch sequence runs eagerly
    for cls in (distill_engine.DistillStep, spade_distill_engine.SpadeDistillStep):
        orig = cls.__init__

        def init(self, *a, _orig=orig, **k):
            k['use_cuda_graph'] = False
            _orig(self, *a, **k)
        monkeypatch.setattr(cls, '__init__', init)
    G.test_mse_distill_steps(golden_dir)


@slow
@pytest.mark.timeout(600)
def test_first_step_eval_mode_body(golden_dir, on_cpu):
    G.test_first_step_with_the_student_in_eval_mode(golden_dir)


@slow
@pytest.mark.timeout(2400)
def test_fullsize_bodies_at_reduced_resolution(monkeypatch):
    """tests/test_zz_fullsize_gpu.py with CATB_FULLSIZE_HW=64 under the bf16 kernel emulation (the benchmark networks at a
    quarter of the benchmark resolution)."""
    from oracle.kernel_emu import emulated_kernels
    Z = importlib.import_module('test_zz_fullsize_gpu')
    monkeypatch.setenv('CATB_FULLSIZE_HW', '64')
    Z.DEV[0] = 'cpu'
    try:
        with emulated_kernels():
            Z.test_ka_invariances_at_benchmark_size()
            Z.test_norm_statistics_at_benchmark_size(True)
            Z.test_norm_statistics_at_benchmark_size(False)
            Z.test_conv_linearity_on_the_widest_patchgan_layer()
            Z.test_pix2pix_5p6B_step_at_benchmark_resolution()
            Z.test_gaugan_5p6B_step_at_benchmark_resolution()
    finally:
        Z.DEV[0] = 'cuda:0'


@slow
@pytest.mark.timeout(900)
def test_spade_first_step_eval_mode_body(golden_dir, on_cpu):
    G.test_spade_first_step_with_the_student_in_eval_mode(golden_dir)
