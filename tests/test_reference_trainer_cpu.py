"""The reference's OWN driver, unmodified, on cat_b200: ``trainer.Trainer('distill')`` (trainer.py:38-175) imported from
the reference checkout after ``cat_b200.install.install()`` -- options parsing through options/distill_options.py (which
asks ``distillers.get_option_setter`` for the flags), the reference DataLoader over PNG files on disk, ``shrink`` +
``init_net`` before the loop, ``set_input`` / ``optimize_parameters`` / ``get_current_losses`` per batch, ``evaluate_model``
and ``save_networks`` at the first iteration and at the end of the epoch, ``update_learning_rate``.  The kernels are
emulated on CPU (exact mode); only tensorboardX -- not installed in this image -- is stubbed.  Needs the reference:
/root/reference in the build container or the copy staged by oracle/make_ref.py; skipped otherwise."""
import os
import sys
import types

import pytest
import torch


def _ref_root():
    from oracle import ref_harness
    root = ref_harness.REF_ROOT
    return root if os.path.isdir(os.path.join(root, 'options')) else None


@pytest.mark.timeout(1200)
def test_unmodified_reference_trainer_runs_on_the_mirrors(tmp_path, monkeypatch):
    root = _ref_root()
    if root is None:
        pytest.skip('no reference checkout (run oracle/make_ref.py in the build container)')
    import numpy as np
    from PIL import Image
    from oracle.kernel_emu import emulated_kernels
    # ---- a tiny aligned dataset on disk: A|B side by side, 5 images -> batches of 2, 2, 1 (the partial batch ends the epoch)
    data = tmp_path / 'data'
    (data / 'train').mkdir(parents=True)
    (data / 'val').mkdir()
    rng = np.random.RandomState(0)
    for i in range(5):
        Image.fromarray(rng.randint(0, 255, (32, 64, 3), dtype=np.uint8)).save(str(data / 'train' / ('%d.png' % i)))
    Image.fromarray(rng.randint(0, 255, (32, 64, 3), dtype=np.uint8)).save(str(data / 'val' / '0.png'))
    np.savez(str(tmp_path / 'real_stat.npz'), mu=np.zeros(4), sigma=np.eye(4))
    saved_modules = dict(sys.modules)
    saved_path = list(sys.path)
    saved_argv = list(sys.argv)
    try:
        tb = types.ModuleType('tensorboardX')

        class SummaryWriter:
            def __init__(self, *a, **k):
                self.scalars = []

            def add_scalar(self, k, v, global_step=None):
                self.scalars.append((k, float(v), global_step))

            def flush(self):
                pass
        tb.SummaryWriter = SummaryWriter
        sys.modules['tensorboardX'] = tb
        import torchvision  # noqa: F401  (before the reference root, whose profile.py shadows the stdlib module)
        sys.path.append(root)
        with emulated_kernels(exact=True):
            import cat_b200.install
            cat_b200.install.install(init_distributed=False)
            import distillers
            assert distillers.__name__ == 'cat_b200.distillers'
            # a teacher checkpoint in the reference's format (seeded synthetic teacher with spread norm scales)
            from cat_b200 import workload as WL
            t_arch = dict(input_nc=3, output_nc=3, widths=[16, 32, 64, 32, 16], kernel_sizes=[1, 3, 5], norm='instance', affine=True,
                          track_running_stats=False, eps=1e-5, momentum=0.1, use_bias=True,
                          blocks=[dict(res=[10, 10, 10], dw=[10, 10, 10]) for _ in range(9)])
            tpath = str(tmp_path / 'teacher_net_G.pth')
            torch.save(WL.init_generator(t_arch, 0, 'uniform'), tpath)
            argv = ['distill.py', '--dataroot', str(data), '--distiller', 'inception', '--log_dir', str(tmp_path / 'logs'),
                    '--restore_teacher_G_path', tpath, '--real_stat_path', str(tmp_path / 'real_stat.npz'),
                    '--gpu_ids', '-1', '--teacher_ngf', '16', '--student_ngf', '8', '--ndf', '8', '--norm', 'instance',
                    '--norm_affine', '--norm_affine_D', '--channels_reduction_factor', '6', '--kernel_sizes', '1', '3', '5',
                    '--distill_G_loss_type', 'ka', '--lambda_distill', '1', '--lambda_recon', '10', '--target_flops', '1.2e7',
                    '--prune_cin_lb', '4', '--batch_size', '2', '--load_size', '32', '--crop_size', '32', '--nepochs', '1',
                    '--nepochs_decay', '0', '--print_freq', '1', '--save_latest_freq', '1000', '--save_epoch_freq', '1',
                    '--num_threads', '0']
            sys.argv = argv
            monkeypatch.setenv('CATB_CUDA_GRAPH', '0')
            import trainer as ref_trainer          # the reference's trainer.py, as shipped
            assert os.path.samefile(os.path.dirname(ref_trainer.__file__), root)
            tr = ref_trainer.Trainer('distill')
            model = tr.model
            assert type(model).__module__ == 'cat_b200.distillers.inception_distiller'
            tr.opt.cuda_graph = False
            seen = []
            orig = model.optimize_parameters

            def spy(steps):
                orig(steps)
                seen.append((steps, model.engine.B, dict(model.get_current_losses())))
            model.optimize_parameters = spy
            tr.start()
        # ---- shrink produced a student within the budget, the loop ran 3 iterations over 2 batch shapes on one optimiser state
        assert model.netG_student.n_macs <= 1.2e7 and model.netG_student.arch()['widths'][2] < 64
        base = tr.opt.iter_base
        assert [s[0] for s in seen] == [base, base + 1, base + 2] and [s[1] for s in seen] == [2, 2, 1]
        for _, _, L in seen:
            assert all(v == v for v in L.values()) and set(L) >= {'G_loss/G_gan', 'G_loss/G_recon', 'G_loss/G_distill', 'D_loss/D_fake', 'D_loss/D_real'}
        assert int(model.engine.step_G.item()) == 3 and int(model.engine.step_D.item()) == 3
        ck = tmp_path / 'logs' / 'checkpoints'
        for f in ('latest_net_G.pth', 'latest_net_D.pth', 'latest_optim-0.pth', 'latest_optim-1.pth', '1_net_G.pth', 'latest_net_A-0.pth'):
            assert (ck / f).exists(), f
        sd = torch.load(str(ck / '1_net_G.pth'))
        assert set(sd) == set(model.netG_student.state_dict())
        osd = torch.load(str(ck / '1_optim-0.pth'), weights_only=False)
        assert set(osd) == {'state', 'param_groups'} and len(osd['param_groups']) == 2
        tr.logger.log_file.flush()      # Logger.print_info does not flush
        log = (tmp_path / 'logs' / 'log.txt').read_text()
        assert 'G_gan' in log and 'End of epoch 1 / 1' in log
        assert tr.logger.writer.scalars, 'losses were plotted through the reference logger'
    finally:
        sys.argv = saved_argv
        sys.path[:] = saved_path
        for k in list(sys.modules):
            if k not in saved_modules:
                del sys.modules[k]
        sys.modules.update(saved_modules)
