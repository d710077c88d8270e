"""Training state across batch shapes and through optimizer checkpoints (the mirrors used to rebuild their
engine -- and restart Adam and the learning-rate schedule -- whenever the batch shape changed, i.e. at the partial batch
that ends every epoch; the reference DataLoader has no drop_last, data/__init__.py:82).  Exact kernel emulation on CPU
against the oracle, which keeps one optimiser state over steps of any batch size."""
import importlib
import os

import pytest
import torch


def _batch(s, n):
    return {'A': s['real_A'][:n], 'B': s['real_B'][:n], 'A_paths': ['a/%d.png' % i for i in range(n)], 'B_paths': ['x'] * n}


KEYS = (('G_loss/G_gan', 'loss_G_gan'), ('G_loss/G_recon', 'loss_G_recon'), ('G_loss/G_distill', 'loss_G_distill'),
        ('D_loss/D_fake', 'loss_D_fake'), ('D_loss/D_real', 'loss_D_real'))


@pytest.mark.timeout(900)
def test_alternating_batch_shapes_keep_adam_state_and_lr(golden_dir, tmp_path):
    from oracle import cat_oracle as O
    from oracle.kernel_emu import emulated_kernels
    _opt = importlib.import_module('test_distiller_flow_emulated_cpu')._opt
    fix = torch.load(os.path.join(golden_dir, 'pix2pix_bn_hinge.pt'), weights_only=False)
    st = dict(teacher_sd=O.clone_sd(fix['teacher_sd']), student_sd=O.clone_sd(fix['student_sd0']), D_sd=O.clone_sd(fix['D_sd0']),
              teacher_arch=fix['teacher_arch'], student_arch=fix['student_arch'], D_arch=fix['D_arch'], adam_G={}, adam_D={})
    s0, s1 = fix['steps']
    B = s0['real_A'].shape[0]
    with emulated_kernels(exact=True):
        from cat_b200.distillers import create_distiller
        opt = _opt(fix, str(tmp_path))
        opt.nepochs, opt.nepochs_decay = 1, 3           # the second update_learning_rate() already decays
        model = create_distiller(opt, verbose=False)
        model.setup(opt, verbose=False)
        model.netG_teacher.load_state_dict(fix['teacher_sd'])
        model.netG_student.load_state_dict(fix['student_sd0'])
        model.netD.load_state_dict(fix['D_sd0'])
        model.netG_student.train()
        lr = fix['hp']['lr']
        engines = []
        # full batch, partial batch (end of the epoch), lr decay, full batch, partial batch
        for it, (s, n) in enumerate(((s0, B), (s1, B - 1), (s1, B), (s0, B - 1))):
            if it == 2:
                model.update_learning_rate()
                model.update_learning_rate()
                lr = model.optimizer_G.param_groups[0]['lr']
                assert lr < fix['hp']['lr']
            ref = O.distill_step(st, s['real_A'][:n], s['real_B'][:n], dict(fix['hp'], lr=lr))
            model.set_input(_batch(s, n))
            model.optimize_parameters(it)
            engines.append(model.engine)
            L = model.get_current_losses()
            for mine, theirs in KEYS:
                r = float(ref[theirs])
                assert abs(L[mine] - r) <= 2e-4 * max(1.0, abs(r)), (it, mine, L[mine], r)
            assert int(model.engine.step_G.item()) == it + 1 and int(model.engine.step_D.item()) == it + 1
        assert engines[0] is engines[2] and engines[1] is engines[3] and engines[0] is not engines[1]     # cached per shape
        # post-step weights follow the oracle's (a restarted Adam would have moved them by ~lr per step)
        mine = model.netG_student.state_dict()
        worst = max(float((mine[k] - v).abs().max()) for k, v in st['student_sd'].items() if v.is_floating_point() and v.dim() == 4)
        assert worst < 2e-5, worst


@pytest.mark.timeout(900)
def test_optimizer_checkpoints_are_torch_adam_state_dicts(golden_dir, tmp_path):
    """save_networks writes <epoch>_optim-<i>.pth in the layout of torch.optim.Adam.state_dict()
    (base_inception_distiller.py:393-396); a real torch Adam over the same parameter lists loads it, its state equals the
    oracle's Adam state, and --restore_O_path of a fresh distiller continues the run exactly."""
    from oracle import cat_oracle as O
    from oracle.kernel_emu import emulated_kernels
    _opt = importlib.import_module('test_distiller_flow_emulated_cpu')._opt
    fix = torch.load(os.path.join(golden_dir, 'pix2pix_bn_hinge.pt'), weights_only=False)
    st = dict(teacher_sd=O.clone_sd(fix['teacher_sd']), student_sd=O.clone_sd(fix['student_sd0']), D_sd=O.clone_sd(fix['D_sd0']),
              teacher_arch=fix['teacher_arch'], student_arch=fix['student_arch'], D_arch=fix['D_arch'], adam_G={}, adam_D={})
    s0, s1 = fix['steps']
    B = s0['real_A'].shape[0]
    with emulated_kernels(exact=True):
        from cat_b200.distillers import create_distiller

        def make(opt):
            model = create_distiller(opt, verbose=False)
            model.setup(opt, verbose=False)
            model.netG_student.train()
            return model
        opt = _opt(fix, str(tmp_path))
        tpath = str(tmp_path / 'teacher.pth')
        torch.save(fix['teacher_sd'], tpath)
        opt.restore_teacher_G_path = tpath
        model = make(opt)
        model.netG_student.load_state_dict(fix['student_sd0'])
        model.netD.load_state_dict(fix['D_sd0'])
        O.distill_step(st, s0['real_A'], s0['real_B'], fix['hp'])
        model.set_input(_batch(s0, B))
        model.optimize_parameters(0)
        model.save_networks('latest')
        ck = os.path.join(str(tmp_path), 'checkpoints')
        # ---- a real torch.optim.Adam (the reference's optimizer objects) accepts the files
        sdG = torch.load(os.path.join(ck, 'latest_optim-0.pth'), weights_only=False)
        sdD = torch.load(os.path.join(ck, 'latest_optim-1.pth'), weights_only=False)
        assert set(sdG) == {'state', 'param_groups'} and len(sdG['param_groups']) == 2 and len(sdD['param_groups']) == 1
        gp = [torch.nn.Parameter(p.detach().clone()) for p in model.netG_student.parameters()]
        ap = [torch.nn.Parameter(p.detach().clone()) for net in model.netAs for p in net.parameters()]
        adam = torch.optim.Adam([{'params': gp}, {'params': ap}], lr=opt.lr, betas=(opt.beta1, 0.999))
        adam.load_state_dict(sdG)
        names = [n for n, _ in model.netG_student.named_parameters()]
        pairs = {'exp_avg': ([], []), 'exp_avg_sq': ([], [])}
        for i, n in enumerate(names):
            ref = st['adam_G'].get(n)
            got = adam.state.get(gp[i])
            if ref is None:
                continue
            assert got is not None, n
            assert float(got['step']) == 1.0 and got['exp_avg'].shape == ref['m'].shape
            for key, rk in (('exp_avg', 'm'), ('exp_avg_sq', 'v')):
                # per tensor: loose (tiny norm-scale gradients are sums with cancellation), over all tensors: tight
                assert float((got[key] - ref[rk]).norm()) <= 5e-2 * float(ref[rk].norm()) + 1e-12, (n, key)
                pairs[key][0].append(got[key].flatten())
                pairs[key][1].append(ref[rk].flatten())
        assert len(pairs['exp_avg'][0]) > 20
        for key, (a, b) in pairs.items():
            a, b = torch.cat(a), torch.cat(b)
            assert float((a - b).norm() / b.norm()) < 2e-3, key
        dp = [torch.nn.Parameter(p.detach().clone()) for p in model.netD.parameters()]
        torch.optim.Adam(dp, lr=opt.lr, betas=(opt.beta1, 0.999)).load_state_dict(sdD)
        # ---- a fresh distiller restored from the files (incl. --restore_O_path) takes the same second step
        ref = O.distill_step(st, s1['real_A'], s1['real_B'], fix['hp'])
        opt2 = _opt(fix, str(tmp_path / 'run2'))
        opt2.restore_teacher_G_path = tpath
        opt2.restore_student_G_path = os.path.join(ck, 'latest_net_G.pth')
        opt2.restore_D_path = os.path.join(ck, 'latest_net_D.pth')
        opt2.restore_A_path = os.path.join(ck, 'latest_net_A')
        opt2.restore_O_path = os.path.join(ck, 'latest_optim')
        model2 = make(opt2)
        model2.set_input(_batch(s1, B))
        assert int(model2.engine.step_G.item()) == 1 and int(model2.engine.step_D.item()) == 1
        model2.optimize_parameters(1)
        L = model2.get_current_losses()
        for mine, theirs in KEYS:
            r = float(ref[theirs])
            assert abs(L[mine] - r) <= 2e-4 * max(1.0, abs(r)), (mine, L[mine], r)
        mine = model2.netG_student.state_dict()
        worst = max(float((mine[k] - v).abs().max()) for k, v in st['student_sd'].items() if v.is_floating_point() and v.dim() == 4)
        assert worst < 2e-5, worst
