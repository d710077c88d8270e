"""Pin oracle/cat_oracle.py against golden vectors produced by the real reference
(oracle/make_golden.py).  fp32 on both sides, same algorithm -> tight tolerances."""
import os

import pytest
import torch

from oracle import cat_oracle as O

CASES = ['pix2pix_bn_hinge', 'cyclegan_in_lsgan', 'pix2pix_bn_lsgan_l2', 'pix2pix_bn_mse']


def _close(a, b, rtol=2e-4, atol=2e-5, what=''):
    a, b = a.double(), b.double()
    err = (a - b).abs().max().item()
    ref = b.abs().max().item()
    assert err <= atol + rtol * ref, f'{what}: max err {err} vs ref scale {ref}'


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name + '.pt'), weights_only=False)


@pytest.mark.parametrize('name', CASES)
def test_forward_matches_reference(golden_dir, name):
    fix = _load(golden_dir, name)
    s0 = fix['steps'][0]
    cap = {}
    T = O.generator_forward(O.clone_sd(fix['teacher_sd']), fix['teacher_arch'], s0['real_A'], False, cap)
    _close(T, s0['Tfake_B'])
    for k in O.MAPPING_LAYERS:
        _close(cap[k], s0['Tacts'][k])
    cap = {}
    S = O.generator_forward(O.clone_sd(fix['student_sd0']), fix['student_arch'], s0['real_A'], True, cap)
    _close(S, s0['Sfake_B'])
    for k in O.MAPPING_LAYERS:
        _close(cap[k], s0['Sacts'][k])


@pytest.mark.parametrize('name', CASES)
def test_two_distill_steps_match_reference(golden_dir, name):
    fix = _load(golden_dir, name)
    state = dict(teacher_sd=O.clone_sd(fix['teacher_sd']), student_sd=O.clone_sd(fix['student_sd0']),
                 D_sd=O.clone_sd(fix['D_sd0']), teacher_arch=fix['teacher_arch'],
                 student_arch=fix['student_arch'], D_arch=fix['D_arch'], adam_G={}, adam_D={})
    if 'netA_sd0' in fix:       # --distill_G_loss_type mse: the adaptor convs netAs
        state['netA_sds'] = [O.clone_sd(sd) for sd in fix['netA_sd0']]
    for it, step in enumerate(fix['steps']):
        out = O.distill_step(state, step['real_A'], step['real_B'], fix['hp'])
        L = step['losses']
        assert abs(float(out['loss_G_gan']) - L['G_loss/G_gan']) < 1e-4 * max(1, abs(L['G_loss/G_gan']))
        assert abs(float(out['loss_G_recon']) - L['G_loss/G_recon']) < 1e-4 * max(1, abs(L['G_loss/G_recon']))
        assert abs(float(out['loss_G_distill']) - L['G_loss/G_distill']) < 1e-4
        assert abs(float(out['loss_D_fake']) - L['D_loss/D_fake']) < 1e-4 * max(1, abs(L['D_loss/D_fake']))
        assert abs(float(out['loss_D_real']) - L['D_loss/D_real']) < 1e-4 * max(1, abs(L['D_loss/D_real']))
        for i in range(4):
            assert abs(float(out['loss_G_distill_terms'][i]) - L['Specific_loss/G_distill%d' % i]) < 1e-4
        if it == 0:
            # conv biases that feed a normalisation layer have an analytically zero gradient; what
            # the reference stores there is rounding noise, so the floor is tied to the global scale
            dscale = max(float(g.abs().max()) for g in step['D_grads'].values())
            sscale = max(float(g.abs().max()) for g in step['S_grads'].values())
            for k, g in step['D_grads'].items():
                _close(out['D_grads'][k], g, rtol=1e-3, atol=1e-5 * dscale)
            for k, g in step['S_grads'].items():
                _close(out['S_grads'][k], g, rtol=1e-3, atol=1e-5 * sscale)
            # Adam turns a rounding-noise gradient into a +-lr update whose sign is arbitrary, so
            # elements whose reference gradient is below 1e-5 of the global scale are only
            # required to stay inside that band.
            noise_D = {k: g.abs() < 1e-5 * dscale for k, g in step['D_grads'].items()}
            noise_S = {k: g.abs() < 1e-5 * sscale for k, g in step['S_grads'].items()}
            for k, g in step['Sact_grads'].items():
                _close(out['Sact_grads'][k], g, rtol=1e-3, atol=1e-7)
            for i, grads in enumerate(step.get('netA_grads', [])):
                for k, g in grads.items():
                    _close(out['A_grads'][f'A{i}.{k}'], g, rtol=1e-3, atol=1e-7, what=f'netA{i}.{k}')
                for k, v in step['netA_sd_after'][i].items():
                    _close(state['netA_sds'][i][k], v, rtol=1e-3, atol=1e-5, what=f'netA{i}.{k} after Adam')
        lr = fix['hp']['lr']
        for sd_key, noise, mine in (('student_sd_after', noise_S, state['student_sd']),
                                   ('D_sd_after', noise_D, state['D_sd'])):
            for k, v in step[sd_key].items():
                if not v.is_floating_point():
                    continue
                if k in noise:
                    m = noise[k]
                    if (~m).any():
                        _close(mine[k][~m], v[~m], rtol=1e-3, atol=1e-5, what=k)
                    if m.any():
                        _close(mine[k][m], v[m], rtol=0, atol=2.1 * lr * (it + 1), what=k + '(noise)')
                else:
                    _close(mine[k], v, rtol=1e-3, atol=1e-5, what=k)


def test_ka_analytic_gradient_matches_autograd():
    torch.manual_seed(0)
    X = torch.randn(5, 7, 4, 4, dtype=torch.float64, requires_grad=True)
    Y = torch.randn(5, 11, 4, 4, dtype=torch.float64)
    v = O.ka(X, Y)
    v.backward()
    _close(O.ka_grad_x(X.detach(), Y), X.grad, rtol=1e-9, atol=1e-12)
    # degenerate batch of one: KA == 1 and the gradient vanishes
    X1 = torch.randn(1, 3, 4, 4, dtype=torch.float64)
    assert abs(float(O.ka(X1, torch.randn(1, 5, 4, 4, dtype=torch.float64))) - 1) < 1e-12
    assert O.ka_grad_x(X1, torch.randn(1, 5, 4, 4, dtype=torch.float64)).abs().max() < 1e-12


def _first_step_state(golden_dir):
    fix = _load(golden_dir, 'pix2pix_bn_lsgan_l2')
    add = _load(golden_dir, 'pix2pix_bn_lsgan_l2_first_step')
    student = O.clone_sd(fix['student_sd0'])
    student.update({k: v.clone() for k, v in add['running_stats'].items()})
    state = dict(teacher_sd=O.clone_sd(fix['teacher_sd']), student_sd=student, D_sd=O.clone_sd(fix['D_sd0']),
                 teacher_arch=fix['teacher_arch'], student_arch=fix['student_arch'], D_arch=fix['D_arch'], adam_G={}, adam_D={})
    return fix, add, state, dict(fix['hp'], student_training=False)


def test_first_step_with_the_student_in_eval_mode_matches_reference(golden_dir):
    """The reference's first optimize_parameters of a run: the pruned student is still in eval() (left there by
    model_profiling, utils/model_profiling.py:299; back to train() only at the end of the first evaluate_model,
    inception_distiller.py:280), so BatchNorm uses -- and is differentiated through -- its running statistics."""
    fix, add, state, hp = _first_step_state(golden_dir)
    s = fix['steps'][0]
    out = O.distill_step(state, s['real_A'], s['real_B'], hp)
    _close(out['Sfake_B'], add['Sfake_B'])
    for k_ref, k in (('loss_G_gan', 'G_loss/G_gan'), ('loss_G_recon', 'G_loss/G_recon'), ('loss_G_distill', 'G_loss/G_distill'),
                     ('loss_D_fake', 'D_loss/D_fake'), ('loss_D_real', 'D_loss/D_real')):
        r = add['losses'][k]
        assert abs(float(out[k_ref]) - r) < 1e-4 * max(1.0, abs(r)), (k, float(out[k_ref]), r)
    scale = max(float(g.abs().max()) for g in add['S_grads'].values())
    for k, g in add['S_grads'].items():
        _close(out['S_grads'][k], g, rtol=1e-3, atol=1e-5 * scale, what=k)
    for k, v in add['running_stats_after'].items():          # eval mode: the running statistics do not move
        assert torch.equal(state['student_sd'][k], add['running_stats'][k]) and torch.equal(v, add['running_stats'][k]), k
