"""Pin oracle/spade_oracle.py against golden vectors produced by the real reference SPADEDistiller
(oracle/make_golden_spade.py).  fp32 on both sides, same algorithm -> tight tolerances."""
import os

import torch

from oracle import spade_oracle as SO
from oracle.cat_oracle import clone_sd

LOSS_KEYS = [('loss_G_gan', 'G_loss/G_gan'), ('loss_G_feat', 'G_loss/G_feat'), ('loss_G_vgg', 'G_loss/G_vgg'),
             ('loss_G_distill', 'G_loss/G_distill'), ('loss_D_real', 'D_loss/D_real'), ('loss_D_fake', 'D_loss/D_fake')]


def spade_state(fix):
    vgg = SO.make_vgg_sd(fix['vgg_seed'])
    check = float(sum(v.double().abs().sum() for v in vgg.values()))
    assert abs(check - fix['vgg_check']) < 1e-6 * fix['vgg_check'], 'seeded VGG19 weights differ from the fixture\'s'
    return dict(teacher_sd=clone_sd(fix['teacher_sd']), student_sd=clone_sd(fix['student_sd0']), D_sd=clone_sd(fix['D_sd0']),
                vgg_sd=vgg, teacher_arch=fix['teacher_arch'], student_arch=fix['student_arch'], D_arch=fix['D_arch'],
                adam_G={}, adam_D={})


def test_two_spade_steps_match_reference(golden_dir):
    fix = torch.load(os.path.join(golden_dir, 'spade_more.pt'), weights_only=False)
    state, hp = spade_state(fix), fix['hp']
    for it, s in enumerate(fix['steps']):
        seg = SO.preprocess_input(s['label'], s['instance'], hp['n_label'])
        assert torch.equal(seg, s['seg'].float())          # one-hot + edges: bit exact
        out = SO.spade_distill_step(state, seg, s['image'], hp)
        for mine, theirs in LOSS_KEYS:
            r = s['losses'][theirs]
            assert abs(float(out[mine]) - r) < 1e-4 * max(1.0, abs(r)), (it, mine, float(out[mine]), r)
        for i in range(3):
            assert abs(float(out['loss_G_distill_terms'][i]) - s['losses']['Specific_loss/G_distill%d' % i]) < 1e-4
        if it:
            continue
        # biases in front of a BatchNorm have an analytically zero gradient (the reference holds rounding noise
        # there): the floor is tied to the global gradient scale
        for name in ('S_grads', 'D_grads'):
            scale = max(float(g.abs().max()) for g in s[name].values())
            assert set(out[name]) == set(s[name])
            for k, g in s[name].items():
                err = float((out[name][k] - g).abs().max())
                assert err <= 1e-3 * float(g.abs().max()) + 1e-5 * scale, (name, k, err)
        for k, v in s['student_buffers_after'].items():
            if v.is_floating_point():
                assert float((state['student_sd'][k] - v).abs().max()) < 1e-4, k
        for k, v in s['D_buffers_after'].items():     # spectral-norm u / v after two power iterations
            assert float((state['D_sd'][k] - v).abs().max()) < 1e-4, k
        # Adam (beta1 = 0) turns the noise gradients of the inert biases into +-lr-sized steps of arbitrary sign,
        # so the parameter checksum is only comparable to that band
        cs = float(sum(v.double().abs().sum() for k, v in state['D_sd'].items() if SO._is_param(k)))
        assert abs(cs - s['D_param_checksum_after']) < 1e-3
        cs = float(sum(v.double().abs().sum() for k, v in state['student_sd'].items() if SO._is_param(k)))
        assert abs(cs - s['student_param_checksum_after']) < 1e-2


def test_spade_mse_distill_step_matches_reference(golden_dir):
    """--distill_G_loss_type mse (spade_distiller_modules.py:23-25) on the networks of spade_more.pt: losses, the gradients
    of the adaptor convs netAs and of the student, and the adaptors after optimizer_G.step, against the real SPADEDistiller
    (oracle/make_golden_spade.py spade_more_mse)."""
    fix = torch.load(os.path.join(golden_dir, 'spade_more.pt'), weights_only=False)
    add = torch.load(os.path.join(golden_dir, 'spade_more_mse.pt'), weights_only=False)
    state = spade_state(fix)
    state['netA_sds'] = [clone_sd(sd) for sd in add['netA_sd0']]
    hp = dict(fix['hp'], distill_loss_type='mse', lambda_distill=add['lambda_distill'])
    s = fix['steps'][0]
    seg = SO.preprocess_input(s['label'], s['instance'], hp['n_label'])
    out = SO.spade_distill_step(state, seg, s['image'], hp)
    for mine, theirs in LOSS_KEYS:
        r = add['losses'][theirs]
        assert abs(float(out[mine]) - r) < 1e-4 * max(1.0, abs(r)), (mine, float(out[mine]), r)
    for i in range(3):
        assert abs(float(out['loss_G_distill_terms'][i]) - add['losses']['Specific_loss/G_distill%d' % i]) < 1e-5
    for i, grads in enumerate(add['netA_grads']):
        for k, g in grads.items():
            err = float((out['A_grads'][f'A{i}.{k}'] - g).abs().max())
            assert err <= 1e-3 * float(g.abs().max()) + 1e-8, (i, k, err)
        for k, v in add['netA_sd_after'][i].items():
            assert float((state['netA_sds'][i][k] - v).abs().max()) < 1e-5, (i, k)
    scale = max(float(g.abs().max()) for g in add['S_grads'].values())
    for k, g in add['S_grads'].items():
        err = float((out['S_grads'][k] - g).abs().max())
        assert err <= 1e-3 * float(g.abs().max()) + 1e-5 * scale, (k, err)


def first_step_state(golden_dir):
    fix = torch.load(os.path.join(golden_dir, 'spade_more.pt'), weights_only=False)
    add = torch.load(os.path.join(golden_dir, 'spade_more_first_step.pt'), weights_only=False)
    state = spade_state(fix)
    state['student_sd'].update({k: v.clone() for k, v in add['running_stats'].items()})
    return fix, add, state, dict(fix['hp'], student_training=False)


def test_spade_first_step_with_the_student_in_eval_mode_matches_reference(golden_dir):
    """The reference's first optimize_parameters of a run (student still in eval(): BaseSPADEDistiller.setup profiles it,
    base_spade_distiller.py:178-190; back to train() at the end of the first evaluate_model, spade_distiller.py:170)."""
    fix, add, state, hp = first_step_state(golden_dir)
    s = fix['steps'][0]
    seg = SO.preprocess_input(s['label'], s['instance'], hp['n_label'])
    out = SO.spade_distill_step(state, seg, s['image'], hp)
    for mine, theirs in LOSS_KEYS:
        r = add['losses'][theirs]
        assert abs(float(out[mine]) - r) < 1e-4 * max(1.0, abs(r)), (mine, float(out[mine]), r)
    scale = max(float(g.abs().max()) for g in add['S_grads'].values())
    assert set(out['S_grads']) == set(add['S_grads'])
    for k, g in add['S_grads'].items():
        err = float((out['S_grads'][k] - g).abs().max())
        assert err <= 1e-3 * float(g.abs().max()) + 1e-5 * scale, (k, err)
    for k, v in add['running_stats_after'].items():           # eval mode: the running statistics do not move
        assert torch.equal(v, add['running_stats'][k]) and torch.equal(state['student_sd'][k], v), k
