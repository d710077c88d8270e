"""Pin oracle/train_oracle.py (teacher-training steps) against golden vectors written by the real
reference Pix2PixModel / CycleGANModel / SPADEModel (oracle/make_golden_train.py).  fp32 on both sides, same
algorithm -> tight tolerances."""
import os
import random

import pytest
import torch

from oracle import spade_oracle as SO
from oracle import train_oracle as TO
from oracle.cat_oracle import clone_sd


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name + '.pt'), weights_only=False)


def _check_grads(mine, theirs, what):
    """Biases in front of a normalisation layer have an analytically zero gradient (the reference holds rounding
    noise there): the floor is tied to the global gradient scale."""
    scale = max(float(g.abs().max()) for g in theirs.values())
    assert set(mine) == set(theirs), what
    for k, g in theirs.items():
        err = float((mine[k] - g).abs().max())
        assert err <= 1e-3 * float(g.abs().max()) + 1e-5 * scale, (what, k, err)


def _check_losses(out, losses, it, rtol=1e-4):
    for k, r in losses.items():
        v = float(out['loss_' + k])
        assert abs(v - r) < rtol * max(1.0, abs(r)), (it, k, v, r)


def _checksum(sd, is_param):
    return float(sum(v.double().abs().sum() for k, v in sd.items() if is_param(k)))


@pytest.mark.parametrize('name', ['train_pix2pix_bn_hinge', 'train_pix2pix_in_lsgan_l2'])
def test_pix2pix_train_steps_match_reference(golden_dir, name):
    fix = _load(golden_dir, name)
    state = dict(G_sd=clone_sd(fix['G_sd0']), D_sd=clone_sd(fix['D_sd0']), G_arch=fix['G_arch'], D_arch=fix['D_arch'],
                 adam_G={}, adam_D={})
    for it, s in enumerate(fix['steps']):
        out = TO.pix2pix_train_step(state, s['real_A'], s['real_B'], fix['hp'])
        _check_losses(out, s['losses'], it)
        if it == 0:
            assert float((out['fake_B'] - s['fake_B']).abs().max()) < 1e-5
            _check_grads(out['G_grads'], s['G_grads'], 'G')
            _check_grads(out['D_grads'], s['D_grads'], 'D')
            for k, v in s['G_sd_after'].items():     # running statistics after one training forward
                if 'running' in k:
                    assert float((state['G_sd'][k] - v).abs().max()) < 1e-5, k
        # Adam turns the noise gradients of the inert biases into +-lr steps of arbitrary sign: checksum band
        n_G = sum(v.numel() for k, v in state['G_sd'].items() if TO.O._is_param(k))
        assert abs(_checksum(state['G_sd'], TO.O._is_param) - s['G_checksum_after']) < 2e-4 * (it + 1) * n_G ** 0.5 + 1e-2
        assert abs(_checksum(state['D_sd'], TO.O._is_param) - s['D_checksum_after']) < 1e-2


@pytest.mark.parametrize('name', ['train_cyclegan_in_lsgan', 'train_cyclegan_bn_lsgan'])
def test_cyclegan_train_steps_match_reference(golden_dir, name):
    fix = _load(golden_dir, name)
    hp = fix['hp']
    state = dict(G_A_sd=clone_sd(fix['G_A_sd0']), G_B_sd=clone_sd(fix['G_B_sd0']), D_A_sd=clone_sd(fix['D_A_sd0']),
                 D_B_sd=clone_sd(fix['D_B_sd0']), G_arch=fix['G_arch'], D_arch=fix['D_arch'], adam_G={}, adam_D={},
                 pool_A=TO.ImagePool(hp['pool_size']), pool_B=TO.ImagePool(hp['pool_size']))
    random.seed(fix['python_random_seed'])
    for it, s in enumerate(fix['steps']):
        out = TO.cyclegan_train_step(state, s['real_A'], s['real_B'], hp)
        # from step 2 on (pool of 3) this also pins the history decisions (a wrong image moves D_A / D_B by O(1)).  The
        # fp32 trajectories separate slowly (Adam amplifies rounding-level gradients to lr-sized steps): measured 1e-6
        # for steps 0-2, 7e-5 at step 3, 6e-4 at step 4
        _check_losses(out, s['losses'], it, 1e-4 if it < 3 else 3e-3)
        if it == 0:
            for k in ('fake_A', 'fake_B', 'rec_A', 'rec_B'):
                assert float((out[k] - s[k]).abs().max()) < 1e-4, k   # two chained generators
            for k in ('G_A', 'G_B', 'D_A', 'D_B'):
                _check_grads(out[k + '_grads'], s[k + '_grads'], k)
            for k, v in s['G_A_buffers_after'].items():   # three training forwards of G_A, in the reference's order
                if v.is_floating_point():
                    assert float((state['G_A_sd'][k] - v).abs().max()) < 1e-5, k
        for k in ('D_A', 'D_B'):
            assert abs(_checksum(state[k + '_sd'], TO.O._is_param) - s['checksums_after'][k]) < 1e-2, k


def test_spade_train_steps_match_reference(golden_dir):
    fix = _load(golden_dir, 'train_spade_more')
    assert fix['G_arch']['active_fn'] == 'nn.LeakyReLU'      # SPADEModel's default (models/spade_model.py:92)
    vgg = SO.make_vgg_sd(fix['vgg_seed'])
    check = float(sum(v.double().abs().sum() for v in vgg.values()))
    assert abs(check - fix['vgg_check']) < 1e-6 * fix['vgg_check']
    hp = fix['hp']
    state = dict(G_sd=clone_sd(fix['G_sd0']), D_sd=clone_sd(fix['D_sd0']), vgg_sd=vgg, G_arch=fix['G_arch'],
                 D_arch=fix['D_arch'], adam_G={}, adam_D={})
    for it, s in enumerate(fix['steps']):
        seg = SO.preprocess_input(s['label'], s['instance'], hp['n_label'])
        assert torch.equal(seg, s['seg'].float())
        if it == 0:
            # Gradients are compared through the oracle evaluated in fp64: the discriminator phase runs on the generator
            # AFTER its Adam step, whose rounding-level differences flip a few LeakyReLU kinks of the first two
            # full-resolution D layers in fp32 (measured: fp32 oracle 1e-4 off there, fp64 oracle 1e-7 off everywhere).
            st64 = dict(state, G_sd=clone_sd(fix['G_sd0'], torch.float64), D_sd=clone_sd(fix['D_sd0'], torch.float64),
                        vgg_sd={k: v.double() for k, v in vgg.items()}, adam_G={}, adam_D={})
            out64 = TO.spade_train_step(st64, seg.double(), s['image'].double(), hp)
            _check_grads(out64['G_grads'], s['G_grads'], 'G')
            _check_grads(out64['D_grads'], s['D_grads'], 'D')
        out = TO.spade_train_step(state, seg, s['image'], hp)
        # second step: both networks have taken one beta1 = 0 Adam step (+-lr per element, the sign of rounding-level
        # gradients is arbitrary), measured 9e-4 on G_feat
        _check_losses(out, s['losses'], it, 1e-4 if it == 0 else 3e-3)
        if it:
            continue
        _check_grads(out['G_grads'], s['G_grads'], 'G')
        for k, v in s['G_buffers_after'].items():
            if v.is_floating_point():
                assert float((state['G_sd'][k] - v).abs().max()) < 1e-4, k
        for k, v in s['D_buffers_after'].items():
            assert float((state['D_sd'][k] - v).abs().max()) < 1e-4, k
        # first Adam step with beta1 = 0 moves every element by +-lr_D = 4e-4 (sign of the gradient): a handful of
        # rounding-level gradient signs differ
        assert abs(_checksum(state['D_sd'], SO._is_param) - s['D_checksum_after']) < 1e-2
