"""End-to-end GPU parity of the SPADE distillation step (cat_b200.spade_distill_engine.SpadeDistillStep through
libcatb200.so) against oracle/spade_oracle.py, which tests/test_spade_oracle_golden.py pins to the real reference.

Stated tolerances (bf16 storage and tensor-core operands, fp32 accumulation / statistics / losses):
  one-hot + edge preprocessing bit-exact; teacher output and mapped activations <= 3e-2 relative L2; losses
  <= 3e-2 * max(1, |ref|) against the fp32 oracle and the oracle with bf16 storage emulated at the same points; KA
  terms <= 5e-3; parameter gradients (relative L2 over all parameters of a network) <= 0.5 -- on this fixture the
  bf16-emulating oracle itself is 5 % (student) / 18 % (discriminator) away from the fp32 oracle because hinge masks,
  ReLU masks and sign(L1) flip under 1 % forward differences, so the bound only excludes real defects; the exact
  (fp32-emulated) CPU test tests/test_spade_engine_emulated_cpu.py carries the tight bound on the same host logic."""
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


LOSSES = (('loss_D_fake', 'D_fake'), ('loss_D_real', 'D_real'), ('loss_G_gan', 'G_gan'), ('loss_G_feat', 'G_feat'),
          ('loss_G_vgg', 'G_vgg'), ('loss_G_distill', 'G_distill'))


@pytest.mark.parametrize('use_graph', [False, True])
def test_spade_step_matches_oracle(golden_dir, use_graph):
    from cat_b200 import ops
    from cat_b200.spade_distill_engine import SpadeDistillStep
    from oracle import cat_oracle as O
    from oracle import spade_oracle as SO
    ops.require_cuda()
    fix = torch.load(os.path.join(golden_dir, 'spade_more.pt'), weights_only=False)
    vgg = SO.make_vgg_sd(fix['vgg_seed'])
    hp = fix['hp']
    B, _, H, W = fix['steps'][0]['image'].shape
    eng = SpadeDistillStep(fix['teacher_arch'], fix['student_arch'], fix['D_arch'], hp, B, H, W, use_cuda_graph=use_graph)
    eng.load(fix['teacher_sd'], fix['student_sd0'], fix['D_sd0'], vgg)

    def state():
        return dict(teacher_sd=O.clone_sd(fix['teacher_sd']), student_sd=O.clone_sd(fix['student_sd0']),
                    D_sd=O.clone_sd(fix['D_sd0']), vgg_sd=vgg, teacher_arch=fix['teacher_arch'],
                    student_arch=fix['student_arch'], D_arch=fix['D_arch'], adam_G={}, adam_D={})
    st32, stq = state(), state()
    for it, s in enumerate(fix['steps']):
        seg = SO.preprocess_input(s['label'], s['instance'], hp['n_label'])
        ref32 = SO.spade_distill_step(st32, seg, s['image'], hp)
        with O.emulate_bf16():
            refq = SO.spade_distill_step(stq, seg, s['image'], hp)
        eng.set_input(s['label'], s['instance'], s['image'])
        eng.step()
        torch.cuda.synchronize()
        L = eng.get_losses()
        assert torch.equal(ops.nhwc_to_nchw(eng.seg, eng.snc).cpu(), seg)
        rep = {}
        for tag, ref in (('fp32', ref32), ('emu', refq)):
            rep[tag + ' T'] = rel_l2(ops.nhwc_to_nchw(eng.T.out, 3).cpu(), ref['Tfake_B'])
            rep[tag + ' S(D phase)'] = rel_l2(ops.nhwc_to_nchw(eng.S.out, 3).cpu(), ref['Sfake_B_D'])
            for n, t in ref['Tacts'].items():
                rep[f'{tag} Tact {n}'] = rel_l2(ops.nhwc_to_nchw(eng.T.acts[n], t.shape[1]).cpu(), t)
            if it == 0:
                for nt, net, key in (('S', eng.S, 'S_grads'), ('D', eng.D, 'D_grads')):
                    ks = [k for k in ref[key] if float(ref32[key][k].abs().max()) > 1e-6]
                    mine = torch.cat([net.arena.view(k, 'g').flatten().cpu() for k in ks])
                    rep[f'{tag} {nt}_grads'] = rel_l2(mine, torch.cat([ref[key][k].flatten() for k in ks]))
        print('spade', 'graph' if use_graph else 'eager', 'step', it, {k: round(v, 4) for k, v in rep.items()}, L)
        for k, v in rep.items():
            if k.endswith('_grads'):
                assert v <= 0.5, (it, k, v)
            elif 'S(D phase)' in k:
                assert v <= (8e-2 if it else 5e-2), (it, k, v)    # after an Adam (beta1 = 0) step of +-lr per weight
            else:
                assert v <= 3e-2, (it, k, v)
        for ref in (ref32, refq):
            for k_ref, k in LOSSES:
                r = float(ref[k_ref])
                assert abs(L[k] - r) <= (6e-2 if it else 3e-2) * max(1.0, abs(r)), (it, k, L[k], r)
            for i in range(3):
                assert abs(L['G_distill%d' % i] - float(ref['loss_G_distill_terms'][i])) <= (2e-2 if it else 5e-3), (it, i)
        if it == 0:
            sd, dsd = eng.S.state_dict(), eng.D.state_dict()
            for k, v in stq['student_sd'].items():
                if 'running_' in k:
                    assert float((sd[k] - v).abs().max()) <= 2e-2 * max(1.0, float(v.abs().max())), k
            for k, v in stq['D_sd'].items():
                if k.endswith('weight_u') or k.endswith('weight_v'):
                    assert float((dsd[k] - v).abs().max()) <= 1e-3, k
