"""End-to-end parity of the CUDA distillation step against the CPU oracle (itself pinned to the
reference by tests/test_oracle_golden.py) on the committed golden fixtures.

Stated tolerances (bf16 operands / activations, fp32 accumulation, networks ~40 convs deep):
  * activations / outputs: relative L2 error <= 3e-2
  * losses: |delta| <= 2e-2 * max(1, |loss|);  KA terms: |delta| <= 5e-3
  * parameter gradients, relative L2 error over all parameters of a network:
      - <= 6e-2 when the loss gradients are smooth functions of the forward values (the l2/lsgan
        fixture) or when the oracle's loss gradients d(loss)/d(pred), d(loss)/d(Sfake) are injected;
      - <= 0.5 otherwise: sign(S - B) of the L1 loss and the hinge mask flip wherever the ~1% forward
        rounding difference crosses zero / the margin; flipping a fraction f of the signs changes the
        gradient by sqrt(4 f) in relative L2 (f = 1% -> 20%).  This is conditioning of the loss, not of
        the kernels, and it is why the injected variant exists.
  * post-Adam weights (smooth / injected runs): within 2.1*lr*(step+1) of the oracle (Adam normalises
    every update to ~lr, so a sign flip of a near-zero gradient moves a weight by up to 2*lr) and mean
    |delta| <= 0.15*lr*(step+1)
"""
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]

CASES = ['pix2pix_bn_lsgan_l2', 'pix2pix_bn_hinge', 'cyclegan_in_lsgan']
SMOOTH = {'pix2pix_bn_lsgan_l2'}


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _inert_biases(sd, tag):
    """Conv biases that feed a normalisation layer (analytically zero gradient; not updated by the
    engine, random +-lr walk in the reference)."""
    conv_biases = [k for k in sd if k.endswith('.bias') and (k[:-5] + '.weight') in sd and sd[k[:-5] + '.weight'].dim() == 4]
    if tag == 'S':
        return {k for k in conv_biases if k != 'up_sampling.7.bias'}
    last = max(int(k.split('.')[1]) for k in conv_biases)
    return {k for k in conv_biases if k not in ('model.0.bias', 'model.%d.bias' % last)}


def _run(golden_dir, name, use_graph, inject):
    from cat_b200 import ops
    from cat_b200.distill_engine import DistillStep
    from oracle import cat_oracle as O
    fix = torch.load(os.path.join(golden_dir, name + '.pt'), weights_only=False)
    s0 = fix['steps'][0]
    B, _, H, W = s0['real_A'].shape
    eng = DistillStep(fix['teacher_arch'], fix['student_arch'], fix['D_arch'], fix['hp'], B, H, W,
                      use_cuda_graph=use_graph)
    eng.load(fix['teacher_sd'], fix['student_sd0'], fix['D_sd0'])
    state = dict(teacher_sd=O.clone_sd(fix['teacher_sd']), student_sd=O.clone_sd(fix['student_sd0']),
                 D_sd=O.clone_sd(fix['D_sd0']), teacher_arch=fix['teacher_arch'], student_arch=fix['student_arch'],
                 D_arch=fix['D_arch'], adam_G={}, adam_D={})
    report = {}
    for it, step in enumerate(fix['steps']):
        ref = O.distill_step(state, step['real_A'], step['real_B'], fix['hp'])
        if inject:
            def put_pred(key):
                def fn(act, key=key):
                    act.t.zero_()
                    act.t[..., 0] = ref[key][:, 0].to(act.t.device, torch.bfloat16)
                return fn

            def put_dS(act):
                act.t.zero_()
                act.t[..., :3] = ref['Sfake_grad'].permute(0, 2, 3, 1).to(act.t.device, torch.bfloat16)
            eng.debug_hooks = {'dpred_fake': put_pred('dpred_fake'), 'dpred_real': put_pred('dpred_real'), 'dS': put_dS}
        eng.set_input(step['real_A'], step['real_B'])
        eng.step()
        torch.cuda.synchronize()
        L = eng.get_losses()
        if it == 0:
            report['Tfake'] = rel_l2(ops.nhwc_to_nchw(eng.T.out, 3).cpu(), ref['Tfake_B'])
            report['Sfake'] = rel_l2(ops.nhwc_to_nchw(eng.S.out, 3).cpu(), ref['Sfake_B'])
            for n in O.MAPPING_LAYERS:
                Ct, Cs = ref['Tacts'][n].shape[1], ref['Sacts'][n].shape[1]
                report['Tact ' + n] = rel_l2(ops.nhwc_to_nchw(eng.T.acts[n], Ct).cpu(), ref['Tacts'][n])
                report['Sact ' + n] = rel_l2(ops.nhwc_to_nchw(eng.S.acts[n], Cs).cpu(), ref['Sacts'][n])
            for tag, net, grads in (('S', eng.S, ref['S_grads']), ('D', eng.D, ref['D_grads'])):
                mine, theirs = [], []
                for k, g in grads.items():
                    if net.arena.has(k) and k not in _inert_biases(grads, tag):
                        mine.append(net.arena.view(k, 'g').flatten().cpu())
                        theirs.append(g.flatten())
                report[tag + '_grads'] = rel_l2(torch.cat(mine), torch.cat(theirs))
        for k_ref, k in (('loss_D_fake', 'D_fake'), ('loss_D_real', 'D_real'), ('loss_G_gan', 'G_gan'),
                         ('loss_G_recon', 'G_recon'), ('loss_G_distill', 'G_distill')):
            r = float(ref[k_ref])
            assert abs(L[k] - r) <= 2e-2 * max(1.0, abs(r)), (name, it, k, L[k], r)
        for i in range(4):
            assert abs(L['G_distill%d' % i] - float(ref['loss_G_distill_terms'][i])) <= 5e-3, (name, it, i)
        lr = fix['hp']['lr']
        for tag, net, sd in (('S', eng.S, state['student_sd']), ('D', eng.D, state['D_sd'])):
            worst, mean_d, cnt = 0.0, 0.0, 0
            inert = _inert_biases(sd, tag)
            mine_sd = net.state_dict()
            for k, v in sd.items():
                if not v.is_floating_point() or k not in mine_sd or k.endswith('num_batches_tracked') or k in inert:
                    continue
                dlt = (mine_sd[k].double() - v.detach().double()).abs()
                if k.endswith('running_mean') or k.endswith('running_var'):
                    assert float(dlt.max()) <= 3e-2 * max(1.0, float(v.abs().max())), (name, it, k, float(dlt.max()))
                    continue
                worst = max(worst, float(dlt.max()))
                mean_d += float(dlt.sum())
                cnt += dlt.numel()
            report[f'{tag}_w_worst_it{it}'] = worst / lr
            report[f'{tag}_w_mean_it{it}'] = mean_d / cnt / lr
    return report


def _check(rep, strict):
    for k, v in rep.items():
        if k.startswith(('Tfake', 'Sfake', 'Tact', 'Sact')):
            assert v <= 3e-2, (k, v)
        elif k.endswith('_grads'):
            assert v <= (6e-2 if strict else 0.5), (k, v)
        elif strict and '_w_worst' in k:
            assert v <= 2.1 * (int(k[-1]) + 1), (k, v)
        elif strict and '_w_mean' in k:
            assert v <= 0.15 * (int(k[-1]) + 1), (k, v)


@pytest.mark.parametrize('name', CASES)
def test_distill_step_with_injected_loss_gradients(golden_dir, name):
    rep = _run(golden_dir, name, use_graph=False, inject=True)
    print(name, 'injected', {k: round(v, 4) for k, v in rep.items()})
    _check(rep, strict=True)


@pytest.mark.parametrize('name', CASES)
@pytest.mark.parametrize('use_graph', [False, True])
def test_distill_step_matches_oracle(golden_dir, name, use_graph):
    rep = _run(golden_dir, name, use_graph, inject=False)
    print(name, 'graph' if use_graph else 'eager', {k: round(v, 4) for k, v in rep.items()})
    _check(rep, strict=name in SMOOTH)
