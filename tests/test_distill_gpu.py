"""End-to-end parity of the CUDA distillation step against the CPU oracle (itself pinned to the
reference by tests/test_oracle_golden.py) on the committed golden fixtures.

Two comparisons, both on the same inputs and weights:

(1) against the fp32 oracle (the reference algorithm):
      activations / outputs: relative L2 error <= 3e-2;   losses: |delta| <= 2e-2*max(1,|loss|) on the
      first step (5e-2 on the second, which starts from weights that already differ by O(lr));
      KA terms: |delta| <= 5e-3 (2e-2 on the second step);   parameter gradients: relative L2 <= 0.5.
(2) against the same oracle with bf16 storage emulated at exactly the points where cat_b200 keeps bf16
    in HBM (oracle.cat_oracle.emulate_bf16): activations <= 3e-2, parameter gradients (relative L2 over
    all parameters of a network) <= 0.5 (measured on an idle GPU with the v1 kernels: 0.3% for D and 3-5% for
    the student on the smooth-loss fixture; up to ~0.3 on the L1 / hinge fixtures, where a one-ulp difference
    in the bf16 student output still flips a few signs of the L1 gradient; the asserted bound only has to
    exclude real defects, which are O(1)), post-Adam weights within 2.1*lr*(step+1) with mean
    |delta| <= 0.25*lr*(step+1), running statistics <= 1e-2.

Why gradients are only loosely comparable with the fp32 oracle: ReLU/LeakyReLU masks, the hinge mask and
sign(S-B) of the L1 loss are discontinuous in the forward values.  A ~1% forward rounding difference flips
the mask of the ~0.5% of elements that sit within 1% of zero, and every flipped element changes the
back-propagated signal by 100% locally, i.e. ~sqrt(0.005) = 7% relative L2 per activation layer.  The
bf16-emulating oracle itself moves by S_grads 26% / D_grads 7% from the fp32 oracle on these fixtures
(measured on CPU), which is the same distance the CUDA engine shows; comparison (2) removes that
conditioning effect and pins the kernels.
"""
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]

CASES = ['pix2pix_bn_lsgan_l2', 'pix2pix_bn_hinge', 'cyclegan_in_lsgan']
SMOOTH = {'pix2pix_bn_lsgan_l2'}   # l2 recon + lsgan: loss gradients are smooth in the forward values


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


def _oracle_state(fix, O):
    return dict(teacher_sd=O.clone_sd(fix['teacher_sd']), student_sd=O.clone_sd(fix['student_sd0']),
                D_sd=O.clone_sd(fix['D_sd0']), teacher_arch=fix['teacher_arch'], student_arch=fix['student_arch'],
                D_arch=fix['D_arch'], adam_G={}, adam_D={})


def _run(golden_dir, name, use_graph, kernels):
    from cat_b200 import ops
    ops.USE_HALO = kernels != 'v1'
    from cat_b200.distill_engine import DistillStep
    from oracle import cat_oracle as O
    fix = torch.load(os.path.join(golden_dir, name + '.pt'), weights_only=False)
    s0 = fix['steps'][0]
    B, _, H, W = s0['real_A'].shape
    eng = DistillStep(fix['teacher_arch'], fix['student_arch'], fix['D_arch'], fix['hp'], B, H, W,
                      use_cuda_graph=use_graph)
    eng.load(fix['teacher_sd'], fix['student_sd0'], fix['D_sd0'])
    st32, stq = _oracle_state(fix, O), _oracle_state(fix, O)
    rep32, repq = {}, {}
    for it, step in enumerate(fix['steps']):
        ref32 = O.distill_step(st32, step['real_A'], step['real_B'], fix['hp'])
        with O.emulate_bf16():
            refq = O.distill_step(stq, step['real_A'], step['real_B'], fix['hp'])
        eng.set_input(step['real_A'], step['real_B'])
        eng.step()
        torch.cuda.synchronize()
        L = eng.get_losses()
        if it == 0:
            for ref, rep in ((ref32, rep32), (refq, repq)):
                rep['Tfake'] = rel_l2(ops.nhwc_to_nchw(eng.T.out, 3).cpu(), ref['Tfake_B'])
                rep['Sfake'] = rel_l2(ops.nhwc_to_nchw(eng.S.out, 3).cpu(), ref['Sfake_B'])
                for n in O.MAPPING_LAYERS:
                    Ct, Cs = ref['Tacts'][n].shape[1], ref['Sacts'][n].shape[1]
                    rep['Tact ' + n] = rel_l2(ops.nhwc_to_nchw(eng.T.acts[n], Ct).cpu(), ref['Tacts'][n])
                    rep['Sact ' + n] = rel_l2(ops.nhwc_to_nchw(eng.S.acts[n], Cs).cpu(), ref['Sacts'][n])
                for tag, net, grads in (('S', eng.S, ref['S_grads']), ('D', eng.D, ref['D_grads'])):
                    mine, theirs = [], []
                    for k, g in grads.items():
                        if net.arena.has(k) and k not in _inert_biases(grads, tag):
                            mine.append(net.arena.view(k, 'g').flatten().cpu())
                            theirs.append(g.flatten())
                    rep[tag + '_grads'] = rel_l2(torch.cat(mine), torch.cat(theirs))
        for k_ref, k in (('loss_D_fake', 'D_fake'), ('loss_D_real', 'D_real'), ('loss_G_gan', 'G_gan'),
                         ('loss_G_recon', 'G_recon'), ('loss_G_distill', 'G_distill')):
            if it == 0:       # first step: within 2e-2 of BOTH oracles
                for ref in (ref32, refq):
                    r = float(ref[k_ref])
                    assert abs(L[k] - r) <= 2e-2 * max(1.0, abs(r)), (name, it, k, L[k], r)
            else:             # later steps start from weights that differ by O(lr) per element (the first Adam step is
                # lr * sign(g), so near-zero gradient components flip with the summation order; the two oracles are
                # themselves several per cent apart here): a sanity band of 15 % around the interval they span -- the
                # run-to-run spread of this step on the device reaches 7 % for the lsgan / InstanceNorm fixture
                # (profiles/r02_cyclegan_spread.txt; an excursion to 7.3 % was seen once in 5 sessions), while a defect
                # shows in the first step, which is held to 2e-2 above
                lo, hi = sorted((float(ref32[k_ref]), float(refq[k_ref])))
                slack = 15e-2 * max(1.0, abs(lo), abs(hi))
                assert lo - slack <= L[k] <= hi + slack, (name, it, k, L[k], lo, hi)
        for i in range(4):
            # the second step starts from weights that already differ by O(lr) (Adam sign flips)
            assert abs(L['G_distill%d' % i] - float(ref32['loss_G_distill_terms'][i])) <= (2e-2 if it else 5e-3), (name, it, i)
            assert abs(L['G_distill%d' % i] - float(refq['loss_G_distill_terms'][i])) <= (2e-2 if it else 2e-3), (name, it, i)
        lr = fix['hp']['lr']
        for tag, net, sd in (('S', eng.S, stq['student_sd']), ('D', eng.D, stq['D_sd'])):
            worst, mean_d, cnt = 0.0, 0.0, 0
            inert = _inert_biases(sd, tag)
            mine_sd = net.state_dict()
            for k, v in sd.items():
                if not v.is_floating_point() or k not in mine_sd or k.endswith('num_batches_tracked') or k in inert:
                    continue
                dlt = (mine_sd[k].double() - v.detach().double()).abs()
                if k.endswith('running_mean') or k.endswith('running_var'):
                    assert float(dlt.max()) <= 1e-2 * max(1.0, float(v.abs().max())), (name, it, k, float(dlt.max()))
                    continue
                worst = max(worst, float(dlt.max()))
                mean_d += float(dlt.sum())
                cnt += dlt.numel()
            repq[f'{tag}_w_worst_it{it}'] = worst / lr
            repq[f'{tag}_w_mean_it{it}'] = mean_d / cnt / lr
    return rep32, repq


@pytest.mark.parametrize('name', CASES)
@pytest.mark.parametrize('use_graph,kernels', [(False, 'v1'), (True, 'v1'), (True, 'auto')])
def test_distill_step_matches_oracle(golden_dir, name, use_graph, kernels):
    """kernels='v1': every GEMM on the gather-per-tap kernel, whose K order is the oracle's (strict bounds).
    kernels='auto': per-GEMM autotuned v1 / v2 (halo) kernels.  v2 sums the K steps in a different order
    (chunk-major), so a few bf16 roundings land one ulp away; after ~40 BN+ReLU layers on these tiny
    fixtures that is a ~1% activation difference, and gradients are then compared with the loose bound."""
    try:
        rep32, repq = _run(golden_dir, name, use_graph, kernels)
    finally:
        from cat_b200 import ops
        ops.USE_HALO = True
    tag = ('graph' if use_graph else 'eager') + '/' + kernels
    print(name, tag, 'vs fp32 oracle', {k: round(v, 4) for k, v in rep32.items()})
    print(name, tag, 'vs bf16-emulating oracle', {k: round(v, 4) for k, v in repq.items()})
    for k, v in rep32.items():
        assert v <= (0.5 if k.endswith('_grads') else 3e-2), ('fp32', k, v)
    # Bounds for the bf16-emulating oracle.  On an idle GPU the v1 kernels reproduce it almost exactly
    # (measured: student output identical to 4 digits, D gradients 0.3 %, student gradients 3-5 % on the
    # smooth-loss fixture), but the fp32 atomics of the norm statistics / weight gradients make the summation
    # order schedule dependent, and one flipped bf16 rounding is amplified by the ReLU / sign conditioning
    # described above; the asserted bounds therefore only exclude real defects (a wrong kernel is O(1) off).
    for k, v in repq.items():
        if k.endswith('_grads'):
            assert v <= 0.5, ('emu', k, v)
        elif '_w_worst' in k:
            assert v <= 2.1 * (int(k[-1]) + 1), ('emu', k, v)
        elif '_w_mean' in k:
            assert v <= 0.25 * (int(k[-1]) + 1), ('emu', k, v)
        else:
            assert v <= 3e-2, ('emu', k, v)


def test_side_stream_branches_do_not_change_the_gradients(golden_dir):
    """The weight-gradient side stream of GenNet.backward and the teacher branch only re-order independent work.  fp32
    atomics make every run of the step slightly different (summation order -> a few flipped bf16 roundings), so the
    yardstick is the run-to-run difference of the step WITHOUT side streams: the step with side streams must not differ
    from it by more than a small multiple of that (a missing dependency corrupts whole tensors, i.e. O(1))."""
    from cat_b200.distill_engine import DistillStep
    fix = torch.load(os.path.join(golden_dir, 'pix2pix_bn_lsgan_l2.pt'), weights_only=False)
    s0 = fix['steps'][0]
    B, _, H, W = s0['real_A'].shape

    def run(overlap):
        eng = DistillStep(fix['teacher_arch'], fix['student_arch'], fix['D_arch'], fix['hp'], B, H, W, use_cuda_graph=True)
        eng.overlap_teacher = overlap
        eng.S.overlap_wgrad = overlap
        eng.load(fix['teacher_sd'], fix['student_sd0'], fix['D_sd0'])
        eng.set_input(s0['real_A'], s0['real_B'])
        eng.step()
        torch.cuda.synchronize()
        return eng.S.arena.g.clone().cpu(), eng.get_losses()
    # three runs without side streams give the yardstick (the weight-gradient GEMMs are deterministic since the two-stage
    # form; the remaining run-to-run noise comes from the atomics of the norm statistics / depthwise gradients flipping
    # bf16 roundings), two runs with side streams are held against it
    (ga, la), (gb, lb), (gc, lc) = run(False), run(False), run(False)
    (g1, l1), (g2, l2) = run(True), run(True)
    noise = max(rel_l2(gb, ga), rel_l2(gc, ga), rel_l2(gc, gb))
    diff = min(rel_l2(g1, ga), rel_l2(g2, ga))
    print('student gradient: run-to-run', noise, 'side streams vs none', diff)
    assert diff <= 3 * noise + 2e-3, (diff, noise)
    for k in la:
        spread = max(abs(lb[k] - la[k]), abs(lc[k] - la[k]))
        assert min(abs(l1[k] - la[k]), abs(l2[k] - la[k])) <= 3 * spread + 1e-3 * max(1.0, abs(la[k])), (k, l1[k], l2[k], la[k], lb[k])
