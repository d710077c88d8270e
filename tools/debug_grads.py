"""Per-tensor gradient error report of the CUDA step vs the CPU oracle (debugging aid, GPU box)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cat_b200 import ops  # noqa: E402
from cat_b200.distill_engine import DistillStep  # noqa: E402
from oracle import cat_oracle as O  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def main(name):
    fix = torch.load(os.path.join(ROOT, 'tests', 'golden', name + '.pt'), weights_only=False)
    step = fix['steps'][0]
    B, _, H, W = step['real_A'].shape
    eng = DistillStep(fix['teacher_arch'], fix['student_arch'], fix['D_arch'], fix['hp'], B, H, W)
    eng.load(fix['teacher_sd'], fix['student_sd0'], fix['D_sd0'])
    state = dict(teacher_sd=O.clone_sd(fix['teacher_sd']), student_sd=O.clone_sd(fix['student_sd0']),
                 D_sd=O.clone_sd(fix['D_sd0']), teacher_arch=fix['teacher_arch'], student_arch=fix['student_arch'],
                 D_arch=fix['D_arch'], adam_G={}, adam_D={})
    ref = O.distill_step(state, step['real_A'], step['real_B'], fix['hp'])
    eng.set_input(step['real_A'], step['real_B'])
    # capture d(activation) at the mapping layers right after the KA gradient is added
    captured = {}
    orig_backward = eng.S.backward

    def spy_backward(d_out, act_grads=None):
        wrapped = {}
        for n, fn in (act_grads or {}).items():
            def w(dact, n=n, fn=fn):
                fn(dact)
                captured[n] = dact.t.clone()
            wrapped[n] = w
        captured['dS'] = d_out.t.clone()
        return orig_backward(d_out, wrapped)
    eng.S.backward = spy_backward
    eng.step()
    torch.cuda.synchronize()
    print('dS (grad wrt student output) rel', rel(captured['dS'][..., :3].permute(0, 3, 1, 2).float().cpu(), ref['Sfake_grad']))
    for n in O.MAPPING_LAYERS:
        g = ref['Sact_grads'][n]
        mine = captured[n][..., :g.shape[1]].permute(0, 3, 1, 2).float().cpu()
        print('d(Sact %s) rel %.4f  |ref| %.3e' % (n, rel(mine, g), float(g.norm())))
    for tag, net, grads in (('D', eng.D, ref['D_grads']),) + ((('S', eng.S, ref['S_grads']),) if '--all' in sys.argv else ()):
        scale = max(float(g.abs().max()) for g in grads.values())
        print(f'---- {tag} grads (global max {scale:.3e})')
        for k, g in grads.items():
            if not net.arena.has(k):
                print('   (not in arena)', k)
                continue
            mine = net.arena.view(k, 'g').cpu()
            print(f'{k:40s} shape {str(tuple(g.shape)):18s} |ref| {float(g.norm()):.3e} rel {rel(mine, g):.4f}')
    print('---- activation gradients at the mapping layers are not stored; pred check:')
    print('pred_fake_D rel', rel(eng.D.pred[..., 0].cpu(), ref['pred_fake_D'][:, 0]) if False else 'n/a')


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else 'pix2pix_bn_hinge')
