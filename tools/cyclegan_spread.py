"""Run-to-run spread of the CycleGAN teacher-training step on the toy fixture (tests/golden/train_cyclegan_in_lsgan.pt):
the same two steps from the same state, repeated with / without the side-stream weight-gradient branches and CUDA graphs,
next to the fp32 and bf16-emulating oracle.  Separates a race (spread only with overlap) from the amplification of
summation-order noise by the first Adam step (update = lr * sign(g) for every weight) on this ill-conditioned fixture."""
import os
import random
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cat_b200 import ops  # noqa: E402
from cat_b200.train_engine import CycleGANTrainStep  # noqa: E402
from oracle import cat_oracle as O  # noqa: E402
from oracle import train_oracle as TO  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else 'train_cyclegan_in_lsgan'
    fix = torch.load(os.path.join(ROOT, 'tests', 'golden', name + '.pt'), weights_only=False)
    hp = fix['hp']
    steps = fix['steps'][:2]
    B, _, H, W = steps[0]['real_A'].shape

    def oracle_run(emulate):
        st = dict(G_A_sd=O.clone_sd(fix['G_A_sd0']), G_B_sd=O.clone_sd(fix['G_B_sd0']), D_A_sd=O.clone_sd(fix['D_A_sd0']),
                  D_B_sd=O.clone_sd(fix['D_B_sd0']), G_arch=fix['G_arch'], D_arch=fix['D_arch'], adam_G={}, adam_D={},
                  pool_A=TO.ImagePool(hp['pool_size']), pool_B=TO.ImagePool(hp['pool_size']))
        random.seed(fix['python_random_seed'])
        if emulate:
            with O.emulate_bf16():
                return [TO.cyclegan_train_step(st, s['real_A'], s['real_B'], hp) for s in steps]
        return [TO.cyclegan_train_step(st, s['real_A'], s['real_B'], hp) for s in steps]

    keys = CycleGANTrainStep.LOSS_NAMES
    for tag, refs in (('oracle fp32', oracle_run(False)), ('oracle bf16-emulated', oracle_run(True))):
        for it, r in enumerate(refs):
            print(f'{tag:28s} step {it}: ' + ' '.join(f'{k}={float(r["loss_" + k]):.4f}' for k in keys))
    ref0 = oracle_run(True)[0]
    for graph in (False, True):
        for overlap in (True, False):
            for rep in range(3):
                eng = CycleGANTrainStep(fix['G_arch'], fix['D_arch'], hp, B, H, W, device='cuda:0', use_cuda_graph=graph)
                for g in eng._gens('A') + eng._gens('B'):
                    g.overlap_wgrad = overlap
                eng.load(fix['G_A_sd0'], fix['G_B_sd0'], fix['D_A_sd0'], fix['D_B_sd0'])
                random.seed(fix['python_random_seed'])
                for it, s in enumerate(steps):
                    eng.set_input(s['real_A'], s['real_B'])
                    eng.step()
                    torch.cuda.synchronize()
                    L = eng.get_losses()
                    extra = ''
                    if it == 0 and not graph:
                        errs = []
                        for net, key in ((eng.G_A, 'G_A_grads'), (eng.G_B, 'G_B_grads')):
                            mine = torch.cat([net.arena.view(k, 'g').flatten().cpu() for k in ref0[key] if net.arena.has(k)])
                            theirs = torch.cat([v.flatten() for k, v in ref0[key].items() if net.arena.has(k)])
                            errs.append(float((mine - theirs).norm() / theirs.norm()))
                        extra = f'  G grads rel-L2 vs bf16 oracle: {errs[0]:.3f} {errs[1]:.3f}'
                    print(f'graph={int(graph)} overlap={int(overlap)} rep {rep} step {it}: ' + ' '.join(f'{k}={L[k]:.4f}' for k in keys) + extra)


if __name__ == '__main__':
    ops.require_cuda()
    main()
