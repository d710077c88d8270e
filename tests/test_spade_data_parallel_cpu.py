"""world_size-2 gloo test of the data-parallel recipe on the SPADE distillation step (CPU, oracle arithmetic).

Recipe (cat_b200/spade_distill_engine.py): every rank runs the step on its shard with per-rank
BatchNorm statistics and its own KA term (ka_scale = 1: the SPADE distiller averages the replica losses,
models/spade_model.py:191), the flat gradient arena is SUM all-reduced once per optimiser and scaled by 1/world_size
inside Adam.  The result must equal the average of the per-shard gradients computed in one process, and both ranks must
end the step with identical weights (no parameter broadcast is ever needed)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cat_b200 import parallel
from oracle import spade_oracle as SO
from oracle.cat_oracle import clone_sd

WORLD = 2
PER_RANK = 2


def _state(fix, vgg):
    return dict(teacher_sd=clone_sd(fix['teacher_sd']), student_sd=clone_sd(fix['student_sd0']), D_sd=clone_sd(fix['D_sd0']),
                vgg_sd=vgg, teacher_arch=fix['teacher_arch'], student_arch=fix['student_arch'], D_arch=fix['D_arch'],
                adam_G={}, adam_D={})


def _batch(fix):
    s = fix['steps']
    lab = torch.cat([s[0]['label'], s[1]['label']])
    inst = torch.cat([s[0]['instance'], s[1]['instance']])
    img = torch.cat([s[0]['image'], s[1]['image']])
    assert lab.shape[0] == WORLD * PER_RANK
    return SO.preprocess_input(lab, inst, fix['hp']['n_label']), img


def _flatten(grads):
    keys = sorted(grads)
    return keys, torch.cat([grads[k].reshape(-1) for k in keys])


def _worker(rank, port, path, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=WORLD)
    torch.set_num_threads(2)
    fix = torch.load(path, weights_only=False)
    seg, img = _batch(fix)
    sl = slice(rank * PER_RANK, (rank + 1) * PER_RANK)

    def hook(tag, grads):
        keys, flat = _flatten(grads)                       # the flat gradient arena of the engine
        parallel.reduce_gradients(flat, WORLD)
        flat *= parallel.grad_scale(WORLD)                 # applied inside catb_adam on the GPU
        out, o = {}, 0
        for k in keys:
            n = grads[k].numel()
            out[k] = flat[o:o + n].view_as(grads[k]).clone()
            o += n
        return out
    state = _state(fix, SO.make_vgg_sd(fix['vgg_seed']))
    res = SO.spade_distill_step(state, seg[sl], img[sl], dict(fix['hp'], ka_scale=1.0), grad_hook=hook)
    torch.save({'S_grads': res['S_grads'], 'D_grads': res['D_grads'],
                'student_params': {k: v for k, v in state['student_sd'].items() if SO._is_param(k)},
                'D_params': {k: v for k, v in state['D_sd'].items() if SO._is_param(k)}}, os.path.join(out_dir, f'rank{rank}.pt'))
    dist.destroy_process_group()


@pytest.mark.timeout(900)
@pytest.mark.slow      # the engine-level 2-rank test (tests/test_engine_data_parallel_cpu.py) covers the same recipe
def test_two_rank_gloo_spade_step(golden_dir, tmp_path):
    path = os.path.join(golden_dir, 'spade_more.pt')
    port = 31500 + os.getpid() % 2000
    mp.spawn(_worker, args=(port, path, str(tmp_path)), nprocs=WORLD, join=True)
    ranks = [torch.load(os.path.join(tmp_path, f'rank{r}.pt'), weights_only=False) for r in range(WORLD)]
    # both ranks hold identical reduced gradients and identical weights after the step
    for key in ('S_grads', 'D_grads', 'student_params', 'D_params'):
        for k, v in ranks[0][key].items():
            assert torch.equal(v, ranks[1][key][k]), (key, k)
    # the reduced student gradient is the average of the per-shard gradients (computed here without any collective);
    # the D gradients additionally depend on each rank's post-update student, which is identical by the check above
    fix = torch.load(path, weights_only=False)
    seg, img = _batch(fix)
    vgg = SO.make_vgg_sd(fix['vgg_seed'])
    shard = []
    for r in range(WORLD):
        sl = slice(r * PER_RANK, (r + 1) * PER_RANK)
        res = SO.spade_distill_step(_state(fix, vgg), seg[sl], img[sl], dict(fix['hp'], ka_scale=1.0, lr_G=0.0, lr_D=0.0))
        shard.append(res['S_grads'])
    scale = max(float(g.abs().max()) for g in ranks[0]['S_grads'].values())
    for k, g in ranks[0]['S_grads'].items():
        avg = sum(s[k] for s in shard) / WORLD
        assert float((g - avg).abs().max()) <= 1e-5 * scale + 1e-9, k
