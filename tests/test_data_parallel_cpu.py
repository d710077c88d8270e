"""world_size-2 gloo test of the data-parallel recipe (cat_b200/parallel.py) on CPU.

Two ranks run the distillation step on their shard of a global batch with the recipe used by
DistillStep (local KA scaled by world_size, SUM all-reduce of the flat gradient arena, 1/world_size in
Adam); the result must equal a single-process evaluation with the reference's nn.DataParallel
semantics (L1 / GAN losses as means over the gathered global batch, KA summed over the replicas'
shards -- distillers/inception_distiller.py:137-170).  The CPU oracle supplies the arithmetic on both
sides; the InstanceNorm fixture is used because its normalisation has no cross-sample coupling, so a
global-batch forward equals the concatenation of the shard forwards (BatchNorm statistics are per
replica in both the reference and cat_b200)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cat_b200 import parallel
from oracle import cat_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORLD = 2
PER_RANK = 2


def _fresh_state(fix):
    return dict(teacher_sd=O.clone_sd(fix['teacher_sd'], torch.float64), student_sd=O.clone_sd(fix['student_sd0'], torch.float64),
                D_sd=O.clone_sd(fix['D_sd0'], torch.float64), teacher_arch=fix['teacher_arch'],
                student_arch=fix['student_arch'], D_arch=fix['D_arch'], adam_G={}, adam_D={})


def _batch(fix):
    g = torch.Generator().manual_seed(5)
    _, _, H, W = fix['steps'][0]['real_A'].shape
    a = (torch.rand(WORLD * PER_RANK, 3, H, W, generator=g, dtype=torch.float64) * 2 - 1)
    b = (torch.rand(WORLD * PER_RANK, 3, H, W, generator=g, dtype=torch.float64) * 2 - 1)
    return a, b


def _flatten(grads):
    keys = sorted(grads)
    return keys, torch.cat([grads[k].reshape(-1) for k in keys])


def _worker(rank, port, path, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=WORLD)
    torch.set_num_threads(2)
    fix = torch.load(path, weights_only=False)
    a, b = _batch(fix)
    sl = slice(rank * PER_RANK, (rank + 1) * PER_RANK)
    hp = dict(fix['hp'], ka_scale=parallel.ka_scale(WORLD))

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
    state = _fresh_state(fix)
    res = O.distill_step(state, a[sl], b[sl], hp, grad_hook=hook)
    torch.save({'S_grads': res['S_grads'], 'D_grads': res['D_grads'], 'student_sd': state['student_sd'],
                'D_sd': state['D_sd']}, os.path.join(out_dir, f'rank{rank}.pt'))
    dist.destroy_process_group()


def _data_parallel_reference(fix):
    """Single process, reference nn.DataParallel semantics on the global batch (see module docstring)."""
    import torch.nn.functional as F
    state = _fresh_state(fix)
    a, b = _batch(fix)
    hp = fix['hp']
    T_sd, S_sd, D_sd = state['teacher_sd'], state['student_sd'], state['D_sd']
    Tacts, Sacts = {}, {}
    with torch.no_grad():
        Tfake = O.generator_forward(T_sd, fix['teacher_arch'], a, False, Tacts)
    S_params = {k: v.requires_grad_(True) for k, v in S_sd.items() if k.endswith(('.weight', '.bias'))}
    Sfake = O.generator_forward(S_sd, fix['student_arch'], a, True, Sacts)
    D_params = {k: v.requires_grad_(True) for k, v in D_sd.items() if k.endswith(('.weight', '.bias'))}
    assert not hp['aligned']
    loss_D = 0.5 * (O.gan_loss(hp['gan_mode'], O.discriminator_forward(D_sd, fix['D_arch'], Sfake.detach()), False, True) +
                    O.gan_loss(hp['gan_mode'], O.discriminator_forward(D_sd, fix['D_arch'], b), True, True))
    loss_D.backward()
    D_grads = {k: p.grad.clone() for k, p in D_params.items()}
    with torch.no_grad():
        O.adam_update(D_params, D_grads, state['adam_D'], hp['lr'], hp['beta1'])
    for p in D_params.values():
        p.requires_grad_(False)
    loss = F.l1_loss(Sfake, Tfake) * hp['lambda_recon']
    loss = loss + O.gan_loss(hp['gan_mode'], O.discriminator_forward(D_sd, fix['D_arch'], Sfake), True, False) * hp['lambda_gan']
    for n in O.MAPPING_LAYERS:                       # KA per replica shard, summed over the replicas
        for r in range(WORLD):
            sl = slice(r * PER_RANK, (r + 1) * PER_RANK)
            loss = loss - O.ka(Sacts[n][sl], Tacts[n][sl]) * hp['lambda_distill']
    loss.backward()
    S_grads = {k: p.grad.clone() for k, p in S_params.items()}
    with torch.no_grad():
        O.adam_update(S_params, S_grads, state['adam_G'], hp['lr'], hp['beta1'])
    return S_grads, D_grads, state


@pytest.mark.timeout(600)
def test_two_rank_gloo_matches_data_parallel_semantics(golden_dir, tmp_path):
    path = os.path.join(golden_dir, 'cyclegan_in_lsgan.pt')
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(port, path, str(tmp_path)), nprocs=WORLD, join=True)
    fix = torch.load(path, weights_only=False)
    S_ref, D_ref, st_ref = _data_parallel_reference(fix)
    ranks = [torch.load(os.path.join(tmp_path, f'rank{r}.pt'), weights_only=False) for r in range(WORLD)]
    sscale = max(float(g.abs().max()) for g in S_ref.values())
    dscale = max(float(g.abs().max()) for g in D_ref.values())
    for r in ranks:
        for k, g in S_ref.items():
            assert float((r['S_grads'][k] - g).abs().max()) <= 1e-9 * sscale + 1e-12, k
        for k, g in D_ref.items():
            assert float((r['D_grads'][k] - g).abs().max()) <= 1e-9 * dscale + 1e-12, k
    # both ranks hold identical weights after the step (no parameter broadcast is ever needed)
    for k, v in ranks[0]['student_sd'].items():
        assert torch.equal(v, ranks[1]['student_sd'][k]), k
    for k, v in ranks[0]['D_sd'].items():
        if v.is_floating_point() and not k.endswith(('running_mean', 'running_var')):
            assert torch.equal(v, ranks[1]['D_sd'][k]), k
