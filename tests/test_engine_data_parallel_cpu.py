"""world_size-2 gloo test of the data-parallel path of the ENGINES (not just the recipe): two processes run
cat_b200.distill_engine.DistillStep / cat_b200.spade_distill_engine.SpadeDistillStep with world_size=2 on their shard of a
global batch, with every kernel wrapper swapped for its torch restatement in exact (fp32) mode (oracle/kernel_emu.py, test
infrastructure).  Checked: the all-reduced gradient arenas equal the reference's nn.DataParallel semantics evaluated in one
process by the oracle (Inception distiller: tests/test_data_parallel_cpu._data_parallel_reference; SPADE distiller: the
average of the per-shard gradients), and both ranks hold bit-identical weights after the step."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
for _p in (HERE, os.path.dirname(HERE)):
    if _p not in sys.path:
        sys.path.insert(0, _p)

WORLD = 2
PER_RANK = 2


def _worker_inception(rank, port, path, out_dir, early='0'):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), CATB_EARLY_REDUCE=early)
    dist.init_process_group('gloo', rank=rank, world_size=WORLD)
    torch.set_num_threads(2)
    from oracle.kernel_emu import emulated_kernels
    from cat_b200 import parallel
    from test_data_parallel_cpu import _batch
    fix = torch.load(path, weights_only=False)
    a, b = _batch(fix)
    sl = slice(rank * PER_RANK, (rank + 1) * PER_RANK)
    hp = dict(fix['hp'], ka_scale=parallel.ka_scale(WORLD))
    _, _, H, W = a.shape
    with emulated_kernels(exact=True):
        from cat_b200.distill_engine import DistillStep
        eng = DistillStep(fix['teacher_arch'], fix['student_arch'], fix['D_arch'], hp, PER_RANK, H, W, device='cpu', world_size=WORLD)
        eng.load(fix['teacher_sd'], fix['student_sd0'], fix['D_sd0'])
        eng.set_input(a[sl].float(), b[sl].float())
        eng.step()
        scale = parallel.grad_scale(WORLD)
        out = {'S_g': {k: eng.S.arena.view(k, 'g').clone() * scale for k in eng.S.arena.entries},
               'D_g': {k: eng.D.arena.view(k, 'g').clone() * scale for k in eng.D.arena.entries},
               'S_p': eng.S.arena.p.clone(), 'D_p': eng.D.arena.p.clone()}
    torch.save(out, os.path.join(out_dir, f'rank{rank}.pt'))
    dist.destroy_process_group()


@pytest.mark.timeout(900)
@pytest.mark.parametrize('early', ['0', '1'])      # the default two all-reduces / the layer-wise reduction inside the D backward pass
def test_engine_two_ranks_inception(golden_dir, tmp_path, early):
    from test_data_parallel_cpu import _data_parallel_reference
    path = os.path.join(golden_dir, 'cyclegan_in_lsgan.pt')
    port = 33500 + os.getpid() % 2000 + int(early)
    mp.spawn(_worker_inception, args=(port, path, str(tmp_path), early), nprocs=WORLD, join=True)
    fix = torch.load(path, weights_only=False)
    S_ref, D_ref, _ = _data_parallel_reference(fix)       # fp64, single process, reference DataParallel semantics
    ranks = [torch.load(os.path.join(tmp_path, f'rank{r}.pt'), weights_only=False) for r in range(WORLD)]
    assert torch.equal(ranks[0]['S_p'], ranks[1]['S_p']) and torch.equal(ranks[0]['D_p'], ranks[1]['D_p'])
    for tag, ref in (('S_g', S_ref), ('D_g', D_ref)):
        scale = max(float(g.abs().max()) for g in ref.values())
        for k, g in ref.items():
            if k not in ranks[0][tag]:
                continue
            err = float((ranks[0][tag][k].double() - g).abs().max())
            assert err <= 2e-3 * float(g.abs().max()) + 2e-5 * scale, (tag, k, err, float(g.abs().max()))


def _worker_spade(rank, port, path, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=WORLD)
    torch.set_num_threads(2)
    from oracle import spade_oracle as SO
    from oracle.kernel_emu import emulated_kernels
    from cat_b200 import parallel
    fix = torch.load(path, weights_only=False)
    s = fix['steps']
    lab, inst, img = (torch.cat([s[0][k], s[1][k]]) for k in ('label', 'instance', 'image'))
    sl = slice(rank * PER_RANK, (rank + 1) * PER_RANK)
    H, W = img.shape[2:]
    with emulated_kernels(exact=True):
        from cat_b200.spade_distill_engine import SpadeDistillStep
        eng = SpadeDistillStep(fix['teacher_arch'], fix['student_arch'], fix['D_arch'], dict(fix['hp'], ka_scale=1.0), PER_RANK, H, W,
                               device='cpu', world_size=WORLD)
        eng.load(fix['teacher_sd'], fix['student_sd0'], fix['D_sd0'], SO.make_vgg_sd(fix['vgg_seed']))
        eng.set_input(lab[sl], inst[sl], img[sl])
        eng.step()
        scale = parallel.grad_scale(WORLD)
        out = {'S_g': {k: eng.S.arena.view(k, 'g').clone() * scale for k in eng.S.arena.entries},
               'S_p': eng.S.arena.p.clone(), 'D_p': eng.D.arena.p.clone()}
    torch.save(out, os.path.join(out_dir, f'rank{rank}.pt'))
    dist.destroy_process_group()


@pytest.mark.timeout(1200)
def test_engine_two_ranks_spade(golden_dir, tmp_path):
    from oracle import spade_oracle as SO
    from test_spade_data_parallel_cpu import _batch, _state
    path = os.path.join(golden_dir, 'spade_more.pt')
    port = 35500 + os.getpid() % 2000
    mp.spawn(_worker_spade, args=(port, path, str(tmp_path)), nprocs=WORLD, join=True)
    ranks = [torch.load(os.path.join(tmp_path, f'rank{r}.pt'), weights_only=False) for r in range(WORLD)]
    assert torch.equal(ranks[0]['S_p'], ranks[1]['S_p']) and torch.equal(ranks[0]['D_p'], ranks[1]['D_p'])
    fix = torch.load(path, weights_only=False)
    seg, img = _batch(fix)
    vgg = SO.make_vgg_sd(fix['vgg_seed'])
    shard = []
    for r in range(WORLD):
        sl = slice(r * PER_RANK, (r + 1) * PER_RANK)
        res = SO.spade_distill_step(_state(fix, vgg), seg[sl], img[sl], dict(fix['hp'], ka_scale=1.0, lr_G=0.0, lr_D=0.0))
        shard.append(res['S_grads'])
    scale = max(float(g.abs().max()) for g in shard[0].values())
    for k in shard[0]:
        avg = sum(s[k] for s in shard) / WORLD
        err = float((ranks[0]['S_g'][k] - avg).abs().max())
        assert err <= 2e-3 * float(avg.abs().max()) + 2e-5 * scale, (k, err, float(avg.abs().max()))


def _worker_cyclegan(rank, port, path, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=WORLD)
    torch.set_num_threads(2)
    from oracle.kernel_emu import emulated_kernels
    from cat_b200 import parallel
    fix = torch.load(path, weights_only=False)
    s = fix['steps'][0]
    _, _, H, W = s['real_A'].shape
    with emulated_kernels(exact=True):
        from cat_b200.train_engine import CycleGANTrainStep
        eng = CycleGANTrainStep(fix['G_arch'], fix['D_arch'], fix['hp'], 1, H, W, device='cpu', world_size=WORLD)
        eng.load(fix['G_A_sd0'], fix['G_B_sd0'], fix['D_A_sd0'], fix['D_B_sd0'])
        eng.set_input(s['real_A'][rank:rank + 1], s['real_B'][rank:rank + 1])
        eng.step()
        scale = parallel.grad_scale(WORLD)
        out = {}
        for tag, net in (('G_A', eng.G_A), ('G_B', eng.G_B), ('D_A', eng.D_A), ('D_B', eng.D_B)):
            out[tag + '_g'] = {k: net.arena.view(k, 'g').clone() * scale for k in net.arena.entries}
            out[tag + '_p'] = net.arena.p.clone()
    torch.save(out, os.path.join(out_dir, f'rank{rank}.pt'))
    dist.destroy_process_group()


@pytest.mark.timeout(900)
def test_engine_two_ranks_cyclegan_training(golden_dir, tmp_path):
    """Teacher training under data parallelism: CycleGANTrainStep with one image per rank.  With
    InstanceNorm and per-sample-mean losses the reference's nn.DataParallel step on the global batch of two
    (models/networks.py:160-161: replicas' outputs are gathered, the losses are means over the gathered batch) has exactly
    the averaged per-shard gradients, so the all-reduced arenas must equal the single-process oracle step on both images;
    four all-reduces per step (two generators, two discriminators), bit-identical weights on both ranks afterwards."""
    import random
    from oracle import train_oracle as TO
    from oracle.cat_oracle import clone_sd
    path = os.path.join(golden_dir, 'train_cyclegan_in_lsgan.pt')
    port = 37500 + os.getpid() % 2000
    mp.spawn(_worker_cyclegan, args=(port, path, str(tmp_path)), nprocs=WORLD, join=True)
    ranks = [torch.load(os.path.join(tmp_path, f'rank{r}.pt'), weights_only=False) for r in range(WORLD)]
    fix = torch.load(path, weights_only=False)
    hp, s = fix['hp'], fix['steps'][0]
    st = dict(G_A_sd=clone_sd(fix['G_A_sd0']), G_B_sd=clone_sd(fix['G_B_sd0']), D_A_sd=clone_sd(fix['D_A_sd0']),
              D_B_sd=clone_sd(fix['D_B_sd0']), G_arch=fix['G_arch'], D_arch=fix['D_arch'], adam_G={}, adam_D={},
              pool_A=TO.ImagePool(hp['pool_size']), pool_B=TO.ImagePool(hp['pool_size']))
    random.seed(0)
    ref = TO.cyclegan_train_step(st, s['real_A'], s['real_B'], hp)
    for tag in ('G_A', 'G_B', 'D_A', 'D_B'):
        assert torch.equal(ranks[0][tag + '_p'], ranks[1][tag + '_p']), tag
        grads = ref[tag + '_grads']
        scale = max(float(g.abs().max()) for g in grads.values())
        n = 0
        for k, g in grads.items():
            if k not in ranks[0][tag + '_g']:
                continue
            n += 1
            err = float((ranks[0][tag + '_g'][k] - g).abs().max())
            assert err <= 2e-3 * float(g.abs().max()) + 2e-5 * scale, (tag, k, err, float(g.abs().max()))
        assert n > 5, tag


def _worker_pix2pix(rank, port, path, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=WORLD)
    torch.set_num_threads(2)
    from oracle.kernel_emu import emulated_kernels
    from cat_b200 import parallel
    fix = torch.load(path, weights_only=False)
    s = fix['steps'][0]
    _, _, H, W = s['real_A'].shape
    with emulated_kernels(exact=True):
        from cat_b200.train_engine import Pix2PixTrainStep
        eng = Pix2PixTrainStep(fix['G_arch'], fix['D_arch'], fix['hp'], 1, H, W, device='cpu', world_size=WORLD)
        eng.load(fix['G_sd0'], fix['D_sd0'])
        eng.set_input(s['real_A'][rank:rank + 1], s['real_B'][rank:rank + 1])
        eng.step()
        scale = parallel.grad_scale(WORLD)
        out = {'G_g': {k: eng.G.arena.view(k, 'g').clone() * scale for k in eng.G.arena.entries},
               'G_p': eng.G.arena.p.clone(), 'D_p': eng.D.arena.p.clone()}
    torch.save(out, os.path.join(out_dir, f'rank{rank}.pt'))
    dist.destroy_process_group()


@pytest.mark.slow      # same recipe and engine base class as the CycleGAN / Inception two-rank tests above
@pytest.mark.timeout(900)
def test_engine_two_ranks_pix2pix_training(golden_dir, tmp_path):
    """Pix2PixTrainStep with one image per rank (InstanceNorm fixture, lsgan + l2): the all-reduced generator gradient equals
    the single-process oracle step on both images (mean losses over the gathered batch, models/networks.py:160-161);
    bit-identical weights on both ranks after the step."""
    from oracle import train_oracle as TO
    from oracle.cat_oracle import clone_sd
    path = os.path.join(golden_dir, 'train_pix2pix_in_lsgan_l2.pt')
    port = 39500 + os.getpid() % 2000
    mp.spawn(_worker_pix2pix, args=(port, path, str(tmp_path)), nprocs=WORLD, join=True)
    ranks = [torch.load(os.path.join(tmp_path, f'rank{r}.pt'), weights_only=False) for r in range(WORLD)]
    assert torch.equal(ranks[0]['G_p'], ranks[1]['G_p']) and torch.equal(ranks[0]['D_p'], ranks[1]['D_p'])
    fix = torch.load(path, weights_only=False)
    s = fix['steps'][0]
    st = dict(G_sd=clone_sd(fix['G_sd0']), D_sd=clone_sd(fix['D_sd0']), G_arch=fix['G_arch'], D_arch=fix['D_arch'],
              adam_G={}, adam_D={})
    # the generator phase runs on the discriminator AFTER its update, which under DP is the all-reduced one: the oracle
    # on the full batch takes exactly that step
    ref = TO.pix2pix_train_step(st, s['real_A'], s['real_B'], fix['hp'])
    grads = ref['G_grads']
    scale = max(float(g.abs().max()) for g in grads.values())
    n = 0
    for k, g in grads.items():
        if k not in ranks[0]['G_g']:
            continue
        n += 1
        err = float((ranks[0]['G_g'][k] - g).abs().max())
        assert err <= 2e-3 * float(g.abs().max()) + 2e-5 * scale, (k, err, float(g.abs().max()))
    assert n > 10


def _worker_mirror(rank, port, path, out_dir):
    """The user-facing distiller mirror under data parallelism: what bench.py's end-to-end loop drives at N > 1."""
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), WORLD_SIZE=str(WORLD), RANK=str(rank))
    dist.init_process_group('gloo', rank=rank, world_size=WORLD)
    torch.set_num_threads(2)
    from oracle.kernel_emu import emulated_kernels
    from test_data_parallel_cpu import _batch
    from test_distiller_flow_emulated_cpu import _opt
    fix = torch.load(path, weights_only=False)
    a, b = _batch(fix)
    sl = slice(rank * PER_RANK, (rank + 1) * PER_RANK)
    with emulated_kernels(exact=True):
        from cat_b200.distillers import create_distiller
        opt = _opt(fix, os.path.join(out_dir, f'log{rank}'))
        os.makedirs(opt.log_dir, exist_ok=True)
        opt.world_size = WORLD
        model = create_distiller(opt, verbose=False)
        model.setup(opt, verbose=False)
        model.netG_teacher.load_state_dict(fix['teacher_sd'])
        model.netG_student.load_state_dict(fix['student_sd0'])
        model.netD.load_state_dict(fix['D_sd0'])
        model.netG_student.train()
        losses = []
        for it in range(2):
            model.set_input({'A': a[sl].float(), 'B': b[sl].float(), 'A_paths': ['x'] * PER_RANK, 'B_paths': ['x'] * PER_RANK})
            model.optimize_parameters(it)
            losses.append(dict(model.get_current_losses()))
        assert model.engine.world_size == WORLD
        out = {'S': {k: v.clone() for k, v in model.netG_student.state_dict().items()},
               'D': {k: v.clone() for k, v in model.netD.state_dict().items()}, 'losses': losses}
    torch.save(out, os.path.join(out_dir, f'mirror{rank}.pt'))
    dist.destroy_process_group()


@pytest.mark.timeout(900)
def test_distiller_mirror_two_ranks(golden_dir, tmp_path):
    """set_input -> optimize_parameters -> get_current_losses of the InceptionDistiller mirror on two ranks (gloo): no rank
    enters a collective alone, and after two steps both ranks hold bit-identical student and discriminator weights (same
    all-reduced gradients, same Adam) although their losses (different half batches) differ."""
    path = os.path.join(golden_dir, 'pix2pix_bn_hinge.pt')
    port = 35500 + os.getpid() % 2000
    mp.spawn(_worker_mirror, args=(port, path, str(tmp_path)), nprocs=WORLD, join=True)
    r0, r1 = (torch.load(os.path.join(tmp_path, f'mirror{r}.pt'), weights_only=False) for r in range(WORLD))
    for tag in ('S', 'D'):
        for k, v in r0[tag].items():
            if k.endswith('running_mean') or k.endswith('running_var') or k.endswith('num_batches_tracked'):
                continue            # per-rank BatchNorm statistics
            assert torch.equal(v, r1[tag][k]), (tag, k)
    for L in r0['losses'] + r1['losses']:
        assert all(v == v for v in L.values())
    assert r0['losses'][0] != r1['losses'][0]
