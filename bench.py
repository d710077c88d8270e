"""Benchmark of the CAT distillation step (BASELINE.json metric: distill-step images/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload pix2pix_5p6B]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is one InceptionDistiller.optimize_parameters (teacher fwd, student fwd/bwd, 3 D fwd + 2 D bwd + 1 D
dgrad, KA/GAN/L1 losses, two Adam updates) on a synthetic batch of the BASELINE.json configs[1] shape:
pix2pix student pruned to 5.6e9 MACs (channel counts from the reference's own shrink(), committed under
tests/golden/), teacher ngf 64, PatchGAN ndf 128, 256x256, batch 16 per GPU.  Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='cat_b200', choices=['cat_b200', 'reference'])
    ap.add_argument('--workload', default='pix2pix_5p6B')
    ap.add_argument('--batch', type=int, default=None, help='images per GPU per step (default: 16; gaugan_5p6B: 4)')
    ap.add_argument('--height', type=int, default=None, help='default 256 (gaugan_5p6B: 512)')
    ap.add_argument('--width', type=int, default=None, help='default 256 (gaugan_5p6B: 512)')
    ap.add_argument('--cpu-batch', type=int, default=None,
                    help='images per step of the CPU arm (default: the GPU batch; gaugan_5p6B: 2 -- KA is degenerate at 1)')
    ap.add_argument('--cpu-threads', type=int, default=None, help='threads of the CPU arm (default: all host cores)')
    ap.add_argument('--nvtx', action='store_true', help='wrap the timed steps in the NVTX range "catb_step" (ncu --nvtx --nvtx-include "catb_step/")')
    ap.add_argument('--engine-e2e', action='store_true', help='end-to-end loop on the bare step engine instead of the distiller protocol')
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-parity-probe', action='store_true')
    ap.add_argument('--profile-gemms', action='store_true', help='print the per-GEMM timing table to stderr')
    args = ap.parse_args()
    spade = is_spade(args.workload)
    args.batch = args.batch or (4 if spade else (8 if args.workload.startswith('cyclegan') else 16))
    args.height = args.height or (512 if spade else 256)
    args.width = args.width or (512 if spade else 256)
    # CPU arm: the GPU arm's batch where a step stays around 5 s (pix2pix 16, CycleGAN 8); the SPADE step (~10 s for two
    # 512x512 images) is sampled at batch 2 (KA is degenerate at batch 1)
    args.cpu_batch = args.cpu_batch or (2 if spade else args.batch)
    return args


def is_spade(workload):
    return workload.startswith('gaugan')


# Teacher-training workloads: the unpruned generator of the same published configuration trained by
# Pix2PixModel / CycleGANModel / SPADEModel.optimize_parameters.  Not the headline metric -- the default workload stays
# the distillation step of BASELINE.json configs[1].
TEACHER = {'pix2pix_teacher': 'pix2pix_5p6B', 'cyclegan_teacher': 'cyclegan_2p6B', 'gaugan_teacher': 'gaugan_5p6B'}


def arch_name(workload):
    return TEACHER.get(workload, workload)


def teacher_hp(workload, arch):
    hp = dict(arch['hp'])
    if workload == 'cyclegan_teacher':      # scripts/cycle_gan/horse2zebra/train_inception_teacher.sh + CycleGANModel defaults
        return dict(gan_mode='lsgan', lambda_A=10.0, lambda_B=10.0, lambda_identity=0.5, lr=hp['lr'], beta1=hp['beta1'], pool_size=50)
    return hp


def teacher_macs(workload, arch, H, W):
    """Algorithmic MACs per image of one teacher-training step (G = the unpruned generator, MAC = 2 FLOP):
    pix2pix 3 G + 8 D (as the distillation step without T and with S = G); CycleGAN 2 generators x 3 applications x
    (fwd + dgrad + wgrad) = 18 G, 2 discriminators x (G phase: fwd + dgrad; D phase: 2 fwd + 2 full bwd) = 16 D;
    SPADE 4 G + 10 D + 3 V (as the distillation step without T and with S = G)."""
    from cat_b200 import workload as WL
    if workload == 'gaugan_teacher':
        m = WL.spade_macs_per_image(arch, H, W)
        return dict(m, step=4 * m['T'] + 10 * m['D'] + 3 * m['V'])
    m = WL.macs_per_image(arch, H, W)
    if workload == 'cyclegan_teacher':
        return dict(m, step=18 * m['T'] + 16 * m['D'])
    return dict(m, step=3 * m['T'] + 8 * m['D'])


# ----------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.samples, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        threading.Thread(target=self._read, daemon=True).start()

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append([x.strip() for x in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for r in self.samples if len(r) >= 7]
        if not rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
        sm = sorted(float(r[0]) for r in rows)
        reasons = []
        for i, name in ((3, 'hw_slowdown'), (4, 'hw_thermal_slowdown'), (5, 'sw_thermal_slowdown'), (6, 'sw_power_cap')):
            if any(r[i].lower().startswith('active') for r in rows):
                reasons.append(name)
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': float(rows[0][1]), 'reasons': reasons,
                'power_w_max': max(float(r[2]) for r in rows), 'samples': len(rows)}


# ----------------------------------------------------------------------------------------------------
# CPU baseline (the oracle port of the reference path, timed on the host cores)
# ----------------------------------------------------------------------------------------------------
CPU_THREADS = [None]


def _pin_threads():
    """The CPU arm runs on ALL host cores of the box (or --cpu-threads), stated in `cores`: no probing, so two runs on the
    same host use the same thread count."""
    n = CPU_THREADS[0] or os.cpu_count() or 1
    torch.set_num_threads(n)
    return n


def cpu_reference_steps(arch, hp, B, H, W, steps, warmup):
    """Times oracle.cat_oracle.distill_step (CPU restatement of optimize_parameters) on a bounded sample
    of the workload: same networks and resolution, `B` images per step."""
    from cat_b200 import workload as WL
    from oracle import cat_oracle as O
    _pin_threads()
    state = dict(teacher_sd=WL.init_generator(arch['teacher_arch'], 0, 'uniform'),
                 student_sd=WL.init_generator(arch['student_arch'], 1), D_sd=WL.init_discriminator(arch['D_arch'], 2),
                 teacher_arch=arch['teacher_arch'], student_arch=arch['student_arch'], D_arch=arch['D_arch'],
                 adam_G={}, adam_D={})
    a, b = WL.synthetic_batch(B, H, W, 233)
    for _ in range(warmup):
        O.distill_step(state, a, b, hp)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.distill_step(state, a, b, hp)
    dt = (time.perf_counter() - t0) / steps
    return B / dt, dt, torch.get_num_threads()


def cpu_reference_spade_steps(arch, hp, B, H, W, steps, warmup):
    """Times oracle.spade_oracle.spade_distill_step (CPU restatement of BaseSPADEDistiller.optimize_parameters) on a
    bounded sample of the workload: same networks and resolution, `B` images per step."""
    from cat_b200 import workload as WL
    from oracle import spade_oracle as SO
    _pin_threads()
    arch = WL.spade_arch_for(arch, H, W)
    state = dict(teacher_sd=WL.init_spade_reference_sd(arch['teacher_arch'], 0), student_sd=WL.init_spade_reference_sd(arch['student_arch'], 1),
                 D_sd=WL.init_multiscale_D_sd(arch['D_arch'], 2), vgg_sd=WL.init_vgg(3), teacher_arch=arch['teacher_arch'],
                 student_arch=arch['student_arch'], D_arch=arch['D_arch'], adam_G={}, adam_D={})
    lab, inst, img = WL.synthetic_spade_batch(B, H, W, hp['n_label'], 233)
    seg = SO.preprocess_input(lab, inst, hp['n_label'])
    for _ in range(warmup):
        SO.spade_distill_step(state, seg, img, hp)
    t0 = time.perf_counter()
    for _ in range(steps):
        seg = SO.preprocess_input(lab, inst, hp['n_label'])
        SO.spade_distill_step(state, seg, img, hp)
    dt = (time.perf_counter() - t0) / steps
    return B / dt, dt, torch.get_num_threads()


def cpu_teacher_steps(workload, arch, hp, B, H, W, steps, warmup):
    """Times oracle.train_oracle (CPU restatement of the reference's teacher-training optimize_parameters) on a bounded
    sample of the workload: same networks and resolution, `B` images per step."""
    from cat_b200 import workload as WL
    from oracle import train_oracle as TO
    _pin_threads()
    G_arch, D_arch = arch['teacher_arch'], arch['D_arch']
    if workload == 'gaugan_teacher':
        from oracle import spade_oracle as SO
        G_arch = dict(WL.spade_arch_for(arch, H, W)['teacher_arch'], active_fn='nn.LeakyReLU')
        state = dict(G_sd=WL.init_spade_reference_sd(G_arch, 1), D_sd=WL.init_multiscale_D_sd(D_arch, 2), vgg_sd=WL.init_vgg(3),
                     G_arch=G_arch, D_arch=D_arch, adam_G={}, adam_D={})
        lab, inst, img = WL.synthetic_spade_batch(B, H, W, hp['n_label'], 233)
        one = lambda: TO.spade_train_step(state, SO.preprocess_input(lab, inst, hp['n_label']), img, hp)
    elif workload == 'cyclegan_teacher':
        state = dict(G_A_sd=WL.init_generator(G_arch, 1), G_B_sd=WL.init_generator(G_arch, 4),
                     D_A_sd=WL.init_discriminator(D_arch, 2), D_B_sd=WL.init_discriminator(D_arch, 5), G_arch=G_arch, D_arch=D_arch,
                     adam_G={}, adam_D={}, pool_A=TO.ImagePool(hp['pool_size']), pool_B=TO.ImagePool(hp['pool_size']))
        a, b = WL.synthetic_batch(B, H, W, 233)
        one = lambda: TO.cyclegan_train_step(state, a, b, hp)
    else:
        state = dict(G_sd=WL.init_generator(G_arch, 1), D_sd=WL.init_discriminator(D_arch, 2), G_arch=G_arch, D_arch=D_arch,
                     adam_G={}, adam_D={})
        a, b = WL.synthetic_batch(B, H, W, 233)
        one = lambda: TO.pix2pix_train_step(state, a, b, hp)
    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = (time.perf_counter() - t0) / steps
    return B / dt, dt, torch.get_num_threads()


def cpu_steps_fn(workload):
    if workload in TEACHER:
        return lambda arch, hp, B, H, W, steps, warmup: cpu_teacher_steps(workload, arch, teacher_hp(workload, arch), B, H, W, steps, warmup)
    return cpu_reference_spade_steps if is_spade(workload) else cpu_reference_steps


def real_reference_available(workload):
    """The reference's own modules staged under oracle/_ref/ by oracle/make_ref.py (the Inception distiller workloads)."""
    return workload in ('pix2pix_5p6B', 'cyclegan_2p6B') and os.path.isdir(os.path.join(ROOT, 'oracle', '_ref', 'distillers'))


def real_reference_steps(workload, B, H, W, steps, warmup, budget_s):
    """Times the REFERENCE's InceptionDistiller.optimize_parameters (distillers/inception_distiller.py:179-188) on the host
    cores: the real distiller object built by oracle/ref_harness.py from oracle/_ref/ with the flags of the published
    script, its own shrink() producing the benchmark student, set_input + optimize_parameters per step.  When the
    projected run exceeds `budget_s` the per-step sample is halved (stated in the result)."""
    import contextlib
    os.environ['CATB_REF_ROOT'] = os.path.join(ROOT, 'oracle', '_ref')
    cores = _pin_threads()
    from cat_b200 import workload as WL
    with contextlib.redirect_stdout(sys.stderr):          # the reference prints its options / networks
        from oracle.make_bench_arch import CONFIGS
        from oracle.ref_harness import build_reference_distiller
        model, _opt = build_reference_distiller(batch_size=B, **CONFIGS[workload])
        model.netG_student.train()       # steady state (the reference's first step of a run is in eval mode)
        a, b = WL.synthetic_batch(B, H, W, 233)

        def one(n):
            model.set_input({'A': a[:n], 'B': b[:n], 'A_paths': ['x'] * n, 'B_paths': ['x'] * n})
            model.optimize_parameters(0)
        n = B
        while True:
            t0 = time.perf_counter()
            one(n)
            t = time.perf_counter() - t0
            if n <= 2 or t * (steps + max(warmup - 1, 0)) <= budget_s:
                break
            n = max(2, n // 2)
        for _ in range(max(warmup - 1, 0)):
            one(n)
        t0 = time.perf_counter()
        for _ in range(steps):
            one(n)
        dt = (time.perf_counter() - t0) / steps
    return n / dt, dt, cores, n


def cpu_arm(args, arch, steps, warmup, budget_s):
    """The CPU arm of the benchmark: the reference itself when it is staged (kind 'reference'), else the CPU oracle port
    (kind 'port'), on all host cores, on a per-step sample of at most --cpu-batch images at the GPU arm's resolution."""
    H, W, B = args.height, args.width, args.cpu_batch
    if real_reference_available(args.workload):
        ips, dt, cores, n = real_reference_steps(args.workload, B, H, W, steps, warmup, budget_s)
        kind, what = 'reference', "the reference's InceptionDistiller.set_input + optimize_parameters (oracle/_ref)"
    else:
        fn = cpu_steps_fn(args.workload)
        ips, dt, cores = fn(arch, dict(arch['hp']), B, H, W, steps, warmup)
        kind, n, what = 'port', B, 'CPU oracle port of optimize_parameters'
    return {'value': ips, 'unit': 'images/s', 'cores': cores, 'kind': kind, 'batch': n, 'same_config': n == args.batch,
            'sample': f'{n} images/step at {H}x{W}, {warmup} warm-up + {steps} timed steps, {dt:.2f} s/step, fp32, '
                      f'{cores} threads; {what}'}, dt


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from cat_b200 import workload as WL
    arch = WL.load_arch(arch_name(args.workload))
    cpu, dt = cpu_arm(args, arch, args.steps, args.warmup, budget_s=240.0)
    ips, B = cpu['value'], cpu['batch']
    print(json.dumps({
        'impl': 'reference', 'metric': metric_name(args.workload), 'value': ips, 'unit': 'images/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args, arch, B, 1),
        'cpu_baseline': cpu,
        'e2e': {'value': ips, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }))


def metric_name(workload):
    return 'teacher-train-step images/sec' if workload in TEACHER else 'distill-step images/sec'


def workload_config(args, arch, B, world):
    if args.workload in TEACHER:
        what = {'pix2pix_teacher': 'Pix2PixModel.optimize_parameters (generator ngf64 + PatchGAN ndf%d, hinge + L1)' % arch['D_arch']['ndf'],
                'cyclegan_teacher': 'CycleGANModel.optimize_parameters (2 generators ngf64 x 3 applications, 2 PatchGANs ndf%d, '
                                    'lsgan + cycle + identity, image pools of 50)' % arch['D_arch']['ndf'],
                'gaugan_teacher': 'SPADEModel.optimize_parameters (SPADE generator ngf64 with LeakyReLU, multi-scale spectral D, '
                                  'feature matching, VGG19 loss)'}[args.workload]
        return {'workload': f'{args.workload}: teacher training step, {what}, {args.height}x{args.width}, batch {B}/GPU',
                'global_batch': B * world, 'height': args.height, 'width': args.width, 'parallelism': f'dp{world}',
                'l2': 'per-step working set (~GBs of activations) exceeds the 126 MB L2; no explicit flush'}
    if is_spade(args.workload):
        return {'workload': f'{args.workload}: GauGAN/SPADE inception student distill step (spade_distiller: teacher ngf64 '
                            f'{arch["teacher_macs"] / 1e9:.1f} GMAC + pruned student {arch["student_macs"] / 1e9:.2f} GMAC @256x512, '
                            f'multi-scale spectral D ndf{arch["D_arch"]["ndf"]}, VGG19 loss, KA), {args.height}x{args.width}, '
                            f'batch {B}/GPU',
                'global_batch': B * world, 'height': args.height, 'width': args.width, 'parallelism': f'dp{world}',
                'l2': 'per-step working set (~GBs of activations) exceeds the 126 MB L2; no explicit flush'}
    return {'workload': f'{args.workload}: pix2pix inception student distill step (teacher ngf64 + pruned student '
                        f'{arch["student_macs"] / 1e9:.2f} GMAC + PatchGAN ndf{arch["D_arch"]["ndf"]}), '
                        f'{args.height}x{args.width}, batch {B}/GPU',
            'global_batch': B * world, 'height': args.height, 'width': args.width, 'parallelism': f'dp{world}',
            'l2': 'per-step working set (~GBs of activations) exceeds the 126 MB L2; no explicit flush'}


# ----------------------------------------------------------------------------------------------------
# per-GEMM instrumentation (roofline of the dominant kernel)
# ----------------------------------------------------------------------------------------------------
def tag_gemms(root, owner, seen=None, depth=0):
    """Mark every ops.Gemm reachable from a compiled network with the network it belongs to (per-network rooflines)."""
    from cat_b200 import ops
    seen = set() if seen is None else seen
    if id(root) in seen or depth > 6:
        return
    seen.add(id(root))
    if isinstance(root, ops.Gemm):
        root._owner = owner
        return
    if isinstance(root, (list, tuple)):
        for v in root:
            tag_gemms(v, owner, seen, depth + 1)
    elif isinstance(root, dict):
        for v in root.values():
            tag_gemms(v, owner, seen, depth + 1)
    elif hasattr(root, '__dict__') and not isinstance(root, torch.Tensor) and type(root).__module__.startswith('cat_b200'):
        for v in vars(root).values():
            tag_gemms(v, owner, seen, depth + 1)


# HBM-bound CUDA-core kernels: wrapper name -> algorithmic bytes per element of the activation they stream (DESIGN.md 3.2:
# bf16 reads + writes that cannot be avoided), evaluated on the first Act argument
HBM_KERNELS = {'norm_stats': 2, 'norm_apply': 4, 'norm_apply_fused': 4, 'norm_bwd_reduce': 6, 'norm_bwd_apply': 8, 'dwconv_fwd': 4,
               'dwconv_bwd_data': 4, 'dwconv_bwd_weight': 4, 'gram': 2, 'ka_bwd': 6, 'reflect_fold': 4, 'act_bwd': 6}


class GemmProfiler:
    """CUDA events around every GEMM launch (and every HBM-bound CUDA-core kernel of HBM_KERNELS) of ONE eager step."""

    def __init__(self):
        self.records, self.hbm = [], []

    def install(self):
        from cat_b200 import ops
        prof = self
        self._orig = (ops.Gemm.fprop, ops.Gemm.wgrad)
        self._orig_hbm = {n: getattr(ops, n) for n in HBM_KERNELS}

        def flops(g):
            if g.seg_raw is not None:     # N-concatenation: each row segment only owns the taps of its own kernel
                return sum(2.0 * g.geo.N * g.geo.OHs * g.geo.OWs * nreal * sum(w[3] for w in su.w) for (_r, _s, nreal, su) in g.seg_raw)
            nv = sum(w[3] for w in g.units.w)
            return 2.0 * g.geo.N * g.geo.OHs * g.geo.OWs * g.n_rows * nv

        def wrap(fn, kind):
            def inner(g, *a, **k):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                r = fn(g, *a, **k)
                e1.record()
                prof.records.append((kind, g, flops(g), e0, e1))
                return r
            return inner

        def wrap_hbm(fn, name, bpe):
            def inner(x, *a, **k):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                r = fn(x, *a, **k)
                e1.record()
                prof.hbm.append((name, float(x.pixels) * x.C * bpe, e0, e1))
                return r
            return inner
        ops.Gemm.fprop, ops.Gemm.wgrad = wrap(self._orig[0], 'fprop'), wrap(self._orig[1], 'wgrad')
        for n, bpe in HBM_KERNELS.items():
            setattr(ops, n, wrap_hbm(self._orig_hbm[n], n, bpe))

    def remove(self):
        from cat_b200 import ops
        ops.Gemm.fprop, ops.Gemm.wgrad = self._orig
        for n, fn in self._orig_hbm.items():
            setattr(ops, n, fn)

    def summary(self, reps=2):
        torch.cuda.synchronize()
        agg, per, net = {}, {}, {}
        # `reps` identical steps were recorded back to back: keep the faster timing of every launch
        def fold(recs, t0, t1):
            n = len(recs) // reps
            if n == 0 or len(recs) != n * reps:
                return [(r, r[t0].elapsed_time(r[t1])) for r in recs]
            return [(recs[i], min(recs[i + k * n][t0].elapsed_time(recs[i + k * n][t1]) for k in range(reps))) for i in range(n)]
        def variant(kind, g):
            if kind != 'fprop':
                return 'halo' if getattr(g, 'w_halo', None) is not None else 'v1'
            if g.halo is None or g.choice == 'v1':
                return 'v1'
            return f"{('v2', 'v3-cpasync', 'v3-tma', 'v3-tma-reflect')[g.h_mode]} TW{g.halo.TW} m{g.halo.m_sub} b{g.hdesc.b_budget // 1024}K"
        for (kind, g, fl, e0, e1), ms in fold(self.records, 3, 4):
            for table, key in ((agg, kind), (per, (kind, g.n_rows, g.n_units, g.geo.N * g.geo.OHs * g.geo.OWs, variant(kind, g))),
                               (net, (getattr(g, '_owner', 'other'), kind))):
                a = table.setdefault(key, [0.0, 0.0, 0])
                a[0] += fl
                a[1] += ms
                a[2] += 1
        hbm = {}
        for (name, nbytes, e0, e1), ms in fold(self.hbm, 2, 3):
            a = hbm.setdefault(name, [0.0, 0.0, 0])
            a[0] += nbytes
            a[1] += ms
            a[2] += 1
        return agg, per, net, hbm


def parity_probe(workload, dev):
    """One step of the benchmarked library on the small committed fixture of the workload's distiller against the CPU
    oracle; returns the measured deviations (and raises if they exceed the tolerances of the GPU test-suite)."""
    from oracle import cat_oracle as O
    out = {}
    if is_spade(workload) or workload == 'gaugan_teacher':
        from cat_b200 import ops
        from cat_b200.spade_distill_engine import SpadeDistillStep
        from oracle import spade_oracle as SO
        fix = torch.load(os.path.join(ROOT, 'tests', 'golden', 'spade_more.pt'), weights_only=False)
        s, hp = fix['steps'][0], fix['hp']
        vgg = SO.make_vgg_sd(fix['vgg_seed'])
        B, _, H, W = s['image'].shape
        eng = SpadeDistillStep(fix['teacher_arch'], fix['student_arch'], fix['D_arch'], hp, B, H, W, device=dev)
        eng.load(fix['teacher_sd'], fix['student_sd0'], fix['D_sd0'], vgg)
        eng.set_input(s['label'], s['instance'], s['image'])
        eng.step()
        torch.cuda.synchronize()
        got = eng.get_losses()
        state = dict(teacher_sd=O.clone_sd(fix['teacher_sd']), student_sd=O.clone_sd(fix['student_sd0']), D_sd=O.clone_sd(fix['D_sd0']),
                     vgg_sd=vgg, teacher_arch=fix['teacher_arch'], student_arch=fix['student_arch'], D_arch=fix['D_arch'],
                     adam_G={}, adam_D={})
        seg = SO.preprocess_input(s['label'], s['instance'], hp['n_label'])
        ref = SO.spade_distill_step(state, seg, s['image'], hp)
        pairs = (('G_gan', 'loss_G_gan'), ('G_feat', 'loss_G_feat'), ('G_vgg', 'loss_G_vgg'), ('G_distill', 'loss_G_distill'),
                 ('D_fake', 'loss_D_fake'), ('D_real', 'loss_D_real'))
        out['onehot_edges_bit_exact'] = bool(torch.equal(ops.nhwc_to_nchw(eng.seg, eng.snc).cpu(), seg))
        tol = 3e-2
    else:
        from cat_b200 import ops
        from cat_b200.distill_engine import DistillStep
        name = 'cyclegan_in_lsgan' if workload.startswith('cyclegan') else 'pix2pix_bn_hinge'
        path = os.path.join(ROOT, 'tests', 'golden', name + '.pt')
        if not os.path.exists(path):
            path = os.path.join(ROOT, 'tests', 'golden', 'pix2pix_bn_hinge.pt')
        fix = torch.load(path, weights_only=False)
        s = fix['steps'][0]
        B, _, H, W = s['real_A'].shape
        eng = DistillStep(fix['teacher_arch'], fix['student_arch'], fix['D_arch'], fix['hp'], B, H, W, device=dev)
        eng.load(fix['teacher_sd'], fix['student_sd0'], fix['D_sd0'])
        eng.set_input(s['real_A'], s['real_B'])
        eng.step()
        torch.cuda.synchronize()
        got = eng.get_losses()
        state = dict(teacher_sd=O.clone_sd(fix['teacher_sd']), student_sd=O.clone_sd(fix['student_sd0']), D_sd=O.clone_sd(fix['D_sd0']),
                     teacher_arch=fix['teacher_arch'], student_arch=fix['student_arch'], D_arch=fix['D_arch'], adam_G={}, adam_D={})
        ref = O.distill_step(state, s['real_A'], s['real_B'], fix['hp'])
        sfake = ops.nhwc_to_nchw(eng.S.out, 3).cpu()
        out['student_output_rel_l2'] = float((sfake - ref['Sfake_B']).norm() / ref['Sfake_B'].norm())
        assert out['student_output_rel_l2'] < 3e-2, out
        pairs = (('G_recon', 'loss_G_recon'), ('G_gan', 'loss_G_gan'), ('G_distill', 'loss_G_distill'), ('D_fake', 'loss_D_fake'),
                 ('D_real', 'loss_D_real'))
        tol = 2e-2
    worst = 0.0
    for mine, theirs in pairs:
        r = float(ref[theirs])
        worst = max(worst, abs(got[mine] - r) / max(1.0, abs(r)))
    out['max_loss_deviation'] = worst
    out['tolerance'] = tol
    out['fixture'] = os.path.basename(path) if not (is_spade(workload) or workload == 'gaugan_teacher') else 'spade_more.pt'
    assert worst <= tol, out
    return out


def main():
    args = parse()
    if args.impl == 'reference':
        return run_reference(args)

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = f'cuda:{local}'
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device(dev))
    from cat_b200 import _C, ops
    from cat_b200 import workload as WL
    from cat_b200.distill_engine import DistillStep
    from cat_b200.engine import GenNet
    ops.require_cuda()      # raises without an sm_100 device / libcatb200.so (sets the kernels' shared-memory attributes)

    arch = WL.load_arch(arch_name(args.workload))
    hp = dict(arch['hp'])
    B, H, W = args.batch, args.height, args.width
    spade = is_spade(args.workload)
    if args.workload in TEACHER:
        eng, host, h2d_bytes = build_teacher(args, arch, teacher_hp(args.workload, arch), B, H, W, dev, world, rank)
        macs = teacher_macs(args.workload, arch, H, W)
    elif spade:
        from cat_b200.spade_distill_engine import SpadeDistillStep
        from cat_b200.spade_engine import SpadeGenNet
        sarch = WL.spade_arch_for(arch, H, W)
        hp['ka_scale'] = 1.0        # the SPADE distiller averages the per-replica losses (spade_model.py:191)
        eng = SpadeDistillStep(sarch['teacher_arch'], sarch['student_arch'], sarch['D_arch'], hp, B, H, W, device=dev,
                               world_size=world, use_cuda_graph=not args.no_graph)
        host = WL.synthetic_spade_batch(B, H, W, hp['n_label'], 233 + rank, pin=True)
        # synthetic "trained" teacher: running statistics calibrated on one synthetic batch (momentum 1)
        eng.set_input(*host)
        eng._preprocess()
        cal = SpadeGenNet(dict(sarch['teacher_arch'], momentum=1.0), eng.seg, dev, training=True, need_grad=False)
        cal.load_state_dict(WL.init_from_entries(cal, 0, 'uniform'))
        cal.forward()
        torch.cuda.synchronize()
        t_sd = cal.state_dict()
        del cal
        eng.load(t_sd, WL.init_from_entries(eng.S, 1), WL.init_from_entries(eng.D, 2), WL.init_vgg(3))
        h2d_bytes = sum(t.numel() * t.element_size() for t in host)
        macs = WL.spade_macs_per_image(arch, H, W)
    else:
        hp['ka_scale'] = float(world)   # the reference sums the per-replica KA terms (inception_distiller.py:145-148)
        if args.engine_e2e:
            eng, host, h2d_bytes, macs = build_pix2pix(args, arch, hp, B, H, W, dev, world, rank)
        else:
            model, eng, host, h2d_bytes, macs = build_inception_distiller(args, arch, hp, B, H, W, dev, world, rank, local)
            return run_job(args, arch, eng, host, h2d_bytes, macs, B, H, W, dev, world, rank, local, model=model)
    return run_job(args, arch, eng, host, h2d_bytes, macs, B, H, W, dev, world, rank, local)


def build_inception_distiller(args, arch, hp, B, H, W, dev, world, rank, local):
    """The user-facing path: cat_b200.distillers.InceptionDistiller driven like trainer.py:79-175 (create_distiller ->
    setup -> set_input -> optimize_parameters -> get_current_losses), with the published script's options, the pruned
    student architecture of the benchmark and a synthetic trained teacher restored through --restore_teacher_G_path."""
    import tempfile
    from cat_b200 import ops
    from cat_b200 import workload as WL
    from cat_b200.distillers import create_distiller
    from cat_b200.engine import GenNet
    Ta, Da = arch['teacher_arch'], arch['D_arch']
    t_sd = WL.init_generator(Ta, 0, 'uniform')
    a_host, b_host = WL.synthetic_batch(B, H, W, 233 + rank, pin=True)
    if Ta['norm'] == 'batch' and Ta['track_running_stats']:
        # synthetic "trained" teacher: running statistics calibrated on one synthetic batch
        cal = GenNet(dict(Ta, momentum=1.0), B, H, W, dev, training=True, need_grad=False)
        cal.load_state_dict(t_sd)
        xa = ops.Act.empty(B, H, W, 3, dev, zero=True)
        ops.nchw_to_nhwc(a_host.to(dev), xa)
        cal.forward(xa)
        torch.cuda.synchronize()
        t_sd = {k: v.cpu() for k, v in cal.state_dict().items()}
        del cal, xa
    work = tempfile.mkdtemp(prefix='catb_bench_')
    tpath = os.path.join(work, 'teacher_net_G.pth')
    torch.save(t_sd, tpath)
    opt = argparse.Namespace(
        isTrain=True, gpu_ids=[local], log_dir=work, distiller='inception', input_nc=3, output_nc=3,
        teacher_ngf=Ta['widths'][0], student_ngf=arch['student_arch']['widths'][0], teacher_netG='inception_9blocks',
        student_netG='inception_9blocks', norm=Ta['norm'], norm_affine=Ta['affine'], norm_affine_D=Da['affine'],
        norm_track_running_stats=Ta['track_running_stats'], norm_momentum=Ta['momentum'], norm_epsilon=Ta['eps'], channels=None,
        channels_reduction_factor=6, kernel_sizes=list(Ta['kernel_sizes']), active_fn='nn.ReLU', active_fn_D='nn.LeakyReLU',
        init_type='normal', init_gain=0.02, netD='n_layers', ndf=Da['ndf'], n_layers_D=Da['n_layers'],
        dataset_mode='aligned' if hp['aligned'] else 'unaligned', direction='AtoB', gan_mode=hp['gan_mode'],
        recon_loss_type=hp.get('recon_loss_type', 'l1'), distill_G_loss_type='ka', lambda_distill=hp['lambda_distill'],
        lambda_recon=hp['lambda_recon'], lambda_gan=hp['lambda_gan'], lr=hp['lr'], beta1=hp['beta1'], nepochs=5, nepochs_decay=15,
        student_arch=arch['student_arch'], restore_teacher_G_path=tpath, restore_student_G_path=None, restore_D_path=None,
        restore_A_path=None, restore_O_path=None, cuda_graph=not args.no_graph, world_size=world)
    model = create_distiller(opt, verbose=False)
    model.setup(opt, verbose=False)
    model.netG_student.load_state_dict(WL.init_generator(arch['student_arch'], 1))    # reference-style init, fixed seeds
    model.netD.load_state_dict(WL.init_discriminator(Da, 2))
    model.netG_student.train()          # steady state: the reference's eval-mode first step is over
    host = {'A': a_host, 'B': b_host, 'A_paths': ['synthetic'] * B, 'B_paths': ['synthetic'] * B}
    model.set_input(host)               # compiles the step engine for this batch shape
    return model, model.engine, host, 2 * B * 3 * H * W * 4, WL.macs_per_image(arch, H, W)


def build_teacher(args, arch, hp, B, H, W, dev, world, rank):
    """Engines of the teacher-training workloads (cat_b200/train_engine.py) with reference-style random initialisation."""
    from cat_b200 import workload as WL
    from cat_b200.train_engine import CycleGANTrainStep, Pix2PixTrainStep, SpadeTrainStep
    kw = dict(device=dev, world_size=world, use_cuda_graph=not args.no_graph)
    G_arch, D_arch = arch['teacher_arch'], arch['D_arch']
    if args.workload == 'gaugan_teacher':
        G_arch = dict(WL.spade_arch_for(arch, H, W)['teacher_arch'], active_fn='nn.LeakyReLU')    # models/spade_model.py:92
        eng = SpadeTrainStep(G_arch, D_arch, hp, B, H, W, **kw)
        host = WL.synthetic_spade_batch(B, H, W, hp['n_label'], 233 + rank, pin=True)
        eng.load(WL.init_from_entries(eng.G, 1), WL.init_from_entries(eng.D, 2), WL.init_vgg(3))
    elif args.workload == 'cyclegan_teacher':
        eng = CycleGANTrainStep(G_arch, D_arch, hp, B, H, W, **kw)
        host = WL.synthetic_batch(B, H, W, 233 + rank, pin=True)
        eng.load(WL.init_generator(G_arch, 1), WL.init_generator(G_arch, 4), WL.init_discriminator(D_arch, 2),
                 WL.init_discriminator(D_arch, 5))
    else:
        eng = Pix2PixTrainStep(G_arch, D_arch, hp, B, H, W, **kw)
        host = WL.synthetic_batch(B, H, W, 233 + rank, pin=True)
        eng.load(WL.init_generator(G_arch, 1), WL.init_discriminator(D_arch, 2))
    return eng, host, sum(t.numel() * t.element_size() for t in host)


def build_pix2pix(args, arch, hp, B, H, W, dev, world, rank):
    from cat_b200 import ops
    from cat_b200 import workload as WL
    from cat_b200.distill_engine import DistillStep
    from cat_b200.engine import GenNet
    eng = DistillStep(arch['teacher_arch'], arch['student_arch'], arch['D_arch'], hp, B, H, W, device=dev,
                      world_size=world, use_cuda_graph=not args.no_graph)
    t_sd = WL.init_generator(arch['teacher_arch'], 0, 'uniform')
    a_host, b_host = WL.synthetic_batch(B, H, W, 233 + rank, pin=True)
    if arch['teacher_arch']['norm'] == 'batch' and arch['teacher_arch']['track_running_stats']:
        # synthetic "trained" teacher: running statistics calibrated on one synthetic batch
        cal_arch = dict(arch['teacher_arch'], momentum=1.0)
        cal = GenNet(cal_arch, B, H, W, dev, training=True, need_grad=False)
        cal.load_state_dict(t_sd)
        xa = ops.Act.empty(B, H, W, 3, dev, zero=True)
        ops.nchw_to_nhwc(a_host.to(dev), xa)
        cal.forward(xa)
        torch.cuda.synchronize()
        t_sd = cal.state_dict()
        del cal
    eng.load(t_sd, WL.init_generator(arch['student_arch'], 1), WL.init_discriminator(arch['D_arch'], 2))
    return eng, (a_host, b_host), 2 * B * 3 * H * W * 4, WL.macs_per_image(arch, H, W)


def run_job(args, arch, eng, host, h2d_bytes, macs, B, H, W, dev, world, rank, local, model=None):
    """model: the distiller mirror the end-to-end loop goes through (None: the bare step engine)."""
    from cat_b200 import _C
    if world > 1:
        import torch.distributed as dist
    torch.cuda.synchronize()
    if model is not None:
        feed, one_step, read_losses = (lambda: model.set_input(host)), (lambda i: model.optimize_parameters(i)), model.get_current_losses
        through = 'cat_b200.distillers.%s: set_input -> optimize_parameters -> get_current_losses' % type(model).__name__
    else:
        feed, one_step, read_losses = (lambda: eng.set_input(*host)), (lambda i: eng.step()), eng.get_losses
        through = '%s: set_input -> step -> get_losses' % type(eng).__name__
    for owner, net in (('teacher', getattr(eng, 'T', None)), ('student', getattr(eng, 'S', None)), ('D', getattr(eng, 'D', None)),
                       ('vgg', getattr(eng, 'V', None)), ('adaptors', getattr(eng, 'A', None))):
        if net is not None:
            tag_gemms(net, owner)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value")
    feed()
    for _ in range(args.warmup):
        eng.step()
    launches0 = _C.LAUNCH_COUNT[0]
    if not args.no_graph:
        launches_per_step = eng.launches_per_step
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if args.nvtx:
        torch.cuda.nvtx.range_push('catb_step')
    e0.record()
    for _ in range(args.steps):
        eng.step()
    e1.record()
    barrier()
    if args.nvtx:
        torch.cuda.nvtx.range_pop()
    ms = e0.elapsed_time(e1)
    if args.no_graph:
        launches_per_step = (_C.LAUNCH_COUNT[0] - launches0) // args.steps
    # ---- end to end: pinned host inputs -> H2D -> step -> D2H of the losses, every step
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(args.steps):
        feed()
        one_step(i)
        losses = read_losses()
    t1.record()
    barrier()
    ms_e2e = t0.elapsed_time(t1)
    clk = clocks.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
    bad = [k for k, v in losses.items() if v != v]
    if bad:
        raise RuntimeError(f'non-finite losses after the timed steps: {bad}')

    # ---- roofline of the dominant kernel: one instrumented eager step (CUDA events around every GEMM launch)
    roof = None
    if rank == 0:
        prof = GemmProfiler()
        prof.install()
        eng.use_cuda_graph = False
        ws, eng.world_size = eng.world_size, 1   # rank 0 only: this extra step must not enter a collective
        # per-kernel timing needs the kernels alone on the GPU: no side-stream branches in this step
        saved = {k: getattr(eng, k) for k in ('overlap', 'overlap_teacher') if hasattr(eng, k)}
        gens = [g for g in (getattr(eng, n, None) for n in ('S', 'GA_real', 'GA_cyc', 'GA_idt', 'GB_real', 'GB_cyc', 'GB_idt'))
                if g is not None]
        saved_w = [g.overlap_wgrad for g in gens]
        for k in saved:
            setattr(eng, k, False)
        for g in gens:
            g.overlap_wgrad = False
        eng.step()
        eng.step()      # twice: every launch is reported at the faster of its two timings (one eager step is noisy --
        #                 the same kernel variant on the same tensors was seen at 93 and 176 us in consecutive sessions)
        for k, v in saved.items():
            setattr(eng, k, v)
        for g, v in zip(gens, saved_w):
            g.overlap_wgrad = v
        eng.world_size = ws
        agg, per, net, hbm_k = prof.summary()
        prof.remove()
        eng.use_cuda_graph = not args.no_graph
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except OSError:
            pass
        peak = peaks.get('bf16_tflops_sustained', 1400.0)
        peak_src = 'MEASURED_PEAKS.json bf16_tflops_sustained' if peaks else 'fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)'
        fl, tms, n = agg['fprop']
        achieved = fl / (tms * 1e-3) / 1e12
        step_ms_eager = sum(v[1] for v in agg.values())
        # DRAM traffic of the same kernel family from the committed ncu launch list of this workload
        # (profiles/traffic_<workload>.json, tools/summarize_launches.py): bytes per launch, averaged over the step
        traffic, traffic_src = None, None
        for tname in (f'r02_traffic_{args.workload}.json', f'traffic_{args.workload}.json'):     # newest capture first
            try:
                tj = json.load(open(os.path.join(ROOT, 'profiles', tname)))
                ks = [v for k, v in tj['kernels'].items() if k in ('igemm_halo_fprop_kernel', 'igemm_halo_persist_kernel', 'igemm_fprop_kernel')]
                traffic = sum(v['dram_read'] + v['dram_write'] for v in ks) / max(1, sum(v['launches'] for v in ks))
                traffic_src = f'profiles/{tname} (ncu dram__bytes_read.sum + dram__bytes_write.sum, mean per launch of the family)'
                break
            except (OSError, KeyError, ValueError):
                continue
        roof = {'bound': 'tensor', 'kernel': 'igemm_halo_persist_kernel + igemm_halo_fprop_kernel + igemm_fprop_kernel (tcgen05 implicit-GEMM conv/dgrad)',
                'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak, 'traffic': traffic,
                'traffic_source': traffic_src, 'algorithmic_flop_per_launch': fl / n,
                'peak_source': peak_src, 'launches': n, 'kernel_ms_per_step': tms,
                'share_of_gemm_time': tms / step_ms_eager,
                'wgrad': {'achieved': agg['wgrad'][0] / (agg['wgrad'][1] * 1e-3) / 1e12, 'launches': agg['wgrad'][2],
                          'kernel_ms_per_step': agg['wgrad'][1]} if 'wgrad' in agg else None}
        # per network (never a single blended number): the frozen teacher's forward GEMMs, the student's
        # forward + input-gradient GEMMs, the discriminator's, and every weight gradient, each against the tensor peak
        per_net = {}
        for (owner, kind), (f_, t_, c_) in sorted(net.items()):
            per_net['%s.%s' % (owner, kind)] = {'achieved': f_ / (t_ * 1e-3) / 1e12, 'frac': f_ / (t_ * 1e-3) / 1e12 / peak,
                                                'launches': c_, 'kernel_ms_per_step': t_}
        roof['per_network'] = per_net
        # HBM-bound CUDA-core kernels against the measured copy bandwidth
        hbm_peak = peaks.get('hbm_gbs', 6500.0)
        roof['hbm_kernels'] = {'peak': hbm_peak, 'unit': 'GB/s',
                               'peak_source': 'MEASURED_PEAKS.json hbm_gbs' if peaks else 'fallback 6.5 TB/s',
                               'kernels': {k: {'achieved': b_ / (t_ * 1e-3) / 1e9, 'frac': b_ / (t_ * 1e-3) / 1e9 / hbm_peak,
                                               'launches': c_, 'kernel_ms_per_step': t_,
                                               'algorithmic_bytes_per_launch': b_ / c_}
                                           for k, (b_, t_, c_) in sorted(hbm_k.items())}}
        if args.profile_gemms:
            for k, v in per_net.items():
                print(f'network {k:18s} {v["launches"]:4d} launches {v["kernel_ms_per_step"]:8.3f} ms {v["achieved"]:8.1f} TF/s', file=sys.stderr)
            for k, v in roof['hbm_kernels']['kernels'].items():
                print(f'hbm     {k:18s} {v["launches"]:4d} launches {v["kernel_ms_per_step"]:8.3f} ms {v["achieved"]:8.1f} GB/s', file=sys.stderr)
            rows = sorted(per.items(), key=lambda kv: -kv[1][1])[:60]
            for (kind, n_rows, n_units, M, var), (f, t, c) in rows:
                print(f'{kind:6s} rows {n_rows:5d} units {n_units:5d} M {M:8d} x{c:3d}  {t:8.3f} ms  {f / (t * 1e-3) / 1e12:8.1f} TF/s  {var}',
                      file=sys.stderr)

    # ---- CPU baseline beside it (rank 0, single-GPU run only): a bounded sample (~30 s) of the same workload
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu, _dt = cpu_arm(args, arch, 2 if is_spade(args.workload) else 3, 1, budget_s=30.0)

    # ---- the benchmarked binary is the checked one: one step of the same engine classes on two images of the committed
    # fixture against the CPU oracle (stated tolerances of tests/test_distill_gpu.py)
    parity = None
    if rank == 0 and not args.no_parity_probe:
        parity = parity_probe(args.workload, dev)

    if rank == 0:
        imgs = B * world * args.steps
        value = imgs / (ms * 1e-3)
        print(json.dumps({
            'metric': metric_name(args.workload), 'value': value, 'unit': 'images/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
            'config': workload_config(args, arch, B, world),
            'e2e': {'value': imgs / (ms_e2e * 1e-3), 'unit': 'images/s',
                    'h2d_bytes_per_step': h2d_bytes, 'd2h_bytes_per_step': (16 + 4) * 4, 'through': through},
            'gpu_launches': launches_per_step * args.steps * 2 + launches_per_step,
            'launches_per_step': launches_per_step,
            'algorithmic_gflop_per_image': 2 * macs['step'] / 1e9,
            'model_tflops': 2 * macs['step'] * value / 1e12,
            'clocks': clk, 'roofline': roof, 'cpu_baseline': cpu, 'parity_probe': parity,
            'losses': {k: round(float(v), 5) for k, v in losses.items()},
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
