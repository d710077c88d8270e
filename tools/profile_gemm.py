"""Stand-alone launcher of single implicit-GEMM shapes from the distillation step (for ncu and quick
CUDA-event timing on the GPU box).

    python tools/profile_gemm.py t5x5 d4 head d4w --iters 20
    ncu --set full --import-source on -k regex:igemm -s 2 -c 2 -o gpurun_out/t5x5 python tools/profile_gemm.py t5x5 --iters 4
"""
import argparse
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cat_b200 import igemm_plan as P  # noqa: E402
from cat_b200 import ops  # noqa: E402

# name: (kind, Cin, Cout, k, stride, pad, reflect, N, H, W)
CASES = {
    't5x5': ('fprop', 256, 42, 5, 1, 2, True, 16, 64, 64),      # teacher block, first 5x5 conv
    't3x3': ('fprop', 256, 42, 3, 1, 1, True, 16, 64, 64),
    't1x1': ('fprop', 256, 42, 1, 1, 0, False, 16, 64, 64),
    't1c': ('fprop', 256, 192, 1, 1, 0, False, 16, 64, 64),     # teacher block, N-concatenated 1x1 first convs
    't5c': ('fprop', 256, 128, 5, 1, 2, True, 16, 64, 64),      # stand-in: all first convs embedded in one 5x5, N = 126 -> 128
    'ts2': ('fprop2', 252, 256, 5, 1, 2, True, 16, 64, 64),     # stand-in for the K-concatenated stage 2
    'd1': ('fprop', 128, 256, 4, 2, 1, False, 16, 128, 128),
    'd2': ('fprop', 256, 512, 4, 2, 1, False, 16, 64, 64),
    'd4': ('fprop', 512, 1024, 4, 1, 1, False, 16, 32, 32),     # PatchGAN 512 -> 1024
    'd5': ('fprop', 1024, 1, 4, 1, 1, False, 16, 31, 31),
    'head': ('fprop', 64, 3, 7, 1, 3, True, 16, 256, 256),      # teacher head 7x7
    'stem': ('fprop', 3, 64, 7, 1, 3, True, 16, 256, 256),
    'd4w': ('wgrad', 512, 1024, 4, 1, 1, False, 16, 32, 32),
    'd2w': ('wgrad', 256, 512, 4, 2, 1, False, 16, 64, 64),
    's5w': ('wgrad', 90, 16, 5, 1, 2, True, 16, 64, 64),        # student-sized wgrad
    # thin GEMMs (HBM bound by rights): student / SPADE full-resolution layers
    'shead': ('fprop', 17, 3, 7, 1, 3, True, 16, 256, 256),     # student head 7x7
    'sstem': ('fprop', 3, 25, 7, 1, 3, True, 16, 256, 256),
    'thin3': ('fprop', 8, 64, 3, 1, 1, False, 4, 512, 512),     # SPADE: 3x3, 8 -> 64 at full resolution
    'thin1': ('fprop', 36, 21, 1, 1, 0, False, 4, 512, 512),    # SPADE gamma/beta first 1x1 on the label map
    'thin5': ('fprop', 36, 21, 5, 1, 2, False, 4, 512, 512),
    's1x1': ('fprop', 62, 90, 1, 1, 0, False, 16, 64, 64),      # student block, fused 1x1 first convs
    's3x3': ('fprop', 62, 15, 3, 1, 1, True, 16, 64, 64),
    'vgg1': ('fprop', 64, 64, 3, 1, 1, False, 4, 512, 512),     # VGG conv1_2
    'vgg3': ('fprop', 256, 256, 3, 1, 1, False, 4, 128, 128),   # VGG conv3_x
}


def run(name, iters, variants=False, tiling=None, timeline=False):
    kind, Cin, Cout, k, stride, pad, reflect, N, H, W = CASES[name]
    dev = 'cuda:0'
    OH = (H + 2 * pad - k) // stride + 1
    OW = (W + 2 * pad - k) // stride + 1
    w = torch.randn(Cout, Cin, k, k) / math.sqrt(Cin * k * k)
    arena = torch.cat([torch.zeros(8), w.flatten()]).to(dev)
    units = P.conv_fprop_units(8, Cout, Cin, k, k, pad)
    geo = P.Geometry(N, H, W, P.cpad(Cin), 0, OH, OW, P.cpad(Cout), 0, sn=stride,
                     pad_mode=P.PAD_REFLECT if reflect else P.PAD_ZERO)
    x = torch.randn(N, H, W, P.cpad(Cin), device=dev).to(torch.bfloat16)
    y = torch.randn(N, OH, OW, P.cpad(Cout), device=dev).to(torch.bfloat16)
    g = ops.Gemm(geo, units, Cout, dev, need_pack=kind != 'wgrad')
    grad = torch.zeros_like(arena)
    if kind != 'wgrad':
        g.pack(arena)
        fn = lambda: g.fprop(x, y)  # noqa: E731
    else:
        fn = lambda: g.wgrad(x, y, grad)  # noqa: E731
    if tiling is not None and kind != 'wgrad':      # pin one v2 variant (for ncu): 'TW,m_sub,budgetK'
        tw, ms_, bk, PIN_MODE = ([int(v) for v in tiling.split(',')] + [0])[:4]   # mode: 0 v2, 1 v3 cp.async, 2 v3 TMA
        ops.AUTOTUNE = False
        pick = [t for t in g.tilings if (t[0], t[1], t[2].b_budget // 1024) == (tw, ms_, bk) and t[4] == PIN_MODE]
        assert pick, [(t[0], t[1], t[2].b_budget // 1024, t[4]) for t in g.tilings]
        g._use_tiling(pick[0])

    def timed(f):
        for _ in range(2):
            f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            f()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    if timeline and kind != 'wgrad':
        import ctypes as C
        from cat_b200 import _C
        if tiling is None:          # let the autotune pick, then insist on the halo kernel (the timeline is its instrumentation)
            fn()
            torch.cuda.synchronize()
            g.choice = 'v2'
            if g.packed is None:
                g.packed = torch.zeros(_C.load().catb_packed_weight_bytes(Cout, len(g.f_units), g.n_tile), dtype=torch.uint8, device=dev)
                g.pack(arena)
            print(f'{name}: tiling TW{g.halo.TW} m{g.halo.m_sub} b{g.hdesc.b_budget // 1024}K')
        buf = torch.zeros(4096 * 16, dtype=torch.int64, device=dev)
        fn()
        torch.cuda.synchronize()
        _C.load().catb_debug_timeline(C.c_void_p(buf.data_ptr()))
        fn()
        torch.cuda.synchronize()
        _C.load().catb_debug_timeline(None)
        t = buf.view(4096, 16).cpu()
        t = t[t[:, 0] > 0]
        t0 = t[:, 7].min()
        rel = (t[:, :7] - t[:, 7:8]).float() / 1e3
        names = ['prologue done', 'fill done', 'mma start', 'last mma issued', 'accum complete', 'epilogue done', 'exit']
        print(f'{name}: {t.shape[0]} CTAs recorded; kernel span {(t[:, 6].max() - t0).item() / 1e3:.1f} us; per-CTA phase times (us after kernel entry) median / p90:')
        for i, nme in enumerate(names):
            col = rel[:, i]
            print(f'    {nme:18s} {col.median().item():8.2f} {col.quantile(0.9).item():8.2f}')
        print('    epilogue cycles of thread 0 (median): TMEM loads %d, pack + staging %d, copy-out %d' % tuple(int(t[:, 8 + q].float().median().item()) for q in range(3)))
        start = (t[:, 7] - t0).float() / 1e3
        srt = start.sort().values
        print('    CTA entry times (us): sorted every 10%', [round(srt[int(q * (len(srt) - 1) / 10)].item(), 1) for q in range(11)])
        return

    flops = 2.0 * N * OH * OW * Cout * Cin * k * k
    hbm_us = (x.numel() + y.numel()) * 2 / 6.5e12 * 1e6          # read x once, write y once at ~6.5 TB/s
    if variants and kind != 'wgrad':
        ops.AUTOTUNE = False
        rows = [('v1', timed(lambda: g.fprop(x, y, force_v1=True)))]
        for t in (g.tilings if g.halo is not None else []):
            g._use_tiling(t)
            rows.append((f'{("v2", "v3-cpasync", "v3-tma", "v3-tma-reflect")[t[4]]} TW{t[0]} m{t[1]} b{t[2].b_budget // 1024}K', timed(lambda: g.fprop(x, y))))
        best = min(r[1] for r in rows)
        print(f'{name:6s} M={N * OH * OW} N={Cout} K={Cin * k * k} n_tile {g.n_tile}  HBM floor {hbm_us:6.1f} us  best {best * 1e3:7.1f} us '
              f'({flops / best / 1e9:6.1f} TF/s, {hbm_us / (best * 1e3) * 100:4.1f}% of the HBM floor rate)')
        for nme, ms in rows:
            print(f'        {nme:22s} {ms * 1e3:9.1f} us')
        return
    ms = timed(fn)
    gathered = N * OH * OW * len(units) * 16
    print(f'{name:6s} {kind:6s} M={N * OH * OW} N={Cout} K={Cin * k * k}: {ms * 1e3:9.1f} us  {flops / ms / 1e9:8.1f} TFLOP/s  '
          f'gather {gathered / ms / 1e9:7.2f} TB/s (n_tile {g.n_tile})  HBM floor {hbm_us:6.1f} us')


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('cases', nargs='*', default=list(CASES))
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--tiling', default=None, help="pin a halo variant: 'TW,m_sub,budgetK[,mode]' (mode 0 v2, 1 v3 persistent + cp.async, 2 v3 persistent + TMA)")
    ap.add_argument('--timeline', action='store_true', help='per-CTA phase timestamps of the halo kernel')
    ap.add_argument('--dbg-mode', type=int, default=0, help='catb_debug_mode bits (epilogue experiments)')
    ap.add_argument('--dbg-sweep', default=None, help='comma separated catb_debug_mode values, run one after the other')
    ap.add_argument('--variants', action='store_true', help='time v1 and every v2 tiling / weight-ring variant')
    a = ap.parse_args()
    ops.require_cuda()
    from cat_b200 import _C as _CC
    for mode in ([a.dbg_mode] if a.dbg_sweep is None else [int(v) for v in a.dbg_sweep.split(',')]):
        _CC.load().catb_debug_mode(mode)
        if a.dbg_sweep is not None:
            print(f'--- catb_debug_mode {mode}')
        for c in a.cases:
            run(c, a.iters, a.variants, a.tiling, a.timeline)
    _CC.load().catb_debug_mode(0)
