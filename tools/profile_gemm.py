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
}


def run(name, iters):
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
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    flops = 2.0 * N * OH * OW * Cout * Cin * k * k
    gathered = N * OH * OW * len(units) * 16
    print(f'{name:6s} {kind:6s} M={N * OH * OW} N={Cout} K={Cin * k * k}: {ms * 1e3:9.1f} us  {flops / ms / 1e9:8.1f} TFLOP/s  '
          f'gather {gathered / ms / 1e9:7.2f} TB/s (n_tile {g.n_tile})')


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('cases', nargs='*', default=list(CASES))
    ap.add_argument('--iters', type=int, default=20)
    a = ap.parse_args()
    ops.require_cuda()
    for c in a.cases:
        run(c, a.iters)
