"""CUDA-event timing of the HBM-bound CUDA-core kernels on shapes of the distillation step, with the achieved fraction
of the measured HBM copy bandwidth (MEASURED_PEAKS.json: hbm_gbs) over the ALGORITHMIC bytes of each kernel.

    python tools/profile_simt.py [--iters 20]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cat_b200 import ops  # noqa: E402
from cat_b200.ops import ACT, Act  # noqa: E402

SHAPES = [(16, 64, 64, 256), (16, 128, 128, 128), (16, 256, 256, 64), (16, 64, 64, 96), (32, 33, 33, 512)]


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--iters', type=int, default=20)
    a = ap.parse_args()
    ops.require_cuda()
    dev = 'cuda:0'
    peak = 6500.0
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))['hbm_gbs']
    except (OSError, KeyError):
        pass
    print(f'HBM copy peak {peak:.0f} GB/s')
    for (N, H, W, C) in SHAPES:
        for per_sample in (False, True):
            x = Act(torch.randn(N, H, W, C, device=dev).to(torch.bfloat16))
            y = Act(torch.empty(N, H, W, C, device=dev, dtype=torch.bfloat16))
            d = Act(torch.randn(N, H, W, C, device=dev).to(torch.bfloat16))
            dx = Act(torch.empty(N, H, W, C, device=dev, dtype=torch.bfloat16))
            G = N if per_sample else 1
            sums = torch.zeros(G, 2, C, device=dev)
            red = torch.zeros(G, 2, C, device=dev)
            scale, shift = torch.ones(G, C, device=dev), torch.zeros(G, C, device=dev)
            mr = torch.zeros(G, 2, C, device=dev)
            mr[:, 1] = 1
            nb = x.t.numel() * 2
            rows = [
                ('norm_stats', 1 * nb, lambda: ops.norm_stats(x, per_sample, sums)),
                ('norm_apply+relu', 2 * nb, lambda: ops.norm_apply(x, y, scale, shift, per_sample, ACT['relu'])),
                ('norm_apply+res', 3 * nb, lambda: ops.norm_apply(x, y, scale, shift, per_sample, ACT['none'], d)),
                ('norm_bwd_reduce', 3 * nb, lambda: ops.norm_bwd_reduce(d, y, x, per_sample, mr, ACT['relu'], red)),
                ('norm_bwd_apply', 4 * nb, lambda: ops.norm_bwd_apply(d, y, x, dx, per_sample, mr, None, red, float(H * W), ACT['relu'], None, None)),
            ]
            for name, bytes_, fn in rows:
                ms = timed(fn, a.iters)
                gbs = bytes_ / ms / 1e6
                print(f'[{N}x{H}x{W}x{C} {"IN" if per_sample else "BN"}] {name:18s} {ms * 1e3:8.1f} us  {gbs:7.0f} GB/s  {100 * gbs / peak:5.1f}% of HBM peak')


if __name__ == '__main__':
    main()
