"""Context number: the REFERENCE's own InceptionDistiller.optimize_parameters (oracle/_ref, stock
PyTorch modules -> cuDNN / cuBLAS kernels) timed on the same B200 at the benchmark configuration, in fp32 (TF32 convs
allowed) and under bf16 autocast, eager (the reference's `.view` calls rule out channels_last).  This is "stock PyTorch 2.x + cuDNN on the same GPU", the bar
SURVEY.md section 2b names; it is not part of bench.py's contract (the reference arm there is the CPU path).

    python tools/torch_gpu_baseline.py [--workload pix2pix_5p6B] [--batch 16] [--steps 10] > profiles/r02_torch_gpu_baseline.json
"""
import argparse
import contextlib
import json
import os
import sys
import time

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='pix2pix_5p6B')
    ap.add_argument('--batch', type=int, default=16)
    ap.add_argument('--height', type=int, default=256)
    ap.add_argument('--width', type=int, default=256)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    a = ap.parse_args()
    os.environ['CATB_REF_ROOT'] = os.path.join(ROOT, 'oracle', '_ref')
    from cat_b200 import workload as WL
    dev = torch.device('cuda:0')
    out = {'workload': a.workload, 'batch': a.batch, 'height': a.height, 'width': a.width, 'gpu': torch.cuda.get_device_name(0),
           'torch': torch.__version__, 'cudnn': torch.backends.cudnn.version(), 'modes': {}}
    for mode in ('fp32_tf32', 'bf16_autocast'):
        with contextlib.redirect_stdout(sys.stderr):
            from oracle.make_bench_arch import CONFIGS
            from oracle.ref_harness import build_reference_distiller
            model, _opt = build_reference_distiller(batch_size=a.batch, **CONFIGS[a.workload])
        model.device = dev
        for name, v in list(vars(model).items()):
            if isinstance(v, nn.Module):
                v.to(dev)
        for v in vars(model).values():      # module lists (the 1x1 adaptors netAs carry the device the KA terms are keyed by)
            if isinstance(v, (list, tuple)):
                for m in v:
                    if isinstance(m, nn.Module):
                        m.to(dev)
        model.netG_student.train()
        torch.backends.cudnn.benchmark = True
        torch.backends.cudnn.allow_tf32 = True
        torch.backends.cuda.matmul.allow_tf32 = True
        xa, xb = WL.synthetic_batch(a.batch, a.height, a.width, 233)
        ctx = (lambda: torch.autocast('cuda', dtype=torch.bfloat16)) if mode != 'fp32_tf32' else contextlib.nullcontext

        def one():
            model.set_input({'A': xa, 'B': xb, 'A_paths': ['x'] * a.batch, 'B_paths': ['x'] * a.batch})
            with ctx():
                model.optimize_parameters(0)
        try:
            for _ in range(a.warmup):
                one()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(a.steps):
                one()
            torch.cuda.synchronize()
            ms = (time.perf_counter() - t0) * 1e3 / a.steps
            out['modes'][mode] = {'ms_per_step': ms, 'images_per_s': a.batch / (ms * 1e-3),
                                  'note': 'host inputs copied to the device inside the timed loop (set_input), eager'}
        except Exception as e:      # noqa: BLE001 -- a context number: report why a mode does not run
            out['modes'][mode] = {'error': repr(e)[:300]}
        del model
        torch.cuda.empty_cache()
    print(json.dumps(out))


if __name__ == '__main__':
    main()
