"""TEST/BENCH INFRASTRUCTURE -- derive the SPADE benchmark network shapes from the real reference.

Runs the reference's own pruning (utils/common.py:710-870, shrink_spade_model) on a seeded synthetic teacher with the
flags of scripts/gaugan/cityscapes/train_inception_student_5p6B.sh (teacher ngf 64, student ngf 48, target 5.6e9 MACs
profiled at the script's 256x512 crop, prune_cin_lb 16, 35 labels + edge map) and commits only the resulting
*architecture* as JSON under tests/golden/.

    python -m oracle.make_bench_arch_spade            (build container only)
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_harness_spade import build_reference_spade_distiller, multiscale_D_arch, spade_generator_arch  # noqa: E402

OUT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def main():
    model, opt = build_reference_spade_distiller(batch_size=1, crop_size=512, aspect_ratio=2.0, teacher_ngf=64,
                                                 student_ngf=48, ndf=64, input_nc=35, target_flops=5.6e9,
                                                 prune_cin_lb=16, lambda_distill=0.5)
    mm = model.modules_on_one_gpu
    out = {
        'name': 'gaugan_5p6B',
        'source': 'reference shrink_spade_model() on a seeded synthetic teacher (oracle/ref_harness_spade.py), flags of '
                  'scripts/gaugan/cityscapes/train_inception_student_5p6B.sh; profiled at 256x512',
        'teacher_arch': spade_generator_arch(mm.netG_teacher),
        'student_arch': spade_generator_arch(mm.netG_student),
        'D_arch': multiscale_D_arch(mm.netD, opt),
        'teacher_macs': int(mm.netG_teacher.n_macs),
        'student_macs': int(mm.netG_student.n_macs),
        'profiled_hw': [256, 512],
        'hp': dict(lambda_gan=float(opt.lambda_gan), lambda_feat=float(opt.lambda_feat), lambda_vgg=float(opt.lambda_vgg),
                   lambda_distill=float(opt.lambda_distill), lr_G=float(opt.lr) / 2, lr_D=float(opt.lr) * 2, beta1=0.0,
                   beta2=0.9, n_label=int(opt.input_nc)),
    }
    path = os.path.join(OUT_DIR, 'arch_gaugan_5p6B.json')
    with open(path, 'w') as f:
        json.dump(out, f, indent=1)
    print('teacher MACs %.3e student MACs %.3e' % (out['teacher_macs'], out['student_macs']))
    print({n: (b['res'], b['dw']) for n, b in out['student_arch']['blocks'].items()})


if __name__ == '__main__':
    main()
