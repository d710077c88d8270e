"""TEST/BENCH INFRASTRUCTURE -- derive the benchmark network shapes from the real reference.

Runs the reference's own pruning (utils/common.py:315-707, shrink) on the seeded synthetic teacher of
SURVEY.md 8(d) with the flags of the published training scripts, and commits only the resulting
*architecture* (channel counts) as JSON under tests/golden/: the GPU box has no /root/reference, and
the benchmark uses random-init weights of exactly these shapes.

    python -m oracle.make_bench_arch            (build container only)
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_harness import build_reference_distiller, discriminator_arch, generator_arch  # noqa: E402

OUT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')

CONFIGS = {
    # BASELINE.json configs[1] (and [4] at 256x512): scripts/pix2pix/cityscapes/train_inception_student_5p6B.sh
    'pix2pix_5p6B': dict(norm='batch', gan_mode='hinge', dataset_mode='aligned', teacher_ngf=64, student_ngf=32,
                         ndf=128, target_flops=5.6e9, prune_cin_lb=16, lambda_distill=0.5, lambda_recon=100.0,
                         height=256, width=256),
    # BASELINE.json configs[2]: scripts/cycle_gan/horse2zebra/train_inception_student_2p6B.sh
    'cyclegan_2p6B': dict(norm='instance', gan_mode='lsgan', dataset_mode='unaligned', teacher_ngf=64,
                          student_ngf=20, ndf=64, target_flops=2.6e9, prune_cin_lb=16, lambda_distill=1.0,
                          lambda_recon=5.0, height=256, width=256),
}


def main():
    only = sys.argv[1:]
    for name, cfg in CONFIGS.items():
        if only and name not in only:
            continue
        model, opt = build_reference_distiller(batch_size=1, **cfg)
        d_in = 6 if cfg['dataset_mode'] == 'aligned' else 3
        out = {
            'name': name,
            'source': 'reference shrink() on the seeded synthetic teacher (oracle/ref_harness.py), flags of the '
                      'published script; profiled at %dx%d' % (cfg['height'], cfg['width']),
            'teacher_arch': generator_arch(model.netG_teacher, opt),
            'student_arch': generator_arch(model.netG_student, opt),
            'D_arch': discriminator_arch(model.netD, opt, d_in),
            'teacher_macs': int(model.netG_teacher.n_macs),
            'student_macs': int(model.netG_student.n_macs),
            'hp': dict(gan_mode=opt.gan_mode, aligned=opt.dataset_mode == 'aligned',
                       lambda_recon=float(opt.lambda_recon), lambda_gan=float(opt.lambda_gan),
                       lambda_distill=float(opt.lambda_distill), lr=float(opt.lr), beta1=float(opt.beta1),
                       student_training=True, recon_loss_type=opt.recon_loss_type),
        }
        path = os.path.join(OUT_DIR, 'arch_%s.json' % name)
        with open(path, 'w') as f:
            json.dump(out, f, indent=1)
        print(name, 'teacher MACs %.3e student MACs %.3e widths %s' % (out['teacher_macs'], out['student_macs'],
                                                                       out['student_arch']['widths']))


if __name__ == '__main__':
    main()
