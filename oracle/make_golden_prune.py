"""TEST INFRASTRUCTURE ONLY -- golden cases for cat_b200/prune.py from the real reference `shrink` (build container only).

    python -m oracle.make_golden_prune

Per case: the teacher's normalisation scales (the only weights the search reads), its architecture, the options, and the
student architecture + MAC counts the reference produced.  A few thousand floats per case -> tests/golden/prune_cases.pt."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_harness import build_reference_distiller, generator_arch  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'prune_cases.pt')

CASES = {
    'small_bn': dict(norm='batch', teacher_ngf=16, student_ngf=8, ndf=8, height=64, width=64, prune_cin_lb=4, frac=0.3, seed=0),
    'small_in': dict(norm='instance', teacher_ngf=12, student_ngf=8, ndf=8, height=32, width=48, prune_cin_lb=2, frac=0.2, seed=3),
    'small_in_half': dict(norm='instance', teacher_ngf=24, student_ngf=8, ndf=8, height=64, width=64, prune_cin_lb=8, frac=0.5, seed=5),
    # the published scripts (BASELINE configs[1] / [2]); the resulting architectures are tests/golden/arch_*.json
    'pix2pix_5p6B': dict(norm='batch', teacher_ngf=64, student_ngf=32, ndf=128, height=256, width=256, prune_cin_lb=16,
                         target_flops=5.6e9, seed=0, gan_mode='hinge', dataset_mode='aligned'),
    'cyclegan_2p6B': dict(norm='instance', teacher_ngf=64, student_ngf=20, ndf=64, height=256, width=256, prune_cin_lb=16,
                          target_flops=2.6e9, seed=0, gan_mode='lsgan', dataset_mode='unaligned'),
}


def main():
    out = {}
    for name, cfg in CASES.items():
        cfg = dict(cfg)
        frac = cfg.pop('frac', None)
        if frac is not None:
            probe, _ = build_reference_distiller(batch_size=1, do_shrink=False, **cfg)
            cfg['target_flops'] = probe.netG_teacher.n_macs * frac
        model, opt = build_reference_distiller(batch_size=1, **cfg)
        tsd = model.netG_teacher.state_dict()
        keep = {k: v.detach().clone() for k, v in tsd.items()
                if k.endswith('.weight') and v.dim() == 1}
        out[name] = {'gammas': keep, 'teacher_arch': generator_arch(model.netG_teacher, opt),
                     'student_arch': generator_arch(model.netG_student, opt), 'teacher_macs': int(model.netG_teacher.n_macs),
                     'student_macs': int(model.netG_student.n_macs), 'target_flops': float(cfg['target_flops']),
                     'H': cfg['height'], 'W': cfg['width'], 'prune_cin_lb': cfg['prune_cin_lb']}
        print(name, out[name]['teacher_macs'], out[name]['student_macs'], out[name]['student_arch']['widths'])
    torch.save(out, OUT)
    print('->', OUT, os.path.getsize(OUT))


SPADE_OUT = os.path.join(os.path.dirname(OUT), 'prune_spade_cases.pt')
SPADE_CASES = {
    'spade_small': dict(crop_size=128, aspect_ratio=2.0, teacher_ngf=6, student_ngf=6, ndf=8, input_nc=6, prune_cin_lb=2, frac=0.4),
    'spade_small_b': dict(crop_size=128, aspect_ratio=1.0, teacher_ngf=8, student_ngf=8, ndf=8, input_nc=5, prune_cin_lb=1, frac=0.25, seed=4),
    # scripts/gaugan/cityscapes/train_inception_student_5p6B.sh -> tests/golden/arch_gaugan_5p6B.json
    'gaugan_5p6B': dict(crop_size=512, aspect_ratio=2.0, teacher_ngf=64, student_ngf=48, ndf=64, input_nc=35, prune_cin_lb=16,
                        target_flops=5.6e9),
}


def main_spade():
    from oracle.ref_harness_spade import build_reference_spade_distiller, spade_generator_arch
    out = {}
    for name, cfg in SPADE_CASES.items():
        cfg = dict(cfg)
        frac = cfg.pop('frac', None)
        if frac is not None:
            probe, _ = build_reference_spade_distiller(batch_size=1, do_shrink=False, **cfg)
            cfg['target_flops'] = probe.modules_on_one_gpu.netG_teacher.n_macs * frac
        model, opt = build_reference_spade_distiller(batch_size=1, **cfg)
        mm = model.modules_on_one_gpu
        keep = {k: v.detach().clone() for k, v in mm.netG_teacher.state_dict().items() if k.endswith('norm.weight')}
        out[name] = {'gammas': keep, 'teacher_arch': spade_generator_arch(mm.netG_teacher), 'student_arch': spade_generator_arch(mm.netG_student),
                     'teacher_macs': int(mm.netG_teacher.n_macs), 'student_macs': int(mm.netG_student.n_macs),
                     'target_flops': float(cfg['target_flops']), 'prune_cin_lb': cfg['prune_cin_lb']}
        print(name, out[name]['teacher_macs'], out[name]['student_macs'], out[name]['student_arch']['fc_out'])
    torch.save(out, SPADE_OUT)
    print('->', SPADE_OUT, os.path.getsize(SPADE_OUT))


if __name__ == '__main__':
    if 'spade' in sys.argv[1:]:
        main_spade()
    else:
        main()
