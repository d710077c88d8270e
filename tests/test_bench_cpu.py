"""bench.py plumbing that needs no GPU: workload tables, algorithmic MAC models and the reference arm (the CPU oracle
timed on the host cores) for the headline workload and the teacher-training workloads."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_teacher_mac_models():
    import bench
    from cat_b200 import workload as WL
    for wl in bench.TEACHER:
        arch = WL.load_arch(bench.arch_name(wl))
        H, W = (256, 512) if wl.startswith('gaugan') else (256, 256)
        m = bench.teacher_macs(wl, arch, H, W)
        if wl == 'pix2pix_teacher':
            assert m['step'] == 3 * m['T'] + 8 * m['D'] and abs(m['T'] - arch['teacher_macs']) < 1
        elif wl == 'cyclegan_teacher':
            assert m['step'] == 18 * m['T'] + 16 * m['D']
        else:
            assert m['step'] == 4 * m['T'] + 10 * m['D'] + 3 * m['V']
        hp = bench.teacher_hp(wl, arch)
        assert ('lambda_A' in hp) == (wl == 'cyclegan_teacher')
        assert bench.metric_name(wl) == 'teacher-train-step images/sec'
    assert bench.metric_name('pix2pix_5p6B') == 'distill-step images/sec'


@pytest.mark.timeout(600)
@pytest.mark.parametrize('workload', ['pix2pix_5p6B', pytest.param('pix2pix_teacher', marks=pytest.mark.slow)])
def test_reference_arm_prints_the_contract_line(workload):
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--workload', workload, '--steps', '1',
                          '--warmup', '0', '--cpu-batch', '2', '--height', '32', '--width', '32'], capture_output=True, text=True,
                         timeout=500, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line['impl'] == 'reference' and line['unit'] == 'images/s' and line['value'] > 0
    # the reference itself when oracle/make_ref.py has staged it (build container: __graft_entry__.build()), else the port
    import bench
    want = 'reference' if bench.real_reference_available(workload) else 'port'
    assert line['cpu_baseline']['kind'] == want and line['cpu_baseline']['cores'] == (os.cpu_count() or 1)
    assert len(out.stdout.strip().splitlines()) == 1, 'the arm prints ONE line on stdout'
    assert line['e2e'] == {'value': line['value'], 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert line['config']['workload'].startswith(workload)
