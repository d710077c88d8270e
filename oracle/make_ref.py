"""TEST / BENCH INFRASTRUCTURE -- recipe that stages the reference implementation for the CPU arm of the benchmark.

The reference (snap-research/CAT) is pure Python: there is nothing to compile.  This script copies the Python modules of
the distillation path from the read-only checkout at /root/reference into ``oracle/_ref/`` (git-ignored, not
gpurun-ignored: it travels to the GPU box like a built ``.so``; reference sources never enter the repository history).
``bench.py --impl reference`` then times the reference's OWN ``InceptionDistiller.optimize_parameters`` on the host cores
(``cpu_baseline.kind = "reference"``); without ``oracle/_ref`` it falls back to the CPU oracle port (``"port"``).

    python -m oracle.make_ref            (build container only; __graft_entry__.build() runs it when /root/reference exists)
"""
import os
import shutil
import sys

SRC = '/root/reference'
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')
PACKAGES = ['distillers', 'models', 'utils', 'options', 'data', 'metric']
FILES = ['common.py', 'trainer.py', 'distill.py', 'train.py', 'profile.py']


def main():
    if not os.path.isdir(SRC):
        print('no reference checkout at %s: nothing staged' % SRC)
        return False
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    n = 0
    for pkg in PACKAGES:
        for root, _dirs, files in os.walk(os.path.join(SRC, pkg)):
            rel = os.path.relpath(root, SRC)
            for f in files:
                if f.endswith('.py'):
                    os.makedirs(os.path.join(DST, rel), exist_ok=True)
                    shutil.copyfile(os.path.join(root, f), os.path.join(DST, rel, f))
                    n += 1
    for f in FILES:
        if os.path.exists(os.path.join(SRC, f)):
            shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
            n += 1
    print('staged %d reference modules under %s' % (n, DST))
    return True


if __name__ == '__main__':
    sys.exit(0 if main() else 1)
