"""Stage the reference's own driver scripts (ddgan.py, test_ddgan.py, ema.py) under git-ignored baseline/_ref/ so that the GPU box
can run them UNCHANGED on top of this repository's drop-in `score_sde` package (tests/test_runs_unchanged_gpu.py).  Runs in the
build container only (reads /root/reference); nothing is copied into tracked paths, and the reference's own `score_sde` package is
deliberately NOT staged: the point of the test is that the scripts import this repository's one."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'
DST = os.path.join(ROOT, 'baseline', '_ref')


def stage():
    if not os.path.isdir(REF):
        return False
    os.makedirs(DST, exist_ok=True)
    for f in ('ddgan.py', 'test_ddgan.py', 'ema.py'):
        shutil.copyfile(os.path.join(REF, f), os.path.join(DST, f))
    return True


if __name__ == '__main__':
    print('staged' if stage() else 'reference not present', DST)
    sys.exit(0)
