"""Put the UNMODIFIED reference (OpenOpt 0.43) under baseline/_ref/ so that it travels to the GPU box.

    python baseline/install_ref.py          # build container only: /root/reference is mounted here

`python -m pip install --no-index --no-build-isolation --no-deps --target baseline/_ref /root/reference`
fails in this image: the reference's setup.py imports numpy.distutils (setup.py:34), which numpy 2 /
Python 3.12 no longer ship.  The package is pure Python (setup.py:90-108 passes no ext_modules), so an
install is exactly its `openopt/` package directory; this script copies that directory, byte for byte, and
records a sha256 manifest.  baseline/_ref/ is git-ignored (never in history) but not gpurun-ignored.
Only bench.py's cpu_baseline / `--impl reference` legs import it (through oracle/ref_shim.py)."""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = '/root/reference/openopt'
DST = os.path.join(HERE, '_ref')


def install(force=False):
    if not os.path.isdir(SRC):
        return os.path.isdir(os.path.join(DST, 'openopt'))
    if os.path.isdir(os.path.join(DST, 'openopt')) and not force:
        return True
    shutil.rmtree(DST, ignore_errors=True)
    shutil.copytree(SRC, os.path.join(DST, 'openopt'), ignore=shutil.ignore_patterns('*.pyc', '__pycache__'))
    lines = []
    for root, _, files in sorted(os.walk(os.path.join(DST, 'openopt'))):
        for f in sorted(files):
            p = os.path.join(root, f)
            lines.append('%s  %s' % (hashlib.sha256(open(p, 'rb').read()).hexdigest(), os.path.relpath(p, DST)))
    open(os.path.join(DST, 'MANIFEST.sha256'), 'w').write('\n'.join(lines) + '\n')
    return True


if __name__ == '__main__':
    ok = install(force='--force' in sys.argv)
    print('baseline/_ref:', 'installed' if ok else 'unavailable (no /root/reference here and no earlier copy)')
