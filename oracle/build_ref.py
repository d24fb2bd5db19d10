"""Stage the UNMODIFIED reference for the GPU box:  python -m oracle.build_ref   (test / bench infrastructure only).

The reference is pure Python -- there is nothing to compile -- so "building" it means copying the modules of the
forecasting path byte for byte from /root/reference into oracle/_ref/reference/ (git-ignored: no reference source
ever enters the history; NOT gpurun-ignored: the staged copy travels to the GPU box like a built .so).  There
``bench.py --impl reference`` runs the reference's own ``utils/evaluate.py::evaluate`` on the host cores
(``cpu_baseline.kind = "reference"``) and the GPU tests can check the oracle against it once more.

Only ``models/`` and ``utils/`` (+ the two ``__init__.py``) are staged; loralib (un-vendored, absent) is provided by
oracle/loralib_restatement.py exactly as in the build container (oracle/ref_harness.py).
"""
import filecmp
import os
import shutil
import sys

SRC = os.environ.get('REF_SRC', '/root/reference')
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref', 'reference')
PACKAGES = ('models', 'utils')


def build(verbose=True):
    if not os.path.isdir(os.path.join(SRC, 'models')):
        if verbose:
            print(f'oracle/_ref: {SRC} not present, keeping the staged copy' if os.path.isdir(DST)
                  else f'oracle/_ref: {SRC} not present and nothing staged')
        return os.path.isdir(DST)
    n = 0
    for pkg in PACKAGES:
        os.makedirs(os.path.join(DST, pkg), exist_ok=True)
        for name in sorted(os.listdir(os.path.join(SRC, pkg))):
            if not name.endswith('.py'):
                continue
            s, d = os.path.join(SRC, pkg, name), os.path.join(DST, pkg, name)
            if not (os.path.exists(d) and filecmp.cmp(s, d, shallow=False)):
                shutil.copyfile(s, d)
            n += 1
    init = os.path.join(SRC, '__init__.py')
    if os.path.exists(init):
        shutil.copyfile(init, os.path.join(DST, '__init__.py'))
    if verbose:
        print(f'oracle/_ref: staged {n} unmodified reference modules under {DST}')
    return True


if __name__ == '__main__':
    sys.exit(0 if build() else 1)
