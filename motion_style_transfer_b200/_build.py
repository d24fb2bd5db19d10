"""In-tree build of libynet_b200.so (nvcc, sm_100a only).

    python -m motion_style_transfer_b200._build [--force]

The shared object is written next to this file so that it travels with the source tree to the GPU
box; objects go to ``build/`` (git-ignored).  nvcc cross-compiles without a GPU.
"""
import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
INCLUDE = os.path.join(ROOT, 'include')
OBJ_DIR = os.path.join(ROOT, 'build', 'ynet_b200')
LIB_PATH = os.path.join(HERE, 'libynet_b200.so')

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
    '-Xcompiler', '-fPIC',
    '--expt-relaxed-constexpr', '-I', INCLUDE, '-I', CSRC,
]


def find_nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    return None


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def _deps_mtime():
    heads = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    heads += [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)]
    return max(os.path.getmtime(h) for h in heads)


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > t for s in sources()) or _deps_mtime() > t


def _compile(nvcc, src, obj):
    cmd = [nvcc] + NVCC_FLAGS + ['-c', src, '-o', obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'nvcc failed for {src}:\n{r.stdout}\n{r.stderr}')
    return r.stderr


def build(force=False, verbose=True):
    if not force and not needs_build():
        return LIB_PATH
    nvcc = find_nvcc()
    if nvcc is None:
        raise RuntimeError('nvcc not found: cannot build libynet_b200.so (and there is no CPU fallback)')
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdr_t = _deps_mtime()
    jobs = []
    objs = []
    for s in sources():
        o = os.path.join(OBJ_DIR, os.path.basename(s)[:-3] + '.o')
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hdr_t):
            jobs.append((s, o))
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        futs = {ex.submit(_compile, nvcc, s, o): s for s, o in jobs}
        for f in concurrent.futures.as_completed(futs):
            msg = f.result()
            if verbose:
                print(f'[ynet_b200 build] compiled {os.path.basename(futs[f])}', file=sys.stderr)
                if msg.strip():
                    print(msg, file=sys.stderr)
    tmp = LIB_PATH + '.tmp'
    link = [nvcc, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-Xcompiler', '-fPIC',
            '-o', tmp] + objs + ['-lcudart_static', '-ldl', '-lpthread', '-lrt']
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'link failed:\n{r.stdout}\n{r.stderr}')
    os.replace(tmp, LIB_PATH)
    if verbose:
        print(f'[ynet_b200 build] linked {LIB_PATH}', file=sys.stderr)
    return LIB_PATH


if __name__ == '__main__':
    build(force='--force' in sys.argv)
