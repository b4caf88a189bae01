"""Builds libbn_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

One translation unit per instantiation group so the build parallelises across host cores.
Objects go to ``build/`` (git-ignored); the shared library lands next to this file so it travels
with the repo snapshot to the GPU box.
"""
import concurrent.futures as cf
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
ROOT = os.path.dirname(HERE)
OBJ_DIR = os.path.join(ROOT, 'build', 'obj' + os.environ.get('BN_B200_OBJ_SUFFIX', ''))
LIB = os.path.join(HERE, os.environ.get('BN_B200_LIBNAME', 'libbn_b200.so'))
EXTRA = os.environ.get('BN_B200_NVCC_EXTRA', '').split()  # e.g. -DBN_UP_TJ=4 -DBN_UP_BLOCKS=6 for tuning variants

NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '--use_fast_math=false', '-Xcompiler', '-fPIC', '-Xcompiler', '-O2']
NVCC_FLAGS = [f for f in NVCC_FLAGS if f != '--use_fast_math=false']  # IEEE fp64 everywhere: never fast-math


def _newest_header():
    hs = glob.glob(os.path.join(CSRC, '*.cuh')) + glob.glob(os.path.join(ROOT, 'include', '*.h'))
    return max(os.path.getmtime(h) for h in hs)


def _compile(src, obj, verbose):
    cmd = [NVCC] + NVCC_FLAGS + EXTRA + ['-c', src, '-o', obj]
    if verbose:
        cmd.insert(1, '-Xptxas=-v')
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (src, r.stdout, r.stderr))
    return r.stderr


def build(force=False, verbose=False, jobs=None):
    os.makedirs(OBJ_DIR, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, '*.cu')))
    hdr_t = _newest_header()
    todo, objs = [], []
    for s in srcs:
        o = os.path.join(OBJ_DIR, os.path.basename(s)[:-3] + '.o')
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hdr_t):
            todo.append((s, o))
    logs = []
    if todo:
        with cf.ThreadPoolExecutor(max_workers=jobs or os.cpu_count() or 4) as ex:
            futs = [ex.submit(_compile, s, o, verbose) for s, o in todo]
            for f in futs:
                logs.append(f.result())
    if todo or not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(o) for o in objs):
        cmd = [NVCC, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-lcudart']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('link failed:\n%s\n%s' % (r.stdout, r.stderr))
    return LIB, logs


if __name__ == '__main__':
    lib, logs = build(force='--force' in sys.argv, verbose='-v' in sys.argv)
    for l in logs:
        if l.strip():
            print(l)
    print(lib)
