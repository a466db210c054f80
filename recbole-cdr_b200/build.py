#!/usr/bin/env python
"""Build libxdr.so (the C-ABI CUDA library) in-tree for sm_100a.

    python recbole-cdr_b200/build.py [--force] [--verbose]

Every ``csrc/*.cu`` is compiled with ``nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo`` into
``build/*.o`` and linked into ``recbole_cdr_b200/lib/libxdr.so`` (git-ignored, but it travels to the GPU box with
the gpurun snapshot).  nvcc cross-compiles without a GPU.  The CUDA runtime is linked statically, so the library
has no dependency on torch or on a system libcudart.
"""
import argparse
import concurrent.futures as cf
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
BUILD = os.path.join(HERE, os.environ.get('XDR_BUILD_DIR', 'build'))
LIB_DIR = os.path.join(HERE, 'recbole_cdr_b200', 'lib')
LIB = os.path.join(LIB_DIR, os.environ.get('XDR_LIB_NAME', 'libxdr.so'))
INCLUDE = os.path.join(os.path.dirname(HERE), 'include')

NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17', '--expt-extended-lambda',
         '--expt-relaxed-constexpr', '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden', '-I', INCLUDE]
FLAGS += os.environ.get('XDR_EXTRA_NVCC_FLAGS', '').split()


def _newer(target, deps):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    os.makedirs(LIB_DIR, exist_ok=True)
    sources = sorted(glob.glob(os.path.join(CSRC, '*.cu')))
    headers = sorted(glob.glob(os.path.join(CSRC, '*.cuh'))) + sorted(glob.glob(os.path.join(INCLUDE, '*.h')))
    if not sources:
        raise RuntimeError('no CUDA sources found under ' + CSRC)
    jobs = []
    for src in sources:
        obj = os.path.join(BUILD, os.path.basename(src)[:-3] + '.o')
        if force or not _newer(obj, [src] + headers + [__file__]):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [NVCC] + FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, r.returncode, r.stdout + r.stderr

    failed = False
    with cf.ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for src, rc, out in ex.map(compile_one, jobs):
            if verbose or rc != 0:
                sys.stderr.write(f'--- nvcc {os.path.basename(src)} (rc={rc})\n{out}\n')
            failed |= rc != 0
    if failed:
        raise RuntimeError('nvcc failed; see messages above')
    objs = [os.path.join(BUILD, os.path.basename(s)[:-3] + '.o') for s in sources]
    if force or jobs or not _newer(LIB, objs):
        cmd = [NVCC, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', LIB] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('link failed:\n' + r.stdout + r.stderr)
    return LIB


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--force', action='store_true')
    ap.add_argument('--verbose', action='store_true')
    a = ap.parse_args()
    print(build(a.force, a.verbose))
