"""In-tree build of libhpb200.so (hand-written sm_100a CUDA + the C-ABI of include/hpb200.h).

nvcc cross-compiles without a GPU.  The built library stays next to this file so that it ships
to the GPU box with the repo snapshot.  Usage:  python -m hipace_b200.build [--force]
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
ROOT = os.path.dirname(HERE)
BUILD = os.path.join(ROOT, 'build')
LIB = os.path.join(HERE, 'libhpb200.so')

ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
COMMON = ['-O3', '-lineinfo', '-std=c++17', '-Xcompiler', '-fPIC', '-Xptxas', '-v']
# per-file extra flags: mg.cu keeps separate multiply/add so that every level operator rounds
# exactly like the oracle (the V-cycle count depends on a 1e-4 threshold)
SOURCES = {
    'context.cu': [],
    'particles.cu': [],
    'fields.cu': [],
    'poisson.cu': [],
    'mg.cu': ['-fmad=false'],
    'sim.cu': [],
    'beam.cu': [],
    'pipeline.cu': [],
    'laser.cu': [],
    'generic_order.cu': [],
    'insitu.cu': [],
    'pc_fields.cu': [],
    'reorder.cu': [],
    'peaks.cu': [],
    'ref_gpu_arm.cu': [],
    'fft2d.cu': [],
    'periodic.cu': [],
}


def _nvcc():
    for c in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError('nvcc not found')


def _stamp():
    h = hashlib.sha256()
    for root, _, files in sorted(os.walk(CSRC)):
        for f in sorted(files):
            h.update(open(os.path.join(root, f), 'rb').read())
    h.update(open(os.path.join(ROOT, 'include', 'hpb200.h'), 'rb').read())
    h.update(repr(sorted(SOURCES.items())).encode())
    return h.hexdigest()


def _compile(src, flags):
    obj = os.path.join(BUILD, src.replace('.cu', '.o'))
    cmd = [_nvcc()] + ARCH + COMMON + flags + ['-c', os.path.join(CSRC, src), '-o', obj]
    p = subprocess.run(cmd, capture_output=True, text=True)
    open(obj.replace('.o', '.log'), 'w').write(' '.join(cmd) + '\n' + p.stdout + p.stderr)
    if p.returncode != 0:
        raise RuntimeError(f'nvcc failed for {src}:\n{p.stdout}\n{p.stderr}')
    return obj


def build_library(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    stamp_file = os.path.join(BUILD, 'libhpb200.stamp')
    stamp = _stamp()
    if (not force and os.path.exists(LIB) and os.path.exists(stamp_file)
            and open(stamp_file).read() == stamp):
        return LIB
    srcs = {s: f for s, f in SOURCES.items() if os.path.exists(os.path.join(CSRC, s))}
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda kv: _compile(*kv), srcs.items()))
    cmd = [_nvcc()] + ARCH + ['-shared', '--cudart', 'shared', '-o', LIB] + objs + ['-ldl']
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError('link failed:\n' + p.stdout + p.stderr)
    # the command-line driver (C++ host over the C-ABI, the reference's `hipace inputs ...`)
    os.makedirs(os.path.join(HERE, 'bin'), exist_ok=True)
    exe = os.path.join(HERE, 'bin', 'hpb200_run')
    cmd = ['g++', '-O2', '-std=c++17', '-I', os.path.join(ROOT, 'include'), os.path.join(CSRC, 'main.cpp'),
           '-L', HERE, '-lhpb200', '-Wl,-rpath,$ORIGIN/..', '-Wl,-rpath,/usr/local/cuda/lib64',
           '-L/usr/local/cuda/lib64', '-lcudart', '-o', exe]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError('driver link failed:\n' + p.stdout + p.stderr)
    open(stamp_file, 'w').write(stamp)
    if verbose:
        print('built', LIB, 'and', exe)
    return LIB


def build_host_check(force: bool = False) -> str:
    """TEST artefact: tests/host_check.cu -> build/libhpb200_hostcheck.so, the __host__ __device__
    arithmetic of the kernels compiled for the CPU (tests/test_device_math_host.py)."""
    os.makedirs(BUILD, exist_ok=True)
    lib = os.path.join(BUILD, 'libhpb200_hostcheck.so')
    deps = [os.path.join(ROOT, 'tests', 'host_check.cu')]
    deps += [os.path.join(CSRC, f) for f in ('generic_order.cuh', 'push_math.cuh', 'shapes.cuh', 'common.cuh',
                                             'insitu.cuh', 'pc_fields.cuh', 'laser_advance.cuh')]
    deps.append(os.path.join(ROOT, 'include', 'hpb200.h'))
    if (not force and os.path.exists(lib)
            and all(os.path.getmtime(lib) >= os.path.getmtime(d) for d in deps)):
        return lib
    # -ffp-contract=off: the host run is compared with the (non-contracting) reference headers
    cmd = [_nvcc()] + ARCH + ['-O2', '-std=c++17', '-Xcompiler', '-fPIC,-ffp-contract=off', '-shared',
                              '--cudart', 'shared', '-I', CSRC, '-o', lib, deps[0]]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError('host_check build failed:\n' + p.stdout + p.stderr)
    return lib


if __name__ == '__main__':
    build_library(force='--force' in sys.argv, verbose=True)
