"""Build the CUDA library in-tree:  python -m quinn_b200.build

One translation unit, compiled for sm_100a only (-gencode arch=compute_100a,code=sm_100a).
The resulting quinn_b200/lib/libquinn_b200.so is git-ignored but travels to the GPU box.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, 'csrc', 'qb_kernels.cu')
DEPS = [SRC, os.path.join(HERE, 'csrc', 'qb_device.cuh'), os.path.join(HERE, 'csrc', 'qb_tc.cuh'), os.path.join(HERE, 'csrc', 'qb_plan.h'),
        os.path.join(ROOT, 'include', 'quinn_b200.h')]
OUT = os.path.join(HERE, 'lib', 'libquinn_b200.so')


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    cmd = [nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
           '--shared', '-Xcompiler', '-fPIC', '-I', os.path.join(ROOT, 'include'), '-I', os.path.join(HERE, 'csrc'),
           '-o', OUT, SRC]
    if verbose:
        cmd.insert(1, '-Xptxas')
        cmd.insert(2, '-v')
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError('nvcc failed building libquinn_b200.so')
    if verbose:
        sys.stderr.write(res.stderr)
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
