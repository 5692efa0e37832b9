"""In-tree build of libcdra (nvcc, sm_100a) and of the CPU logic-check build used by the CPU tests.

    python carla-driving-rl-agent_b200/build.py            # both
    python carla-driving-rl-agent_b200/build.py cuda|emu
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
CUDA_LIB = os.path.join(HERE, 'cdra', 'libcdra.so')
EMU_LIB = os.path.join(ROOT, 'tests', 'emu', 'libcdra_emu.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')

SOURCES = ['cdra_lib.cu', 'plan.cpp']


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _deps():
    out = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    out.append(os.path.join(ROOT, 'include', 'cdra.h'))
    out.append(os.path.join(ROOT, 'tests', 'emu', 'cuda_emu.h'))
    return out


def build_cuda(force=False, verbose=False):
    if not force and not _newer(CUDA_LIB, _deps()):
        return CUDA_LIB
    cmd = [NVCC, '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
           '-Xcompiler', '-fPIC', '-shared', '--use_fast_math' if False else '-DCDRA_CUDA=1',
           '-o', CUDA_LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, '-Xptxas=-v')
    subprocess.check_call(cmd, cwd=CSRC)
    return CUDA_LIB


def build_emu(force=False):
    if not force and not _newer(EMU_LIB, _deps()):
        return EMU_LIB
    cmd = ['g++', '-O2', '-std=c++17', '-fPIC', '-shared', '-DCDRA_EMU=1', '-ffp-contract=off', '-w',
           '-I/usr/local/cuda/include', '-I' + os.path.join(ROOT, 'tests', 'emu'),
           '-x', 'c++', os.path.join(CSRC, 'cdra_lib.cu'), os.path.join(CSRC, 'plan.cpp'), '-o', EMU_LIB]
    subprocess.check_call(cmd, cwd=CSRC)
    return EMU_LIB


if __name__ == '__main__':
    what = sys.argv[1] if len(sys.argv) > 1 else 'all'
    if what in ('cuda', 'all'):
        print(build_cuda(force=True, verbose='-v' in sys.argv))
    if what in ('emu', 'all'):
        print(build_emu(force=True))
