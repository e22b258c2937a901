"""Build libmrphy_b200.so (C ABI, include/mrphy_b200.h) in-tree with nvcc for sm_100a.

    python mrphy.py_b200/build.py [--force] [--verbose]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libmrphy_b200.so')
SOURCES = ['blochsim_fused.cu', 'blochsim_beff.cu', 'aux_ops.cu', 'design_ops.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-std=c++17', '-lineinfo', '-ftz=true',
              '--threads', '4', '--shared', '-Xcompiler', '-fPIC', '-I', os.path.join(ROOT, 'include')]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, 'include', 'mrphy_b200.h'), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=(), out=None):
    """`defines`/`out` build a tuning variant (e.g. -DMRPHY_BWD_MINB=10) next to the default library."""
    if out is not None:
        return _compile(list(defines), out, verbose)
    if not force and not _stale():
        return LIB
    return _compile([], LIB, verbose)


def _compile(defines, LIB, verbose):
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    cmd = [nvcc] + NVCC_FLAGS + list(defines) + (['-Xptxas', '-v'] if verbose else []) + \
        [os.path.join(CSRC, s) for s in SOURCES] + ['-o', LIB + '.tmp']
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    os.replace(LIB + '.tmp', LIB)
    return LIB


if __name__ == '__main__':
    defs = [a for a in sys.argv[1:] if a.startswith('-D')]
    outs = [a.split('=', 1)[1] for a in sys.argv[1:] if a.startswith('--out=')]
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv, defines=defs,
                out=os.path.abspath(outs[0]) if outs else None))
