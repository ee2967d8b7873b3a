"""Builds libdmlb.so in-tree with nvcc for sm_100a (no torch headers: ~10 s, cross-compiles without a GPU).

    python -m dmlcloud_b200.csrc.build [--force] [--ptxas-v]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import hashlib
import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
SOURCES = ['core.cu', 'bucket_kernels.cu', 'bucket_tma.cu', 'peer_comm.cu', 'metric_kernels.cu', 'shard_kernels.cu',
           'optim_kernels.cu', 'vmm.cu']
HEADERS = ['dmlb_common.cuh', 'peer_comm.cuh', 'metric_dev.cuh', '../../include/dmlb.h']
LIB = HERE / 'libdmlb.so'
STAMP = HERE / '.libdmlb.stamp'

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-O3', '-lineinfo', '-std=c++17',
    '--shared', '-Xcompiler', '-fPIC',
    '-cudart', 'static',
]


def nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (Path(cand).exists() or cand == 'nvcc'):
            return cand
    raise RuntimeError('nvcc not found')


def _digest():
    h = hashlib.sha256()
    for name in SOURCES + HEADERS:
        h.update((HERE / name).read_bytes())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False, ptxas_v=False):
    digest = _digest()
    if not force and LIB.exists() and STAMP.exists() and STAMP.read_text() == digest:
        return LIB
    cmd = [nvcc()] + NVCC_FLAGS + (['-Xptxas', '-v'] if ptxas_v else []) + \
          ['-o', str(LIB)] + [str(HERE / s) for s in SOURCES]
    if verbose:
        print(' '.join(cmd), flush=True)
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + proc.stdout)
    if verbose or ptxas_v:
        print(proc.stdout)
    STAMP.write_text(digest)
    return LIB


if __name__ == '__main__':
    build(force='--force' in sys.argv, verbose=True, ptxas_v='--ptxas-v' in sys.argv)
    print(LIB)
