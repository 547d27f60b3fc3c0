"""Build liblbt_b200.so in-tree with nvcc for sm_100a (B200) — no JIT, no torch C++ extension.

    python -m lbt_b200.build [--force] [--verbose]

The .so carries sm_100a SASS only (tcgen05 kind::i8 exists on no other target) and is loaded by
lbt_b200._lib through ctypes.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OUT = os.environ.get('LBT_BUILD_OUT') or os.path.join(HERE, 'liblbt_b200.so')      # experiments: a second library with other -D flags
OBJ = os.path.join(HERE, 'build' + os.environ.get('LBT_BUILD_TAG', ''))
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
         '-Xcompiler', '-fPIC,-O2', '--expt-relaxed-constexpr', '-Xptxas', '-v',
         '-I' + os.path.join(os.path.dirname(HERE), 'include')] + os.environ.get('LBT_NVCC_DEFS', '').split()


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def _digest():
    h = hashlib.sha256()
    for root, _, files in sorted(os.walk(CSRC)):
        for f in sorted(files):
            with open(os.path.join(root, f), 'rb') as fh:
                h.update(f.encode())
                h.update(fh.read())
    with open(os.path.join(os.path.dirname(HERE), 'include', 'lbt.h'), 'rb') as fh:
        h.update(fh.read())
    h.update(' '.join(FLAGS).encode())
    return h.hexdigest()


def _compile(src, verbose):
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + '.o')
    cmd = [NVCC] + FLAGS + ['-c', src, '-o', obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = r.stdout + r.stderr
    with open(obj + '.log', 'w') as f:
        f.write(' '.join(cmd) + '\n' + log)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed for %s:\n%s' % (src, log))
    if verbose:
        print(log)
    return obj


def build(force=False, verbose=False):
    """Compile every csrc/*.cu for sm_100a and link liblbt_b200.so.  Returns the path."""
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(OBJ, 'digest.txt')
    dig = _digest()
    if not force and os.path.exists(OUT) and os.path.exists(stamp) and open(stamp).read() == dig:
        return OUT
    if not os.path.exists(NVCC):
        raise RuntimeError('nvcc not found at %s and no prebuilt %s' % (NVCC, OUT))
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), sources()))
    cmd = [NVCC, '-shared', '-o', OUT] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-lcudart_static',
                                                   '-lpthread', '-ldl', '-lrt']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n' + r.stdout + r.stderr)
    with open(stamp, 'w') as f:
        f.write(dig)
    return OUT


if __name__ == '__main__':
    p = build(force='--force' in sys.argv, verbose='--verbose' in sys.argv)
    print(p)
