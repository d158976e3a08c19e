"""Builds libanatomask_b200.so in-tree with nvcc for sm_100a (no torch headers, no JIT cache)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libanatomask_b200.so')
SOURCES = ['misc.cu', 'norm.cu', 'conv_direct.cu', 'conv_igemm.cu', 'conv_igemm3.cu', 'conv_igemm4.cu', 'conv_igemm4t.cu', 'conv_wgrad_tc.cu',
           'conv_wgrad_halo.cu', 'conv_wgrad_ns.cu', 'sparse_layers.cu', 'augment.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17', '--use_fast_math',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-O2', '-Wno-deprecated-gpu-targets']


STAMP = LIB + '.src-sha256'


def source_hash() -> str:
    """sha256 over every file the library is compiled from (csrc/ + the C header) and the compiler flags."""
    import hashlib
    h = hashlib.sha256(' '.join(NVCC_FLAGS + SOURCES).encode())
    deps = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)) + [os.path.join(HERE, '..', 'include', 'anatomask_b200.h')]
    for d in deps:
        h.update(os.path.basename(d).encode())
        with open(d, 'rb') as f:
            h.update(f.read())
    return h.hexdigest()


def built_hash() -> str:
    try:
        with open(STAMP) as f:
            return f.read().strip()
    except OSError:
        return ''


def needs_build() -> bool:
    """By content, not by mtime: a `git checkout` of a source file after the last build must trigger a rebuild, and the copy
    of the tree on a GPU box (fresh mtimes) must not."""
    return not os.path.exists(LIB) or built_hash() != source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, 'build'), exist_ok=True)
    for s in SOURCES:
        o = os.path.join(HERE, 'build', s.replace('.cu', '.o'))
        cmd = [nvcc, *NVCC_FLAGS, '-c', os.path.join(CSRC, s), '-o', o] + (['-Xptxas', '-v'] if verbose else [])
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    fail = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            print(f'--- {s}\n{out}', file=sys.stderr)
        fail |= p.returncode != 0
    if fail:
        raise RuntimeError('nvcc failed')
    cmd = [nvcc, '-shared', '-o', LIB, *objs, '-cudart', 'static', '-Wno-deprecated-gpu-targets']
    subprocess.check_call(cmd)
    with open(STAMP, 'w') as f:
        f.write(source_hash() + '\n')
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
