"""In-tree build of libgvk_sm100a.so (nvcc, sm_100a only).  `python -m gaviko_b200.build [--force]`."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libgvk_sm100a.so')
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17']


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, '*.cuh')) + [os.path.join(os.path.dirname(HERE), 'include', 'gvk.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    flags = [f for f in NVCC_FLAGS if 'placeholder' not in f]
    objs = []
    os.makedirs(os.path.join(HERE, 'build'), exist_ok=True)
    procs = []
    hdrs = glob.glob(os.path.join(CSRC, '*.cuh')) + [os.path.join(os.path.dirname(HERE), 'include', 'gvk.h')]
    for src in sources():
        obj = os.path.join(HERE, 'build', os.path.basename(src)[:-3] + '.o')
        objs.append(obj)
        if not force and os.path.exists(obj) and all(os.path.getmtime(d) < os.path.getmtime(obj) for d in [src] + hdrs):
            continue   # object is newer than its source and every header
        cmd = [nvcc] + flags + ['-Xcompiler', '-fPIC', '-c', src, '-o', obj] + (['-Xptxas', '-v'] if verbose else [])
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f'nvcc failed for {src}:\n{out}\n')
        elif verbose:
            print(out)
    if failed:
        raise RuntimeError('libgvk_sm100a.so build failed')
    subprocess.check_call([nvcc, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a'])
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
