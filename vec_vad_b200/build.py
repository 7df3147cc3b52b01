"""In-tree build of libvecvad.so (sm_100a only) with plain nvcc -- no JIT cache, the .so travels with the tree."""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(PKG)
CSRC = os.path.join(PKG, 'csrc')
LIB = os.path.join(PKG, 'libvecvad.so')
SOURCES = ['net.cu', 'single_ops.cu', 'unet_kernels.cu', 'igemm_simt.cu', 'tc_support.cu', 'igemm_flat.cu', 'igemm_tc3.cu', 'wgrad_tc2.cu', 'wgrad_flat.cu', 'flow_ops.cu', 'corr_tma.cu', 'prof.cu', 'crop_resize.cu', 'flownet_ops.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-Xcompiler', '-fPIC',
              '-I' + os.path.join(REPO, 'include'), '-I' + CSRC]


def _newer(a, b):
    return (not os.path.exists(b)) or os.path.getmtime(a) > os.path.getmtime(b)


def build(force=False, verbose=False):
    """Compile every .cu to an object (only when stale) and link the shared library."""
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    objdir = os.path.join(PKG, 'build')
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.h', '.cuh'))]
    headers.append(os.path.join(REPO, 'include', 'vecvad.h'))
    objs, procs = [], []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(objdir, s.replace('.cu', '.o'))
        objs.append(obj)
        if force or _newer(src, obj) or any(_newer(h, obj) for h in headers):
            cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', obj]
            procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write('[nvcc %s]\n%s\n' % (s, out))
        failed = failed or p.returncode != 0
    if failed:
        raise RuntimeError('nvcc failed')
    if procs or not os.path.exists(LIB):
        cmd = [nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-o', LIB] + objs + ['-lcudart']
        subprocess.check_call(cmd)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
