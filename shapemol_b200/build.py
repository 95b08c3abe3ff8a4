"""Builds libshapemol_b200.so in-tree with nvcc for sm_100a (no torch / pybind dependency)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libshapemol_b200.so')
SOURCES = ['smb_host.cu', 'smb_api.cu', 'smb_small_kernels.cu', 'smb_node_mlp.cu', 'smb_node_tc5.cu', 'smb_edge_attn.cu', 'smb_edge_ws.cu', 'smb_generic.cu', 'smb_encoder.cu', 'smb_tc_gemm.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden', '--expt-relaxed-constexpr']


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.h', '.cuh'))]
    headers.append(os.path.join(HERE, '..', 'include', 'shapemol_b200.h'))
    objdir = os.path.join(HERE, 'build')
    os.makedirs(objdir, exist_ok=True)
    objs, procs = [], []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace('.cu', '.o'))
        objs.append(o)
        if force or _newer(o, [s] + headers):
            cmd = [nvcc] + NVCC_FLAGS + os.environ.get('SMB_NVCC_EXTRA', '').split() + (['-Xptxas', '-v'] if verbose else []) + ['-c', s, '-o', o]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write('nvcc failed for %s:\n%s\n' % (src, out))
        elif verbose:
            sys.stderr.write(out)
    if failed:
        raise RuntimeError('nvcc compilation failed')
    if force or procs or _newer(LIB, objs):
        cmd = [nvcc, '-shared', '-o', LIB] + objs + ['-lcudart_static', '-Xcompiler', '-fPIC']
        subprocess.check_call(cmd)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
