"""Build libldc_b200.so (sm_100a) in-tree with nvcc.

    python lid_driven_cavity_problem_b200/csrc/build.py [--force] [--verbose]

One translation unit per .cu, compiled in parallel, linked into
lid_driven_cavity_problem_b200/csrc/libldc_b200.so.  nvcc cross-compiles without a GPU.
"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ['common.cu', 'residual.cu', 'pack.cu', 'jacobian.cu', 'spmv.cu', 'vector_ops.cu',
           'precond.cu', 'solver.cu', 'host_pipeline.cu']
LIB = os.path.join(HERE, 'libldc_b200.so')
COMM_LIB = os.path.join(HERE, 'libldc_b200_comm.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
         '-Xcompiler', '-fPIC,-fvisibility=hidden', '--expt-relaxed-constexpr',
         '-DLDC_BUILDING=1']


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(HERE, s))]
    headers = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith(('.cuh', '.h'))]
    headers.append(os.path.join(HERE, '..', '..', 'include', 'ldc_b200.h'))
    objdir = os.path.join(HERE, 'build')
    os.makedirs(objdir, exist_ok=True)
    jobs = []
    for s in srcs:
        src = os.path.join(HERE, s)
        obj = os.path.join(objdir, s.replace('.cu', '.o'))
        if force or _stale(obj, [src] + headers):
            cmd = [NVCC] + FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', obj]
            jobs.append((s, cmd))
    def run(job):
        name, cmd = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        return name, r
    failed = False
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for name, r in ex.map(run, jobs):
            if verbose or r.returncode != 0:
                sys.stderr.write('--- %s\n%s%s' % (name, r.stdout, r.stderr))
            if r.returncode != 0:
                failed = True
    if failed:
        raise RuntimeError('nvcc failed')
    objs = [os.path.join(objdir, s.replace('.cu', '.o')) for s in srcs]
    if force or jobs or _stale(LIB, objs):
        cmd = [NVCC, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError('link failed')
    build_comm(force=force)
    return LIB


def _nccl_paths():
    import importlib.util
    spec = importlib.util.find_spec('nvidia.nccl')
    if spec is not None and spec.submodule_search_locations:
        root = list(spec.submodule_search_locations)[0]
        inc, lib = os.path.join(root, 'include'), os.path.join(root, 'lib')
        if os.path.exists(os.path.join(inc, 'nccl.h')) and os.path.exists(os.path.join(lib, 'libnccl.so.2')):
            return inc, lib
    if os.path.exists('/usr/include/nccl.h'):
        return '/usr/include', '/usr/lib/x86_64-linux-gnu'
    return None, None


def build_comm(force=False):
    """libldc_b200_comm.so: the NCCL transport (links the same libnccl.so.2 torch loads)."""
    src = os.path.join(HERE, 'comm.cu')
    if not os.path.exists(src):
        return None
    inc, lib = _nccl_paths()
    if inc is None:
        sys.stderr.write('nccl.h not found: libldc_b200_comm.so not built\n')
        return None
    deps = [src, os.path.join(HERE, '..', '..', 'include', 'ldc_b200_comm.h'),
            os.path.join(HERE, '..', '..', 'include', 'ldc_b200.h')]
    if force or _stale(COMM_LIB, deps):
        cmd = [NVCC] + FLAGS + ['-shared', '-I', inc, src, '-o', COMM_LIB, '-L', lib, '-l:libnccl.so.2',
                                '-Xlinker', '-rpath,' + lib]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError('comm library build failed')
    return COMM_LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
