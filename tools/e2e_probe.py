"""Where the time of the host-buffer residual call goes (run on the GPU box):
python tools/e2e_probe.py [n]"""
import ctypes
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402
from lid_driven_cavity_problem_b200 import _native  # noqa: E402
from lid_driven_cavity_problem_b200.residual_function import cuda_residual_function as crf  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
g, X = bench.synthetic_case(n, n)
lib = _native.lib()
grid = _native.grid_from_graph(g)
U, V = g.ns_x_mesh.phi_old, g.ns_y_mesh.phi_old
F = np.empty_like(X)


def t(fn, reps=30):
    for _ in range(3):
        fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e3


def raw():
    lib.ldc_residual_host(grid, X.ctypes.data, U.ctypes.data, V.ctypes.data, F.ctypes.data, 0)


print("ldc_residual_host, reused F buffer : %.3f ms" % t(raw))
sys.stdout.flush()
print("residual_function(X, graph)        : %.3f ms" % t(lambda: crf.residual_function(X, g)))
print("np.empty + touch                   : %.3f ms" % t(lambda: np.empty_like(X).fill(0.0)))
os.environ['LDC_HOST_PROFILE'] = '1'
