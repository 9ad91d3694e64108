"""Distributed strip solve vs single-GPU solve on ONE GPU through the in-process transport of
tests/thread_transport.py (debugging aid for bench.py's strip_solve_vs_single_gpu)."""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'tests'))
import torch  # noqa: E402
from thread_transport import run_strips  # noqa: E402
from lid_driven_cavity_problem_b200.strip_solver import CudaStripSolver  # noqa: E402

n, bc, steps = int(sys.argv[1]) if len(sys.argv) > 1 else 256, 1000.0, 3
opts = dict(snes_rtol=1e-13, snes_atol=1e-9, snes_stol=1e-15, ksp_rtol=1e-8, snes_max_it=30, gmres_refine=1)
for fp32 in (1, 0):
    o = dict(opts, pc_fp32=fp32)
    one = CudaStripSolver(n, n, 1e-2, 1.0, 1.0, bc, **o)
    one.set_uniform_state()
    t1 = one.run(max_steps=steps)
    X1 = one.get_state().cpu().numpy()
    print('fp32=%d single: t=%.6f steps=%d retries=%d newton=%d krylov=%d' % (fp32, t1, one.steps, one.retries, one.newton_iterations, one.krylov_iterations))
    one.close()
    for world in (2, 4, 8):
        def strip(rank, comm):
            s = CudaStripSolver(n, n, 1e-2, 1.0, 1.0, bc, comm=comm, rank=rank, world=world, use_cuda_graphs=0, **o)
            try:
                s.set_uniform_state()
                t = s.run(max_steps=steps)
                torch.cuda.synchronize()
                return s.get_state().cpu().numpy(), t, s.steps, s.retries, s.newton_iterations, s.krylov_iterations
            finally:
                s.close()
        parts, _ = run_strips(world, strip)
        X = np.concatenate([p[0] for p in parts])
        print('fp32=%d world=%d: t=%.6f steps=%d retries=%d newton=%d krylov=%d  max|dU|rel=%.2e' % (
            fp32, world, parts[0][1], parts[0][2], parts[0][3], parts[0][4], parts[0][5],
            np.abs(X[1::3] - X1[1::3]).max() / np.abs(X1[1::3]).max()))
