"""Wall-time sweep of full steady runs -- the table the reference's experiments/benchmarks.py
(:11,35-57,59-75) only plots: mean of N_RUNS `run_simulation(final_time=None)` calls at Re=400,
dt=1e-2 on grids 8..128, here for the `cuda` solver (and, with --cpu, for the PETSc-equivalent CPU
restatement of the reference path up to --cpu-max).

    python tools/benchmarks.py [--runs 5] [--grids 8,16,32,64,128,256] [--cpu] [--cpu-max 64]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--runs', type=int, default=5)
    ap.add_argument('--grids', default='8,16,32,64,128,256')
    ap.add_argument('--re', type=float, default=400.0)
    ap.add_argument('--cpu', action='store_true')
    ap.add_argument('--cpu-max', type=int, default=64)
    args = ap.parse_args()
    import torch
    from lid_driven_cavity_problem_b200.nonlinear_solver import solver_builder
    from lid_driven_cavity_problem_b200.staggered_grid import Graph
    from lid_driven_cavity_problem_b200.time_stepper import run_simulation
    rows = []
    for n in [int(v) for v in args.grids.split(',')]:
        solver = solver_builder.build('cuda', None)
        times = []
        for _ in range(args.runs):
            g = Graph(1.0, 1.0, n, n, 1e-2, 1.0, 1.0, args.re)
            torch.cuda.synchronize()
            t0 = time.time()
            out = run_simulation(g, None, solver)
            _ = out.ns_x_mesh.phi     # results on the host, like the reference's runs
            times.append(time.time() - t0)
        row = dict(grid='%dx%d' % (n, n), Re=args.re, cuda_mean_s=round(float(np.mean(times[1:] or times)), 4),
                   cuda_first_run_s=round(times[0], 4))
        if args.cpu and n <= args.cpu_max:
            from oracle import ldc_oracle as orc
            t0 = time.time()
            orc.run_simulation(Graph(1.0, 1.0, n, n, 1e-2, 1.0, 1.0, args.re), None, nthreads=0)
            row['cpu_petsc_equivalent_s'] = round(time.time() - t0, 3)
            row['speedup'] = round(row['cpu_petsc_equivalent_s'] / row['cuda_mean_s'], 1)
        rows.append(row)
        print(json.dumps(row), flush=True)


if __name__ == '__main__':
    main()
