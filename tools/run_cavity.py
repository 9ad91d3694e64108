"""Run one steady (or fixed-time) cavity with the cuda backend and print timing + solver counts.

    python tools/run_cavity.py --n 1024 --re 1000 [--dt 1e-2] [--final-time T] [--opt key=value ...]

Equivalent of the reference's experiments/run_solver.py (wall time around run_simulation,
run_solver.py:65-67) with numbers instead of plots.
"""
import argparse
import json
import logging
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--n', type=int, default=100)
    ap.add_argument('--ny', type=int, default=0)
    ap.add_argument('--re', type=float, default=1000.0)
    ap.add_argument('--dt', type=float, default=1e-2)
    ap.add_argument('--final-time', type=float, default=None)
    ap.add_argument('--max-steps', type=int, default=None)
    ap.add_argument('--opt', nargs='*', default=[])
    ap.add_argument('--verbose', action='store_true')
    ap.add_argument('--ghia', action='store_true')
    args = ap.parse_args()
    if args.verbose:
        logging.basicConfig(stream=sys.stdout, level=logging.INFO)
    import torch
    from lid_driven_cavity_problem_b200.nonlinear_solver import solver_builder, cuda_solver_wrapper
    from lid_driven_cavity_problem_b200.staggered_grid import Graph
    from lid_driven_cavity_problem_b200.time_stepper import run_simulation
    opts = {}
    for kv in args.opt:
        k, v = kv.split('=')
        opts[k] = float(v) if ('.' in v or 'e' in v) else int(v)
    nx, ny = args.n, args.ny or args.n
    wrapper = cuda_solver_wrapper.CudaSolverWrapper(None, **opts)
    graph = Graph(1.0, 1.0, nx, ny, args.dt, 1.0, 1.0, args.re)
    log = []

    def on_step(g, t):
        r = g.newton_report
        log.append(dict(t=t, dt=g.dt, newton=r.iterations, krylov=r.linear_iterations, reason=r.reason,
                        fnorm0=r.fnorm0, fnorm=r.fnorm, du=g.steady_state_norm()))
        if args.verbose:
            print(json.dumps(log[-1]), flush=True)
    torch.cuda.synchronize()
    t0 = time.time()
    result = run_simulation(graph, args.final_time, wrapper.solve, max_steps=args.max_steps, on_step=on_step)
    torch.cuda.synchronize()
    elapsed = time.time() - t0
    out = dict(grid='%dx%d' % (nx, ny), Re=args.re, seconds=round(elapsed, 4), steps=len(log),
               newton_its=wrapper.total_newton_iterations, krylov_its=wrapper.total_linear_iterations,
               function_evaluations=wrapper.total_function_evaluations,
               jacobian_evaluations=wrapper.total_jacobian_evaluations,
               final_dt=log[-1]['dt'] if log else None, final_du=log[-1]['du'] if log else None,
               options=opts)
    if args.ghia:
        from lid_driven_cavity_problem_b200 import validation
        out['ghia'] = validation.compare_with_ghia(result, args.re)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
