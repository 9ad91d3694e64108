"""Row-strip distributed steady solve of one cavity (BASELINE config 5).

    python -m torch.distributed.run --nproc-per-node 8 tools/run_strips.py --grid 4096 --re 1000
    (single GPU: python tools/run_strips.py --grid 1024)

--check gathers the distributed solution and compares it with a single-GPU solve on rank 0.
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
    ap.add_argument('--grid', type=int, default=1024)
    ap.add_argument('--re', type=float, default=1000.0)
    ap.add_argument('--dt', type=float, default=1e-2)
    ap.add_argument('--max-steps', type=int, default=None)
    ap.add_argument('--check', action='store_true')
    ap.add_argument('--opt', nargs='*', default=[])
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from lid_driven_cavity_problem_b200 import _comm
    from lid_driven_cavity_problem_b200.strip_solver import CudaStripSolver
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    comm = None
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
        comm = _comm.Comm()
    opts = {}
    for kv in args.opt:
        k, v = kv.split('=')
        opts[k] = float(v) if ('.' in v or 'e' in v) else int(v)
    n = args.grid
    solver = CudaStripSolver(n, n, args.dt, 1.0, 1.0, args.re, comm=comm, rank=rank, world=world, **opts)
    solver.set_uniform_state()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.time()
    t_end = solver.run(max_steps=args.max_steps)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sec = time.time() - t0
    out = dict(grid='%dx%d' % (n, n), Re=args.re, gpus=world, seconds=round(sec, 3), steps=solver.steps,
               retries=solver.retries, newton_its=solver.newton_iterations, krylov_its=solver.krylov_iterations,
               pseudo_time=t_end, rows_per_gpu=solver.row_count)
    if args.check:
        X = solver.gather_state().cpu().numpy()
        if rank == 0:
            from lid_driven_cavity_problem_b200.nonlinear_solver import cuda_solver_wrapper
            from lid_driven_cavity_problem_b200.staggered_grid import Graph
            from lid_driven_cavity_problem_b200.time_stepper import run_simulation
            w = cuda_solver_wrapper.CudaSolverWrapper(None, **opts)
            ref = run_simulation(Graph(1.0, 1.0, n, n, args.dt, 1.0, 1.0, args.re), None, w.solve,
                                 max_steps=args.max_steps)
            U = X[1::3].reshape(n, n)[:, :n - 1].ravel()
            V = X[2::3][:n * (n - 1)]
            out['check'] = dict(max_dU_over_bc=float(np.abs(U - ref.ns_x_mesh.phi).max() / args.re),
                                max_dV_over_bc=float(np.abs(V - ref.ns_y_mesh.phi).max() / args.re),
                                single_gpu_krylov=w.total_linear_iterations, single_gpu_newton=w.total_newton_iterations)
    if rank == 0:
        print(json.dumps(out))
    solver.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
