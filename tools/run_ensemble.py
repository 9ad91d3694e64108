"""Reynolds ensemble (BASELINE config 4): `--count` independent n x n cavities, Re log-spaced
100..10000, batch-sharded over the visible GPUs (launch with torchrun for N > 1).

    python tools/run_ensemble.py --count 512 --grid 64
    python -m torch.distributed.run --nproc-per-node 8 tools/run_ensemble.py --count 512 --grid 64
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
    ap.add_argument('--count', type=int, default=512)
    ap.add_argument('--grid', type=int, default=64)
    ap.add_argument('--re-min', type=float, default=100.0)
    ap.add_argument('--re-max', type=float, default=10000.0)
    ap.add_argument('--max-steps', type=int, default=None)
    ap.add_argument('--interleave', action='store_true', help='deal cavities round-robin to the ranks')
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from lid_driven_cavity_problem_b200.ensemble import run_ensemble, STEADY, FAILED
    from lid_driven_cavity_problem_b200.staggered_grid import Graph
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    K = args.count
    res = args.re_min * (args.re_max / args.re_min) ** (np.arange(K) / max(K - 1, 1))
    graphs = [Graph(1.0, 1.0, args.grid, args.grid, 1e-2, 1.0, 1.0, float(r)) for r in res]
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.time()
    summ, _ = run_ensemble(graphs, max_steps=args.max_steps, distributed=world > 1, interleave=args.interleave)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sec = time.time() - t0
    if int(os.environ.get('RANK', '0')) == 0:
        st = np.array([s['status'] for s in summ])
        print(json.dumps(dict(cavities=K, grid='%dx%d' % (args.grid, args.grid), gpus=world, seconds=round(sec, 3),
                              steady=int((st == STEADY).sum()), failed=int((st == FAILED).sum()),
                              steps_total=int(sum(s['steps'] for s in summ)),
                              retries_total=int(sum(s['retries'] for s in summ)),
                              newton_total=int(sum(s['newton'] for s in summ)),
                              krylov_total=int(sum(s['krylov'] for s in summ)),
                              cavities_per_s=round(K / sec, 2), rank0_counters=getattr(run_ensemble, 'last_counters', None))))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
