"""Where the time of a Reynolds ensemble goes: the same 64 cavities (Re 100..10000, 64x64) as one
batch, and split into sub-batches by difficulty.  python tools/ensemble_probe.py"""
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402
from lid_driven_cavity_problem_b200.ensemble import CudaEnsembleSolver  # noqa: E402
from lid_driven_cavity_problem_b200.staggered_grid import Graph  # noqa: E402

K, n = 64, 64
res = 100.0 * 100.0 ** (np.arange(K) / (K - 1))


def run(idx, tag):
    graphs = [Graph(1.0, 1.0, n, n, 1e-2, 1.0, 1.0, float(res[i])) for i in idx]
    torch.cuda.synchronize()
    t0 = time.time()
    s = CudaEnsembleSolver(graphs)
    ts = s.run()
    torch.cuda.synchronize()
    sec = time.time() - t0
    c = s.counters()
    print(json.dumps(dict(tag=tag, cavities=len(idx), seconds=round(sec, 3), sweeps=int((ts.steps + ts.retries).max()),
                          lockstep_krylov=c['lockstep_krylov'], lockstep_newton=c['lockstep_newton'],
                          krylov_sum=c['krylov_sum'], retries=[int(v) for v in ts.retries],
                          steps=[int(v) for v in ts.steps])))
    s.close()


run(list(range(8)), 'warm-up')
run(list(range(K)), 'all 64')
run(list(range(0, 32)), 'easiest 32')
run(list(range(32, 56)), 'next 24')
run(list(range(56, 64)), 'hardest 8')
run([63], 'hardest 1')
