for k in 300 150 100; do
python - <<PY
import sys, json, time
sys.path.insert(0,'.')
import numpy as np, torch
from lid_driven_cavity_problem_b200.ensemble import run_ensemble
from lid_driven_cavity_problem_b200.staggered_grid import Graph
K=64
res = 100.0 * (100.0) ** (np.arange(K) / (K - 1))
graphs = [Graph(1.0, 1.0, 64, 64, 1e-2, 1.0, 1.0, float(r)) for r in res]
t0=time.time(); summ,_ = run_ensemble(graphs, ksp_max_it=$k); torch.cuda.synchronize(); sec=time.time()-t0
print('ksp_max_it=$k', round(sec,2), 'steady', sum(s['status']==1 for s in summ), 'retries', sum(s['retries'] for s in summ), 'krylov', sum(s['krylov'] for s in summ), run_ensemble.last_counters)
PY
done
for k in 150 100; do echo -n "1024 ksp_max_it=$k: "; timeout 300 python tools/run_cavity.py --n 1024 --re 1000 --opt ksp_max_it=$k 2>&1 | tail -1 | cut -c1-200; done
