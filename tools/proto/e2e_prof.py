import sys, time
sys.path.insert(0,'/root/repo')
import numpy as np, torch
import bench
from lid_driven_cavity_problem_b200.residual_function import cuda_residual_function as crf
g, X = bench.synthetic_case(1024,1024)
n = X.size
for _ in range(3): crf.residual_function(X, g)
def t(fn, reps=20):
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/reps*1e3
pin = torch.empty(n, dtype=torch.float64, pin_memory=True)
src = torch.from_numpy(X)
print('threads', torch.get_num_threads())
print('stage X (copy_ into pinned) ms', t(lambda: pin.copy_(src)))
d = torch.empty(n, dtype=torch.float64, device='cuda')
print('H2D pinned 25MB ms', t(lambda: d.copy_(pin, non_blocking=True)))
print('D2H pinned 25MB ms', t(lambda: pin.copy_(d, non_blocking=True)))
out = np.empty(n); outt = torch.from_numpy(out)
print('copy out of pinned ms', t(lambda: outt.copy_(pin)))
print('np.empty+copy ms', t(lambda: torch.from_numpy(np.empty(n)).copy_(pin)))
print('H2D pageable 25MB ms', t(lambda: d.copy_(src)))
print('full call ms', t(lambda: crf.residual_function(X, g)))
