"""Host<->device copy bandwidth of this box (pinned memory), alone and both directions at once:
the floor of the host-buffer plug-in's e2e time.  python tools/pcie_probe.py"""
import json
import time

import torch

n = 42 * 1024 * 1024 // 8
h_in = torch.empty(n, dtype=torch.float64).pin_memory()
h_out = torch.empty(n, dtype=torch.float64).pin_memory()
d_in = torch.empty(n, dtype=torch.float64, device='cuda')
d_out = torch.empty(n, dtype=torch.float64, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def h2d():
    d_in.copy_(h_in, non_blocking=True)


def d2h():
    h_out.copy_(d_out, non_blocking=True)


def both():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


pageable = torch.empty(n, dtype=torch.float64)


def h2d_pageable():
    d_in.copy_(pageable)


def stage():
    h_in.copy_(pageable)


gb = n * 8 / 1e9
out = {"bytes": n * 8, "h2d_pinned_gbs": gb / timed(h2d), "d2h_pinned_gbs": gb / timed(d2h),
       "both_each_gbs": gb / timed(both), "h2d_pageable_gbs": gb / timed(h2d_pageable),
       "host_memcpy_pageable_to_pinned_gbs": gb / timed(stage), "torch_threads": torch.get_num_threads()}
print(json.dumps(out))
