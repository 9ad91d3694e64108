"""Multi-GPU check of the collective-free strip residual (ldc_residual_peer_push +
ldc_residual_peer) on uneven row strips.

    python -m torch.distributed.run --nproc-per-node 4 --master-addr 127.0.0.1 tools/check_peer_halo.py --nx 250 --ny 203

Every rank pushes its boundary rows into the neighbours' ghost rows (peer-mapped memory), evaluates
its strip once their flag words arrive and compares it, bit for bit, with its slice of the
full-grid residual (ldc_residual on the same GPU).  Repeats for a few
epochs with changed X (an all-reduce in between keeps the buffers safe to overwrite).
"""
import argparse
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--nx', type=int, default=250)
    ap.add_argument('--ny', type=int, default=203)
    ap.add_argument('--epochs', type=int, default=3)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from lid_driven_cavity_problem_b200 import _comm, _native, peer_halo
    from lid_driven_cavity_problem_b200.residual_function import cuda_residual_function as crf
    from lid_driven_cavity_problem_b200.staggered_grid import Graph
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
    if world == 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29631')
    dist.init_process_group('nccl', rank=rank, world_size=world,
                            device_id=torch.device('cuda', torch.cuda.current_device()))
    comm = _comm.Comm()
    nx, ny = args.nx, args.ny
    rng = np.random.default_rng(7)          # same numbers on every rank
    g = Graph(1.0, 1.0, nx, ny, 1e-2, 1.0, 1.0, 1000.0)
    g.ns_x_mesh.phi_old[:] = rng.uniform(-1, 1, (nx - 1) * ny)
    g.ns_y_mesh.phi_old[:] = rng.uniform(-1, 1, nx * (ny - 1))
    parts = [_comm.strip_partition(ny, world, r) for r in range(world)]
    j0, rows = parts[rank]
    rows_max = max(c for _, c in parts)
    region = comm.peer_alloc(peer_halo.region_bytes(nx, rows_max, 1))
    grid = _native.grid_from_graph(g, row_begin=j0, row_count=rows)
    strip = peer_halo.PeerStripResidual(grid, region[0], region[1], region[2],
                                        parts[rank - 1][1] if rank > 0 else 0, rows_max)
    U = torch.from_numpy(g.ns_x_mesh.phi_old.reshape(ny, nx - 1)[j0:j0 + rows].copy().ravel()).cuda()
    Vl = g.ns_y_mesh.phi_old.reshape(ny - 1, nx)[j0:min(j0 + rows, ny - 1)].copy().ravel()
    V = torch.from_numpy(Vl if Vl.size else np.zeros(1)).cuda()
    F = torch.empty(3 * nx * rows, dtype=torch.float64, device='cuda')
    token = torch.zeros(1, dtype=torch.float64, device='cuda')
    bad = 0
    for epoch in range(args.epochs):
        X = rng.uniform(-2, 2, 3 * nx * ny) * (1.0 + epoch)
        strip.x_tensor(0).copy_(torch.from_numpy(X[3 * nx * j0:3 * nx * (j0 + rows)].copy()))
        torch.cuda.synchronize()
        strip.residual(0, U, V, F)
        F_full = crf.residual_function_device(torch.from_numpy(X).cuda(), g)
        same = torch.equal(F, F_full[3 * nx * j0:3 * nx * (j0 + rows)])
        bad += 0 if same and not strip.status() else 1
        comm.allreduce(token)               # nobody overwrites X before everybody has read it
        torch.cuda.synchronize()
    t = torch.tensor([float(bad)], dtype=torch.float64, device='cuda')
    dist.all_reduce(t)
    if rank == 0:
        print('peer halo check %dx%d on %d strips %s: %s' % (nx, ny, world, [c for _, c in parts],
                                                           'bit-identical' if t.item() == 0 else 'MISMATCH'))
    comm.peer_free(region[0])
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 0 else 1)


if __name__ == '__main__':
    main()
