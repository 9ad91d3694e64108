import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # a gpu-marked test on a box without a device is an error of the invocation, not a pass
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name))
    return load


def make_graph(params, phi_old_scale=None, U=None, V=None, P=None):
    """params = [size_x, size_y, nx, ny, dt, rho, mi, bc]"""
    from lid_driven_cavity_problem_b200.staggered_grid import Graph
    sx, sy, nx, ny, dt, rho, mi, bc = [float(v) for v in params[:8]]
    g = Graph(sx, sy, int(nx), int(ny), dt, rho, mi, bc)
    if U is not None:
        g.ns_x_mesh.phi = np.array(U, dtype=np.float64)
        g.ns_y_mesh.phi = np.array(V, dtype=np.float64)
        g.pressure_mesh.phi = np.array(P, dtype=np.float64)
        if phi_old_scale is not None:
            g.ns_x_mesh.phi_old = phi_old_scale * g.ns_x_mesh.phi
            g.ns_y_mesh.phi_old = phi_old_scale * g.ns_y_mesh.phi
    return g


def synthetic_case(nx, ny, bc=1000.0, dt=1e-2, seed=None):
    """SURVEY 8(d) synthetic throughput inputs.  Returns (graph, X)."""
    from lid_driven_cavity_problem_b200.staggered_grid import Graph
    rng = np.random.default_rng(1000 + nx if seed is None else seed)
    g = Graph(1.0, 1.0, nx, ny, dt, 1.0, 1.0, bc)
    U = bc * rng.uniform(-1, 1, (nx - 1) * ny)
    V = bc * rng.uniform(-1, 1, nx * (ny - 1))
    P = bc * bc * rng.uniform(-1, 1, nx * ny)
    g.ns_x_mesh.phi, g.ns_y_mesh.phi, g.pressure_mesh.phi = U, V, P
    g.ns_x_mesh.phi_old = 0.9 * U
    g.ns_y_mesh.phi_old = 0.9 * V
    X = np.zeros(3 * nx * ny)
    X[0::3] = P
    Xu = np.zeros((ny, nx))
    Xu[:, :nx - 1] = U.reshape(ny, nx - 1)
    X[1::3] = Xu.ravel()
    Xv = np.zeros(nx * ny)
    Xv[:nx * (ny - 1)] = V
    X[2::3] = Xv
    return g, X
