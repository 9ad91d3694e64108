"""State containers of the staggered grid, API-compatible with the reference
(staggered_grid.py:1-50): ``Mesh2d(nx, ny, initial_phi)`` with ``phi`` / ``phi_old`` and
``Graph(size_x, size_y, nx, ny, dt, rho, mi, bc, initial_P, initial_U, initial_V)``.

Three meshes share one graph: pressure at cell centres (nx x ny), U on east faces
((nx-1) x ny) and V on north faces (nx x (ny-1)); all are flat row-major arrays.
Arrays are always float64 (the reference lets the dtype follow the Python scalar).
"""
import numpy as np


class Mesh2d(object):
    __slots__ = ('nx', 'ny', 'phi', 'phi_old')

    def __init__(self, nx, ny, initial_phi):
        self.nx = int(nx)
        self.ny = int(ny)
        self.phi = np.full(self.nx * self.ny, initial_phi, dtype=np.float64)
        self.phi_old = self.phi.copy()

    def __len__(self):
        return self.nx * self.ny


class Graph(object):
    __slots__ = ('dt', 'dx', 'dy', 'rho', 'mi', 'pressure_mesh', 'ns_x_mesh', 'ns_y_mesh', 'bc', 'use_cds')

    def __init__(self, size_x, size_y, nx, ny, dt, rho, mi, bc, initial_P=1.0, initial_U=1.0, initial_V=1.0):
        if nx < 3 or ny < 2:
            # nx == 2 makes two mask diagonals collide (reference TODO,
            # tests/test_calculate_jacobian_mask.py:20-22); ny == 1 leaves no V face.
            raise ValueError("staggered grid needs nx >= 3 and ny >= 2, got %sx%s" % (nx, ny))
        self.dt = dt
        self.dx = size_x / nx
        self.dy = size_y / ny
        self.rho = rho
        self.mi = mi
        self.bc = bc          # lid speed (U above the top row)
        self.use_cds = False  # central differencing for advection instead of first-order upwind
        self.pressure_mesh = Mesh2d(nx, ny, initial_P)
        self.ns_x_mesh = Mesh2d(nx - 1, ny, initial_U)
        self.ns_y_mesh = Mesh2d(nx, ny - 1, initial_V)

    def __len__(self):
        return len(self.pressure_mesh) + len(self.ns_x_mesh) + len(self.ns_y_mesh)
