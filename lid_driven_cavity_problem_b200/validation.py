"""Centre-line validation against Ghia, Ghia & Shin (1982) as NUMBERS.

The reference only plots (experiments/run_solver.py:90-118, experiments/validation.py:70-99): it
extracts u along the vertical centre line and v along the horizontal centre line, normalises by
the lid speed and overlays the Ghia table points.  ``centre_lines`` reproduces that extraction
exactly (same reshape, same ``len(U) // 2`` index, same ``linspace`` abscissa that ignores the
half-cell offset of the staggered positions); ``compare_with_ghia`` interpolates the extracted
curves at the table positions and reports rms / max differences.

Tables: position, Re=100, Re=400 are the columns shipped with the reference
(experiments/ghia_ghia_shin_results/{U,V}.txt); the Re=1000 column is NOT shipped by the
reference -- it is the paper's Table I/II as recalled in SURVEY.md App. F (treat the last digit
with care).
"""
import numpy as np

# y, u(Re=100), u(Re=400), u(Re=1000) along x = 0.5
GHIA_U = np.array([
    [1.0000, 1.00000, 1.00000, 1.00000],
    [0.9766, 0.84123, 0.75837, 0.65928],
    [0.9688, 0.78871, 0.68439, 0.57492],
    [0.9609, 0.73722, 0.61756, 0.51117],
    [0.9531, 0.68717, 0.55892, 0.46604],
    [0.8516, 0.23151, 0.29093, 0.33304],
    [0.7344, 0.00332, 0.16256, 0.18719],
    [0.6172, -0.13641, 0.02135, 0.05702],
    [0.5000, -0.20581, -0.11477, -0.06080],
    [0.4531, -0.21090, -0.17119, -0.10648],
    [0.2813, -0.15662, -0.32726, -0.27805],
    [0.1719, -0.10150, -0.24299, -0.38289],
    [0.1016, -0.06434, -0.14612, -0.29730],
    [0.0703, -0.04775, -0.10338, -0.22220],
    [0.0625, -0.04192, -0.09266, -0.20196],
    [0.0547, -0.03717, -0.08186, -0.18109],
    [0.0000, 0.00000, 0.00000, 0.00000]])
# x, v(Re=100), v(Re=400), v(Re=1000) along y = 0.5
GHIA_V = np.array([
    [1.0000, 0.00000, 0.00000, 0.00000],
    [0.9688, -0.05906, -0.12146, -0.21388],
    [0.9609, -0.07391, -0.15663, -0.27669],
    [0.9531, -0.08864, -0.19254, -0.33714],
    [0.9453, -0.10313, -0.22847, -0.39188],
    [0.9063, -0.16914, -0.23827, -0.51550],
    [0.8594, -0.22445, -0.44993, -0.42665],
    [0.8047, -0.24533, -0.38598, -0.31966],
    [0.5000, 0.05454, 0.05188, 0.02526],
    [0.2344, 0.17527, 0.30174, 0.32235],
    [0.2266, 0.17507, 0.30203, 0.33075],
    [0.1563, 0.16077, 0.28124, 0.37095],
    [0.0938, 0.12317, 0.22965, 0.32627],
    [0.0781, 0.10890, 0.20920, 0.30353],
    [0.0703, 0.10091, 0.19713, 0.29012],
    [0.0625, 0.09233, 0.18360, 0.27485],
    [0.0000, 0.00000, 0.00000, 0.00000]])
_COLUMN = {100: 1, 400: 2, 1000: 3}


def centre_lines(graph, size_x=1.0, size_y=1.0):
    """(y, u/bc) along the vertical centre line and (x, v/bc) along the horizontal one, extracted
    as experiments/run_solver.py:90-118 does."""
    nx, ny = graph.pressure_mesh.nx, graph.pressure_mesh.ny
    U = np.asarray(graph.ns_x_mesh.phi, dtype=np.float64).reshape(ny, nx - 1) / graph.bc
    V = np.asarray(graph.ns_y_mesh.phi, dtype=np.float64).reshape(ny - 1, nx) / graph.bc
    u_line = U[:, min(len(U) // 2, nx - 2)]
    v_line = V[len(V) // 2, :]
    return (np.linspace(0.0, size_y, len(u_line)), u_line), (np.linspace(0.0, size_x, len(v_line)), v_line)


def centre_lines_staggered(graph, size_x=1.0, size_y=1.0):
    """the same two profiles on their TRUE coordinates: U[i,j] sits at y=(j+1/2)dy, V[i,j] at
    x=(i+1/2)dx; the wall / lid values close the curves (u(0)=0, u(size_y)=1, v(0)=v(size_x)=0)."""
    nx, ny = graph.pressure_mesh.nx, graph.pressure_mesh.ny
    (_, u), (_, v) = centre_lines(graph, size_x, size_y)
    y = np.r_[0.0, (np.arange(ny) + 0.5) * (size_y / ny), size_y]
    x = np.r_[0.0, (np.arange(nx) + 0.5) * (size_x / nx), size_x]
    return (y, np.r_[0.0, u, 1.0]), (x, np.r_[0.0, v, 0.0])


def compare_with_ghia(graph, reynolds, size_x=1.0, size_y=1.0):
    """rms and max |difference| between the centre-line profiles and the Ghia table points, on the
    true staggered coordinates; the ``*_as_plotted`` numbers use the reference's plotting abscissa
    (``linspace``, run_solver.py:108,117), which shifts the curves by up to half a cell."""
    key = int(round(reynolds))
    if key not in _COLUMN:
        raise ValueError("Ghia tables are available for Re in %s" % sorted(_COLUMN))
    col = _COLUMN[key]
    (yp, up), (xp, vp) = centre_lines(graph, size_x, size_y)
    dup = np.interp(GHIA_U[:, 0] * size_y, yp, up) - GHIA_U[:, col]
    dvp = np.interp(GHIA_V[:, 0] * size_x, xp, vp) - GHIA_V[:, col]
    (y, u), (x, v) = centre_lines_staggered(graph, size_x, size_y)
    du = np.interp(GHIA_U[:, 0] * size_y, y, u) - GHIA_U[:, col]
    dv = np.interp(GHIA_V[:, 0] * size_x, x, v) - GHIA_V[:, col]
    return dict(Re=key, rms_u=float(np.sqrt(np.mean(du ** 2))), rms_v=float(np.sqrt(np.mean(dv ** 2))),
                max_u=float(np.abs(du).max()), max_v=float(np.abs(dv).max()),
                rms_u_as_plotted=float(np.sqrt(np.mean(dup ** 2))), rms_v_as_plotted=float(np.sqrt(np.mean(dvp ** 2))),
                u_min=float(u.min()), v_min=float(v.min()), v_max=float(v.max()),
                source='reference table' if col < 3 else 'Ghia et al. 1982 Re=1000 (SURVEY App. F, not shipped by the reference)')
