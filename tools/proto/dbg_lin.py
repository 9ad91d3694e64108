import sys, ctypes
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
from test_cuda_solver import _developed_graph
from oracle import ldc_oracle as orc
from lid_driven_cavity_problem_b200 import _native
lib = _native.lib()
for nx, ny, pc in [(32,32,2),(48,40,2),(48,40,1)]:
    g = _developed_graph(nx, ny, 100.0)
    X = orc.create_X(g.ns_x_mesh.phi, g.ns_y_mesh.phi, g.pressure_mesh.phi, g)
    g.ns_x_mesh.phi_old = g.ns_x_mesh.phi.copy(); g.ns_y_mesh.phi_old = g.ns_y_mesh.phi.copy(); g.dt = 0.05
    F = orc.residual_function(X, g)
    mask = orc.jacobian_mask_arrays(nx, ny, 3)
    J = orc.fd_jacobian(X, g, F0=F, mask=mask)
    y_ref = orc._direct_solve(mask, J, F)
    for rtol in [1e-5, 1e-7, 1e-9]:
        opt = _native.LdcSolverOptions(); lib.ldc_solver_default_options(ctypes.byref(opt))
        opt.ksp_rtol, opt.pc_type, opt.ksp_max_it = rtol, pc, 3000
        grid = _native.grid_from_graph(g); h = ctypes.c_void_p()
        _native.check(lib.ldc_solver_create(ctypes.byref(grid), ctypes.byref(opt), ctypes.byref(h)))
        st = _native.current_stream()
        Xd, Fd = torch.from_numpy(X).cuda(), torch.from_numpy(F).cuda(); yd = torch.empty_like(Xd)
        _native.check(lib.ldc_solver_set_state(h, _native.ptr(Xd), st))
        its, reason, rel = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_double()
        _native.check(lib.ldc_solver_linear_solve(h, _native.ptr(Fd), _native.ptr(yd), ctypes.byref(its), ctypes.byref(reason), ctypes.byref(rel), st))
        y = yd.cpu().numpy()
        r = F - orc.spmv(mask, J, y)
        print(nx, ny, 'pc', pc, 'rtol', rtol, 'its', its.value, 'reason', reason.value, 'est', rel.value, 'true', np.linalg.norm(r)/np.linalg.norm(F),
              '|y|', np.linalg.norm(y), 'meanP', y[0::3].mean(), '|y_ref|', np.linalg.norm(y_ref), 'ref meanP', y_ref[0::3].mean())
        lib.ldc_solver_destroy(h)
