"""Drop-in for the reference's pylamp_stokes module (pylamp_stokes.py:17-35, :86-101, :104-563).

`makeStokesMatrix` keeps the reference's signature but returns a light *operator handle* in place
of the (3N x 3N) scipy lil_matrix: the system is matrix-free on the GPU (at 4097^2 the assembled
matrix alone would be ~5 GB).  The handle supports `A @ x` / `A.dot(x)` (what
`A_ref @ x` gives, every row class included) and `A.solve(rhs)`; `pylamp_b200.solve.spsolve(A, rhs)`
is the replacement for the driver's `scipy.sparse.linalg.spsolve(scipy.sparse.csc_matrix(A), rhs)`
(pylamp2.py:360).  No CPU path.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .pylamp_const import *  # noqa: F401,F403
from .pylamp_const import DIM, G, IP, IX, IZ

# pylamp_stokes.py:17-20
BC_TYPE_NOSLIP = 0
BC_TYPE_FREESLIP = 1
BC_TYPE_CYCLIC = 2
BC_TYPE_FLOWTHRU = 4

DEFAULT_RTOL = 1e-12
DEFAULT_MAXIT = 600


def gidx(idxs, nx, dim):
    """Global DOF index of node `idxs` (pylamp_stokes.py:22-35)."""
    if len(idxs) != dim:
        raise Exception("num of idxs != dimensions")
    if dim == 2:
        return idxs[IZ] * nx[IX] * (dim + 1) + idxs[IX] * (dim + 1)
    print("!!! NOT IMPLEMENTED")                                             # :32-33


def x2vp(x, nx):
    """De-interleave the solution vector into ([vz, vx], P/Kcont) -- pylamp_stokes.py:86-101.
    Works on NumPy arrays (views, like the reference) and on CUDA tensors (device kernel)."""
    if isinstance(x, torch.Tensor) and x.is_cuda:
        ctx = _lib.default_context(x.device.index)
        nz, nxx = int(nx[IZ]), int(nx[IX])
        out = [torch.empty((nz, nxx), dtype=torch.float64, device=x.device) for _ in range(3)]
        ctx.call("plb_x2vp", nz, nxx, nxx, x.data_ptr(), out[0].data_ptr(), out[1].data_ptr(),
                 out[2].data_ptr())
        return [out[0], out[1]], out[2]
    newvel = [[]] * DIM
    newvel[IZ] = x[(IZ)::(DIM + 1)].reshape(nx)
    newvel[IX] = x[(IX)::(DIM + 1)].reshape(nx)
    newpres = x[(IP)::(DIM + 1)].reshape(nx)
    return newvel, newpres


class StokesOperator:
    """Device-resident Stokes system in the reference's DOF layout and row scaling."""

    def __init__(self, nx, grid, f_etas, f_etan, f_rho, bc, ctx=None):
        self.host = not (isinstance(f_etas, torch.Tensor) and f_etas.is_cuda)
        self.ctx = ctx or _lib.default_context(None if self.host else f_etas.device.index)
        ctx = self.ctx
        self.nz, self.nxx = int(nx[IZ]), int(nx[IX])
        self.shape = (3 * self.nz * self.nxx,) * 2
        self.dtype = np.float64
        self.bc = [int(b) for b in bc]
        gz = np.ascontiguousarray(_np(grid[IZ]), dtype=np.float64)
        gx = np.ascontiguousarray(_np(grid[IX]), dtype=np.float64)
        h = C.c_void_p()
        ctx.check(ctx.lib.plb_stokes_create(ctx.h, self.nz, self.nxx, self.nxx,
                                            gz.ctypes.data_as(_lib.DP), gx.ctypes.data_as(_lib.DP),
                                            _lib.int_array(self.bc), C.byref(h)))
        self.h = h
        self.warn_unconverged = True
        self.converged = None
        self.set_coeffs(f_etas, f_etan, f_rho)

    def set_coeffs(self, f_etas, f_etan, f_rho):
        """(Re)bind the viscosity/density fields; the grid and BCs of the handle are kept."""
        ctx = self.ctx
        self._fields = [_dev(f, ctx) for f in (f_etas, f_etan, f_rho)]     # keep alive
        for f in self._fields:
            assert tuple(f.shape) == (self.nz, self.nxx)
        ctx.check(ctx.lib.plb_stokes_set_coeffs(self.h, self._fields[0].data_ptr(),
                                                self._fields[1].data_ptr(), self._fields[2].data_ptr(),
                                                float(G[IZ]), float(G[IX])))

    def set_param(self, name, value):
        self.ctx.check(self.ctx.lib.plb_stokes_set_param(self.h, name.encode(), float(value)))

    def set_surfstab(self, tstep=None, surfstab_theta=0.5):
        """Free-surface stabilisation terms of ``makeStokesMatrix(surfstab=True, tstep, surfstab_theta)``
        (pylamp_stokes.py:422-426, :483-487) for the current density field; ``tstep=None`` switches
        them off.  `set_coeffs` switches them off as well (they belong to a density field)."""
        theta_dt = 0.0 if tstep is None else float(surfstab_theta) * float(tstep)
        self.ctx.check(self.ctx.lib.plb_stokes_set_surfstab(self.h, theta_dt))

    @property
    def scaling(self):
        """(Kcont, Kbond) of pylamp_stokes.py:116-122."""
        out = (C.c_double * 2)()
        self.ctx.check(self.ctx.lib.plb_stokes_scaling(self.h, out))
        return out[0], out[1]

    def rhs(self, device=False):
        r = torch.empty(self.shape[0], dtype=torch.float64, device=self.ctx.torch_device)
        self.ctx.check(self.ctx.lib.plb_stokes_rhs(self.h, r.data_ptr()))
        return r if device else r.cpu().numpy()

    def dot(self, x):
        host = not isinstance(x, torch.Tensor)
        xd = _dev(x, self.ctx).reshape(-1)
        assert xd.shape[0] == self.shape[0]
        y = torch.empty_like(xd)
        self.ctx.check(self.ctx.lib.plb_stokes_apply(self.h, xd.data_ptr(), y.data_ptr()))
        return y.cpu().numpy() if host else y

    __matmul__ = dot

    def vcycle(self, b2):
        """Test hook: one velocity-block V-cycle on planar [vz | vx] data."""
        bd = _dev(b2, self.ctx).reshape(-1)
        x = torch.empty_like(bd)
        self.ctx.check(self.ctx.lib.plb_stokes_vcycle(self.h, bd.data_ptr(), x.data_ptr()))
        if self.ctx.comm_info()[1] > 1:
            self.ctx.allreduce(x)
        return x

    def solve(self, rhs=None, rtol=DEFAULT_RTOL, maxit=DEFAULT_MAXIT, raise_on_fail=True):
        """x = A^-1 rhs on the GPU; same layout as the reference's spsolve result."""
        host = self.host if rhs is None else not isinstance(rhs, torch.Tensor)
        rd = None if rhs is None else _dev(rhs, self.ctx).reshape(-1)
        # (slab-local fields: the solver writes the own rows and the halo rows only; the rest stays zero)
        new = torch.zeros if self.ctx.slab is not None else torch.empty
        x = new(self.shape[0], dtype=torch.float64, device=self.ctx.torch_device)
        it, rr = C.c_int(0), C.c_double(0)
        rc = self.ctx.lib.plb_stokes_solve(self.h, None if rd is None else rd.data_ptr(), float(rtol),
                                           int(maxit), x.data_ptr(), C.byref(it), C.byref(rr))
        self.iterations, self.relres = it.value, rr.value
        if rc != 0 and raise_on_fail:
            self.ctx.check(rc)
        self.converged = rr.value <= 1.5 * float(rtol)
        if rc == 0 and not self.converged and self.warn_unconverged:
            import warnings
            warnings.warn("Stokes solve stopped at relres %.2e > rtol %.1e (fp64 residual floor or stagnation; "
                          "accepted below rtol_accept) -- see StokesOperator.stats" % (rr.value, rtol))
        if self.ctx.comm_info()[1] > 1 and self.ctx.slab is None:
            self.ctx.allreduce(x)       # every slab rank filled its own rows: sum the pieces
        return x.cpu().numpy() if host else x

    @property
    def stats(self):
        out = (C.c_double * 6)()
        self.ctx.check(self.ctx.lib.plb_stokes_last_stats(self.h, out))
        return {"iterations": int(out[0]), "vcycles": int(out[1]), "relres": out[2], "floor": out[3],
                "status": ("converged", "accepted_above_rtol", "not_converged")[int(out[4])], "rtol_eff": out[5]}

    def close(self):
        if getattr(self, "h", None):
            self.ctx.lib.plb_stokes_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _np(a):
    return a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)


def _dev(a, ctx):
    if isinstance(a, torch.Tensor):
        return a.to(device=ctx.torch_device, dtype=torch.float64).contiguous()
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).to(ctx.torch_device)


def makeStokesMatrix(nx, grid, f_etas, f_etan, f_rho, bc, surfstab=False, tstep=None,
                     surfstab_theta=0.5):
    """Build the Stokes system for the given viscosity/density fields and wall types.

    Reference: pylamp_stokes.py:104-563.  Returns ``(A, rhs)`` like the reference; ``A`` is a
    :class:`StokesOperator` (see the module docstring), ``rhs`` the (3N,) right-hand side."""
    if surfstab and tstep is None:
        raise Exception("surface stabilization needs predetermined tstep")       # :423-424
    if (bc[DIM * 0 + IZ] & BC_TYPE_FLOWTHRU) or (bc[DIM * 1 + IZ] & BC_TYPE_FLOWTHRU):
        raise Exception("BC_TYPE_FLOWTHRU not implemented for z-direction")      # :546
    A = StokesOperator(nx, grid, f_etas, f_etan, f_rho, bc)
    if surfstab:
        A.set_surfstab(tstep, surfstab_theta)                                    # :422-426, :483-487
    return A, A.rhs(device=not A.host)
