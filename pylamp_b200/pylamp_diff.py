"""Drop-in for the reference's pylamp_diff module (pylamp_diff.py:12-28, :78-83, :85-183).

`makeDiffusionMatrix` keeps the reference's signature and returns ``(A, rhs)`` where ``A`` is a
device-resident operator handle (``A @ x``, ``A.solve(rhs)``) instead of the (N x N) lil_matrix;
`pylamp_b200.solve.spsolve(A, rhs)` replaces the driver's spsolve call at pylamp2.py:419.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .pylamp_const import *  # noqa: F401,F403
from .pylamp_const import DIM, IX, IZ

# pylamp_diff.py:12-13
BC_TYPE_FIXTEMP = 0
BC_TYPE_FIXFLOW = 1

DEFAULT_RTOL = 1e-13
DEFAULT_MAXIT = 400


def gidx(idxs, nx):
    """Global DOF index (pylamp_diff.py:15-28: two arguments, the dimension is len(nx))."""
    dim = len(nx)
    if dim == 2:
        return idxs[IZ] * nx[IX] + idxs[IX]
    print("!!! NOT IMPLEMENTED")                                             # :26


def x2t(x, nx):
    """pylamp_diff.py:78-83"""
    return x.reshape(tuple(int(n) for n in nx))


def _np(a):
    return a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)


def _dev(a, ctx):
    if isinstance(a, torch.Tensor):
        return a.to(device=ctx.torch_device, dtype=torch.float64).contiguous()
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).to(ctx.torch_device)


class DiffusionOperator:
    """Device-resident implicit-Euler heat system in the reference's row layout."""

    def __init__(self, nx, grid, gridmp, f_T, f_k, f_Cp, f_rho, f_H, bc, bcvalue, tstep, ctx=None):
        self.host = not (isinstance(f_T, torch.Tensor) and f_T.is_cuda)
        self.ctx = ctx or _lib.default_context(None if self.host else f_T.device.index)
        ctx = self.ctx
        self.nz, self.nxx = int(nx[IZ]), int(nx[IX])
        self.shape = (self.nz * self.nxx,) * 2
        self.dtype = np.float64
        arrs = [np.ascontiguousarray(_np(a), dtype=np.float64)
                for a in (grid[IZ], grid[IX], gridmp[IZ], gridmp[IX])]
        h = C.c_void_p()
        ctx.check(ctx.lib.plb_diff_create(ctx.h, self.nz, self.nxx, self.nxx,
                                          *[a.ctypes.data_as(_lib.DP) for a in arrs],
                                          _lib.int_array(bc), _lib.dbl_array(bcvalue), C.byref(h)))
        self.h = h
        self.set_coeffs(f_T, f_k, f_Cp, f_rho, f_H, tstep)

    def set_coeffs(self, f_T, f_k, f_Cp, f_rho, f_H, tstep):
        ctx = self.ctx
        self._fields = [_dev(f, ctx) for f in (f_T, f_k[IZ], f_k[IX], f_Cp, f_rho, f_H)]
        for f in self._fields:
            assert tuple(f.shape) == (self.nz, self.nxx)
        ctx.check(ctx.lib.plb_diff_set_coeffs(self.h, *[f.data_ptr() for f in self._fields], float(tstep)))

    def rhs(self, device=False):
        r = torch.empty(self.shape[0], dtype=torch.float64, device=self.ctx.torch_device)
        self.ctx.check(self.ctx.lib.plb_diff_rhs(self.h, r.data_ptr()))
        return r if device else r.cpu().numpy()

    def dot(self, x):
        host = not isinstance(x, torch.Tensor)
        xd = _dev(x, self.ctx).reshape(-1)
        y = torch.empty_like(xd)
        self.ctx.check(self.ctx.lib.plb_diff_apply(self.h, xd.data_ptr(), y.data_ptr()))
        return y.cpu().numpy() if host else y

    __matmul__ = dot

    def solve(self, rhs=None, rtol=DEFAULT_RTOL, maxit=DEFAULT_MAXIT, guess=None):
        """x = A^-1 rhs.  `guess`: optional (nz,nxx) CUDA tensor to start from (default: f_T)."""
        host = self.host if rhs is None else not isinstance(rhs, torch.Tensor)
        if guess is not None:
            self._guess = _dev(guess, self.ctx)
            self.ctx.check(self.ctx.lib.plb_diff_set_initial_guess(self.h, self._guess.data_ptr()))
        rd = None if rhs is None else _dev(rhs, self.ctx).reshape(-1)
        # (slab-local fields: the solver writes the own rows and the halo rows only; the rest stays zero)
        new = torch.zeros if self.ctx.slab is not None else torch.empty
        x = new(self.shape[0], dtype=torch.float64, device=self.ctx.torch_device)
        it, rr = C.c_int(0), C.c_double(0)
        rc = self.ctx.lib.plb_diff_solve(self.h, None if rd is None else rd.data_ptr(), float(rtol),
                                         int(maxit), x.data_ptr(), C.byref(it), C.byref(rr))
        self.iterations, self.relres = it.value, rr.value
        self.ctx.check(rc)
        if self.ctx.comm_info()[1] > 1 and self.ctx.slab is None:
            self.ctx.allreduce(x)       # every slab rank filled its own rows
        return x.cpu().numpy() if host else x

    def close(self):
        if getattr(self, "h", None):
            self.ctx.lib.plb_diff_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def makeDiffusionMatrix(nx, grid, gridmp, f_T, f_k, f_Cp, f_rho, f_H, bc, bcvalue, tstep):
    """Implicit-Euler heat-conduction system; reference: pylamp_diff.py:85-183.
    Returns ``(A, rhs)`` with ``A`` a :class:`DiffusionOperator`."""
    A = DiffusionOperator(nx, grid, gridmp, f_T, f_k, f_Cp, f_rho, f_H, bc, bcvalue, tstep)
    return A, A.rhs(device=not A.host)
