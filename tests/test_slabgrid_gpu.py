"""plb_trac2grid_scatter + plb_trac2grid_finalise (the two halves `trac2grid_slab` is built from) on one
GPU: finalising all rows of the scattered planes must reproduce plb_trac2grid bit for bit (same
kernels, same planes), and a row range must fill exactly that range.

First run on a B200 in round 2 (profiles/r02_gpu_unverified.log).  The communication part (slabgrid.py) is
tested under gloo in tests/test_slabgrid_cpu.py; the 2-GPU run is
`scripts/multi_gpu_driver_check.py 64 3 slab reduce` (tests/test_multi_gpu.py)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import pylamp_oracle as O

pytestmark = [pytest.mark.gpu]


def test_scatter_finalise_equals_trac2grid():
    from pylamp_b200 import _lib, pylamp_trac as T
    ctx = _lib.default_context()
    rng = np.random.default_rng(2)
    nx, L = [41, 33], [1.0, 0.75]
    grid, mesh, gridmp, meshmp = O.make_grids(nx, L)
    M = 50001
    x = torch.as_tensor(rng.random((M, 2)) * L).cuda()
    cols = [torch.as_tensor(rng.uniform(1, 2, M)).cuda(), torch.as_tensor(10 ** rng.uniform(18, 24, M)).cuda()]
    for target, schemes in (([grid[0], grid[1]], [5, 6]), ([gridmp[0], grid[1]], [5, 6]), ([grid[0], grid[1]], [1, 2])):
        want = [torch.zeros(tuple(nx), dtype=torch.float64, device="cuda") for _ in cols]
        mm = T.trac2grid_device(ctx, x, cols, schemes, target, want)
        axz, lz, _ = T._extended_axis(np.asarray(target[0]), mm[0], mm[1])
        axx, lx, _ = T._extended_axis(np.asarray(target[1]), mm[2], mm[3])
        axz_d, axx_d = torch.as_tensor(axz).cuda(), torch.as_tensor(axx).cuda()
        nze, nxe = len(axz), len(axx)
        planes = torch.empty((len(cols) + 2, nze, nxe), dtype=torch.float64, device="cuda")
        npl = C.c_int(0)
        ctx.call("plb_trac2grid_scatter", M, x.data_ptr(), len(cols), _lib.ptr_array(cols), _lib.int_array(schemes),
                 axz_d.data_ptr(), nze, axx_d.data_ptr(), nxe, float(axz[0]), float(axz[-1] - axz[0]), float(axx[0]),
                 float(axx[-1] - axx[0]), planes.data_ptr(), C.byref(npl))
        assert npl.value == len(cols) + 1
        got = [torch.full(tuple(nx), -7.0, dtype=torch.float64, device="cuda") for _ in cols]
        for r0, r1 in ((0, 13), (13, nx[0])):
            ctx.call("plb_trac2grid_finalise", len(cols), _lib.int_array(schemes), planes.data_ptr(), nze, nxe, lz, lx,
                     nx[0], nx[1], nx[1], r0, r1, _lib.ptr_array(got))
            if r1 < nx[0]:
                assert bool((got[0][r1:] == -7.0).all())          # rows outside the range untouched
        for a, b in zip(got, want):
            # atomics: the two scatters may add in different orders
            assert torch.allclose(a, b, rtol=1e-12, atol=0, equal_nan=True)
