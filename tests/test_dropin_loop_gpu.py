"""Drop-in at the function boundary, with HOST (NumPy) arrays like the reference driver holds them:
the loop body of pylamp2.py:273-594 (the oracle's restatement of it, which makes exactly the
reference's call sequence: fancy-indexed `tr_f[:, [cols]]` copies, lists of (nz,nxx) arrays mutated
in place, `spsolve(A, rhs)` on whatever `make*Matrix` returned) is run twice -- once on the
oracle's own functions, once with the seven module-level functions + spsolve swapped for
pylamp_b200's -- and the results are compared.  Tolerances as in test_driver_gpu.py."""
import contextlib

import numpy as np
import pytest

from oracle import pylamp_oracle as O
from pylamp_b200 import setups

pytestmark = pytest.mark.gpu


@contextlib.contextmanager
def dropin():
    from pylamp_b200 import pylamp_trac as T, pylamp_stokes as S, pylamp_diff as D, solve
    names = {"trac2grid": T.trac2grid, "grid2trac": T.grid2trac, "RK": T.RK,
             "makeStokesMatrix": S.makeStokesMatrix, "x2vp": S.x2vp,
             "makeDiffusionMatrix": D.makeDiffusionMatrix, "x2t": D.x2t}
    saved = {k: getattr(O, k) for k in names}
    try:
        for k, v in names.items():
            setattr(O, k, v)
        yield solve.spsolve
    finally:
        for k, v in saved.items():
            setattr(O, k, v)


@pytest.mark.parametrize("name", ["thermo", "c1"])
def test_reference_loop_on_dropin_modules(name):
    setup = setups.thermo_variant(4321) if name == "thermo" else setups.c1_shipped(1234)
    nx, L, tr_x, tr_f, opts = setup
    ref = O.State(nx, L, tr_x.copy(), tr_f.copy())
    oref = O.Options(solve=O.solve_refined, **opts)
    new = O.State(nx, L, tr_x.copy(), tr_f.copy())
    rel = lambda a, b: np.linalg.norm(np.asarray(a) - b) / np.linalg.norm(b)
    for it in range(2):
        O.timestep(ref, oref)
        with dropin() as gpu_spsolve:
            O.timestep(new, O.Options(solve=gpu_spsolve, **opts))
        A, rhs = O.makeStokesMatrix(nx, ref.grid, ref.f_etas, ref.f_etan, ref.f_rho, oref.bcstokes)
        (fz, fx), fp = O.x2vp(O.spsolve(A, rhs), nx)
        floor = [rel(fz, ref.newvel[0]), rel(fx, ref.newvel[1]), rel(fp, ref.newpres)]
        e = [rel(new.newvel[0], ref.newvel[0]), rel(new.newvel[1], ref.newvel[1]), rel(new.newpres, ref.newpres)]
        print(name, "step", it + 1, "err", ["%.1e" % v for v in e], "floor", ["%.1e" % v for v in floor],
              "x %.1e" % rel(new.tr_x, ref.tr_x))
        for a, f in zip(e, floor):
            assert a <= max(1e-8, 3 * f)
        assert rel(new.f_rho, ref.f_rho) < 1e-9
        assert rel(new.tr_x, ref.tr_x) < (1e-10 if name == "thermo" else 1e-8)
        if opts["do_heatdiff"]:
            assert rel(new.newtemp, ref.newtemp) < 1e-8
            assert rel(new.tr_f[:, O.TR_TMP], ref.tr_f[:, O.TR_TMP]) < 1e-10
        assert np.array_equal(new.count, ref.count) or np.mean(new.kelem == ref.kelem) > 0.9999
