"""Run the UNMODIFIED reference driver (`pylamp2.py`) on top of the drop-in modules.

    python -m pylamp_b200.launcher /path/to/reference/pylamp2.py [nsteps]

pylamp2.py is a script without functions: it imports `pylamp_trac`, `pylamp_stokes`, `pylamp_diff`
(`from ... import *`, pylamp2.py:12-17) and calls `scipy.sparse.linalg.spsolve(scipy.sparse.csc_matrix(A),
rhs)` through module attributes resolved at call time (pylamp2.py:10, :360, :394, :419).  The launcher
 1. registers this package's drop-ins under the reference's module names, so the script's imports bind
    to them (the reference's own `pylamp_const` / `pylamp_tool` are used as they are),
 2. rebinds `scipy.sparse.csc_matrix` and `scipy.sparse.linalg.spsolve` to wrappers that pass the
    operator handles of `makeStokesMatrix` / `makeDiffusionMatrix` to the GPU solvers and defer to
    SciPy for everything else,
 3. supplies what the reference needs from its environment on a current Python (a single-rank
    `mpi4py.MPI` if mpi4py is absent, `time.clock`, the `out/` directory),
 4. executes the script's source as `__main__`.  `substitutions` are (old, new) source-text edits -- the
    reference is configured by editing its source (README:32-34) -- and `nsteps` stops the run after
    that many `np.savez` pairs (the script's own limit is max_it = 1e10, pylamp2.py:67).
The driver's inline NumPy steps stay on the host, so every drop-in call copies its arrays in and out:
the parity path, not the throughput path (pylamp_b200.driver keeps the state in HBM).
"""
import contextlib
import io
import os
import sys
import time
import types

import numpy as np


def _ensure_mpi4py():
    try:
        import mpi4py.MPI  # noqa: F401
        return
    except Exception:
        pass
    pkg, mpi = types.ModuleType("mpi4py"), types.ModuleType("mpi4py.MPI")

    class _Comm:          # rank 0 of 1: what the reference's six MPI call sites see (pylamp2.py:30-32, 122, 446, 454, 554-555)
        def Get_rank(self):
            return 0

        def Get_size(self):
            return 1

        def Bcast(self, buf, root=0):
            return None

        def Allreduce(self, send, recv, op=None):
            recv[0][...] = send[0]

    mpi.COMM_WORLD, mpi.DOUBLE, mpi.SUM = _Comm(), "DOUBLE", "SUM"
    pkg.MPI = mpi
    sys.modules["mpi4py"], sys.modules["mpi4py.MPI"] = pkg, mpi


class _Stop(Exception):
    pass


def run(script, nsteps=None, substitutions=(), seed=None, workdir=None, quiet=False):
    """Execute the reference driver at `script` against the drop-in modules.  Returns the list of
    (griddata dict, tracs dict) the script handed to `np.savez`, one pair per time step."""
    import scipy.sparse
    import scipy.sparse.linalg
    from . import pylamp_diff, pylamp_stokes, pylamp_trac
    refdir = os.path.dirname(os.path.abspath(script))
    src = open(script).read()
    for old, new in substitutions:
        if src.count(old) != 1:
            raise ValueError("substitution %r occurs %d times in %s" % (old, src.count(old), script))
        src = src.replace(old, new)
    _ensure_mpi4py()
    if not hasattr(time, "clock"):
        time.clock = time.process_time                  # pylamp_tool.py:12, 16 (removed in Python 3.8)
    saved_modules = {n: sys.modules.get(n) for n in ("pylamp_trac", "pylamp_stokes", "pylamp_diff")}
    sys.modules["pylamp_trac"], sys.modules["pylamp_stokes"], sys.modules["pylamp_diff"] = \
        pylamp_trac, pylamp_stokes, pylamp_diff
    real_csc, real_spsolve, real_savez = scipy.sparse.csc_matrix, scipy.sparse.linalg.spsolve, np.savez
    handle = lambda A: isinstance(A, (pylamp_stokes.StokesOperator, pylamp_diff.DiffusionOperator))
    scipy.sparse.csc_matrix = lambda A, *a, **k: A if handle(A) else real_csc(A, *a, **k)
    scipy.sparse.linalg.spsolve = lambda A, b, *a, **k: A.solve(b) if handle(A) else real_spsolve(A, b, *a, **k)
    captured = []

    def savez(fname, **kw):
        kw = {k: np.array(v, copy=True) for k, v in kw.items()}
        if "griddata" in str(fname):
            captured.append([kw, None])
        else:
            captured[-1][1] = kw
            if nsteps is not None and len(captured) >= nsteps:
                raise _Stop()
        if nsteps is None:
            real_savez(fname, **kw)

    cwd = os.getcwd()
    if workdir is None:
        import tempfile
        workdir = tempfile.mkdtemp(prefix="pylamp_b200_run_")
    os.makedirs(os.path.join(workdir, "out"), exist_ok=True)      # pylamp2.py:58, :644
    if seed is not None:
        np.random.seed(seed)                                      # pylamp2.py:119 draws from the global stream
    sys.path.insert(0, refdir)                                    # pylamp_const, pylamp_tool of the reference
    glb = {"__name__": "__main__", "__file__": os.path.abspath(script)}
    try:
        os.chdir(workdir)
        np.savez = savez
        with (contextlib.redirect_stdout(io.StringIO()) if quiet else contextlib.nullcontext()):
            try:
                exec(compile(src, os.path.abspath(script), "exec"), glb)
            except _Stop:
                pass
    finally:
        np.savez = real_savez
        scipy.sparse.csc_matrix, scipy.sparse.linalg.spsolve = real_csc, real_spsolve
        os.chdir(cwd)
        sys.path.remove(refdir)
        for n, m in saved_modules.items():
            if m is None:
                sys.modules.pop(n, None)
            else:
                sys.modules[n] = m
    return captured


if __name__ == "__main__":
    run(sys.argv[1], nsteps=int(sys.argv[2]) if len(sys.argv) > 2 else None)
