"""Replacement for the driver's ``scipy.sparse.linalg.spsolve(scipy.sparse.csc_matrix(A), rhs)``
(pylamp2.py:360, :394, :419): routes the operator handles returned by
``pylamp_stokes.makeStokesMatrix`` / ``pylamp_diff.makeDiffusionMatrix`` to the GPU solvers.

There is no CPU path: anything that is not one of those handles raises TypeError."""


def csc_matrix(A):
    """Pass-through used in place of scipy.sparse.csc_matrix for operator handles."""
    return A


def spsolve(A, rhs, **kw):
    solve = getattr(A, "solve", None)
    if solve is None or not getattr(A, "_plb_handle", True):
        raise TypeError("pylamp_b200.solve.spsolve expects a StokesOperator / DiffusionOperator "
                        "(got %r); there is no CPU fallback" % type(A))
    return solve(rhs, **kw)
