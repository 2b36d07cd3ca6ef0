// fgmres.cuh -- restarted flexible GMRES(m) on device vectors (right preconditioning).
//
// Minimises || r ||_2 of r = b^ - A^ x over the span of the preconditioned directions; the
// caller's operator already contains any row scaling, so the norm is the caller's weighted norm.
// The preconditioner may change between iterations (multigrid cycles).  The Arnoldi basis V lives
// in the residual space, the directions Z in the solution space; the small Hessenberg
// least-squares problem is updated on the host with Givens rotations (two scalar read-backs per
// iteration: the Gram-Schmidt coefficients, then the norm of the new basis vector).
// Orthogonalisation: classical Gram-Schmidt with one fused multi-dot, and a second pass whenever
// the first one cancelled the vector by more than two digits (selective re-orthogonalisation).
// Restarting rather than a sliding window matters here: the preconditioned saddle-point operator
// has many outlying eigenvalues and truncated recurrences stagnate (oracle/mg_prototype.py).
// Basis vectors are allocated on first use.
#pragma once
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "blas1.cuh"

struct plb_fgmres_ws {
    int m = 0;                 // restart length
    long long n = 0;
    std::vector<double*> V, Z; // m+1 basis vectors, m directions (device, lazily allocated)
    double* d_scal = nullptr;  // >= m + 8 device scalars
    double* h_coef = nullptr;  // pinned host staging, >= m + 8 doubles
    double reorth_thresh = 1e-4;   // second Gram-Schmidt pass if |w_after|^2 < thresh * |w_before|^2
    double pyth_thresh = 1e-2;     // norm by Pythagoras only if |w_after|^2 > thresh * |w_before|^2
};

struct plb_fgmres_result {
    int iters = 0;
    double relres = 0;         // Arnoldi estimate of || r || / bnorm at exit
    bool converged = false;
    double floor = 0;          // > 0: the true residual stopped following the Arnoldi estimate here
    bool stagnated = false;    // restart cycles stopped reducing the true residual (NOT a floor: no status change)
};

inline int plb_fgmres_alloc(plb_ctx* ctx, plb_fgmres_ws* ws, int m, long long n, double* d_scal) {
    for (double* p : ws->V) if (p) cudaFree(p);
    for (double* p : ws->Z) if (p) cudaFree(p);
    if (ws->h_coef) cudaFreeHost(ws->h_coef);
    ws->V.assign(m + 1, nullptr), ws->Z.assign(m, nullptr);
    ws->m = m, ws->n = n, ws->d_scal = d_scal;
    PLB_CUDA(ctx, cudaMallocHost((void**)&ws->h_coef, sizeof(double) * (m + 8)));
    return 0;
}

inline void plb_fgmres_free(plb_fgmres_ws* ws) {
    for (double* p : ws->V) if (p) cudaFree(p);
    for (double* p : ws->Z) if (p) cudaFree(p);
    if (ws->h_coef) cudaFreeHost(ws->h_coef);
    ws->V.clear(), ws->Z.clear(), ws->h_coef = nullptr;
}

inline int plb_fgmres_read(plb_ctx* ctx, plb_fgmres_ws* ws, const double* d, int k) {
    PLB_CUDA(ctx, cudaMemcpyAsync(ws->h_coef, d, sizeof(double) * k, cudaMemcpyDeviceToHost, ctx->stream));
    PLB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

// residual(r): r = b^ - A^ x (true residual of the current x).  apply(z, w): w = A^ z.
// precond(v, z): z = M v.   x is updated in place; r is scratch (n doubles).
template <class Residual, class Apply, class Precond>
int plb_fgmres(plb_ctx* ctx, plb_reduce_ws* rws, plb_fgmres_ws* ws, Residual residual, Apply apply,
               Precond precond, double* x, double bnorm, double rtol, int maxit,
               plb_fgmres_result* res) {
    const long long n = ws->n;
    const int m = ws->m;
    double* S = ws->d_scal;
    std::vector<double> H((size_t)(m + 1) * m, 0.0), cs(m), sn(m), g(m + 1), y(m), hcol(m + 2);
    std::vector<const double*> Vp;
    auto need = [&](std::vector<double*>& A, int i) -> int {
        if (!A[i]) PLB_CUDA(ctx, cudaMalloc(&A[i], sizeof(double) * n));
        return 0;
    };
    res->iters = 0, res->converged = false, res->relres = 1;
    int total = 0, stalls = 0;
    double beta_prev = INFINITY, est_prev = -1;
    res->floor = 0, res->stagnated = false;
    while (total < maxit) {
        if (need(ws->V, 0)) return 2;
        if (residual(ws->V[0])) return 2;
        if (plb_dot(ctx, rws, n, ws->V[0], ws->V[0], S)) return 2;
        if (plb_fgmres_read(ctx, ws, S, 1)) return 2;
        double beta = sqrt(ws->h_coef[0]);
        if (!(beta == beta)) PLB_FAIL(ctx, "FGMRES: residual is NaN");
        res->relres = beta / bnorm;
        if (getenv("PLB_DEBUG_FGMRES"))
            fprintf(stderr, "  fgmres restart at it %d: TRUE relres %.6e (Arnoldi estimate at the end of the cycle before: %.6e)\n",
                    total, res->relres, est_prev >= 0 ? est_prev / bnorm : -1.0);
        // (after at least one cycle a true residual within 1.5x of the target is accepted: another
        // restart cycle would spend tens of iterations on a few percent)
        if (beta <= rtol * bnorm || (total > 0 && beta <= 1.5 * rtol * bnorm)) {
            res->converged = true;
            break;
        }
        // fp64 floor of the residual evaluation: in exact arithmetic the TRUE residual at a restart equals
        // the Arnoldi estimate at the end of the cycle before.  A true residual several times above that
        // estimate is rounding -- either lost orthogonality in a long cycle (the restart repairs that: the
        // next cycle makes progress again) or the accuracy of b^ - A^ x itself (no cycle can go below it).
        // Only the second is a floor, so both are required: the gap AND a cycle that failed to halve the
        // true residual.  Slow convergence alone (no gap) is never reported as a floor.
        if (est_prev >= 0 && beta > 5 * est_prev && beta > 0.5 * beta_prev) {
            res->floor = beta / bnorm;
            break;
        }
        // stagnation without a gap: three cycles in a row that did not halve the residual -- give up
        // (not converged; the caller sees relres and decides)
        stalls = (beta > 0.5 * beta_prev) ? stalls + 1 : 0;
        if (stalls >= 3) {
            res->stagnated = true;
            break;
        }
        beta_prev = beta;
        if (plb_scale_rsqrt2(ctx, n, S, ws->V[0], nullptr)) return 2;
        g.assign(m + 1, 0.0);
        g[0] = beta;
        int k = 0;
        bool inner_conv = false;
        for (; k < m && total < maxit; k++) {
            if (need(ws->Z, k) || need(ws->V, k + 1)) return 2;
            double* w = ws->V[k + 1];
            if (precond(ws->V[k], ws->Z[k])) return 2;
            if (apply(ws->Z[k], w)) return 2;
            Vp.clear();
            for (int j = 0; j <= k; j++) Vp.push_back(ws->V[j]);
            Vp.push_back(w);                                   // last entry: w.w
            for (int j = 0; j <= k + 1; j++) hcol[j] = 0;
            // Pass 1: h = V^T w and w.w in one fused multi-dot.  By Pythagoras the norm after the
            // projection is w.w - sum h^2; when that keeps at least two digits (the normal case) it is
            // used directly and the normalisation is folded into the projection kernel (no separate
            // dot and scale passes).  After heavy cancellation: explicit norm and a second pass.
            if (plb_multi_dot(ctx, rws, n, k + 2, Vp.data(), w, S)) return 2;
            if (plb_fgmres_read(ctx, ws, S, k + 2)) return 2;
            double w2_before = ws->h_coef[k + 1], sumh2 = 0;
            for (int j = 0; j <= k; j++) hcol[j] = ws->h_coef[j], sumh2 += hcol[j] * hcol[j];
            double w2_after = w2_before - sumh2;
            double hn;
            if (w2_after > ws->pyth_thresh * w2_before) {
                hn = sqrt(w2_after);
                if (plb_multi_axpy2(ctx, n, k + 1, S, Vp.data(), w, nullptr, nullptr, 1.0 / hn)) return 2;
            } else {
                if (plb_multi_axpy2(ctx, n, k + 1, S, Vp.data(), w, nullptr, nullptr)) return 2;
                if (plb_dot(ctx, rws, n, w, w, S + k + 2)) return 2;
                if (plb_fgmres_read(ctx, ws, S + k + 2, 1)) return 2;
                w2_after = ws->h_coef[0];
                if (w2_after < ws->reorth_thresh * w2_before) {       // second Gram-Schmidt pass
                    if (plb_multi_dot(ctx, rws, n, k + 1, Vp.data(), w, S)) return 2;
                    if (plb_multi_axpy2(ctx, n, k + 1, S, Vp.data(), w, nullptr, nullptr)) return 2;
                    if (plb_dot(ctx, rws, n, w, w, S + k + 2)) return 2;
                    if (plb_fgmres_read(ctx, ws, S, k + 3)) return 2;
                    for (int j = 0; j <= k; j++) hcol[j] += ws->h_coef[j];
                    w2_after = ws->h_coef[k + 2];
                }
                hn = sqrt(w2_after);
                if (hn > 0 && plb_scale_rsqrt2(ctx, n, S + k + 2, w, nullptr)) return 2;
            }
            if (!(hn == hn)) PLB_FAIL(ctx, "FGMRES: NaN in Arnoldi step %d", total + 1);
            hcol[k + 1] = hn;
            // Givens update of column k
            for (int j = 0; j < k; j++) {
                double t = cs[j] * hcol[j] + sn[j] * hcol[j + 1];
                hcol[j + 1] = -sn[j] * hcol[j] + cs[j] * hcol[j + 1];
                hcol[j] = t;
            }
            double den = hypot(hcol[k], hcol[k + 1]);
            cs[k] = den > 0 ? hcol[k] / den : 1.0;
            sn[k] = den > 0 ? hcol[k + 1] / den : 0.0;
            hcol[k] = den;
            g[k + 1] = -sn[k] * g[k];
            g[k] = cs[k] * g[k];
            for (int j = 0; j <= k; j++) H[(size_t)j * m + k] = hcol[j];
            total++;
            res->iters = total;
            res->relres = fabs(g[k + 1]) / bnorm;
            if (getenv("PLB_DEBUG_FGMRES")) fprintf(stderr, "  fgmres it %d est %.6e hn %.6e h0 %.6e\n", total, res->relres, hn, hcol[0]);
            if (fabs(g[k + 1]) <= rtol * bnorm || hn == 0) {
                k++;
                inner_conv = true;
                break;
            }
        }
        // y = H^-1 g (upper triangular), x += Z y
        for (int i = k - 1; i >= 0; i--) {
            double s = g[i];
            for (int j = i + 1; j < k; j++) s -= H[(size_t)i * m + j] * y[j];
            y[i] = s / H[(size_t)i * m + i];
        }
        for (int j = 0; j < k; j++) ws->h_coef[j] = -y[j];     // multi_axpy2 subtracts
        PLB_CUDA(ctx, cudaMemcpyAsync(S, ws->h_coef, sizeof(double) * k, cudaMemcpyHostToDevice, ctx->stream));
        std::vector<const double*> Zp(ws->Z.begin(), ws->Z.begin() + k);
        if (k > 0 && plb_multi_axpy2(ctx, n, k, S, Zp.data(), x, nullptr, nullptr)) return 2;
        PLB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));      // h_coef is reused by the next read
        (void)inner_conv;    // the loop head re-evaluates the TRUE residual and decides
        est_prev = k > 0 ? fabs(g[k]) : beta;
    }
    if (!res->converged) {
        // final true residual for the report
        if (need(ws->V, 0)) return 2;
        if (residual(ws->V[0])) return 2;
        if (plb_dot(ctx, rws, n, ws->V[0], ws->V[0], S)) return 2;
        if (plb_fgmres_read(ctx, ws, S, 1)) return 2;
        res->relres = sqrt(ws->h_coef[0]) / bnorm;
        res->converged = res->relres <= rtol;
    }
    return 0;
}
