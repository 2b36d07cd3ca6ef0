// diff.cu -- the reference's implicit-Euler heat-conduction system (pylamp_diff.makeDiffusionMatrix,
// pylamp_diff.py:85-183) as a matrix-free sm_100a operator, and the solver that replaces
// scipy.sparse.linalg.spsolve at pylamp2.py:419.
//
// Row r = i*nxx + j.  Interior rows (pylamp_diff.py:157-179), walls FIXTEMP / FIXFLOW (:99-152; the
// z-walls own the corners).  Because the driver's dt never exceeds 0.67*dx^2/max(2*kappa)
// (pylamp2.py:339-343), the row-scaled operator D^-1 A is I + (a perturbation of norm < 1): plain
// restarted GMRES on the row-scaled system converges in a few tens of iterations independent of
// the grid size, so no multigrid is needed here.
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include <vector>

#include "blas1.cuh"
#include "comm.cuh"
#include "fgmres.cuh"

namespace {

constexpr int BX = 32, BY = 8;
inline dim3 grid2d(int nz, int nxx) { return dim3((nxx + BX - 1) / BX, (nz + BY - 1) / BY); }
inline dim3 block2d() { return dim3(BX, BY); }

struct DiffDev {
    int nz, nxx, ld;
    int i0, i1;                   // rows handled by this rank (z-slab); single GPU: [0, nz)
    const double *idz, *idx;      // 1/(g[k+1]-g[k])
    const double *idzm, *idxm;    // 1/(gmp[k]-gmp[k-1]), k >= 1
    const double *T, *kz, *kx, *cp, *rho, *H;
    double dt;
    int bc[4];                    // [z=0, x=0, z=L, x=L]
    double bcval[4];
};

// coefficients of row (i,j): value = cC*T[i,j] + cE*T[i,j+1] + cW*T[i,j-1] + cS*T[i+1,j] + cN*T[i-1,j]
struct Row {
    double cC, cE, cW, cS, cN, rhs;
};

__device__ __forceinline__ Row diff_row(const DiffDev& D, int i, int j) {
    Row r = {0, 0, 0, 0, 0, 0};
    const int nz = D.nz, nxx = D.nxx;
    const long long o = (long long)i * D.ld + j;
    if (i == 0 || i == nz - 1) {                                  // z-walls, :99-124
        const int w = (i == 0) ? 0 : 2;
        r.rhs = D.bcval[w];
        if (D.bc[w] == PLB_BC_FIXTEMP) {
            r.cC = 1.0;
        } else if (i == 0) {
            double c = D.kz[o] * D.idz[0];
            r.cS = c, r.cC = -c;
        } else {
            double c = D.kz[o - D.ld] * D.idz[nz - 2];
            r.cC = c, r.cN = -c;
        }
    } else if (j == 0 || j == nxx - 1) {                          // x-walls, :126-152
        const int w = (j == 0) ? 1 : 3;
        r.rhs = D.bcval[w];
        if (D.bc[w] == PLB_BC_FIXTEMP) {
            r.cC = 1.0;
        } else if (j == 0) {
            double c = D.kx[o] * D.idx[0];
            r.cE = c, r.cC = -c;
        } else {
            double c = D.kx[o - 1] * D.idx[nxx - 2];
            r.cC = c, r.cW = -c;
        }
    } else {                                                       // interior, :157-179
        const double rc = D.rho[o] * D.cp[o];
        const double pre = D.dt / rc;
        r.cE = pre * D.kx[o] * D.idx[j] * D.idxm[j];
        r.cW = pre * D.kx[o - 1] * D.idx[j - 1] * D.idxm[j];
        r.cS = pre * D.kz[o] * D.idz[i] * D.idzm[i];
        r.cN = pre * D.kz[o - D.ld] * D.idz[i - 1] * D.idzm[i];
        r.cC = -(r.cE + r.cW + r.cS + r.cN) - 1.0;
        r.rhs = -D.T[o] - D.dt * D.H[o] / rc;
    }
    return r;
}

__device__ __forceinline__ double row_apply(const DiffDev& D, const Row& r, const double* __restrict__ x,
                                            int i, int j) {
    const long long o = (long long)i * D.ld + j;
    double v = r.cC * x[o];
    if (r.cE != 0) v += r.cE * x[o + 1];
    if (r.cW != 0) v += r.cW * x[o - 1];
    if (r.cS != 0) v += r.cS * x[o + D.ld];
    if (r.cN != 0) v += r.cN * x[o - D.ld];
    return v;
}

// MODE 0: y = A x (reference rows, unscaled)   MODE 1: y = rhs   MODE 2: y = D^-1 (b - A x)
// MODE 3: y = D^-1 A x        (b == nullptr in MODE 2 means the system's own rhs)
template <int MODE>
__global__ void __launch_bounds__(BX* BY)
k_diff(DiffDev D, const double* __restrict__ x, const double* __restrict__ b, double* __restrict__ y) {
    const int j = blockIdx.x * BX + threadIdx.x, i = D.i0 + blockIdx.y * BY + threadIdx.y;
    if (i >= D.i1 || j >= D.nxx) return;
    const long long o = (long long)i * D.ld + j;
    Row r = diff_row(D, i, j);
    if (MODE == 1) {
        y[o] = r.rhs;
        return;
    }
    double a = row_apply(D, r, x, i, j);
    if (MODE == 0) y[o] = a;
    else if (MODE == 2) y[o] = ((b ? b[o] : r.rhs) - a) / r.cC;
    else y[o] = a / r.cC;
}


// ---------------------------------------------------------------------------------------------
// Chebyshev-Jacobi iteration for the heat system (the default path of plb_diff_solve).
//
// Every wall row is an explicit relation between a wall node and its normal neighbour (FIXTEMP: T_w = value;
// FIXFLOW: k (T_1 - T_0)/d = value; the z-walls own the corners, pylamp_diff.py:99-152), so the wall nodes are
// eliminated: they are kept consistent with their interior neighbours ("slaves"), and the iteration runs on the
// interior rows, whose eliminated diagonal is diag' = cC + sum over FIXFLOW wall neighbours of their coefficient.
// Row-scaled by diag' the eliminated operator is I + N with ||N||_inf = rho = max_i S_i / (S_i + 1) < 1, S_i = the
// row's sum of interior off-diagonal coefficients (dt <= 0.67 dx^2 / max(2 kappa) in the time loop gives
// rho ~ 0.57): the spectrum lies in [1 - rho, 1 + rho], which is all a Chebyshev iteration needs -- no inner
// products, no host read-back per iteration, one halo exchange per sweep on several GPUs.
// ---------------------------------------------------------------------------------------------
struct WallRel {       // T_wall = a + b * T_neighbour
    double a, b;
};

// relation of wall node (iw, jw) to its normal neighbour (from the wall row)
__device__ __forceinline__ WallRel wall_rel(const DiffDev& D, int iw, int jw) {
    const Row r = diff_row(D, iw, jw);
    const double off = r.cE + r.cW + r.cS + r.cN;      // a wall row has at most one off-diagonal entry
    WallRel w;
    w.a = r.rhs / r.cC, w.b = -off / r.cC;
    return w;
}

__device__ __forceinline__ bool is_wall(const DiffDev& D, int i, int j) {
    return i == 0 || i == D.nz - 1 || j == 0 || j == D.nxx - 1;
}

// eliminated diagonal and Gershgorin radius of interior row (i,j)
__device__ __forceinline__ void elim_row(const DiffDev& D, const Row& r, int i, int j, double& diag, double& offsum) {
    diag = r.cC, offsum = 0;
    if (j == 1) diag += r.cW * wall_rel(D, i, 0).b; else offsum += fabs(r.cW);
    if (j == D.nxx - 2) diag += r.cE * wall_rel(D, i, D.nxx - 1).b; else offsum += fabs(r.cE);
    if (i == 1) diag += r.cN * wall_rel(D, 0, j).b; else offsum += fabs(r.cN);
    if (i == D.nz - 2) diag += r.cS * wall_rel(D, D.nz - 1, j).b; else offsum += fabs(r.cS);
}

// the wall nodes next to interior node (i,j), from its value v: x-wall nodes first, then the z-wall nodes incl. the
// corners (a corner's z-wall row refers to the x-wall node of the row next to it)
__device__ __forceinline__ void write_slaves(const DiffDev& D, double* __restrict__ x, int i, int j, double v) {
    const int nz = D.nz, nxx = D.nxx, ld = D.ld;
    const long long o = (long long)i * ld + j;
    double vl = 0, vr = 0;
    if (j == 1) { const WallRel w = wall_rel(D, i, 0); vl = w.a + w.b * v; x[o - 1] = vl; }
    if (j == nxx - 2) { const WallRel w = wall_rel(D, i, nxx - 1); vr = w.a + w.b * v; x[o + 1] = vr; }
    if (i == 1) {
        const WallRel w = wall_rel(D, 0, j);
        x[o - ld] = w.a + w.b * v;
        if (j == 1) { const WallRel c = wall_rel(D, 0, 0); x[o - ld - 1] = c.a + c.b * vl; }
        if (j == nxx - 2) { const WallRel c = wall_rel(D, 0, nxx - 1); x[o - ld + 1] = c.a + c.b * vr; }
    }
    if (i == nz - 2) {
        const WallRel w = wall_rel(D, nz - 1, j);
        x[o + ld] = w.a + w.b * v;
        if (j == 1) { const WallRel c = wall_rel(D, nz - 1, 0); x[o + ld - 1] = c.a + c.b * vl; }
        if (j == nxx - 2) { const WallRel c = wall_rel(D, nz - 1, nxx - 1); x[o + ld + 1] = c.a + c.b * vr; }
    }
}

// MODE 0: make the wall nodes of x consistent with its interior values.  MODE 1: one Chebyshev-Jacobi step
// d = cd d + cr diag'^-1 (b - A x), xout = x + d (FIRST: d = cr ...).  MODE 2: rho = max radius (atomic max into out).
template <int MODE>
__global__ void __launch_bounds__(BX* BY)
k_diff_cheb(DiffDev D, const double* x, const double* __restrict__ b, double* __restrict__ d,
            double* xout, double cd, double cr, int first, double* __restrict__ out) {      // (MODE 0: xout == x)
    const int j = blockIdx.x * BX + threadIdx.x, i = D.i0 + blockIdx.y * BY + threadIdx.y;
    double rad = 0;
    if (i < D.i1 && j < D.nxx && !is_wall(D, i, j)) {
        const long long o = (long long)i * D.ld + j;
        if (MODE == 0) {
            write_slaves(D, xout, i, j, x[o]);
        } else {
            const Row r = diff_row(D, i, j);
            double diag, offsum;
            elim_row(D, r, i, j, diag, offsum);
            if (MODE == 2) {
                rad = offsum / fabs(diag);
            } else {
                const double res = ((b ? b[o] : r.rhs) - row_apply(D, r, x, i, j)) / diag;
                const double dn = first ? cr * res : cd * d[o] + cr * res;
                d[o] = dn;
                const double v = x[o] + dn;
                xout[o] = v;
                write_slaves(D, xout, i, j, v);
            }
        }
    }
    if (MODE == 2) {
        rad = warp_max(rad);
        __shared__ double sm[BX * BY / 32];
        const int t = threadIdx.y * BX + threadIdx.x;
        if ((t & 31) == 0) sm[t >> 5] = rad;
        __syncthreads();
        if (t == 0) {
            for (int q = 1; q < BX * BY / 32; q++) rad = fmax(rad, sm[q]);
            atomic_max_double(out, rad);
        }
    }
}

}  // namespace

struct plb_diff {
    plb_ctx* ctx = nullptr;
    int device = 0;
    int nz = 0, nxx = 0, ld = 0;
    double *idz = nullptr, *idx = nullptr, *idzm = nullptr, *idxm = nullptr;
    DiffDev dev{};
    bool coeffs = false;
    plb_reduce_ws rws{};
    double* d_scal = nullptr;
    plb_fgmres_ws kry;
    double *xs = nullptr, *xl = nullptr;   // local work vectors (slab rows + halos)
    // z-slab ownership when the context has a communicator: rows [i0, i1), stored rows [lo, hi]
    int i0 = 0, i1 = 0, lo = 0, hi = 0;
    bool dist = false;
    bool slab_fields = false;              // fields and result are slab-local (plb_ctx_set_slab)
    size_t plane = 0;                      // local plane size
    long long shift = 0;
    int m = 40;
    int use_cheb = 1;                      // Chebyshev-Jacobi on the wall-eliminated system (fallback: GMRES)
    double *cd_ = nullptr, *cx2 = nullptr; // Chebyshev direction and second iterate (local planes)
    int last_sweeps = 0;
    double last_rho = 0;
    const double* guess = nullptr;         // full-size initial guess for the next solve (NULL: T)
    int last_iters = 0;
    double last_relres = 0;
};

namespace {
int upload(plb_ctx* ctx, const std::vector<double>& h, double** d) {
    PLB_CUDA(ctx, cudaMalloc(d, sizeof(double) * h.size()));
    PLB_CUDA(ctx, cudaMemcpyAsync(*d, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice, ctx->stream));
    PLB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}
}  // namespace

extern "C" {

int plb_diff_create(plb_ctx* ctx, int nz, int nxx, int ld, const double* h_grid_z,
                    const double* h_grid_x, const double* h_gridmp_z, const double* h_gridmp_x,
                    const int* h_bc, const double* h_bcvalue, plb_diff** out) {
    if (!ctx || !out) return 1;
    *out = nullptr;
    if (nz < 3 || nxx < 3) PLB_FAIL(ctx, "plb_diff_create: grid too small");
    if (ld != nxx) PLB_FAIL(ctx, "plb_diff_create: only ld == nxx is supported");
    for (int w = 0; w < 4; w++)
        if (h_bc[w] != PLB_BC_FIXTEMP && h_bc[w] != PLB_BC_FIXFLOW)
            PLB_FAIL(ctx, "plb_diff_create: unknown heat BC %d on wall %d", h_bc[w], w);
    PLB_CUDA(ctx, cudaSetDevice(ctx->device));
    plb_diff* op = new plb_diff();
    op->ctx = ctx, op->device = ctx->device, op->nz = nz, op->nxx = nxx, op->ld = ld;
    std::vector<double> idz(nz, 0.0), idx(nxx, 0.0), idzm(nz, 0.0), idxm(nxx, 0.0);
    for (int i = 0; i + 1 < nz; i++) idz[i] = 1.0 / (h_grid_z[i + 1] - h_grid_z[i]);
    for (int j = 0; j + 1 < nxx; j++) idx[j] = 1.0 / (h_grid_x[j + 1] - h_grid_x[j]);
    for (int i = 1; i < nz; i++) idzm[i] = 1.0 / (h_gridmp_z[i] - h_gridmp_z[i - 1]);
    for (int j = 1; j < nxx; j++) idxm[j] = 1.0 / (h_gridmp_x[j] - h_gridmp_x[j - 1]);
    if (upload(ctx, idz, &op->idz) || upload(ctx, idx, &op->idx) || upload(ctx, idzm, &op->idzm) ||
        upload(ctx, idxm, &op->idxm) || plb_reduce_ws_init(ctx, &op->rws)) {
        plb_diff_destroy(op);
        return 2;
    }
    PLB_CUDA(ctx, cudaMalloc(&op->d_scal, sizeof(double) * 1024));
    const int R = plb_comm_size(ctx), rank = plb_comm_rank(ctx);
    if (const char* e = getenv("PLB_DIFF_GMRES")) op->use_cheb = atoi(e) ? 0 : 1;     // A/B switch: GMRES only
    op->dist = R > 1 && nz >= 4 * R;
    op->slab_fields = op->dist && ctx->slab_on;
    if (op->slab_fields) {
        // slab-local fields: the rows the context declared (those of the Stokes slabs)
        if (ctx->slab_i1 > nz) PLB_FAIL(ctx, "plb_diff_create: slab rows [%d, %d) exceed %d node rows", ctx->slab_i0, ctx->slab_i1, nz);
        op->i0 = ctx->slab_i0, op->i1 = ctx->slab_i1;
        op->lo = rank > 0 ? op->i0 - 1 : 0, op->hi = rank < R - 1 ? op->i1 : nz - 1;
    } else if (op->dist) {
        op->i0 = (int)((long long)rank * nz / R), op->i1 = (int)((long long)(rank + 1) * nz / R);
        op->lo = rank > 0 ? op->i0 - 1 : 0, op->hi = rank < R - 1 ? op->i1 : nz - 1;
    } else {
        op->i0 = 0, op->i1 = nz, op->lo = 0, op->hi = nz - 1;
    }
    op->plane = (size_t)(op->hi - op->lo + 1) * ld;
    op->shift = (long long)op->lo * ld;
    PLB_CUDA(ctx, cudaMalloc(&op->xs, sizeof(double) * op->plane));
    PLB_CUDA(ctx, cudaMalloc(&op->xl, sizeof(double) * op->plane));
    DiffDev& D = op->dev;
    D.nz = nz, D.nxx = nxx, D.ld = ld, D.i0 = op->i0, D.i1 = op->i1;
    D.idz = op->idz, D.idx = op->idx, D.idzm = op->idzm, D.idxm = op->idxm;
    for (int w = 0; w < 4; w++) D.bc[w] = h_bc[w], D.bcval[w] = h_bcvalue[w];
    *out = op;
    return 0;
}

void plb_diff_destroy(plb_diff* op) {
    if (!op) return;
    cudaSetDevice(op->device);
    cudaDeviceSynchronize();
    double* ptrs[] = {op->idz, op->idx, op->idzm, op->idxm, op->d_scal, op->xs, op->xl, op->cd_, op->cx2};
    for (double* p : ptrs) if (p) cudaFree(p);
    plb_fgmres_free(&op->kry);
    plb_reduce_ws_free(&op->rws);
    delete op;
}

int plb_diff_set_coeffs(plb_diff* op, const double* d_T, const double* d_kz, const double* d_kx,
                        const double* d_cp, const double* d_rho, const double* d_H, double tstep) {
    if (!op) return 1;
    DiffDev& D = op->dev;
    D.T = d_T, D.kz = d_kz, D.kx = d_kx, D.cp = d_cp, D.rho = d_rho, D.H = d_H, D.dt = tstep;
    op->coeffs = true;
    return 0;
}

int plb_diff_set_initial_guess(plb_diff* op, const double* d_x0) {
    if (!op) return 1;
    op->guess = d_x0;
    return 0;
}

int plb_diff_rhs(plb_diff* op, double* d_rhs) {
    if (!op) return 1;
    plb_ctx* ctx = op->ctx;
    if (!op->coeffs) PLB_FAIL(ctx, "plb_diff_rhs: coefficients not set");
    PLB_CUDA(ctx, cudaSetDevice(ctx->device));
    DiffDev D = op->dev;
    D.i0 = 0, D.i1 = op->nz;                 // whole grid (fields are replicated on every rank)
    k_diff<1><<<grid2d(op->nz, op->nxx), block2d(), 0, ctx->stream>>>(D, nullptr, nullptr, d_rhs);
    PLB_LAUNCHED(ctx);
    return 0;
}

int plb_diff_apply(plb_diff* op, const double* d_x, double* d_y) {
    if (!op) return 1;
    plb_ctx* ctx = op->ctx;
    if (!op->coeffs) PLB_FAIL(ctx, "plb_diff_apply: coefficients not set");
    PLB_CUDA(ctx, cudaSetDevice(ctx->device));
    DiffDev D = op->dev;
    D.i0 = 0, D.i1 = op->nz;
    k_diff<0><<<grid2d(op->nz, op->nxx), block2d(), 0, ctx->stream>>>(D, d_x, nullptr, d_y);
    PLB_LAUNCHED(ctx);
    return 0;
}

int plb_diff_solve(plb_diff* op, const double* d_rhs, double rtol, int maxit, double* d_x,
                   int* h_iters, double* h_relres) {
    if (!op) return 1;
    plb_ctx* ctx = op->ctx;
    if (!op->coeffs) PLB_FAIL(ctx, "plb_diff_solve: coefficients not set");
    PLB_CUDA(ctx, cudaSetDevice(ctx->device));
    const long long n = (long long)op->plane;            // local vector length (slab rows + halos)
    if (op->kry.m != op->m && plb_fgmres_alloc(ctx, &op->kry, op->m, n, op->d_scal)) return 2;
    const DiffDev D = op->dev;
    const dim3 g = grid2d(op->i1 - op->i0, op->nxx), blk = block2d();
    const long long sh = op->shift;
    const size_t own_off = (size_t)(op->i0 - op->lo) * op->ld, own_len = (size_t)(op->i1 - op->i0) * op->ld;
    if (op->dist) plb_reduce_shape(&op->rws, 1, n, (long long)own_off, (long long)own_len, true);
    else plb_reduce_shape(&op->rws, 0, 0, 0, 0, false);
    auto halo = [&](double* v) -> int {
        if (!op->dist) return 0;
        return plb_comm_halo_exchange(ctx, v, 1, op->plane, op->ld, op->nxx, op->lo, op->i0, op->i1);
    };
    double* x = op->xl;
    // the caller's right-hand side is a full-size vector: global indexing, no shift
    auto resid_of = [&](double* xx, double* out) -> int {
        if (halo(xx)) return 2;
        k_diff<2><<<g, blk, 0, ctx->stream>>>(D, xx - sh, d_rhs, out - sh);
        PLB_LAUNCHED(ctx);
        return 0;
    };
    auto residual = [&](double* out) -> int { return resid_of(x, out); };
    // bnorm = || D^-1 b ||: residual of x = 0
    PLB_CUDA(ctx, cudaMemsetAsync(x, 0, sizeof(double) * n, ctx->stream));
    if (resid_of(x, op->xs)) return 2;
    double bn2;
    if (plb_dot(ctx, &op->rws, n, op->xs, op->xs, op->d_scal + 900)) return 2;
    if (plb_read_scalars(ctx, op->d_scal + 900, 1, &bn2)) return 2;
    const double bnorm = sqrt(bn2);
    // initial guess: the current temperature field (the rhs is -T_old - dt*H/(rho*cp)); local rows
    PLB_CUDA(ctx, cudaMemcpyAsync(x, (op->guess ? op->guess : D.T) + sh, sizeof(double) * n,
                                  cudaMemcpyDeviceToDevice, ctx->stream));
    op->guess = nullptr;
    plb_fgmres_result res;
    bool done = false;
    op->last_sweeps = 0;
    // wall rows with a dependent right-hand side (a caller-supplied b) keep the general path
    if (op->use_cheb && bnorm > 0 && !d_rhs) {
        if (!op->cd_) {
            PLB_CUDA(ctx, cudaMalloc(&op->cd_, sizeof(double) * op->plane));
            PLB_CUDA(ctx, cudaMalloc(&op->cx2, sizeof(double) * op->plane));
        }
        // the rows whose unknown this rank iterates on: its own rows; wall rows are written by their neighbours
        double* s2 = op->d_scal + 910;
        PLB_CUDA(ctx, cudaMemsetAsync(s2, 0, sizeof(double), ctx->stream));
        k_diff_cheb<2><<<g, blk, 0, ctx->stream>>>(D, nullptr, nullptr, nullptr, nullptr, 0, 0, 0, s2);
        PLB_LAUNCHED(ctx);
        if (plb_comm_allreduce(ctx, s2, 1, PLB_OP_MAX)) return 2;
        // consistent wall values for the start vector, then its residual
        k_diff_cheb<0><<<g, blk, 0, ctx->stream>>>(D, x - sh, nullptr, nullptr, x - sh, 0, 0, 0, nullptr);
        PLB_LAUNCHED(ctx);
        if (resid_of(x, op->xs)) return 2;
        if (plb_dot(ctx, &op->rws, n, op->xs, op->xs, op->d_scal + 911)) return 2;
        double h[2];
        if (plb_read_scalars(ctx, op->d_scal + 910, 2, h)) return 2;
        const double rho = h[0];
        double rel = sqrt(h[1]) / bnorm;
        op->last_rho = rho;
        if (rho < 0.97 && rel == rel) {
            const double lmin = 1 - rho, lmax = 1 + rho;
            const double theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin);
            const double sigma1 = theta / delta;                       // > 1
            const double conv = sigma1 - sqrt(sigma1 * sigma1 - 1);    // asymptotic factor per sweep
            double* cur = x;
            double* oth = op->cx2;
            for (int round = 0; round < 4 && rel > rtol && res.iters < maxit; round++) {
                // sweeps for the reduction still needed (Chebyshev bound 2 conv^n), a few to spare
                int nsw = rho > 1e-14 ? (int)ceil(log(2.0 * rel / (0.5 * rtol)) / -log(conv)) + 1 : 2;
                nsw = std::max(2, std::min(nsw, maxit - res.iters));
                double rk = 1.0 / sigma1;
                for (int k = 0; k < nsw; k++) {
                    double cd, cr;
                    if (k == 0) cd = 0, cr = 1.0 / theta;
                    else {
                        const double rn = 1.0 / (2 * sigma1 - rk);
                        cd = rn * rk, cr = 2 * rn / delta, rk = rn;
                    }
                    if (halo(cur)) return 2;
                    // (the other buffer's wall rows not adjacent to this rank's rows are never read)
                    k_diff_cheb<1><<<g, blk, 0, ctx->stream>>>(D, cur - sh, nullptr, op->cd_ - sh, oth - sh, cd, cr, k == 0, nullptr);
                    PLB_LAUNCHED(ctx);
                    std::swap(cur, oth);
                }
                res.iters += nsw;
                if (resid_of(cur, op->xs)) return 2;
                if (plb_dot(ctx, &op->rws, n, op->xs, op->xs, op->d_scal + 911)) return 2;
                if (plb_read_scalars(ctx, op->d_scal + 911, 1, h)) return 2;
                rel = sqrt(h[0]) / bnorm;
            }
            if (cur != x) PLB_CUDA(ctx, cudaMemcpyAsync(x, cur, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
            op->last_sweeps = res.iters;
            res.relres = rel;
            res.converged = rel <= rtol;
            done = res.converged;
        }
    }
    auto apply = [&](const double* z, double* w) -> int {
        if (halo(const_cast<double*>(z))) return 2;
        k_diff<3><<<g, blk, 0, ctx->stream>>>(D, z - sh, nullptr, w - sh);
        PLB_LAUNCHED(ctx);
        return 0;
    };
    auto precond = [&](const double* v, double* z) -> int { return plb_copy(ctx, n, v, z); };
    op->kry.pyth_thresh = 0.5;      // rtol is near the fp64 limit here: keep the Arnoldi norms exact
    if (done) {
        // converged by the Chebyshev iteration
    } else if (bnorm > 0) {
        const int spent = res.iters;
        if (plb_fgmres(ctx, &op->rws, &op->kry, residual, apply, precond, x, bnorm, rtol, maxit, &res)) return 2;
        res.iters += spent;
    } else {
        PLB_CUDA(ctx, cudaMemsetAsync(x, 0, sizeof(double) * n, ctx->stream));
        res.converged = true;
    }
    op->last_iters = res.iters, op->last_relres = res.relres;
    if (h_iters) *h_iters = res.iters;
    if (h_relres) *h_relres = res.relres;
    // full-size result: a slab rank fills its own rows, the rest stays zero (summed by the host side)
    // (slab-local fields: own rows only, then halo rows from the neighbours)
    if (op->dist && !op->slab_fields)
        PLB_CUDA(ctx, cudaMemsetAsync(d_x, 0, sizeof(double) * (size_t)op->nz * op->ld, ctx->stream));
    PLB_CUDA(ctx, cudaMemcpyAsync(d_x + (size_t)op->i0 * op->ld, x + own_off, sizeof(double) * own_len,
                                  cudaMemcpyDeviceToDevice, ctx->stream));
    if (op->slab_fields) {
        double* arrs[1] = {d_x};
        const long long rd[1] = {op->ld};
        if (plb_comm_halo_rows(ctx, 1, arrs, rd, op->i0, op->i1, ctx->slab_halo)) return 2;
    }
    if (!res.converged && res.relres > 1e3 * rtol)
        PLB_FAIL(ctx, "plb_diff_solve: not converged after %d iterations (relres %.3e > rtol %.3e)", res.iters,
                 res.relres, rtol);
    return 0;
}

}  // extern "C"
