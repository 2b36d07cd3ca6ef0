// stokes.cu -- the reference's Stokes system (pylamp_stokes.makeStokesMatrix, pylamp_stokes.py:
// 104-563) as matrix-free sm_100a kernels, and the solver that replaces
// scipy.sparse.linalg.spsolve at pylamp2.py:360: flexible GCR on the BC-eliminated saddle-point
// system with a block-triangular preconditioner (viscosity-scaled pressure mass + one geometric
// multigrid V-cycle with Chebyshev-Jacobi smoothing on the velocity block).
#include <cuda_pipeline.h>
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "blas1.cuh"
#include "comm.cuh"
#include "fgmres.cuh"
#include "stencil.cuh"

namespace {

constexpr int BX = 32, BY = 8;
inline dim3 grid2d(int nz, int nxx) { return dim3((nxx + BX - 1) / BX, (nz + BY - 1) / BY); }
inline dim3 block2d() { return dim3(BX, BY); }

struct Level {
    int nz = 0, nxx = 0, ld = 0;
    // z-slab ownership (global row numbers): owned node rows [i0, i1), stored rows [lo, hi]
    // (one halo row towards each existing neighbour).  Replicated levels: i0 = lo = 0, i1 = nz.
    int i0 = 0, i1 = 0, lo = 0, hi = 0;
    bool dist = false;
    size_t plane = 0;      // doubles per LOCAL plane = (hi - lo + 1) * ld
    size_t full = 0;       // doubles per full plane = nz * ld (coefficient fields are replicated)
    long long shift = 0;   // lo * ld: kernels index with global rows through shifted pointers
    std::vector<double> gz, gx;
    double *idz = nullptr, *idzc = nullptr, *idx = nullptr, *idxc = nullptr;
    const double *etas = nullptr, *etan = nullptr;
    double *etas_own = nullptr, *etan_own = nullptr;
    int proper = 0;
    int vz_i0, vz_i1, vz_j0, vz_j1, vx_i0, vx_i1, vx_j0, vx_j1;
    double sl_z0 = 1, sl_z1 = 1;
    int ns_z0 = 0, ns_z1 = 0;
    int ft_x0 = 0;
    double *X = nullptr, *T = nullptr, *b = nullptr, *r = nullptr, *d = nullptr;   // 2 local planes each
    double lmax = 0;
    LevelDev dev() const {
        LevelDev L;
        L.nz = nz, L.nxx = nxx, L.ld = ld, L.i0 = i0, L.i1 = i1;
        L.idz = idz, L.idzc = idzc, L.idx = idx, L.idxc = idxc;
        L.etas = etas, L.etan = etan, L.proper = proper;
        L.vz_i0 = vz_i0, L.vz_i1 = vz_i1, L.vz_j0 = vz_j0, L.vz_j1 = vz_j1;
        L.vx_i0 = vx_i0, L.vx_i1 = vx_i1, L.vx_j0 = vx_j0, L.vx_j1 = vx_j1;
        L.sl_z0 = sl_z0, L.sl_z1 = sl_z1;
        L.ns_z0 = ns_z0, L.ns_z1 = ns_z1;
        L.ft_x0 = ft_x0;
        return L;
    }
    // pointer to a local array as the kernels see it (global row indexing)
    double* sh(double* p) const { return p - shift; }
    const double* sh(const double* p) const { return p - shift; }
    dim3 grid() const { return grid2d(i1 - i0, nxx); }
};

}  // namespace

struct plb_stokes {
    plb_ctx* ctx = nullptr;
    int device = 0;
    int nz = 0, nxx = 0, ld = 0;
    int bc[4] = {1, 1, 1, 1};
    std::vector<Level> lv;
    const double* rho = nullptr;
    double g_z = 9.81, g_x = 0;
    // free-surface stabilisation (pylamp_stokes.py:422-426, :483-487): theta*dt (0 = off) and the two
    // full-size planes of centred density gradients Dz | Dx the extra momentum-row terms are built from
    double surf = 0;
    double* surf_d = nullptr;
    double* wsc_d = nullptr;      // finest level: W_p = sqrt(eta_n)/Kc, sqrt(-diag K_vz), sqrt(-diag K_vx) (k_scale_planes)
    int ai = 3, aj = 2;           // pressure anchor cell (pylamp_stokes.py:525-551: (3,2), or (nz/2, 0) on a flow-through wall)
    double Kc = 0, Kb = 0;
    bool coeffs = false, hierarchy = false;
    bool slab_fields = false;     // coefficient fields and results are slab-local (plb_ctx_set_slab): own rows + halo rows
    plb_reduce_ws rws{};
    double* d_scal = nullptr;     // 128 device scalars
    // dense coarse solve
    int nc = 0;
    double* cinv = nullptr;       // nc x nc inverse (row-major)
    double *probes = nullptr, *gjM = nullptr;   // resident scratch of the dense inverse (n probes, [A | I])
    int scratch_n = 0;
    // Krylov storage (3 planes per vector)
    plb_fgmres_ws kry;
    double *xs = nullptr, *r3 = nullptr, *b3 = nullptr, *t3 = nullptr, *gz_d = nullptr, *gx_d = nullptr;
    // parameters
    int hydrostatic = 1, warm_start = 0, debug_halo = 0, tile_smoother = 1;
    int lmax_every = 1, lmax_age = -1;   // eigenvalue estimates: recompute every n-th set_coeffs
    // CUDA graph of the V-cycle below level `graph_level` (small, launch-latency-bound, replicated levels)
    int graph_level = -1, graph_launches = 0, use_graph = 1;
    int graph_all = -1;           // capture the WHOLE V-cycle (-1: on a single GPU only; 1 also captures NCCL calls)
    double* zv = nullptr;         // fixed output buffer of the whole-cycle graph (2 local planes)
    cudaGraphExec_t graph_exec = nullptr;
    // everything the captured cycle bakes in (kernel arguments): coefficient pointers of level 0, the dense
    // inverse, smoothing parameters, eigenvalue estimates.  A replay is only valid while these are unchanged.
    struct GraphSig {
        const void *etas = nullptr, *etan = nullptr, *cinv = nullptr;
        int nu = 0, nu_coarse = 0, nc = 0, tile = 0, all = 0, level = -1;
        double cheb_ratio = 0;
        std::vector<double> lmax;
        bool operator==(const GraphSig& o) const {
            return etas == o.etas && etan == o.etan && cinv == o.cinv && nu == o.nu && nu_coarse == o.nu_coarse &&
                   nc == o.nc && tile == o.tile && all == o.all && level == o.level && cheb_ratio == o.cheb_ratio &&
                   lmax == o.lmax;
        }
    } graph_sig;
    bool have_prev = false;
    std::vector<double*> hist;    // previous converged iterates (warm_start >= 2), newest first
    int nhist = 0;
    // fp64 floor of the scaled residual met by the last solve on the CURRENT coefficient fields (0 = none):
    // reset by plb_stokes_set_coeffs / plb_stokes_set_surfstab, so one solve never loosens a later system
    double floor_est = 0;
    int nu = 3, gcr_m = 50, coarsen_wide = 1, dense_max = 640, nu_coarse = 60, reorth = 0;
    double cheb_ratio = 8.0, kry_reorth = 1e-4;
    double rtol_accept = 1e-8;    // a solve stalled at its fp64 floor is accepted below this true residual
    // statistics of the last solve
    int last_iters = 0, last_vcycles = 0;
    double last_relres = 0, last_rtol_eff = 0;
    int last_status = 0;          // 0 converged to rtol, 1 accepted at the fp64 floor (> rtol, <= rtol_accept), 2 not converged
};

namespace {

// -------------------------------------------------------------------------------------------
// scaling constants: mineta reduction (pylamp_stokes.py:116-122)
// -------------------------------------------------------------------------------------------
__global__ void k_set1(double* p, double v) { *p = v; }

__global__ void __launch_bounds__(256)
k_min2(long long n, const double* __restrict__ a, const double* __restrict__ b, double* out) {
    double m = INFINITY;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n;
         t += (long long)gridDim.x * blockDim.x)
        m = fmin(m, fmin(a[t], b[t]));
    m = warp_min(m);
    __shared__ double s[8];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; i++) m = fmin(m, s[i]);
        atomic_min_double(out, m);
    }
}

// -------------------------------------------------------------------------------------------
// full-row operator in the reference's layout (parity path): y = A x for ALL rows, planar in/out
// -------------------------------------------------------------------------------------------
struct FullArgs {
    const double *gz, *gx;
    int bz0, bz1;          // z-wall types (NOSLIP / FREESLIP); x-walls are FREESLIP (x = 0 optionally flow-through)
    double Kc, Kb;
    int ft_x0, ai, aj;     // flow-through wall at x = 0; pressure anchor cell
};

// Free-surface stabilisation terms (SURF): with q = Dz*vz(i,j) + Dx*vx(i,j), the interior z-momentum row
// gains theta*dt*g_z*q and the interior x-momentum row theta*dt*g_x*q, where
//   Dz(i,j) = (rho[i+1,j]+rho[i+1,j+1]-rho[i-1,j]-rho[i-1,j+1]) / 2 / (z[i+1]-z[i-1]),
//   Dx(i,j) = (rho[i,j+1]+rho[i+1,j+1]-rho[i,j-1]-rho[i+1,j-1]) / 2 / (x[j+1]-x[j-1])
// (pylamp_stokes.py:422-426, :483-487).  sd = [Dz | Dx], two full-size planes indexed with global rows.
struct SurfArgs {
    const double* dz;
    const double* dx;
    double cz, cx;        // theta*dt*g_z, theta*dt*g_x
};

__global__ void __launch_bounds__(BX* BY)
k_surfstab_planes(LevelDev L, const double* __restrict__ rho, double* __restrict__ dzp, double* __restrict__ dxp) {
    const int j = blockIdx.x * BX + threadIdx.x, i = blockIdx.y * BY + threadIdx.y;
    if (i >= L.nz || j >= L.nxx) return;
    const int ld = L.ld;
    const long long o = (long long)i * ld + j;
    double a = 0, b = 0;
    if (i >= 1 && i <= L.nz - 2 && j >= 1 && j <= L.nxx - 2) {
        a = 0.5 * (rho[o + ld] + rho[o + ld + 1] - rho[o - ld] - rho[o - ld + 1]) * L.idzc[i];
        b = 0.5 * (rho[o + 1] + rho[o + ld + 1] - rho[o - 1] - rho[o + ld - 1]) * L.idxc[j];
    }
    dzp[o] = a, dxp[o] = b;
}

template <bool SURF>
__global__ void __launch_bounds__(BX* BY)
k_stokes_full(LevelDev L, FullArgs a, SurfArgs sf, const double* __restrict__ vz, const double* __restrict__ vx,
              const double* __restrict__ p, double* __restrict__ yz, double* __restrict__ yx,
              double* __restrict__ yp) {
    const int j = blockIdx.x * BX + threadIdx.x, i = L.i0 + blockIdx.y * BY + threadIdx.y;
    if (i >= L.i1 || j >= L.nxx) return;
    const int nz = L.nz, nxx = L.nxx, ld = L.ld;
    const long long o = (long long)i * ld + j;
    const double Kc = a.Kc;
    // ---- vz row
    double r;
    if (j == nxx - 1 || i == 0 || i == nz - 1) {
        r = Kc * vz[o];                                              // ghost / wall-normal rows
    } else if (j == 0) {
        r = Kc * (vz[o] - vz[o + 1]);                                // free slip, x=0
    } else if (j == nxx - 2) {
        r = Kc * (vz[o] - vz[o - 1]);                                // free slip, x=L
    } else {
        VzCoef c = vz_coef(L, i, j);
        r = kvz_apply(L, c, vz, vx, i, j) - 2 * Kc * L.idzc[i] * (p[o] - p[o - ld]);
        if (SURF) r += sf.cz * (sf.dz[o] * vz[o] + sf.dx[o] * vx[o]);
    }
    yz[o] = r;
    // ---- vx row
    if (j == 0 && a.ft_x0 && i <= nz - 2) {
        r = Kc * (vx[o + 1] - vx[o]);                                // flow-through wall: dvx/dx = 0, :268-273
    } else if (i == nz - 1 || j == 0 || j == nxx - 1) {
        r = Kc * vx[o];
    } else if (i == 0) {
        if (a.bz0 == PLB_BC_FREESLIP) {
            r = Kc * (vx[o] - vx[o + ld]);
        } else {                                                     // no slip, :163-168
            double d2 = a.gz[2] - a.gz[0], d1 = a.gz[1] - a.gz[0];
            r = Kc * (-1 / d2 + (-1) / d1) * vx[o] + Kc * (1 / d2) * vx[o + ld];
        }
    } else if (i == nz - 2) {
        if (a.bz1 == PLB_BC_FREESLIP) {
            r = Kc * (vx[o] - vx[o - ld]);
        } else {                                                     // :202-207
            double d2 = a.gz[nz - 3] - a.gz[nz - 1], d1 = a.gz[nz - 2] - a.gz[nz - 1];
            r = Kc * (-1 / d2 + (-1) / d1) * vx[o] + Kc * (1 / d2) * vx[o - ld];
        }
    } else {
        VxCoef c = vx_coef(L, i, j);
        r = kvx_apply(L, c, vz, vx, i, j) - 2 * Kc * L.idxc[j] * (p[o] - p[o - 1]);
        if (SURF) r += sf.cx * (sf.dz[o] * vz[o] + sf.dx[o] * vx[o]);
    }
    yx[o] = r;
    // ---- pressure row
    if (i == nz - 1 || j == nxx - 1) {
        r = Kc * p[o];
    } else if ((i == 0 || i == nz - 2) && (j == 0 || j == nxx - 2)) {
        r = (j == 0) ? a.Kb * (p[o + 1] - p[o]) : a.Kb * (p[o - 1] - p[o]);      // :358-369
    } else if (i == a.ai && j == a.aj) {
        r = Kc * p[o];                                               // anchor, :525-551
    } else {
        r = Kc * (L.idx[j] * (vx[o + 1] - vx[o]) + L.idz[i] * (vz[o + ld] - vz[o]));
    }
    yp[o] = r;
}

// rhs planes (pylamp_stokes.py:429, :490); zero on every non-momentum row
__global__ void __launch_bounds__(BX* BY)
k_stokes_rhs(LevelDev L, const double* __restrict__ rho, double g_z, double g_x,
             double* __restrict__ bz, double* __restrict__ bx, double* __restrict__ bp) {
    const int j = blockIdx.x * BX + threadIdx.x, i = L.i0 + blockIdx.y * BY + threadIdx.y;
    if (i >= L.i1 || j >= L.nxx) return;
    const long long o = (long long)i * L.ld + j;
    bz[o] = is_vz_row(L, i, j) ? -0.5 * (rho[o] + rho[o + 1]) * g_z : 0.0;
    bx[o] = is_vx_row(L, i, j) ? -0.5 * (rho[o] + rho[o + L.ld]) * g_x : 0.0;
    bp[o] = 0.0;
}

__global__ void __launch_bounds__(256)
k_interleave(long long n, const double* __restrict__ a, const double* __restrict__ b,
             const double* __restrict__ c, double* __restrict__ x) {
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n;
         t += (long long)gridDim.x * blockDim.x) {
        x[3 * t] = a[t], x[3 * t + 1] = b[t], x[3 * t + 2] = c[t];
    }
}
__global__ void __launch_bounds__(256)
k_deinterleave(long long n, const double* __restrict__ x, double* __restrict__ a,
               double* __restrict__ b, double* __restrict__ c) {
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n;
         t += (long long)gridDim.x * blockDim.x) {
        a[t] = x[3 * t], b[t] = x[3 * t + 1], c[t] = x[3 * t + 2];
    }
}

// full-size interleaved vector (reference layout) -> local planes, owned rows only
__global__ void __launch_bounds__(BX* BY)
k_deinterleave_rows(LevelDev L, const double* __restrict__ x, double* __restrict__ a, double* __restrict__ b,
                    double* __restrict__ c) {
    const int j = blockIdx.x * BX + threadIdx.x, i = L.i0 + blockIdx.y * BY + threadIdx.y;
    if (i >= L.i1 || j >= L.nxx) return;
    const long long o = (long long)i * L.ld + j, t = (long long)i * L.nxx + j;
    a[o] = x[3 * t], b[o] = x[3 * t + 1], c[o] = x[3 * t + 2];
}

// -------------------------------------------------------------------------------------------
// reduced (BC-eliminated) saddle-point operator with the weighted-norm row scaling
//   RESID: out = W (b - A x)   else   out = W (A x);   W_v = 1/sqrt|diag K|, W_p = sqrt(eta_n)/Kc
// x must satisfy the homogeneous BC rows (slaves filled); out is zero on non-rows.
// -------------------------------------------------------------------------------------------
template <bool RESID, bool SURF>
__global__ void __launch_bounds__(BX* BY)
k_stokes_op(LevelDev L, SurfArgs sf, double Kc, const double* __restrict__ vz, const double* __restrict__ vx,
            const double* __restrict__ p, const double* __restrict__ bz, const double* __restrict__ bx,
            const double* __restrict__ bp, double* __restrict__ oz, double* __restrict__ ox,
            double* __restrict__ op) {
    const int j = blockIdx.x * BX + threadIdx.x, i = L.i0 + blockIdx.y * BY + threadIdx.y;
    if (i >= L.i1 || j >= L.nxx) return;
    const int ld = L.ld;
    const long long o = (long long)i * ld + j;
    if (is_interior(L, i, j)) {
        // interior fast path: all loads up front (see rows_interior), stores predicated
        const double p00 = p[o], pM0 = p[o - ld], p0M = p[o - 1];
        const double en = L.etan[o];
        double bzv = 0, bxv = 0, bpv = 0;
        if (RESID) bzv = bz[o], bxv = bx[o], bpv = bp[o];
        const double vz00 = vz[o], vzP0 = vz[o + ld], vx00 = vx[o], vx0P = vx[o + 1];
        const Rows2 q = rows_interior<true>(L, vz, vx, i, j);
        double sq = 0;
        if (SURF) sq = sf.dz[o] * vz00 + sf.dx[o] * vx00;
        double a = q.kz - 2 * Kc * L.idzc[i] * (p00 - pM0);
        if (SURF) a += sf.cz * sq;
        oz[o] = is_vz_row(L, i, j) ? (RESID ? bzv - a : a) * rsqrt(-q.dz) : 0.0;
        a = q.kx - 2 * Kc * L.idxc[j] * (p00 - p0M);
        if (SURF) a += sf.cx * sq;
        ox[o] = is_vx_row(L, i, j) ? (RESID ? bxv - a : a) * rsqrt(-q.dx) : 0.0;
        a = Kc * (L.idx[j] * (vx0P - vx00) + L.idz[i] * (vzP0 - vz00));
        op[o] = is_p_row(L, i, j) ? (RESID ? bpv - a : a) * (sqrt(en) / Kc) : 0.0;
        return;
    }
    double r = 0;
    if (is_vz_row(L, i, j)) {
        VzCoef c = vz_coef(L, i, j);
        double a = kvz_apply(L, c, vz, vx, i, j) - 2 * Kc * L.idzc[i] * (p[o] - p[o - ld]);
        if (SURF) a += sf.cz * (sf.dz[o] * vz[o] + sf.dx[o] * vx[o]);
        r = (RESID ? bz[o] - a : a) * rsqrt(-c.diag);
    }
    oz[o] = r;
    r = 0;
    if (is_vx_row(L, i, j)) {
        VxCoef c = vx_coef(L, i, j);
        double a = kvx_apply(L, c, vz, vx, i, j) - 2 * Kc * L.idxc[j] * (p[o] - p[o - 1]);
        if (SURF) a += sf.cx * (sf.dz[o] * vz[o] + sf.dx[o] * vx[o]);
        r = (RESID ? bx[o] - a : a) * rsqrt(-c.diag);
    }
    ox[o] = r;
    r = 0;
    if (is_p_row(L, i, j)) {
        double a = Kc * (L.idx[j] * (vx[o + 1] - vx[o]) + L.idz[i] * (vz[o + ld] - vz[o]));
        r = (RESID ? bp[o] - a : a) * (sqrt(L.etan[o]) / Kc);
    }
    op[o] = r;
}

// preconditioner, stage 1: dp = S^-1 r_p with S = Kc^2/eta_n (diagonal), and the right-hand side
// of the velocity-block solve  bv = r_v - G dp   (r = W^-1 r^ un-scaled on the fly)
// The three scalings depend on the coefficients only: tabulated once per plb_stokes_set_coeffs instead of five square
// roots and three divisions per node and Krylov iteration (ncu: k_precond_rhs was bound by exactly those -- issue
// slots 67 % busy at 2.6 TB/s).  wp = sqrt(eta_n)/Kc where a pressure lives, sdz / sdx = sqrt(-diag) on the momentum rows.
__global__ void __launch_bounds__(BX* BY)
k_scale_planes(LevelDev L, int row0, double invKc, double* __restrict__ wp, double* __restrict__ sdz,
               double* __restrict__ sdx) {
    const int j = blockIdx.x * BX + threadIdx.x, i = row0 + blockIdx.y * BY + threadIdx.y;
    if (i >= L.i1 || j >= L.nxx) return;
    const long long o = (long long)i * L.ld + j;
    wp[o] = (i <= L.nz - 2 && j <= L.nxx - 2) ? sqrt(L.etan[o]) * invKc : 0.0;
    if (i < L.i0) return;                       // (the halo row below: only its pressure scaling is read)
    sdz[o] = is_vz_row(L, i, j) ? sqrt(-vz_coef(L, i, j).diag) : 0.0;
    sdx[o] = is_vx_row(L, i, j) ? sqrt(-vx_coef(L, i, j).diag) : 0.0;
}

__global__ void __launch_bounds__(BX* BY)
k_precond_rhs(LevelDev L, double Kc, const double* __restrict__ wp, const double* __restrict__ sdz,
              const double* __restrict__ sdx, const double* __restrict__ rz, const double* __restrict__ rx,
              const double* __restrict__ rp, double* __restrict__ zp, double* __restrict__ bvz,
              double* __restrict__ bvx) {
    const int j = blockIdx.x * BX + threadIdx.x, i = L.i0 + blockIdx.y * BY + threadIdx.y;
    if (i >= L.i1 || j >= L.nxx) return;
    const int ld = L.ld, nz = L.nz, nxx = L.nxx;
    const long long o = (long long)i * ld + j;
    if (is_interior(L, i, j) && i <= nz - 3 && j <= nxx - 3) {
        // interior fast path (no corner cell here): all loads up front
        const double r0 = rp[o], rM = rp[o - ld], rW = rp[o - 1];
        const double w0 = wp[o], wM = wp[o - ld], wW = wp[o - 1];
        const double rzv = rz[o], rxv = rx[o], sz = sdz[o], sx = sdx[o];
        const double d0 = r0 * w0, dM = rM * wM, dW = rW * wW;     // (rp/W_p) / (Kc^2/eta) = rp * sqrt(eta)/Kc
        zp[o] = d0;
        bvz[o] = is_vz_row(L, i, j) ? rzv * sz + 2 * Kc * L.idzc[i] * (d0 - dM) : 0.0;
        bvx[o] = is_vx_row(L, i, j) ? rxv * sx + 2 * Kc * L.idxc[j] * (d0 - dW) : 0.0;
        return;
    }
    double dp = 0;
    if (i <= nz - 2 && j <= nxx - 2) {
        bool corner = (i == 0 || i == nz - 2) && (j == 0 || j == nxx - 2);
        // corner pressures are slaves of their x-neighbour (pylamp_stokes.py:358-369)
        const long long q = corner ? (j == 0 ? o + 1 : o - 1) : o;
        dp = rp[q] * wp[q];
    }
    zp[o] = dp;
    double b = 0;
    if (is_vz_row(L, i, j)) b = rz[o] * sdz[o] + 2 * Kc * L.idzc[i] * (rp[o] * wp[o] - rp[o - ld] * wp[o - ld]);
    bvz[o] = b;
    b = 0;
    if (is_vx_row(L, i, j)) b = rx[o] * sdx[o] + 2 * Kc * L.idxc[j] * (rp[o] * wp[o] - rp[o - 1] * wp[o - 1]);
    bvx[o] = b;
}

// -------------------------------------------------------------------------------------------
// multigrid on the velocity block K
// -------------------------------------------------------------------------------------------
// one Chebyshev-Jacobi step:  r = D^-1 (b - K x);  d = cd*d + cr*r;  xout = x + d
// FIRST (x == 0): d = cr * D^-1 b; xout = d
template <bool FIRST>
__global__ void __launch_bounds__(BX* BY)
k_cheb(LevelDev L, const double* __restrict__ xz, const double* __restrict__ xx,
       const double* __restrict__ bz, const double* __restrict__ bx, double* __restrict__ dz,
       double* __restrict__ dx, double* __restrict__ oz, double* __restrict__ ox, double cd, double cr) {
    const int j = blockIdx.x * BX + threadIdx.x, i = L.i0 + blockIdx.y * BY + threadIdx.y;
    if (i >= L.i1 || j >= L.nxx) return;
    const long long o = (long long)i * L.ld + j;
    const bool rz = is_vz_row(L, i, j), rx = is_vx_row(L, i, j);
    if (is_interior(L, i, j)) {
        // all loads up front, stores predicated
        const double bzv = bz[o], bxv = bx[o];
        double dzv = 0, dxv = 0, xzv = 0, xxv = 0;
        if (!FIRST) dzv = dz[o], dxv = dx[o], xzv = xz[o], xxv = xx[o];
        const Rows2 r = rows_interior<!FIRST>(L, xz, xx, i, j);
        if (rz) {
            double d = FIRST ? cr * bzv / r.dz : cd * dzv + cr * ((bzv - r.kz) / r.dz);
            dz[o] = d;
            store_vz(L, oz, i, j, FIRST ? d : xzv + d);
        }
        if (rx) {
            double d = FIRST ? cr * bxv / r.dx : cd * dxv + cr * ((bxv - r.kx) / r.dx);
            dx[o] = d;
            store_vx(L, ox, i, j, FIRST ? d : xxv + d);
        }
        return;
    }
    if (rz) {
        VzCoef c = vz_coef(L, i, j);
        double d;
        if (FIRST) {
            d = cr * bz[o] / c.diag;
            dz[o] = d;
            store_vz(L, oz, i, j, d);
        } else {
            double r = (bz[o] - kvz_apply(L, c, xz, xx, i, j)) / c.diag;
            d = cd * dz[o] + cr * r;
            dz[o] = d;
            store_vz(L, oz, i, j, xz[o] + d);
        }
    }
    if (rx) {
        VxCoef c = vx_coef(L, i, j);
        double d;
        if (FIRST) {
            d = cr * bx[o] / c.diag;
            dx[o] = d;
            store_vx(L, ox, i, j, d);
        } else {
            double r = (bx[o] - kvx_apply(L, c, xz, xx, i, j)) / c.diag;
            d = cd * dx[o] + cr * r;
            dx[o] = d;
            store_vx(L, ox, i, j, xx[o] + d);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Shared-memory-staged variant of the (non-first) Chebyshev-Jacobi sweep.  A CTA of 256 threads
// owns a 16 x 64 tile of nodes; the four stencil inputs (vz, vx with a one-node halo; eta_s, eta_n
// with their one-sided halos) are staged with asynchronous 8-byte global->shared copies (LDGSTS,
// no register staging), b and d are read straight into registers meanwhile, and each thread then
// computes 4 nodes from shared memory.  Nodes on the domain boundary take the generic path.
// Same arithmetic as rows_interior().
// ---------------------------------------------------------------------------------------------
constexpr int TLX = 64, TLZ = 16, TLP = TLX + 2, TLE = TLX + 1;

// issue the asynchronous copies of the velocity and viscosity tiles of the tile whose first node is
// (i00, j00); rows outside the stored range [row_lo,row_hi] or the grid read as 0.  No commit/wait.
__device__ __forceinline__ void stage_tiles(const LevelDev& L, int row_lo, int row_hi, int i00, int j00,
                                            const double* __restrict__ xz, const double* __restrict__ xx,
                                            double* sz, double* sx, double* ses, double* sen) {
    const int tid = threadIdx.x, ld = L.ld;
    for (int e = tid; e < (TLZ + 2) * TLP; e += 256) {
        const int r = e / TLP, c = e - r * TLP;
        const int gi = i00 - 1 + r, gj = j00 - 1 + c;
        if (gi >= row_lo && gi <= row_hi && gj >= 0 && gj < L.nxx) {
            const long long o = (long long)gi * ld + gj;
            __pipeline_memcpy_async(&sz[e], &xz[o], 8);
            __pipeline_memcpy_async(&sx[e], &xx[o], 8);
        } else {
            sz[e] = 0.0, sx[e] = 0.0;
        }
    }
    for (int e = tid; e < (TLZ + 1) * TLE; e += 256) {
        const int r = e / TLE, c = e - r * TLE;
        {   // eta_s rows i00..i00+TLZ, cols j00..j00+TLX (full-size replicated field: global rows)
            const int gi = i00 + r, gj = j00 + c;
            if (gi < L.nz && gj < L.nxx) __pipeline_memcpy_async(&ses[e], &L.etas[(long long)gi * ld + gj], 8);
            else ses[e] = 0.0;
        }
        {   // eta_n rows i00-1.., cols j00-1..
            const int gi = i00 - 1 + r, gj = j00 - 1 + c;
            if (gi >= 0 && gi < L.nz && gj >= 0 && gj < L.nxx) __pipeline_memcpy_async(&sen[e], &L.etan[(long long)gi * ld + gj], 8);
            else sen[e] = 0.0;
        }
    }
}

struct TileRows {
    double kz, dgz, kx, dgx;      // (K v)_vz, its diagonal, (K v)_vx, its diagonal
    double z00, x00, zP0, x0P;    // centre values and the two neighbours the continuity row needs
};

// both momentum rows of interior node (i,j) = tile-local (r, tx) from the staged tiles
__device__ __forceinline__ TileRows rows_from_tiles(const LevelDev& L, int i, int j, int r, int tx, const double* sz,
                                                    const double* sx, const double* ses, const double* sen,
                                                    double idx_j, double idx_m, double idxc_j, double idxc_p) {
    const double* pz = &sz[(r + 1) * TLP + tx + 1];
    const double* px = &sx[(r + 1) * TLP + tx + 1];
    const double* pes = &ses[r * TLE + tx];
    const double* pen = &sen[(r + 1) * TLE + tx + 1];
    const double enC = pen[0], enM0 = pen[-TLE], en0M = pen[-1];
    const double esC = pes[0], es0P = pes[1], esP0 = pes[TLE];
    const double idz_i = L.idz[i], idz_m = L.idz[i - 1], idzc_i = L.idzc[i], idzc_p = L.idzc[i + 1];
    const double z00 = pz[0], zP0 = pz[TLP], zM0 = pz[-TLP], z0P = pz[1], z0M = pz[-1], zPM = pz[TLP - 1];
    const double x00 = px[0], x0P = px[1], x0M = px[-1], xP0 = px[TLP], xM0 = px[-TLP], xMP = px[-TLP + 1];
    TileRows t;
    t.z00 = z00, t.x00 = x00, t.zP0 = zP0, t.x0P = x0P;
    {
        double cN = 4 * enC * idz_i * idzc_i, cS = 4 * enM0 * idz_m * idzc_i;
        double cE = 2 * es0P * idxc_p * idx_j, cW = 2 * esC * idxc_j * idx_j;
        double xE = 2 * es0P * idzc_i * idx_j, xW = 2 * esC * idzc_i * idx_j;
        if (L.proper && j + 1 == L.nxx - 1) cE = 0, xE = 0;
        t.dgz = -(cN + cS + cE + cW);
        t.kz = cN * zP0 + cS * zM0 + cE * z0P + cW * z0M + t.dgz * z00 + xE * (x0P - xMP) - xW * (x00 - xM0);
    }
    {
        double cE = 4 * enC * idx_j * idxc_j, cW = 4 * en0M * idx_m * idxc_j;
        double cS = 2 * esP0 * idzc_p * idz_i, cN = 2 * esC * idzc_i * idz_i;
        double xS = 2 * esP0 * idxc_j * idz_i, xN = 2 * esC * idxc_j * idz_i;
        double wall = 0;
        if (L.proper && i + 1 == L.nz - 1) {
            if (L.ns_z1) wall = 2 * esP0 * idz_i * idz_i;
            cS = 0, xS = 0;
        }
        t.dgx = -(cE + cW + cS + cN + wall);
        t.kx = cE * x0P + cW * x0M + cS * xP0 + cN * xM0 + t.dgx * x00 + xS * (zP0 - zPM) - xN * (z00 - z0M);
    }
    return t;
}

__global__ void __launch_bounds__(256)
k_cheb_tile(LevelDev L, int row_lo, int row_hi, const double* __restrict__ xz, const double* __restrict__ xx,
            const double* __restrict__ bz, const double* __restrict__ bx, double* __restrict__ dz,
            double* __restrict__ dx, double* __restrict__ oz, double* __restrict__ ox, double cd, double cr) {
    __shared__ double sz[(TLZ + 2) * TLP], sx[(TLZ + 2) * TLP], ses[(TLZ + 1) * TLE], sen[(TLZ + 1) * TLE];
    const int tid = threadIdx.x;
    const int i00 = L.i0 + blockIdx.y * TLZ, j00 = blockIdx.x * TLX;
    const int ld = L.ld;
    stage_tiles(L, row_lo, row_hi, i00, j00, xz, xx, sz, sx, ses, sen);
    __pipeline_commit();
    // ---- per-thread nodes: column tx, rows ty + 4k
    const int tx = tid & 63, ty = tid >> 6;
    const int j = j00 + tx;
    const double idx_j = (j < L.nxx) ? L.idx[j] : 0.0, idx_m = (j >= 1 && j < L.nxx) ? L.idx[j - 1] : 0.0;
    const double idxc_j = (j < L.nxx) ? L.idxc[j] : 0.0, idxc_p = (j + 1 < L.nxx) ? L.idxc[j + 1] : 0.0;
    __pipeline_wait_prior(0);
    __syncthreads();
#pragma unroll 1
    for (int k = 0; k < 4; k++) {
        const int r = ty + 4 * k, i = i00 + r;
        if (i >= L.i1 || j >= L.nxx) continue;
        const long long o = (long long)i * ld + j;
        const double vbz = bz[o], vbx = bx[o], vdz = dz[o], vdx = dx[o];     // issued before the smem reads
        const bool rowz = is_vz_row(L, i, j), rowx = is_vx_row(L, i, j);
        if (!is_interior(L, i, j)) {
            // boundary nodes: generic path on global memory (same as k_cheb<false>)
            if (rowz) {
                VzCoef c = vz_coef(L, i, j);
                double q = (vbz - kvz_apply(L, c, xz, xx, i, j)) / c.diag;
                double d = cd * vdz + cr * q;
                dz[o] = d;
                store_vz(L, oz, i, j, xz[o] + d);
            }
            if (rowx) {
                VxCoef c = vx_coef(L, i, j);
                double q = (vbx - kvx_apply(L, c, xz, xx, i, j)) / c.diag;
                double d = cd * vdx + cr * q;
                dx[o] = d;
                store_vx(L, ox, i, j, xx[o] + d);
            }
            continue;
        }
        const TileRows t = rows_from_tiles(L, i, j, r, tx, sz, sx, ses, sen, idx_j, idx_m, idxc_j, idxc_p);
        if (rowz) {
            double d = cd * vdz + cr * ((vbz - t.kz) / t.dgz);
            dz[o] = d;
            store_vz(L, oz, i, j, t.z00 + d);
        }
        if (rowx) {
            double d = cd * vdx + cr * ((vbx - t.kx) / t.dgx);
            dx[o] = d;
            store_vx(L, ox, i, j, t.x00 + d);
        }
    }
}

// shared-memory-staged variant of k_stokes_op (pressure staged as a fifth tile)
template <bool RESID, bool SURF>
__global__ void __launch_bounds__(256)
k_stokes_op_tile(LevelDev L, SurfArgs sf, int row_lo, int row_hi, double Kc, const double* __restrict__ vz,
                 const double* __restrict__ vx, const double* __restrict__ p, const double* __restrict__ bz,
                 const double* __restrict__ bx, const double* __restrict__ bp, double* __restrict__ oz,
                 double* __restrict__ ox, double* __restrict__ op) {
    __shared__ double sz[(TLZ + 2) * TLP], sx[(TLZ + 2) * TLP], ses[(TLZ + 1) * TLE], sen[(TLZ + 1) * TLE],
        sp[(TLZ + 1) * TLE];
    const int tid = threadIdx.x;
    const int i00 = L.i0 + blockIdx.y * TLZ, j00 = blockIdx.x * TLX;
    const int ld = L.ld;
    stage_tiles(L, row_lo, row_hi, i00, j00, vz, vx, sz, sx, ses, sen);
    for (int e = tid; e < (TLZ + 1) * TLE; e += 256) {      // p rows i00-1.., cols j00-1.. (like eta_n)
        const int r = e / TLE, c = e - r * TLE;
        const int gi = i00 - 1 + r, gj = j00 - 1 + c;
        if (gi >= row_lo && gi <= row_hi && gj >= 0 && gj < L.nxx) __pipeline_memcpy_async(&sp[e], &p[(long long)gi * ld + gj], 8);
        else sp[e] = 0.0;
    }
    __pipeline_commit();
    const int tx = tid & 63, ty = tid >> 6;
    const int j = j00 + tx;
    const double idx_j = (j < L.nxx) ? L.idx[j] : 0.0, idx_m = (j >= 1 && j < L.nxx) ? L.idx[j - 1] : 0.0;
    const double idxc_j = (j < L.nxx) ? L.idxc[j] : 0.0, idxc_p = (j + 1 < L.nxx) ? L.idxc[j + 1] : 0.0;
    const double invKc = 1.0 / Kc;
    __pipeline_wait_prior(0);
    __syncthreads();
#pragma unroll 1
    for (int k = 0; k < 4; k++) {
        const int r = ty + 4 * k, i = i00 + r;
        if (i >= L.i1 || j >= L.nxx) continue;
        const long long o = (long long)i * ld + j;
        double vbz = 0, vbx = 0, vbp = 0;
        if (RESID) vbz = bz[o], vbx = bx[o], vbp = bp[o];
        if (!is_interior(L, i, j)) {
            // boundary nodes: generic path (same as k_stokes_op)
            double q = 0;
            if (is_vz_row(L, i, j)) {
                VzCoef c = vz_coef(L, i, j);
                double a = kvz_apply(L, c, vz, vx, i, j) - 2 * Kc * L.idzc[i] * (p[o] - p[o - ld]);
                if (SURF) a += sf.cz * (sf.dz[o] * vz[o] + sf.dx[o] * vx[o]);
                q = (RESID ? vbz - a : a) * rsqrt(-c.diag);
            }
            oz[o] = q;
            q = 0;
            if (is_vx_row(L, i, j)) {
                VxCoef c = vx_coef(L, i, j);
                double a = kvx_apply(L, c, vz, vx, i, j) - 2 * Kc * L.idxc[j] * (p[o] - p[o - 1]);
                if (SURF) a += sf.cx * (sf.dz[o] * vz[o] + sf.dx[o] * vx[o]);
                q = (RESID ? vbx - a : a) * rsqrt(-c.diag);
            }
            ox[o] = q;
            q = 0;
            if (is_p_row(L, i, j)) {
                double a = Kc * (L.idx[j] * (vx[o + 1] - vx[o]) + L.idz[i] * (vz[o + ld] - vz[o]));
                q = (RESID ? vbp - a : a) * (sqrt(L.etan[o]) / Kc);
            }
            op[o] = q;
            continue;
        }
        const TileRows t = rows_from_tiles(L, i, j, r, tx, sz, sx, ses, sen, idx_j, idx_m, idxc_j, idxc_p);
        const double* pp = &sp[(r + 1) * TLE + tx + 1];
        const double p00 = pp[0], pM0 = pp[-TLE], p0M = pp[-1];
        const double en = sen[(r + 1) * TLE + tx + 1];
        double sq = 0;
        if (SURF) sq = sf.dz[o] * t.z00 + sf.dx[o] * t.x00;
        double a = t.kz - 2 * Kc * L.idzc[i] * (p00 - pM0);
        if (SURF) a += sf.cz * sq;
        oz[o] = is_vz_row(L, i, j) ? (RESID ? vbz - a : a) * rsqrt(-t.dgz) : 0.0;
        a = t.kx - 2 * Kc * idxc_j * (p00 - p0M);
        if (SURF) a += sf.cx * sq;
        ox[o] = is_vx_row(L, i, j) ? (RESID ? vbx - a : a) * rsqrt(-t.dgx) : 0.0;
        a = Kc * (idx_j * (t.x0P - t.x00) + L.idz[i] * (t.zP0 - t.z00));
        op[o] = is_p_row(L, i, j) ? (RESID ? vbp - a : a) * (sqrt(en) * invKc) : 0.0;
    }
}

// MODE 0: r = b - K x;  MODE 1: r = D^-1 K x (power iteration);  zero on non-rows
template <int MODE>
__global__ void __launch_bounds__(BX* BY)
k_vel_op(LevelDev L, const double* __restrict__ xz, const double* __restrict__ xx,
         const double* __restrict__ bz, const double* __restrict__ bx, double* __restrict__ rz,
         double* __restrict__ rx) {
    const int j = blockIdx.x * BX + threadIdx.x, i = L.i0 + blockIdx.y * BY + threadIdx.y;
    if (i >= L.i1 || j >= L.nxx) return;
    const long long o = (long long)i * L.ld + j;
    if (MODE == 0 && is_interior(L, i, j)) {
        const double bzv = bz[o], bxv = bx[o];
        const Rows2 q = rows_interior<true>(L, xz, xx, i, j);
        rz[o] = is_vz_row(L, i, j) ? bzv - q.kz : 0.0;
        rx[o] = is_vx_row(L, i, j) ? bxv - q.kx : 0.0;
        return;
    }
    double r = 0;
    if (is_vz_row(L, i, j)) {
        VzCoef c = vz_coef(L, i, j);
        double a = kvz_apply(L, c, xz, xx, i, j);
        r = MODE == 0 ? bz[o] - a : a / c.diag;
    }
    if (MODE == 1) { if (is_vz_row(L, i, j)) store_vz(L, rz, i, j, r); }
    else rz[o] = r;
    r = 0;
    if (is_vx_row(L, i, j)) {
        VxCoef c = vx_coef(L, i, j);
        double a = kvx_apply(L, c, xz, xx, i, j);
        r = MODE == 0 ? bx[o] - a : a / c.diag;
    }
    if (MODE == 1) { if (is_vx_row(L, i, j)) store_vx(L, rx, i, j, r); }
    else rx[o] = r;
}

// deterministic pseudo-random start vector for the power iteration (rows + slaves)
__global__ void __launch_bounds__(BX* BY)
k_fill_random(LevelDev L, double* __restrict__ xz, double* __restrict__ xx) {
    const int j = blockIdx.x * BX + threadIdx.x, i = L.i0 + blockIdx.y * BY + threadIdx.y;
    if (i >= L.i1 || j >= L.nxx) return;
    unsigned long long h = ((unsigned long long)i * 2654435761ull) ^ ((unsigned long long)j * 40503ull + 12345ull);
    h ^= h >> 13, h *= 0x9E3779B97F4A7C15ull, h ^= h >> 29;
    double u = (double)(h & 0xFFFFFF) / 16777216.0 - 0.5;
    if (is_vz_row(L, i, j)) store_vz(L, xz, i, j, u);
    if (is_vx_row(L, i, j)) store_vx(L, xx, i, j, 0.7 * u + 0.1);
}

// restriction of the fine residual (zero on non-rows) to the coarse right-hand side:
// b_c = 1/4 P^T r  with bilinear P (nodal direction: 1/2,1,1/2; staggered direction: 1/4,3/4,3/4,1/4
// with constant extrapolation at the ends)
__global__ void __launch_bounds__(BX* BY)
k_restrict(LevelDev F, LevelDev Cc, const double* __restrict__ rz, const double* __restrict__ rx,
           double* __restrict__ bz, double* __restrict__ bx) {
    const int J = blockIdx.x * BX + threadIdx.x, I = Cc.i0 + blockIdx.y * BY + threadIdx.y;
    if (I >= Cc.i1 || J >= Cc.nxx) return;
    const long long oc = (long long)I * Cc.ld + J;
    const int ldf = F.ld;
    double s = 0;
    if (is_vz_row(Cc, I, J)) {
        const double wn[3] = {0.5, 1.0, 0.5};
        double wm[4] = {0.25, 0.75, 0.75, 0.25};
        if (J == 0) wm[1] = 1.0;                       // fine j=0 interpolates from coarse 0 only
        if (J == Cc.nxx - 2) wm[2] = 1.0;              // fine j=nxx_f-2 from the last coarse mid only
#pragma unroll
        for (int a = 0; a < 3; a++) {
            int i = 2 * I - 1 + a;
            if (i < 0 || i >= F.nz) continue;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                int j = 2 * J - 1 + b;
                if (j < 0 || j >= F.nxx) continue;
                s += wn[a] * wm[b] * rz[(long long)i * ldf + j];
            }
        }
        s *= 0.25;
    }
    bz[oc] = s;
    s = 0;
    if (is_vx_row(Cc, I, J)) {
        const double wn[3] = {0.5, 1.0, 0.5};
        double wm[4] = {0.25, 0.75, 0.75, 0.25};
        if (I == 0) wm[1] = 1.0;
        if (I == Cc.nz - 2) wm[2] = 1.0;
#pragma unroll
        for (int a = 0; a < 4; a++) {
            int i = 2 * I - 1 + a;
            if (i < 0 || i >= F.nz) continue;
#pragma unroll
            for (int b = 0; b < 3; b++) {
                int j = 2 * J - 1 + b;
                if (j < 0 || j >= F.nxx) continue;
                s += wm[a] * wn[b] * rx[(long long)i * ldf + j];
            }
        }
        s *= 0.25;
    }
    bx[oc] = s;
}

// xout = xin + P e_c on the fine rows (+ slaves)
__device__ __forceinline__ void mid_parents(int j, int ncm, int& a, int& b, double& wa, double& wb) {
    int Jp = j >> 1;
    if ((j & 1) == 0) a = Jp - 1, b = Jp, wa = 0.25, wb = 0.75;
    else a = Jp, b = Jp + 1, wa = 0.75, wb = 0.25;
    a = a < 0 ? 0 : (a > ncm ? ncm : a);
    b = b < 0 ? 0 : (b > ncm ? ncm : b);
}

__global__ void __launch_bounds__(BX* BY)
k_prolong_add(LevelDev F, LevelDev Cc, const double* __restrict__ ez, const double* __restrict__ ex,
              const double* __restrict__ xz, const double* __restrict__ xx, double* __restrict__ oz,
              double* __restrict__ ox) {
    const int j = blockIdx.x * BX + threadIdx.x, i = F.i0 + blockIdx.y * BY + threadIdx.y;
    if (i >= F.i1 || j >= F.nxx) return;
    const long long o = (long long)i * F.ld + j;
    const int ldc = Cc.ld;
    if (is_vz_row(F, i, j)) {
        int a, b;
        double wa, wb;
        mid_parents(j, Cc.nxx - 2, a, b, wa, wb);
        int I = i >> 1;
        double v;
        if ((i & 1) == 0) v = wa * ez[(long long)I * ldc + a] + wb * ez[(long long)I * ldc + b];
        else v = 0.5 * (wa * ez[(long long)I * ldc + a] + wb * ez[(long long)I * ldc + b] +
                        wa * ez[(long long)(I + 1) * ldc + a] + wb * ez[(long long)(I + 1) * ldc + b]);
        store_vz(F, oz, i, j, xz[o] + v);
    }
    if (is_vx_row(F, i, j)) {
        int a, b;
        double wa, wb;
        mid_parents(i, Cc.nz - 2, a, b, wa, wb);
        int J = j >> 1;
        double v;
        if ((j & 1) == 0) v = wa * ex[(long long)a * ldc + J] + wb * ex[(long long)b * ldc + J];
        else v = 0.5 * (wa * ex[(long long)a * ldc + J] + wb * ex[(long long)b * ldc + J] +
                        wa * ex[(long long)a * ldc + J + 1] + wb * ex[(long long)b * ldc + J + 1]);
        store_vx(F, ox, i, j, xx[o] + v);
    }
}

// viscosity coarsening (arithmetic): nodes = 9-point full weighting with edge clamping; centres =
// (1,3,3,1)/8 x (1,3,3,1)/8 over the 4x4 block of fine cells around the coarse cell (wide, stable
// with sharp contrasts) or the plain mean of its 2x2 fine cells (narrow)
__global__ void __launch_bounds__(BX* BY)
k_coarsen_eta(int nzf, int nxf, int ldf, const double* __restrict__ es, const double* __restrict__ en,
              int nzc, int nxc, int ldc, double* __restrict__ cs, double* __restrict__ cn, int wide, int I0, int I1) {
    const int J = blockIdx.x * BX + threadIdx.x, I = I0 + blockIdx.y * BY + threadIdx.y;
    if (I >= I1 || J >= nxc) return;
    const double w3[3] = {0.25, 0.5, 0.25};
    double s = 0;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        int i = min(max(2 * I - 1 + a, 0), nzf - 1);
#pragma unroll
        for (int b = 0; b < 3; b++) {
            int j = min(max(2 * J - 1 + b, 0), nxf - 1);
            s += w3[a] * w3[b] * es[(long long)i * ldf + j];
        }
    }
    cs[(long long)I * ldc + J] = s;
    // centres: real fine cells are [0,nzf-2] x [0,nxf-2]; ghost coarse cells copy the edge
    int Ic = min(I, nzc - 2), Jc = min(J, nxc - 2);
    s = 0;
    if (wide) {
        const double w4[4] = {0.125, 0.375, 0.375, 0.125};
#pragma unroll
        for (int a = 0; a < 4; a++) {
            int i = min(max(2 * Ic - 1 + a, 0), nzf - 2);
#pragma unroll
            for (int b = 0; b < 4; b++) {
                int j = min(max(2 * Jc - 1 + b, 0), nxf - 2);
                s += w4[a] * w4[b] * en[(long long)i * ldf + j];
            }
        }
    } else {
        long long o = (long long)(2 * Ic) * ldf + 2 * Jc;
        s = 0.25 * (en[o] + en[o + 1] + en[o + ldf] + en[o + ldf + 1]);
    }
    cn[(long long)I * ldc + J] = s;
}

// ---- dense coarse-level solve ----------------------------------------------------------------
__device__ __forceinline__ int unk_vz(const LevelDev& L, int i, int j) {
    return (i - L.vz_i0) * (L.vz_j1 - L.vz_j0 + 1) + (j - L.vz_j0);
}
__device__ __forceinline__ int n_vz(const LevelDev& L) {
    return (L.vz_i1 - L.vz_i0 + 1) * (L.vz_j1 - L.vz_j0 + 1);
}
__device__ __forceinline__ int unk_vx(const LevelDev& L, int i, int j) {
    return n_vz(L) + (i - L.vx_i0) * (L.vx_j1 - L.vx_j0 + 1) + (j - L.vx_j0);
}

// probe k = unit vector of unknown k (with its slaves) in its own pair of planes
__global__ void k_probe_set(LevelDev L, int n, size_t plane, double* __restrict__ probes) {
    const int j = blockIdx.x * BX + threadIdx.x, i = L.i0 + blockIdx.y * BY + threadIdx.y;
    if (i >= L.i1 || j >= L.nxx) return;
    if (is_vz_row(L, i, j)) store_vz(L, probes + (size_t)unk_vz(L, i, j) * 2 * plane, i, j, 1.0);
    if (is_vx_row(L, i, j)) store_vx(L, probes + (size_t)unk_vx(L, i, j) * 2 * plane + plane, i, j, 1.0);
}

// A[row][k] = (K e_k)[row], augmented with the identity: M = [A | I], n x 2n row-major
__global__ void k_probe_apply(LevelDev L, int n, size_t plane, const double* __restrict__ probes,
                              double* __restrict__ M) {
    const int j = blockIdx.x * BX + threadIdx.x, i = blockIdx.y * BY + threadIdx.y, k = blockIdx.z;
    if (i >= L.nz || j >= L.nxx) return;
    const double* xz = probes + (size_t)k * 2 * plane;
    const double* xx = xz + plane;
    if (is_vz_row(L, i, j)) {
        VzCoef c = vz_coef(L, i, j);
        int row = unk_vz(L, i, j);
        M[(size_t)row * 2 * n + k] = kvz_apply(L, c, xz, xx, i, j);
        M[(size_t)row * 2 * n + n + k] = (row == k) ? 1.0 : 0.0;
    }
    if (is_vx_row(L, i, j)) {
        VxCoef c = vx_coef(L, i, j);
        int row = unk_vx(L, i, j);
        M[(size_t)row * 2 * n + k] = kvx_apply(L, c, xz, xx, i, j);
        M[(size_t)row * 2 * n + n + k] = (row == k) ? 1.0 : 0.0;
    }
}

// Gauss-Jordan with partial pivoting on [A | I] in one CTA; the right half becomes A^-1
__global__ void __launch_bounds__(1024) k_gauss_jordan(int n, double* __restrict__ M, int* __restrict__ fail) {
    extern __shared__ double sh[];
    double* col = sh;                       // n entries: column k before elimination
    __shared__ double pv[32];
    __shared__ int pi[32];
    __shared__ int prow;
    const int tid = threadIdx.x, nt = blockDim.x, w2 = 2 * n;
    for (int k = 0; k < n; k++) {
        double best = -1;
        int bi = k;
        for (int r = k + tid; r < n; r += nt) {
            double v = fabs(M[(size_t)r * w2 + k]);
            if (v > best) best = v, bi = r;
        }
        for (int o = 16; o > 0; o >>= 1) {
            double ob = __shfl_down_sync(0xffffffffu, best, o);
            int oi = __shfl_down_sync(0xffffffffu, bi, o);
            if (ob > best) best = ob, bi = oi;
        }
        if ((tid & 31) == 0) pv[tid >> 5] = best, pi[tid >> 5] = bi;
        __syncthreads();
        if (tid == 0) {
            for (int q = 1; q < (nt + 31) / 32; q++)
                if (pv[q] > pv[0]) pv[0] = pv[q], pi[0] = pi[q];
            prow = pi[0];
            if (!(pv[0] > 0)) *fail = 1;
        }
        __syncthreads();
        const int p = prow;
        if (p != k) {
            for (int c = tid; c < w2; c += nt) {
                double t = M[(size_t)k * w2 + c];
                M[(size_t)k * w2 + c] = M[(size_t)p * w2 + c];
                M[(size_t)p * w2 + c] = t;
            }
        }
        __syncthreads();
        const double inv = 1.0 / M[(size_t)k * w2 + k];
        __syncthreads();
        for (int c = tid; c < w2; c += nt) M[(size_t)k * w2 + c] *= inv;
        for (int r = tid; r < n; r += nt) col[r] = M[(size_t)r * w2 + k];
        __syncthreads();
        // eliminate: only columns > k of the left half and all touched columns of the right half
        for (long long e = tid; e < (long long)n * w2; e += nt) {
            int r = (int)(e / w2), c = (int)(e % w2);
            if (r == k) continue;
            double f = col[r];
            if (f != 0.0) M[(size_t)r * w2 + c] -= f * M[(size_t)k * w2 + c];
        }
        __syncthreads();
    }
}

__global__ void k_extract_inverse(int n, const double* __restrict__ M, double* __restrict__ inv) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= (long long)n * n) return;
    int r = (int)(t / n), c = (int)(t % n);
    inv[t] = M[(size_t)r * 2 * n + n + c];
}

// x = inv * b on the coarsest level (b, x in plane layout): one warp per unknown
__global__ void __launch_bounds__(256)
k_dense_solve(LevelDev L, int n, const double* __restrict__ inv, const double* __restrict__ bz,
              const double* __restrict__ bx, double* __restrict__ xz, double* __restrict__ xx) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n) return;
    const int nvz = n_vz(L), wz = L.vz_j1 - L.vz_j0 + 1, wx = L.vx_j1 - L.vx_j0 + 1;
    double s = 0;
    for (int k = lane; k < n; k += 32) {
        double bv;
        if (k < nvz) bv = bz[(long long)(L.vz_i0 + k / wz) * L.ld + L.vz_j0 + k % wz];
        else bv = bx[(long long)(L.vx_i0 + (k - nvz) / wx) * L.ld + L.vx_j0 + (k - nvz) % wx];
        s += inv[(size_t)warp * n + k] * bv;
    }
    s = warp_sum(s);
    if (lane == 0) {
        if (warp < nvz) store_vz(L, xz, L.vz_i0 + warp / wz, L.vz_j0 + warp % wz, s);
        else store_vx(L, xx, L.vx_i0 + (warp - nvz) / wx, L.vx_j0 + (warp - nvz) % wx, s);
    }
}

// ---- hydrostatic splitting --------------------------------------------------------------------
// The z-momentum right-hand side is dominated by the horizontally uniform (lithostatic) load, which
// is balanced exactly by a pressure P_h(z): A [0,0,P_h] has entries only on the vz rows and they
// do not depend on j.  Solving for the deviation from P_h makes the residual norm relative to the
// flow-driving part of the load, so that near-hydrostatic states are resolved to the same
// relative accuracy as a direct solve.  m[i] = mean_j b_vz(i,j) over the vz rows of grid row i.
__global__ void __launch_bounds__(256) k_row_mean(LevelDev L, const double* __restrict__ bz, double* __restrict__ m) {
    const int i = blockIdx.x;
    double s = 0;
    const bool mine = i >= L.i0 && i < L.i1 && i >= L.vz_i0 && i <= L.vz_i1;
    if (mine)
        for (int j = L.vz_j0 + threadIdx.x; j <= L.vz_j1; j += blockDim.x) s += bz[(long long)i * L.ld + j];
    s = warp_sum(s);
    __shared__ double sm[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int q = 1; q < 8; q++) s += sm[q];
        m[i] = mine ? s / (L.vz_j1 - L.vz_j0 + 1) : 0.0;
    }
}
// P_h per cell row with P_h(row 3) = 0 (the anchor row): -2 Kc idzc[i] (P_h[i] - P_h[i-1]) = m[i]
__global__ void k_scan_ph(LevelDev L, double Kc, const double* __restrict__ m, double* __restrict__ ph, int ai) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int nz = L.nz;
    ph[0] = 0;
    for (int i = 1; i <= nz - 2; i++) ph[i] = ph[i - 1] - m[i] / (2 * Kc * L.idzc[i]);
    ph[nz - 1] = 0;
    const double p3 = ph[ai];
    for (int i = 0; i <= nz - 2; i++) ph[i] -= p3;
}
__global__ void __launch_bounds__(BX* BY)
k_sub_row_mean(LevelDev L, const double* __restrict__ m, double* __restrict__ bz) {
    const int j = blockIdx.x * BX + threadIdx.x, i = L.i0 + blockIdx.y * BY + threadIdx.y;
    if (i >= L.i1 || j >= L.nxx) return;
    if (is_vz_row(L, i, j)) bz[(long long)i * L.ld + j] -= m[i];
}

// warm start by polynomial extrapolation in time of the last p converged iterates (equal step
// lengths assumed): x0 = sum_j c_j h_j, c_j = (-1)^j C(p, j+1): (2,-1), (3,-3,1), (4,-6,4,-1), ...; h_0 = newest
struct HistList {
    const double* h[6];
    double c[6];
    int p;
};
__global__ void __launch_bounds__(256) k_extrapolate(long long n, HistList H, double* __restrict__ x) {
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n;
         t += (long long)gridDim.x * blockDim.x) {
        double v = 0;
        for (int j = 0; j < H.p; j++) v += H.c[j] * H.h[j][t];
        x[t] = v;
    }
}

// value of the iterate's pressure in the anchor cell (3,2) (zero on ranks that do not own row 3)
__global__ void k_get_anchor(LevelDev L, const double* __restrict__ p, double* out, int ai, int aj) {
    *out = (ai >= L.i0 && ai < L.i1) ? p[(long long)ai * L.ld + aj] : 0.0;
}

// final solution: planar -> interleaved with the slaved corner pressures filled in
__global__ void __launch_bounds__(BX* BY)
k_solution_out(LevelDev L, const double* __restrict__ vz, const double* __restrict__ vx,
               const double* __restrict__ p, const double* __restrict__ ph, const double* __restrict__ panchor,
               double* __restrict__ x) {
    const int j = blockIdx.x * BX + threadIdx.x, i = L.i0 + blockIdx.y * BY + threadIdx.y;
    if (i >= L.i1 || j >= L.nxx) return;
    const long long o = (long long)i * L.ld + j, t = (long long)i * L.nxx + j;
    double pv = p[o];
    if ((i == 0 || i == L.nz - 2) && (j == 0 || j == L.nxx - 2)) pv = (j == 0) ? p[o + 1] : p[o - 1];
    pv -= *panchor;                        // P(3,2) = 0 like the reference's anchor row
    if (ph) pv += ph[i];
    if (i == L.nz - 1 || j == L.nxx - 1) pv = 0;
    x[3 * t] = vz[o], x[3 * t + 1] = vx[o], x[3 * t + 2] = pv;
}

// -------------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------------
int upload(plb_ctx* ctx, const std::vector<double>& h, double** d) {
    PLB_CUDA(ctx, cudaMalloc(d, sizeof(double) * h.size()));
    PLB_CUDA(ctx, cudaMemcpyAsync(*d, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice, ctx->stream));
    PLB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

int zalloc(plb_ctx* ctx, double** d, size_t n) {
    PLB_CUDA(ctx, cudaMalloc(d, sizeof(double) * n));
    PLB_CUDA(ctx, cudaMemsetAsync(*d, 0, sizeof(double) * n, ctx->stream));
    return 0;
}

SurfArgs surf_args(const plb_stokes* op) {
    SurfArgs s;
    s.dz = op->surf_d;
    s.dx = op->surf_d ? op->surf_d + op->lv[0].full : nullptr;
    s.cz = op->surf * op->g_z, s.cx = op->surf * op->g_x;
    return s;
}

int level_metrics(plb_ctx* ctx, Level& L) {
    const int nz = L.nz, nxx = L.nxx;
    std::vector<double> idz(nz, 0.0), idzc(nz, 0.0), idx(nxx, 0.0), idxc(nxx, 0.0);
    for (int i = 0; i + 1 < nz; i++) idz[i] = 1.0 / (L.gz[i + 1] - L.gz[i]);
    for (int i = 1; i + 1 < nz; i++) idzc[i] = 1.0 / (L.gz[i + 1] - L.gz[i - 1]);
    for (int j = 0; j + 1 < nxx; j++) idx[j] = 1.0 / (L.gx[j + 1] - L.gx[j]);
    for (int j = 1; j + 1 < nxx; j++) idxc[j] = 1.0 / (L.gx[j + 1] - L.gx[j - 1]);
    if (upload(ctx, idz, &L.idz) || upload(ctx, idzc, &L.idzc) || upload(ctx, idx, &L.idx) ||
        upload(ctx, idxc, &L.idxc))
        return 2;
    return 0;
}

void free_level(Level& L) {
    double* ptrs[] = {L.idz, L.idzc, L.idx, L.idxc, L.etas_own, L.etan_own, L.X, L.T, L.b, L.r, L.d};
    for (double* p : ptrs)
        if (p) cudaFree(p);
}

int halo(plb_stokes* op, const Level& L, double* local, int nplanes) {
    if (!L.dist) return 0;
    return plb_comm_halo_exchange(op->ctx, local, nplanes, L.plane, L.ld, L.nxx, L.lo, L.i0, L.i1);
}

// reductions over 2-plane (velocity) or 3-plane vectors of level L: owned rows only, summed over ranks
void reduce_shape(plb_stokes* op, const Level& L, int nplanes) {
    if (L.dist)
        plb_reduce_shape(&op->rws, nplanes, (long long)L.plane, (long long)(L.i0 - L.lo) * L.ld,
                         (long long)(L.i1 - L.i0) * L.ld, true);
    else
        plb_reduce_shape(&op->rws, 0, 0, 0, 0, false);
}

int build_levels(plb_stokes* op, const double* h_gz, const double* h_gx) {
    plb_ctx* ctx = op->ctx;
    const int R = plb_comm_size(ctx), rank = plb_comm_rank(ctx);
    // grid sizes of all levels
    std::vector<std::vector<double>> GZ, GX;
    {
        std::vector<double> gz(h_gz, h_gz + op->nz), gx(h_gx, h_gx + op->nxx);
        for (;;) {
            GZ.push_back(gz), GX.push_back(gx);
            int cz = (int)gz.size() - 1, cx = (int)gx.size() - 1;
            if ((cz % 2) || (cx % 2) || std::min(cz, cx) / 2 < 4) break;
            std::vector<double> ngz, ngx;
            for (size_t i = 0; i < gz.size(); i += 2) ngz.push_back(gz[i]);
            for (size_t j = 0; j < gx.size(); j += 2) ngx.push_back(gx[j]);
            gz.swap(ngz), gx.swap(ngx);
        }
    }
    const int nlev = (int)GZ.size();
    int min_rows = 64;
    if (const char* e = getenv("PLB_DIST_MIN_ROWS")) min_rows = atoi(e);
    for (int l = 0; l < nlev; l++) {
        Level L;
        const int nz = (int)GZ[l].size(), nxx = (int)GX[l].size();
        L.nz = nz, L.nxx = nxx, L.ld = nxx, L.full = (size_t)nz * nxx;
        L.gz = GZ[l], L.gx = GX[l];
        // slab-distributed while every rank keeps an even number (>= 4) of cell rows; the coarsest
        // level is always replicated (dense solve)
        const int cells = nz - 1;
        const bool prev_dist = l == 0 ? true : op->lv[l - 1].dist;
        // ... and, below `min_rows` rows per rank, a level costs less replicated (every rank smooths the
        // whole small grid) than the latency of its halo exchanges
        L.dist = R > 1 && prev_dist && l < nlev - 1 && cells % R == 0 && (cells / R) >= 4 && (cells / R) % 2 == 0 &&
                 (l == 0 || cells / R >= min_rows);
        if (R > 1 && l == 0 && !L.dist)
            PLB_FAIL(ctx, "plb_stokes_create: %d cell rows cannot be split into %d even slabs of >= 4 rows", cells, R);
        if (L.dist) {
            const int c = cells / R;
            L.i0 = rank * c, L.i1 = (rank + 1) * c + (rank == R - 1 ? 1 : 0);
            L.lo = rank > 0 ? L.i0 - 1 : 0, L.hi = rank < R - 1 ? L.i1 : nz - 1;
        } else {
            L.i0 = 0, L.i1 = nz, L.lo = 0, L.hi = nz - 1;
        }
        L.plane = (size_t)(L.hi - L.lo + 1) * L.ld;
        L.shift = (long long)L.lo * L.ld;
        L.proper = l > 0;
        if (l == 0) {
            L.vz_i0 = 1, L.vz_i1 = nz - 2, L.vz_j0 = 1, L.vz_j1 = nxx - 3;
            L.vx_i0 = 1, L.vx_i1 = nz - 3, L.vx_j0 = 1, L.vx_j1 = nxx - 2;
            L.ft_x0 = (op->bc[1] & PLB_BC_FLOWTHRU) ? 1 : 0;
            // slave factors of the tangential wall rows (pylamp_stokes.py:163-175, :202-214)
            const std::vector<double>& gz = L.gz;
            if (op->bc[0] == PLB_BC_NOSLIP) {
                double d2 = gz[2] - gz[0], d1 = gz[1] - gz[0];
                L.sl_z0 = (1 / d2) / (1 / d2 + 1 / d1);
            }
            if (op->bc[2] == PLB_BC_NOSLIP) {
                double d2 = gz[nz - 3] - gz[nz - 1], d1 = gz[nz - 2] - gz[nz - 1];
                L.sl_z1 = (1 / d2) / (1 / d2 + 1 / d1);
            }
        } else {
            L.vz_i0 = 1, L.vz_i1 = nz - 2, L.vz_j0 = 0, L.vz_j1 = nxx - 2;
            L.vx_i0 = 0, L.vx_i1 = nz - 2, L.vx_j0 = 1, L.vx_j1 = nxx - 2;
            L.ns_z0 = op->bc[0] == PLB_BC_NOSLIP, L.ns_z1 = op->bc[2] == PLB_BC_NOSLIP;
        }
        if (level_metrics(ctx, L)) return 2;
        if (l > 0) {
            // coefficient fields are full-size on every rank (coarsened redundantly: no communication)
            if (zalloc(ctx, &L.etas_own, L.full) || zalloc(ctx, &L.etan_own, L.full)) return 2;
            L.etas = L.etas_own, L.etan = L.etan_own;
        }
        if (zalloc(ctx, &L.X, 2 * L.plane) || zalloc(ctx, &L.T, 2 * L.plane) || zalloc(ctx, &L.b, 2 * L.plane) ||
            zalloc(ctx, &L.r, 2 * L.plane) || zalloc(ctx, &L.d, 2 * L.plane))
            return 2;
        op->lv.push_back(L);
    }
    return 0;
}

int n_unknowns(const Level& L) {
    return (L.vz_i1 - L.vz_i0 + 1) * (L.vz_j1 - L.vz_j0 + 1) + (L.vx_i1 - L.vx_i0 + 1) * (L.vx_j1 - L.vx_j0 + 1);
}

// nu Chebyshev steps on level l.  `from_zero`: the iterate is zero on entry.  Writes alternate
// between the two buffers; returns the buffer holding the result (halo rows exchanged).
double* smooth(plb_stokes* op, int l, const double* b, double* cur, double* other, bool from_zero, int nu) {
    Level& L = op->lv[l];
    plb_ctx* ctx = op->ctx;
    const LevelDev D = L.dev();
    const size_t P = L.plane;
    const double lmax = L.lmax, lmin = lmax / op->cheb_ratio;
    const double theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin), sigma = theta / delta;
    double rho = 1.0 / sigma;
    for (int k = 0; k < nu; k++) {
        double cd, cr;
        if (k == 0) cd = 0, cr = 1.0 / theta;
        else {
            double rn = 1.0 / (2 * sigma - rho);
            cd = rn * rho, cr = 2 * rn / delta, rho = rn;
        }
        {
            plb_prof_scope prof_(ctx, l == 0 ? PLB_K_CHEB0 : -1, ((k == 0 && from_zero) ? 64.0 : 96.0) * (double)P);
            if (k == 0 && from_zero) {
                // result goes to `cur` (no input needed)
                k_cheb<true><<<L.grid(), block2d(), 0, ctx->stream>>>(D, nullptr, nullptr, L.sh(b), L.sh(b + P), L.sh(L.d),
                                                                      L.sh(L.d + P), L.sh(cur), L.sh(cur + P), cd, cr);
            } else {
                if (op->tile_smoother && L.nxx >= 1024) {
                    const dim3 tg((L.nxx + TLX - 1) / TLX, (L.i1 - L.i0 + TLZ - 1) / TLZ);
                    k_cheb_tile<<<tg, 256, 0, ctx->stream>>>(D, L.lo, L.hi, L.sh(cur), L.sh(cur + P), L.sh(b), L.sh(b + P),
                                                             L.sh(L.d), L.sh(L.d + P), L.sh(other), L.sh(other + P), cd,
                                                             cr);
                } else {
                    k_cheb<false><<<L.grid(), block2d(), 0, ctx->stream>>>(D, L.sh(cur), L.sh(cur + P), L.sh(b),
                                                                           L.sh(b + P), L.sh(L.d), L.sh(L.d + P),
                                                                           L.sh(other), L.sh(other + P), cd, cr);
                }
                std::swap(cur, other);
            }
            ctx->launches++;
        }
        if (halo(op, L, cur, 2)) return nullptr;
    }
    return cur;
}

// one V-cycle for K x = b on level l; the result lands in `xout` (2 local planes, halos valid).
// Buffers alternate so that the (2 nu + 1)-th write hits xout.
int vcycle(plb_stokes* op, int l, const double* b, double* xout) {
    Level& L = op->lv[l];
    plb_ctx* ctx = op->ctx;
    const LevelDev D = L.dev();
    const size_t P = L.plane;
    const int nlev = (int)op->lv.size();
    if (l == nlev - 1) {
        if (op->cinv) {
            k_dense_solve<<<(op->nc * 32 + 255) / 256, 256, 0, ctx->stream>>>(D, op->nc, op->cinv, b, b + P,
                                                                             xout, xout + P);
            PLB_LAUNCHED(ctx);
        } else {
            // large coarsest level (odd cell counts): many smoothing steps instead of a direct solve
            int nu = op->nu_coarse | 1;          // odd: lands in xout
            double* res = smooth(op, l, b, xout, L.T, true, nu);
            if (res != xout) PLB_FAIL(ctx, "internal: coarse smoother parity");
            PLB_CUDA(ctx, cudaGetLastError());
        }
        return 0;
    }
    double* other = L.T;
    // write 1 -> xout, 2 -> T, 3 -> xout ...: after nu writes the iterate is in (nu odd ? xout : T)
    double* cur = smooth(op, l, b, xout, other, true, op->nu);
    if (!cur) return 2;
    double* oth = (cur == xout) ? other : xout;
    Level& Cl = op->lv[l + 1];
    LevelDev DC = Cl.dev();
    {
        plb_prof_scope prof_(ctx, l == 0 ? PLB_K_MGXFER0 : -1, 84.0 * (double)P);
        k_vel_op<0><<<L.grid(), block2d(), 0, ctx->stream>>>(D, L.sh(cur), L.sh(cur + P), L.sh(b), L.sh(b + P),
                                                            L.sh(L.r), L.sh(L.r + P));
        PLB_LAUNCHED(ctx);
        if (halo(op, L, L.r, 2)) return 2;
        if (L.dist && !Cl.dist) {
            // transition to the replicated levels: every rank restricts the coarse rows under its own
            // fine rows into the full-size coarse right-hand side, then the pieces are summed
            PLB_CUDA(ctx, cudaMemsetAsync(Cl.b, 0, sizeof(double) * 2 * Cl.plane, ctx->stream));
            DC.i0 = L.i0 / 2, DC.i1 = (L.i1 + 1) / 2;
        }
        k_restrict<<<grid2d(DC.i1 - DC.i0, Cl.nxx), block2d(), 0, ctx->stream>>>(D, DC, L.sh(L.r), L.sh(L.r + P),
                                                                              Cl.sh(Cl.b), Cl.sh(Cl.b + Cl.plane));
        PLB_LAUNCHED(ctx);
        if (L.dist && !Cl.dist && plb_comm_allreduce(ctx, Cl.b, 2 * Cl.plane, PLB_OP_SUM)) return 2;
    }
    {
        plb_prof_scope prof_(ctx, l == 0 ? PLB_K_MGCOARSE : -1);
        if (op->graph_exec && l + 1 == op->graph_level) {
            PLB_CUDA(ctx, cudaGraphLaunch(op->graph_exec, ctx->stream));
            ctx->launches += op->graph_launches;
        } else if (vcycle(op, l + 1, Cl.b, Cl.X)) {
            return 2;
        }
    }
    {
        plb_prof_scope prof_(ctx, l == 0 ? PLB_K_MGXFER0 : -1, 36.0 * (double)P);
        k_prolong_add<<<L.grid(), block2d(), 0, ctx->stream>>>(D, Cl.dev(), Cl.sh(Cl.X), Cl.sh(Cl.X + Cl.plane),
                                                              L.sh(cur), L.sh(cur + P), L.sh(oth), L.sh(oth + P));
        PLB_LAUNCHED(ctx);
        if (halo(op, L, oth, 2)) return 2;
    }
    std::swap(cur, oth);
    cur = smooth(op, l, b, cur, oth, false, op->nu);
    if (!cur) return 2;
    PLB_CUDA(ctx, cudaGetLastError());
    if (cur != xout) PLB_FAIL(ctx, "internal: V-cycle buffer parity");
    return 0;
}

int setup_hierarchy(plb_stokes* op) {
    plb_ctx* ctx = op->ctx;
    const int nlev = (int)op->lv.size();
    // coarse viscosities.  Replicated coefficient fields: full grids, computed redundantly on every rank.
    // Slab-local fields (plb_ctx_set_slab): a rank coarsens the rows of its own slab -- coarse row I needs the fine
    // rows 2I-1 .. 2I+2, i.e. one halo row either side -- and then exchanges one halo row of the coarse fields;
    // at the first replicated level every rank contributes its rows and the (small) planes are summed.
    for (int l = 1; l < nlev; l++) {
        Level &F = op->lv[l - 1], &Cc = op->lv[l];
        int I0 = 0, I1 = Cc.nz;
        const bool slabf = op->slab_fields && F.dist;
        if (slabf) {
            I0 = F.i0 / 2, I1 = (F.i1 + 1) / 2;
            if (!Cc.dist) {
                PLB_CUDA(ctx, cudaMemsetAsync(Cc.etas_own, 0, sizeof(double) * Cc.full, ctx->stream));
                PLB_CUDA(ctx, cudaMemsetAsync(Cc.etan_own, 0, sizeof(double) * Cc.full, ctx->stream));
            }
        }
        k_coarsen_eta<<<grid2d(I1 - I0, Cc.nxx), block2d(), 0, ctx->stream>>>(
            F.nz, F.nxx, F.ld, F.etas, F.etan, Cc.nz, Cc.nxx, Cc.ld, Cc.etas_own, Cc.etan_own, op->coarsen_wide, I0, I1);
        PLB_LAUNCHED(ctx);
        if (slabf && Cc.dist) {
            double* arrs[2] = {Cc.etas_own, Cc.etan_own};
            const long long rd[2] = {Cc.ld, Cc.ld};
            if (plb_comm_halo_rows(ctx, 2, arrs, rd, Cc.i0, Cc.i1, 1)) return 2;
        } else if (slabf) {
            if (plb_comm_allreduce(ctx, Cc.etas_own, Cc.full, PLB_OP_SUM) || plb_comm_allreduce(ctx, Cc.etan_own, Cc.full, PLB_OP_SUM))
                return 2;
        }
    }
    // largest eigenvalue of D^-1 K per level by power iteration.  The estimate (with its 10 % safety
    // margin) is reused for `lmax_every` consecutive coefficient updates: the viscosity field of a
    // time-stepping run changes by a fraction of a cell per step and the Chebyshev interval only
    // needs an upper bound.
    const int npow = 12;
    const bool fresh = op->lmax_age < 0 || op->lmax_age >= op->lmax_every;
    op->lmax_age = fresh ? 1 : op->lmax_age + 1;
    for (int l = 0; l < nlev && fresh; l++) {
        Level& L = op->lv[l];
        const LevelDev D = L.dev();
        const size_t P = L.plane;
        double *x = L.X, *y = L.T;
        reduce_shape(op, L, 2);
        PLB_CUDA(ctx, cudaMemsetAsync(x, 0, sizeof(double) * 2 * P, ctx->stream));
        PLB_CUDA(ctx, cudaMemsetAsync(y, 0, sizeof(double) * 2 * P, ctx->stream));
        k_fill_random<<<L.grid(), block2d(), 0, ctx->stream>>>(D, L.sh(x), L.sh(x + P));
        PLB_LAUNCHED(ctx);
        if (halo(op, L, x, 2)) return 2;
        double* s = op->d_scal + 900;
        for (int it = 0; it <= npow; it++) {
            k_vel_op<1><<<L.grid(), block2d(), 0, ctx->stream>>>(D, L.sh(x), L.sh(x + P), nullptr, nullptr, L.sh(y),
                                                                L.sh(y + P));
            PLB_LAUNCHED(ctx);
            if (plb_dot(ctx, &op->rws, 2 * P, y, y, s)) return 2;
            if (it == npow) break;          // last application: Rayleigh-type estimate below
            if (plb_scale_rsqrt2(ctx, 2 * P, s, y, nullptr)) return 2;
            if (halo(op, L, y, 2)) return 2;
            std::swap(x, y);
        }
        // || D^-1 K x || / || x || (over the momentum rows and their slaves)
        if (plb_dot(ctx, &op->rws, 2 * P, x, x, s + 1)) return 2;
        double h[2];
        if (plb_read_scalars(ctx, s, 2, h)) return 2;
        L.lmax = 1.1 * sqrt(h[0] / h[1]);
        if (!(L.lmax > 0) || !(L.lmax < 1e6)) PLB_FAIL(ctx, "Stokes MG: bad eigenvalue estimate %g on level %d", L.lmax, l);
        PLB_CUDA(ctx, cudaMemsetAsync(L.X, 0, sizeof(double) * 2 * P, ctx->stream));
        PLB_CUDA(ctx, cudaMemsetAsync(L.T, 0, sizeof(double) * 2 * P, ctx->stream));
    }
    reduce_shape(op, op->lv[0], 3);
    // dense inverse on the coarsest level (replicated)
    Level& Lc = op->lv[nlev - 1];
    const int n = n_unknowns(Lc);
    if (op->cinv && op->nc != n) cudaFree(op->cinv), op->cinv = nullptr;     // (kept: the CUDA graph holds it)
    if (n > op->dense_max && op->cinv) cudaFree(op->cinv), op->cinv = nullptr;
    op->nc = 0;
    if (n <= op->dense_max) {
        const LevelDev D = Lc.dev();
        const size_t P = Lc.plane;
        // scratch stays resident (one allocation per operator, not per coefficient update)
        if (op->scratch_n != n) {
            if (op->probes) cudaFree(op->probes), op->probes = nullptr;
            if (op->gjM) cudaFree(op->gjM), op->gjM = nullptr;
            PLB_CUDA(ctx, cudaMalloc(&op->probes, sizeof(double) * (size_t)n * 2 * P));
            PLB_CUDA(ctx, cudaMalloc(&op->gjM, sizeof(double) * (size_t)n * 2 * n));
            op->scratch_n = n;
        }
        double *probes = op->probes, *M = op->gjM;
        int* fail = (int*)(op->d_scal + 940);        // read back with the right-hand-side norm of the next solve
        PLB_CUDA(ctx, cudaMemsetAsync(probes, 0, sizeof(double) * (size_t)n * 2 * P, ctx->stream));
        PLB_CUDA(ctx, cudaMemsetAsync(M, 0, sizeof(double) * (size_t)n * 2 * n, ctx->stream));
        PLB_CUDA(ctx, cudaMemsetAsync(fail, 0, sizeof(double), ctx->stream));
        if (!op->cinv) PLB_CUDA(ctx, cudaMalloc(&op->cinv, sizeof(double) * (size_t)n * n));
        k_probe_set<<<grid2d(Lc.nz, Lc.nxx), block2d(), 0, ctx->stream>>>(D, n, P, probes);
        PLB_LAUNCHED(ctx);
        dim3 g = grid2d(Lc.nz, Lc.nxx);
        g.z = n;
        k_probe_apply<<<g, block2d(), 0, ctx->stream>>>(D, n, P, probes, M);
        PLB_LAUNCHED(ctx);
        k_gauss_jordan<<<1, 1024, sizeof(double) * n, ctx->stream>>>(n, M, fail);
        PLB_LAUNCHED(ctx);
        k_extract_inverse<<<(int)(((size_t)n * n + 255) / 256), 256, 0, ctx->stream>>>(n, M, op->cinv);
        PLB_LAUNCHED(ctx);
        op->nc = n;
    }
    // capture the V-cycle of the small levels (<= 1025 rows, replicated: no NCCL calls inside) into a
    // CUDA graph: those levels are launch-latency-bound (~10 launches of a few microseconds each)
    // (kernel arguments only change with the eigenvalue estimates: re-capture only then)
    int lg = -1;
    for (int l = 1; l < nlev; l++)
        if (!op->lv[l].dist && op->lv[l].nz <= 1025) { lg = l; break; }
    const bool all = op->graph_all == 1 || (op->graph_all < 0 && !op->lv[0].dist);
    if (all) lg = 0;
    plb_stokes::GraphSig sig;
    sig.etas = lg == 0 ? op->lv[0].etas : nullptr, sig.etan = lg == 0 ? op->lv[0].etan : nullptr;   // levels >= 1 own theirs
    sig.cinv = op->cinv, sig.nu = op->nu, sig.nu_coarse = op->nu_coarse, sig.nc = op->nc;
    sig.tile = op->tile_smoother, sig.all = all ? 1 : 0, sig.level = lg, sig.cheb_ratio = op->cheb_ratio;
    for (int l = 0; l < nlev; l++) sig.lmax.push_back(op->lv[l].lmax);
    if (op->graph_exec && (!op->use_graph || !(sig == op->graph_sig)))
        cudaGraphExecDestroy(op->graph_exec), op->graph_exec = nullptr;
    if (!op->graph_exec) op->graph_level = -1;
    if (op->use_graph && !op->graph_exec) {
        // Default on one GPU: record the WHOLE cycle (~70 launches) -- small and medium grids are bound by
        // the host's enqueue rate (2049^2: 67 -> 53 ms per solve).  With slabs the cycle also contains
        // ~35 grouped NCCL send/recv calls; capturing them works ("graph_all" = 1, tested on 2/4/8 GPUs)
        // but brings nothing there (the NCCL kernels' own latency on the device timeline is the bound),
        // so slab solves keep the graph for the replicated small levels only.
        if (all && !op->zv && zalloc(ctx, &op->zv, 2 * op->lv[0].plane)) return 2;
        if (lg >= 0 && lg < nlev) {
            Level& Lg = op->lv[lg];
            double* gout = lg == 0 ? op->zv : Lg.X;
            if (lg == 0 && vcycle(op, 0, Lg.b, gout)) return 2;      // warm-up outside the capture (NCCL set-up)
            const bool prof_was = ctx->prof_on;
            ctx->prof_on = false;                                      // no event records inside the graph
            cudaGraph_t graph = nullptr;
            const long long before = ctx->launches;
            PLB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            // the context usually runs on the legacy default stream (torch's), which cannot capture:
            // record on a private stream, replay on the context's
            cudaStream_t run_stream = ctx->stream, cap = nullptr;
            PLB_CUDA(ctx, cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
            ctx->stream = cap;
            cudaError_t e = cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal);
            int rc = e == cudaSuccess ? vcycle(op, lg, Lg.b, gout) : 2;
            if (e == cudaSuccess) e = cudaStreamEndCapture(cap, &graph);
            ctx->stream = run_stream;
            ctx->prof_on = prof_was;
            cudaStreamDestroy(cap);
            if (rc || e != cudaSuccess || !graph) {
                if (graph) cudaGraphDestroy(graph);
                PLB_FAIL(ctx, "Stokes MG: CUDA graph capture of the coarse V-cycle failed (%s)", cudaGetErrorString(e));
            }
            op->graph_launches = (int)(ctx->launches - before);
            ctx->launches = before;
            PLB_CUDA(ctx, cudaGraphInstantiate(&op->graph_exec, graph, 0));
            cudaGraphDestroy(graph);
            op->graph_level = lg;
            op->graph_sig = sig;
        }
    }
    op->hierarchy = true;
    return 0;
}

int ensure_krylov(plb_stokes* op) {
    plb_ctx* ctx = op->ctx;
    const size_t n3 = 3 * op->lv[0].plane;
    if (!op->xs) {
        if (zalloc(ctx, &op->xs, n3) || zalloc(ctx, &op->r3, n3) || zalloc(ctx, &op->b3, n3) ||
            zalloc(ctx, &op->t3, n3))
            return 2;
    }
    if (op->kry.m != op->gcr_m && plb_fgmres_alloc(ctx, &op->kry, op->gcr_m, (long long)n3, op->d_scal)) return 2;
    return 0;
}

}  // namespace

extern "C" {

int plb_stokes_create(plb_ctx* ctx, int nz, int nxx, int ld, const double* h_grid_z,
                      const double* h_grid_x, const int* h_bc, plb_stokes** out) {
    if (!ctx || !out) return 1;
    *out = nullptr;
    if (nz < 6 || nxx < 6) PLB_FAIL(ctx, "plb_stokes_create: grid %dx%d too small (anchor cell (3,2) must exist)", nz, nxx);
    if (ld != nxx) PLB_FAIL(ctx, "plb_stokes_create: only ld == nxx is supported");
    for (int w = 0; w < 4; w++) {
        int b = h_bc[w];
        bool zwall = (w % 2) == 0;
        if (zwall && b != PLB_BC_NOSLIP && b != PLB_BC_FREESLIP)
            PLB_FAIL(ctx, "plb_stokes_create: z-wall BC %d not supported (NOSLIP|FREESLIP; CYCLIC is SURVEY 8f-4)", b);
        if (!zwall && b != PLB_BC_FREESLIP && !(w == 1 && b == (PLB_BC_FREESLIP | PLB_BC_FLOWTHRU)))
            PLB_FAIL(ctx, "plb_stokes_create: x-wall BC %d not supported (FREESLIP, or FLOWTHRU|FREESLIP on the x = 0 wall; "
                          "the reference's own matrix is singular for NOSLIP, CYCLIC and pure FLOWTHRU x-walls and has "
                          "no pressure anchor with a flow-through x = L wall: tests/test_reference_bc_probe.py)", b);
    }
    PLB_CUDA(ctx, cudaSetDevice(ctx->device));
    plb_stokes* op = new plb_stokes();
    op->ctx = ctx, op->device = ctx->device, op->nz = nz, op->nxx = nxx, op->ld = ld;
    for (int w = 0; w < 4; w++) op->bc[w] = h_bc[w];
    if (op->bc[1] & PLB_BC_FLOWTHRU) op->ai = nz / 2, op->aj = 0;      // pylamp_stokes.py:539-541
    if (plb_reduce_ws_init(ctx, &op->rws)) { delete op; return 2; }
    if (zalloc(ctx, &op->d_scal, 1024 + 2 * (size_t)nz)) { delete op; return 2; }
    if (build_levels(op, h_grid_z, h_grid_x)) { plb_stokes_destroy(op); return 2; }
    std::vector<double> gz(h_grid_z, h_grid_z + nz), gx(h_grid_x, h_grid_x + nxx);
    if (upload(ctx, gz, &op->gz_d) || upload(ctx, gx, &op->gx_d)) { plb_stokes_destroy(op); return 2; }
    *out = op;
    return 0;
}

void plb_stokes_destroy(plb_stokes* op) {
    if (!op) return;
    // (does not touch the context: the host side may release it first at interpreter shutdown)
    cudaSetDevice(op->device);
    cudaDeviceSynchronize();
    for (Level& L : op->lv) free_level(L);
    plb_fgmres_free(&op->kry);
    if (op->graph_exec) cudaGraphExecDestroy(op->graph_exec);
    double* ptrs[] = {op->d_scal, op->cinv, op->probes, op->gjM, op->xs, op->r3, op->b3, op->t3, op->gz_d, op->gx_d, op->zv, op->surf_d, op->wsc_d};
    for (double* p : ptrs) if (p) cudaFree(p);
    for (double* p : op->hist) if (p) cudaFree(p);
    plb_reduce_ws_free(&op->rws);
    delete op;
}

int plb_stokes_set_param(plb_stokes* op, const char* name, double value) {
    if (!op || !name) return 1;
    plb_ctx* ctx = op->ctx;
    if (!strcmp(name, "nu")) op->nu = (int)value, op->hierarchy = false;
    else if (!strcmp(name, "gcr_m")) {
        if (value < 2 || value > 800) PLB_FAIL(ctx, "plb_stokes_set_param: gcr_m out of range 2..800");
        op->gcr_m = (int)value;
    }
    else if (!strcmp(name, "coarsen_wide")) op->coarsen_wide = (int)value, op->hierarchy = false;
    else if (!strcmp(name, "cheb_ratio")) op->cheb_ratio = value, op->hierarchy = false;
    else if (!strcmp(name, "dense_max")) op->dense_max = (int)value, op->hierarchy = false;
    else if (!strcmp(name, "nu_coarse")) op->nu_coarse = (int)value, op->hierarchy = false;
    else if (!strcmp(name, "reorth")) op->reorth = (int)value;
    else if (!strcmp(name, "rtol_accept")) op->rtol_accept = value;
    else if (!strcmp(name, "hydrostatic")) op->hydrostatic = (int)value;
    else if (!strcmp(name, "warm_start")) op->warm_start = (int)value;
    else if (!strcmp(name, "debug_halo")) op->debug_halo = (int)value;
    else if (!strcmp(name, "tile_smoother")) op->tile_smoother = (int)value, op->hierarchy = false, op->lmax_age = -1;
    else if (!strcmp(name, "lmax_every")) op->lmax_every = (int)value;
    else if (!strcmp(name, "use_graph")) op->use_graph = (int)value, op->hierarchy = false, op->lmax_age = -1;
    else if (!strcmp(name, "graph_all")) op->graph_all = (int)value, op->hierarchy = false, op->lmax_age = -1;
    else if (!strcmp(name, "reorth_thresh")) op->kry_reorth = value;
    else PLB_FAIL(ctx, "plb_stokes_set_param: unknown parameter '%s'", name);
    return 0;
}

int plb_stokes_set_coeffs(plb_stokes* op, const double* d_etas, const double* d_etan,
                          const double* d_rho, double gz, double gx) {
    if (!op) return 1;
    plb_ctx* ctx = op->ctx;
    PLB_CUDA(ctx, cudaSetDevice(ctx->device));
    Level& L = op->lv[0];
    L.etas = d_etas, L.etan = d_etan;
    op->rho = d_rho, op->g_z = gz, op->g_x = gx;
    op->surf = 0;      // the stabilisation planes belong to the previous density field: off until set again
    op->floor_est = 0; // a residual floor belongs to the system it was met on
    // mineta over both fields INCLUDING the ghost row/column of etan, like np.min (:116)
    double* d = op->d_scal + 920;
    k_set1<<<1, 1, 0, ctx->stream>>>(d, INFINITY);
    PLB_LAUNCHED(ctx);
    op->slab_fields = ctx->slab_on && L.dist;
    if (op->slab_fields) {
        if (ctx->slab_i0 != L.i0 || ctx->slab_i1 != L.i1)
            PLB_FAIL(ctx, "plb_stokes_set_coeffs: the context's slab rows [%d, %d) are not the solver's [%d, %d)", ctx->slab_i0,
                     ctx->slab_i1, L.i0, L.i1);
        // the own rows of both fields (the last rank's include the ghost row), then the minimum over the ranks
        const long long off = (long long)L.i0 * L.ld, cnt = (long long)(L.i1 - L.i0) * L.ld;
        k_min2<<<plb_grid_for(ctx, cnt, 256, 8), 256, 0, ctx->stream>>>(cnt, d_etas + off, d_etan + off, d);
        PLB_LAUNCHED(ctx);
        if (plb_comm_allreduce(ctx, d, 1, PLB_OP_MIN)) return 2;
    } else {
        k_min2<<<plb_grid_for(ctx, (long long)L.full, 256, 8), 256, 0, ctx->stream>>>((long long)L.full, d_etas, d_etan, d);
        PLB_LAUNCHED(ctx);
    }
    double mineta;
    if (plb_read_scalars(ctx, d, 1, &mineta)) return 2;
    // avgd = L/n with n = number of NODES (sic), pylamp_stokes.py:119-120
    double avgdx = (L.gx[op->nxx - 1] - L.gx[0]) / op->nxx;
    double avgdz = (L.gz[op->nz - 1] - L.gz[0]) / op->nz;
    op->Kc = 2 * mineta / (avgdx + avgdz);
    op->Kb = 4 * mineta / ((avgdx + avgdz) * (avgdx + avgdz));
    op->coeffs = true, op->hierarchy = false;
    // scalings of the preconditioner's first stage (own rows; the pressure scaling also on the halo row below)
    if (!op->wsc_d && zalloc(ctx, &op->wsc_d, 3 * L.full)) return 2;
    {
        const int row0 = L.i0 > 0 ? L.i0 - 1 : 0;
        const dim3 g((L.nxx + BX - 1) / BX, (L.i1 - row0 + BY - 1) / BY);
        k_scale_planes<<<g, block2d(), 0, ctx->stream>>>(L.dev(), row0, 1.0 / op->Kc, op->wsc_d, op->wsc_d + L.full,
                                                       op->wsc_d + 2 * L.full);
        PLB_LAUNCHED(ctx);
    }
    return 0;
}

int plb_stokes_set_surfstab(plb_stokes* op, double theta_dt) {
    if (!op) return 1;
    plb_ctx* ctx = op->ctx;
    if (!op->coeffs) PLB_FAIL(ctx, "plb_stokes_set_surfstab: coefficients not set");
    op->floor_est = 0;
    if (!(theta_dt > 0)) {
        op->surf = 0;
        return 0;
    }
    PLB_CUDA(ctx, cudaSetDevice(ctx->device));
    Level& L = op->lv[0];
    if (!op->surf_d && zalloc(ctx, &op->surf_d, 2 * L.full)) return 2;
    k_surfstab_planes<<<grid2d(L.nz, L.nxx), block2d(), 0, ctx->stream>>>(L.dev(), op->rho, op->surf_d,
                                                                        op->surf_d + L.full);
    PLB_LAUNCHED(ctx);
    op->surf = theta_dt;
    return 0;
}

int plb_stokes_scaling(plb_stokes* op, double* h_out) {
    if (!op || !op->coeffs) return 1;
    h_out[0] = op->Kc, h_out[1] = op->Kb;
    return 0;
}

int plb_stokes_rhs(plb_stokes* op, double* d_rhs) {
    if (!op) return 1;
    plb_ctx* ctx = op->ctx;
    if (!op->coeffs) PLB_FAIL(ctx, "plb_stokes_rhs: coefficients not set");
    if (op->lv[0].dist) PLB_FAIL(ctx, "plb_stokes_rhs: not available on a slab-distributed operator");
    PLB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ensure_krylov(op)) return 2;
    Level& L = op->lv[0];
    const size_t P = L.plane;
    k_stokes_rhs<<<grid2d(L.nz, L.nxx), block2d(), 0, ctx->stream>>>(L.dev(), op->rho, op->g_z, op->g_x,
                                                                   op->t3, op->t3 + P, op->t3 + 2 * P);
    PLB_LAUNCHED(ctx);
    k_interleave<<<plb_grid_for(ctx, (long long)P, 256, 8), 256, 0, ctx->stream>>>((long long)P, op->t3, op->t3 + P,
                                                                                op->t3 + 2 * P, d_rhs);
    PLB_LAUNCHED(ctx);
    return 0;
}

int plb_stokes_apply(plb_stokes* op, const double* d_x, double* d_y) {
    if (!op) return 1;
    plb_ctx* ctx = op->ctx;
    if (!op->coeffs) PLB_FAIL(ctx, "plb_stokes_apply: coefficients not set");
    if (op->lv[0].dist) PLB_FAIL(ctx, "plb_stokes_apply: not available on a slab-distributed operator");
    PLB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ensure_krylov(op)) return 2;
    Level& L = op->lv[0];
    const size_t P = L.plane;
    double *in = op->t3, *outp = op->r3;
    k_deinterleave<<<plb_grid_for(ctx, (long long)P, 256, 8), 256, 0, ctx->stream>>>((long long)P, d_x, in, in + P, in + 2 * P);
    PLB_LAUNCHED(ctx);
    FullArgs a = {op->gz_d, op->gx_d, op->bc[0], op->bc[2], op->Kc, op->Kb, op->lv[0].ft_x0, op->ai, op->aj};
    const SurfArgs sf = surf_args(op);
    if (op->surf > 0)
        k_stokes_full<true><<<grid2d(L.nz, L.nxx), block2d(), 0, ctx->stream>>>(L.dev(), a, sf, in, in + P, in + 2 * P,
                                                                              outp, outp + P, outp + 2 * P);
    else
        k_stokes_full<false><<<grid2d(L.nz, L.nxx), block2d(), 0, ctx->stream>>>(L.dev(), a, sf, in, in + P, in + 2 * P,
                                                                               outp, outp + P, outp + 2 * P);
    PLB_LAUNCHED(ctx);
    k_interleave<<<plb_grid_for(ctx, (long long)P, 256, 8), 256, 0, ctx->stream>>>((long long)P, outp, outp + P,
                                                                                outp + 2 * P, d_y);
    PLB_LAUNCHED(ctx);
    return 0;
}

// one multigrid V-cycle on the velocity block of level 0: test hook.  b, x: two full-size planes
// [vz | vx]; a slab rank reads its own rows of b and fills its own rows of x (others zero).
int plb_stokes_vcycle(plb_stokes* op, const double* d_b2, double* d_x2) {
    if (!op) return 1;
    plb_ctx* ctx = op->ctx;
    if (!op->coeffs) PLB_FAIL(ctx, "plb_stokes_vcycle: coefficients not set");
    PLB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ensure_krylov(op)) return 2;
    if (!op->hierarchy && setup_hierarchy(op)) return 2;
    Level& L = op->lv[0];
    const size_t P = L.plane, rows = (size_t)(L.i1 - L.i0) * L.ld, off_l = (size_t)(L.i0 - L.lo) * L.ld,
                 off_g = (size_t)L.i0 * L.ld;
    double* z = op->t3;
    PLB_CUDA(ctx, cudaMemsetAsync(z, 0, sizeof(double) * 2 * P, ctx->stream));
    PLB_CUDA(ctx, cudaMemsetAsync(L.b, 0, sizeof(double) * 2 * P, ctx->stream));
    for (int p = 0; p < 2; p++)
        PLB_CUDA(ctx, cudaMemcpyAsync(L.b + p * P + off_l, d_b2 + p * L.full + off_g, sizeof(double) * rows,
                                      cudaMemcpyDeviceToDevice, ctx->stream));
    if (vcycle(op, 0, L.b, z)) return 2;
    PLB_CUDA(ctx, cudaMemsetAsync(d_x2, 0, sizeof(double) * 2 * L.full, ctx->stream));
    for (int p = 0; p < 2; p++)
        PLB_CUDA(ctx, cudaMemcpyAsync(d_x2 + p * L.full + off_g, z + p * P + off_l, sizeof(double) * rows,
                                      cudaMemcpyDeviceToDevice, ctx->stream));
    return 0;
}

int plb_stokes_last_stats(plb_stokes* op, double* h_out) {
    if (!op) return 1;
    h_out[0] = op->last_iters, h_out[1] = op->last_vcycles, h_out[2] = op->last_relres, h_out[3] = op->floor_est;
    h_out[4] = op->last_status, h_out[5] = op->last_rtol_eff;
    return 0;
}

int plb_stokes_solve(plb_stokes* op, const double* d_rhs, double rtol, int maxit, double* d_x,
                     int* h_iters, double* h_relres) {
    if (!op) return 1;
    plb_ctx* ctx = op->ctx;
    if (!op->coeffs) PLB_FAIL(ctx, "plb_stokes_solve: coefficients not set");
    PLB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ensure_krylov(op)) return 2;
    if (!op->hierarchy && setup_hierarchy(op)) return 2;
    Level& L = op->lv[0];
    const LevelDev D = L.dev();
    const size_t P = L.plane;
    const double Kc = op->Kc;
    const SurfArgs sf = surf_args(op);
    double *x = op->xs, *r = op->r3, *b = op->b3;
    const dim3 g = L.grid(), blk = block2d();
    const dim3 tg((L.nxx + TLX - 1) / TLX, (L.i1 - L.i0 + TLZ - 1) / TLZ);     // tile-staged kernels
    // kernels index with global rows: shifted views of the local (slab) vectors
    auto V3 = [&](double* v, int pl) { return L.sh(v + (size_t)pl * P); };
    auto C3 = [&](const double* v, int pl) { return L.sh(v + (size_t)pl * P); };
    reduce_shape(op, L, 3);
    if (d_rhs) {
        // caller's right-hand side in the reference layout (full size); entries on wall/ghost/corner/
        // anchor rows are homogeneous in the reference (pylamp_stokes.py never writes them): ignored
        k_deinterleave_rows<<<g, blk, 0, ctx->stream>>>(D, d_rhs, V3(b, 0), V3(b, 1), V3(b, 2));
    } else {
        k_stokes_rhs<<<g, blk, 0, ctx->stream>>>(D, op->rho, op->g_z, op->g_x, V3(b, 0), V3(b, 1), V3(b, 2));
    }
    PLB_LAUNCHED(ctx);
    double* ph = nullptr;
    if (op->hydrostatic) {
        double* m = op->d_scal + 1024;        // [nz] row means, then [nz] P_h
        ph = m + L.nz;
        k_row_mean<<<L.nz, 256, 0, ctx->stream>>>(D, C3(b, 0), m);
        PLB_LAUNCHED(ctx);
        if (L.dist && plb_comm_allreduce(ctx, m, (size_t)L.nz, PLB_OP_SUM)) return 2;   // rows of other slabs
        k_scan_ph<<<1, 1, 0, ctx->stream>>>(D, Kc, m, ph, op->ai);
        PLB_LAUNCHED(ctx);
        k_sub_row_mean<<<g, blk, 0, ctx->stream>>>(D, m, V3(b, 0));
        PLB_LAUNCHED(ctx);
    }
    auto stokes_resid = [&](double* xx, double* out) -> int {
        if (halo(op, L, xx, 3)) return 2;
        plb_prof_scope prof_(ctx, PLB_K_STOKES_OP, 88.0 * (double)P);
        if (op->surf > 0) {      // free-surface stabilisation terms in the outer operator
            if (op->tile_smoother && L.nxx >= 1024)
                k_stokes_op_tile<true, true><<<tg, 256, 0, ctx->stream>>>(D, sf, L.lo, L.hi, Kc, C3(xx, 0), C3(xx, 1),
                                                                          C3(xx, 2), C3(b, 0), C3(b, 1), C3(b, 2),
                                                                          V3(out, 0), V3(out, 1), V3(out, 2));
            else
                k_stokes_op<true, true><<<g, blk, 0, ctx->stream>>>(D, sf, Kc, C3(xx, 0), C3(xx, 1), C3(xx, 2), C3(b, 0),
                                                                    C3(b, 1), C3(b, 2), V3(out, 0), V3(out, 1), V3(out, 2));
        } else if (op->tile_smoother && L.nxx >= 1024)
            k_stokes_op_tile<true, false><<<tg, 256, 0, ctx->stream>>>(D, sf, L.lo, L.hi, Kc, C3(xx, 0), C3(xx, 1), C3(xx, 2),
                                                                       C3(b, 0), C3(b, 1), C3(b, 2), V3(out, 0), V3(out, 1),
                                                                       V3(out, 2));
        else
            k_stokes_op<true, false><<<g, blk, 0, ctx->stream>>>(D, sf, Kc, C3(xx, 0), C3(xx, 1), C3(xx, 2), C3(b, 0), C3(b, 1),
                                                                 C3(b, 2), V3(out, 0), V3(out, 1), V3(out, 2));
        PLB_LAUNCHED(ctx);
        return 0;
    };
    auto residual = [&](double* out) -> int { return stokes_resid(x, out); };
    // || W b ||: the scaled norm of the right-hand side (W b is the residual of x = 0 and the
    // operator is linear, so evaluate it with a zero iterate in the scratch vector)
    PLB_CUDA(ctx, cudaMemsetAsync(op->t3, 0, sizeof(double) * 3 * P, ctx->stream));
    if (stokes_resid(op->t3, r)) return 2;
    double bn2;
    if (plb_dot(ctx, &op->rws, 3 * P, r, r, op->d_scal + 910)) return 2;
    {
        // one read-back: the norm and the failure flag of the coarse-level inverse built by setup_hierarchy
        double h[31];
        if (plb_read_scalars(ctx, op->d_scal + 910, 31, h)) return 2;
        bn2 = h[0];
        int hfail;
        memcpy(&hfail, &h[30], sizeof(int));
        if (hfail) PLB_FAIL(ctx, "Stokes MG: singular coarse-level operator");
    }
    const double bnorm = sqrt(bn2);
    // initial guess: the previous solve's iterate (deviation from its hydrostatic pressure) when
    // warm starts are enabled, otherwise zero
    if (!(op->warm_start && op->have_prev)) {
        PLB_CUDA(ctx, cudaMemsetAsync(x, 0, sizeof(double) * 3 * P, ctx->stream));
        op->nhist = 0;
    } else if (op->warm_start >= 2) {
        // keep the last (warm_start - 1) iterates besides the live one; newest first
        const int keep = std::min(op->warm_start, 6);
        if ((int)op->hist.size() < keep) op->hist.resize(keep, nullptr);
        double* recycled = op->hist[keep - 1];
        for (int j = keep - 1; j > 0; j--) op->hist[j] = op->hist[j - 1];
        op->hist[0] = recycled;
        if (!op->hist[0] && zalloc(ctx, &op->hist[0], 3 * P)) return 2;
        PLB_CUDA(ctx, cudaMemcpyAsync(op->hist[0], x, sizeof(double) * 3 * P, cudaMemcpyDeviceToDevice, ctx->stream));
        op->nhist = std::min(op->nhist + 1, keep);
        if (op->nhist >= 2) {
            HistList H;
            H.p = op->nhist;
            double binom = 1;                                  // C(p, j+1), alternating sign
            for (int j = 0; j < 6; j++) {
                if (j < H.p) binom = binom * (H.p - j) / (j + 1);
                H.h[j] = j < H.p ? op->hist[j] : nullptr;
                H.c[j] = j < H.p ? ((j & 1) ? -binom : binom) : 0.0;
            }
            k_extrapolate<<<plb_grid_for(ctx, (long long)(3 * P), 256, 8), 256, 0, ctx->stream>>>((long long)(3 * P), H, x);
            PLB_LAUNCHED(ctx);
        }
    }
    int vcycles = 0;
    auto apply = [&](const double* z, double* c) -> int {
        // z comes out of `precond` with valid halo rows
        if (op->debug_halo && halo(op, L, const_cast<double*>(z), 3)) return 2;
        plb_prof_scope prof_(ctx, PLB_K_STOKES_OP, 64.0 * (double)P);
        if (op->surf > 0) {
            if (op->tile_smoother && L.nxx >= 1024)
                k_stokes_op_tile<false, true><<<tg, 256, 0, ctx->stream>>>(D, sf, L.lo, L.hi, Kc, C3(z, 0), C3(z, 1), C3(z, 2),
                                                                           nullptr, nullptr, nullptr, V3(c, 0), V3(c, 1),
                                                                           V3(c, 2));
            else
                k_stokes_op<false, true><<<g, blk, 0, ctx->stream>>>(D, sf, Kc, C3(z, 0), C3(z, 1), C3(z, 2), nullptr, nullptr,
                                                                     nullptr, V3(c, 0), V3(c, 1), V3(c, 2));
        } else if (op->tile_smoother && L.nxx >= 1024)
            k_stokes_op_tile<false, false><<<tg, 256, 0, ctx->stream>>>(D, sf, L.lo, L.hi, Kc, C3(z, 0), C3(z, 1), C3(z, 2), nullptr,
                                                                        nullptr, nullptr, V3(c, 0), V3(c, 1), V3(c, 2));
        else
            k_stokes_op<false, false><<<g, blk, 0, ctx->stream>>>(D, sf, Kc, C3(z, 0), C3(z, 1), C3(z, 2), nullptr, nullptr,
                                                                  nullptr, V3(c, 0), V3(c, 1), V3(c, 2));
        PLB_LAUNCHED(ctx);
        return 0;
    };
    auto precond = [&](const double* rr, double* z) -> int {
        // the pressure residual of the row below is read across the slab boundary
        if (halo(op, L, const_cast<double*>(rr) + 2 * P, 1)) return 2;
        {
            plb_prof_scope prof_(ctx, PLB_K_PRECRHS, 80.0 * (double)P);
            k_precond_rhs<<<g, blk, 0, ctx->stream>>>(D, Kc, op->wsc_d, op->wsc_d + L.full, op->wsc_d + 2 * L.full, C3(rr, 0),
                                                      C3(rr, 1), C3(rr, 2), V3(z, 2), L.sh(L.b), L.sh(L.b + P));
            PLB_LAUNCHED(ctx);
            PLB_CUDA(ctx, cudaMemsetAsync(z, 0, sizeof(double) * 2 * P, ctx->stream));
        }
        vcycles++;
        if (op->graph_exec && op->graph_level == 0) {
            plb_prof_scope prof_(ctx, PLB_K_MGCOARSE);            // (whole cycle: no per-kernel classes)
            PLB_CUDA(ctx, cudaGraphLaunch(op->graph_exec, ctx->stream));
            ctx->launches += op->graph_launches;
            PLB_CUDA(ctx, cudaMemcpyAsync(z, op->zv, sizeof(double) * 2 * P, cudaMemcpyDeviceToDevice, ctx->stream));
        } else if (vcycle(op, 0, L.b, z)) {
            return 2;
        }
        reduce_shape(op, L, 3);                      // the V-cycle set per-level shapes
        return halo(op, L, z + 2 * P, 1);            // velocity halos are valid after the last sweep
    };
    plb_fgmres_result res;
    op->kry.reorth_thresh = op->kry_reorth;
    // a floor met by an earlier solve on these SAME coefficient fields (re-solves of the surfstab loop,
    // repeated solves of one system) bounds what this one can reach; new coefficients reset it
    const double rtol_eff = std::max(rtol, 3 * op->floor_est);
    if (bnorm > 0) {
        if (plb_fgmres(ctx, &op->rws, &op->kry, residual, apply, precond, x, bnorm, rtol_eff, maxit, &res)) return 2;
        if (res.floor > 0 && res.floor <= op->rtol_accept) op->floor_est = res.floor;
    } else {
        res.converged = true;
    }
    const int total = res.iters;
    op->last_iters = total, op->last_vcycles = vcycles, op->last_relres = res.relres;
    op->last_rtol_eff = rtol_eff;
    op->last_status = (res.converged && res.relres <= 1.5 * rtol) ? 0 : (res.relres <= op->rtol_accept ? 1 : 2);
    op->have_prev = res.converged || res.relres <= op->rtol_accept;
    if (h_iters) *h_iters = total;
    if (h_relres) *h_relres = res.relres;
    // full-size interleaved solution; a slab rank fills its own rows and leaves the rest zero (the
    // host side sums the pieces, e.g. with plb_allreduce)
    // host side sums the pieces, e.g. with plb_allreduce); with slab-local fields only the own rows are written
    // and halo rows exchanged with the neighbours below
    if (L.dist && !op->slab_fields) PLB_CUDA(ctx, cudaMemsetAsync(d_x, 0, sizeof(double) * 3 * L.full, ctx->stream));
    double* panchor = op->d_scal + 930;
    k_get_anchor<<<1, 1, 0, ctx->stream>>>(D, C3(x, 2), panchor, op->ai, op->aj);
    PLB_LAUNCHED(ctx);
    if (L.dist && plb_comm_allreduce(ctx, panchor, 1, PLB_OP_SUM)) return 2;
    k_solution_out<<<g, blk, 0, ctx->stream>>>(D, C3(x, 0), C3(x, 1), C3(x, 2), ph, panchor, d_x);
    PLB_LAUNCHED(ctx);
    if (op->slab_fields) {
        double* arrs[1] = {d_x};
        const long long rd[1] = {3LL * L.nxx};
        if (plb_comm_halo_rows(ctx, 1, arrs, rd, L.i0, L.i1, ctx->slab_halo)) return 2;
    }
    if (!res.converged && res.relres > op->rtol_accept)
        PLB_FAIL(ctx, "plb_stokes_solve: not converged after %d iterations (relres %.3e > rtol %.3e, "
                      "accept %.1e)", total, res.relres, rtol, op->rtol_accept);
    return 0;
}

}  // extern "C"
