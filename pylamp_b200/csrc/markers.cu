// markers.cu -- marker-in-cell kernels: marker->grid averaging, grid->marker interpolation,
// RK4 advection with Meyer-Jenny conservative velocity interpolation, fence, cell index/count,
// marker property update, subgrid-diffusion marker stages.
//
// Reference behaviour restated (never copied): pylamp_trac.py:30-158 (grid2trac), :161-318
// (trac2grid), :321-388 (RK); pylamp2.py:291-303, :471-476, :558-572, :588-593.
#include <algorithm>
#include <vector>

#include "comm.cuh"

namespace {

// floor((n-1)*(x-lmin)/len) with IEEE multiply then divide and no FMA contraction, so that the
// cell a marker falls into is bit-identical to NumPy's (pylamp_trac.py:46-47, :226-227;
// pylamp2.py:588-589).
__device__ __forceinline__ long long cell_of(double x, double lmin, double len, int n) {
    double t = __ddiv_rn(__dmul_rn((double)(n - 1), __dsub_rn(x, lmin)), len);
    return (long long)floor(t);
}

// ---------------------------------------------------------------------------------------------
// min/max of marker coordinates
// ---------------------------------------------------------------------------------------------
__global__ void k_minmax_init(double* out) {
    out[0] = out[2] = 1e300;
    out[1] = out[3] = -1e300;
}

__global__ void k_minmax_flip(double* out) {
    out[0] = -out[0];
    out[2] = -out[2];
}

__global__ void __launch_bounds__(256) k_marker_minmax(long long M, const double2* __restrict__ x,
                                                      double* out) {
    double zmin = 1e300, zmax = -1e300, xmin = 1e300, xmax = -1e300;
    for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < M;
         m += (long long)gridDim.x * blockDim.x) {
        double2 p = x[m];
        zmin = fmin(zmin, p.x), zmax = fmax(zmax, p.x);
        xmin = fmin(xmin, p.y), xmax = fmax(xmax, p.y);
    }
    zmin = warp_min(zmin), zmax = warp_max(zmax), xmin = warp_min(xmin), xmax = warp_max(xmax);
    __shared__ double s[4][8];
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) s[0][w] = zmin, s[1][w] = zmax, s[2][w] = xmin, s[3][w] = xmax;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; i++) {
            s[0][0] = fmin(s[0][0], s[0][i]), s[1][0] = fmax(s[1][0], s[1][i]);
            s[2][0] = fmin(s[2][0], s[2][i]), s[3][0] = fmax(s[3][0], s[3][i]);
        }
        atomic_min_double(out + 0, s[0][0]), atomic_max_double(out + 1, s[1][0]);
        atomic_min_double(out + 2, s[2][0]), atomic_max_double(out + 3, s[3][0]);
    }
}

// ---------------------------------------------------------------------------------------------
// trac2grid: scatter phase + finalise phase
// ---------------------------------------------------------------------------------------------
struct T2GArgs {
    const double* f[PLB_MAX_FIELDS];
    double* acc[PLB_MAX_FIELDS];   // accumulation planes nze x nxe
    double* out[PLB_MAX_FIELDS];
    int scheme[PLB_MAX_FIELDS];
    double* wsum;                  // sum of weights (weighted schemes)
    double* cnt;                   // marker count per node (unweighted schemes)
    const double* axz;
    const double* axx;
    const double* riz;             // 1/(axz[i+1]-axz[i])
    const double* rix;
    const double2* tz;             // {axz[i], riz[i]} (one 16-byte load per marker and axis)
    const double2* tx;
    int merge_first;               // chunk kernel: combine the first runs across lanes too (t2g_variant 2)
    int nze, nxe;
    double z0, zlen, x0, xlen;
    double sz, sx;                 // (nze-1)/zlen, (nxe-1)/xlen
    int k;
};

__global__ void k_axis_recip(int n, const double* __restrict__ ax, double* __restrict__ r,
                             double2* __restrict__ tab) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = (i < n - 1) ? 1.0 / (ax[i + 1] - ax[i]) : 0.0;
    r[i] = v;
    if (tab) tab[i] = make_double2(ax[i], v);
}

// (A shared-memory transpose variant of this reduction was measured slower: 50 vs 37 ms per 4096^2
// step -- 105 registers and 35 KB smem per block cost more occupancy than the shuffles cost issue slots.)
// Warp-aggregated scatter: contiguous lanes of a warp that fall in the same cell (the common
// case once the markers are cell-ordered) are combined by a segmented shuffle reduction, so
// only the first lane of each run issues the global fp64 reductions.  Correct for any order.
// `span` (warp-uniform) = longest remaining run length in lanes - 1: shuffle distances beyond it
// cannot contribute, so short runs (the cell-ordered case: 16 markers = 4 lanes) need 2 steps, not 5
__device__ __forceinline__ double seg_reduce(double v, int lane, int run_end, int span = 31) {
    for (int o = 1; o <= span; o <<= 1) {
        double t = __shfl_down_sync(0xffffffffu, v, o);
        if (lane + o <= run_end) v += t;
    }
    return v;
}

// per-marker contribution: cell, the four corner weights and the (transformed) field values
template <int K>
struct T2GMarker {
    long long cell;        // ie * nxe + je, or -1 if outside (cannot happen after the ghost extension)
    double w[4];
    double v[K];
};

template <int K>
__device__ __forceinline__ T2GMarker<K> t2g_load(const T2GArgs& a, const double2* __restrict__ trx, long long m) {
    T2GMarker<K> r;
    const double2 p = trx[m];
    // Cell lookup by multiplication: unlike the per-cell count (plb_cell_index_count), the node sums do
    // not depend on which of two cells a marker sitting exactly on a cell face is assigned to (its
    // weights towards the far nodes are 0 either way), so the exact mul-then-divide of the reference
    // is not needed here; fp64 division is what bounds this kernel.
    // (Unweighted schemes count markers per assigned cell, so they keep the exact lookup.)
    long long ie, je;
    if (a.cnt) {
        ie = cell_of(p.x, a.z0, a.zlen, a.nze);
        je = cell_of(p.y, a.x0, a.xlen, a.nxe);
    } else {
        ie = (long long)floor((p.x - a.z0) * a.sz);
        je = (long long)floor((p.y - a.x0) * a.sx);
        // a marker on the upper edge of the (extended) axis belongs to the last cell
        if (ie == a.nze - 1 && p.x <= a.axz[a.nze - 1]) ie = a.nze - 2;
        if (je == a.nxe - 1 && p.y <= a.axx[a.nxe - 1]) je = a.nxe - 2;
    }
    r.cell = -1;
#pragma unroll
    for (int c = 0; c < 4; c++) r.w[c] = 0;
#pragma unroll
    for (int f = 0; f < K; f++) r.v[f] = 0;
    if (ie < 0 || ie > a.nze - 2 || je < 0 || je > a.nxe - 2) return r;
    const double az = (p.x - a.axz[ie]) * a.riz[ie];   // pylamp_trac.py:247 (reciprocal spacing)
    const double ax = (p.y - a.axx[je]) * a.rix[je];
    const double bz = 1 - az, bx = 1 - ax;            // :249
    r.w[0] = (1 - ax) * (1 - az);                     // node (i  , j  )   :252
    r.w[1] = (1 - ax) * (1 - bz);                     // node (i+1, j  )
    r.w[2] = (1 - bx) * (1 - az);                     // node (i  , j+1)
    r.w[3] = (1 - bx) * (1 - bz);                     // node (i+1, j+1)
    r.cell = ie * a.nxe + je;
#pragma unroll
    for (int f = 0; f < K; f++) {
        const double val = a.f[f][m];
        r.v[f] = (a.scheme[f] & PLB_AVG_ARITHMETIC) ? val : log(val);
    }
    return r;
}

// add an aggregate (sums over some markers of ONE cell) straight to the planes
template <int K>
__device__ __forceinline__ void t2g_flush(const T2GArgs& a, long long cell, const double (&w)[4], double cntv,
                                          const double (&vw)[K][4]) {
    const long long idx[4] = {cell, cell + a.nxe, cell + 1, cell + a.nxe + 1};
#pragma unroll
    for (int c = 0; c < 4; c++) {
        if (a.wsum) atomicAdd(a.wsum + idx[c], w[c]);
        if (a.cnt) atomicAdd(a.cnt + idx[c], cntv);
    }
#pragma unroll
    for (int f = 0; f < K; f++)
#pragma unroll
        for (int c = 0; c < 4; c++)
            atomicAdd(a.acc[f] + idx[c], (a.scheme[f] & PLB_AVG_WEIGHTED) ? vw[f][c] : vw[f][0]);
}

// Each thread takes T2G_U consecutive markers and sums those of one cell in registers (while the
// cloud is cell-ordered that is all of them); the per-thread sums then go through the warp-level
// segmented reduction, so a run of markers of one cell costs one atomic per (node, quantity) and
// the shuffle count per marker drops by T2G_U.  When a thread's markers straddle cells, the sums
// of the earlier cells are added with plain atomics.  Correct for any marker order.
constexpr int T2G_U = 4;

template <int K>
__global__ void __launch_bounds__(256)
k_t2g_scatter(long long M, const double2* __restrict__ trx, T2GArgs a) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const long long nchunk = (M + T2G_U - 1) / T2G_U;             // one chunk per thread
    long long base = (blockIdx.x * (long long)blockDim.x + threadIdx.x) - lane;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; base < nchunk; base += stride) {
        const long long ch = base + lane;
        long long cell = -1 - lane;          // lanes without a valid aggregate form runs of their own
        double w[4] = {0, 0, 0, 0}, cntv = 0;
        double vw[K][4];                     // weighted: sum v*w[c]; unweighted: sum v in [f][0]
#pragma unroll
        for (int f = 0; f < K; f++)
#pragma unroll
            for (int c = 0; c < 4; c++) vw[f][c] = 0;
        bool valid = false;
        if (ch < nchunk) {
            const long long m0 = ch * T2G_U;
#pragma unroll
            for (int u = 0; u < T2G_U; u++) {
                if (m0 + u >= M) break;
                const T2GMarker<K> r = t2g_load<K>(a, trx, m0 + u);
                if (r.cell < 0) continue;
                if (valid && r.cell != cell) {           // cell change inside the chunk: flush, restart
                    t2g_flush<K>(a, cell, w, cntv, vw);
                    cntv = 0;
#pragma unroll
                    for (int c = 0; c < 4; c++) w[c] = 0;
#pragma unroll
                    for (int f = 0; f < K; f++)
#pragma unroll
                        for (int c = 0; c < 4; c++) vw[f][c] = 0;
                }
                valid = true;
                cell = r.cell;
                cntv += 1.0;
#pragma unroll
                for (int c = 0; c < 4; c++) w[c] += r.w[c];
#pragma unroll
                for (int f = 0; f < K; f++) {
                    if (a.scheme[f] & PLB_AVG_WEIGHTED) {
#pragma unroll
                        for (int c = 0; c < 4; c++) vw[f][c] += r.v[f] * r.w[c];
                    } else {
                        vw[f][0] += r.v[f];
                    }
                }
            }
        }
        const long long prev = __shfl_up_sync(full, cell, 1);
        const bool head = (lane == 0) || (prev != cell);
        const unsigned heads = __ballot_sync(full, head);
        const unsigned after = (lane == 31) ? 0u : (heads >> (lane + 1));
        const int run_end = after ? lane + __ffs(after) - 1 : 31;
        const int span = __reduce_max_sync(full, run_end - lane);
        const bool emit = head && valid;
        const long long idx[4] = {cell, cell + a.nxe, cell + 1, cell + a.nxe + 1};
        if (a.wsum) {
#pragma unroll
            for (int c = 0; c < 4; c++) {
                double s = seg_reduce(w[c], lane, run_end, span);
                if (emit) atomicAdd(a.wsum + idx[c], s);
            }
        }
        if (a.cnt) {
            double s = seg_reduce(cntv, lane, run_end, span);
            if (emit) {
#pragma unroll
                for (int c = 0; c < 4; c++) atomicAdd(a.cnt + idx[c], s);
            }
        }
#pragma unroll
        for (int f = 0; f < K; f++) {
            if (a.scheme[f] & PLB_AVG_WEIGHTED) {
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    double s = seg_reduce(vw[f][c], lane, run_end, span);
                    if (emit) atomicAdd(a.acc[f] + idx[c], s);
                }
            } else {
                double s = seg_reduce(vw[f][0], lane, run_end, span);
                if (emit) {
#pragma unroll
                    for (int c = 0; c < 4; c++) atomicAdd(a.acc[f] + idx[c], s);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Chunk kernel for the weighted schemes (every trac2grid call of the time loop).  Same algorithm as
// k_t2g_scatter -- 4 consecutive markers per thread, cell runs combined across lanes, one atomic
// per (run, node, quantity) -- restructured around what ncu showed to bound that kernel (the L1
// load/store pipe at 67 % with 8-byte loads at a 32-byte lane stride, and ~200 issued instructions
// per marker):
//  * a thread's 4 positions arrive as two 256-bit loads, its 4 values of a field as one;
//  * the four corner weights of the 4 markers are computed once and kept; the fields are then
//    processed one after the other (load, sum, reduce, add to the plane), so no accumulator lives
//    across fields and the register count does not grow with K;
//  * 32-bit cell indices; axis coordinate and reciprocal spacing share one 16-byte table entry;
//  * a chunk that straddles cells keeps two aggregates (its first and its last run of equal cells,
//    selected by masking the values, no branches); markers between those runs (three or more cells
//    in one chunk) and chunks with a marker outside the grid take the one-marker path.
// Summation order differs from the generic kernel, results agree to rounding.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void ld256(const double* p, double& a, double& b, double& c, double& d) {
    // volatile: never hoisted above the guard of a lane whose chunk lies beyond the arrays
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}

// one marker straight to the planes (weighted schemes; rare paths only)
__device__ __forceinline__ void t2g_single(const T2GArgs& a, const double2* __restrict__ trx, long long m) {
    const T2GMarker<1> r = t2g_load<1>(a, trx, m);     // cell and weights; field 0 is reloaded below
    if (r.cell < 0) return;
    const long long idx[4] = {r.cell, r.cell + a.nxe, r.cell + 1, r.cell + a.nxe + 1};
#pragma unroll
    for (int c = 0; c < 4; c++) atomicAdd(a.wsum + idx[c], r.w[c]);
#pragma unroll 1
    for (int f = 0; f < a.k; f++) {
        double val = a.f[f][m];
        if (!(a.scheme[f] & PLB_AVG_ARITHMETIC)) val = log(val);
#pragma unroll
        for (int c = 0; c < 4; c++) atomicAdd(a.acc[f] + idx[c], val * r.w[c]);
    }
}

// the first / last run of a lane's chunk as seen by the warp-level segmented reduction: head of a run
// of lanes with the same cell, last lane of that run, longest remaining run in the warp
struct RunInfo {
    int run_end, span;
    bool emit;
};

__device__ __forceinline__ RunInfo t2g_runs(int cell, bool has, int lane) {
    const unsigned full = 0xffffffffu;
    const int prev = __shfl_up_sync(full, cell, 1);
    const bool head = (lane == 0) || (prev != cell);
    const unsigned heads = __ballot_sync(full, head);
    const unsigned after = (lane == 31) ? 0u : (heads >> (lane + 1));
    RunInfo r;
    r.run_end = after ? lane + __ffs(after) - 1 : 31;
    r.span = __reduce_max_sync(full, r.run_end - lane);
    r.emit = head && has;
    return r;
}

// segmented reduction of the four corner sums over runs of lanes, then one atomic per corner by the head
__device__ __forceinline__ void t2g_reduce_emit(double* __restrict__ plane, double (&S)[4], const RunInfo& r,
                                                int lane, int cell, int nxe) {
    for (int o = 1; o <= r.span; o <<= 1) {
        const bool take = lane + o <= r.run_end;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const double t = __shfl_down_sync(0xffffffffu, S[c], o);
            if (take) S[c] += t;
        }
    }
    if (r.emit) {
        const int off[4] = {0, nxe, 1, nxe + 1};
#pragma unroll
        for (int c = 0; c < 4; c++) atomicAdd(plane + (cell + off[c]), S[c]);
    }
}

// sums of one quantity over the chunk's last run (u >= s) and first run (u < e); the last runs are
// combined across the lanes of the warp, the first runs too when `merge` is set (otherwise every
// lane adds its own first run)
__device__ __forceinline__ void t2g_chunk_quantity(double* __restrict__ plane, const double (&v)[4],
                                                   const double (&wu)[4][4], int s, int e, int lane,
                                                   const RunInfo& rl, const RunInfo& rf, bool any_first, bool merge,
                                                   int cell_last, int cell_first, int nxe) {
    double L[4] = {0, 0, 0, 0};
#pragma unroll
    for (int u = 0; u < 4; u++) {
        const double vl = (u >= s) ? v[u] : 0.0;
#pragma unroll
        for (int c = 0; c < 4; c++) L[c] = fma(vl, wu[u][c], L[c]);
    }
    t2g_reduce_emit(plane, L, rl, lane, cell_last, nxe);
    if (any_first) {
        double F[4] = {0, 0, 0, 0};
#pragma unroll
        for (int u = 0; u < 3; u++) {
            const double vf = (u < e) ? v[u] : 0.0;
#pragma unroll
            for (int c = 0; c < 4; c++) F[c] = fma(vf, wu[u][c], F[c]);
        }
        if (merge) {
            t2g_reduce_emit(plane, F, rf, lane, cell_first, nxe);
        } else if (e > 0) {
            const int off[4] = {0, nxe, 1, nxe + 1};
#pragma unroll
            for (int c = 0; c < 4; c++) atomicAdd(plane + (cell_first + off[c]), F[c]);
        }
    }
}

template <bool MERGE>
__global__ void __launch_bounds__(256, 3)
k_t2g_chunk(long long nchunk, const double2* __restrict__ trx, T2GArgs a) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const double zlast = a.axz[a.nze - 1], xlast = a.axx[a.nxe - 1];
    long long base = (blockIdx.x * (long long)blockDim.x + threadIdx.x) - lane;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; base < nchunk; base += stride) {
        const long long ch = base + lane;
        const bool live = ch < nchunk;
        int cell[4] = {0, 0, 0, 0};
        double wu[4][4];
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int c = 0; c < 4; c++) wu[u][c] = 0;
        bool ok = live;
        if (live) {
            double pz[4], px[4];
            const double* xp = (const double*)(trx + 4 * ch);
            ld256(xp, pz[0], px[0], pz[1], px[1]);
            ld256(xp + 4, pz[2], px[2], pz[3], px[3]);
#pragma unroll
            for (int u = 0; u < 4; u++) {
                int ie = __double2int_rd((pz[u] - a.z0) * a.sz);
                int je = __double2int_rd((px[u] - a.x0) * a.sx);
                // a marker on the upper edge of the (extended) axis belongs to the last cell
                if (ie == a.nze - 1 && pz[u] <= zlast) ie = a.nze - 2;
                if (je == a.nxe - 1 && px[u] <= xlast) je = a.nxe - 2;
                const bool in = (unsigned)ie <= (unsigned)(a.nze - 2) && (unsigned)je <= (unsigned)(a.nxe - 2);
                ok = ok && in;
                ie = in ? ie : 0, je = in ? je : 0;
                const double2 tz = __ldg(a.tz + ie), tx = __ldg(a.tx + je);
                const double az = (pz[u] - tz.x) * tz.y;          // pylamp_trac.py:247
                const double ax = (px[u] - tx.x) * tx.y;
                const double bz = 1 - az, bx = 1 - ax;            // :249
                wu[u][0] = (1 - ax) * (1 - az);                   // node (i  , j  )   :252-255
                wu[u][1] = (1 - ax) * (1 - bz);                   // node (i+1, j  )
                wu[u][2] = (1 - bx) * (1 - az);                   // node (i  , j+1)
                wu[u][3] = (1 - bx) * (1 - bz);                   // node (i+1, j+1)
                cell[u] = ie * a.nxe + je;
            }
        }
        // last run = markers [s, 4), first run = markers [0, e) (only if it is not the last run too),
        // markers [e, s) in between take the one-marker path; a chunk with a marker outside: all four do
        int s = 4, e = live ? 0 : 4;
        if (ok) {
            s = 3;
            if (cell[2] == cell[3]) s = 2;
            if (s == 2 && cell[1] == cell[2]) s = 1;
            if (s == 1 && cell[0] == cell[1]) s = 0;
            if (s > 0) {
                e = 1;
                if (s > 1 && cell[1] == cell[0]) e = 2;
                if (e == 2 && s > 2 && cell[2] == cell[0]) e = 3;
            }
        }
        const int cell_last = ok ? cell[3] : -1 - lane;   // lanes without a last run form runs of their own
        const bool any_first = __any_sync(full, ok && e > 0);
        const bool any_mid = __any_sync(full, live && s > e);
        const int ef = ok ? e : 0;                        // first run only exists for clean chunks
        const int cell_first = ef > 0 ? cell[0] : -1 - lane;
        constexpr bool merge = MERGE;
        const RunInfo rl = t2g_runs(cell_last, ok, lane);
        RunInfo rf = {lane, 0, false};
        if (any_first && merge) rf = t2g_runs(cell_first, ef > 0, lane);
        // the fields one after the other (rolled: one copy of the code), the next field's values in flight
        double vn[4] = {0, 0, 0, 0};
        if (ok) ld256(a.f[0] + 4 * ch, vn[0], vn[1], vn[2], vn[3]);
        {
            const double one[4] = {1, 1, 1, 1};
            t2g_chunk_quantity(a.wsum, one, wu, s, ef, lane, rl, rf, any_first, merge, cell_last, cell_first, a.nxe);
        }
#pragma unroll 1
        for (int f = 0; f < a.k; f++) {
            double v[4] = {vn[0], vn[1], vn[2], vn[3]};
            if (ok && f + 1 < a.k) ld256(a.f[f + 1] + 4 * ch, vn[0], vn[1], vn[2], vn[3]);
            if (!(a.scheme[f] & PLB_AVG_ARITHMETIC)) {
#pragma unroll
                for (int u = 0; u < 4; u++) v[u] = log(v[u]);
            }
            t2g_chunk_quantity(a.acc[f], v, wu, s, ef, lane, rl, rf, any_first, merge, cell_last, cell_first, a.nxe);
        }
        if (any_mid) {
#pragma unroll 1
            for (int u = 0; u < 4; u++)
                if (live && u >= e && u < s) t2g_single(a, trx, 4 * ch + u);
        }
    }
}

__global__ void __launch_bounds__(256)
k_t2g_finalise(T2GArgs a, int cz0, int cx0, int nz, int nxx, int ld) {
    long long n = (long long)nz * nxx;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n;
         t += (long long)gridDim.x * blockDim.x) {
        int i = (int)(t / nxx), j = (int)(t % nxx);
        long long src = (long long)(i + cz0) * a.nxe + (j + cx0);
        for (int f = 0; f < a.k; f++) {
            int sc = a.scheme[f];
            double den = (sc & PLB_AVG_WEIGHTED) ? a.wsum[src] : a.cnt[src];
            double s = a.acc[f][src];
            double r;
            if ((sc & PLB_AVG_GEOMETRIC) && !(sc & PLB_AVG_ARITHMETIC)) {
                if (isinf(s)) s = 0;                 // pylamp_trac.py:301
                r = exp(s / den);                    // :304, :306
            } else {
                r = s / den;                         // :281, :287
            }
            a.out[f][(long long)i * ld + j] = r;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// grid2trac
// ---------------------------------------------------------------------------------------------
struct G2TGrid {
    const double* gz;
    const double* gx;
    int nz, nxx, ld;
    double z0, zlen, x0, xlen;
};

struct Cell {
    long long ie, je;
    double dzn, dxn;      // normalised local coordinates (pylamp_trac.py:89-90)
    double dz0, dz1, dx0, dx1;
    bool bad;
};

__device__ __forceinline__ Cell locate(const G2TGrid& g, double z, double x) {
    Cell c;
    c.ie = cell_of(z, g.z0, g.zlen, g.nz);
    c.je = cell_of(x, g.x0, g.xlen, g.nxx);
    // the reference tests ie > n-1 (pylamp_trac.py:52); ie == n-1 would index past the axis
    // there (NumPy raises), so it is treated as outside as well
    c.bad = (c.ie < 0) || (c.ie > g.nz - 2) || (c.je < 0) || (c.je > g.nxx - 2);
    if (c.bad) c.ie = 0, c.je = 0;
    c.dz0 = z - g.gz[c.ie];
    c.dz1 = -(z - g.gz[c.ie + 1]);
    c.dx0 = x - g.gx[c.je];
    c.dx1 = -(x - g.gx[c.je + 1]);
    c.dxn = c.dx0 / (c.dx0 + c.dx1);
    c.dzn = c.dz0 / (c.dz0 + c.dz1);
    return c;
}

__device__ __forceinline__ double bilin(const double* __restrict__ f, int ld, const Cell& c) {
    const double* p = f + c.ie * ld + c.je;
    double f00 = __ldg(p), f01 = __ldg(p + 1), f10 = __ldg(p + ld), f11 = __ldg(p + ld + 1);
    return (1 - c.dxn) * (1 - c.dzn) * f00 + c.dxn * (1 - c.dzn) * f01 +
           (1 - c.dxn) * c.dzn * f10 + c.dxn * c.dzn * f11;               // :92-96
}

// Meyer & Jenny (2004) divergence-conserving correction, pylamp_trac.py:98-154
__device__ __forceinline__ void veldiv(const double* __restrict__ fz, const double* __restrict__ fx,
                                       const G2TGrid& g, const Cell& c, double& vz, double& vx) {
    const double* pz = fz + c.ie * g.ld + c.je;
    const double* px = fx + c.ie * g.ld + c.je;
    double z00 = __ldg(pz), z01 = __ldg(pz + 1), z10 = __ldg(pz + g.ld), z11 = __ldg(pz + g.ld + 1);
    double x00 = __ldg(px), x01 = __ldg(px + 1), x10 = __ldg(px + g.ld), x11 = __ldg(px + g.ld + 1);
    double hz = g.gz[c.ie + 1] - g.gz[c.ie];
    double hx = g.gx[c.je + 1] - g.gx[c.je];
    double c10 = (0.5 * hx / hz) * (z00 - z10 + z11 - z01);
    double c20 = (0.5 * hz / hx) * (x00 - x01 + x11 - x10);
    double w00 = (1 - c.dxn) * (1 - c.dzn), w01 = c.dxn * (1 - c.dzn), w10 = (1 - c.dxn) * c.dzn,
           w11 = c.dxn * c.dzn;
    double ux = w00 * x00 + w01 * x01 + w10 * x10 + w11 * x11;
    double uz = w00 * z00 + w01 * z01 + w10 * z10 + w11 * z11;
    vx = ux + c.dxn * (1 - c.dxn) * c10;
    vz = uz + c.dzn * (1 - c.dzn) * c20;
}

struct G2TArgs {
    const double* f[PLB_MAX_FIELDS];
    double* out[PLB_MAX_FIELDS];
    int k;
};

__global__ void __launch_bounds__(256)
k_grid2trac(long long M, const double2* __restrict__ trx, int method, G2TGrid g, G2TArgs a,
            double defval, unsigned long long* n_outside) {
    unsigned long long bad_local = 0;
    for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < M;
         m += (long long)gridDim.x * blockDim.x) {
        double2 p = trx[m];
        Cell c = locate(g, p.x, p.y);
        if (c.bad) {
            bad_local++;
            for (int f = 0; f < a.k; f++) a.out[f][m] = defval;
            continue;
        }
        if (method & PLB_METHOD_NEAREST) {                 // :77-85
            double d[4] = {c.dz0 * c.dz0 + c.dx0 * c.dx0, c.dz0 * c.dz0 + c.dx1 * c.dx1,
                           c.dz1 * c.dz1 + c.dx0 * c.dx0, c.dz1 * c.dz1 + c.dx1 * c.dx1};
            int best = 0;
            for (int q = 1; q < 4; q++)
                if (d[q] < d[best]) best = q;
            long long off = (c.ie + best / 2) * g.ld + c.je + best % 2;
            for (int f = 0; f < a.k; f++) a.out[f][m] = a.f[f][off];
        } else if (method & PLB_METHOD_LINEAR) {
            for (int f = 0; f < a.k; f++) a.out[f][m] = bilin(a.f[f], g.ld, c);
        } else {                                           // VELDIV, fields = (vz, vx)
            double vz, vx;
            veldiv(a.f[0], a.f[1], g, c, vz, vx);
            a.out[0][m] = vz;
            a.out[1][m] = vx;
        }
    }
    bad_local = __reduce_add_sync(0xffffffffu, (unsigned)bad_local);
    if ((threadIdx.x & 31) == 0 && bad_local) atomicAdd(n_outside, bad_local);
}

// ---------------------------------------------------------------------------------------------
// RK4 with the reference's (1/6)(k1+k2+k3+k4) update, pylamp_trac.py:347-388
// ---------------------------------------------------------------------------------------------
// RK stage velocity: same arithmetic as locate() + veldiv() with the divisions hoisted -- the two
// reciprocal cell sizes serve the local coordinates and both Meyer-Jenny coefficients (per-axis tables filled
// by k_axis_recip with the same IEEE division: bit-identical results, no fp64 division per stage, which is
// what bounded this kernel).  The cell of a position is the reference's floor((n-1)(x-x0)/L)
// (pylamp_trac.py:46-47): evaluated by a multiplication, and re-evaluated with the exact multiply-then-divide
// whenever the position lies within 1e-7 cells of a face -- the Meyer-Jenny correction of the tangential
// component is discontinuous across faces, so the cell choice matters there.

// locate() for the marker-temperature kernels of the time loop: the cell by multiplication (exact re-evaluation
// within 1e-7 cells of a face, like RK4 -- bilinear interpolation is continuous across faces, so the choice does
// not even matter here) and the local coordinates by a multiplication with the tabulated reciprocal cell size
// instead of the reference's (x - g0)/((x - g0) + (g1 - x)) (pylamp_trac.py:74-75): four fp64 divisions per marker
// less; the two forms differ by a rounding error (1e-16 relative), the public plb_grid2trac keeps the reference's.
struct CellF {
    int o;                // ie * ld + je
    double dzn, dxn;
    bool bad;
};

__device__ __forceinline__ CellF locate_fast(const G2TGrid& g, const double* __restrict__ riz, const double* __restrict__ rix,
                                             double sz, double sx, double z, double x) {
    CellF c;
    int ie = __double2int_rd((z - g.z0) * sz), je = __double2int_rd((x - g.x0) * sx);
    c.bad = (unsigned)ie > (unsigned)(g.nz - 2) || (unsigned)je > (unsigned)(g.nxx - 2);
    if (c.bad) {                                     // (rare: let the reference's formula decide what is outside)
        const long long a = cell_of(z, g.z0, g.zlen, g.nz), b = cell_of(x, g.x0, g.xlen, g.nxx);
        c.bad = (a < 0) || (a > g.nz - 2) || (b < 0) || (b > g.nxx - 2);
        ie = c.bad ? 0 : (int)a, je = c.bad ? 0 : (int)b;
    }
    c.o = ie * g.ld + je;
    c.dzn = (z - g.gz[ie]) * riz[ie];
    c.dxn = (x - g.gx[je]) * rix[je];
    return c;
}

// pylamp_trac.py:92-96 with the products summed by fused multiply-adds
__device__ __forceinline__ double bilin_fast(const double* __restrict__ f, int ld, const CellF& c) {
    const double* p = f + c.o;
    const double f00 = __ldg(p), f01 = __ldg(p + 1), f10 = __ldg(p + ld), f11 = __ldg(p + ld + 1);
    const double omz = 1 - c.dzn, omx = 1 - c.dxn;
    return fma(omx * omz, f00, fma(c.dxn * omz, f01, fma(omx * c.dzn, f10, (c.dxn * c.dzn) * f11)));
}

// EXACT = false: the cell by a multiplication; returns true when the position lies within 1e-7 cells of a face (or
// outside the grid), where the multiplication may have picked the neighbour of the reference's cell -- and the
// Meyer-Jenny term is discontinuous across faces.  EXACT = true: the reference's own formula (multiply, then divide,
// pylamp_trac.py:46-47).  k_rk4 runs the four stages with EXACT = false and repeats the (rare) markers that
// reported a face with EXACT = true.
template <bool EXACT>
__device__ __forceinline__ bool vel_at(const double* __restrict__ fz, const double* __restrict__ fx,
                                       const G2TGrid& g, const double* __restrict__ riz,
                                       const double* __restrict__ rix, double sz, double sx, double z, double x,
                                       double& vz, double& vx) {
    int ie, je;                                            // (the host side checks nz * ld < 2^31)
    if (EXACT) {
        const long long a = cell_of(z, g.z0, g.zlen, g.nz), b = cell_of(x, g.x0, g.xlen, g.nxx);
        ie = a < 0 ? -1 : (a > g.nz ? g.nz : (int)a);
        je = b < 0 ? -1 : (b > g.nxx ? g.nxx : (int)b);
    } else {
        ie = __double2int_rd((z - g.z0) * sz), je = __double2int_rd((x - g.x0) * sx);     // (saturating)
    }
    if ((unsigned)ie > (unsigned)(g.nz - 2) || (unsigned)je > (unsigned)(g.nxx - 2)) {
        vz = 0, vx = 0;                                    // defval=0, :361
        return true;
    }
    const double gz0 = g.gz[ie], gz1 = g.gz[ie + 1], gx0 = g.gx[je], gx1 = g.gx[je + 1];
    const double hz = gz1 - gz0, hx = gx1 - gx0;
    const double rz = riz[ie], rx = rix[je];          // 1/hz, 1/hx
    const double dzn = (z - gz0) * rz, dxn = (x - gx0) * rx;
    const int o = ie * g.ld + je;
    const double* pz = fz + o;
    const double* px = fx + o;
    const double z00 = __ldg(pz), z01 = __ldg(pz + 1), z10 = __ldg(pz + g.ld), z11 = __ldg(pz + g.ld + 1);
    const double x00 = __ldg(px), x01 = __ldg(px + 1), x10 = __ldg(px + g.ld), x11 = __ldg(px + g.ld + 1);
    const double c10 = (0.5 * hx * rz) * (z00 - z10 + z11 - z01);
    const double c20 = (0.5 * hz * rx) * (x00 - x01 + x11 - x10);
    const double omz = 1 - dzn, omx = 1 - dxn;
    const double w00 = omx * omz, w01 = dxn * omz, w10 = omx * dzn, w11 = dxn * dzn;
    const double bz = dzn * omz, bx = dxn * omx;
    // (this file is compiled without fp contraction for the exact index formulas; the sums of products here are
    // fused by hand: one rounding less per term than the reference's, 1e-16 relative)
    vx = fma(w00, x00, fma(w01, x01, fma(w10, x10, fma(w11, x11, bx * c10))));
    vz = fma(w00, z00, fma(w01, z01, fma(w10, z10, fma(w11, z11, bz * c20))));
    // within ~1e-7 cells of a face (d (1 - d) <= min(d, 1 - d); negative when the multiplication picked a neighbour)
    return fmin(bz, bx) < 1e-7;
}

template <bool EXACT>
__device__ __forceinline__ bool rk4_stages(const double* __restrict__ fz, const double* __restrict__ fx, const G2TGrid& g,
                                           const double* __restrict__ riz, const double* __restrict__ rix, double sz,
                                           double sx, double dt, const double2& p, double2& q) {
    const double hdt = 0.5 * dt, sixth_dt = (1.0 / 6.0) * dt;
    double k1z, k1x, k2z, k2x, k3z, k3x, k4z, k4x;
    bool face = vel_at<EXACT>(fz, fx, g, riz, rix, sz, sx, p.x, p.y, k1z, k1x);
    face |= vel_at<EXACT>(fz, fx, g, riz, rix, sz, sx, fma(hdt, k1z, p.x), fma(hdt, k1x, p.y), k2z, k2x);
    face |= vel_at<EXACT>(fz, fx, g, riz, rix, sz, sx, fma(hdt, k2z, p.x), fma(hdt, k2x, p.y), k3z, k3x);
    face |= vel_at<EXACT>(fz, fx, g, riz, rix, sz, sx, fma(dt, k3z, p.x), fma(dt, k3x, p.y), k4z, k4x);
    q.x = p.x + sixth_dt * (((k1z + k2z) + k3z) + k4z);   // :385 (unweighted sum)
    q.y = p.y + sixth_dt * (((k1x + k2x) + k3x) + k4x);
    return face;
}

// FENCE: the fence of pylamp2.py:558-572 and the cell index / per-cell count of :588-593 in the same pass (the new
// positions are final here) -- same operations as k_fence_count, one read of the coordinates less.
struct FenceArgs {
    double Lz, Lx, eps;
    int nz, nxx;                 // NODE grid of the per-cell count
    long long* kelem;
    unsigned long long* count;
};

template <bool FENCE>
__global__ void __launch_bounds__(256, 4)
k_rk4(long long M, const double2* __restrict__ trx, const double* __restrict__ fz,
      const double* __restrict__ fx, G2TGrid g, const double* __restrict__ riz,
      const double* __restrict__ rix, double dt, double2* __restrict__ xout, double2* __restrict__ vout, FenceArgs fa) {
    const double sz = (double)(g.nz - 1) / g.zlen, sx = (double)(g.nxx - 1) / g.xlen;
    const double rdt = 1.0 / dt;
    const long long ncell = FENCE ? (long long)(fa.nz - 1) * (fa.nxx - 1) : 0;
    for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < M;
         m += (long long)gridDim.x * blockDim.x) {
        const double2 p = trx[m];
        double2 q;
        if (rk4_stages<false>(fz, fx, g, riz, rix, sz, sx, dt, p, q)) rk4_stages<true>(fz, fx, g, riz, rix, sz, sx, dt, p, q);
        if (vout) {
            double2 v;
            v.x = (q.x - p.x) * rdt;                          // :386
            v.y = (q.y - p.y) * rdt;
            vout[m] = v;
        }
        if (FENCE) {
            if (q.x <= 0) q.x = fa.eps;                       // pylamp2.py:558-572
            if (q.x >= fa.Lz) q.x = fa.Lz - fa.eps;
            if (q.y <= 0) q.y = fa.eps;
            if (q.y >= fa.Lx) q.y = fa.Lx - fa.eps;
            const long long ie = (long long)floor(__ddiv_rn(__dmul_rn((double)(fa.nz - 1), q.x), fa.Lz));   // :588-589
            const long long je = (long long)floor(__ddiv_rn(__dmul_rn((double)(fa.nxx - 1), q.y), fa.Lx));
            const long long k = ie * (fa.nxx - 1) + je;
            if (fa.kelem) fa.kelem[m] = k;
            if (fa.count && k >= 0 && k < ncell) atomicAdd(fa.count + k, 1ull);
        }
        xout[m] = q;
    }
}

// ---------------------------------------------------------------------------------------------
// fence, cell index + count, property update, subgrid stages
// ---------------------------------------------------------------------------------------------
// walls: bit 0 z = 0, bit 1 x = 0, bit 2 z = L, bit 3 x = L (the order of the reference's bc lists); a wall
// whose bit is clear is a flow-through wall: markers beyond it are left where they are (and removed by the caller,
// pylamp2.py:563-581)
__global__ void __launch_bounds__(256)
k_fence(long long M, double2* __restrict__ trx, double Lz, double Lx, double eps, int walls) {
    for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < M;
         m += (long long)gridDim.x * blockDim.x) {
        double2 p = trx[m];
        bool ch = false;
        if (p.x <= 0 && (walls & 1)) p.x = eps, ch = true;
        if (p.x >= Lz && (walls & 4)) p.x = Lz - eps, ch = true;
        if (p.y <= 0 && (walls & 2)) p.y = eps, ch = true;
        if (p.y >= Lx && (walls & 8)) p.y = Lx - eps, ch = true;
        if (ch) trx[m] = p;
    }
}

__global__ void __launch_bounds__(256)
k_cell_index_count(long long M, const double2* __restrict__ trx, int nz, int nxx, double Lz,
                   double Lx, long long* __restrict__ kelem, unsigned long long* __restrict__ count) {
    const long long ncell = (long long)(nz - 1) * (nxx - 1);
    for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < M;
         m += (long long)gridDim.x * blockDim.x) {
        double2 p = trx[m];
        // pylamp2.py:588-589: floor((n-1)*x/L) -- multiply, then divide
        long long ie = (long long)floor(__ddiv_rn(__dmul_rn((double)(nz - 1), p.x), Lz));
        long long je = (long long)floor(__ddiv_rn(__dmul_rn((double)(nxx - 1), p.y), Lx));
        long long k = ie * (nxx - 1) + je;
        if (kelem) kelem[m] = k;
        if (count && k >= 0 && k < ncell) atomicAdd(count + k, 1ull);
    }
}

// fence + cell index + per-cell count in one pass over the coordinates (same operations as k_fence
// followed by k_cell_index_count: bit-identical results, one read of tr_x instead of two)
__global__ void __launch_bounds__(256)
k_fence_count(long long M, double2* __restrict__ trx, double Lz, double Lx, double eps, int nz, int nxx,
              long long* __restrict__ kelem, unsigned long long* __restrict__ count) {
    const long long ncell = (long long)(nz - 1) * (nxx - 1);
    for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < M;
         m += (long long)gridDim.x * blockDim.x) {
        double2 p = trx[m];
        bool ch = false;
        if (p.x <= 0) p.x = eps, ch = true;
        if (p.x >= Lz) p.x = Lz - eps, ch = true;
        if (p.y <= 0) p.y = eps, ch = true;
        if (p.y >= Lx) p.y = Lx - eps, ch = true;
        if (ch) trx[m] = p;
        long long ie = (long long)floor(__ddiv_rn(__dmul_rn((double)(nz - 1), p.x), Lz));
        long long je = (long long)floor(__ddiv_rn(__dmul_rn((double)(nxx - 1), p.y), Lx));
        long long k = ie * (nxx - 1) + je;
        if (kelem) kelem[m] = k;
        if (count && k >= 0 && k < ncell) atomicAdd(count + k, 1ull);
    }
}

__global__ void __launch_bounds__(256)
k_update_properties(long long M, int tdep_rho, int tdep_eta, double Tref, double etamin,
                    double etamax, double gasr, const double* __restrict__ T,
                    const double* __restrict__ rho0, const double* __restrict__ alpha,
                    const double* __restrict__ Ea, const double* __restrict__ eta0,
                    double* __restrict__ rho, double* __restrict__ eta) {
    for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < M;
         m += (long long)gridDim.x * blockDim.x) {
        if (tdep_rho) {
            rho[m] = 1.0 / ((alpha[m] * (T[m] - Tref) + 1) / rho0[m]);        // pylamp2.py:294
        } else {
            rho[m] = rho0[m];
        }
        if (tdep_eta) {
            double e = eta0[m] * exp(Ea[m] / (gasr * T[m]) - Ea[m] / (gasr * Tref));   // :298
            if (e < etamin) e = etamin;
            if (e > etamax) e = etamax;
            eta[m] = e;
        } else {
            eta[m] = eta0[m];
        }
    }
}

__global__ void __launch_bounds__(256)
k_subgrid1(long long M, double dt, double fac, const double* __restrict__ Told,
           const double* __restrict__ T, const double* __restrict__ cp,
           const double* __restrict__ rho, const double* __restrict__ k, double* __restrict__ Tsg,
           double* __restrict__ dT) {
    const double d = 0.5;                                                      // pylamp2.py:472
    for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < M;
         m += (long long)gridDim.x * blockDim.x) {
        double tau = cp[m] * rho[m] / (k[m] * fac);                            // :473
        double tsg = Told[m] - (Told[m] - T[m]) * exp(-d * dt / tau);          // :474
        Tsg[m] = tsg;
        dT[m] = tsg - T[m];                                                    // :475
    }
}

__global__ void __launch_bounds__(256)
k_sub(long long M, const double* __restrict__ a, const double* __restrict__ b,
      double* __restrict__ out) {
    for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < M;
         m += (long long)gridDim.x * blockDim.x)
        out[m] = a[m] - b[m];
}

// Fused marker temperature update of pylamp2.py:448-475: T1 = T + interp(dT_grid) (:453-455), then the
// subgrid relaxation Tsg = Told - (Told - T1) exp(-d dt / tau), dT = Tsg - T1 (:472-475) with Told = T.
// One pass over the markers instead of clone + grid2trac + add + stage 1 (same arithmetic).
// (both kernels: the streaming loads of a thread's NEXT marker are issued before the gathers of the current one --
// ncu on the plain loop showed 24 of 32 resident warps waiting on memory per issue slot at 0.58 of the HBM peak)
__global__ void __launch_bounds__(256, 4)
k_subgrid_fused1(long long M, const double2* __restrict__ trx, G2TGrid g, const double* __restrict__ riz,
                 const double* __restrict__ rix, const double* __restrict__ dTg,
                 double dt, double fac, const double* __restrict__ T, const double* __restrict__ cp,
                 const double* __restrict__ rho, const double* __restrict__ k, double* __restrict__ Tsg,
                 double* __restrict__ dT, unsigned long long* n_outside) {
    const double d = 0.5;
    const double sz = (double)(g.nz - 1) / g.zlen, sx = (double)(g.nxx - 1) / g.xlen;
    const long long stride = (long long)gridDim.x * blockDim.x;
    unsigned bad_local = 0;
    long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (m < M) {
        double2 p = trx[m];
        double told = T[m], vcp = cp[m], vrho = rho[m], vk = k[m];
        for (;;) {
            const long long mn = m + stride;
            const bool more = mn < M;
            double2 pn = p;
            double toldn = 0, vcpn = 0, vrhon = 0, vkn = 0;
            if (more) pn = trx[mn], toldn = T[mn], vcpn = cp[mn], vrhon = rho[mn], vkn = k[mn];
            const CellF c = locate_fast(g, riz, rix, sz, sx, p.x, p.y);
            if (c.bad) {
                bad_local++;
            } else {
                const double t1 = told + bilin_fast(dTg, g.ld, c);
                const double tau = vcp * vrho / (vk * fac);
                const double tsg = told - (told - t1) * exp(-d * dt / tau);
                Tsg[m] = tsg;
                dT[m] = tsg - t1;
            }
            if (!more) break;
            m = mn, p = pn, told = toldn, vcp = vcpn, vrho = vrhon, vk = vkn;
        }
    }
    bad_local = __reduce_add_sync(0xffffffffu, bad_local);
    if ((threadIdx.x & 31) == 0 && bad_local) atomicAdd(n_outside, (unsigned long long)bad_local);
}

// T = Tsg - interp(f_sgc), pylamp2.py:479-480
__global__ void __launch_bounds__(256, 4)
k_subgrid_fused2(long long M, const double2* __restrict__ trx, G2TGrid g, const double* __restrict__ riz,
                 const double* __restrict__ rix, const double* __restrict__ sgc,
                 const double* __restrict__ Tsg, double* __restrict__ T, unsigned long long* n_outside) {
    const double sz = (double)(g.nz - 1) / g.zlen, sx = (double)(g.nxx - 1) / g.xlen;
    const long long stride = (long long)gridDim.x * blockDim.x;
    unsigned bad_local = 0;
    long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (m < M) {
        double2 p = trx[m];
        double tsg = Tsg[m];
        for (;;) {
            const long long mn = m + stride;
            const bool more = mn < M;
            double2 pn = p;
            double tsgn = 0;
            if (more) pn = trx[mn], tsgn = Tsg[mn];
            const CellF c = locate_fast(g, riz, rix, sz, sx, p.x, p.y);
            if (c.bad) bad_local++;
            else T[m] = tsg - bilin_fast(sgc, g.ld, c);
            if (!more) break;
            m = mn, p = pn, tsg = tsgn;
        }
    }
    bad_local = __reduce_add_sync(0xffffffffu, bad_local);
    if ((threadIdx.x & 31) == 0 && bad_local) atomicAdd(n_outside, (unsigned long long)bad_local);
}


// ---------------------------------------------------------------------------------------------
// Fused marker->grid pass for the targets of one time step (pylamp2.py:309-313: nodes, cell centres
// and the two half-staggered grids; :478: the subgrid term on the nodes).
//
// A CTA stages a chunk of TF_NM consecutive markers -- the coordinates and every distinct property
// column, each read from HBM exactly once per step -- in shared memory with 1-D bulk async copies
// (TMA: cp.async.bulk + mbarrier), takes the logarithm of the geometrically averaged columns in
// place, and splits the chunk into RUNS of consecutive markers of one node-grid cell (after the
// marker-by-cell sort a run is a cell's whole population).  A thread then owns a (run, target) pair:
// it walks the run's markers in shared memory and keeps the sums of that cell's destination nodes
// (2x2 on the node grid; 3x3 / 3x2 / 2x3 on the staggered grids, where the half cell a marker lies
// in selects two of three candidate nodes per staggered axis) in registers -- one FMA per
// (marker, node, quantity), no shuffles, no masks -- and adds them to the planes once per run.
// Correct for any marker order (an unsorted cloud just has runs of length one).  Same weights as the
// reference: (1-a), 1-(1-a) per axis, multiplied x then z (pylamp_trac.py:247-255).
// ---------------------------------------------------------------------------------------------
constexpr int TF_THREADS = 256, TF_MAXC = 8, TF_MAXT = 8, TF_NFMAX = 6;

struct TFTask {
    int type;              // 0: 2x2 (node axes), 1: 3x3 (both axes staggered), 2: 3x2 (z staggered), 3: 2x3 (x staggered)
    int nf, ws;            // fields of this task (type 0: <= TF_NFMAX, else 1); accumulate the weight sums too
    int col[TF_NFMAX];     // staged column of each field
    const double2* tz;     // {coordinate, 1/spacing} of the target's (ghost-extended) axes
    const double2* tx;
    int lz, lx;            // extended index of the first destination row / column of node-grid cell 0
    int nxe;               // row length of the target's planes
    double* wsum;
    double* acc[TF_NFMAX];
};

struct TFArgs {
    const double* col[TF_MAXC];
    int ncol;
    unsigned logmask;      // columns averaged geometrically: staged as log(value)
    double z0, x0, sz, sx; // node grid: first node and cells per unit length
    int ncz, ncx;          // node-grid cells
    TFTask t[TF_MAXT];
    int nt;
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tf_bulk_load(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ int tf_cell(double v, double v0, double s, int nc) {
    int i = __double2int_rd((v - v0) * s);
    return i < 0 ? 0 : (i > nc - 1 ? nc - 1 : i);
}

template <int R>
__device__ __forceinline__ void tf_axis_weights(double v, const double2& t0, const double2& t1, double (&w)[R]) {
    if (R == 2) {
        const double a = (v - t0.x) * t0.y;          // pylamp_trac.py:247
        w[0] = 1 - a;                                 // :252 (1 - a)
        w[1] = 1 - w[0];                              // :249, :253 (1 - b), b = 1 - a
    } else {
        const bool lower = v < t1.x;                  // which half of the node-grid cell
        const double2 t = lower ? t0 : t1;
        const double a = (v - t.x) * t.y;
        const double u = 1 - a, q = 1 - u;
        w[0] = lower ? u : 0.0;
        w[1] = lower ? q : u;
        w[R - 1] = lower ? 0.0 : q;
    }
}

// One (run, task) work item.  PARTS lanes may share a run: each walks 1/PARTS of its markers, the partial
// sums are combined with xor-shuffles and the first lane of the group adds them to the planes.  Every lane
// of the warp calls this (lanes without a run pass len = 0): the shuffles are warp-wide.
template <int RZ, int RX, int NF, int PARTS, int NM>
__device__ __forceinline__ void tf_run(const TFArgs& a, const TFTask& t, const double2* __restrict__ sx,
                                       const double* __restrict__ sv, int s, int len, int part) {
    const double2 p0 = sx[s];                            // (lanes without a run pass s = 0, len = 0)
    const int ie = tf_cell(p0.x, a.z0, a.sz, a.ncz), je = tf_cell(p0.y, a.x0, a.sx, a.ncx);
    const int ez = ie + t.lz, ex = je + t.lx;
    const double2 tz0 = __ldg(t.tz + ez), tx0 = __ldg(t.tx + ex);
    double2 tz1 = tz0, tx1 = tx0;
    if (RZ == 3) tz1 = __ldg(t.tz + ez + 1);
    if (RX == 3) tx1 = __ldg(t.tx + ex + 1);
    double W[RZ * RX], A[NF][RZ * RX];
#pragma unroll
    for (int c = 0; c < RZ * RX; c++) {
        W[c] = 0;
#pragma unroll
        for (int f = 0; f < NF; f++) A[f][c] = 0;
    }
    const bool ws = t.ws != 0;
    const int b0 = PARTS == 1 ? 0 : (len * part) / PARTS, sub = (PARTS == 1 ? len : (len * (part + 1)) / PARTS) - b0;
    const double2* px = sx + s + b0;
    // the fields of a task are staged in consecutive columns (host side): one address, NF immediate offsets
    const double* cf0 = sv + t.col[0] * NM + s + b0;
    // The lanes of a warp walk different runs.  Each starts at the marker whose shared-memory bank matches its
    // lane number and wraps around, so that the 16 lanes of a memory phase read 16 different banks whatever the
    // run starts are (a linear walk of freshly sorted runs, which start 16 markers apart, would put all lanes
    // on one bank: measured 1.6x slower).
    int idx = ((threadIdx.x & 15) - (s + b0)) & 15;
    if (idx >= sub) idx = 0;
    // two markers per trip: both sets of loads are issued before the first is used
    int it = 0;
    for (; it + 2 <= sub; it += 2) {
        int idx2 = idx + 1;
        if (idx2 == sub) idx2 = 0;
        const double2 pa = px[idx], pb = px[idx2];
        double va[NF], vb[NF];
#pragma unroll
        for (int f = 0; f < NF; f++) va[f] = cf0[f * NM + idx], vb[f] = cf0[f * NM + idx2];
        double wza[RZ], wxa[RX], wzb[RZ], wxb[RX];
        tf_axis_weights<RZ>(pa.x, tz0, tz1, wza);
        tf_axis_weights<RX>(pa.y, tx0, tx1, wxa);
        tf_axis_weights<RZ>(pb.x, tz0, tz1, wzb);
        tf_axis_weights<RX>(pb.y, tx0, tx1, wxb);
#pragma unroll
        for (int r = 0; r < RZ; r++)
#pragma unroll
            for (int c = 0; c < RX; c++) {
                const double w1 = wxa[c] * wza[r], w2 = wxb[c] * wzb[r];       // :252-255
                W[r * RX + c] += w1, W[r * RX + c] += w2;       // (kept even when the weight sums go unused: no selects)
#pragma unroll
                for (int f = 0; f < NF; f++) {
                    A[f][r * RX + c] = fma(va[f], w1, A[f][r * RX + c]);
                    A[f][r * RX + c] = fma(vb[f], w2, A[f][r * RX + c]);
                }
            }
        idx = idx2 + 1;
        if (idx == sub) idx = 0;
    }
    if (it < sub) {
        const double2 p = px[idx];
        double v[NF];
#pragma unroll
        for (int f = 0; f < NF; f++) v[f] = cf0[f * NM + idx];
        double wz[RZ], wx[RX];
        tf_axis_weights<RZ>(p.x, tz0, tz1, wz);
        tf_axis_weights<RX>(p.y, tx0, tx1, wx);
#pragma unroll
        for (int r = 0; r < RZ; r++)
#pragma unroll
            for (int c = 0; c < RX; c++) {
                const double w = wx[c] * wz[r];
                W[r * RX + c] += w;
#pragma unroll
                for (int f = 0; f < NF; f++) A[f][r * RX + c] = fma(v[f], w, A[f][r * RX + c]);
            }
    }
    if (PARTS > 1) {
#pragma unroll
        for (int o = 1; o < PARTS; o <<= 1) {
#pragma unroll
            for (int c = 0; c < RZ * RX; c++) {
                W[c] += __shfl_xor_sync(0xffffffffu, W[c], o);
#pragma unroll
                for (int f = 0; f < NF; f++) A[f][c] += __shfl_xor_sync(0xffffffffu, A[f][c], o);
            }
        }
        if (part != 0) return;
    }
    if (len == 0) return;
    const long long base = (long long)ez * t.nxe + ex;
    // A non-finite property value (log of 0, NaN of a marker injected into an empty cell) must reach the nodes
    // the reference adds it to and no others (pylamp_trac.py:276-298 multiplies by the four real corner weights
    // only); above it has also met the structurally zero weights of the staggered patterns.  Such a run -- it
    // shows as a non-finite sum -- is redone marker by marker with the four real corners.
    double chk = 0;
#pragma unroll
    for (int c = 0; c < RZ * RX; c++)
#pragma unroll
        for (int f = 0; f < NF; f++) chk += A[f][c] * 0.0;          // 0 for finite sums, NaN otherwise
    if (chk != 0.0) {
        for (int it = 0; it < len; it++) {
            const double2 p = sx[s + it];
            const bool lowz = RZ == 2 || p.x < tz1.x, lowx = RX == 2 || p.y < tx1.x;
            const double2 az = lowz ? tz0 : tz1, ax = lowx ? tx0 : tx1;
            const double uz = 1 - (p.x - az.x) * az.y, ux = 1 - (p.y - ax.x) * ax.y;
            const double w4[4] = {ux * uz, ux * (1 - uz), (1 - ux) * uz, (1 - ux) * (1 - uz)};
            const long long o = base + (lowz ? 0 : t.nxe) + (lowx ? 0 : 1);
            const long long o4[4] = {o, o + t.nxe, o + 1, o + t.nxe + 1};
#pragma unroll
            for (int c = 0; c < 4; c++) {
                if (ws) atomicAdd(t.wsum + o4[c], w4[c]);
#pragma unroll
                for (int f = 0; f < NF; f++) atomicAdd(t.acc[f] + o4[c], sv[(t.col[0] + f) * NM + s + it] * w4[c]);
            }
        }
        return;
    }
#pragma unroll
    for (int r = 0; r < RZ; r++)
#pragma unroll
        for (int c = 0; c < RX; c++) {
            const long long o = base + (long long)r * t.nxe + c;
            // (3-wide patterns: a destination none of the run's markers reaches keeps exact zeros -- nothing to add)
            const bool any = (RZ == 2 && RX == 2) || W[r * RX + c] != 0.0;
            if (ws && any) atomicAdd(t.wsum + o, W[r * RX + c]);
#pragma unroll
            for (int f = 0; f < NF; f++)
                if (any) atomicAdd(t.acc[f] + o, A[f][r * RX + c]);
        }
}

template <int PARTS, int NM, int MINB>
__global__ void __launch_bounds__(TF_THREADS, MINB)
k_t2g_fused(long long M, const double2* __restrict__ trx, const TFArgs a, int use_tma) {
    constexpr int ROUNDS = (NM + TF_THREADS - 1) / TF_THREADS;
    extern __shared__ __align__(128) unsigned char tf_smem[];
    double2* sx = (double2*)tf_smem;                                     // [NM]
    double* sv = (double*)(tf_smem + (size_t)NM * 16);                   // [ncol][NM]
    unsigned short* rstart = (unsigned short*)(sv + (size_t)a.ncol * NM);   // [NM + 1]
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ int cnt[ROUNDS * (TF_THREADS / 32)];
    __shared__ int nrun_s;
    const unsigned full = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long m0 = (long long)blockIdx.x * NM;
    const int n = (int)((M - m0) < (long long)NM ? (M - m0) : (long long)NM);
    // ---- stage the chunk
    if (use_tma && n == NM) {
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0) {
            const unsigned bytes = (unsigned)NM * 16u + (unsigned)a.ncol * NM * 8u;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&mbar)), "r"(bytes) : "memory");
            tf_bulk_load(sx, trx + m0, NM * 16u, &mbar);
            for (int c = 0; c < a.ncol; c++) tf_bulk_load(sv + (size_t)c * NM, a.col[c] + m0, NM * 8u, &mbar);
        }
        unsigned done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&mbar)) : "memory");
        }
    } else {
        for (int m = tid; m < n; m += TF_THREADS) sx[m] = trx[m0 + m];
        for (int c = 0; c < a.ncol; c++)
            for (int m = tid; m < n; m += TF_THREADS) sv[(size_t)c * NM + m] = a.col[c][m0 + m];
        __syncthreads();
    }
    // ---- logarithm of the geometrically averaged columns, in place (once per marker and step)
    for (int c = 0; c < a.ncol; c++)
        if ((a.logmask >> c) & 1u)
            for (int m = tid; m < n; m += TF_THREADS) sv[(size_t)c * NM + m] = log(sv[(size_t)c * NM + m]);
    // ---- runs of equal node-grid cells: heads -> ordered list of run starts
    unsigned hb[ROUNDS];
#pragma unroll
    for (int r = 0; r < ROUNDS; r++) {
        const int m = r * TF_THREADS + tid;
        const bool valid = m < n;
        int key = -2;
        if (valid) {
            const double2 p = sx[m];
            key = tf_cell(p.x, a.z0, a.sz, a.ncz) * a.ncx + tf_cell(p.y, a.x0, a.sx, a.ncx);
        }
        int prev = __shfl_up_sync(full, key, 1);
        if (lane == 0) {
            prev = -1;
            if (valid && m > 0) {
                const double2 p = sx[m - 1];
                prev = tf_cell(p.x, a.z0, a.sz, a.ncz) * a.ncx + tf_cell(p.y, a.x0, a.sx, a.ncx);
            }
        }
        hb[r] = __ballot_sync(full, valid && key != prev);
        if (lane == 0) cnt[r * (TF_THREADS / 32) + warp] = __popc(hb[r]);
    }
    __syncthreads();
    if (warp == 0) {
        const int v = lane < ROUNDS * (TF_THREADS / 32) ? cnt[lane] : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(full, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane < ROUNDS * (TF_THREADS / 32)) cnt[lane] = inc - v;
        if (lane == 31) nrun_s = inc;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < ROUNDS; r++)
        if ((hb[r] >> lane) & 1u)
            rstart[cnt[r * (TF_THREADS / 32) + warp] + __popc(hb[r] & ((1u << lane) - 1u))] = (unsigned short)(r * TF_THREADS + tid);
    if (tid == 0) rstart[nrun_s] = (unsigned short)n;
    __syncthreads();
    // ---- work items (task, run, part); the item count of a task is padded to whole warps, so a warp's 32
    // items share the task (uniform switch) and all its lanes reach the shuffles of tf_run
    const int nrun = nrun_s, nip = (nrun * PARTS + 31) & ~31, total = a.nt * nip;
    for (int q = tid; q < total; q += TF_THREADS) {
        const int ty = q / nip, item = q - ty * nip;
        const int run = item / PARTS, part = item - run * PARTS;
        const bool live = run < nrun;
        const int s = live ? rstart[run] : 0, len = live ? (int)rstart[run + 1] - s : 0;
        const TFTask& t = a.t[ty];
        switch (t.type) {
            case 0:
                switch (t.nf) {
                    case 1: tf_run<2, 2, 1, PARTS, NM>(a, t, sx, sv, s, len, part); break;
                    case 2: tf_run<2, 2, 2, PARTS, NM>(a, t, sx, sv, s, len, part); break;
                    case 3: tf_run<2, 2, 3, PARTS, NM>(a, t, sx, sv, s, len, part); break;
                    case 4: tf_run<2, 2, 4, PARTS, NM>(a, t, sx, sv, s, len, part); break;
                    case 5: tf_run<2, 2, 5, PARTS, NM>(a, t, sx, sv, s, len, part); break;
                    default: tf_run<2, 2, 6, PARTS, NM>(a, t, sx, sv, s, len, part); break;
                }
                break;
            case 1: tf_run<3, 3, 1, PARTS, NM>(a, t, sx, sv, s, len, part); break;
            case 2: tf_run<3, 2, 1, PARTS, NM>(a, t, sx, sv, s, len, part); break;
            default: tf_run<2, 3, 1, PARTS, NM>(a, t, sx, sv, s, len, part); break;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// From raw node sums to finished fields, for one or several targets whose planes lie in one block:
//  * one rank: divide / exponentiate (k_t2g_finalise);
//  * several ranks, replicated fields (any marker on any rank): all-reduce the raw planes first;
//  * several ranks, slab-local fields (plb_ctx_set_slab: a rank's markers lie in its own cell rows, so its
//    sums reach at most one node row into a neighbour's slab): add the boundary rows of the neighbouring
//    slabs (grouped ncclSend/ncclRecv), finish the own rows only, exchange halo rows of the results.
// ---------------------------------------------------------------------------------------------
constexpr int T2G_BOUNDARY_ROWS = 2;

// zero `nplanes` accumulation planes of nze x nxe doubles; with slab-local fields only the rows this rank's markers
// can reach (its own rows and the boundary rows either side)
int t2g_zero_planes(plb_ctx* ctx, double* planes, size_t nplanes, int nze, int nxe, int crop_z0) {
    const size_t plane = (size_t)nze * nxe;
    if (!(plb_comm_size(ctx) > 1 && ctx->slab_on)) {
        PLB_CUDA(ctx, cudaMemsetAsync(planes, 0, nplanes * plane * sizeof(double), ctx->stream));
        return 0;
    }
    const int lo = std::max(ctx->slab_i0 + crop_z0 - T2G_BOUNDARY_ROWS - 1, 0);
    const int hi = std::min(ctx->slab_i1 + crop_z0 + T2G_BOUNDARY_ROWS + 1, nze);
    // (the first rank's markers can lie in the ghost rows below, the last rank's above: keep those)
    const int a = plb_comm_rank(ctx) == 0 ? 0 : lo, b = plb_comm_rank(ctx) == plb_comm_size(ctx) - 1 ? nze : hi;
    for (size_t p = 0; p < nplanes; p++)
        PLB_CUDA(ctx, cudaMemsetAsync(planes + p * plane + (size_t)a * nxe, 0, (size_t)(b - a) * nxe * sizeof(double), ctx->stream));
    return 0;
}

struct T2GFinish {
    T2GArgs a;
    int crop_z0, crop_x0;
};

int t2g_finish(plb_ctx* ctx, int n, T2GFinish* fin, double* planes, size_t nplane_dbl, int nz, int nxx, int ld) {
    const int R = plb_comm_size(ctx);
    int r0 = 0, r1 = nz;
    const bool slab = R > 1 && ctx->slab_on;
    if (slab) {
        if (ctx->slab_i1 > nz) PLB_FAIL(ctx, "trac2grid: slab rows [%d, %d) exceed the %d node rows", ctx->slab_i0, ctx->slab_i1, nz);
        r0 = ctx->slab_i0, r1 = ctx->slab_i1;
        std::vector<double*> arrs;
        std::vector<long long> rd;
        std::vector<int> i0, i1, nr;
        size_t need = 0;
        for (int t = 0; t < n; t++) {
            const T2GArgs& a = fin[t].a;
            auto add = [&](double* p) {
                if (!p) return;
                arrs.push_back(p), rd.push_back(a.nxe), i0.push_back(r0 + fin[t].crop_z0), i1.push_back(r1 + fin[t].crop_z0);
                nr.push_back(a.nze);
                need += 2 * (size_t)T2G_BOUNDARY_ROWS * a.nxe;
            };
            for (int f = 0; f < a.k; f++) add(a.acc[f]);
            add(a.wsum), add(a.cnt);
        }
        if (need > ctx->slab_scratch_dbl) {
            PLB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            if (ctx->slab_scratch) cudaFree(ctx->slab_scratch);
            ctx->slab_scratch = nullptr, ctx->slab_scratch_dbl = 0;
            PLB_CUDA(ctx, cudaMalloc(&ctx->slab_scratch, need * sizeof(double)));
            ctx->slab_scratch_dbl = need;
        }
        if (plb_comm_accumulate_rows(ctx, (int)arrs.size(), arrs.data(), rd.data(), i0.data(), i1.data(), nr.data(),
                                     T2G_BOUNDARY_ROWS, ctx->slab_scratch))
            return 2;
    } else if (R > 1) {
        if (plb_comm_allreduce(ctx, planes, nplane_dbl, PLB_OP_SUM)) return 2;
    }
    std::vector<double*> outs;
    std::vector<long long> ord;
    for (int t = 0; t < n; t++) {
        T2GArgs a = fin[t].a;
        for (int f = 0; f < a.k; f++) {
            outs.push_back(a.out[f]), ord.push_back(ld);
            a.out[f] += (size_t)r0 * ld;
        }
        const int rows = r1 - r0;
        if (rows > 0) {
            k_t2g_finalise<<<plb_grid_for(ctx, (long long)rows * nxx, 256, 8), 256, 0, ctx->stream>>>(
                a, fin[t].crop_z0 + r0, fin[t].crop_x0, rows, nxx, ld);
            PLB_LAUNCHED(ctx);
        }
    }
    if (slab && plb_comm_halo_rows(ctx, (int)outs.size(), outs.data(), ord.data(), r0, r1, ctx->slab_halo)) return 2;
    return 0;
}

template <int K>
void launch_scatter(plb_ctx* ctx, long long M, const double2* x, const T2GArgs& a) {
    int threads = 256;
    int grid = plb_grid_for(ctx, (M + T2G_U - 1) / T2G_U, threads, 8);
    k_t2g_scatter<K><<<grid, threads, 0, ctx->stream>>>(M, x, a);
}

void launch_chunk(plb_ctx* ctx, long long nchunk, const double2* x, const T2GArgs& a) {
    int grid = plb_grid_for(ctx, nchunk, 256, 6);     // 80 registers: 3 resident CTAs per SM, two full waves
    if (a.merge_first) k_t2g_chunk<true><<<grid, 256, 0, ctx->stream>>>(nchunk, x, a);
    else k_t2g_chunk<false><<<grid, 256, 0, ctx->stream>>>(nchunk, x, a);
}

template <int K>
void scatter_k(plb_ctx* ctx, long long M, const double2* x, const T2GArgs& a, bool chunked) {
    if (!chunked) {
        launch_scatter<K>(ctx, M, x, a);
        return;
    }
    const long long nchunk = M / 4, rest = M - 4 * nchunk;
    launch_chunk(ctx, nchunk, x, a);
    if (rest) {       // the < 4 markers after the last whole chunk: generic kernel on the tail
        ctx->launches++;
        T2GArgs t = a;
        for (int f = 0; f < K; f++) t.f[f] = a.f[f] + 4 * nchunk;
        launch_scatter<K>(ctx, rest, x + 4 * nchunk, t);
    }
}

}  // namespace

extern "C" {

int plb_marker_minmax(plb_ctx* ctx, long long M, const double* d_tr_x, double* h_out) {
    if (!ctx || M < 0) return 1;
    PLB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (plb_ws_reserve(ctx, 64)) return 2;
    double* d = (double*)ctx->ws;
    k_minmax_init<<<1, 1, 0, ctx->stream>>>(d);
    PLB_LAUNCHED(ctx);
    if (M > 0) {
        k_marker_minmax<<<plb_grid_for(ctx, M, 256, 8), 256, 0, ctx->stream>>>(M, (const double2*)d_tr_x, d);
        PLB_LAUNCHED(ctx);
    }
    if (plb_comm_size(ctx) > 1) {
        // global extent over all ranks' markers: one MAX all-reduce of (-zmin, zmax, -xmin, xmax)
        k_minmax_flip<<<1, 1, 0, ctx->stream>>>(d);
        PLB_LAUNCHED(ctx);
        if (plb_comm_allreduce(ctx, d, 4, PLB_OP_MAX)) return 2;
        k_minmax_flip<<<1, 1, 0, ctx->stream>>>(d);
        PLB_LAUNCHED(ctx);
    }
    PLB_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned, d, 4 * sizeof(double), cudaMemcpyDeviceToHost,
                                  ctx->stream));
    PLB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    // order: zmin, zmax, xmin, xmax
    for (int i = 0; i < 4; i++) h_out[i] = ctx->h_pinned[i];
    return 0;
}

int plb_trac2grid(plb_ctx* ctx, long long M, const double* d_tr_x, int k,
                  const double* const* h_fields, const int* h_scheme, const double* d_axis_z,
                  int nze, const double* d_axis_x, int nxe, double z0, double zlen, double x0,
                  double xlen, int crop_z0, int crop_x0, int nz, int nxx, int ld,
                  double* const* h_out) {
    if (!ctx) return 1;
    if (k < 1 || k > PLB_MAX_FIELDS) PLB_FAIL(ctx, "plb_trac2grid: k=%d out of range 1..%d", k, PLB_MAX_FIELDS);
    if (crop_z0 + nz > nze || crop_x0 + nxx > nxe) PLB_FAIL(ctx, "plb_trac2grid: crop outside extended grid");
    PLB_CUDA(ctx, cudaSetDevice(ctx->device));
    T2GArgs a;
    memset(&a, 0, sizeof(a));
    bool any_w = false, any_c = false;
    for (int f = 0; f < k; f++) {
        int sc = h_scheme[f];
        if (!(sc & (PLB_AVG_ARITHMETIC | PLB_AVG_GEOMETRIC)))
            PLB_FAIL(ctx, "plb_trac2grid: invalid averaging scheme %d", sc);
        (sc & PLB_AVG_WEIGHTED) ? any_w = true : any_c = true;
        a.f[f] = h_fields[f];
        a.out[f] = h_out[f];
        a.scheme[f] = sc;
    }
    size_t plane = (size_t)nze * nxe;
    size_t nplanes = k + (any_w ? 1 : 0) + (any_c ? 1 : 0);
    // scratch: accumulation planes | {coordinate, reciprocal spacing} tables (16-byte aligned) | reciprocals
    const size_t tab_off = (nplanes * plane + 1) & ~(size_t)1;
    if (plb_ws_reserve(ctx, (tab_off + 3 * (size_t)(nze + nxe)) * sizeof(double))) return 2;
    double* w = (double*)ctx->ws;
    double2* tab = (double2*)(w + tab_off);
    double* recip = w + tab_off + 2 * (size_t)(nze + nxe);
    k_axis_recip<<<plb_blocks(nze, 256), 256, 0, ctx->stream>>>(nze, d_axis_z, recip, tab);
    PLB_LAUNCHED(ctx);
    k_axis_recip<<<plb_blocks(nxe, 256), 256, 0, ctx->stream>>>(nxe, d_axis_x, recip + nze, tab + nze);
    PLB_LAUNCHED(ctx);
    a.riz = recip, a.rix = recip + nze;
    a.tz = tab, a.tx = tab + nze;
    // the chunk kernel needs weighted schemes only, 32-bit plane indices and 32-byte aligned arrays
    a.merge_first = ctx->t2g_variant == 2;
    bool chunked = ctx->t2g_variant >= 1 && !any_c && M >= 4 && plane < ((size_t)1 << 31) &&
                   ((uintptr_t)d_tr_x & 31) == 0;
    for (int f = 0; f < k; f++) chunked = chunked && ((uintptr_t)h_fields[f] & 31) == 0;
    a.sz = (double)(nze - 1) / zlen, a.sx = (double)(nxe - 1) / xlen;
    for (int f = 0; f < k; f++) a.acc[f] = w + (size_t)f * plane;
    size_t nxt = k;
    if (any_w) a.wsum = w + (nxt++) * plane;
    if (any_c) a.cnt = w + (nxt++) * plane;
    a.axz = d_axis_z, a.axx = d_axis_x, a.nze = nze, a.nxe = nxe;
    a.z0 = z0, a.zlen = zlen, a.x0 = x0, a.xlen = xlen, a.k = k;
    if (t2g_zero_planes(ctx, w, nplanes, nze, nxe, crop_z0)) return 2;
    if (M > 0) {
        plb_prof_scope prof_(ctx, PLB_K_T2G, (16.0 + 8.0 * k) * (double)M);
        const double2* x = (const double2*)d_tr_x;
        switch (k) {
            case 1: scatter_k<1>(ctx, M, x, a, chunked); break;
            case 2: scatter_k<2>(ctx, M, x, a, chunked); break;
            case 3: scatter_k<3>(ctx, M, x, a, chunked); break;
            case 4: scatter_k<4>(ctx, M, x, a, chunked); break;
            case 5: scatter_k<5>(ctx, M, x, a, chunked); break;
            case 6: scatter_k<6>(ctx, M, x, a, chunked); break;
            case 7: scatter_k<7>(ctx, M, x, a, chunked); break;
            default: scatter_k<8>(ctx, M, x, a, chunked); break;
        }
        PLB_LAUNCHED(ctx);
    }
    T2GFinish fin;
    fin.a = a, fin.crop_z0 = crop_z0, fin.crop_x0 = crop_x0;
    return t2g_finish(ctx, 1, &fin, w, nplanes * plane, nz, nxx, ld);
}

// The step's marker->grid targets in ONE pass over the markers (k_t2g_fused above).  Every target is
// a (nz x nxx) field on the node grid or on one of its half-staggered companions (kind 1..3: the
// staggered axes hold the midpoints of the node axis, pylamp2.py:92-95, ghost-extended by >= 1 node on
// the low side); all schemes must be weighted (ARITHW / GEOMW); at most TF_MAXC distinct columns.
// Returns 3 (and touches nothing) when the request does not fit these rules: the caller then uses
// plb_trac2grid once per target.
int plb_trac2grid_fused(plb_ctx* ctx, long long M, const double* d_tr_x, int nz, int nxx, int ld, double z0,
                        double zlen, double x0, double xlen, int ntargets, const plb_t2g_target* tg) {
    if (!ctx || !tg || ntargets < 1) return 1;
    PLB_CUDA(ctx, cudaSetDevice(ctx->device));
    TFArgs a;
    memset(&a, 0, sizeof(a));
    const int nfmax = (ctx->t2g_nfmax >= 1 && ctx->t2g_nfmax <= TF_NFMAX) ? ctx->t2g_nfmax : TF_NFMAX;
    // ---- distinct columns, task list, plane layout
    size_t plane_off[8], tab_off[8], ndbl = 0;
    if (ntargets > 8) return 3;
    for (int i = 0; i < ntargets; i++) {
        const plb_t2g_target& T = tg[i];
        if (T.kind < 0 || T.kind > 3 || T.k < 1 || T.k > PLB_MAX_FIELDS) return 3;
        const bool sz_ = T.kind == 1 || T.kind == 2, sx_ = T.kind == 1 || T.kind == 3;
        if ((sz_ && T.crop_z0 < 1) || (sx_ && T.crop_x0 < 1)) return 3;
        if (T.crop_z0 + nz > T.nze || T.crop_x0 + nxx > T.nxe) return 3;
        if ((size_t)T.nze * T.nxe >= ((size_t)1 << 31)) return 3;
        for (int f = 0; f < T.k; f++)
            if ((T.scheme[f] & (PLB_AVG_ARITHMETIC | PLB_AVG_GEOMETRIC | PLB_AVG_WEIGHTED)) != (PLB_AVG_ARITHMETIC | PLB_AVG_WEIGHTED) &&
                (T.scheme[f] & (PLB_AVG_ARITHMETIC | PLB_AVG_GEOMETRIC | PLB_AVG_WEIGHTED)) != (PLB_AVG_GEOMETRIC | PLB_AVG_WEIGHTED))
                return 3;
        plane_off[i] = ndbl;
        ndbl += (size_t)(1 + T.k) * T.nze * T.nxe;
    }
    const size_t nplane_dbl = ndbl;
    ndbl = (ndbl + 1) & ~(size_t)1;
    for (int i = 0; i < ntargets; i++) {
        tab_off[i] = ndbl;
        ndbl += 2 * (size_t)(tg[i].nze + tg[i].nxe);          // double2 tables of both axes
    }
    const size_t recip_off = ndbl;
    size_t maxax = 0;
    for (int i = 0; i < ntargets; i++) maxax = std::max(maxax, (size_t)(tg[i].nze + tg[i].nxe));
    ndbl += maxax;
    bool aligned = ((uintptr_t)d_tr_x & 15) == 0;
    for (int i = 0; i < ntargets; i++) {
        const plb_t2g_target& T = tg[i];
        const bool sz_ = T.kind == 1 || T.kind == 2, sx_ = T.kind == 1 || T.kind == 3;
        const int type = T.kind == 0 ? 0 : (T.kind == 1 ? 1 : (T.kind == 2 ? 2 : 3));
        const int per = type == 0 ? nfmax : 1;                  // fields per task
        for (int f0 = 0; f0 < T.k; f0 += per) {
            if (a.nt >= TF_MAXT) return 3;
            TFTask& t = a.t[a.nt++];
            t.type = type, t.nf = std::min(per, T.k - f0), t.ws = f0 == 0;
            t.lz = T.crop_z0 - (sz_ ? 1 : 0), t.lx = T.crop_x0 - (sx_ ? 1 : 0), t.nxe = T.nxe;
            // the kernel addresses a task's fields as staged columns col[0], col[0]+1, ...: a single-field task
            // shares an already staged column; a multi-field task takes a fresh run of columns (a column that
            // two multi-field tasks share is staged twice -- not a pattern of the time loop)
            for (int f = 0; f < t.nf; f++) {
                const double* p = T.fields[f0 + f];
                const bool lg = !(T.scheme[f0 + f] & PLB_AVG_ARITHMETIC);
                int c = -1;
                if (t.nf == 1)
                    for (int q = 0; q < a.ncol; q++)
                        if (a.col[q] == p && lg == (((a.logmask >> q) & 1u) != 0)) c = q;
                if (c < 0) {
                    if (a.ncol >= TF_MAXC) return 3;
                    c = a.ncol++;
                    a.col[c] = p;
                    if (lg) a.logmask |= 1u << c;
                    aligned = aligned && ((uintptr_t)p & 15) == 0;
                }
                t.col[f] = c;
            }
        }
    }
    if (plb_ws_reserve(ctx, ndbl * sizeof(double))) return 2;
    double* w = (double*)ctx->ws;
    // planes, tables
    for (int i = 0; i < ntargets; i++)
        if (t2g_zero_planes(ctx, w + plane_off[i], (size_t)(1 + tg[i].k), tg[i].nze, tg[i].nxe, tg[i].crop_z0)) return 2;
    int ti = 0;
    for (int i = 0; i < ntargets; i++) {
        const plb_t2g_target& T = tg[i];
        double2* tab = (double2*)(w + tab_off[i]);
        double* recip = w + recip_off;
        k_axis_recip<<<plb_blocks(T.nze, 256), 256, 0, ctx->stream>>>(T.nze, T.axis_z, recip, tab);
        PLB_LAUNCHED(ctx);
        k_axis_recip<<<plb_blocks(T.nxe, 256), 256, 0, ctx->stream>>>(T.nxe, T.axis_x, recip + T.nze, tab + T.nze);
        PLB_LAUNCHED(ctx);
        const size_t plane = (size_t)T.nze * T.nxe;
        const int per = T.kind == 0 ? nfmax : 1;
        for (int f0 = 0; f0 < T.k; f0 += per, ti++) {
            TFTask& t = a.t[ti];
            t.tz = tab, t.tx = tab + T.nze;
            t.wsum = w + plane_off[i];
            for (int f = 0; f < t.nf; f++) t.acc[f] = w + plane_off[i] + (size_t)(1 + f0 + f) * plane;
        }
    }
    a.z0 = z0, a.x0 = x0, a.ncz = nz - 1, a.ncx = nxx - 1;
    a.sz = (double)(nz - 1) / zlen, a.sx = (double)(nxx - 1) / xlen;
    if (M > 0) {
        const int nm = ctx->t2g_nm == 1024 ? 1024 : 960;          // markers per CTA (960 = 60 full cells of 16: <= 64 runs)
        const size_t smem = (size_t)nm * 16 + (size_t)a.ncol * nm * 8 + (nm + 2) * sizeof(unsigned short);
        plb_prof_scope prof_(ctx, PLB_K_T2G, (16.0 + 8.0 * a.ncol) * (double)M);
        const long long nchunk = (M + nm - 1) / nm;
        if (nchunk > 0x7fffffffLL) PLB_FAIL(ctx, "plb_trac2grid_fused: too many markers");
        const int parts = ctx->t2g_parts > 0 ? ctx->t2g_parts : 1;     // lanes per run
        const double2* xx = (const double2*)d_tr_x;
        const unsigned g = (unsigned)nchunk;
        const int tma = aligned ? 1 : 0;
#define TF_LAUNCH(P, N, B)                                                                                           \
    do {                                                                                                             \
        PLB_CUDA(ctx, cudaFuncSetAttribute(k_t2g_fused<P, N, B>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)); \
        k_t2g_fused<P, N, B><<<g, TF_THREADS, smem, ctx->stream>>>(M, xx, a, tma);                                   \
    } while (0)
        // 3 resident CTAs per SM (<= 80 registers) while the staged chunk allows it, else 2
        int nf_big = 0;
        for (int i = 0; i < a.nt; i++) nf_big = std::max(nf_big, a.t[i].nf);
        // (measured, 2048^2: 6 fields per node item want the registers of 2 CTAs/SM -- 2.48 vs 3.29 ms; <= 3 fields
        // per item run as fast with 3 CTAs/SM, and a single-field call is 25 % faster with them)
        const bool three = (ctx->t2g_minb == 3 || (ctx->t2g_minb != 2 && nf_big <= 3)) && 3 * (smem + 1280) <= 227 * 1024;
        if (nm == 1024) {
            if (parts >= 4) TF_LAUNCH(4, 1024, 2);
            else if (parts == 2) TF_LAUNCH(2, 1024, 2);
            else TF_LAUNCH(1, 1024, 2);
        } else if (three) {
            if (parts >= 4) TF_LAUNCH(4, 960, 3);
            else if (parts == 2) TF_LAUNCH(2, 960, 3);
            else TF_LAUNCH(1, 960, 3);
        } else {
            if (parts >= 4) TF_LAUNCH(4, 960, 2);
            else if (parts == 2) TF_LAUNCH(2, 960, 2);
            else TF_LAUNCH(1, 960, 2);
        }
#undef TF_LAUNCH
        PLB_LAUNCHED(ctx);
    }
    T2GFinish fin[8];
    for (int i = 0; i < ntargets; i++) {
        const plb_t2g_target& T = tg[i];
        T2GArgs& fa = fin[i].a;
        memset(&fa, 0, sizeof(fa));
        const size_t plane = (size_t)T.nze * T.nxe;
        fa.wsum = w + plane_off[i];
        for (int f = 0; f < T.k; f++) {
            fa.acc[f] = w + plane_off[i] + (size_t)(1 + f) * plane;
            fa.out[f] = T.out[f];
            fa.scheme[f] = T.scheme[f];
        }
        fa.nze = T.nze, fa.nxe = T.nxe, fa.k = T.k;
        fin[i].crop_z0 = T.crop_z0, fin[i].crop_x0 = T.crop_x0;
    }
    return t2g_finish(ctx, ntargets, fin, w, nplane_dbl, nz, nxx, ld);
}

// plb_trac2grid in two halves for slab-owned markers (pylamp_b200/slabgrid.py): the raw sums land in
// caller-owned planes [field 0 .. k-1 | sum of weights | marker count] (the last two only if a
// weighted / an unweighted scheme is present), nothing is all-reduced and nothing divided; the
// caller combines the boundary rows of neighbouring slabs and then finalises a row range.
int plb_trac2grid_scatter(plb_ctx* ctx, long long M, const double* d_tr_x, int k,
                          const double* const* h_fields, const int* h_scheme, const double* d_axis_z,
                          int nze, const double* d_axis_x, int nxe, double z0, double zlen, double x0,
                          double xlen, double* d_planes, int* h_nplanes) {
    if (!ctx || !d_planes) return 1;
    if (k < 1 || k > PLB_MAX_FIELDS) PLB_FAIL(ctx, "plb_trac2grid_scatter: k=%d out of range 1..%d", k, PLB_MAX_FIELDS);
    PLB_CUDA(ctx, cudaSetDevice(ctx->device));
    T2GArgs a;
    memset(&a, 0, sizeof(a));
    bool any_w = false, any_c = false;
    for (int f = 0; f < k; f++) {
        int sc = h_scheme[f];
        if (!(sc & (PLB_AVG_ARITHMETIC | PLB_AVG_GEOMETRIC)))
            PLB_FAIL(ctx, "plb_trac2grid_scatter: invalid averaging scheme %d", sc);
        (sc & PLB_AVG_WEIGHTED) ? any_w = true : any_c = true;
        a.f[f] = h_fields[f];
        a.scheme[f] = sc;
    }
    const size_t plane = (size_t)nze * nxe;
    const size_t nplanes = k + (any_w ? 1 : 0) + (any_c ? 1 : 0);
    if (h_nplanes) *h_nplanes = (int)nplanes;
    if (plb_ws_reserve(ctx, 3 * (size_t)(nze + nxe) * sizeof(double))) return 2;
    double2* tab = (double2*)ctx->ws;
    double* recip = (double*)ctx->ws + 2 * (size_t)(nze + nxe);
    k_axis_recip<<<plb_blocks(nze, 256), 256, 0, ctx->stream>>>(nze, d_axis_z, recip, tab);
    PLB_LAUNCHED(ctx);
    k_axis_recip<<<plb_blocks(nxe, 256), 256, 0, ctx->stream>>>(nxe, d_axis_x, recip + nze, tab + nze);
    PLB_LAUNCHED(ctx);
    a.riz = recip, a.rix = recip + nze, a.tz = tab, a.tx = tab + nze;
    a.sz = (double)(nze - 1) / zlen, a.sx = (double)(nxe - 1) / xlen;
    for (int f = 0; f < k; f++) a.acc[f] = d_planes + (size_t)f * plane;
    size_t nxt = k;
    if (any_w) a.wsum = d_planes + (nxt++) * plane;
    if (any_c) a.cnt = d_planes + (nxt++) * plane;
    a.axz = d_axis_z, a.axx = d_axis_x, a.nze = nze, a.nxe = nxe;
    a.z0 = z0, a.zlen = zlen, a.x0 = x0, a.xlen = xlen, a.k = k;
    a.merge_first = ctx->t2g_variant == 2;
    bool chunked = ctx->t2g_variant >= 1 && !any_c && M >= 4 && plane < ((size_t)1 << 31) &&
                   ((uintptr_t)d_tr_x & 31) == 0;
    for (int f = 0; f < k; f++) chunked = chunked && ((uintptr_t)h_fields[f] & 31) == 0;
    PLB_CUDA(ctx, cudaMemsetAsync(d_planes, 0, nplanes * plane * sizeof(double), ctx->stream));
    if (M > 0) {
        plb_prof_scope prof_(ctx, PLB_K_T2G, (16.0 + 8.0 * k) * (double)M);
        const double2* x = (const double2*)d_tr_x;
        switch (k) {
            case 1: scatter_k<1>(ctx, M, x, a, chunked); break;
            case 2: scatter_k<2>(ctx, M, x, a, chunked); break;
            case 3: scatter_k<3>(ctx, M, x, a, chunked); break;
            case 4: scatter_k<4>(ctx, M, x, a, chunked); break;
            case 5: scatter_k<5>(ctx, M, x, a, chunked); break;
            case 6: scatter_k<6>(ctx, M, x, a, chunked); break;
            case 7: scatter_k<7>(ctx, M, x, a, chunked); break;
            default: scatter_k<8>(ctx, M, x, a, chunked); break;
        }
        PLB_LAUNCHED(ctx);
    }
    return 0;
}

// rows [row0, row1) of the (nz x nxx) target from raw planes laid out as by plb_trac2grid_scatter
int plb_trac2grid_finalise(plb_ctx* ctx, int k, const int* h_scheme, const double* d_planes, int nze, int nxe,
                           int crop_z0, int crop_x0, int nz, int nxx, int ld, int row0, int row1,
                           double* const* h_out) {
    if (!ctx || !d_planes) return 1;
    if (k < 1 || k > PLB_MAX_FIELDS) PLB_FAIL(ctx, "plb_trac2grid_finalise: k=%d out of range", k);
    if (row0 < 0 || row1 > nz || row0 > row1 || crop_z0 + nz > nze || crop_x0 + nxx > nxe)
        PLB_FAIL(ctx, "plb_trac2grid_finalise: rows/crop outside the grids");
    PLB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (row1 == row0) return 0;
    T2GArgs a;
    memset(&a, 0, sizeof(a));
    bool any_w = false, any_c = false;
    for (int f = 0; f < k; f++) (h_scheme[f] & PLB_AVG_WEIGHTED) ? any_w = true : any_c = true;
    const size_t plane = (size_t)nze * nxe;
    double* planes = const_cast<double*>(d_planes);
    for (int f = 0; f < k; f++) {
        a.scheme[f] = h_scheme[f];
        a.acc[f] = planes + (size_t)f * plane;
        a.out[f] = h_out[f] + (size_t)row0 * ld;
    }
    size_t nxt = k;
    if (any_w) a.wsum = planes + (nxt++) * plane;
    if (any_c) a.cnt = planes + (nxt++) * plane;
    a.nze = nze, a.nxe = nxe, a.k = k;
    const int rows = row1 - row0;
    k_t2g_finalise<<<plb_grid_for(ctx, (long long)rows * nxx, 256, 8), 256, 0, ctx->stream>>>(
        a, crop_z0 + row0, crop_x0, rows, nxx, ld);
    PLB_LAUNCHED(ctx);
    return 0;
}

int plb_grid2trac(plb_ctx* ctx, long long M, const double* d_tr_x, int method, int k,
                  const double* const* h_fields, const double* d_grid_z, int nz,
                  const double* d_grid_x, int nxx, int ld, double z0, double zlen, double x0,
                  double xlen, double defval, double* const* h_out, long long* h_n_outside) {
    if (!ctx) return 1;
    if (k < 1 || k > PLB_MAX_FIELDS) PLB_FAIL(ctx, "plb_grid2trac: k=%d out of range", k);
    if (!(method & (PLB_METHOD_NEAREST | PLB_METHOD_LINEAR | PLB_METHOD_VELDIV)))
        PLB_FAIL(ctx, "plb_grid2trac: unknown method %d", method);
    if ((method & PLB_METHOD_VELDIV) && !(method & (PLB_METHOD_NEAREST | PLB_METHOD_LINEAR)) && k != 2)
        PLB_FAIL(ctx, "grid2trac(): method INTERP_METHOD_VELDIV only works in 2D and expects field to be (vz,vx)");
    PLB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (plb_ws_reserve(ctx, 64)) return 2;
    unsigned long long* d_bad = (unsigned long long*)ctx->ws;
    PLB_CUDA(ctx, cudaMemsetAsync(d_bad, 0, sizeof(unsigned long long), ctx->stream));
    G2TGrid g = {d_grid_z, d_grid_x, nz, nxx, ld, z0, zlen, x0, xlen};
    G2TArgs a;
    memset(&a, 0, sizeof(a));
    a.k = k;
    for (int f = 0; f < k; f++) a.f[f] = h_fields[f], a.out[f] = h_out[f];
    if (M > 0) {
        plb_prof_scope prof_(ctx, PLB_K_G2T, (16.0 + 8.0 * k) * (double)M);
        k_grid2trac<<<plb_grid_for(ctx, M, 256, 8), 256, 0, ctx->stream>>>(
            M, (const double2*)d_tr_x, method, g, a, defval, d_bad);
        PLB_LAUNCHED(ctx);
    }
    if (h_n_outside) {
        unsigned long long* hp = (unsigned long long*)ctx->h_pinned;
        PLB_CUDA(ctx, cudaMemcpyAsync(hp, d_bad, sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                                      ctx->stream));
        PLB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        *h_n_outside = (long long)*hp;
    }
    return 0;
}

namespace {
int rk4_launch(plb_ctx* ctx, long long M, const double* d_tr_x, const double* d_vz_c, const double* d_vx_c,
               const double* d_gc_z, int nzc, const double* d_gc_x, int nxc, int ld, double z0, double zlen, double x0,
               double xlen, double dt, double* d_x_out, double* d_v_out, const FenceArgs* fence) {
    PLB_CUDA(ctx, cudaSetDevice(ctx->device));
    G2TGrid g = {d_gc_z, d_gc_x, nzc, nxc, ld, z0, zlen, x0, xlen};
    if ((long long)nzc * ld >= (1LL << 31)) PLB_FAIL(ctx, "plb_rk4: velocity fields beyond 2^31 nodes");
    if (fence && fence->count)
        PLB_CUDA(ctx, cudaMemsetAsync(fence->count, 0, (size_t)(fence->nz - 1) * (fence->nxx - 1) * sizeof(long long), ctx->stream));
    if (M <= 0) return 0;
    // reciprocal cell sizes of both axes, once per call instead of twice per marker and stage
    if (plb_ws_reserve(ctx, (size_t)(nzc + nxc) * sizeof(double))) return 2;
    double* recip = (double*)ctx->ws;
    k_axis_recip<<<plb_blocks(nzc, 256), 256, 0, ctx->stream>>>(nzc, d_gc_z, recip, nullptr);
    PLB_LAUNCHED(ctx);
    k_axis_recip<<<plb_blocks(nxc, 256), 256, 0, ctx->stream>>>(nxc, d_gc_x, recip + nzc, nullptr);
    PLB_LAUNCHED(ctx);
    plb_prof_scope prof_(ctx, PLB_K_RK4, (d_v_out ? 48.0 : 32.0) * (double)M);
    FenceArgs fa;
    memset(&fa, 0, sizeof(fa));
    const int grid = plb_grid_for(ctx, M, 256, 8);
    if (fence) {
        fa = *fence;
        k_rk4<true><<<grid, 256, 0, ctx->stream>>>(M, (const double2*)d_tr_x, d_vz_c, d_vx_c, g, recip, recip + nzc, dt,
                                                   (double2*)d_x_out, (double2*)d_v_out, fa);
    } else {
        k_rk4<false><<<grid, 256, 0, ctx->stream>>>(M, (const double2*)d_tr_x, d_vz_c, d_vx_c, g, recip, recip + nzc, dt,
                                                    (double2*)d_x_out, (double2*)d_v_out, fa);
    }
    PLB_LAUNCHED(ctx);
    return 0;
}
}  // namespace

int plb_rk4(plb_ctx* ctx, long long M, const double* d_tr_x, const double* d_vz_c,
            const double* d_vx_c, const double* d_gc_z, int nzc, const double* d_gc_x, int nxc,
            int ld, double z0, double zlen, double x0, double xlen, double dt, double* d_x_out,
            double* d_v_out) {
    if (!ctx) return 1;
    return rk4_launch(ctx, M, d_tr_x, d_vz_c, d_vx_c, d_gc_z, nzc, d_gc_x, nxc, ld, z0, zlen, x0, xlen, dt, d_x_out, d_v_out,
                      nullptr);
}

// plb_rk4 followed by plb_fence_count on the new positions, in one pass (pylamp2.py:550, :558-572, :588-593)
int plb_rk4_fence_count(plb_ctx* ctx, long long M, const double* d_tr_x, const double* d_vz_c, const double* d_vx_c,
                        const double* d_gc_z, int nzc, const double* d_gc_x, int nxc, int ld, double z0, double zlen,
                        double x0, double xlen, double dt, double* d_x_out, double* d_v_out, double Lz, double Lx,
                        double eps, int nz, int nxx, long long* d_kelem, long long* d_count) {
    if (!ctx) return 1;
    FenceArgs fa = {Lz, Lx, eps, nz, nxx, d_kelem, (unsigned long long*)d_count};
    return rk4_launch(ctx, M, d_tr_x, d_vz_c, d_vx_c, d_gc_z, nzc, d_gc_x, nxc, ld, z0, zlen, x0, xlen, dt, d_x_out, d_v_out, &fa);
}

int plb_fence_walls(plb_ctx* ctx, long long M, double* d_tr_x, double Lz, double Lx, double eps, int walls) {
    if (!ctx) return 1;
    PLB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (M > 0) {
        k_fence<<<plb_grid_for(ctx, M, 256, 8), 256, 0, ctx->stream>>>(M, (double2*)d_tr_x, Lz, Lx, eps, walls);
        PLB_LAUNCHED(ctx);
    }
    return 0;
}

int plb_fence(plb_ctx* ctx, long long M, double* d_tr_x, double Lz, double Lx, double eps) {
    return plb_fence_walls(ctx, M, d_tr_x, Lz, Lx, eps, 15);
}

int plb_cell_index_count(plb_ctx* ctx, long long M, const double* d_tr_x, int nz, int nxx,
                         double Lz, double Lx, long long* d_kelem, long long* d_count) {
    if (!ctx) return 1;
    PLB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (d_count)
        PLB_CUDA(ctx, cudaMemsetAsync(d_count, 0, (size_t)(nz - 1) * (nxx - 1) * sizeof(long long),
                                      ctx->stream));
    if (M > 0) {
        k_cell_index_count<<<plb_grid_for(ctx, M, 256, 8), 256, 0, ctx->stream>>>(
            M, (const double2*)d_tr_x, nz, nxx, Lz, Lx, d_kelem, (unsigned long long*)d_count);
        PLB_LAUNCHED(ctx);
    }
    return 0;
}

int plb_fence_count(plb_ctx* ctx, long long M, double* d_tr_x, double Lz, double Lx, double eps, int nz,
                    int nxx, long long* d_kelem, long long* d_count) {
    if (!ctx) return 1;
    PLB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (d_count)
        PLB_CUDA(ctx, cudaMemsetAsync(d_count, 0, (size_t)(nz - 1) * (nxx - 1) * sizeof(long long),
                                      ctx->stream));
    if (M > 0) {
        k_fence_count<<<plb_grid_for(ctx, M, 256, 8), 256, 0, ctx->stream>>>(
            M, (double2*)d_tr_x, Lz, Lx, eps, nz, nxx, d_kelem, (unsigned long long*)d_count);
        PLB_LAUNCHED(ctx);
    }
    return 0;
}

int plb_update_properties(plb_ctx* ctx, long long M, int tdep_rho, int tdep_eta, double Tref,
                          double etamin, double etamax, double gasr, const double* d_T,
                          const double* d_rho0, const double* d_alpha, const double* d_Ea,
                          const double* d_eta0, double* d_rho, double* d_eta) {
    if (!ctx) return 1;
    PLB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (M > 0) {
        k_update_properties<<<plb_grid_for(ctx, M, 256, 8), 256, 0, ctx->stream>>>(
            M, tdep_rho, tdep_eta, Tref, etamin, etamax, gasr, d_T, d_rho0, d_alpha, d_Ea, d_eta0,
            d_rho, d_eta);
        PLB_LAUNCHED(ctx);
    }
    return 0;
}

int plb_subgrid_stage1(plb_ctx* ctx, long long M, double dt, double dz, double dx,
                       const double* d_Told, const double* d_T, const double* d_cp,
                       const double* d_rho, const double* d_k, double* d_Tsg, double* d_dT) {
    if (!ctx) return 1;
    PLB_CUDA(ctx, cudaSetDevice(ctx->device));
    double fac = (2 / dx) * (2 / dx) + (2 / dz) * (2 / dz);                    // pylamp2.py:473
    if (M > 0) {
        k_subgrid1<<<plb_grid_for(ctx, M, 256, 8), 256, 0, ctx->stream>>>(M, dt, fac, d_Told, d_T, d_cp,
                                                                       d_rho, d_k, d_Tsg, d_dT);
        PLB_LAUNCHED(ctx);
    }
    return 0;
}

int plb_subgrid_stage2(plb_ctx* ctx, long long M, const double* d_Tsg, const double* d_back,
                       double* d_T) {
    if (!ctx) return 1;
    PLB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (M > 0) {
        k_sub<<<plb_grid_for(ctx, M, 256, 8), 256, 0, ctx->stream>>>(M, d_Tsg, d_back, d_T);   // :480
        PLB_LAUNCHED(ctx);
    }
    return 0;
}

int plb_subgrid_fused(plb_ctx* ctx, int stage, long long M, const double* d_tr_x, const double* d_field,
                      const double* d_grid_z, int nz, const double* d_grid_x, int nxx, int ld, double z0,
                      double zlen, double x0, double xlen, double dt, double dz, double dx, double* d_T,
                      const double* d_cp, const double* d_rho, const double* d_k, double* d_Tsg, double* d_dT,
                      long long* h_n_outside) {
    if (!ctx) return 1;
    PLB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (plb_ws_reserve(ctx, 64 + (size_t)(nz + nxx) * sizeof(double))) return 2;
    unsigned long long* d_bad = (unsigned long long*)ctx->ws;
    double* recip = (double*)ctx->ws + 8;
    PLB_CUDA(ctx, cudaMemsetAsync(d_bad, 0, sizeof(unsigned long long), ctx->stream));
    G2TGrid g = {d_grid_z, d_grid_x, nz, nxx, ld, z0, zlen, x0, xlen};
    if ((long long)nz * ld >= (1LL << 31)) PLB_FAIL(ctx, "plb_subgrid_fused: field beyond 2^31 nodes");
    if (M > 0) {
        k_axis_recip<<<plb_blocks(nz, 256), 256, 0, ctx->stream>>>(nz, d_grid_z, recip, nullptr);
        PLB_LAUNCHED(ctx);
        k_axis_recip<<<plb_blocks(nxx, 256), 256, 0, ctx->stream>>>(nxx, d_grid_x, recip + nz, nullptr);
        PLB_LAUNCHED(ctx);
        plb_prof_scope prof_(ctx, PLB_K_G2T, (stage == 1 ? 64.0 : 32.0) * (double)M);
        const int grid = plb_grid_for(ctx, M, 256, 8);
        if (stage == 1) {
            const double fac = (2 / dx) * (2 / dx) + (2 / dz) * (2 / dz);        // pylamp2.py:473
            k_subgrid_fused1<<<grid, 256, 0, ctx->stream>>>(M, (const double2*)d_tr_x, g, recip, recip + nz, d_field, dt, fac,
                                                            d_T, d_cp, d_rho, d_k, d_Tsg, d_dT, d_bad);
        } else {
            k_subgrid_fused2<<<grid, 256, 0, ctx->stream>>>(M, (const double2*)d_tr_x, g, recip, recip + nz, d_field, d_Tsg, d_T,
                                                            d_bad);
        }
        PLB_LAUNCHED(ctx);
    }
    if (h_n_outside) {
        unsigned long long* hp = (unsigned long long*)ctx->h_pinned;
        PLB_CUDA(ctx, cudaMemcpyAsync(hp, d_bad, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
        PLB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        *h_n_outside = (long long)*hp;
    }
    return 0;
}

}  // extern "C"
