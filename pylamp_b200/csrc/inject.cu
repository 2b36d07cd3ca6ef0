// inject.cu -- marker injection into under-populated cells on the device (pylamp2.py:594-633).
//
// The reference loops over the cells with fewer than `tracdens_min` markers in ascending cell order and appends
// `tracdens - count` markers per cell at np.random positions inside the cell, with the mean properties of the
// cell's existing markers (0/0 = NaN for an empty cell) and ids that continue from the running maximum (each
// cell's first new id repeats it, :614-615).  Here:
//   plan : per cell the number of markers to add (cells [cell0, cell1) only: a slab serves its own rows), three
//          exclusive scans over the cells -- slot of every deficient cell, first new marker of every cell, id offset
//          of every cell -- and the totals read back;
//   sums : one pass over the markers adds the properties of those that sit in a deficient cell to that cell's slot;
//   fill : one thread per new marker finds its cell (binary search in the scan), draws its position from a
//          counter-based Philox4x32-10 stream (seed, marker number: reproducible, order-independent), and writes
//          coordinates, cell-mean properties and id behind the existing markers.
// Counts, cells, properties and ids equal the reference's; positions differ (its global Mersenne-Twister stream
// cannot be reproduced on a device), which is the parity SURVEY.md 8f-1 defines.
#include <algorithm>

#include "common.cuh"

int plb_exclusive_scan_u32(plb_ctx* ctx, long long n, const unsigned* in, unsigned* out, unsigned* tiles);

struct plb_inject_ws {
    unsigned *need = nullptr, *flag = nullptr, *idn = nullptr;     // per cell: markers to add, deficient (0/1), need - 1
    unsigned *off = nullptr, *slot = nullptr, *idoff = nullptr;    // their exclusive scans (ncell + 1 entries)
    unsigned* tiles = nullptr;
    long long ncell = 0;
    double* sums = nullptr;        // [ncols][ncell_m]
    size_t cap_sums = 0;
    long long n_def = 0, n_new = 0;
    int cell0 = 0, cell1 = 0;
};

namespace {

__global__ void __launch_bounds__(256)
k_inj_need(long long ncell, const long long* __restrict__ count, int tracdens, int tracdens_min, long long c0, long long c1,
           unsigned* __restrict__ need, unsigned* __restrict__ flag, unsigned* __restrict__ idn) {
    for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < ncell; c += (long long)gridDim.x * blockDim.x) {
        const long long n = count[c];
        const bool d = c >= c0 && c < c1 && n < tracdens_min;
        const unsigned add = d ? (unsigned)(tracdens - n) : 0u;
        need[c] = add, flag[c] = d ? 1u : 0u, idn[c] = d ? add - 1u : 0u;
    }
}

struct InjCols {
    double* p[16];
    int n;
};

__global__ void __launch_bounds__(256)
k_inj_sums(long long M, const long long* __restrict__ kelem, const unsigned* __restrict__ flag, const unsigned* __restrict__ slot,
           long long ncell, long long n_def, InjCols C, double* __restrict__ sums) {
    for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < M; m += (long long)gridDim.x * blockDim.x) {
        const long long c = kelem[m];
        if (c < 0 || c >= ncell || !flag[c]) continue;
        const unsigned s = slot[c];
        for (int a = 0; a < C.n; a++) atomicAdd(sums + (size_t)a * n_def + s, C.p[a][m]);
    }
}

// Philox4x32-10 (Salmon et al. 2011): counter (t, 0, 0, 0), key (seed lo, seed hi)
__device__ __forceinline__ void philox4x32_10(unsigned long long t, unsigned long long seed, unsigned (&r)[4]) {
    unsigned c0 = (unsigned)t, c1 = (unsigned)(t >> 32), c2 = 0, c3 = 0;
    unsigned k0 = (unsigned)seed, k1 = (unsigned)(seed >> 32);
#pragma unroll
    for (int i = 0; i < 10; i++) {
        const unsigned long long p0 = 0xD2511F53ull * c0, p1 = 0xCD9E8D57ull * c2;
        const unsigned n0 = (unsigned)(p1 >> 32) ^ c1 ^ k0, n1 = (unsigned)p1;
        const unsigned n2 = (unsigned)(p0 >> 32) ^ c3 ^ k1, n3 = (unsigned)p0;
        c0 = n0, c1 = n1, c2 = n2, c3 = n3;
        k0 += 0x9E3779B9u, k1 += 0xBB67AE85u;
    }
    r[0] = c0, r[1] = c1, r[2] = c2, r[3] = c3;
}

__device__ __forceinline__ double u01(unsigned hi, unsigned lo) {      // uniform in [0, 1), 53 bits
    return (double)((((unsigned long long)hi << 32) | lo) >> 11) * (1.0 / 9007199254740992.0);
}

__global__ void __launch_bounds__(256)
k_inj_fill(long long n_new, long long M, long long ncell, const unsigned* __restrict__ off, const unsigned* __restrict__ slot,
           const unsigned* __restrict__ idoff, const long long* __restrict__ count, long long n_def,
           const double* __restrict__ sums, InjCols C, int id_col, double id_start, double2* __restrict__ trx,
           const double* __restrict__ gz, const double* __restrict__ gx, int ncx, unsigned long long seed,
           unsigned long long stream0) {
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n_new; t += (long long)gridDim.x * blockDim.x) {
        // cell of new marker t: the last c with off[c] <= t (cells without need have off[c] == off[c+1])
        long long lo = 0, hi = ncell;                  // invariant: off[lo] <= t < off[hi]
        while (hi - lo > 1) {
            const long long mid = (lo + hi) >> 1;
            if (off[mid] <= t) lo = mid; else hi = mid;
        }
        const long long c = lo;
        const unsigned s = slot[c];
        const long long within = t - off[c];
        const int i = (int)(c / ncx), j = (int)(c % ncx);
        unsigned r[4];
        philox4x32_10(stream0 + (unsigned long long)t, seed, r);
        double2 p;
        p.x = u01(r[0], r[1]) * (gz[i + 1] - gz[i]) + gz[i];           // pylamp2.py:618-619
        p.y = u01(r[2], r[3]) * (gx[j + 1] - gx[j]) + gx[j];
        trx[M + t] = p;
        const double cnt = (double)count[c];
        for (int a = 0; a < C.n; a++) {
            double v = sums[(size_t)a * n_def + s] / cnt;              // 0/0 = NaN for an empty cell, like the reference
            if (a == id_col) v = id_start + (double)idoff[c] + (double)within;     // :614-615
            C.p[a][M + t] = v;
        }
    }
}

template <typename T>
int regrow(plb_ctx* ctx, T** p, size_t n) {
    if (*p) cudaFree(*p);
    *p = nullptr;
    PLB_CUDA(ctx, cudaMalloc(p, n * sizeof(T)));
    return 0;
}

}  // namespace

extern "C" {

// h_out[2] = {cells below tracdens_min among the cells [cell0, cell1), markers to add}.  Synchronises.
int plb_inject_plan(plb_ctx* ctx, long long ncell, const long long* d_count, int tracdens, int tracdens_min,
                    long long cell0, long long cell1, long long* h_out) {
    if (!ctx || !d_count || !h_out) return 1;
    if (ncell >= (1LL << 31)) PLB_FAIL(ctx, "plb_inject_plan: too many cells");
    PLB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->inj) ctx->inj = new plb_inject_ws();
    plb_inject_ws* w = ctx->inj;
    if (w->ncell != ncell) {
        PLB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        const size_t n1 = (size_t)ncell + 1;
        if (regrow(ctx, &w->need, n1) || regrow(ctx, &w->flag, n1) || regrow(ctx, &w->idn, n1) || regrow(ctx, &w->off, n1) ||
            regrow(ctx, &w->slot, n1) || regrow(ctx, &w->idoff, n1) || regrow(ctx, &w->tiles, n1 / 4096 + 8))
            return 2;
        w->ncell = ncell;
    }
    k_inj_need<<<plb_grid_for(ctx, ncell, 256, 8), 256, 0, ctx->stream>>>(ncell, d_count, tracdens, tracdens_min, cell0, cell1,
                                                                        w->need, w->flag, w->idn);
    PLB_LAUNCHED(ctx);
    const size_t tl = (size_t)ncell / 4096 + 8;
    const unsigned* in[3] = {w->need, w->flag, w->idn};
    unsigned* out[3] = {w->off, w->slot, w->idoff};
    for (int q = 0; q < 3; q++) {
        PLB_CUDA(ctx, cudaMemsetAsync(w->tiles, 0, tl * sizeof(unsigned), ctx->stream));
        if (plb_exclusive_scan_u32(ctx, ncell, in[q], out[q], w->tiles)) return 2;
    }
    unsigned* hp = (unsigned*)(ctx->h_pinned + 24);
    PLB_CUDA(ctx, cudaMemcpyAsync(hp, w->off + ncell, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
    PLB_CUDA(ctx, cudaMemcpyAsync(hp + 1, w->slot + ncell, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
    PLB_CUDA(ctx, cudaMemcpyAsync(hp + 2, w->idoff + ncell, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
    PLB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    w->n_new = hp[0], w->n_def = hp[1];
    h_out[0] = w->n_def, h_out[1] = w->n_new;
    h_out[2] = hp[2];                       // sum of (need - 1): what this rank's cells advance the ids by
    return 0;
}

// Appends the planned markers behind the M existing ones: d_tr_x and the ncols columns must have room for
// M + n_new rows.  id_col: index (in h_cols) of the id column (-1: none); id_start: the first new id
// (the current maximum id, pylamp2.py:614, plus what lower ranks' cells add).
int plb_inject_apply(plb_ctx* ctx, long long M, const long long* d_kelem, const long long* d_count, double* d_tr_x,
                     int ncols, double* const* h_cols, int id_col, double id_start, const double* d_grid_z,
                     const double* d_grid_x, int nxx, unsigned long long seed, unsigned long long stream0) {
    if (!ctx || !ctx->inj) return 1;
    plb_inject_ws* w = ctx->inj;
    if (ncols < 0 || ncols > 16) PLB_FAIL(ctx, "plb_inject_apply: at most 16 distinct columns");
    PLB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (w->n_new == 0) return 0;
    InjCols Cc;
    Cc.n = ncols;
    for (int a = 0; a < ncols; a++) Cc.p[a] = h_cols[a];
    const size_t ns = (size_t)std::max(1, ncols) * (size_t)w->n_def;
    if (ns > w->cap_sums) {
        PLB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (regrow(ctx, &w->sums, ns + ns / 4 + 64)) return 2;
        w->cap_sums = ns + ns / 4 + 64;
    }
    PLB_CUDA(ctx, cudaMemsetAsync(w->sums, 0, ns * sizeof(double), ctx->stream));
    if (M > 0 && ncols > 0) {
        k_inj_sums<<<plb_grid_for(ctx, M, 256, 8), 256, 0, ctx->stream>>>(M, d_kelem, w->flag, w->slot, w->ncell, w->n_def, Cc,
                                                                        w->sums);
        PLB_LAUNCHED(ctx);
    }
    k_inj_fill<<<plb_grid_for(ctx, w->n_new, 256, 8), 256, 0, ctx->stream>>>(w->n_new, M, w->ncell, w->off, w->slot, w->idoff,
                                                                           d_count, w->n_def, w->sums, Cc, id_col, id_start,
                                                                           (double2*)d_tr_x, d_grid_z, d_grid_x, nxx - 1, seed,
                                                                           stream0);
    PLB_LAUNCHED(ctx);
    return 0;
}

void plb_inject_free(plb_ctx* ctx) {
    if (!ctx || !ctx->inj) return;
    plb_inject_ws* w = ctx->inj;
    void* ptrs[] = {w->need, w->flag, w->idn, w->off, w->slot, w->idoff, w->tiles, w->sums};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    delete w;
    ctx->inj = nullptr;
}

}  // extern "C"
