// sort.cu -- marker-by-cell counting sort on the device (north_star: "after a marker-by-cell sort").
//
// The reference never orders its markers (np.add.at does not care, pylamp_trac.py:257-298); on the
// GPU the order decides how well the marker->node sums aggregate (one reduction per run of markers
// of one cell instead of one per marker) and how local the grid gathers of grid2trac / RK4 are.
// Markers move < 0.45 cells per step, so the cloud is re-sorted every few steps:
//   1. key = cell index  ie*(nxx-1)+je  with the exact floor((n-1)*x/L) of pylamp2.py:588-589
//      (markers outside the box are clamped into the nearest cell) + per-cell histogram,
//   2. exclusive scan of the histogram -> first slot of every cell,
//   3. slot of every marker = first slot of its cell + its rank among the cell's markers
//      (warp-aggregated: equal keys inside a warp take ONE atomic and are ranked by lane),
//   4. every marker array is permuted by a scatter through the slot table (plb_permute).
// Results of all marker kernels are order-independent (up to the fp64 summation order in trac2grid).
#include "common.cuh"

namespace {

__device__ __forceinline__ int clampi(long long v, int hi) { return v < 0 ? 0 : (v > hi ? hi : (int)v); }

__global__ void __launch_bounds__(256)
k_sort_keys(long long M, const double2* __restrict__ trx, int nz, int nxx, double Lz, double Lx,
            unsigned* __restrict__ key, unsigned* __restrict__ count) {
    for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < M; m += (long long)gridDim.x * blockDim.x) {
        const double2 p = trx[m];
        // pylamp2.py:588-589: floor((n-1)*x/L) -- multiply, then divide (markers.cu is built without FMA
        // contraction; here the intrinsics make the order explicit)
        const long long ie = (long long)floor(__ddiv_rn(__dmul_rn((double)(nz - 1), p.x), Lz));
        const long long je = (long long)floor(__ddiv_rn(__dmul_rn((double)(nxx - 1), p.y), Lx));
        const unsigned k = (unsigned)clampi(ie, nz - 2) * (unsigned)(nxx - 1) + (unsigned)clampi(je, nxx - 2);
        key[m] = k;
        // equal keys in a warp: one atomic for the group
        const unsigned act = __activemask();
        const unsigned grp = __match_any_sync(act, k);
        if ((int)(__ffs(grp) - 1) == (int)(threadIdx.x & 31)) atomicAdd(count + k, (unsigned)__popc(grp));
    }
}

// exclusive scan, three kernels: tiles of 4096 counts (256 threads x 16), tile totals, add-back
constexpr int SCAN_T = 256, SCAN_I = 16, SCAN_TILE = SCAN_T * SCAN_I;

__device__ __forceinline__ unsigned block_exclusive(unsigned v, unsigned* total) {
    __shared__ unsigned ws[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    unsigned inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) ws[w] = inc;
    __syncthreads();
    if (w == 0) {
        unsigned s = lane < nw ? ws[lane] : 0, si = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned t = __shfl_up_sync(0xffffffffu, si, o);
            if (lane >= o) si += t;
        }
        ws[lane] = si - s;                 // exclusive prefix of the warp totals
        if (lane == 31 && total) *total = si;
    }
    __syncthreads();
    const unsigned r = ws[w] + inc - v;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(SCAN_T)
k_scan_tiles(long long n, const unsigned* __restrict__ in, unsigned* __restrict__ out, unsigned* __restrict__ tile_sum) {
    const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_I;
    unsigned v[SCAN_I], s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_I; i++) {
        v[i] = (base + i < n) ? in[base + i] : 0u;
        s += v[i];
    }
    __shared__ unsigned tot;
    unsigned pre = block_exclusive(s, &tot);
#pragma unroll
    for (int i = 0; i < SCAN_I; i++) {
        if (base + i < n) out[base + i] = pre;
        pre += v[i];
    }
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(1024) k_scan_tops(int ntiles, unsigned* __restrict__ tile_sum) {
    __shared__ unsigned carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int b = 0; b < ntiles; b += 1024) {
        const int t = b + threadIdx.x;
        const unsigned v = t < ntiles ? tile_sum[t] : 0u;
        __shared__ unsigned tot;
        const unsigned pre = block_exclusive(v, &tot);
        const unsigned carry = carry_s;
        if (t < ntiles) tile_sum[t] = carry + pre;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + tot;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
k_scan_add(long long n, unsigned* __restrict__ out, const unsigned* __restrict__ tile_sum, unsigned* __restrict__ cursor,
           int* __restrict__ start_out, unsigned total) {
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t <= n; t += (long long)gridDim.x * blockDim.x) {
        if (t == n) {
            if (start_out) start_out[n] = (int)total;
            continue;
        }
        const unsigned v = out[t] + tile_sum[t / SCAN_TILE];
        out[t] = v;
        cursor[t] = v;
        if (start_out) start_out[t] = (int)v;
    }
}

__global__ void __launch_bounds__(256)
k_scan_finish(long long n, unsigned* __restrict__ out, const unsigned* __restrict__ tile_sum) {
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t <= n; t += (long long)gridDim.x * blockDim.x)
        out[t] = t == n ? tile_sum[(n + SCAN_TILE - 1) / SCAN_TILE] : out[t] + tile_sum[t / SCAN_TILE];
}

// slot = cursor[key]++ (warp-aggregated, ranks inside a warp follow the lane order)
__global__ void __launch_bounds__(256)
k_sort_slots(long long M, unsigned* __restrict__ key_dest, unsigned* __restrict__ cursor) {
    const int lane = threadIdx.x & 31;
    for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < M; m += (long long)gridDim.x * blockDim.x) {
        const unsigned k = key_dest[m];
        const unsigned act = __activemask();
        const unsigned grp = __match_any_sync(act, k);
        const int leader = __ffs(grp) - 1;
        unsigned base = 0;
        if (lane == leader) base = atomicAdd(cursor + k, (unsigned)__popc(grp));
        base = __shfl_sync(act, base, leader);
        key_dest[m] = base + (unsigned)__popc(grp & ((1u << lane) - 1u));
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
k_permute(long long M, const unsigned* __restrict__ dest, const T* __restrict__ in, T* __restrict__ out) {
    for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < M; m += (long long)gridDim.x * blockDim.x)
        out[dest[m]] = in[m];
}

}  // namespace

// out[t] = sum of in[0..t), t = 0..n (n + 1 entries); `tiles`: ceil(n / 4096) + 1 words of scratch
int plb_exclusive_scan_u32(plb_ctx* ctx, long long n, const unsigned* in, unsigned* out, unsigned* tiles) {
    const int ntiles = (int)((n + SCAN_TILE - 1) / SCAN_TILE);
    if (ntiles == 0) {
        PLB_CUDA(ctx, cudaMemsetAsync(out, 0, sizeof(unsigned), ctx->stream));
        return 0;
    }
    k_scan_tiles<<<ntiles, SCAN_T, 0, ctx->stream>>>(n, in, out, tiles);
    PLB_LAUNCHED(ctx);
    k_scan_tops<<<1, 1024, 0, ctx->stream>>>(ntiles + 1, tiles);       // tiles[ntiles] (zeroed by the caller) = grand total
    PLB_LAUNCHED(ctx);
    k_scan_finish<<<plb_grid_for(ctx, n + 1, 256, 8), 256, 0, ctx->stream>>>(n, out, tiles);
    PLB_LAUNCHED(ctx);
    return 0;
}

extern "C" {

int plb_sort_plan(plb_ctx* ctx, long long M, const double* d_tr_x, int nz, int nxx, double Lz, double Lx,
                  unsigned* d_dest, int* d_cell_start) {
    if (!ctx || !d_dest) return 1;
    if (M >= (1LL << 31)) PLB_FAIL(ctx, "plb_sort_plan: %lld markers on one GPU exceed the 32-bit slot table", M);
    if (nz < 2 || nxx < 2) PLB_FAIL(ctx, "plb_sort_plan: grid too small");
    PLB_CUDA(ctx, cudaSetDevice(ctx->device));
    const long long ncell = (long long)(nz - 1) * (nxx - 1);
    const int ntiles = (int)((ncell + SCAN_TILE - 1) / SCAN_TILE);
    // scratch: histogram | first slots | cursors | tile totals
    if (plb_ws_reserve(ctx, sizeof(unsigned) * (size_t)(3 * ncell + ntiles + 64))) return 2;
    unsigned* count = (unsigned*)ctx->ws;
    unsigned* start = count + ncell;
    unsigned* cursor = start + ncell;
    unsigned* tiles = cursor + ncell;
    PLB_CUDA(ctx, cudaMemsetAsync(count, 0, sizeof(unsigned) * (size_t)ncell, ctx->stream));
    if (M > 0) {
        k_sort_keys<<<plb_grid_for(ctx, M, 256, 8), 256, 0, ctx->stream>>>(M, (const double2*)d_tr_x, nz, nxx, Lz, Lx,
                                                                         d_dest, count);
        PLB_LAUNCHED(ctx);
    }
    k_scan_tiles<<<ntiles, SCAN_T, 0, ctx->stream>>>(ncell, count, start, tiles);
    PLB_LAUNCHED(ctx);
    k_scan_tops<<<1, 1024, 0, ctx->stream>>>(ntiles, tiles);
    PLB_LAUNCHED(ctx);
    k_scan_add<<<plb_grid_for(ctx, ncell + 1, 256, 8), 256, 0, ctx->stream>>>(ncell, start, tiles, cursor, d_cell_start,
                                                                           (unsigned)M);
    PLB_LAUNCHED(ctx);
    if (M > 0) {
        k_sort_slots<<<plb_grid_for(ctx, M, 256, 8), 256, 0, ctx->stream>>>(M, d_dest, cursor);
        PLB_LAUNCHED(ctx);
    }
    return 0;
}

int plb_permute(plb_ctx* ctx, long long M, const unsigned* d_dest, const double* d_in, double* d_out, int width) {
    if (!ctx || !d_dest || !d_in || !d_out) return 1;
    if (d_in == d_out) PLB_FAIL(ctx, "plb_permute: in-place permutation is not supported");
    if (width != 1 && width != 2) PLB_FAIL(ctx, "plb_permute: width must be 1 or 2 doubles per marker");
    PLB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (M <= 0) return 0;
    plb_prof_scope prof_(ctx, PLB_K_SORT, (8.0 * width * 2 + 4.0) * (double)M);
    const int grid = plb_grid_for(ctx, M, 256, 8);
    if (width == 1) k_permute<double><<<grid, 256, 0, ctx->stream>>>(M, d_dest, d_in, d_out);
    else k_permute<double2><<<grid, 256, 0, ctx->stream>>>(M, d_dest, (const double2*)d_in, (double2*)d_out);
    PLB_LAUNCHED(ctx);
    return 0;
}

}  // extern "C"
