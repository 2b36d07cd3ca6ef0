// common.cuh -- shared host/device helpers for libpylamp_b200 (sm_100a, fp64).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/pylamp_b200.h"

struct plb_comm;
struct plb_migrate_ws;
struct plb_inject_ws;

struct plb_ctx {
    int device;
    plb_comm* comm;        // z-slab communicator (NULL: single GPU)
    cudaStream_t stream;
    bool own_stream;
    char err[1024];
    long long launches;
    // growable scratch (accumulation planes, reduction partials)
    void* ws;
    size_t ws_bytes;
    double* h_pinned;      // small pinned staging buffer for scalar read-backs
    int num_sms;
    // optional per-kernel-class timing with CUDA event pairs on the launching stream
    bool prof_on;
    cudaEvent_t* prof_ev;  // 2 * prof_cap events
    int* prof_cls;
    int prof_cap, prof_used;
    long long prof_skipped[16];
    double* prof_bytes;    // algorithmic bytes per recorded pair
    // slab-local fields (several ranks): this rank keeps node rows [slab_i0, slab_i1) of every grid field current,
    // plus slab_halo rows of each z-neighbour; its markers lie in the cell rows [slab_i0, min(slab_i1, nz-1))
    bool slab_on;
    int slab_i0, slab_i1, slab_halo;
    double* slab_scratch;  // receive buffer of the boundary-row accumulate (grown on demand)
    plb_migrate_ws* mig;   // marker migration scratch (migrate.cu)
    plb_inject_ws* inj;    // marker injection scratch (inject.cu)
    size_t slab_scratch_dbl;
    // tuning knobs (plb_ctx_set_param)
    int t2g_variant;       // 1: wide-load chunk kernel for weighted schemes (default), 0: generic kernel only
    int t2g_parts;         // fused step kernel: lanes per run (0 = 1; 1, 2, 4)
    int t2g_nm;            // fused step kernel: markers per CTA (960 default, 1024)
    int t2g_nfmax;         // fused step kernel: fields per node-target work item (default 6)
    int t2g_minb;          // fused step kernel: 2 = at most two resident CTAs per SM (default: three when they fit)
};

// kernel classes for plb_profile_* (bench.py's roofline block)
enum {
    PLB_K_CHEB0 = 0,     // Chebyshev-Jacobi smoother sweep, finest level
    PLB_K_STOKES_OP = 1, // coupled Stokes residual / apply, finest level
    PLB_K_MDOT = 2,      // fused multi-dot (Gram-Schmidt)
    PLB_K_MAXPY = 3,     // fused multi-axpy
    PLB_K_T2G = 4,       // marker -> grid scatter
    PLB_K_RK4 = 5,       // RK4 advection
    PLB_K_G2T = 6,       // grid -> marker interpolation
    PLB_K_MGCOARSE = 7,  // everything below the finest MG level (whole coarse part of a V-cycle)
    PLB_K_MGXFER0 = 8,   // finest-level residual + restriction + prolongation
    PLB_K_PRECRHS = 9,   // preconditioner right-hand side
    PLB_K_DIFF = 10,     // heat operator
    PLB_K_MARKER_MISC = 11,
    PLB_K_SORT = 12,     // marker-by-cell sort: permutation of the marker arrays
    PLB_K_NCLASS = 16
};

void plb_prof_begin(plb_ctx* ctx, int cls, double bytes);
void plb_prof_end(plb_ctx* ctx);
struct plb_prof_scope {   // cls < 0: no-op (scopes must not nest)
    plb_ctx* c;
    bool live;
    plb_prof_scope(plb_ctx* ctx, int cls, double bytes = 0) : c(ctx), live(ctx->prof_on && cls >= 0) {
        if (live) plb_prof_begin(c, cls, bytes);
    }
    ~plb_prof_scope() { if (live) plb_prof_end(c); }
};

#define PLB_FAIL(ctx, ...)                                     \
    do {                                                       \
        snprintf((ctx)->err, sizeof((ctx)->err), __VA_ARGS__); \
        return 1;                                              \
    } while (0)

#define PLB_CUDA(ctx, call)                                                              \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) {                                                         \
            snprintf((ctx)->err, sizeof((ctx)->err), "%s:%d: %s: %s", __FILE__, __LINE__, \
                     #call, cudaGetErrorString(e_));                                     \
            return 2;                                                                    \
        }                                                                                \
    } while (0)

// count + check a kernel launch
#define PLB_LAUNCHED(ctx)                  \
    do {                                   \
        (ctx)->launches++;                 \
        PLB_CUDA(ctx, cudaGetLastError()); \
    } while (0)

int plb_ws_reserve(plb_ctx* ctx, size_t bytes);

static inline int plb_blocks(long long n, int threads) { return (int)((n + threads - 1) / threads); }

// persistent-style grid: enough CTAs to fill the machine a few times over, capped by the work
static inline int plb_grid_for(const plb_ctx* ctx, long long n, int threads, int ctas_per_sm = 8) {
    long long need = (n + threads - 1) / threads;
    long long cap = (long long)ctx->num_sms * ctas_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_down_sync(0xffffffffu, v, o));
    return v;
}

// atomic max / min for doubles (CAS loop; used once per block)
__device__ __forceinline__ void atomic_max_double(double* addr, double v) {
    unsigned long long* a = (unsigned long long*)addr;
    unsigned long long old = *a, assumed;
    do {
        assumed = old;
        if (__longlong_as_double(assumed) >= v) break;
        old = atomicCAS(a, assumed, __double_as_longlong(v));
    } while (assumed != old);
}
__device__ __forceinline__ void atomic_min_double(double* addr, double v) {
    unsigned long long* a = (unsigned long long*)addr;
    unsigned long long old = *a, assumed;
    do {
        assumed = old;
        if (__longlong_as_double(assumed) <= v) break;
        old = atomicCAS(a, assumed, __double_as_longlong(v));
    } while (assumed != old);
}
