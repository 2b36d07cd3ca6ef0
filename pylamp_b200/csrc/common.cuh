// common.cuh -- shared host/device helpers for libpylamp_b200 (sm_100a, fp64).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/pylamp_b200.h"

struct plb_ctx {
    int device;
    cudaStream_t stream;
    bool own_stream;
    char err[1024];
    long long launches;
    // growable scratch (accumulation planes, reduction partials)
    void* ws;
    size_t ws_bytes;
    double* h_pinned;      // small pinned staging buffer for scalar read-backs
    int num_sms;
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
