// blas1.cu -- see blas1.cuh
#include "blas1.cuh"
#include "comm.cuh"

namespace {

constexpr int TB = 256;

template <int J>
__device__ __forceinline__ void block_reduce_store(double (&acc)[J], double* partials, int nblocks,
                                                   unsigned int* ticket, double* out) {
    __shared__ double sm[J][TB / 32];
    __shared__ bool last;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < J; j++) {
        double v = warp_sum(acc[j]);
        if (lane == 0) sm[j][w] = v;
    }
    __syncthreads();
    if (threadIdx.x < J) {
        double s = 0;
#pragma unroll
        for (int q = 0; q < TB / 32; q++) s += sm[threadIdx.x][q];
        partials[(size_t)threadIdx.x * nblocks + (blockIdx.y * gridDim.x + blockIdx.x)] = s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = atomicAdd(ticket, 1u);
        last = (t == (unsigned)nblocks - 1);
    }
    __syncthreads();
    if (last) {
        __threadfence();
        // fixed-order final sum: warp j handles output j
        for (int j = w; j < J; j += TB / 32) {
            double s = 0;
            for (int b = lane; b < nblocks; b += 32) s += partials[(size_t)j * nblocks + b];
            s = warp_sum(s);
            if (lane == 0) out[j] = s;
        }
        if (threadIdx.x == 0) *ticket = 0;
    }
}

struct VecList {
    const double* v[PLB_DOT_CHUNK];
};

// blockIdx.y selects the plane; each plane contributes its segment [base, base + n)
template <int J>
__global__ void __launch_bounds__(TB)
k_multi_dot(long long n, long long pstride, long long seg_off, VecList V, const double* __restrict__ w,
            double* partials, unsigned int* ticket, double* out) {
    double acc[J];
#pragma unroll
    for (int j = 0; j < J; j++) acc[j] = 0;
    const long long base = blockIdx.y * pstride + seg_off;
    for (long long t = blockIdx.x * (long long)TB + threadIdx.x; t < n; t += (long long)gridDim.x * TB) {
        const long long i = base + t;
        double wi = w[i];
#pragma unroll
        for (int j = 0; j < J; j++) acc[j] += V.v[j][i] * wi;
    }
    block_reduce_store<J>(acc, partials, gridDim.x * gridDim.y, ticket, out);
}

struct VecList2 {
    const double* v[PLB_DOT_CHUNK];
    const double* u[PLB_DOT_CHUNK];
};

template <int J, bool TWO>
__global__ void __launch_bounds__(TB)
k_multi_axpy2(long long n, const double* __restrict__ h, VecList2 L, double* __restrict__ w,
              double* __restrict__ u, double post_scale) {
    double c[J];
#pragma unroll
    for (int j = 0; j < J; j++) c[j] = h[j];
    for (long long i = blockIdx.x * (long long)TB + threadIdx.x; i < n; i += (long long)gridDim.x * TB) {
        double a = w[i];
#pragma unroll
        for (int j = 0; j < J; j++) a -= c[j] * L.v[j][i];
        w[i] = a * post_scale;
        if (TWO) {
            double b = u[i];
#pragma unroll
            for (int j = 0; j < J; j++) b -= c[j] * L.u[j][i];
            u[i] = b * post_scale;
        }
    }
}

__global__ void __launch_bounds__(TB)
k_axpy_dev(long long n, const double* __restrict__ a, double sign, const double* __restrict__ x,
           double* __restrict__ y) {
    const double s = sign * (*a);
    for (long long i = blockIdx.x * (long long)TB + threadIdx.x; i < n; i += (long long)gridDim.x * TB)
        y[i] += s * x[i];
}

__global__ void __launch_bounds__(TB)
k_scale_rsqrt2(long long n, const double* __restrict__ n2, double* __restrict__ a, double* __restrict__ b) {
    const double s = 1.0 / sqrt(*n2);
    for (long long i = blockIdx.x * (long long)TB + threadIdx.x; i < n; i += (long long)gridDim.x * TB) {
        a[i] *= s;
        if (b) b[i] *= s;
    }
}

__global__ void __launch_bounds__(TB)
k_gcr_update(long long n, const double* __restrict__ pa, const double* __restrict__ z,
             const double* __restrict__ c, double* __restrict__ x, double* __restrict__ r) {
    const double a = *pa;
    for (long long i = blockIdx.x * (long long)TB + threadIdx.x; i < n; i += (long long)gridDim.x * TB) {
        x[i] += a * z[i];
        r[i] -= a * c[i];
    }
}

int vec_grid(const plb_ctx* ctx, long long n) { return plb_grid_for(ctx, n, TB, 6); }

}  // namespace

int plb_reduce_ws_init(plb_ctx* ctx, plb_reduce_ws* ws) {
    ws->max_blocks = ctx->num_sms * 6;
    ws->nplanes = 0, ws->pstride = 0, ws->seg_off = 0, ws->seg_len = 0, ws->allreduce = false;
    PLB_CUDA(ctx, cudaMalloc(&ws->partials, sizeof(double) * PLB_DOT_CHUNK * ws->max_blocks));
    PLB_CUDA(ctx, cudaMalloc(&ws->ticket, sizeof(unsigned int)));
    PLB_CUDA(ctx, cudaMemsetAsync(ws->ticket, 0, sizeof(unsigned int), ctx->stream));
    return 0;
}

void plb_reduce_ws_free(plb_reduce_ws* ws) {
    if (ws->partials) cudaFree(ws->partials);
    if (ws->ticket) cudaFree(ws->ticket);
    ws->partials = nullptr;
    ws->ticket = nullptr;
}

int plb_multi_dot(plb_ctx* ctx, plb_reduce_ws* ws, long long n, int k, const double* const* h_V,
                  const double* w, double* d_out) {
    const int np = ws->nplanes > 0 ? ws->nplanes : 1;
    const long long seg = ws->nplanes > 0 ? ws->seg_len : n;
    const long long pstride = ws->nplanes > 0 ? ws->pstride : 0, seg_off = ws->nplanes > 0 ? ws->seg_off : 0;
    int gx = vec_grid(ctx, seg * np) / np;
    if (gx < 1) gx = 1;
    if (gx * np > ws->max_blocks) gx = ws->max_blocks / np;
    const dim3 grid(gx, np);
    n = seg;
    // k vectors read once, w once per chunk of 8
    plb_prof_scope prof_(ctx, PLB_K_MDOT, 8.0 * (double)seg * np * (k + (k + PLB_DOT_CHUNK - 1) / PLB_DOT_CHUNK));
    for (int j0 = 0; j0 < k; j0 += PLB_DOT_CHUNK) {
        int J = k - j0 < PLB_DOT_CHUNK ? k - j0 : PLB_DOT_CHUNK;
        VecList V;
        for (int j = 0; j < PLB_DOT_CHUNK; j++) V.v[j] = h_V[j0 + (j < J ? j : 0)];
#define MD(JJ) case JJ: k_multi_dot<JJ><<<grid, TB, 0, ctx->stream>>>(n, pstride, seg_off, V, w, ws->partials, ws->ticket, d_out + j0); break;
        switch (J) { MD(1) MD(2) MD(3) MD(4) MD(5) MD(6) MD(7) MD(8) }
#undef MD
        PLB_LAUNCHED(ctx);
    }
    if (ws->allreduce && plb_comm_allreduce(ctx, d_out, (size_t)k, PLB_OP_SUM)) return 2;
    return 0;
}

int plb_dot(plb_ctx* ctx, plb_reduce_ws* ws, long long n, const double* a, const double* b, double* d_out) {
    const double* v[1] = {a};
    return plb_multi_dot(ctx, ws, n, 1, v, b, d_out);
}

int plb_multi_axpy2(plb_ctx* ctx, long long n, int k, const double* d_h, const double* const* h_V,
                    double* w, const double* const* h_U, double* u, double post_scale) {
    int grid = vec_grid(ctx, n);
    // k (2k) vectors read once, w (and u) read+written once per chunk of 8
    plb_prof_scope prof_(ctx, PLB_K_MAXPY, 8.0 * (double)n * (h_U ? 2 : 1) * (k + 2 * ((k + PLB_DOT_CHUNK - 1) / PLB_DOT_CHUNK)));
    for (int j0 = 0; j0 < k; j0 += PLB_DOT_CHUNK) {
        int J = k - j0 < PLB_DOT_CHUNK ? k - j0 : PLB_DOT_CHUNK;
        VecList2 L;
        const double ps = (j0 + J >= k) ? post_scale : 1.0;      // scale once, in the last chunk
        for (int j = 0; j < PLB_DOT_CHUNK; j++) {
            L.v[j] = h_V[j0 + (j < J ? j : 0)];
            L.u[j] = h_U ? h_U[j0 + (j < J ? j : 0)] : nullptr;
        }
#define MA(JJ)                                                                                   \
    case JJ:                                                                                     \
        if (h_U) k_multi_axpy2<JJ, true><<<grid, TB, 0, ctx->stream>>>(n, d_h + j0, L, w, u, ps);     \
        else k_multi_axpy2<JJ, false><<<grid, TB, 0, ctx->stream>>>(n, d_h + j0, L, w, u, ps);        \
        break;
        switch (J) { MA(1) MA(2) MA(3) MA(4) MA(5) MA(6) MA(7) MA(8) }
#undef MA
        PLB_LAUNCHED(ctx);
    }
    return 0;
}

int plb_axpy_dev(plb_ctx* ctx, long long n, const double* d_a, double sign, const double* x, double* y) {
    k_axpy_dev<<<vec_grid(ctx, n), TB, 0, ctx->stream>>>(n, d_a, sign, x, y);
    PLB_LAUNCHED(ctx);
    return 0;
}

int plb_scale_rsqrt2(plb_ctx* ctx, long long n, const double* d_n2, double* a, double* b) {
    k_scale_rsqrt2<<<vec_grid(ctx, n), TB, 0, ctx->stream>>>(n, d_n2, a, b);
    PLB_LAUNCHED(ctx);
    return 0;
}

int plb_copy(plb_ctx* ctx, long long n, const double* x, double* y) {
    PLB_CUDA(ctx, cudaMemcpyAsync(y, x, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
    return 0;
}

int plb_gcr_update(plb_ctx* ctx, long long n, const double* d_a, const double* z, const double* c,
                   double* x, double* r) {
    k_gcr_update<<<vec_grid(ctx, n), TB, 0, ctx->stream>>>(n, d_a, z, c, x, r);
    PLB_LAUNCHED(ctx);
    return 0;
}

int plb_read_scalars(plb_ctx* ctx, const double* d, int k, double* h) {
    if (k > 64) PLB_FAIL(ctx, "plb_read_scalars: k > 64");
    PLB_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned, d, sizeof(double) * k, cudaMemcpyDeviceToHost, ctx->stream));
    PLB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < k; i++) h[i] = ctx->h_pinned[i];
    return 0;
}
