// blas1.cuh -- vector kernels of the Krylov solvers (fp64, device-resident scalars).
//
// Every reduction is deterministic: each CTA writes one partial, the last CTA to finish sums the
// partials in a fixed order (threadfence + ticket counter), so no memset and no float atomics.
// Scalars (dot products, norms, Krylov coefficients) stay on the device; vector updates read their
// coefficient through a pointer, so a Krylov step needs no host round trip except the
// convergence test.
#pragma once
#include "common.cuh"

#define PLB_DOT_CHUNK 8

struct plb_reduce_ws {
    double* partials;          // [PLB_DOT_CHUNK * max_blocks]
    unsigned int* ticket;      // zero-initialised, self-resetting
    int max_blocks;
    // slab-distributed vectors: `nplanes` planes of `pstride` doubles whose owned rows are the
    // segment [seg_off, seg_off + seg_len) of each plane (halo rows are skipped by the reductions);
    // nplanes == 0: plain contiguous vectors of n doubles.  `allreduce`: sum the results over the
    // ranks of the context's communicator before they are used.
    int nplanes;
    long long pstride, seg_off, seg_len;
    bool allreduce;
};

inline void plb_reduce_shape(plb_reduce_ws* ws, int nplanes, long long pstride, long long seg_off,
                             long long seg_len, bool allreduce) {
    ws->nplanes = nplanes, ws->pstride = pstride, ws->seg_off = seg_off, ws->seg_len = seg_len;
    ws->allreduce = allreduce;
}

int plb_reduce_ws_init(plb_ctx* ctx, plb_reduce_ws* ws);
void plb_reduce_ws_free(plb_reduce_ws* ws);

// out[0] = sum a[i]*b[i]
int plb_dot(plb_ctx* ctx, plb_reduce_ws* ws, long long n, const double* a, const double* b, double* d_out);
// d_out[j] = sum V[j][i]*w[i], j < k (any k; processed in chunks of PLB_DOT_CHUNK)
int plb_multi_dot(plb_ctx* ctx, plb_reduce_ws* ws, long long n, int k, const double* const* h_V,
                  const double* w, double* d_out);
// w = post_scale * (w - sum_j h[j] V[j])   and (if U) the same for u with U     (h on the device)
int plb_multi_axpy2(plb_ctx* ctx, long long n, int k, const double* d_h, const double* const* h_V,
                    double* w, const double* const* h_U, double* u, double post_scale = 1.0);
// y += sign * (*d_a) * x
int plb_axpy_dev(plb_ctx* ctx, long long n, const double* d_a, double sign, const double* x, double* y);
// a *= 1/sqrt(*d_n2), b *= 1/sqrt(*d_n2) (b may be NULL)
int plb_scale_rsqrt2(plb_ctx* ctx, long long n, const double* d_n2, double* a, double* b);
// y = x
int plb_copy(plb_ctx* ctx, long long n, const double* x, double* y);
// fused GCR update: x += a*z, r -= a*c with a = *d_a
int plb_gcr_update(plb_ctx* ctx, long long n, const double* d_a, const double* z, const double* c,
                   double* x, double* r);
// read k doubles back (synchronises the stream)
int plb_read_scalars(plb_ctx* ctx, const double* d, int k, double* h);
