// gridops.cu -- driver-inline grid steps of the reference loop body on the device:
// cell-centre velocities + BC ghost ring (pylamp2.py:491-545), signed field maximum and the
// diffusivity maximum of the dt selection (pylamp2.py:339-343, :364-366), x2vp de-interleave
// (pylamp_stokes.py:86-101).
#include "common.cuh"

namespace {

// interior of the (nz+1) x (nxx+1) cell-centre fields, zero elsewhere (pylamp2.py:491-499)
__global__ void __launch_bounds__(256)
k_centre_interior(int nz, int nxx, int ld, const double* __restrict__ vz,
                  const double* __restrict__ vx, int ldc, double* __restrict__ vzc,
                  double* __restrict__ vxc) {
    long long n = (long long)(nz + 1) * (nxx + 1);
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n;
         t += (long long)gridDim.x * blockDim.x) {
        int I = (int)(t / (nxx + 1)), J = (int)(t % (nxx + 1));
        double a = 0, b = 0;
        if (I >= 1 && I <= nz - 1 && J >= 1 && J <= nxx - 1) {
            a = 0.5 * (vz[(long long)I * ld + J - 1] + vz[(long long)(I - 1) * ld + J - 1]);
            b = 0.5 * (vx[(long long)(I - 1) * ld + J] + vx[(long long)(I - 1) * ld + J - 1]);
        }
        vzc[(long long)I * ldc + J] = a;
        vxc[(long long)I * ldc + J] = b;
    }
}

// One ghost line of the ring.  axis 0: row `gh` <- row `src`; axis 1: column `gh` <- column `src`.
// mode 1 (FREESLIP): tangential copied, normal negated; mode 2 (CYCLIC): both copied;
// flow (FLOWTHRU, x-walls only): vx copied.  pylamp2.py:503-545.
__global__ void k_ring_line(int axis, int gh, int src, int n, int ldc, int mode, int flow,
                            double* vzc, double* vxc) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        long long g = axis == 0 ? (long long)gh * ldc + t : (long long)t * ldc + gh;
        long long s = axis == 0 ? (long long)src * ldc + t : (long long)t * ldc + src;
        double* tang = axis == 0 ? vxc : vzc;
        double* norm = axis == 0 ? vzc : vxc;
        if (mode == 1) {
            tang[g] = tang[s];
            norm[g] = -norm[s];
        } else if (mode == 2) {
            tang[g] = tang[s];
            norm[g] = norm[s];
        }
        if (flow) vxc[g] = vxc[s];
    }
}

__global__ void __launch_bounds__(256)
k_field_max(int nz, int nxx, int ld, const double* __restrict__ f, double* out) {
    long long n = (long long)nz * nxx;
    double m = -INFINITY;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n;
         t += (long long)gridDim.x * blockDim.x) {
        int i = (int)(t / nxx), j = (int)(t % nxx);
        m = fmax(m, f[(long long)i * ld + j]);
    }
    m = warp_max(m);
    __shared__ double s[8];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; i++) m = fmax(m, s[i]);
        atomic_max_double(out, m);
    }
}

__global__ void __launch_bounds__(256)
k_max_diffusivity2(int nz, int nxx, int ld, const double* __restrict__ kz,
                   const double* __restrict__ rho, const double* __restrict__ cp, double* out) {
    long long n = (long long)nz * nxx;
    double m = -INFINITY;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n;
         t += (long long)gridDim.x * blockDim.x) {
        int i = (int)(t / nxx), j = (int)(t % nxx);
        long long o = (long long)i * ld + j;
        m = fmax(m, 2 * (kz[o] / (rho[o] * cp[o])));            // pylamp2.py:340
    }
    m = warp_max(m);
    __shared__ double s[8];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; i++) m = fmax(m, s[i]);
        atomic_max_double(out, m);
    }
}

__global__ void k_set(double* p, double v) { *p = v; }

__global__ void __launch_bounds__(256)
k_x2vp(int nz, int nxx, int ld, const double* __restrict__ x, double* __restrict__ vz,
       double* __restrict__ vx, double* __restrict__ p) {
    long long n = (long long)nz * nxx;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n;
         t += (long long)gridDim.x * blockDim.x) {
        int i = (int)(t / nxx), j = (int)(t % nxx);
        long long o = (long long)i * ld + j;
        vz[o] = x[3 * t];
        vx[o] = x[3 * t + 1];
        p[o] = x[3 * t + 2];
    }
}

int read_back_scalar(plb_ctx* ctx, double* d, double* h_out) {
    PLB_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned, d, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    PLB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *h_out = ctx->h_pinned[0];
    return 0;
}

}  // namespace

extern "C" {

int plb_centre_velocities(plb_ctx* ctx, int nz, int nxx, int ld, const double* d_vz,
                          const double* d_vx, const int* h_bc, int ldc, double* d_vz_c,
                          double* d_vx_c) {
    if (!ctx) return 1;
    if (ldc < nxx + 1) PLB_FAIL(ctx, "plb_centre_velocities: ldc < nxx+1");
    PLB_CUDA(ctx, cudaSetDevice(ctx->device));
    k_centre_interior<<<plb_grid_for(ctx, (long long)(nz + 1) * (nxx + 1), 256, 8), 256, 0, ctx->stream>>>(
        nz, nxx, ld, d_vz, d_vx, ldc, d_vz_c, d_vx_c);
    PLB_LAUNCHED(ctx);
    // ring order of the reference: z=0, x=0, z=L, x=L (later lines see earlier ones)
    const int wall_axis[4] = {0, 1, 0, 1};
    const int wall_side[4] = {0, 0, 1, 1};
    for (int w = 0; w < 4; w++) {
        int axis = wall_axis[w], side = wall_side[w];
        int b = h_bc[2 * side + axis];
        int nlines = axis == 0 ? nz + 1 : nxx + 1;      // ghost index range along the normal
        int len = axis == 0 ? nxx + 1 : nz + 1;
        int gh = side == 0 ? 0 : nlines - 1;
        int mode = 0, src = side == 0 ? 1 : nlines - 2;
        if (b & PLB_BC_FREESLIP) mode = 1;
        else if (b & PLB_BC_CYCLIC) mode = 2, src = side == 0 ? nlines - 2 : 1;
        int flow = (axis == 1 && (b & PLB_BC_FLOWTHRU)) ? 1 : 0;
        if (flow && mode == 2) PLB_FAIL(ctx, "plb_centre_velocities: CYCLIC|FLOWTHRU wall");
        if (!mode && !flow) continue;
        k_ring_line<<<plb_blocks(len, 256), 256, 0, ctx->stream>>>(axis, gh, src, len, ldc, mode, flow,
                                                                 d_vz_c, d_vx_c);
        PLB_LAUNCHED(ctx);
    }
    return 0;
}

int plb_field_max(plb_ctx* ctx, int nz, int nxx, int ld, const double* d_f, double* h_out) {
    if (!ctx) return 1;
    PLB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (plb_ws_reserve(ctx, 64)) return 2;
    double* d = (double*)ctx->ws;
    k_set<<<1, 1, 0, ctx->stream>>>(d, -INFINITY);
    PLB_LAUNCHED(ctx);
    k_field_max<<<plb_grid_for(ctx, (long long)nz * nxx, 256, 8), 256, 0, ctx->stream>>>(nz, nxx, ld, d_f, d);
    PLB_LAUNCHED(ctx);
    return read_back_scalar(ctx, d, h_out);
}

int plb_max_diffusivity2(plb_ctx* ctx, int nz, int nxx, int ld, const double* d_kz,
                         const double* d_rho, const double* d_cp, double* h_out) {
    if (!ctx) return 1;
    PLB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (plb_ws_reserve(ctx, 64)) return 2;
    double* d = (double*)ctx->ws;
    k_set<<<1, 1, 0, ctx->stream>>>(d, -INFINITY);
    PLB_LAUNCHED(ctx);
    k_max_diffusivity2<<<plb_grid_for(ctx, (long long)nz * nxx, 256, 8), 256, 0, ctx->stream>>>(
        nz, nxx, ld, d_kz, d_rho, d_cp, d);
    PLB_LAUNCHED(ctx);
    return read_back_scalar(ctx, d, h_out);
}

int plb_x2vp(plb_ctx* ctx, int nz, int nxx, int ld, const double* d_x, double* d_vz,
             double* d_vx, double* d_p) {
    if (!ctx) return 1;
    PLB_CUDA(ctx, cudaSetDevice(ctx->device));
    k_x2vp<<<plb_grid_for(ctx, (long long)nz * nxx, 256, 8), 256, 0, ctx->stream>>>(nz, nxx, ld, d_x, d_vz,
                                                                                d_vx, d_p);
    PLB_LAUNCHED(ctx);
    return 0;
}

}  // extern "C"
