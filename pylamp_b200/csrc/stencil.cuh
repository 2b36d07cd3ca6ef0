// stencil.cuh -- matrix-free rows of the reference's staggered-grid Stokes system.
//
// Row formulas restate pylamp_stokes.py:376-429 (z-momentum), :435-490 (x-momentum), :496-518 and
// :333-354 (continuity) for a rectilinear grid; SURVEY.md Appendix C lists every row class.
// Planar storage: vz, vx, p are separate (nz x ld) planes, index [i (z)][j (x)].
//   vz(i,j) at (z_i, x_{j+1/2}),  vx(i,j) at (z_{i+1/2}, x_j),  p(i,j) at the centre of cell (i,j),
//   etas at nodes, etan at cell centres.
#pragma once
#include "common.cuh"

struct LevelDev {
    int nz, nxx, ld;
    int i0, i1;           // node rows [i0, i1) handled by this rank (z-slab); single GPU: [0, nz)
    const double* idz;    // [nz]  1/(gz[i+1]-gz[i]),   i <= nz-2   (0 beyond)
    const double* idzc;   // [nz]  1/(gz[i+1]-gz[i-1]), 1 <= i <= nz-2 (0 elsewhere)
    const double* idx;    // [nxx] 1/(gx[j+1]-gx[j])
    const double* idxc;   // [nxx] 1/(gx[j+1]-gx[j-1])
    const double* etas;
    const double* etan;
    // closure: 0 = the reference's (tangential velocity slaved at the first interior staggered
    // point, pylamp_stokes.py:170-175, :249-255); 1 = standard free slip (zero shear stress on the
    // wall nodes; every in-domain staggered velocity is an unknown) used on coarse MG levels
    int proper;
    int vz_i0, vz_i1, vz_j0, vz_j1;   // inclusive ranges of the momentum rows
    int vx_i0, vx_i1, vx_j0, vx_j1;
    double sl_z0, sl_z1;              // vx(0,j) = sl_z0*vx(1,j), vx(nz-2,j) = sl_z1*vx(nz-3,j)
    int ns_z0, ns_z1;                 // proper levels: no-slip (instead of free-slip) z-walls
    // reference closure only: flow-through wall at x = 0 (BC_TYPE_FLOWTHRU|FREESLIP, pylamp_stokes.py:268-273):
    // vx(i,0) = vx(i,1) instead of vx(i,0) = 0
    int ft_x0;
};

__device__ __forceinline__ bool is_vz_row(const LevelDev& L, int i, int j) {
    return i >= L.vz_i0 && i <= L.vz_i1 && j >= L.vz_j0 && j <= L.vz_j1;
}
__device__ __forceinline__ bool is_vx_row(const LevelDev& L, int i, int j) {
    return i >= L.vx_i0 && i <= L.vx_i1 && j >= L.vx_j0 && j <= L.vx_j1;
}
// continuity rows of the SOLVER: every real cell except the four corners.  The reference replaces
// the continuity row of cell (3,2) by the pressure anchor P(3,2) = 0 (pylamp_stokes.py:525-551);
// pinning one pressure leaves a near-null mode (eigenvalue ~ 1/N) that costs GMRES a long plateau
// (iteration counts halve without it: oracle/mg_prototype.py).  The solver therefore keeps that
// cell's continuity row -- it is a linear combination of the others for closed boundaries, the
// reference's solution satisfies it -- iterates on the consistent singular system
// (null space: constant pressure) and shifts the pressure to P(3,2) = 0 at the end.
__device__ __forceinline__ bool is_p_row(const LevelDev& L, int i, int j) {
    if (i > L.nz - 2 || j > L.nxx - 2) return false;
    bool corner = (i == 0 || i == L.nz - 2) && (j == 0 || j == L.nxx - 2);
    return !corner;
}

struct VzCoef {
    double cN, cS, cE, cW, xE, xW, diag;
};
struct VxCoef {
    double cE, cW, cS, cN, xS, xN, diag;
};

__device__ __forceinline__ VzCoef vz_coef(const LevelDev& L, int i, int j) {
    const long long o = (long long)i * L.ld + j;
    const double rz = L.idzc[i], rx = L.idx[j];
    VzCoef c;
    c.cN = 4 * L.etan[o] * L.idz[i] * rz;               // to vz(i+1,j)
    c.cS = 4 * L.etan[o - L.ld] * L.idz[i - 1] * rz;    // to vz(i-1,j)
    const double eE = L.etas[o + 1], eW = L.etas[o];
    c.cE = 2 * eE * L.idxc[j + 1] * rx;                 // to vz(i,j+1), shear at node (i,j+1)
    c.cW = 2 * eW * L.idxc[j] * rx;                     // to vz(i,j-1), shear at node (i,j)
    c.xE = 2 * eE * rz * rx;
    c.xW = 2 * eW * rz * rx;
    if (L.proper) {
        if (j + 1 == L.nxx - 1) c.cE = 0, c.xE = 0;     // node on the x=L wall: no shear stress
        if (j == 0) c.cW = 0, c.xW = 0;                 // node on the x=0 wall
    }
    c.diag = -(c.cN + c.cS + c.cE + c.cW);
    return c;
}

__device__ __forceinline__ VxCoef vx_coef(const LevelDev& L, int i, int j) {
    const long long o = (long long)i * L.ld + j;
    const double rx = L.idxc[j], rz = L.idz[i];
    VxCoef c;
    c.cE = 4 * L.etan[o] * L.idx[j] * rx;               // to vx(i,j+1)
    c.cW = 4 * L.etan[o - 1] * L.idx[j - 1] * rx;       // to vx(i,j-1)
    const double eS = L.etas[o + L.ld], eN = L.etas[o];
    c.cS = 2 * eS * L.idzc[i + 1] * rz;                 // to vx(i+1,j), shear at node (i+1,j)
    c.cN = 2 * eN * L.idzc[i] * rz;                     // to vx(i-1,j), shear at node (i,j)
    c.xS = 2 * eS * rx * rz;
    c.xN = 2 * eN * rx * rz;
    double wall = 0;
    if (L.proper) {
        // wall nodes: free slip = no shear stress; no slip = mirrored ghost value -vx, i.e. the
        // wall shear 2*eta*vx/dz acts on the diagonal only
        if (i + 1 == L.nz - 1) {
            if (L.ns_z1) wall += 2 * eS * L.idz[i] * rz;
            c.cS = 0, c.xS = 0;
        }
        if (i == 0) {
            if (L.ns_z0) wall += 2 * eN * L.idz[i] * rz;
            c.cN = 0, c.xN = 0;
        }
    }
    c.diag = -(c.cE + c.cW + c.cS + c.cN + wall);
    return c;
}

// (K v)_vz at (i,j): viscous part of the z-momentum row
__device__ __forceinline__ double kvz_apply(const LevelDev& L, const VzCoef& c,
                                            const double* __restrict__ vz,
                                            const double* __restrict__ vx, int i, int j) {
    const long long o = (long long)i * L.ld + j;
    const double w = (j > 0) ? vz[o - 1] : 0.0;
    return c.cN * vz[o + L.ld] + c.cS * vz[o - L.ld] + c.cE * vz[o + 1] + c.cW * w + c.diag * vz[o] +
           c.xE * (vx[o + 1] - vx[o - L.ld + 1]) - c.xW * (vx[o] - vx[o - L.ld]);
}

__device__ __forceinline__ double kvx_apply(const LevelDev& L, const VxCoef& c,
                                            const double* __restrict__ vz,
                                            const double* __restrict__ vx, int i, int j) {
    const long long o = (long long)i * L.ld + j;
    const double n = (i > 0) ? vx[o - L.ld] : 0.0;
    return c.cE * vx[o + 1] + c.cW * vx[o - 1] + c.cS * vx[o + L.ld] + c.cN * n + c.diag * vx[o] +
           c.xS * (vz[o + L.ld] - vz[o + L.ld - 1]) - c.xN * (vz[o] - vz[o - 1]);
}

// write a velocity value together with the slave(s) tied to it by the reference's wall rows
__device__ __forceinline__ void store_vz(const LevelDev& L, double* __restrict__ vz, int i, int j, double v) {
    const long long o = (long long)i * L.ld + j;
    vz[o] = v;
    if (!L.proper) {
        if (j == L.vz_j0) vz[o - 1] = v;                // vz(i,0) = vz(i,1), pylamp_stokes.py:249-255
        if (j == L.vz_j1) vz[o + 1] = v;                // vz(i,nxx-2) = vz(i,nxx-3), :296-301
    }
}
__device__ __forceinline__ void store_vx(const LevelDev& L, double* __restrict__ vx, int i, int j, double v) {
    const long long o = (long long)i * L.ld + j;
    vx[o] = v;
    if (!L.proper) {
        if (i == L.vx_i0) vx[o - L.ld] = L.sl_z0 * v;   // :163-175
        if (i == L.vx_i1) vx[o + L.ld] = L.sl_z1 * v;   // :202-214
        if (L.ft_x0 && j == L.vx_j0) {                  // dvx/dx = 0 on the x = 0 wall, rows i = 0 .. nz-2 (:268-273)
            vx[o - 1] = v;
            if (i == L.vx_i0) vx[o - L.ld - 1] = L.sl_z0 * v;
            if (i == L.vx_i1) vx[o + L.ld - 1] = L.sl_z1 * v;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Interior fast path.  For 1 <= i <= nz-2 and 1 <= j <= nxx-2 every neighbour index is in range, so
// all loads of both momentum rows are issued unconditionally and up front (one round of memory
// latency per thread instead of one per sub-expression; shared values are loaded once), and only
// the final stores are predicated on the row ranges.  Same arithmetic as vz_coef/kvz_apply and
// vx_coef/kvx_apply.
// ---------------------------------------------------------------------------------------------
struct Rows2 {
    double kz, dz;    // (K v)_vz and the vz diagonal
    double kx, dx;    // (K v)_vx and the vx diagonal
};

__device__ __forceinline__ bool is_interior(const LevelDev& L, int i, int j) {
    return i >= 1 && i <= L.nz - 2 && j >= 1 && j <= L.nxx - 2;
}

template <bool APPLY>
__device__ __forceinline__ Rows2 rows_interior(const LevelDev& L, const double* __restrict__ vz,
                                               const double* __restrict__ vx, int i, int j) {
    const int ld = L.ld;
    const long long o = (long long)i * ld + j;
    // coefficients
    const double enC = L.etan[o], enM0 = L.etan[o - ld], en0M = L.etan[o - 1];
    const double esC = L.etas[o], es0P = L.etas[o + 1], esP0 = L.etas[o + ld];
    const double idz_i = L.idz[i], idz_m = L.idz[i - 1], idzc_i = L.idzc[i], idzc_p = L.idzc[i + 1];
    const double idx_j = L.idx[j], idx_m = L.idx[j - 1], idxc_j = L.idxc[j], idxc_p = L.idxc[j + 1];
    double z00 = 0, zP0 = 0, zM0 = 0, z0P = 0, z0M = 0, zPM = 0, x00 = 0, x0P = 0, x0M = 0, xP0 = 0, xM0 = 0,
           xMP = 0;
    if (APPLY) {
        z00 = vz[o], zP0 = vz[o + ld], zM0 = vz[o - ld], z0P = vz[o + 1], z0M = vz[o - 1], zPM = vz[o + ld - 1];
        x00 = vx[o], x0P = vx[o + 1], x0M = vx[o - 1], xP0 = vx[o + ld], xM0 = vx[o - ld], xMP = vx[o - ld + 1];
    }
    Rows2 r;
    {   // z-momentum row
        double cN = 4 * enC * idz_i * idzc_i, cS = 4 * enM0 * idz_m * idzc_i;
        double cE = 2 * es0P * idxc_p * idx_j, cW = 2 * esC * idxc_j * idx_j;
        double xE = 2 * es0P * idzc_i * idx_j, xW = 2 * esC * idzc_i * idx_j;
        if (L.proper && j + 1 == L.nxx - 1) cE = 0, xE = 0;
        r.dz = -(cN + cS + cE + cW);
        r.kz = APPLY ? cN * zP0 + cS * zM0 + cE * z0P + cW * z0M + r.dz * z00 + xE * (x0P - xMP) - xW * (x00 - xM0)
                     : 0.0;
    }
    {   // x-momentum row
        double cE = 4 * enC * idx_j * idxc_j, cW = 4 * en0M * idx_m * idxc_j;
        double cS = 2 * esP0 * idzc_p * idz_i, cN = 2 * esC * idzc_i * idz_i;
        double xS = 2 * esP0 * idxc_j * idz_i, xN = 2 * esC * idxc_j * idz_i;
        double wall = 0;
        if (L.proper && i + 1 == L.nz - 1) {
            if (L.ns_z1) wall = 2 * esP0 * idz_i * idz_i;
            cS = 0, xS = 0;
        }
        r.dx = -(cE + cW + cS + cN + wall);
        r.kx = APPLY ? cE * x0P + cW * x0M + cS * xP0 + cN * xM0 + r.dx * x00 + xS * (zP0 - zPM) - xN * (z00 - z0M)
                     : 0.0;
    }
    return r;
}
