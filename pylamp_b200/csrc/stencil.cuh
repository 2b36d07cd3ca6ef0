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
};

__device__ __forceinline__ bool is_vz_row(const LevelDev& L, int i, int j) {
    return i >= L.vz_i0 && i <= L.vz_i1 && j >= L.vz_j0 && j <= L.vz_j1;
}
__device__ __forceinline__ bool is_vx_row(const LevelDev& L, int i, int j) {
    return i >= L.vx_i0 && i <= L.vx_i1 && j >= L.vx_j0 && j <= L.vx_j1;
}
// continuity rows: every real cell except the four corners and the pressure anchor (3,2)
__device__ __forceinline__ bool is_p_row(const LevelDev& L, int i, int j) {
    if (i > L.nz - 2 || j > L.nxx - 2) return false;
    bool corner = (i == 0 || i == L.nz - 2) && (j == 0 || j == L.nxx - 2);
    bool anchor = (i == 3 && j == 2);
    return !corner && !anchor;
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
    }
}
