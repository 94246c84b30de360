// hex8_dev.cuh -- device-side stress of a uniformly strained cuboid in a half-space.
//
// Replaces `stress_vol_hex8!(out, x,y,z, qx,qy,qz, Δx,Δy,Δz, 0, ε..., μ, ν)` as called at
// /root/reference/src/BEM/GF.jl:215-221 and :277-283 (GeoGreensFunctions.jl, Barbot et al. 2017; un-vendored).
// The closed form is this repository's own derivation from the kernel's definition
// (derive/hex8_derive.py generates hex8_gen.cuh): the strain is a signed sum over the 8 corners of
// the cuboid, for the real and the image source, of explicit functions of the corner vector.
//
// B200-first restructuring: the field is linear in the eigenstrain, so ONE geometry evaluation (36 strain
// kernels Q[(il),(jk)]) serves all six unit strains; the reference re-evaluates the whole kernel six times
// (GF.jl:206-208, :262-264).
#pragma once
#include <math.h>

#include "hex8_gen.cuh"

namespace oq {

struct Hex8Corner {
    double R, iR, w[3], q[3], iw[3], iq[3], L[3], A[3];
};

__device__ __forceinline__ void hex8_corner_inputs(double r1, double r2, double r3, Hex8Corner& c)
{
    const double rs[3] = {r1, r2, r3};
    const double s1 = r1 * r1, s2 = r2 * r2, s3 = r3 * r3;
    c.q[0] = s2 + s3; c.q[1] = s1 + s3; c.q[2] = s1 + s2;
    const double n = sqrt(c.q[0] + s1);
    c.R = n;
    c.iR = 1.0 / n;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int a = (k + 1) % 3, b = (k + 2) % 3;
        // R + R_k without cancellation when R_k < 0
        c.w[k] = rs[k] >= 0.0 ? n + rs[k] : c.q[k] / (n - rs[k]);
        c.L[k] = log(c.w[k]);
        c.A[k] = atan(rs[a] * rs[b] / (rs[k] * n));
        c.iw[k] = 1.0 / c.w[k];
        c.iq[k] = 1.0 / c.q[k];
    }
}

// Q[36] (times 8πμ): strain component (il) per unit moment component (jk), pairs ordered xx,xy,xz,yy,yz,zz
__device__ __forceinline__ void hex8_strain_kernels(double x, double y, double z, double qx, double qy, double qz,
                                                    double dx, double dy, double dz, double alpha, double (&Q)[36])
{
#pragma unroll
    for (int k = 0; k < 36; ++k) Q[k] = 0.0;
    const double x0 = qx - 0.5 * dx;
#pragma unroll 1
    for (int corner = 0; corner < 8; ++corner) {
        const int c1 = corner & 1, c2 = (corner >> 1) & 1, c3 = corner >> 2;
        const double sgn = ((c1 + c2 + c3) & 1) ? 1.0 : -1.0;          // s1 s2 s3, s = -1 at the lower limit
        const double r1 = x - (c1 ? x0 + dx : x0);
        const double r2 = y - (c2 ? qy + dy : qy);
        const double zc = c3 ? qz : qz - dz;
        Hex8Corner c;
        hex8_corner_inputs(r1, r2, z - zc, c);                          // real source
        hex8_corner_real(r1, r2, z - zc, c.R, c.w[0], c.w[1], c.w[2], c.q[0], c.q[1], c.q[2], c.iR, c.iw[0], c.iw[1], c.iw[2],
                         c.iq[0], c.iq[1], c.iq[2], c.L[0], c.L[1], c.L[2],
                         c.A[0], c.A[1], c.A[2], alpha, sgn, Q);
        const double r3 = -z - zc;                                      // image source
        hex8_corner_inputs(r1, r2, r3, c);
        hex8_corner_image(r1, r2, r3, c.R, c.w[0], c.w[1], c.w[2], c.q[0], c.q[1], c.q[2], c.iR, c.iw[0], c.iw[1], c.iw[2],
                         c.iq[0], c.iq[1], c.iq[2], c.L[0], c.L[1], c.L[2],
                          c.A[0], c.A[1], c.A[2], atan(r1 / r2), atan(r2 / r1), z, alpha, sgn, Q);
    }
}

// S[p][k]: stress component k (xx,xy,xz,yy,yz,zz) at (x,y,z) for unit eigenstrain component p of the cuboid
// x∈[qx-dx/2,qx+dx/2], y∈[qy,qy+dy], z∈[qz-dz,qz]  (mesh.jl:181-183, GF.jl:218).
__device__ __forceinline__ void hex8_stress_all(double x, double y, double z, double qx, double qy, double qz,
                                                double dx, double dy, double dz, double mu, double nu,
                                                double (&S)[6][6])
{
    const double lam = 2.0 * mu * nu / (1.0 - 2.0 * nu);
    const double alpha = (lam + mu) / (lam + 2.0 * mu);
    double Q[36];
    hex8_strain_kernels(x, y, z, qx, qy, qz, dx, dy, dz, alpha, Q);
    const double pref = 1.0 / (8.0 * 3.14159265358979323846 * mu);
    const bool inside = x > qx - 0.5 * dx && x < qx + 0.5 * dx && y > qy && y < qy + dy && z > qz - dz && z < qz;
#pragma unroll
    for (int p = 0; p < 6; ++p) {
        // moment density of the unit eigenstrain p: m = λ tr(ε) I + 2 μ ε
        const bool diag = (p == 0 || p == 3 || p == 5);
        double e[6];
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            double s = 2.0 * mu * Q[6 * a + p];
            if (diag) s += lam * (Q[6 * a + 0] + Q[6 * a + 3] + Q[6 * a + 5]);
            e[a] = pref * s;
        }
        if (inside) e[p] -= 1.0;
        const double ekk = e[0] + e[3] + e[5];
        S[p][0] = lam * ekk + 2.0 * mu * e[0];
        S[p][1] = 2.0 * mu * e[1];
        S[p][2] = 2.0 * mu * e[2];
        S[p][3] = lam * ekk + 2.0 * mu * e[3];
        S[p][4] = 2.0 * mu * e[4];
        S[p][5] = lam * ekk + 2.0 * mu * e[5];
    }
}

}  // namespace oq
