// hex8_dev.cuh -- device-side stress of a uniformly strained cuboid in a half-space.
//
// Replaces `stress_vol_hex8!(out, x,y,z, qx,qy,qz, Δx,Δy,Δz, 0, ε..., μ, ν)` as called at
// /root/reference/src/BEM/GF.jl:215-221 and :277-283 (GeoGreensFunctions.jl, Barbot et al. 2017; un-vendored).
// The closed form is this repository's own derivation from the kernel's definition (derive/hex8_derive.py
// generates hex8_gen.cuh): the strain is a linear combination -- with coefficients that depend only on the
// elastic constant α and the receiver depth -- of 8-corner signed sums of 21 (real source) + 70 (image source)
// basis functions, each a derivative of the triple antiderivative of 1/R, R or R - R3 ln(R+R3).
//
// B200-first restructuring:
//   * the field is linear in the eigenstrain, so ONE geometry evaluation serves all six unit strains (the
//     reference re-evaluates the whole kernel six times, GF.jl:206-208, :262-264);
//   * per corner only the basis functions are evaluated (polynomials in the corner vector, 7 reciprocals,
//     3 logs, 3+2 atans); the 36 x 91 combination is applied once per pair, after the corner sums;
//   * the 91 running sums live in shared memory (basis-major, conflict-free), which keeps the register file for
//     the arithmetic: no spills.
#pragma once
#include <math.h>

#include "hex8_gen.cuh"

// Optimisation barrier expanded at the start of every group of basis functions: the inputs pass through an empty
// asm, so temporaries of one group cannot be kept alive for the next (each group is register-allocated on its
// own; the hoisting/sharing across groups is what pushed the image basis to 250 registers).
#ifndef HEX8_GROUP_BARRIER
#define HEX8_GROUP_BARRIER                                                                                         \
    asm volatile("" : "+d"(R1), "+d"(R2), "+d"(R3), "+d"(R), "+d"(iR), "+d"(iw1), "+d"(iw2), "+d"(iw3), "+d"(iq1),  \
                 "+d"(iq2), "+d"(iq3), "+d"(q1), "+d"(q2), "+d"(q3));
#endif

namespace oq {

#ifndef OQ_HEX8_THREADS
#define OQ_HEX8_THREADS 128
#endif
#ifndef OQ_HEX8_MINB
#define OQ_HEX8_MINB 2
#endif
constexpr int kHex8Threads = OQ_HEX8_THREADS;
constexpr int kHex8Acc = HEX8_NB_REAL + HEX8_NB_IMAGE;
constexpr size_t kHex8SmemBytes = (size_t)kHex8Acc * kHex8Threads * sizeof(double);

struct Hex8Corner {
    double R, iR, w[3], q[3], iw[3], iq[3], L[3], A[3];
};

// The closed form is singular on the lines through the cuboid's edges (two components of the corner vector
// vanish); the field itself is regular there outside the cuboid.  Receivers within `nudge` of such a line are
// moved off it along ONE axis (keeping the other component exactly zero keeps the cancelling terms exactly
// zero); the reference formulas return non-finite values at these points.
__host__ __device__ inline void hex8_regularise(double& r1, double& r2, double& r3, double nudge)
{
    const bool t1 = fabs(r1) < nudge, t2 = fabs(r2) < nudge, t3 = fabs(r3) < nudge;
    if (t1 && (t2 || t3)) r1 = nudge;
    else if (t2 && t3) r2 = nudge;
}

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
        // w = R + R_k without cancellation when R_k < 0: with d = R + |R_k|, w = d or q/d and 1/w = 1/d or d/q --
        // one reciprocal serves both branches
        const double d = n + fabs(rs[k]);
        const double rd = 1.0 / d;
        c.iq[k] = 1.0 / c.q[k];
        const bool pos = rs[k] >= 0.0;
        c.w[k] = pos ? d : c.q[k] * rd;
        c.iw[k] = pos ? rd : d * c.iq[k];
        c.L[k] = log(c.w[k]);
        c.A[k] = atan(rs[a] * rs[b] / (rs[k] * n));
    }
}

template <int NEED>
__device__ __forceinline__ void hex8_basis_real(double R1, double R2, double R3, const Hex8Corner& c, double sgn,
                                                double* acc)
{
    double R = c.R, iR = c.iR, w1 = c.w[0], w2 = c.w[1], w3 = c.w[2], q1 = c.q[0], q2 = c.q[1], q3 = c.q[2];
    double iw1 = c.iw[0], iw2 = c.iw[1], iw3 = c.iw[2], iq1 = c.iq[0], iq2 = c.iq[1], iq3 = c.iq[2];
    const double L1 = c.L[0], L2 = c.L[1], L3 = c.L[2], A1 = c.A[0], A2 = c.A[1], A3 = c.A[2];
    (void)R; (void)iR; (void)w1; (void)w2; (void)w3; (void)q1; (void)q2; (void)q3; (void)iw1; (void)iw2; (void)iw3;
    (void)iq1; (void)iq2; (void)iq3; (void)L1; (void)L2; (void)L3; (void)A1; (void)A2; (void)A3;
#define ACC(b) (*(volatile double*)&acc[(b) * kHex8Threads])   // volatile: keeps each load next to its use
#define HEX8_NEED(mask) (((mask) & NEED) != 0)
    HEX8_BASIS_REAL_BODY
#undef HEX8_NEED
#undef ACC
}

template <int NEED>
__device__ __forceinline__ void hex8_basis_image(double R1, double R2, double R3, const Hex8Corner& c, double Ba,
                                                 double Bb, double sgn, double* acc)
{
    double R = c.R, iR = c.iR, w1 = c.w[0], w2 = c.w[1], w3 = c.w[2], q1 = c.q[0], q2 = c.q[1], q3 = c.q[2];
    double iw1 = c.iw[0], iw2 = c.iw[1], iw3 = c.iw[2], iq1 = c.iq[0], iq2 = c.iq[1], iq3 = c.iq[2];
    const double L1 = c.L[0], L2 = c.L[1], L3 = c.L[2], A1 = c.A[0], A2 = c.A[1], A3 = c.A[2];
    (void)R; (void)iR; (void)w1; (void)w2; (void)w3; (void)q1; (void)q2; (void)q3; (void)iw1; (void)iw2; (void)iw3;
    (void)iq1; (void)iq2; (void)iq3; (void)L1; (void)L2; (void)L3; (void)A1; (void)A2; (void)A3; (void)Ba; (void)Bb;
#define ACC(b) (*(volatile double*)&acc[(b) * kHex8Threads])   // volatile: keeps each load next to its use
#define HEX8_NEED(mask) (((mask) & NEED) != 0)
    HEX8_BASIS_IMAGE_BODY
#undef HEX8_NEED
#undef ACC
}

// Q[36] (times 8πμ): strain component (il) per unit moment component (jk), pairs ordered xx,xy,xz,yy,yz,zz.
// `acc` points at this thread's column of the CTA's shared accumulator array [kHex8Acc][kHex8Threads].
// NEED: bit mask of the strain rows (il) the caller will read; the other rows of Q are left unspecified.
template <int NEED>
__device__ __forceinline__ void hex8_strain_kernels(double x, double y, double z, double qx, double qy, double qz,
                                                    double dx, double dy, double dz, double al, double* acc,
                                                    double (&Q)[36])
{
#pragma unroll 1
    for (int b = 0; b < kHex8Acc; ++b) acc[b * kHex8Threads] = 0.0;
    const double x0 = qx - 0.5 * dx;
    const double nudge = 1e-6 * fmin(dx, fmin(dy, dz));
#pragma unroll 1
    for (int corner = 0; corner < 8; ++corner) {
        const int c1 = corner & 1, c2 = (corner >> 1) & 1, c3 = corner >> 2;
        const double sgn = ((c1 + c2 + c3) & 1) ? 1.0 : -1.0;          // s1 s2 s3, s = -1 at the lower limit
        const double zc = c3 ? qz : qz - dz;
        Hex8Corner c;
        {
            double r1 = x - (c1 ? x0 + dx : x0), r2 = y - (c2 ? qy + dy : qy), r3 = z - zc;   // real source
            hex8_regularise(r1, r2, r3, nudge);
            hex8_corner_inputs(r1, r2, r3, c);
            hex8_basis_real<NEED>(r1, r2, r3, c, sgn, acc);
        }
        {
            double r1 = x - (c1 ? x0 + dx : x0), r2 = y - (c2 ? qy + dy : qy), r3 = -z - zc;  // image source
            hex8_regularise(r1, r2, r3, nudge);
            hex8_corner_inputs(r1, r2, r3, c);
            hex8_basis_image<NEED>(r1, r2, r3, c, atan(r1 / r2), atan(r2 / r1), sgn, acc + HEX8_NB_REAL * kHex8Threads);
        }
    }
    const double x3 = z, ial = 1.0 / al;
#define ACCR(b) acc[(b) * kHex8Threads]
#define ACCI(b) acc[(HEX8_NB_REAL + (b)) * kHex8Threads]
    HEX8_COMBINE_BODY
#undef ACCR
#undef ACCI
}

// ---- vertex form --------------------------------------------------------------------------------------------
// The strain kernels are LINEAR in the corner sums, so the 36 x 91 combination can be applied to the basis values of
// ONE corner: Qv[36] of a mesh VERTEX (sign +1).  A cell's kernels are then the signed sum of the Qv of its eight
// vertices -- and conforming cells share every vertex up to eight ways, so a tile of cells needs one basis
// evaluation per (receiver, vertex) instead of one per (receiver, cell, corner): ~4x fewer for a 4x4x4 tile
// (125 vertices for 64 cells = 512 corners).  `acc` as in hex8_strain_kernels.
template <int NEED>
__device__ __forceinline__ void hex8_vertex_kernels(double x, double y, double z, double vx, double vy, double vz,
                                                    double nudge, double al, double* acc, double (&Q)[36])
{
#pragma unroll 1
    for (int b = 0; b < kHex8Acc; ++b) acc[b * kHex8Threads] = 0.0;
    Hex8Corner c;
    {
        double r1 = x - vx, r2 = y - vy, r3 = z - vz;                 // real source
        hex8_regularise(r1, r2, r3, nudge);
        hex8_corner_inputs(r1, r2, r3, c);
        hex8_basis_real<NEED>(r1, r2, r3, c, 1.0, acc);
    }
    {
        double r1 = x - vx, r2 = y - vy, r3 = -z - vz;                // image source
        hex8_regularise(r1, r2, r3, nudge);
        hex8_corner_inputs(r1, r2, r3, c);
        hex8_basis_image<NEED>(r1, r2, r3, c, atan(r1 / r2), atan(r2 / r1), 1.0, acc + HEX8_NB_REAL * kHex8Threads);
    }
    const double x3 = z, ial = 1.0 / al;
#define ACCR(b) acc[(b) * kHex8Threads]
#define ACCI(b) acc[(HEX8_NB_REAL + (b)) * kHex8Threads]
    HEX8_COMBINE_BODY
#undef ACCR
#undef ACCI
}

// out(p, S) for the six unit eigenstrains from the strain kernels Q of one (receiver, cuboid) pair; `inside`: the
// receiver lies strictly inside the cuboid (the eigenstrain itself is subtracted there)
template <class Out>
__device__ __forceinline__ void hex8_stress_from_kernels(const double (&Q)[36], bool inside, double mu, double nu, Out&& out)
{
    const double lam = 2.0 * mu * nu / (1.0 - 2.0 * nu);
    const double pref = 1.0 / (8.0 * 3.14159265358979323846 * mu);
#pragma unroll
    for (int p = 0; p < 6; ++p) {
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
        const double S[6] = {lam * ekk + 2.0 * mu * e[0], 2.0 * mu * e[1], 2.0 * mu * e[2],
                             lam * ekk + 2.0 * mu * e[3], 2.0 * mu * e[4], lam * ekk + 2.0 * mu * e[5]};
        out(p, S);
    }
}

// Calls out(p, S) with S[k] = stress component k (xx,xy,xz,yy,yz,zz) at (x,y,z) for unit eigenstrain component
// p = 0..5 of the cuboid x∈[qx-dx/2,qx+dx/2], y∈[qy,qy+dy], z∈[qz-dz,qz]  (mesh.jl:181-183, GF.jl:218).
// The six stresses of one p are handed over as soon as they exist, so callers never hold all 36.  With NEED a
// subset of strain rows, only S[k] for those rows k is meaningful (off-diagonal rows are independent; the
// diagonal rows 0,3,5 need each other through the trace).
template <int NEED = 0x3f, class Out>
__device__ __forceinline__ void hex8_stress_emit(double x, double y, double z, double qx, double qy, double qz,
                                                 double dx, double dy, double dz, double mu, double nu, double* acc,
                                                 Out&& out)
{
    const double lam = 2.0 * mu * nu / (1.0 - 2.0 * nu);
    const double alpha = (lam + mu) / (lam + 2.0 * mu);
    double Q[36];
    hex8_strain_kernels<NEED>(x, y, z, qx, qy, qz, dx, dy, dz, alpha, acc, Q);
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
        const double S[6] = {lam * ekk + 2.0 * mu * e[0], 2.0 * mu * e[1], 2.0 * mu * e[2],
                             lam * ekk + 2.0 * mu * e[3], 2.0 * mu * e[4], lam * ekk + 2.0 * mu * e[5]};
        out(p, S);
    }
}

}  // namespace oq
