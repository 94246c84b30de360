// okada_dev.cuh -- device-side Okada (1992) rectangular dislocation, gradient-only form.
//
// Replaces the call `dc3d(x, y, z, α, dep, dip, al1, al2, aw1, aw2, d1, d2, d3, cache)` made at
// /root/reference/src/BEM/GF.jl:49-54 and :156-161 (the arithmetic lives in the un-vendored
// GeoGreensFunctions.jl; this is written from the published tables of Okada 1992, BSSA 82(2)).
//
// B200-first restructuring (not a transliteration of DC3D):
//   * every caller in the hot path consumes only the 9 displacement GRADIENTS (GF.jl:77-85,163-169),
//     never the 3 displacements, so the log/atan terms of the closed form (which appear only in the
//     displacement rows) are never evaluated: one sqrt and a handful of reciprocals per corner;
//   * the slip type is a template parameter (unit_dislocation has one non-zero, GF.jl:73-74), which
//     removes two thirds of the table;
//   * divisions are hoisted into reciprocals shared across the table rows;
//   * the 1/(2π) prefactor is applied once by the caller after the image sum.
#pragma once
#include <math.h>

namespace oq {

constexpr double kOkadaEps = 1e-6;
constexpr double kInv2Pi = 0.15915494309189533576888376337251;

enum : int { kStrikeSlip = 0, kDipSlip = 1 };

struct OkadaMedium {
    double a1, a2, a3, a4, a5;     // (1-α)/2, α/2, (1-α)/α, 1-α, α
    double sd, cd, sdsd, cdcd, sdcd;
    double cdi;                    // 1/cd (unused when cd == 0)
};

// Host+device constructor; sd, cd are the host's sind/cosd of the dip (exact at multiples of 90°).
__host__ __device__ inline OkadaMedium make_okada_medium(double alpha, double sd, double cd)
{
    OkadaMedium m;
    m.a1 = (1.0 - alpha) / 2.0; m.a2 = alpha / 2.0; m.a3 = (1.0 - alpha) / alpha;
    m.a4 = 1.0 - alpha; m.a5 = alpha;
    if (fabs(cd) < kOkadaEps) { cd = 0.0; sd = (sd > 0.0) ? 1.0 : -1.0; }
    m.sd = sd; m.cd = cd; m.sdsd = sd * sd; m.cdcd = cd * cd; m.sdcd = sd * cd;
    m.cdi = (cd != 0.0) ? 1.0 / cd : 0.0;
    return m;
}

__device__ __forceinline__ double snap(double v) { return fabs(v) < kOkadaEps ? 0.0 : v; }

// Quantities shared by the A, B and C parts at one corner (ξ, η, q).
struct OkadaCorner {
    double xi, et, q, xi2, et2, q2;
    double ri, r3i, r5i, r;        // 1/R, 1/R^3, 1/R^5, R
    double y, d;
    double x11, x32, y11, y32;
    double ey, ez, fy, fz, gy, gz;
};

template <int SLIP>
__device__ __forceinline__ void corner_terms(const OkadaMedium& m, double xi, double et, double q,
                                             double r, bool kxi, bool ket, OkadaCorner& c)
{
    c.xi = xi; c.et = et; c.q = q;
    c.xi2 = xi * xi; c.et2 = et * et; c.q2 = q * q;
    c.r = r;
    c.ri = 1.0 / r;
    const double r2i = c.ri * c.ri;
    c.r3i = c.ri * r2i;
    c.r5i = c.r3i * r2i;
    c.y = et * m.cd + q * m.sd;
    c.d = et * m.sd - q * m.cd;
    if (kxi) { c.x11 = 0.0; c.x32 = 0.0; }
    else {
        const double rxi = r + xi;
        c.x11 = c.ri * (1.0 / rxi);        // reciprocal + multiply: cheaper than a full fp64 division
        c.x32 = (r + rxi) * c.x11 * c.x11 * c.ri;
    }
    if (ket) { c.y11 = 0.0; c.y32 = 0.0; }
    else {
        const double ret = r + et;
        c.y11 = c.ri * (1.0 / ret);
        c.y32 = (r + ret) * c.y11 * c.y11 * c.ri;
    }
    c.ey = m.sd * c.ri - c.y * q * c.r3i;
    c.ez = m.cd * c.ri + c.d * q * c.r3i;
    if (SLIP == kStrikeSlip) {
        c.fy = c.d * c.r3i + c.xi2 * c.y32 * m.sd;
        c.fz = c.y * c.r3i + c.xi2 * c.y32 * m.cd;
        c.gy = 0.0; c.gz = 0.0;
    } else {
        c.gy = 2.0 * c.x11 * m.sd - c.y * q * c.x32;
        c.gz = 2.0 * c.x11 * m.cd + c.d * q * c.x32;
        c.fy = 0.0; c.fz = 0.0;
    }
}

// Gradient rows (d/dx, d/dy', d/dz' in the fault-aligned frame) of the infinite-medium part.
template <int SLIP>
__device__ __forceinline__ void part_a(const OkadaMedium& m, const OkadaCorner& c, double (&A)[9])
{
    const double xy = c.xi * c.y11, qy = c.q * c.y11;
    if (SLIP == kStrikeSlip) {
        A[0] = -m.a1 * qy - m.a2 * c.xi2 * c.q * c.y32;
        A[1] = -m.a2 * c.xi * c.q * c.r3i;
        A[2] = m.a1 * xy + m.a2 * c.xi * c.q2 * c.y32;
        A[3] = m.a1 * xy * m.sd + m.a2 * c.xi * c.fy + 0.5 * c.d * c.x11;
        A[4] = m.a2 * c.ey;
        A[5] = m.a1 * (m.cd * c.ri + qy * m.sd) - m.a2 * c.q * c.fy;
        A[6] = m.a1 * xy * m.cd + m.a2 * c.xi * c.fz + 0.5 * c.y * c.x11;
        A[7] = m.a2 * c.ez;
        A[8] = -m.a1 * (m.sd * c.ri - qy * m.cd) - m.a2 * c.q * c.fz;
    } else {
        A[0] = -m.a2 * c.xi * c.q * c.r3i;
        A[1] = -0.5 * qy - m.a2 * c.et * c.q * c.r3i;
        A[2] = m.a1 * c.ri + m.a2 * c.q2 * c.r3i;
        A[3] = m.a2 * c.ey;
        A[4] = m.a1 * c.d * c.x11 + 0.5 * xy * m.sd + m.a2 * c.et * c.gy;
        A[5] = m.a1 * c.y * c.x11 - m.a2 * c.q * c.gy;
        A[6] = m.a2 * c.ez;
        A[7] = m.a1 * c.y * c.x11 + 0.5 * xy * m.cd + m.a2 * c.et * c.gz;
        A[8] = -m.a1 * c.d * c.x11 - m.a2 * c.q * c.gz;
    }
}

// Free-surface part (image source only).
template <int SLIP>
__device__ __forceinline__ void part_b(const OkadaMedium& m, const OkadaCorner& c, double (&B)[9])
{
    const double rd = c.r + c.d;
    const double rdi = 1.0 / rd;
    const double d11 = c.ri * rdi;
    const double aj2 = c.xi * c.y * rdi * d11;
    const double aj5 = -(c.d + c.y * c.y * rdi) * d11;
    double ak1, ak3, aj3, aj6;
    if (m.cd != 0.0) {
        ak1 = c.xi * (d11 - c.y11 * m.sd) * m.cdi;
        ak3 = (c.q * c.y11 - c.y * d11) * m.cdi;
        aj3 = (ak1 - aj2 * m.sd) * m.cdi;
        aj6 = (ak3 - aj5 * m.sd) * m.cdi;
    } else {
        const double rd2i = rdi * rdi;
        ak1 = c.xi * c.q * rdi * d11;
        ak3 = m.sd * rdi * (c.xi2 * d11 - 1.0);
        aj3 = -c.xi * rd2i * (c.q2 * d11 - 0.5);
        aj6 = -c.y * rd2i * (c.xi2 * d11 - 0.5);
    }
    const double xy = c.xi * c.y11, qy = c.q * c.y11;
    const double ak2 = c.ri + ak3 * m.sd;
    const double ak4 = xy * m.cd - ak1 * m.sd;
    const double aj1 = aj5 * m.cd - aj6 * m.sd;
    const double aj4 = -xy - aj2 * m.cd + aj3 * m.sd;
    if (SLIP == kStrikeSlip) {
        const double a3s = m.a3 * m.sd;
        B[0] = c.xi2 * c.q * c.y32 - a3s * aj1;
        B[1] = c.xi * c.q * c.r3i - a3s * aj2;
        B[2] = -c.xi * c.q2 * c.y32 - a3s * aj3;
        B[3] = -c.xi * c.fy - c.d * c.x11 + a3s * (xy + aj4);
        B[4] = -c.ey + a3s * (c.ri + aj5);
        B[5] = c.q * c.fy - a3s * (qy - aj6);
        B[6] = -c.xi * c.fz - c.y * c.x11 + a3s * ak1;
        B[7] = -c.ez + a3s * c.y * d11;
        B[8] = c.q * c.fz + a3s * ak2;
    } else {
        const double a3sc = m.a3 * m.sdcd;
        B[0] = c.xi * c.q * c.r3i + a3sc * aj4;
        B[1] = c.et * c.q * c.r3i + qy + a3sc * aj5;
        B[2] = -c.q2 * c.r3i + a3sc * aj6;
        B[3] = -c.ey + a3sc * aj1;
        B[4] = -c.et * c.gy - xy * m.sd + a3sc * aj2;
        B[5] = c.q * c.gy + a3sc * aj3;
        B[6] = -c.ez - a3sc * ak3;
        B[7] = -c.et * c.gz - xy * m.cd - a3sc * c.xi * d11;
        B[8] = c.q * c.gz - a3sc * ak4;
    }
}

// Depth-dependent part (image source only).  C0[3] are the displacement rows, which enter the
// z-derivative of the total field (Okada 1992 eq. for du/dz); C[9] are the gradient rows.
template <int SLIP>
__device__ __forceinline__ void part_c(const OkadaMedium& m, const OkadaCorner& c, double z,
                                       double (&C0)[3], double (&C)[9])
{
    const double cc = c.d + z;
    const double r2i = c.ri * c.ri;
    const double h = c.q * m.cd - z;
    const double y53 = (8.0 * c.r * c.r + 9.0 * c.r * c.et + 3.0 * c.et2) * c.y11 * c.y11 * c.y11 * r2i;
    const double z32 = m.sd * c.r3i - h * c.y32;
    const double z53 = 3.0 * m.sd * c.r5i - h * y53;
    const double y0 = c.y11 - c.xi2 * c.y32;
    const double z0 = z32 - c.xi2 * z53;
    const double xy = c.xi * c.y11, qy = c.q * c.y11;
    const double qr = 3.0 * c.q * c.r5i;
    const double cdr = (cc + c.d) * c.r3i;
    if (SLIP == kStrikeSlip) {
        const double ppy = m.cd * c.r3i + c.q * c.y32 * m.sd;
        const double ppz = m.sd * c.r3i - c.q * c.y32 * m.cd;
        const double qq = z * c.y32 + z32 + z0;
        const double qqy = 3.0 * cc * c.d * c.r5i - qq * m.sd;
        const double qqz = 3.0 * cc * c.y * c.r5i - qq * m.cd + c.q * c.y32;
        const double yy0 = c.y * c.r3i - y0 * m.cd;
        C0[0] = m.a4 * xy * m.cd - m.a5 * c.xi * c.q * z32;
        C0[1] = m.a4 * (m.cd * c.ri + 2.0 * qy * m.sd) - m.a5 * cc * c.q * c.r3i;
        C0[2] = m.a4 * qy * m.cd - m.a5 * (cc * c.et * c.r3i - z * c.y11 + c.xi2 * z32);
        C[0] = m.a4 * y0 * m.cd - m.a5 * c.q * z0;
        C[1] = -m.a4 * c.xi * (m.cd * c.r3i + 2.0 * c.q * c.y32 * m.sd) + m.a5 * cc * c.xi * qr;
        C[2] = -m.a4 * c.xi * c.q * c.y32 * m.cd + m.a5 * c.xi * (3.0 * cc * c.et * c.r5i - qq);
        C[3] = -m.a4 * c.xi * ppy * m.cd - m.a5 * c.xi * qqy;
        C[4] = m.a4 * 2.0 * (c.d * c.r3i - y0 * m.sd) * m.sd - c.y * c.r3i * m.cd
               - m.a5 * (cdr * m.sd - c.et * c.r3i - cc * c.y * qr);
        C[5] = -m.a4 * c.q * c.r3i + yy0 * m.sd + m.a5 * (cdr * m.cd + cc * c.d * qr - (y0 * m.cd + c.q * z0) * m.sd);
        C[6] = m.a4 * c.xi * ppz * m.cd - m.a5 * c.xi * qqz;
        C[7] = m.a4 * 2.0 * (c.y * c.r3i - y0 * m.cd) * m.sd + c.d * c.r3i * m.cd - m.a5 * (cdr * m.cd + cc * c.d * qr);
        C[8] = yy0 * m.cd - m.a5 * (cdr * m.sd - cc * c.y * qr - y0 * m.sdsd + c.q * z0 * m.cd);
    } else {
        const double x53 = (8.0 * c.r * c.r + 9.0 * c.r * c.xi + 3.0 * c.xi2) * c.x11 * c.x11 * c.x11 * r2i;
        const double ppy = m.cd * c.r3i + c.q * c.y32 * m.sd;
        const double ppz = m.sd * c.r3i - c.q * c.y32 * m.cd;
        C0[0] = m.a4 * m.cd * c.ri - qy * m.sd - m.a5 * cc * c.q * c.r3i;
        C0[1] = m.a4 * c.y * c.x11 - m.a5 * cc * c.et * c.q * c.x32;
        C0[2] = -c.d * c.x11 - xy * m.sd - m.a5 * cc * (c.x11 - c.q2 * c.x32);
        C[0] = -m.a4 * c.xi * c.r3i * m.cd + m.a5 * cc * c.xi * qr + c.xi * c.q * c.y32 * m.sd;
        C[1] = -m.a4 * c.y * c.r3i + m.a5 * cc * c.et * qr;
        C[2] = c.d * c.r3i - y0 * m.sd + m.a5 * cc * c.r3i * (1.0 - 3.0 * c.q2 * r2i);
        C[3] = -m.a4 * c.et * c.r3i + y0 * m.sdsd - m.a5 * (cdr * m.sd - cc * c.y * qr);
        C[4] = m.a4 * (c.x11 - c.y * c.y * c.x32) - m.a5 * cc * ((c.d + 2.0 * c.q * m.cd) * c.x32 - c.y * c.et * c.q * x53);
        C[5] = c.xi * ppy * m.sd + c.y * c.d * c.x32 + m.a5 * cc * ((c.y + 2.0 * c.q * m.sd) * c.x32 - c.y * c.q2 * x53);
        C[6] = -c.q * c.r3i + y0 * m.sdcd - m.a5 * (cdr * m.cd + cc * c.d * qr);
        C[7] = m.a4 * c.y * c.d * c.x32 - m.a5 * cc * ((c.y - 2.0 * c.q * m.sd) * c.x32 + c.d * c.et * c.q * x53);
        C[8] = -c.xi * ppz * m.sd + c.x11 - c.d * c.d * c.x32 - m.a5 * cc * ((c.d - 2.0 * c.q * m.cd) * c.x32 - c.d * c.q2 * x53);
    }
}

// One half (real or image source) of the solution: sums the four corners with Chinnery signs into
// g[9] = (uxx,uyx,uzx, uxy,uyy,uzy, uxz,uyz,uzz) WITHOUT the 1/(2π) factor.
// Returns false if the receiver sits on a fault edge (the closed form is singular; DC3D returns zeros).
template <int SLIP, bool IMAGE>
__device__ __forceinline__ bool okada_half(const OkadaMedium& m, double x, double y, double z, double dd,
                                           double al1, double al2, double aw1, double aw2, double (&g)[9])
{
    const double xi0 = snap(x - al1), xi1 = snap(x - al2);
    const double p = y * m.cd + dd * m.sd;
    const double q = snap(y * m.sd - dd * m.cd);
    const double et0 = snap(p - aw1), et1 = snap(p - aw2);
    if (q == 0.0 && ((xi0 * xi1 <= 0.0 && et0 * et1 == 0.0) || (et0 * et1 <= 0.0 && xi0 * xi1 == 0.0)))
        return false;
    const double q2 = q * q;
    const double r11 = sqrt(xi0 * xi0 + et0 * et0 + q2);
    const double r12 = sqrt(xi0 * xi0 + et1 * et1 + q2);   // (xi0, et1)
    const double r21 = sqrt(xi1 * xi1 + et0 * et0 + q2);   // (xi1, et0)
    const double r22 = sqrt(xi1 * xi1 + et1 * et1 + q2);
    // receiver on the negative extension of a fault edge: the regular 1/(R+ξ), 1/(R+η) terms vanish
    const bool kxi0 = xi0 < 0.0 && r21 + xi1 < kOkadaEps;   // used with η = et0
    const bool kxi1 = xi0 < 0.0 && r22 + xi1 < kOkadaEps;   // used with η = et1
    const bool ket0 = et0 < 0.0 && r12 + et1 < kOkadaEps;   // used with ξ = xi0
    const bool ket1 = et0 < 0.0 && r22 + et1 < kOkadaEps;   // used with ξ = xi1

#pragma unroll
    for (int k = 0; k < 2; ++k) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const double xi = j ? xi1 : xi0, et = k ? et1 : et0;
            const double r = k ? (j ? r22 : r12) : (j ? r21 : r11);
            const bool kxi = k ? kxi1 : kxi0, ket = j ? ket1 : ket0;
            const double sgn = (j + k == 1) ? -1.0 : 1.0;
            OkadaCorner c;
            corner_terms<SLIP>(m, xi, et, q, r, kxi, ket, c);
            double A[9];
            part_a<SLIP>(m, c, A);
            if (!IMAGE) {
#pragma unroll
                for (int i = 0; i < 9; i += 3) {
                    double s = (i == 6) ? -sgn : sgn;      // d/dz block flips sign for the real source
                    g[i]     += s * (-A[i]);
                    g[i + 1] += s * (-A[i + 1] * m.cd + A[i + 2] * m.sd);
                    g[i + 2] += s * (-A[i + 1] * m.sd - A[i + 2] * m.cd);
                }
            } else {
                double B[9], C0[3], C[9];
                part_b<SLIP>(m, c, B);
                part_c<SLIP>(m, c, z, C0, C);
#pragma unroll
                for (int i = 0; i < 9; i += 3) {
                    const double ab0 = A[i] + B[i], ab1 = A[i + 1] + B[i + 1], ab2 = A[i + 2] + B[i + 2];
                    double d0 = ab0 + z * C[i];
                    double d1 = (ab1 + z * C[i + 1]) * m.cd - (ab2 + z * C[i + 2]) * m.sd;
                    double d2 = (ab1 - z * C[i + 1]) * m.sd + (ab2 - z * C[i + 2]) * m.cd;
                    if (i == 6) {
                        d0 += C0[0];
                        d1 += C0[1] * m.cd - C0[2] * m.sd;
                        d2 -= C0[1] * m.sd + C0[2] * m.cd;
                    }
                    g[i] += sgn * d0; g[i + 1] += sgn * d1; g[i + 2] += sgn * d2;
                }
            }
        }
    }
    return true;
}

// Full dc3d gradient for one source rectangle: real + image halves.  g must be zero-initialised by
// the caller if it is not accumulating.  On a singular receiver nothing is added (DC3D's IRET=1);
// a receiver above the free surface (z > 0) adds nothing either (IRET=2).
template <int SLIP>
__device__ __forceinline__ void okada_gradient(const OkadaMedium& m, double x, double y, double z, double dep,
                                               double al1, double al2, double aw1, double aw2, double (&g)[9])
{
    if (z > 0.0) return;
    double t[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) t[i] = 0.0;
    if (!okada_half<SLIP, false>(m, x, y, z, dep + z, al1, al2, aw1, aw2, t)) return;
    if (!okada_half<SLIP, true>(m, x, y, z, dep - z, al1, al2, aw1, aw2, t)) return;
#pragma unroll
    for (int i = 0; i < 9; ++i) g[i] += t[i];
}

}  // namespace oq
