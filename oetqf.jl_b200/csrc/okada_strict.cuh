// okada_strict.cuh -- device-side Okada (1992) DC3D gradient rows in the PUBLISHED OPERATION ORDER.
//
// Replaces the call `dc3d(x, y, z, α, dep, dip, al1, al2, aw1, aw2, d1, d2, d3, cache)` made at
// /root/reference/src/BEM/GF.jl:49-54 and :156-161 (arithmetic in the un-vendored GeoGreensFunctions.jl, a
// transcription of Okada's DC3D; specification: SURVEY.md Appendix A).
//
// Why a second form next to okada_dev.cuh: the closed form is badly conditioned far from the source (R + ξ with
// ξ ≈ -R, differences of 1/(R(R+ξ)) between neighbouring corners): at 1000 cell sizes the fp64 result carries
// relative errors up to 1e-2 of the (tiny) entry, and two fp64 evaluations agree to 1e-10 per entry only if they
// round identically.  This form therefore keeps every division, every product order and every sum order of the
// published routine (DCCON0/DCCON2/UA/UB/UC/DC3D), is compiled WITHOUT floating-point contraction (own
// translation unit, --fmad=false), and uses only IEEE-exact operations (+ - * / sqrt) -- the log/atan terms of
// DC3D occur only in the displacement rows, which no caller of the hot path reads (GF.jl:77-85,163-169).  Its
// results are bit-identical to a scalar CPU evaluation of the same routine; okada_dev.cuh (reciprocals shared,
// FMA-contracted, ~2x fewer fp64 instructions) is kept as the opt-in fast form (OQ_OKADA=fast).
//
// Only the rows of the selected slip type are evaluated (unit_dislocation has one non-zero, GF.jl:73-74), and
// rows the caller does not read are removed by the compiler (everything is inlined, no contraction to respect).
#pragma once
#include <math.h>

#include "okada_dev.cuh"   // OkadaMedium, kOkadaEps, kStrikeSlip / kDipSlip

namespace oq {

constexpr double kOkadaPi2 = 6.283185307179586476925286766559;

// Correctly rounded quotients that SHARE the reciprocal of their denominator.  With y = RN(1/b) (one IEEE
// division), q0 = RN(a*y) is within one ulp of a/b, the residual r = a - b*q0 is exact in one FMA, and
// RN(q0 + r*y) is the correctly rounded quotient a/b (Markstein's theorem; it holds for every normal b whose
// significand is not all ones -- probability 2^-52 per division).  DC3D divides ~60 times per corner by a
// handful of denominators (R, R^3, R^5, R^2, R+d, (R+d)^2, cos(dip)); this keeps every quotient bit-identical to
// `a / b` for 3 flops instead of a ~20-instruction division sequence.
struct Recip {
    double b, y;
    __device__ __forceinline__ explicit Recip(double den) : b(den), y(1.0 / den) {}
};
__device__ __forceinline__ double sdiv(double a, const Recip& r)
{
    const double q0 = a * r.y;
    const double res = fma(-r.b, q0, a);
    return fma(res, r.y, q0);
}

struct StrictGeo {
    double xi, et, q, xi2, et2, q2, r, r2, r3, r5, y, d;
    double x11, y11, x32, y32;
    double ey, ez, fy, fz, gy, gz;
};

// reciprocals of the denominators a corner divides by repeatedly
struct StrictRecips {
    Recip r, r2, r3, r5;
    __device__ __forceinline__ explicit StrictRecips(const StrictGeo& g) : r(g.r), r2(g.r2), r3(g.r3), r5(g.r5) {}
};

__device__ __forceinline__ void strict_corner_base(const OkadaMedium& m, double xi, double et, double q, StrictGeo& g)
{
    const double sd = m.sd, cd = m.cd;
    g.xi = xi; g.et = et; g.q = q;
    g.xi2 = xi * xi; g.et2 = et * et; g.q2 = q * q;
    g.r2 = g.xi2 + g.et2 + g.q2;
    g.r = sqrt(g.r2);
    g.r3 = g.r * g.r2;
    g.r5 = g.r3 * g.r2;
    g.y = et * cd + q * sd;
    g.d = et * sd - q * cd;
}

__device__ __forceinline__ void strict_corner(const OkadaMedium& m, bool kxi, bool ket, const StrictRecips& R,
                                              StrictGeo& g)
{
    const double sd = m.sd, cd = m.cd, xi = g.xi, et = g.et, q = g.q;
    if (kxi) { g.x11 = 0.0; g.x32 = 0.0; }
    else {
        const double rxi = g.r + xi;
        g.x11 = 1.0 / (g.r * rxi);
        g.x32 = sdiv((g.r + rxi) * g.x11 * g.x11, R.r);
    }
    if (ket) { g.y11 = 0.0; g.y32 = 0.0; }
    else {
        const double ret = g.r + et;
        g.y11 = 1.0 / (g.r * ret);
        g.y32 = sdiv((g.r + ret) * g.y11 * g.y11, R.r);
    }
    g.ey = sdiv(sd, R.r) - sdiv(g.y * q, R.r3);
    g.ez = sdiv(cd, R.r) + sdiv(g.d * q, R.r3);
    g.fy = sdiv(g.d, R.r3) + g.xi2 * g.y32 * sd;
    g.fz = sdiv(g.y, R.r3) + g.xi2 * g.y32 * cd;
    g.gy = 2.0 * g.x11 * sd - g.y * q * g.x32;
    g.gz = 2.0 * g.x11 * cd + g.d * q * g.x32;
}

// rows 3..11 of UA for one slip type, already scaled by disl/(2π) with disl = 1
template <int SLIP>
__device__ __forceinline__ void strict_ua(const OkadaMedium& m, const StrictGeo& g, const StrictRecips& R, double (&u)[9])
{
    const double xi = g.xi, et = g.et, q = g.q, xi2 = g.xi2, q2 = g.q2;
    const double y = g.y, d = g.d, x11 = g.x11, y11 = g.y11, y32 = g.y32;
    const double ey = g.ey, ez = g.ez, fy = g.fy, fz = g.fz, gy = g.gy, gz = g.gz;
    const double a1 = m.a1, a2 = m.a2, sd = m.sd, cd = m.cd;
    const double xy = xi * y11, qy = q * y11;
    const double f = 1.0 / kOkadaPi2;
    if (SLIP == kStrikeSlip) {
        u[0] = f * (-a1 * qy - a2 * xi2 * q * y32);
        u[1] = f * (sdiv(-a2 * xi * q, R.r3));
        u[2] = f * (a1 * xy + a2 * xi * q2 * y32);
        u[3] = f * (a1 * xy * sd + a2 * xi * fy + d / 2 * x11);
        u[4] = f * (a2 * ey);
        u[5] = f * (a1 * (sdiv(cd, R.r) + qy * sd) - a2 * q * fy);
        u[6] = f * (a1 * xy * cd + a2 * xi * fz + y / 2 * x11);
        u[7] = f * (a2 * ez);
        u[8] = f * (-a1 * (sdiv(sd, R.r) - qy * cd) - a2 * q * fz);
    } else {
        u[0] = f * (sdiv(-a2 * xi * q, R.r3));
        u[1] = f * (-qy / 2 - sdiv(a2 * et * q, R.r3));
        u[2] = f * (sdiv(a1, R.r) + sdiv(a2 * q2, R.r3));
        u[3] = f * (a2 * ey);
        u[4] = f * (a1 * d * x11 + xy / 2 * sd + a2 * et * gy);
        u[5] = f * (a1 * y * x11 - a2 * q * gy);
        u[6] = f * (a2 * ez);
        u[7] = f * (a1 * y * x11 + xy / 2 * cd + a2 * et * gz);
        u[8] = f * (-a1 * d * x11 - a2 * q * gz);
    }
}

// rows 3..11 of UB
template <int SLIP>
__device__ __forceinline__ void strict_ub(const OkadaMedium& m, const StrictGeo& g, const StrictRecips& R, double (&u)[9])
{
    const double xi = g.xi, et = g.et, q = g.q, xi2 = g.xi2, q2 = g.q2, r = g.r;
    const double y = g.y, d = g.d, x11 = g.x11, y11 = g.y11, y32 = g.y32;
    const double ey = g.ey, ez = g.ez, fy = g.fy, fz = g.fz, gy = g.gy, gz = g.gz;
    const double a3 = m.a3, sd = m.sd, cd = m.cd, sdcd = m.sdcd;
    const double rd = r + d, d11 = 1.0 / (r * rd);
    const Recip Rrd(rd);
    const double aj2 = sdiv(xi * y, Rrd) * d11, aj5 = -(d + sdiv(y * y, Rrd)) * d11;
    double ak1, ak3, aj3, aj6;
    if (cd != 0.0) {
        const Recip Rcd(cd);
        ak1 = sdiv(xi * (d11 - y11 * sd), Rcd);
        ak3 = sdiv(q * y11 - y * d11, Rcd);
        aj3 = sdiv(ak1 - aj2 * sd, Rcd);
        aj6 = sdiv(ak3 - aj5 * sd, Rcd);
    } else {
        const double rd2 = rd * rd;
        const Recip Rrd2(rd2);
        ak1 = sdiv(xi * q, Rrd) * d11;
        ak3 = sdiv(sd, Rrd) * (xi2 * d11 - 1.0);
        aj3 = sdiv(-xi, Rrd2) * (q2 * d11 - 0.5);
        aj6 = sdiv(-y, Rrd2) * (xi2 * d11 - 0.5);
    }
    const double xy = xi * y11;
    const double ir = R.r.y;                                 // 1.0 / r
    const double ak2 = ir + ak3 * sd;
    const double ak4 = xy * cd - ak1 * sd;
    const double aj1 = aj5 * cd - aj6 * sd;
    const double aj4 = -xy - aj2 * cd + aj3 * sd;
    const double qy = q * y11;
    const double f = 1.0 / kOkadaPi2;
    if (SLIP == kStrikeSlip) {
        u[0] = f * (xi2 * q * y32 - a3 * aj1 * sd);
        u[1] = f * (sdiv(xi * q, R.r3) - a3 * aj2 * sd);
        u[2] = f * (-xi * q2 * y32 - a3 * aj3 * sd);
        u[3] = f * (-xi * fy - d * x11 + a3 * (xy + aj4) * sd);
        u[4] = f * (-ey + a3 * (ir + aj5) * sd);
        u[5] = f * (q * fy - a3 * (qy - aj6) * sd);
        u[6] = f * (-xi * fz - y * x11 + a3 * ak1 * sd);
        u[7] = f * (-ez + a3 * y * d11 * sd);
        u[8] = f * (q * fz + a3 * ak2 * sd);
    } else {
        u[0] = f * (sdiv(xi * q, R.r3) + a3 * aj4 * sdcd);
        u[1] = f * (sdiv(et * q, R.r3) + qy + a3 * aj5 * sdcd);
        u[2] = f * (sdiv(-q2, R.r3) + a3 * aj6 * sdcd);
        u[3] = f * (-ey + a3 * aj1 * sdcd);
        u[4] = f * (-et * gy - xy * sd + a3 * aj2 * sdcd);
        u[5] = f * (q * gy + a3 * aj3 * sdcd);
        u[6] = f * (-ez - a3 * ak3 * sdcd);
        u[7] = f * (-et * gz - xy * cd - a3 * xi * d11 * sdcd);
        u[8] = f * (q * gz - a3 * ak4 * sdcd);
    }
}

// rows 0..2 (needed for du/dz, DC3D adds them to the z-derivative block) and rows 3..11 of UC
template <int SLIP>
__device__ __forceinline__ void strict_uc(const OkadaMedium& m, const StrictGeo& g, const StrictRecips& R, double z,
                                          double (&u0)[3], double (&u)[9])
{
    const double xi = g.xi, et = g.et, q = g.q, xi2 = g.xi2, et2 = g.et2, q2 = g.q2;
    const double r = g.r, r2 = g.r2, y = g.y, d = g.d;
    const double x11 = g.x11, y11 = g.y11, x32 = g.x32, y32 = g.y32;
    const double a4 = m.a4, a5 = m.a5, sd = m.sd, cd = m.cd, sdsd = m.sdsd, sdcd = m.sdcd;
    const double c = d + z;
    const double x53 = sdiv((8.0 * r2 + 9.0 * r * xi + 3.0 * xi2) * x11 * x11 * x11, R.r2);
    const double y53 = sdiv((8.0 * r2 + 9.0 * r * et + 3.0 * et2) * y11 * y11 * y11, R.r2);
    const double h = q * cd - z;
    const double sd_r3 = sdiv(sd, R.r3), cd_r3 = sdiv(cd, R.r3), y_r3 = sdiv(y, R.r3), d_r3 = sdiv(d, R.r3);
    const double z32 = sd_r3 - h * y32;
    const double z53 = sdiv(3.0 * sd, R.r5) - h * y53;
    const double y0 = y11 - xi2 * y32;
    const double z0 = z32 - xi2 * z53;
    const double ppy = cd_r3 + q * y32 * sd;
    const double ppz = sd_r3 - q * y32 * cd;
    const double qq = z * y32 + z32 + z0;
    const double qqy = sdiv(3.0 * c * d, R.r5) - qq * sd;
    const double qqz = sdiv(3.0 * c * y, R.r5) - qq * cd + q * y32;
    const double xy = xi * y11, qy = q * y11;
    const double qr = sdiv(3.0 * q, R.r5);
    const double cdr = sdiv(c + d, R.r3);
    const double yy0 = y_r3 - y0 * cd;
    const double f = 1.0 / kOkadaPi2;
    if (SLIP == kStrikeSlip) {
        u0[0] = f * (a4 * xy * cd - a5 * xi * q * z32);
        u0[1] = f * (a4 * (sdiv(cd, R.r) + 2.0 * qy * sd) - sdiv(a5 * c * q, R.r3));
        u0[2] = f * (a4 * qy * cd - a5 * (sdiv(c * et, R.r3) - z * y11 + xi2 * z32));
        u[0] = f * (a4 * y0 * cd - a5 * q * z0);
        u[1] = f * (-a4 * xi * (cd_r3 + 2.0 * q * y32 * sd) + a5 * c * xi * qr);
        u[2] = f * (-a4 * xi * q * y32 * cd + a5 * xi * (sdiv(3.0 * c * et, R.r5) - qq));
        u[3] = f * (-a4 * xi * ppy * cd - a5 * xi * qqy);
        u[4] = f * (a4 * 2.0 * (d_r3 - y0 * sd) * sd - y_r3 * cd - a5 * (cdr * sd - sdiv(et, R.r3) - c * y * qr));
        u[5] = f * (sdiv(-a4 * q, R.r3) + yy0 * sd + a5 * (cdr * cd + c * d * qr - (y0 * cd + q * z0) * sd));
        u[6] = f * (a4 * xi * ppz * cd - a5 * xi * qqz);
        u[7] = f * (a4 * 2.0 * (y_r3 - y0 * cd) * sd + d_r3 * cd - a5 * (cdr * cd + c * d * qr));
        u[8] = f * (yy0 * cd - a5 * (cdr * sd - c * y * qr - y0 * sdsd + q * z0 * cd));
    } else {
        u0[0] = f * (sdiv(a4 * cd, R.r) - qy * sd - sdiv(a5 * c * q, R.r3));
        u0[1] = f * (a4 * y * x11 - a5 * c * et * q * x32);
        u0[2] = f * (-d * x11 - xy * sd - a5 * c * (x11 - q2 * x32));
        u[0] = f * (sdiv(-a4 * xi, R.r3) * cd + a5 * c * xi * qr + xi * q * y32 * sd);
        u[1] = f * (sdiv(-a4 * y, R.r3) + a5 * c * et * qr);
        u[2] = f * (d_r3 - y0 * sd + sdiv(a5 * c, R.r3) * (1.0 - sdiv(3.0 * q2, R.r2)));
        u[3] = f * (sdiv(-a4 * et, R.r3) + y0 * sdsd - a5 * (cdr * sd - c * y * qr));
        u[4] = f * (a4 * (x11 - y * y * x32) - a5 * c * ((d + 2.0 * q * cd) * x32 - y * et * q * x53));
        u[5] = f * (xi * ppy * sd + y * d * x32 + a5 * c * ((y + 2.0 * q * sd) * x32 - y * q2 * x53));
        u[6] = f * (sdiv(-q, R.r3) + y0 * sdcd - a5 * (cdr * cd + c * d * qr));
        u[7] = f * (a4 * y * d * x32 - a5 * c * ((y - 2.0 * q * sd) * x32 + d * et * q * x53));
        u[8] = f * (-xi * ppz * sd + x11 - d * d * x32 - a5 * c * ((d - 2.0 * q * cd) * x32 - d * q2 * x53));
    }
}

struct StrictSetup {
    double xi[2], et[2], q;
    bool kxi[2], ket[2], singular;
};

__device__ __forceinline__ void strict_setup(const OkadaMedium& m, double x, double y, double dd, double al1,
                                             double al2, double aw1, double aw2, StrictSetup& s)
{
    const double sd = m.sd, cd = m.cd;
    s.xi[0] = x - al1; s.xi[1] = x - al2;
#pragma unroll
    for (int k = 0; k < 2; ++k) if (fabs(s.xi[k]) < kOkadaEps) s.xi[k] = 0.0;
    const double p = y * cd + dd * sd;
    double q = y * sd - dd * cd;
    s.et[0] = p - aw1; s.et[1] = p - aw2;
    if (fabs(q) < kOkadaEps) q = 0.0;
#pragma unroll
    for (int k = 0; k < 2; ++k) if (fabs(s.et[k]) < kOkadaEps) s.et[k] = 0.0;
    s.q = q;
    s.singular = (q == 0.0) && ((s.xi[0] * s.xi[1] <= 0.0 && s.et[0] * s.et[1] == 0.0) ||
                                (s.et[0] * s.et[1] <= 0.0 && s.xi[0] * s.xi[1] == 0.0));
    const double r12 = sqrt(s.xi[0] * s.xi[0] + s.et[1] * s.et[1] + q * q);
    const double r21 = sqrt(s.xi[1] * s.xi[1] + s.et[0] * s.et[0] + q * q);
    const double r22 = sqrt(s.xi[1] * s.xi[1] + s.et[1] * s.et[1] + q * q);
    s.kxi[0] = s.xi[0] < 0.0 && r21 + s.xi[1] < kOkadaEps;
    s.kxi[1] = s.xi[0] < 0.0 && r22 + s.xi[1] < kOkadaEps;
    s.ket[0] = s.et[0] < 0.0 && r12 + s.et[1] < kOkadaEps;
    s.ket[1] = s.et[0] < 0.0 && r22 + s.et[1] < kOkadaEps;
}

// One dc3d call, gradient rows only: g[0..8] = entries 4..12 of dc3d's 12-vector (1/(2π) included), ADDED to
// the caller's running image sum exactly as GF.jl:55 does (`u .+= cache[1]`).  A receiver above the surface or
// on a fault edge contributes zeros (IRET = 2 / 1).
template <int SLIP>
__device__ __forceinline__ void okada_gradient_strict(const OkadaMedium& m, double x, double y, double z, double dep,
                                                      double al1, double al2, double aw1, double aw2, double (&g)[9])
{
    if (z > 0.0) return;
    const double sd = m.sd, cd = m.cd;
    double acc[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) acc[i] = 0.0;
    StrictSetup s;
    StrictGeo geo;
    strict_setup(m, x, y, dep + z, al1, al2, aw1, aw2, s);                    // real source
    if (s.singular) return;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            strict_corner_base(m, s.xi[j], s.et[k], s.q, geo);
            const StrictRecips R(geo);
            strict_corner(m, s.kxi[k], s.ket[j], R, geo);
            double A[9], du[9];
            strict_ua<SLIP>(m, geo, R, A);
#pragma unroll
            for (int i = 0; i < 9; i += 3) {
                du[i] = -A[i];
                du[i + 1] = -A[i + 1] * cd + A[i + 2] * sd;
                du[i + 2] = -A[i + 1] * sd - A[i + 2] * cd;
                if (i == 6) { du[6] = -du[6]; du[7] = -du[7]; du[8] = -du[8]; }
            }
            const double sgn = (j + k == 1) ? -1.0 : 1.0;
#pragma unroll
            for (int i = 0; i < 9; ++i) acc[i] += sgn * du[i];
        }
    }
    strict_setup(m, x, y, dep - z, al1, al2, aw1, aw2, s);                    // image source
    if (s.singular) return;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            strict_corner_base(m, s.xi[j], s.et[k], s.q, geo);
            const StrictRecips R(geo);
            strict_corner(m, s.kxi[k], s.ket[j], R, geo);
            double A[9], B[9], C0[3], C[9], du[9];
            strict_ua<SLIP>(m, geo, R, A);
            strict_ub<SLIP>(m, geo, R, B);
            strict_uc<SLIP>(m, geo, R, z, C0, C);
#pragma unroll
            for (int i = 0; i < 9; i += 3) {
                du[i] = A[i] + B[i] + z * C[i];
                du[i + 1] = (A[i + 1] + B[i + 1] + z * C[i + 1]) * cd - (A[i + 2] + B[i + 2] + z * C[i + 2]) * sd;
                du[i + 2] = (A[i + 1] + B[i + 1] - z * C[i + 1]) * sd + (A[i + 2] + B[i + 2] - z * C[i + 2]) * cd;
                if (i == 6) {
                    du[6] += C0[0];
                    du[7] += C0[1] * cd - C0[2] * sd;
                    du[8] -= C0[1] * sd + C0[2] * cd;
                }
            }
            const double sgn = (j + k == 1) ? -1.0 : 1.0;
#pragma unroll
            for (int i = 0; i < 9; ++i) acc[i] += sgn * du[i];
        }
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) g[i] += acc[i];
}

}  // namespace oq
