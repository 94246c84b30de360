// greens_okada.cuh -- device views and the two Okada assembly kernels (K1 fault->fault, K2 fault->mantle), shared
// by greens.cu (fast form, FMA-contracted) and greens_strict.cu (published operation order, --fmad=false).
#pragma once
#include "common.cuh"
#include "okada_dev.cuh"
#include "okada_strict.cuh"

namespace oq {

// ---- device views ---------------------------------------------------------------------------
struct FaultGeom {
    const double *x, *ax0, *ax1;            // [nx]
    const double *y, *z, *axi0, *axi1;      // [nxi]
    int nx, nxi;
    double dep;
};

struct Hex8Geom {
    const double *cx, *cy, *cz, *qx, *qy, *qz, *dx, *dy, *dz;
    int n;
};

struct OkadaParams {
    OkadaMedium m;
    double lam, mu;
    double s1, c1, s2, c2;   // sind(dip), cosd(dip), sind(2dip), cosd(2dip) for the traction projection
    double lrept;            // image period (GF.jl:38,135)
    int nrept;
};

// one periodic image of the source added to the running gradient sum (GF.jl:47-56, :154-162): the published
// operation order (okada_strict.cuh, bit-reproducible; needs a translation unit compiled with --fmad=false) or
// the restructured fast form (okada_dev.cuh, 1/(2π) applied once after the image sum by okada_finish)
template <int SLIP, bool STRICT>
__device__ __forceinline__ void okada_image(const OkadaMedium& m, double x, double y, double z, double dep,
                                            double al1, double al2, double aw1, double aw2, double (&g)[9])
{
    if (STRICT) okada_gradient_strict<SLIP>(m, x, y, z, dep, al1, al2, aw1, aw2, g);
    else okada_gradient<SLIP>(m, x, y, z, dep, al1, al2, aw1, aw2, g);
}

template <bool STRICT>
__device__ __forceinline__ void okada_finish(double (&g)[9])
{
    if (!STRICT) {
#pragma unroll
        for (int k = 0; k < 9; ++k) g[k] *= kInv2Pi;
    }
}

// GF.jl:76-87 on the gradient 9-vector g = u[4..12]
template <int SLIP>
__device__ __forceinline__ double shear_traction_grad(const double (&g)[9], const OkadaParams& p)
{
    if (SLIP == kStrikeSlip) {
        const double sxy = p.mu * (g[1] + g[3]);
        const double sxz = p.mu * (g[2] + g[6]);
        return -sxy * p.s1 + sxz * p.c1;
    } else {
        const double l2m = p.lam + 2.0 * p.mu;
        const double szz = l2m * g[8] + p.lam * g[0] + p.lam * g[4];
        const double syy = l2m * g[4] + p.lam * g[0] + p.lam * g[8];
        const double syz = p.mu * (g[7] + g[5]);
        return (szz - syy) / 2.0 * p.s2 + syz * p.c2;
    }
}

// GF.jl:89-96
__device__ __forceinline__ double shear_traction_stress(int slip, const double (&s)[6], double s1, double c1,
                                                        double s2, double c2)
{
    return slip == kStrikeSlip ? (-s[1] * s1 + s[2] * c1) : ((s[5] - s[3]) / 2.0 * s2 + s[4] * c2);
}

// ---- K1: fault -> fault, Toeplitz-unique entries st[i,j,l] (GF.jl:41-58) -------------------------
// thread t -> (i, j, l) with i fastest: writes are coalesced, the receiver depth (j) and the source
// row (l) are warp-uniform for nx >= 32 so the EPS / edge branches of the closed form do not diverge.
#ifndef OQ_OKADA_STRICT_MINB
#define OQ_OKADA_STRICT_MINB 3   // the published operation order keeps more values live (168-register cap)
#endif
#ifndef OQ_OKADA_MINB
#define OQ_OKADA_MINB 4      // resident CTAs per SM the Okada kernels are compiled for (caps registers at 128; measured fastest)
#endif
#ifndef OQ_OKADA_STRICT_MINB1
#define OQ_OKADA_STRICT_MINB1 4
#endif
template <int SLIP, bool STRICT>
__global__ void __launch_bounds__(128, STRICT ? OQ_OKADA_STRICT_MINB1 : OQ_OKADA_MINB)
gf_fault_fault_kernel(FaultGeom f, OkadaParams p, double* __restrict__ st)
{
    extern __shared__ double sm[];
    double* sy = sm;
    double* sz = sy + f.nxi;
    double* sa0 = sz + f.nxi;
    double* sa1 = sa0 + f.nxi;
    for (int k = threadIdx.x; k < f.nxi; k += blockDim.x) {
        sy[k] = f.y[k]; sz[k] = f.z[k]; sa0[k] = f.axi0[k]; sa1[k] = f.axi1[k];
    }
    __syncthreads();
    const size_t total = (size_t)f.nx * f.nxi * f.nxi;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int i = (int)(t % f.nx);
    const int j = (int)((t / f.nx) % f.nxi);
    const int l = (int)(t / ((size_t)f.nx * f.nxi));
    const double x = f.x[i], y = sy[j], z = sz[j];
    const double al1 = f.ax0[0], al2 = f.ax1[0], aw1 = sa0[l], aw2 = sa1[l];
    double g[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) g[k] = 0.0;
    for (int r = -p.nrept; r <= p.nrept; ++r) {
        const double jump = r * p.lrept;
        okada_image<SLIP, STRICT>(p.m, x, y, z, f.dep, al1 + jump, al2 + jump, aw1, aw2, g);
    }
    okada_finish<STRICT>(g);
    st[t] = shear_traction_grad<SLIP>(g, p);
}

// ---- K2: fault -> mantle (GF.jl:123-174) -----------------------------------------------------------
// thread t -> (source fault cell j fastest, receiver element e); writes row-major G[(k*nel+el), j].
template <int SLIP, bool STRICT>
__global__ void __launch_bounds__(128, STRICT ? OQ_OKADA_STRICT_MINB : OQ_OKADA_MINB)
gf_fault_mantle_kernel(FaultGeom f, Hex8Geom a, OkadaParams p, const double* __restrict__ qc,
                       const double* __restrict__ qw, int nq, int e_begin, int nel, size_t ld,
                       double* __restrict__ G)
{
    const int nf = f.nx * f.nxi;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)nf * nel) return;
    const int j = (int)(t % nf);
    const int el = (int)(t / nf);
    const int e = e_begin + el;
    const int q1 = j % f.nx, q2 = j / f.nx;
    const double al1 = f.ax0[q1], al2 = f.ax1[q1], aw1 = f.axi0[q2], aw2 = f.axi1[q2];
    const double cx = a.cx[e], cy = a.cy[e], cz = a.cz[e];
    const double hx = a.dx[e] / 2, hy = a.dy[e] / 2, hz = a.dz[e] / 2;
    double s[6] = {0, 0, 0, 0, 0, 0};
    for (int w = 0; w < nq; ++w) {
        // strike coordinates: product and sum rounded separately in BOTH builds (no FMA contraction), so that the
        // host can form the very same differences x - al when it sorts the pairs into classes (greens_classes.cuh)
        const double rx = __dadd_rn(cx, __dmul_rn(qc[3 * w], hx));
        const double ry = cy + qc[3 * w + 1] * hy;
        const double rz = cz + qc[3 * w + 2] * hz;
        double g[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) g[k] = 0.0;
        for (int r = -p.nrept; r <= p.nrept; ++r) {
            const double jump = __dmul_rn((double)r, p.lrept);
            okada_image<SLIP, STRICT>(p.m, rx, ry, rz, f.dep, __dadd_rn(al1, jump), __dadd_rn(al2, jump), aw1, aw2, g);
        }
        okada_finish<STRICT>(g);
        const double lekk = p.lam * (g[0] + g[4] + g[8]);
        const double wt = qw[w];
        s[0] += wt * (lekk + 2.0 * p.mu * g[0]);
        s[1] += wt * (p.mu * (g[1] + g[3]));
        s[2] += wt * (p.mu * (g[2] + g[6]));
        s[3] += wt * (lekk + 2.0 * p.mu * g[4]);
        s[4] += wt * (p.mu * (g[5] + g[7]));
        s[5] += wt * (lekk + 2.0 * p.mu * g[8]);
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) G[((size_t)k * nel + el) * ld + j] = s[k];
}


// ---- K2'': fault -> mantle on CLASSES of pairs (greens_classes.cuh) ----------------------------------------------
// dc3d reads the strike coordinate of receiver and source only through x - al1, x - al2 (DC3D: XI(1) = X - AL1,
// XI(2) = X - AL2), for every periodic image.  Pairs whose differences are BITWISE equal therefore have bitwise
// equal entries: one evaluation per class on the coordinates of a representative pair, the same code as K2, and
// the table entry is bit-identical to what K2 writes for every pair of the class.  T[k][u23][u1].
template <int SLIP, bool STRICT>
__global__ void __launch_bounds__(128, STRICT ? OQ_OKADA_STRICT_MINB : OQ_OKADA_MINB)
gf_fault_mantle_class_kernel(FaultGeom f, Hex8Geom a, OkadaParams p, const double* __restrict__ qc,
                             const double* __restrict__ qw, int nq, const int* __restrict__ rep_r1,
                             const int* __restrict__ rep_s1, const int* __restrict__ rep_r23,
                             const int* __restrict__ rep_s23, int n1, int n23, double* __restrict__ T)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)n1 * n23) return;
    const int u1 = (int)(t % n1), u23 = (int)(t / n1);
    const int e1 = rep_r1[u1], e23 = rep_r23[u23];
    const int q1 = rep_s1[u1] % f.nx, q2 = rep_s23[u23] / f.nx;
    const double al1 = f.ax0[q1], al2 = f.ax1[q1], aw1 = f.axi0[q2], aw2 = f.axi1[q2];
    const double cx = a.cx[e1], cy = a.cy[e23], cz = a.cz[e23];
    const double hx = a.dx[e1] / 2, hy = a.dy[e23] / 2, hz = a.dz[e23] / 2;
    double s[6] = {0, 0, 0, 0, 0, 0};
    for (int w = 0; w < nq; ++w) {
        // strike coordinates: product and sum rounded separately in BOTH builds (no FMA contraction), so that the
        // host can form the very same differences x - al when it sorts the pairs into classes (greens_classes.cuh)
        const double rx = __dadd_rn(cx, __dmul_rn(qc[3 * w], hx));
        const double ry = cy + qc[3 * w + 1] * hy;
        const double rz = cz + qc[3 * w + 2] * hz;
        double g[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) g[k] = 0.0;
        for (int r = -p.nrept; r <= p.nrept; ++r) {
            const double jump = __dmul_rn((double)r, p.lrept);
            okada_image<SLIP, STRICT>(p.m, rx, ry, rz, f.dep, __dadd_rn(al1, jump), __dadd_rn(al2, jump), aw1, aw2, g);
        }
        okada_finish<STRICT>(g);
        const double lekk = p.lam * (g[0] + g[4] + g[8]);
        const double wt = qw[w];
        s[0] += wt * (lekk + 2.0 * p.mu * g[0]);
        s[1] += wt * (p.mu * (g[1] + g[3]));
        s[2] += wt * (p.mu * (g[2] + g[6]));
        s[3] += wt * (lekk + 2.0 * p.mu * g[4]);
        s[4] += wt * (p.mu * (g[5] + g[7]));
        s[5] += wt * (lekk + 2.0 * p.mu * g[8]);
    }
    const size_t stride = (size_t)n1 * n23;
#pragma unroll
    for (int k = 0; k < 6; ++k) T[(size_t)k * stride + (size_t)u23 * n1 + u1] = s[k];
}

struct OkadaClassLaunch {
    const double *qc, *qw;
    int nq;
    const int *rep_r1, *rep_s1, *rep_r23, *rep_s23;
    int n1, n23;
    double* T;
};

// launchers of the bit-reproducible instantiations (greens_strict.cu)
void launch_fault_fault_strict(int ftype, unsigned blocks, size_t smem, const FaultGeom& f, const OkadaParams& p, double* st);
void launch_fault_mantle_strict(int ftype, unsigned blocks, const FaultGeom& f, const Hex8Geom& a, const OkadaParams& p,
                                const double* qc, const double* qw, int nq, int e_begin, int nel, size_t ld, double* G);
void launch_fault_mantle_class_strict(int ftype, const FaultGeom& f, const Hex8Geom& a, const OkadaParams& p,
                                      const OkadaClassLaunch& c);
void launch_dc3d_gradient_strict(int ftype, int n, const double* x, const double* y, const double* z, const OkadaMedium& m,
                                 double dep, double al1, double al2, double aw1, double aw2, double* out);

}  // namespace oq
