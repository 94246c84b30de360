// greens_strict.cu -- the Okada assembly kernels in the published operation order (okada_strict.cuh).
//
// This translation unit is compiled with --fmad=false: no multiply-add contraction, so every +, -, *, / and sqrt
// rounds exactly as in a scalar IEEE evaluation of DC3D -- the results are bit-identical to the CPU restatement
// of /root/reference/src/BEM/GF.jl:31-58 and :123-174 (tests/test_gpu_greens.py compares with ==).
#include "greens_okada.cuh"

namespace oq {

template <int SLIP>
__global__ void dc3d_gradient_strict_kernel(int n, const double* x, const double* y, const double* z, OkadaMedium m,
                                            double dep, double al1, double al2, double aw1, double aw2, double* out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    double g[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) g[k] = 0.0;
    okada_gradient_strict<SLIP>(m, x[t], y[t], z[t], dep, al1, al2, aw1, aw2, g);
#pragma unroll
    for (int k = 0; k < 9; ++k) out[(size_t)t * 9 + k] = g[k];
}

void launch_fault_fault_strict(int ftype, unsigned blocks, size_t smem, const FaultGeom& f, const OkadaParams& p, double* st)
{
    if (ftype == OQ_STRIKE_SLIP) gf_fault_fault_kernel<kStrikeSlip, true><<<blocks, 128, smem>>>(f, p, st);
    else gf_fault_fault_kernel<kDipSlip, true><<<blocks, 128, smem>>>(f, p, st);
}

void launch_fault_mantle_strict(int ftype, unsigned blocks, const FaultGeom& f, const Hex8Geom& a, const OkadaParams& p,
                                const double* qc, const double* qw, int nq, int e_begin, int nel, size_t ld, double* G)
{
    if (ftype == OQ_STRIKE_SLIP)
        gf_fault_mantle_kernel<kStrikeSlip, true><<<blocks, 128>>>(f, a, p, qc, qw, nq, e_begin, nel, ld, G);
    else
        gf_fault_mantle_kernel<kDipSlip, true><<<blocks, 128>>>(f, a, p, qc, qw, nq, e_begin, nel, ld, G);
}

void launch_fault_mantle_class_strict(int ftype, const FaultGeom& f, const Hex8Geom& a, const OkadaParams& p,
                                      const OkadaClassLaunch& c)
{
    const unsigned blocks = (unsigned)(((size_t)c.n1 * c.n23 + 127) / 128);
    if (ftype == OQ_STRIKE_SLIP)
        gf_fault_mantle_class_kernel<kStrikeSlip, true><<<blocks, 128>>>(f, a, p, c.qc, c.qw, c.nq, c.rep_r1, c.rep_s1, c.rep_r23,
                                                                         c.rep_s23, c.n1, c.n23, c.T);
    else
        gf_fault_mantle_class_kernel<kDipSlip, true><<<blocks, 128>>>(f, a, p, c.qc, c.qw, c.nq, c.rep_r1, c.rep_s1, c.rep_r23,
                                                                      c.rep_s23, c.n1, c.n23, c.T);
}

void launch_dc3d_gradient_strict(int ftype, int n, const double* x, const double* y, const double* z, const OkadaMedium& m,
                                 double dep, double al1, double al2, double aw1, double aw2, double* out)
{
    const int blocks = (n + 127) / 128;
    if (ftype == OQ_STRIKE_SLIP) dc3d_gradient_strict_kernel<kStrikeSlip><<<blocks, 128>>>(n, x, y, z, m, dep, al1, al2, aw1, aw2, out);
    else dc3d_gradient_strict_kernel<kDipSlip><<<blocks, 128>>>(n, x, y, z, m, dep, al1, al2, aw1, aw2, out);
}

}  // namespace oq
