// toeplitz_fft.cuh -- the reference's own algorithm for the fault-fault interaction, on the device.
//
// /root/reference/src/BEM/equation.jl:44-61 evaluates dτ/dt[i,j] = Σ_l Σ_k st[|i-k|,j,l] (v-vpl)[k,l] as a linear
// convolution along strike through FFTs: rfft of the zero-padded forcing, a per-frequency nξ x nξ contraction
// with the transformed kernel (GF.jl:60-68), inverse rfft, first nx rows.  The same three steps here:
//   fft_forward_kernel      one CTA per source row l: shared-memory Stockham FFT of length N
//   spectral_contract_kernel  T[f,j] = Σ_l Ĝ[f,j,l] R[f,l]   (Ĝ is REAL: the kernel's extension is even)
//   fft_inverse_kernel      one CTA per receiver row j: inverse FFT, first nx samples (+ optional fused
//                           rate-and-state epilogue when no dense operand follows)
// N is the power of two >= 2nx-1 (the reference uses exactly 2nx-1; any N >= 2nx-1 yields the same linear
// convolution).  The spectrum Ĝ[l][j][f] (f fastest, (N/2+1) doubles) is built once per problem.
#pragma once

namespace oq {

struct cplx { double re, im; };

// twiddle table W[j] = exp(-2 pi i j / N), j < N/2, built once per problem with sincospi (exact argument
// reduction); the transforms read it through shared memory
__global__ void __launch_bounds__(256) twiddle_kernel(int N, cplx* __restrict__ W)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N / 2) return;
    double s, c;
    sincospi(-2.0 * (double)j / (double)N, &s, &c);
    W[j] = {c, s};
}

// in-place-by-ping-pong Stockham radix-2 FFT of length N (power of two) in shared memory; returns the buffer
// holding the result.  All threads of the CTA participate; tw[] holds the N/2 twiddles in shared memory.
__device__ __forceinline__ cplx* stockham_fft(cplx* a, cplx* b, const cplx* tw, int N)
{
    const int half = N >> 1;
    int shift = 0;
    while ((1 << shift) < half) ++shift;                 // log2(N/2)
    for (int Ns = 1, ls = 0; Ns < N; Ns <<= 1, ++ls) {
        for (int j = threadIdx.x; j < half; j += blockDim.x) {
            const int k = j & (Ns - 1);
            const cplx w = tw[k << (shift - ls)];          // exp(-i pi k / Ns) = W[k * (N/2) / Ns]
            const cplx u0 = a[j], v = a[j + half];
            const cplx u1 = {v.re * w.re - v.im * w.im, v.re * w.im + v.im * w.re};
            const int j0 = ((j - k) << 1) + k;
            b[j0] = {u0.re + u1.re, u0.im + u1.im};
            b[j0 + Ns] = {u0.re - u1.re, u0.im - u1.im};
        }
        __syncthreads();
        cplx* t = a; a = b; b = t;
    }
    return a;
}

// Ĝ[l][j][f] = st[0,j,l] + 2 Σ_{m=1}^{nx-1} st[m,j,l] cos(2π f m / N),  f = 0..N/2
__global__ void __launch_bounds__(256)
toeplitz_spectrum_kernel(const double* __restrict__ st, int nx, int nxi, int N, int j0, int nj, double* __restrict__ Gh)
{
    const int nfreq = N / 2 + 1;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)nfreq * nj * nxi) return;
    const int f = (int)(t % nfreq);
    const int jl = (int)((t / nfreq) % nj);
    const int l = (int)(t / ((size_t)nfreq * nj));
    const double* s = st + (size_t)nx * ((j0 + jl) + (size_t)nxi * l);
    double acc = 0.0;
    for (int m = nx - 1; m >= 1; --m) {
        const int fm = (int)(((long long)f * m) % N);
        acc = fma(s[m], cospi(2.0 * (double)fm / (double)N), acc);
    }
    Gh[t] = s[0] + 2.0 * acc;
}

// R[l][f] = FFT_N(zero-padded relv[:, l])[f],  f = 0..N/2
// `v_direct` != nullptr (single rank, no dense operand needs the forcing vector): the forcing v - vpl is formed
// here from the state and the separate forcing kernel is skipped.
__global__ void __launch_bounds__(256)
fft_forward_kernel(const double* relv0, size_t relv_stride, PeerWait pw, const double* __restrict__ v_direct,
                   double vpl, int nx, int N, const cplx* __restrict__ W, cplx* __restrict__ Rh)
{
    extern __shared__ __align__(16) unsigned char fsm[];
    cplx* a = reinterpret_cast<cplx*>(fsm);
    cplx* b = a + N;
    cplx* tw = b + N;
    for (int k = threadIdx.x; k < N / 2; k += blockDim.x) tw[k] = W[k];
    const int l = blockIdx.x;
    if (v_direct) {
        for (int k = threadIdx.x; k < N; k += blockDim.x)
            a[k] = {k < nx ? v_direct[k + (size_t)nx * l] - vpl : 0.0, 0.0};
    } else {
        const double* relv = relv0 + consumer_parity(pw) * relv_stride;
        for (int k = threadIdx.x; k < N; k += blockDim.x) a[k] = {k < nx ? relv[k + (size_t)nx * l] : 0.0, 0.0};
    }
    __syncthreads();
    const cplx* r = stockham_fft(a, b, tw, N);
    const int nfreq = N / 2 + 1;
    for (int f = threadIdx.x; f < nfreq; f += blockDim.x) Rh[(size_t)l * nfreq + f] = r[f];
}

// T[jl][f] = Σ_l Ĝ[l][jl][f] R[l][f]
__global__ void __launch_bounds__(256)
spectral_contract_kernel(const double* __restrict__ Gh, const cplx* __restrict__ Rh, int nxi, int nj, int nfreq,
                         cplx* __restrict__ Th)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)nfreq * nj) return;
    const int f = (int)(t % nfreq);
    double re = 0.0, im = 0.0;
#pragma unroll 4
    for (int l = 0; l < nxi; ++l) {
        const double g = Gh[(size_t)l * nfreq * nj + t];
        const cplx r = Rh[(size_t)l * nfreq + f];
        re = fma(g, r.re, re);
        im = fma(g, r.im, im);
    }
    Th[t] = {re, im};
}

// dτ[i, j0+jl] = real(IFFT_N(Hermitian extension of T[jl][:]))[i], i < nx; rows outside [f0, f0+nfl) are dropped.
// The per-frequency contraction T[jl][f] = Σ_l Ĝ[l][jl][f] R[l][f] is done here by the CTA that transforms the
// row (Ĝ is read exactly once per evaluation; R comes from L2): the source index l is split over kFftLGroups
// thread groups so that many independent loads are in flight, partial sums are folded in a fixed order.
constexpr int kFftLGroups = 4;
constexpr int kFftInvThreads = 1024;

__global__ void __launch_bounds__(kFftInvThreads)
fft_inverse_kernel(const double* __restrict__ Gh, const cplx* __restrict__ Rh, int nxi, int nj, int nx, int N, int j0,
                   int f0, int nfl, const cplx* __restrict__ W, double* __restrict__ dtau, int fuse_epilogue,
                   FaultEpilogue fe)
{
    extern __shared__ __align__(16) unsigned char fsm[];
    cplx* a = reinterpret_cast<cplx*>(fsm);
    cplx* b = a + N;
    cplx* tw = b + N;
    cplx* part = tw + N / 2;                       // [kFftLGroups][nfreq]
    const int jl = blockIdx.x;
    const int nfreq = N / 2 + 1;
    for (int k = threadIdx.x; k < N / 2; k += blockDim.x) tw[k] = W[k];
    const int per = blockDim.x / kFftLGroups;      // threads per l-group
    const int grp = threadIdx.x / per, tf = threadIdx.x % per;
    const int lchunk = (nxi + kFftLGroups - 1) / kFftLGroups;
    const int l0 = grp * lchunk, l1 = min(nxi, l0 + lchunk);
    for (int f = tf; f < nfreq; f += per) {
        double re = 0.0, im = 0.0;
#pragma unroll 8
        for (int l = l0; l < l1; ++l) {
            const double g = Gh[((size_t)l * nj + jl) * nfreq + f];
            const cplx r = Rh[(size_t)l * nfreq + f];
            re = fma(g, r.re, re);
            im = fma(g, r.im, im);
        }
        part[grp * nfreq + f] = {re, im};
    }
    __syncthreads();
    for (int f = threadIdx.x; f < nfreq; f += blockDim.x) {
        double re = 0.0, im = 0.0;
#pragma unroll
        for (int g = 0; g < kFftLGroups; ++g) { re += part[g * nfreq + f].re; im += part[g * nfreq + f].im; }
        // ifft(X) = conj(fft(conj(X))) / N; X[N-f] = conj(X[f])
        a[f] = {re, -im};
        if (f > 0 && f < N - f) a[N - f] = {re, im};
    }
    __syncthreads();
    const cplx* r = stockham_fft(a, b, tw, N);
    const double inv = 1.0 / (double)N;
    for (int i = threadIdx.x; i < nx; i += blockDim.x) {
        const int row = i + nx * (j0 + jl) - f0;
        if (row >= 0 && row < nfl) {
            const double v = r[i].re * inv;
            if (fuse_epilogue) update_fault_row(fe, row, v);
            else dtau[row] = v;
        }
    }
}

}  // namespace oq
