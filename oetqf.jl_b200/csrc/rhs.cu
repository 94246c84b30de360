// rhs.cu -- the per-step ODE right-hand side (hot path 2).
//
// Replaces ode(du,u,p,t) of /root/reference/src/BEM/equation.jl:156-205 and the kernels it calls:
//   relative_velocity!      :35-42     -> forcing_kernel
//   update_strain_rate!     :207-222   -> forcing_kernel   (power-law / composite, :285-292)
//   relative_strain_rate!   :224-230   -> forcing_kernel
//   dτ_dt! (FFT conv)       :44-61     -> dense rows of G11 in matvec_fused_kernel, or toeplitz_conv_kernel
//   3x matvecmul!           :201-203   -> matvec_fused_kernel (both operands of a row in one pass)
//   update_fault!           :233-246   -> epilogue of matvec_fused_kernel
//   update_fault_with_dilatancy! :248-276 -> same epilogue, dilatancy branch
//
// The matvec is HBM-bound: fp64 matrices are streamed once with 128-bit no-allocate loads, the
// forcing-vector segment of each CTA is staged in shared memory by a 1-D bulk TMA copy, rows are
// reduced with warp shuffles, and the last CTA to finish a row block applies the pointwise physics.
#include "comm.cuh"
#include "problem.cuh"
#include "tma.cuh"

namespace oq {

constexpr int kMvThreads = 256;
constexpr int kMvRows = 4;            // rows per CTA item (x segment reused across them)
constexpr int kMvMaxSeg = 4096;       // columns per segment (32 KB of shared memory)
constexpr int kMvColStep = 2 * kMvThreads;

// ---- pointwise: forcing vectors, with the all-gather fused in -----------------------------------------
struct ForcingArgs {
    const double* v;       // [nfl]
    const double* sig;     // [6*nel] or null
    double* deps_out;      // du.eps [6*nel] or null
    StageSpec stage;       // optional fused Runge-Kutta stage combination
    double* y;             // stage state to fill (base of the state-shaped buffer) when stage.nk > 0
    size_t off_fault[4];   // offsets of v, θ, δ, 𝓅 inside a state-shaped buffer
    size_t off_eps, off_sig;
    int n_fault_parts;     // 3, or 4 with dilatancy
    PeerTargets peers;     // every rank's window (own included)
    WindowLayout wl;
    unsigned long long* epochs;   // local counters
    int nfl, f0, nel, e0, ne;
    double vpl;
    MantleParams mp;
};

__global__ void __launch_bounds__(1024) forcing_kernel(const __grid_constant__ ForcingArgs a)
{
    // steps enqueued past the end of the integration do nothing -- identically on every rank, so the epoch
    // protocol stays in lockstep
    if (a.stage.nk > 0 && a.stage.done && *a.stage.done) return;
    pdl_launch_dependents();       // the matvec may start (and prefetch its matrix chunks) while we compute
    // buffer copy for this evaluation: parity of the number of publications so far
    const unsigned long long ep = a.epochs[kEpForcing];
    const size_t par = (size_t)(ep & 1ull);
    const int world = a.peers.world;
    const int stride = gridDim.x * blockDim.x;
    const bool staged = a.stage.nk > 0;
    const double dt = staged ? (a.stage.dt ? *a.stage.dt : 1.0) : 0.0;
    const double* __restrict__ adev = a.stage.adev;
    auto combine = [&](size_t idx) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < 6; ++j)
            if (j < a.stage.nk) acc = fma(adev ? adev[j] : a.stage.a[j], a.stage.k[j][idx], acc);
        const double y = fma(dt, acc, a.stage.u[idx]);
        a.y[idx] = y;
        return y;
    };
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < a.nfl; t += stride) {
        double vt;
        if (staged) {
            vt = combine(a.off_fault[0] + t);
            for (int q = 1; q < a.n_fault_parts; ++q) combine(a.off_fault[q] + t);
        } else {
            vt = a.v[t];
        }
        const double rv = vt - a.vpl;                                        // equation.jl:38
        const size_t off = a.wl.off_relv + par * a.wl.relv_len + a.f0 + t;
        for (int r = 0; r < world; ++r) a.peers.base[r][off] = rv;           // local + NVLink peer stores
    }
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < a.nel; t += stride) {
        const size_t n = a.nel;
        double s1, s2, s3, s4, s5, s6;
        if (staged) {
#pragma unroll
            for (int k = 0; k < 6; ++k) combine(a.off_eps + t + k * n);
            s1 = combine(a.off_sig + t); s2 = combine(a.off_sig + t + n); s3 = combine(a.off_sig + t + 2 * n);
            s4 = combine(a.off_sig + t + 3 * n); s5 = combine(a.off_sig + t + 4 * n); s6 = combine(a.off_sig + t + 5 * n);
        } else {
            s1 = a.sig[t]; s2 = a.sig[t + n]; s3 = a.sig[t + 2 * n];
            s4 = a.sig[t + 3 * n]; s5 = a.sig[t + 4 * n]; s6 = a.sig[t + 5 * n];
        }
        const double skk = (s1 + s4 + s6) / 3;                               // equation.jl:209
        const double sxx = s1 - skk, syy = s4 - skk, szz = s6 - skk;
        const double tn = sqrt(sxx * sxx + syy * syy + szz * szz + 2 * (s2 * s2 + s3 * s3 + s5 * s5));
        const double comp[6] = {sxx, s2, s3, syy, s5, szz};
        double de[6] = {0, 0, 0, 0, 0, 0};
        for (int l = 0; l < a.mp.nlaws; ++l) {                               // equation.jl:285-292
            const double g = a.mp.gamma[(size_t)l * n + t];
            const double pw = pow(tn, a.mp.npow[(size_t)l * n + t]);
#pragma unroll
            for (int k = 0; k < 6; ++k) de[k] += g * comp[k] * pw;
        }
        const size_t base = a.wl.off_reldeps + par * a.wl.reldeps_len + a.e0 + t;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            a.deps_out[t + k * n] = de[k];
            const double rel = de[k] - a.mp.deps0[k];                        // equation.jl:227
            for (int r = 0; r < world; ++r) a.peers.base[r][base + (size_t)k * a.ne] = rel;
        }
    }
    // publication: the last block to finish bumps the local epoch and tells every peer
    // (one fence per block: the CTA barrier orders every thread's stores before thread 0's fence, which is
    // cumulative; a system-scope fence costs microseconds over NVLink, so it is not executed per thread)
    __shared__ int last;
    __syncthreads();
    if (threadIdx.x == 0) {
        if (gridDim.x == 1) {
            last = 1;                                  // single block: no cross-block handshake
            if (world > 1) __threadfence_system();
        } else {
            if (world > 1) __threadfence_system(); else __threadfence();
            const unsigned long long prev = atomicAdd(&a.epochs[kEpBlocksF], 1ull);
            last = (prev == (unsigned long long)gridDim.x - 1ull);
        }
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        if (gridDim.x > 1) a.epochs[kEpBlocksF] = 0ull;
        if (world > 1) {
            if (gridDim.x > 1) __threadfence_system();     // acquire the other blocks' publications
            for (int r = 0; r < world; ++r) {
                if (r == a.peers.rank) continue;
                unsigned long long* f = reinterpret_cast<unsigned long long*>(a.peers.base[r] + a.wl.off_flags);
                publish_flag(f + a.peers.rank, ep + 1ull);
            }
        }
        if (gridDim.x > 1) __threadfence();
        a.epochs[kEpForcing] = ep + 1ull;              // read by the next kernel in stream order
    }
}

// ---- pointwise: rate-and-state friction (equation.jl:233-246, :248-276, :279, :282) --------------
struct FaultEpilogue {
    FaultParams fp;
    const double *v, *theta, *pr;      // state in
    double *dv, *dtheta, *ddelta, *dpr;
    int dilatancy;
};

// state and properties of one fault row: loaded BEFORE the row's traction rate is known (they do not depend on
// it), so that at the end of a span only arithmetic separates the last partial sum from the stored derivative
struct FaultRowInputs {
    double v, th, a, b, L, sg;
};

__device__ __forceinline__ FaultRowInputs load_fault_row(const FaultEpilogue& e, int i)
{
    const FaultParams& p = e.fp;
    return FaultRowInputs{e.v[i], e.theta[i], p.a[i], p.b[i], p.L[i], p.sigma[i]};
}

__device__ __forceinline__ void update_fault_row(const FaultEpilogue& e, int i, double dtau, const FaultRowInputs& in);

__device__ __forceinline__ void update_fault_row(const FaultEpilogue& e, int i, double dtau)
{
    update_fault_row(e, i, dtau, load_fault_row(e, i));
}

__device__ __forceinline__ void update_fault_row(const FaultEpilogue& e, int i, double dtau, const FaultRowInputs& in)
{
    const FaultParams& p = e.fp;
    const double v = in.v, th = in.th, a = in.a, b = in.b, L = in.L, sg = in.sg;
    const double dth = 1.0 - v * th / L;                                     // aging law, :279
    if (!e.dilatancy) {
        const double psi1 = exp((p.f0 + b * log(p.v0 * fmax(0.0, th) / L)) / a) / (2.0 * p.v0);
        const double psi2 = sg * psi1 / hypot(1.0, v * psi1);
        const double dmu_dv = a * psi2;
        const double dmu_dth = b / th * v * psi2;
        e.dtheta[i] = dth;
        e.dv[i] = (dtau - dmu_dth * dth) / (dmu_dv + p.eta);
        e.ddelta[i] = v;
    } else {
        const double pr = e.pr[i];
        const double dpr = -(pr - p.p0[i]) / p.tp[i] + p.epsd[i] / p.beta[i] / th * dth;   // :282
        const double af = a / p.f0, bf = b / p.f0;
        const double vf = fmax(0.0, v / p.v0);
        const double tf = fmax(0.0, th * p.v0 / L);
        const double vfa1 = pow(vf, af - 1.0), tfb1 = pow(tf, bf - 1.0);
        const double vfa = pow(vf, af), tfb = pow(tf, bf);
        e.dtheta[i] = dth;
        e.dpr[i] = dpr;
        e.dv[i] = (dtau + p.f0 * dpr * vfa * tfb - p.f0 * (sg - pr) * vfa * tfb1 * bf * p.v0 / L * dth) /
                  (p.f0 * (sg - pr) * vfa1 * tfb * af / p.v0);
        e.ddelta[i] = v;
    }
}

__global__ void __launch_bounds__(256) fault_epilogue_kernel(FaultEpilogue e, const double* dtau, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) update_fault_row(e, i, dtau[i]);
}

// ---- Toeplitz form of the fault-fault interaction (the reference's algorithm, equation.jl:44-61) ----
// out[i,j] = sum_l sum_k st[|i-k|, j, l] * relv[k, l]  -- the linear convolution the reference evaluates
// with FFTs, done directly: one CTA per (receiver row j, strike tile); the kernel column st[:,j,l] and
// the forcing column relv[:,l] are staged in shared memory; each thread produces one receiver i.
struct PeerWait {
    const unsigned long long* flags;   // local forcing flags [kMaxWorld]
    unsigned long long* epochs;        // local counters
    int world, rank;
    int warp_poll;                     // poll all peers at once, one lane each (OQ_WAIT=serial: one after the other)
};

// every consumer CTA: find the buffer copy of the current evaluation and make sure all peers delivered
__device__ __forceinline__ size_t consumer_parity(const PeerWait& w)
{
    __shared__ unsigned long long ep_s;
    if (threadIdx.x == 0) {
        unsigned long long ep = 1ull;
        if (w.epochs) {
            ep = *(volatile unsigned long long*)(w.epochs + kEpForcing);
            if (w.world > 1) wait_peers(w.flags, w.world, w.rank, ep, w.epochs + kEpError);
        }
        ep_s = ep;
    }
    __syncthreads();
    return (size_t)((ep_s - 1ull) & 1ull);
}

}  // namespace oq
#include "classmat.cuh"
namespace oq {

__global__ void __launch_bounds__(256)
toeplitz_conv_kernel(const double* __restrict__ st, const double* relv0, size_t relv_stride, PeerWait pw, int nx,
                     int nxi, int f0, int nfl, double* __restrict__ out)
{
    extern __shared__ double sm[];
    double* sk = sm;            // st[0..nx)
    double* sr = sm + nx;       // relv[0..nx)
    const double* relv = relv0 + consumer_parity(pw) * relv_stride;
    const int j = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double acc = 0.0;
    for (int l = 0; l < nxi; ++l) {
        __syncthreads();
        for (int k = threadIdx.x; k < nx; k += blockDim.x) {
            sk[k] = st[k + (size_t)nx * (j + (size_t)nxi * l)];
            sr[k] = relv[k + (size_t)nx * l];
        }
        __syncthreads();
        if (i < nx) {
            for (int k = 0; k < nx; ++k) {
                const int dk = i > k ? i - k : k - i;
                acc = fma(sk[dk], sr[k], acc);
            }
        }
    }
    const int f = i + nx * j;
    if (i < nx && f >= f0 && f < f0 + nfl) out[f - f0] = acc;
}

// ---- the fused matvec -------------------------------------------------------------------------
enum : int { kEpiFault = 0, kEpiStore = 1 };

struct MatvecJob {
    MatOperand op[2];
    int nrows;            // local rows
    int nsegTotal;        // op[0].nseg + op[1].nseg
    double* partial;      // [nrows * nsegTotal]
    unsigned* counters;   // [ceil(nrows / kMvRows)]
    const double* y0;     // optional initial value per row (Toeplitz-form traction rate / accumulate)
    double* yout;         // kEpiStore target
    int epilogue;
    int nitems;           // row blocks * nsegTotal
    // streaming (TMA) plan, see matvec_stream.cuh
    int nrb;              // row blocks of kStR rows
    int nch[2];           // chunks per operand
    int chunks_per_rb;
    int slots;            // partial-sum slots per row (max contributing CTAs / pieces of a row block)
    long long chunk_begin;
};

struct MatvecArgs {
    MatvecJob job[2];     // fault rows, mantle rows
    FaultEpilogue fe;
    PeerWait pw;
    long long total_chunks;
    const int* done;      // optional device flag: skip the evaluation (integration complete)
    // traversal direction: pass[0] counts completed launches (its parity picks forward / reverse row-block order so
    // that an evaluation starts on what the previous one left in L2), pass[1] counts CTAs that finished this launch
    unsigned long long* pass;
    // L2 residency: chunks [0, keep_chunks) of every CTA span are loaded evict-last, the rest evict-first
    // (negative: no eviction hints)
    int keep_chunks;
};

__global__ void __launch_bounds__(kMvThreads)
matvec_fused_kernel(const __grid_constant__ MatvecArgs args)
{
    extern __shared__ __align__(16) double sx[];              // x segment, up to kMvMaxSeg doubles
    __shared__ __align__(8) uint64_t bar;
    __shared__ double red[kMvThreads / 32][kMvRows];
    __shared__ int is_last;

    int item = blockIdx.x;
    const int jsel = item < args.job[0].nitems ? 0 : 1;
    const MatvecJob& job = args.job[jsel];
    if (jsel) item -= args.job[0].nitems;
    const int rb = item / job.nsegTotal;
    int s = item % job.nsegTotal;
    const int osel = s < job.op[0].nseg ? 0 : 1;
    const MatOperand& op = job.op[osel];
    if (osel) s -= job.op[0].nseg;

    const int tid = threadIdx.x;
    const int row0 = rb * kMvRows;
    const int c_begin = s * op.seg_len;
    const int c_end = min(c_begin + op.seg_len, (int)op.ld);   // padding columns are zero in G and x
    const int ncol = c_end - c_begin;

    // stage the forcing-vector segment with one bulk TMA copy
    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    const size_t par = consumer_parity(args.pw);               // includes the CTA-wide barrier
    if (tid == 0) {
        if (args.pw.world > 1) fence_proxy_async();            // peer stores -> async-proxy read
        mbar_arrive_expect_tx(&bar, (unsigned)(ncol * sizeof(double)));
        tma_load_1d(sx, op.x + par * op.x_stride + c_begin, (unsigned)(ncol * sizeof(double)), &bar);
    }

    const double* rowp[kMvRows];
#pragma unroll
    for (int r = 0; r < kMvRows; ++r) {
        const int row = min(row0 + r, job.nrows - 1);          // clamp: tail rows are computed and dropped
        rowp[r] = op.G + (size_t)row * op.ld + c_begin;
    }
    double acc[kMvRows];
#pragma unroll
    for (int r = 0; r < kMvRows; ++r) acc[r] = 0.0;

    // first batch of matrix loads is issued before waiting for x
    int c = 2 * tid;
    double2 g0[kMvRows], g1[kMvRows];
    const bool h0 = c < ncol, h1 = c + kMvColStep < ncol;
#pragma unroll
    for (int r = 0; r < kMvRows; ++r) {
        g0[r] = h0 ? ldg_stream(rowp[r] + c) : make_double2(0.0, 0.0);
        g1[r] = h1 ? ldg_stream(rowp[r] + c + kMvColStep) : make_double2(0.0, 0.0);
    }
    mbar_wait(&bar, 0);
    const double2* sx2 = reinterpret_cast<const double2*>(sx);
    while (c < ncol) {
        const int cn = c + 2 * kMvColStep;
        double2 n0[kMvRows], n1[kMvRows];
        const bool p0 = cn < ncol, p1 = cn + kMvColStep < ncol;
#pragma unroll
        for (int r = 0; r < kMvRows; ++r) {
            n0[r] = p0 ? ldg_stream(rowp[r] + cn) : make_double2(0.0, 0.0);
            n1[r] = p1 ? ldg_stream(rowp[r] + cn + kMvColStep) : make_double2(0.0, 0.0);
        }
        const double2 x0 = sx2[c >> 1];
        const double2 x1 = (c + kMvColStep < ncol) ? sx2[(c + kMvColStep) >> 1] : make_double2(0.0, 0.0);
#pragma unroll
        for (int r = 0; r < kMvRows; ++r) {
            acc[r] = fma(g0[r].x, x0.x, acc[r]);
            acc[r] = fma(g0[r].y, x0.y, acc[r]);
            acc[r] = fma(g1[r].x, x1.x, acc[r]);
            acc[r] = fma(g1[r].y, x1.y, acc[r]);
            g0[r] = n0[r];
            g1[r] = n1[r];
        }
        c = cn;
    }

    // warp-shuffle row reduction, then across the CTA's warps in a fixed order
#pragma unroll
    for (int r = 0; r < kMvRows; ++r) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], off);
    }
    const int warp = tid >> 5, lane = tid & 31;
    if (lane == 0) {
#pragma unroll
        for (int r = 0; r < kMvRows; ++r) red[warp][r] = acc[r];
    }
    __syncthreads();
    double mine = 0.0;
    const int myrow = row0 + tid;
    if (tid < kMvRows && myrow < job.nrows) {
#pragma unroll
        for (int w = 0; w < kMvThreads / 32; ++w) mine += red[w][tid];
    }

    if (job.nsegTotal > 1) {
        if (tid < kMvRows && myrow < job.nrows)
            job.partial[(size_t)myrow * job.nsegTotal + (osel ? job.op[0].nseg : 0) + s] = mine;
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            const unsigned prev = atomicAdd(&job.counters[rb], 1u);
            is_last = (prev == (unsigned)job.nsegTotal - 1u);
            if (is_last) job.counters[rb] = 0u;               // re-arm for the next evaluation
        }
        __syncthreads();
        if (!is_last) return;
        __threadfence();
        if (tid < kMvRows && myrow < job.nrows) {
            mine = 0.0;
            const double* pp = job.partial + (size_t)myrow * job.nsegTotal;
            for (int q = 0; q < job.nsegTotal; ++q) mine += ld_cg(pp + q);   // fixed order: deterministic
        }
    }
    if (tid < kMvRows && myrow < job.nrows) {
        if (job.y0) mine += job.y0[myrow];
        if (job.epilogue == kEpiFault) update_fault_row(args.fe, myrow, mine);
        else job.yout[myrow] = mine;
    }
}

}  // namespace oq
#include "matvec_stream.cuh"
#include "matvec_panel.cuh"
#include "toeplitz_fft.cuh"
namespace oq {

static bool use_direct_toeplitz()
{
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("OQ_TOEPLITZ");
        v = (e && strcmp(e, "direct") == 0) ? 1 : 0;
    }
    return v == 1;
}

static int sm_count()
{
    static int n = 0;
    if (n == 0) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, current_device() >= 0 ? current_device() : 0) == cudaSuccess)
            n = prop.multiProcessorCount;
        else n = 148;
    }
    return n;
}

// streaming plan over both row sets; returns the grid size.  by_cols: chunk the columns that exist (rounded up to
// 16; the panel kernel) instead of the padded leading dimension (the first streaming kernel)
static int plan_stream(MatvecArgs& a, bool by_cols = false)
{
    long long total = 0;
    for (int jb = 0; jb < 2; ++jb) {
        MatvecJob& j = a.job[jb];
        j.nrb = j.nrows > 0 ? (j.nrows + kStR - 1) / kStR : 0;
        for (int o = 0; o < 2; ++o) {
            const size_t width = by_cols ? round_up((size_t)j.op[o].cols, 16) : j.op[o].ld;
            j.nch[o] = (j.op[o].G && j.op[o].cols > 0 && j.nrows > 0) ? (int)((width + kStCH - 1) / kStCH) : 0;
        }
        j.chunks_per_rb = j.nch[0] + j.nch[1];
        if (j.chunks_per_rb == 0) j.nrb = 0;
        j.chunk_begin = total;
        total += (long long)j.nrb * j.chunks_per_rb;
    }
    a.total_chunks = total;
    if (total == 0) return 0;
    const int grid = (int)(total < sm_count() ? total : sm_count());
    const long long min_span = total / grid;
    for (int jb = 0; jb < 2; ++jb) {
        MatvecJob& j = a.job[jb];
        long long s = j.chunks_per_rb > 0 ? (j.chunks_per_rb - 1) / min_span + 2 : 1;
        if (s > j.chunks_per_rb) s = j.chunks_per_rb > 0 ? j.chunks_per_rb : 1;
        j.slots = (int)s;
    }
    return grid;
}

// 0: fused panel kernel (matvec_panel.cuh), 1: streaming kernel behind a forcing kernel (matvec_stream.cuh),
// 2: first LDG kernel.  Default (OQ_MATVEC unset): the fused kernel on one GPU -- one launch per evaluation, fastest
// end to end -- and the streaming pair on row shards: there the evaluation is 40-50 us long and the measured timeline
// of the fused kernel (profiles/r02_panel_timeline_n8.md) shows its grid-wide publication (per-CTA system fence,
// counter, second fence, flags: first multiply 14 us after the launch) costing more than the launch it saves.
static int matvec_variant(int world = 1)
{
    static int v = -2;
    if (v == -2) {
        const char* e = getenv("OQ_MATVEC");
        v = !e ? -1 : strcmp(e, "ldg") == 0 ? 2 : strcmp(e, "stream") == 0 ? 1 : strcmp(e, "panel") == 0 ? 0 : -1;
    }
    return v >= 0 ? v : (world > 1 ? 1 : 0);
}

// chunks per CTA span to keep L2-resident between evaluations: OQ_MATVEC_KEEP_MB megabytes over the grid
// (default: three quarters of the L2; 0 disables the eviction hints)
static int keep_chunks_per_cta(int grid)
{
    static double mb = -1.0;
    if (mb < 0.0) {
        const char* e = getenv("OQ_MATVEC_KEEP_MB");
        if (e) mb = atof(e);
        else {
            cudaDeviceProp prop;
            mb = cudaGetDeviceProperties(&prop, current_device() >= 0 ? current_device() : 0) == cudaSuccess
                     ? 0.75 * prop.l2CacheSize / 1048576.0 : 90.0;
        }
        if (mb < 0.0) mb = 0.0;
    }
    if (mb == 0.0) return -1;
    const double chunk_mb = (double)kStR * kStCH * sizeof(double) / 1048576.0;
    return (int)(mb / (grid * chunk_mb));
}

static bool pingpong_enabled()
{
    static const bool on = [] { const char* e = getenv("OQ_MATVEC_PINGPONG"); return !(e && e[0] == '0'); }();
    return on;
}

static bool use_ldg_matvec() { return matvec_variant() == 2; }

// choose the column split so that the grid has enough CTAs to balance 148 SMs
static void plan_operand(MatOperand& op, int nrowblocks, int other_min_seg)
{
    if (!op.G || op.cols == 0) { op.nseg = 0; op.seg_len = kMvMaxSeg; return; }
    (void)other_min_seg;
    const long target_items = 148L * 4 * 16;
    int want = (int)((target_items + nrowblocks - 1) / (nrowblocks > 0 ? nrowblocks : 1));
    const int max_seg_count = (int)((op.ld + 2 * kMvColStep - 1) / (2 * kMvColStep));
    const int min_seg_count = (int)((op.ld + kMvMaxSeg - 1) / kMvMaxSeg);
    if (want < min_seg_count) want = min_seg_count;
    if (want > max_seg_count) want = max_seg_count;
    int seg_len = (int)round_up((op.ld + want - 1) / want, 2 * kMvColStep);
    if (seg_len > kMvMaxSeg) seg_len = kMvMaxSeg;
    op.seg_len = seg_len;
    op.nseg = (int)((op.ld + seg_len - 1) / seg_len);
}

int plan_job(MatvecJob& job, int nrows)
{
    job.nrows = nrows;
    const int nrb = (nrows + kMvRows - 1) / kMvRows;
    plan_operand(job.op[0], nrb, 0);
    plan_operand(job.op[1], nrb, 0);
    job.nsegTotal = job.op[0].nseg + job.op[1].nseg;
    job.nitems = nrows > 0 ? nrb * job.nsegTotal : 0;
    return nrb;
}

static unsigned long long* g_timeline = nullptr;

static void dump_timeline(int rank)
{
    const char* f = getenv("OQ_TIMELINE");
    if (!f || !g_timeline) return;
    std::vector<unsigned long long> h(160 * 32);
    cudaMemcpy(h.data(), g_timeline, h.size() * 8, cudaMemcpyDeviceToHost);
    static int calls = 0;
    char name[512];
    snprintf(name, sizeof name, "%s.rank%d.call%d", f, rank, calls++);
    if (FILE* fp = fopen(name, "wb")) { fwrite(h.data(), 8, h.size(), fp); fclose(fp); }
}

// the fused kernel: `pro` non-null folds the forcing front end into the launch
int launch_panel(MatvecArgs& a, const ForcingArgs* pro, const ColOwners& own, int ne, int f0, unsigned seq,
                 cudaStream_t stream)
{
    PanelArgs A{};
    int grid = plan_stream(a, true);
    if (grid == 0) {
        if (!pro) return 0;
        grid = 1;      // a rank without rows still takes part in the exchange (publishes an empty slice)
    }
    // row blocks per panel (OQ_PANEL_P = 1, 2, 4 or 6; default 2: measured fastest on one GPU, 3 193 vs 3 086-3 117
    // evaluations/s of the 256x64 problem for 1, 4, 6)
    static const int panel_p = [] { const char* e = getenv("OQ_PANEL_P"); const int v = e ? atoi(e) : 2;
                                    return v == 1 || v == 4 || v == 6 ? v : 2; }();
    void (*kern)(const PanelArgs) = panel_p == 1 ? matvec_panel_kernel<1> : panel_p == 2 ? matvec_panel_kernel<2>
                                    : panel_p == 4 ? matvec_panel_kernel<4> : matvec_panel_kernel<6>;
    static bool attr_set = false;
    if (!attr_set) {
        OQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPnSmemBytes));
        attr_set = true;
    }
    a.pass = nullptr;
    a.keep_chunks = keep_chunks_per_cta(grid);
    A.mv = a;
    A.pro.enabled = pro ? 1 : 0;
    if (pro) A.pro.fa = *pro;
    A.own = own;
    A.ne = ne > 0 ? ne : 1;
    A.reverse = pingpong_enabled() ? (int)(seq & 1u) : 0;
    // debug timeline (OQ_TIMELINE=<file>): globaltimer stamps of the last launch, dumped by oq_rhs_resident
    A.timeline = g_timeline;
    for (int jb = 0; jb < 2; ++jb) {
        const int n0 = a.job[jb].nch[0];
        A.cstart[jb] = n0 > 0 ? (f0 / kPnCH < n0 ? f0 / kPnCH : n0 - 1) : 0;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kPnThreads);
    cfg.dynamicSmemBytes = kPnSmemBytes; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    OQ_CUDA(cudaLaunchKernelEx(&cfg, kern, A));
    OQ_LAUNCHED();
    return 0;
}

struct PanelLaunch {
    const ForcingArgs* pro = nullptr;   // fold the forcing front end into the launch
    ColOwners own;
    int ne = 1, f0 = 0;
    unsigned seq = 0;
};

int launch_matvec(MatvecArgs& a, cudaStream_t stream, const PanelLaunch* pl = nullptr)
{
    if (matvec_variant(pl ? pl->own.world : 1) == 0) {
        if (pl) return launch_panel(a, pl->pro, pl->own, pl->ne, pl->f0, pl->seq, stream);
        ColOwners own;
        own.world = 1; own.fb[1] = 0x7fffffff; own.eb[1] = 0x7fffffff;
        return launch_panel(a, nullptr, own, 1, 0, 0u, stream);
    }
    if (!use_ldg_matvec()) {
        const int grid = plan_stream(a);
        if (grid == 0) return 0;
        static bool stream_attr_set = false;
        if (!stream_attr_set) {
            OQ_CUDA(cudaFuncSetAttribute(matvec_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kStSmemBytes));
            stream_attr_set = true;
        }
        if (!pingpong_enabled()) a.pass = nullptr;
        a.keep_chunks = keep_chunks_per_cta(grid);
        // programmatic dependent launch: start while the forcing kernel still runs; the kernel prefetches its
        // first ring of matrix chunks and only then waits for the predecessor (pdl_wait)
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kStThreads);
        cfg.dynamicSmemBytes = kStSmemBytes; cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        OQ_CUDA(cudaLaunchKernelEx(&cfg, matvec_stream_kernel, a));
        OQ_LAUNCHED();
        return 0;
    }
    const int items = a.job[0].nitems + a.job[1].nitems;
    if (items == 0) return 0;
    static bool attr_set = false;
    const size_t smem = kMvMaxSeg * sizeof(double);
    if (!attr_set) {
        OQ_CUDA(cudaFuncSetAttribute(matvec_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    matvec_fused_kernel<<<items, kMvThreads, smem, stream>>>(a);
    OQ_LAUNCHED();
    return 0;
}

StateView view_of(const OqProblem* p, double* b)
{
    StateView s{};
    s.v = b + p->part_off[0];
    s.theta = b + p->part_off[1];
    if (p->kind == kFaultOnly) s.delta = b + p->part_off[2];
    else if (p->kind == kDilatancy) { s.delta = b + p->part_off[2]; s.pr = b + p->part_off[3]; }
    else { s.eps = b + p->part_off[2]; s.sig = b + p->part_off[3]; s.delta = b + p->part_off[4]; }
    return s;
}

// views over five separately allocated partitions (the host's ArrayPartition, mapped into the device's address space)
static StateView view_of_parts(const OqProblem* p, double* const* b)
{
    StateView s{};
    s.v = b[0];
    s.theta = b[1];
    if (p->kind == kFaultOnly) s.delta = b[2];
    else if (p->kind == kDilatancy) { s.delta = b[2]; s.pr = b[3]; }
    else { s.eps = b[2]; s.sig = b[3]; s.delta = b[4]; }
    return s;
}

static int rhs_views(OqProblem* p, const StateView& in, const StateView& out, const double* uin, const StageSpec* stage);

int rhs_device(OqProblem* p, const double* uin, double* du, const StageSpec* stage)
{
    return rhs_views(p, view_of(p, const_cast<double*>(uin)), view_of(p, du), uin, stage);
}

// `uin`: base of the state-shaped buffer the fused stage combination fills (needed only with a stage spec)
static int rhs_views(OqProblem* p, const StateView& in, const StateView& out, const double* uin, const StageSpec* stage)
{
    cudaStream_t st = p->stream;
    // 1. forcing vectors; every rank's slice is stored straight into all windows (fused all-gather)
    ForcingArgs fa{};
    fa.v = in.v; fa.sig = in.sig; fa.deps_out = out.eps;
    fa.peers = comm_targets(p); fa.wl = p->wl; fa.epochs = p->epochs;
    fa.nfl = p->nfl; fa.f0 = p->f0; fa.nel = p->kind == kViscoelastic ? p->nel : 0; fa.e0 = p->e0; fa.ne = p->ne;
    fa.vpl = p->fp.vpl; fa.mp = p->mp;
    if (stage) fa.stage = *stage;
    fa.y = const_cast<double*>(uin);
    fa.n_fault_parts = p->kind == kDilatancy ? 4 : 3;
    fa.off_fault[0] = p->part_off[0]; fa.off_fault[1] = p->part_off[1];
    if (p->kind == kViscoelastic) {
        fa.off_fault[2] = p->part_off[4]; fa.off_fault[3] = 0;
        fa.off_eps = p->part_off[2]; fa.off_sig = p->part_off[3];
    } else {
        fa.off_fault[2] = p->part_off[2]; fa.off_fault[3] = p->kind == kDilatancy ? p->part_off[3] : 0;
        fa.off_eps = fa.off_sig = 0;
    }
    const int nthr = p->nfl > fa.nel ? p->nfl : fa.nel;
    // single-rank, FFT form, no dense operand: the forward transform forms v - vpl itself (one launch fewer)
    const bool direct_fft = p->gf11_form == OQ_GF11_FFT && p->kind != kViscoelastic && p->world == 1 && !stage &&
                            !use_direct_toeplitz();
    // dense fault-fault operand + panel kernel: the forcing front end runs inside the matvec launch (one launch per
    // evaluation); the FFT form needs the forcing vector before its transforms, the older kernels have no prologue
    static const bool split_forcing = [] { const char* e = getenv("OQ_FORCING"); return e && strcmp(e, "split") == 0; }();
    const bool class_ops = (p->g12 && p->g12->cls) || (p->g21 && p->g21->cls) || (p->g22 && p->g22->cls);
    const bool fused = matvec_variant(p->world) == 0 && p->gf11_form == OQ_GF11_DENSE && !split_forcing && p->nfl + fa.nel > 0 &&
                       !class_ops;
    if (!direct_fft && !fused) {
    // small shards: ONE block (no cross-block handshake before the publication); large ones: 256-thread blocks
    if (nthr <= 4096) forcing_kernel<<<1, 1024, 0, st>>>(fa);
    else forcing_kernel<<<(nthr + 255) / 256, 256, 0, st>>>(fa);
    OQ_LAUNCHED();
    }
    static const int warp_poll = [] { const char* e = getenv("OQ_WAIT"); return (e && strcmp(e, "serial") == 0) ? 0 : 1; }();
    PeerWait pw{p->flags, p->epochs, p->world, p->rank, warp_poll};
    // 2. fault-fault interaction in its translation-invariant form (the reference's algorithm) when requested
    FaultEpilogue fe{};
    fe.fp = p->fp; fe.v = in.v; fe.theta = in.theta; fe.pr = in.pr;
    fe.dv = out.v; fe.dtheta = out.theta; fe.ddelta = out.delta; fe.dpr = out.pr;
    fe.dilatancy = p->kind == kDilatancy;
    const double* y0 = nullptr;
    bool epilogue_done = false;
    if (p->gf11_form == OQ_GF11_FFT && p->nfl > 0) {
        if (use_direct_toeplitz()) {
            dim3 grid((p->nx + 255) / 256, p->nxi);
            toeplitz_conv_kernel<<<grid, 256, 2 * p->nx * sizeof(double), st>>>(p->st.p, p->relv, p->wl.relv_len, pw,
                                                                                p->nx, p->nxi, p->f0, p->nfl, p->dtau0.p);
            OQ_LAUNCHED();
        } else {
            const int N = p->fftN, nfreq = N / 2 + 1;
            const size_t fsmem = (2 * (size_t)N + N / 2) * sizeof(cplx);
            const size_t ismem = fsmem + (size_t)kFftLGroups * nfreq * sizeof(cplx);
            const cplx* Wtw = reinterpret_cast<const cplx*>(p->Wtw.p);
            fft_forward_kernel<<<p->nxi, 256, fsmem, st>>>(p->relv, p->wl.relv_len, pw, direct_fft ? in.v : nullptr,
                                                         p->fp.vpl, p->nx, N, Wtw, reinterpret_cast<cplx*>(p->Rhat.p));
            OQ_LAUNCHED();
            // with no dense operand on the fault rows the pointwise physics is fused into the inverse transform
            const bool fuse = p->kind != kViscoelastic;
            fft_inverse_kernel<<<p->fnj, kFftInvThreads, ismem, st>>>(
                p->Ghat.p, reinterpret_cast<const cplx*>(p->Rhat.p), p->nxi, p->fnj, p->nx, N, p->fj0, p->f0, p->nfl,
                Wtw, p->dtau0.p, fuse ? 1 : 0, fe);
            OQ_LAUNCHED();
            epilogue_done = fuse;
        }
        y0 = p->dtau0.p;
    }
    // 2b. operands kept in class form (classmat.cuh) multiply from their tables into the vectors the dense kernel
    //     starts from (or, with no dense operand on those rows, straight into the result)
    const int* done = stage ? stage->done : nullptr;
    const double* ym0 = nullptr;
    if (class_ops) {
        const bool dense_f = p->opf[0].G || p->opf[1].G, dense_m = p->opm[0].G || p->opm[1].G;
        (void)dense_f;
        if (p->g21->cls && p->nfl > 0) {
            OQ_TRY(class_matvec(p->g21, p->reldeps, p->wl.reldeps_len, y0, p->dtau0.p, pw, done, st));
            y0 = p->dtau0.p;
        }
        double* ym = dense_m ? p->dsig0.p : out.sig;
        if (p->g12->cls && p->nel > 0) {
            OQ_TRY(class_matvec(p->g12, p->relv, p->wl.relv_len, ym0, ym, pw, done, st));
            ym0 = ym;
        }
        if (p->g22->cls && p->nel > 0) {
            OQ_TRY(class_matvec(p->g22, p->reldeps, p->wl.reldeps_len, ym0, ym, pw, done, st));
            ym0 = ym;
        }
        if (!dense_m) ym0 = nullptr;          // already in out.sig
    }
    // 3. fused matvec + pointwise physics
    MatvecArgs a{};
    a.fe = fe;
    a.pw = pw;
    a.done = done;
    a.job[0].op[0] = p->opf[0]; a.job[0].op[1] = p->opf[1];
    a.job[0].partial = p->partial_f.p; a.job[0].counters = p->counters.p;
    a.job[0].y0 = y0; a.job[0].epilogue = kEpiFault;
    const int nrbf = plan_job(a.job[0], p->nfl);
    a.job[1].op[0] = p->opm[0]; a.job[1].op[1] = p->opm[1];
    a.job[1].partial = p->partial_m.p; a.job[1].counters = p->counters.p + nrbf;
    a.pass = p->ticket.p;
    a.job[1].yout = out.sig; a.job[1].epilogue = kEpiStore; a.job[1].y0 = ym0;
    plan_job(a.job[1], p->kind == kViscoelastic ? 6 * p->nel : 0);
    if (a.job[0].nitems == 0 && p->nfl > 0 && !epilogue_done) {
        // no dense operand on the fault rows (Toeplitz form, fault-only): standalone epilogue
        OQ_CHECK(y0 != nullptr, "fault rows have no Green's operand");
        fault_epilogue_kernel<<<(p->nfl + 255) / 256, 256, 0, st>>>(fe, y0, p->nfl);
        OQ_LAUNCHED();
    }
    const bool capturing = [&] {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(st, &cs);
        return cs != cudaStreamCaptureStatusNone;
    }();
    const bool prof = p->prof_on && !capturing && p->prof_used + 2 <= p->prof_ev.size();
    if (prof) OQ_CUDA(cudaEventRecord(p->prof_ev[p->prof_used], st));
    PanelLaunch pl;
    pl.pro = fused ? &fa : nullptr;
    pl.own = comm_owners(p);
    pl.ne = p->ne; pl.f0 = p->f0; pl.seq = p->mv_seq++;
    OQ_TRY(launch_matvec(a, st, &pl));
    if (prof) {
        OQ_CUDA(cudaEventRecord(p->prof_ev[p->prof_used + 1], st));
        p->prof_used += 2;
    }
    return 0;
}

// plain y = A x / y += A x on a shard (the matvecmul! slot)
int gemv_device(const OqMatrix* A, const double* x_dev_padded, const double* y_in, double* y_out,
                double* partial, unsigned* counters, unsigned long long* ticket, cudaStream_t st)
{
    MatvecArgs a{};
    a.pass = ticket;
    a.job[0].op[0].G = A->d.p; a.job[0].op[0].ld = A->ld; a.job[0].op[0].x = x_dev_padded;
    a.job[0].op[0].cols = A->cols;
    a.job[0].partial = partial; a.job[0].counters = counters; a.job[0].y0 = y_in; a.job[0].yout = y_out;
    a.job[0].epilogue = kEpiStore;
    plan_job(a.job[0], A->local_rows);
    return launch_matvec(a, st);
}

int gemv_scratch_sizes(const OqMatrix* A, size_t* npartial, size_t* ncounters)
{
    MatvecArgs a{};
    MatvecJob& j = a.job[0];
    j.op[0].G = A->d.p; j.op[0].ld = A->ld; j.op[0].cols = A->cols;
    const int nrb = plan_job(j, A->local_rows);
    plan_stream(a, true);
    const int slots_cols = j.slots;
    plan_stream(a, false);
    if (slots_cols > j.slots) j.slots = slots_cols;
    const int per_row = j.nsegTotal > j.slots ? j.nsegTotal : j.slots;
    *npartial = (size_t)A->local_rows * (per_row > 0 ? per_row : 1);
    *ncounters = nrb;
    return 0;
}

}  // namespace oq

using namespace oq;

OqProblem::~OqProblem()
{
    for (cudaEvent_t e : prof_ev) cudaEventDestroy(e);
    if (rhs_graph) cudaGraphExecDestroy(rhs_graph);
    comm_release(this);
    if (err_host) cudaFreeHost(err_host);
    if (stream) cudaStreamDestroy(stream);
}

// ---- problem construction ----------------------------------------------------------------------
static int finish_problem(OqProblem* p, const OqFaultProperty* pf, const OqDilatancyProperty* dila,
                          const OqMantleProperty* pa, const double* st_host)
{
    OQ_CHECK(pf && pf->a && pf->b && pf->L && pf->sigma, "fault property is NULL");
    // property.jl:20-24
    OQ_CHECK(pf->f0 > 0, "f0 must be > 0");
    OQ_CHECK(pf->v0 > 0, "v0 must be > 0");
    OQ_CHECK(pf->eta > 0, "eta must be > 0");
    OQ_CHECK(pf->vpl > 0, "vpl must be > 0");
    const int nfl = p->nfl, nel = p->nel;
    const int nlaws = pa ? pa->nlaws : 0;
    OQ_CHECK(!pa || (nlaws >= 1 && pa->gamma && pa->n && pa->deps0), "mantle property is malformed");
    // pack local slices of the property arrays
    std::vector<double> h;
    auto push_fault = [&](const double* src) { h.insert(h.end(), src + p->f0, src + p->f1); };
    push_fault(pf->a); push_fault(pf->b); push_fault(pf->L); push_fault(pf->sigma);
    if (dila) {
        OQ_CHECK(dila->tp && dila->eps && dila->beta && dila->p0, "dilatancy property is NULL");
        push_fault(dila->tp); push_fault(dila->eps); push_fault(dila->beta); push_fault(dila->p0);
    }
    for (int arr = 0; arr < 2 && pa; ++arr)
        for (int l = 0; l < nlaws; ++l) {
            const double* src = (arr == 0 ? pa->gamma : pa->n) + (size_t)l * p->ne;
            h.insert(h.end(), src + p->e0, src + p->e1);
        }
    if (h.empty()) h.push_back(0.0);
    OQ_TRY(p->props.upload(h.data(), h.size()));
    const double* q = p->props.p;
    p->fp.a = q; p->fp.b = q + nfl; p->fp.L = q + 2 * (size_t)nfl; p->fp.sigma = q + 3 * (size_t)nfl;
    q += 4 * (size_t)nfl;
    if (dila) {
        p->fp.tp = q; p->fp.epsd = q + nfl; p->fp.beta = q + 2 * (size_t)nfl; p->fp.p0 = q + 3 * (size_t)nfl;
        q += 4 * (size_t)nfl;
    }
    p->fp.eta = pf->eta; p->fp.vpl = pf->vpl; p->fp.f0 = pf->f0; p->fp.v0 = pf->v0;
    if (pa) {
        p->mp.gamma = q; p->mp.npow = q + (size_t)nlaws * nel; p->mp.nlaws = nlaws;
        for (int k = 0; k < 6; ++k) p->mp.deps0[k] = pa->deps0[k];
    }
    // state partitions
    if (p->kind == kFaultOnly) { p->nparts = 3; p->part_len[0] = p->part_len[1] = p->part_len[2] = nfl; }
    else if (p->kind == kDilatancy) { p->nparts = 4; for (int i = 0; i < 4; ++i) p->part_len[i] = nfl; }
    else {
        p->nparts = 5;
        p->part_len[0] = p->part_len[1] = p->part_len[4] = nfl;
        p->part_len[2] = p->part_len[3] = 6 * nel;
    }
    size_t off = 0;
    for (int i = 0; i < p->nparts; ++i) { p->part_off[i] = off; off += round_up((size_t)p->part_len[i], 2); }
    p->nstate = off;
    p->nstate_global = p->kind == kViscoelastic ? 3 * (size_t)p->nf + 12 * (size_t)p->ne
                                                : (size_t)p->nparts * p->nf;
    OQ_TRY(p->u.alloc(p->nstate + 2)); OQ_TRY(p->u.zero());
    for (int i = 0; i < 7; ++i) { OQ_TRY(p->k[i].alloc(p->nstate + 2)); OQ_TRY(p->k[i].zero()); }
    OQ_TRY(p->utmp.alloc(p->nstate + 2)); OQ_TRY(p->utmp.zero());
    OQ_TRY(p->unew.alloc(p->nstate + 2)); OQ_TRY(p->unew.zero());
    // forcing vectors, reduction slots and flags: one peer-visible window
    OQ_TRY(comm_alloc_window(p));
    // operands
    if (p->gf11_form == OQ_GF11_DENSE) {
        OQ_CHECK(p->g11, "dense gf11 requested but no matrix given");
        p->opf[0].G = p->g11->d.p; p->opf[0].ld = p->g11->ld; p->opf[0].x = p->relv; p->opf[0].x_stride = p->wl.relv_len; p->opf[0].cols = p->g11->cols;
    } else {
        OQ_CHECK(st_host, "Toeplitz gf11 requested but no kernel given");
        OQ_TRY(p->st.upload(st_host, (size_t)p->nx * p->nxi * p->nxi));
        OQ_TRY(p->dtau0.alloc(nfl > 0 ? nfl : 1));
        OQ_CHECK(2 * (size_t)p->nx * sizeof(double) <= 48 * 1024, "nx too large for the Toeplitz kernel");
        // FFT form: transform length = power of two >= 2nx-1; receiver rows j that intersect this rank's shard
        int N = 2;
        while (N < 2 * p->nx - 1) N <<= 1;
        OQ_CHECK((2 * (size_t)N + N / 2 + (size_t)kFftLGroups * (N / 2 + 1)) * 16 <= 200 * 1024,
                 "nx = %d too large for the shared-memory FFT", p->nx);
        p->fftN = N;
        p->fj0 = nfl > 0 ? p->f0 / p->nx : 0;
        p->fnj = nfl > 0 ? (p->f1 - 1) / p->nx + 1 - p->fj0 : 0;
        const size_t nfreq = N / 2 + 1;
        OQ_TRY(p->Rhat.alloc(2 * nfreq * p->nxi + 2));
        OQ_TRY(p->That.alloc(2 * nfreq * (p->fnj > 0 ? p->fnj : 1) + 2));
        OQ_TRY(p->Ghat.alloc(nfreq * (p->fnj > 0 ? p->fnj : 1) * p->nxi + 1));
        if (p->fnj > 0) {
            const size_t nt = nfreq * p->fnj * p->nxi;
            toeplitz_spectrum_kernel<<<(unsigned)((nt + 255) / 256), 256>>>(p->st.p, p->nx, p->nxi, N, p->fj0, p->fnj,
                                                                            p->Ghat.p);
            OQ_LAUNCHED();
            OQ_TRY(p->Wtw.alloc((size_t)N + 2));
            twiddle_kernel<<<(N / 2 + 255) / 256, 256>>>(N, reinterpret_cast<cplx*>(p->Wtw.p));
            OQ_LAUNCHED();
            const size_t fsmem = (2 * (size_t)N + N / 2) * sizeof(cplx);
            const size_t ismem = fsmem + (size_t)kFftLGroups * nfreq * sizeof(cplx);
            if (ismem > 48 * 1024) {
                OQ_CUDA(cudaFuncSetAttribute(fft_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
                OQ_CUDA(cudaFuncSetAttribute(fft_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ismem));
            }
        }
    }
    // class-form operands (classmat.cuh) accumulate into the vectors the dense kernel starts from
    if (p->kind == kViscoelastic && p->g21->cls && !p->dtau0.p) OQ_TRY(p->dtau0.alloc(nfl > 0 ? nfl : 1));
    if (p->kind == kViscoelastic && (bool)p->g12->cls != (bool)p->g22->cls) OQ_TRY(p->dsig0.alloc(nel > 0 ? 6 * (size_t)nel : 1));
    if (p->kind == kViscoelastic) {
        p->opf[1].G = p->g21->d.p; p->opf[1].ld = p->g21->ld; p->opf[1].x = p->reldeps; p->opf[1].x_stride = p->wl.reldeps_len; p->opf[1].cols = p->g21->cols; p->opf[1].x_kind = 1;
        p->opm[0].G = p->g12->d.p; p->opm[0].ld = p->g12->ld; p->opm[0].x = p->relv; p->opm[0].x_stride = p->wl.relv_len; p->opm[0].cols = p->g12->cols;
        p->opm[1].G = p->g22->d.p; p->opm[1].ld = p->g22->ld; p->opm[1].x = p->reldeps; p->opm[1].x_stride = p->wl.reldeps_len; p->opm[1].cols = p->g22->cols; p->opm[1].x_kind = 1;
    }
    // matvec scratch
    MatvecArgs plan{};
    MatvecJob &jf = plan.job[0], &jm = plan.job[1];
    jf.op[0] = p->opf[0]; jf.op[1] = p->opf[1];
    jm.op[0] = p->opm[0]; jm.op[1] = p->opm[1];
    const int nrbf = plan_job(jf, nfl);
    const int nrbm = plan_job(jm, p->kind == kViscoelastic ? 6 * nel : 0);
    // (the kernel variant is chosen per launch -- the world size is not known yet: size the scratch for both plans)
    plan_stream(plan, true);
    const int sf = jf.slots, sm_ = jm.slots;
    plan_stream(plan, false);
    if (sf > jf.slots) jf.slots = sf;
    if (sm_ > jm.slots) jm.slots = sm_;
    p->nseg_f = jf.nsegTotal > jf.slots ? jf.nsegTotal : jf.slots;
    p->nseg_m = jm.nsegTotal > jm.slots ? jm.nsegTotal : jm.slots;
    OQ_TRY(p->partial_f.alloc((size_t)nfl * (p->nseg_f > 0 ? p->nseg_f : 1) + 1));
    OQ_TRY(p->partial_m.alloc((size_t)6 * nel * (p->nseg_m > 0 ? p->nseg_m : 1) + 1));
    OQ_TRY(p->counters.alloc((size_t)nrbf + nrbm + 1)); OQ_TRY(p->counters.zero());
    OQ_TRY(p->ticket.alloc(2)); OQ_TRY(p->ticket.zero());
    OQ_TRY(p->errpart.alloc(1024)); OQ_TRY(p->errpart.zero());
    OQ_TRY(p->ctl.alloc(32)); OQ_TRY(p->ctl.zero());
    OQ_CUDA(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
    if (getenv("OQ_TIMELINE") && !g_timeline) {          // debug timeline of the fused kernel (never inside a capture)
        OQ_CUDA(cudaMalloc(&g_timeline, 160 * 32 * sizeof(unsigned long long)));
        OQ_CUDA(cudaMemset(g_timeline, 0, 160 * 32 * sizeof(unsigned long long)));
    }
    OQ_CUDA(cudaDeviceSynchronize());
    return 0;
}

static int check_matrix(const OqMatrix* m, const char* name, int kind, int rows_global, int cols)
{
    OQ_CHECK(m, "%s is NULL", name);
    OQ_CHECK(m->row_kind == kind, "%s has the wrong row kind", name);
    OQ_CHECK(m->global_rows == rows_global && m->cols == cols, "%s is %dx%d, expected %dx%d", name, m->global_rows,
             m->cols, rows_global, cols);
    return 0;
}

extern "C" {

int oq_problem_create_fault(int nx, int nxi, int gf11_form, const OqMatrix* g11, const double* st_toeplitz,
                            const OqFaultProperty* pf, const OqDilatancyProperty* dila, OqProblem** out)
{
    OQ_CHECK(out && nx > 0 && nxi > 0, "bad argument");
    OQ_TRY(enter());
    OqProblem* p = new OqProblem();
    p->kind = dila ? kDilatancy : kFaultOnly;
    p->nx = nx; p->nxi = nxi; p->nf = nx * nxi; p->gf11_form = gf11_form;
    int rc = 0;
    if (gf11_form == OQ_GF11_DENSE) {
        rc = check_matrix(g11, "g11", OQ_ROWS_FAULT, p->nf, p->nf);
        if (!rc) { p->g11 = g11; p->f0 = g11->row_begin; p->f1 = g11->row_end; }
    } else if (gf11_form == OQ_GF11_FFT) {
        p->f0 = 0; p->f1 = p->nf;
    } else rc = fail("unknown gf11 form %d", gf11_form);
    p->nfl = p->f1 - p->f0;
    if (!rc) rc = finish_problem(p, pf, dila, nullptr, st_toeplitz);
    if (rc) { delete p; return rc; }
    *out = p;
    return 0;
}

int oq_problem_create_viscoelastic(int nx, int nxi, int ne, int gf11_form, const OqMatrix* g11,
                                   const double* st_toeplitz, const OqMatrix* g12, const OqMatrix* g21,
                                   const OqMatrix* g22, const OqFaultProperty* pf, const OqMantleProperty* pa,
                                   OqProblem** out)
{
    OQ_CHECK(out && nx > 0 && nxi > 0 && ne > 0 && pa, "bad argument");
    OQ_TRY(enter());
    OqProblem* p = new OqProblem();
    p->kind = kViscoelastic;
    p->nx = nx; p->nxi = nxi; p->nf = nx * nxi; p->ne = ne; p->gf11_form = gf11_form;
    int rc = check_matrix(g12, "g12", OQ_ROWS_MANTLE, 6 * ne, p->nf);
    if (!rc) rc = check_matrix(g21, "g21", OQ_ROWS_FAULT, p->nf, 6 * ne);
    if (!rc) rc = check_matrix(g22, "g22", OQ_ROWS_MANTLE, 6 * ne, 6 * ne);
    if (!rc) {
        p->g12 = g12; p->g21 = g21; p->g22 = g22;
        p->f0 = g21->row_begin; p->f1 = g21->row_end;
        p->e0 = g22->row_begin; p->e1 = g22->row_end;
        if (g12->row_begin != p->e0 || g12->row_end != p->e1) rc = fail("g12 and g22 shard different elements");
    }
    if (!rc && gf11_form == OQ_GF11_DENSE) {
        rc = check_matrix(g11, "g11", OQ_ROWS_FAULT, p->nf, p->nf);
        if (!rc && (g11->row_begin != p->f0 || g11->row_end != p->f1)) rc = fail("g11 and g21 shard different rows");
        p->g11 = g11;
    } else if (!rc && gf11_form != OQ_GF11_FFT) rc = fail("unknown gf11 form %d", gf11_form);
    p->nfl = p->f1 - p->f0; p->nel = p->e1 - p->e0;
    if (!rc) rc = finish_problem(p, pf, nullptr, pa, st_toeplitz);
    if (rc) { delete p; return rc; }
    *out = p;
    return 0;
}

int oq_problem_destroy(OqProblem* p)
{
    if (p) { enter(); cudaDeviceSynchronize(); delete p; }
    return 0;
}

int oq_problem_layout(const OqProblem* p, int* nparts, int* lengths)
{
    OQ_CHECK(p && nparts && lengths, "NULL argument");
    *nparts = p->nparts;
    for (int i = 0; i < 5; ++i) lengths[i] = i < p->nparts ? p->part_len[i] : 0;
    return 0;
}

static int upload_parts(OqProblem* p, const double* const* parts, double* dst)
{
    for (int i = 0; i < p->nparts; ++i) {
        OQ_CHECK(parts[i] || p->part_len[i] == 0, "state partition %d is NULL", i);
        if (p->part_len[i])
            OQ_CUDA(cudaMemcpyAsync(dst + p->part_off[i], parts[i], p->part_len[i] * sizeof(double),
                                    cudaMemcpyHostToDevice, p->stream));
    }
    return 0;
}

static int download_parts(const OqProblem* p, const double* src, double* const* parts)
{
    for (int i = 0; i < p->nparts; ++i) {
        OQ_CHECK(parts[i] || p->part_len[i] == 0, "state partition %d is NULL", i);
        if (p->part_len[i])
            OQ_CUDA(cudaMemcpyAsync(parts[i], src + p->part_off[i], p->part_len[i] * sizeof(double),
                                    cudaMemcpyDeviceToHost, p->stream));
    }
    OQ_CUDA(cudaStreamSynchronize(p->stream));
    return 0;
}

// Device addresses of host partitions that are page-locked and mapped (cudaHostAlloc / cudaHostRegister, e.g. a
// pinned torch tensor or a registered Julia array); false if any partition is ordinary pageable memory.
static bool mapped_parts(const OqProblem* p, const double* const* parts, double** dev)
{
    for (int i = 0; i < 5; ++i) dev[i] = nullptr;
    for (int i = 0; i < p->nparts; ++i) {
        if (!p->part_len[i]) continue;
        if (!parts[i]) return false;
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, parts[i]) != cudaSuccess) { cudaGetLastError(); return false; }
        if (at.type != cudaMemoryTypeHost || !at.devicePointer) return false;
        dev[i] = static_cast<double*>(at.devicePointer);
    }
    return true;
}

static bool zero_copy_enabled()
{
    static const bool on = [] { const char* e = getenv("OQ_RHS_ZEROCOPY"); return !(e && e[0] == '0'); }();
    return on;
}

int oq_rhs(OqProblem* p, double t, const double* const* u_parts, double* const* du_parts)
{
    (void)t;   // the system is autonomous (equation.jl:156-205 never reads t)
    OQ_CHECK(p && u_parts && du_parts, "NULL argument");
    OQ_TRY(enter());
    // Page-locked host arrays are read and written by the kernels themselves over PCIe (no staging copies, no
    // copy-engine launches: the evaluation is two kernel launches and one synchronisation); pageable arrays are
    // staged through device buffers.
    double *du_dev[5], *u_dev[5];
    // (operands in class form accumulate into the result vector: they take the staged path, their evaluations are
    // milliseconds long and two copies of the state do not matter)
    const bool class_ops = (p->g12 && p->g12->cls) || (p->g21 && p->g21->cls) || (p->g22 && p->g22->cls);
    if (zero_copy_enabled() && !class_ops && mapped_parts(p, u_parts, u_dev) &&
        mapped_parts(p, const_cast<const double* const*>(du_parts), du_dev)) {
        comm_clear_error(p);
        OQ_TRY(rhs_views(p, view_of_parts(p, u_dev), view_of_parts(p, du_dev), nullptr, nullptr));
        OQ_CUDA(cudaStreamSynchronize(p->stream));
        return comm_check_error(p, "oq_rhs");
    }
    comm_clear_error(p);
    OQ_TRY(upload_parts(p, u_parts, p->utmp.p));
    OQ_TRY(rhs_device(p, p->utmp.p, p->unew.p));
    OQ_TRY(download_parts(p, p->unew.p, du_parts));
    return comm_check_error(p, "oq_rhs");
}

int oq_state_set(OqProblem* p, const double* const* u_parts)
{
    OQ_CHECK(p && u_parts, "NULL argument");
    OQ_TRY(enter());
    OQ_TRY(upload_parts(p, u_parts, p->u.p));
    OQ_CUDA(cudaStreamSynchronize(p->stream));
    return 0;
}

int oq_state_get(const OqProblem* p, double* const* u_parts)
{
    OQ_CHECK(p && u_parts, "NULL argument");
    OQ_TRY(enter());
    return download_parts(p, p->u.p, u_parts);
}

int oq_state_get_du(const OqProblem* p, double* const* du_parts)
{
    OQ_CHECK(p && du_parts, "NULL argument");
    OQ_TRY(enter());
    return download_parts(p, p->k[0].p, du_parts);
}

int oq_rhs_resident(OqProblem* p, int nevals, double* ms_total)
{
    OQ_CHECK(p && nevals >= 0, "bad argument");
    OQ_TRY(enter());
    // Without per-launch profiling the evaluation is replayed from a CUDA graph (every evaluation-dependent
    // scalar -- buffer parity, epochs -- lives in device memory), which removes the launch gaps that dominate
    // small problems.
    const bool graph = !p->prof_on && nevals >= 4;
    comm_clear_error(p);
    if (graph && !p->rhs_graph) {
        cudaGraph_t g = nullptr;
        p->mv_seq &= ~1u;                             // the captured pair starts with a forward traversal
        OQ_CUDA(cudaStreamBeginCapture(p->stream, cudaStreamCaptureModeThreadLocal));
        const int64_t before = g_launches.load();
        int rc = rhs_device(p, p->u.p, p->k[0].p);
        if (!rc) rc = rhs_device(p, p->u.p, p->k[0].p);
        p->rhs_graph_launches = g_launches.load() - before;
        g_launches.fetch_sub(p->rhs_graph_launches);
        const cudaError_t ce = cudaStreamEndCapture(p->stream, &g);
        if (rc) { if (g) cudaGraphDestroy(g); return rc; }
        OQ_CUDA(ce);
        const cudaError_t ie = cudaGraphInstantiate(&p->rhs_graph, g, 0);
        cudaGraphDestroy(g);
        OQ_CUDA(ie);
    }
    // Row-sharded runs: one untimed evaluation first.  Its exchange lines the ranks' streams up on the device, so the
    // event pair below brackets nevals evaluations of device work and not the few tens of microseconds by which
    // the host threads of the ranks reach this call apart.
    if (p->world > 1 && nevals > 0) OQ_TRY(rhs_device(p, p->u.p, p->k[0].p));
    EventTimer tm;
    OQ_TRY(tm.start(p->stream));
    int i = 0;
    if (graph) {
        if (p->mv_seq & 1u) { OQ_TRY(rhs_device(p, p->u.p, p->k[0].p)); ++i; }   // keep the direction alternating
        for (; i + 2 <= nevals; i += 2) {
            OQ_CUDA(cudaGraphLaunch(p->rhs_graph, p->stream));
            g_launches.fetch_add(p->rhs_graph_launches);
        }
    }
    for (; i < nevals; ++i) OQ_TRY(rhs_device(p, p->u.p, p->k[0].p));
    OQ_TRY(tm.stop(ms_total, p->stream));
    dump_timeline(p->rank);
    return comm_check_error(p, "oq_rhs_resident");
}

int oq_profile_enable(OqProblem* p, int on)
{
    OQ_CHECK(p, "NULL problem");
    OQ_TRY(enter());
    if (on && p->prof_ev.empty()) {
        p->prof_ev.resize(2 * 4096);
        for (auto& e : p->prof_ev) OQ_CUDA(cudaEventCreate(&e));
    }
    p->prof_on = on != 0;
    p->prof_used = 0;
    return 0;
}

int oq_profile_read(OqProblem* p, double* matvec_ms_total, int64_t* launches)
{
    OQ_CHECK(p && matvec_ms_total && launches, "NULL argument");
    OQ_TRY(enter());
    OQ_CUDA(cudaStreamSynchronize(p->stream));
    double tot = 0.0;
    for (size_t i = 0; i + 1 < p->prof_used; i += 2) {
        float ms = 0;
        OQ_CUDA(cudaEventElapsedTime(&ms, p->prof_ev[i], p->prof_ev[i + 1]));
        tot += ms;
    }
    *matvec_ms_total = tot;
    *launches = (int64_t)(p->prof_used / 2);
    p->prof_used = 0;
    return 0;
}

int oq_rhs_bytes(const OqProblem* p, double* bytes)
{
    OQ_CHECK(p && bytes, "NULL argument");
    double b = 0.0;
    const OqMatrix* ms[4] = {p->g11, p->g12, p->g21, p->g22};
    for (const OqMatrix* m : ms)
        if (m) b += m->cls ? m->cls->table_bytes : 8.0 * (double)m->local_rows * (double)m->cols;
    if (p->gf11_form == OQ_GF11_FFT) b += 8.0 * (double)p->Ghat.n + 16.0 * ((double)p->Rhat.n + (double)p->That.n) / 2;
    // vectors: state in, derivative out, properties, forcing vectors
    b += 8.0 * (2.0 * (double)p->nstate + 4.0 * p->nfl + (double)p->nf + 6.0 * p->ne);
    *bytes = b;
    return 0;
}

int oq_gemv(const OqMatrix* A, const double* x, double* y, int accumulate)
{
    OQ_CHECK(A && x && y, "NULL argument");
    OQ_TRY(enter());
    if (A->local_rows == 0) return 0;
    if (A->cls) {                                       // class form: multiply from the table
        DevBuf<double> cx, cy;
        OQ_TRY(cx.upload(x, A->cols));
        OQ_TRY(cy.alloc(A->local_rows));
        if (accumulate) OQ_CUDA(cudaMemcpy(cy.p, y, A->local_rows * sizeof(double), cudaMemcpyHostToDevice));
        const PeerWait none{nullptr, nullptr, 1, 0, 0};
        OQ_TRY(class_matvec(A, cx.p, 0, accumulate ? cy.p : nullptr, cy.p, none, nullptr, 0));
        OQ_CUDA(cudaMemcpy(y, cy.p, A->local_rows * sizeof(double), cudaMemcpyDeviceToHost));
        return 0;
    }
    DevBuf<double> dx, dy, partial;
    DevBuf<unsigned> counters;
    DevBuf<unsigned long long> ticket;
    OQ_TRY(ticket.alloc(2)); OQ_TRY(ticket.zero());
    OQ_TRY(dx.alloc(A->ld + kMvMaxSeg)); OQ_TRY(dx.zero());
    OQ_CUDA(cudaMemcpy(dx.p, x, A->cols * sizeof(double), cudaMemcpyHostToDevice));
    OQ_TRY(dy.alloc(A->local_rows));
    if (accumulate) OQ_CUDA(cudaMemcpy(dy.p, y, A->local_rows * sizeof(double), cudaMemcpyHostToDevice));
    size_t np = 0, nc = 0;
    gemv_scratch_sizes(A, &np, &nc);
    OQ_TRY(partial.alloc(np + 1));
    OQ_TRY(counters.alloc(nc + 1)); OQ_TRY(counters.zero());
    OQ_TRY(gemv_device(A, dx.p, accumulate ? dy.p : nullptr, dy.p, partial.p, counters.p, ticket.p, 0));
    OQ_CUDA(cudaMemcpy(y, dy.p, A->local_rows * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

}  // extern "C"
