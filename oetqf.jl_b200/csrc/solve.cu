// solve.cu -- device-resident adaptive integrators around the RHS of rhs.cu: Tsit5 (what the reference's tests
// use) and a variable-coefficient Adams-Bashforth-Moulton PECE of order 5 (the VCABM5 class the reference's
// example uses, examples/otf-with-mantle.jl:160-162): 2 RHS evaluations per step instead of 6.
//
// Takes the place of OrdinaryDiffEq's `solve(prob, Tsit5(); reltol, abstol, dtmax, dt, maxiters)` as the
// reference drives it (/root/reference/src/io.jl:128-130, test/tests.jl:11, examples/otf-with-mantle.jl:160-162):
// stage combinations, the scaled error norm, the PI step-size controller and the accept/reject copy all
// run on the GPU; the host reads one small control record per step.  Semantics follow OrdinaryDiffEq:
//   EEst = sqrt( mean_i ( err_i / (abstol + reltol*max(|uprev_i|,|u_i|)) )^2 ) over ALL state scalars,
//   PI controller beta1 = 7/50, beta2 = 2/25, gamma = 0.9, qmin = 0.2, qmax = 10, qoldinit = 1e-4.
#include <cmath>

#include "comm.cuh"
#include "problem.cuh"

namespace oq {

// Tsitouras 5(4) tableau (Tsitouras 2011; the coefficients OrdinaryDiffEq's Tsit5 uses)
__constant__ double cA[7][6] = {
    {0, 0, 0, 0, 0, 0},
    {0.161, 0, 0, 0, 0, 0},
    {-0.008480655492356989, 0.335480655492357, 0, 0, 0, 0},
    {2.8971530571054935, -6.359448489975075, 4.3622954328695815, 0, 0, 0},
    {5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525, 0, 0},
    {5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383, 0},
    {0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774}};
__constant__ double cBtilde[7] = {-0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995,
                                  -0.1447110071732629,     0.5823571654525552,     -0.45808210592918697,
                                  0.015151515151515152};

static const double hA[7][6] = {
    {0, 0, 0, 0, 0, 0},
    {0.161, 0, 0, 0, 0, 0},
    {-0.008480655492356989, 0.335480655492357, 0, 0, 0, 0},
    {2.8971530571054935, -6.359448489975075, 4.3622954328695815, 0, 0, 0},
    {5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525, 0, 0},
    {5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383, 0},
    {0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774}};

struct StepCtl {
    double t, dt, qold, eest, tstop, dtmax, reltol, abstol;
    double dt_last;
    double hdt[4];                   // sizes of the last accepted steps, newest first (multistep history grid)
    long long naccept, nreject, nrhs;
    int accepted, done, retcode, fixed;
    int nhist;                       // accepted steps recorded in the history so far (saturates at 4)
    int snap_slot;                   // ring slot the accept kernel copies this step's state into (-1: none)
    long long snap_stride;           // device-side snapshot ring: every snap_stride-th accepted step (0: off)
    long long nsnap;                 // snapshots written to the ring so far
};
static_assert(sizeof(StepCtl) <= 32 * sizeof(double), "ctl buffer too small");

struct Stages {
    const double* k[7];
};

// y = u + dt * sum_{j<nk} A[s][j] k_j
__global__ void __launch_bounds__(256)
stage_kernel(const double* __restrict__ u, Stages ks, int s, int nk, const StepCtl* __restrict__ ctl, size_t n,
             double* __restrict__ y)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double dt = ctl->dt;
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < 6; ++j)
        if (j < nk) acc = fma(cA[s][j], ks.k[j][i], acc);
    y[i] = fma(dt, acc, u[i]);
}

struct ErrArgs {
    const double *u, *unew;
    Stages ks;
    const double* wdev;              // multistep: 6 device weights (step size included); null: Tsit5's btilde * dt
    const StepCtl* ctl;
    size_t n;
    double* errpart;                 // [gridDim.x]
    unsigned long long* epochs;
    PeerTargets peers;
    WindowLayout wl;
};

// scaled error partial sums; the last block folds them (fixed order) and publishes this rank's sum
__global__ void __launch_bounds__(256) error_kernel(const __grid_constant__ ErrArgs a)
{
    __shared__ double red[8];
    __shared__ int last;
    if (a.ctl->done) return;          // steps enqueued past completion do nothing (identically on every rank)
    const double dt = a.ctl->dt, atol = a.ctl->abstol, rtol = a.ctl->reltol;
    double s = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (size_t)gridDim.x * blockDim.x) {
        double e = 0.0;
        if (a.wdev) {
#pragma unroll
            for (int j = 0; j < 6; ++j) e = fma(a.wdev[j], a.ks.k[j][i], e);
        } else {
#pragma unroll
            for (int j = 0; j < 7; ++j) e = fma(cBtilde[j], a.ks.k[j][i], e);
            e *= dt;
        }
        const double sk = atol + rtol * fmax(fabs(a.u[i]), fabs(a.unew[i]));
        const double q = e / sk;
        s = fma(q, q, s);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double b = 0.0;
        for (int w = 0; w < 8; ++w) b += red[w];
        a.errpart[blockIdx.x] = b;
        __threadfence();
        const unsigned long long prev = atomicAdd(&a.epochs[kEpBlocksR], 1ull);
        last = (prev == (unsigned long long)gridDim.x - 1ull);
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        a.epochs[kEpBlocksR] = 0ull;
        __threadfence();
        double tot = 0.0;
        for (unsigned b = 0; b < gridDim.x; ++b) tot += *(volatile double*)(a.errpart + b);
        const unsigned long long ep = a.epochs[kEpReduce];
        const size_t par = (size_t)(ep & 1ull);
        const int world = a.peers.world, rank = a.peers.rank;
        for (int r = 0; r < world; ++r) a.peers.base[r][a.wl.off_red + par * kMaxWorld + rank] = tot;
        if (world > 1) {
            __threadfence_system();
            for (int r = 0; r < world; ++r) {
                if (r == rank) continue;
                unsigned long long* f = reinterpret_cast<unsigned long long*>(a.peers.base[r] + a.wl.off_flags);
                publish_flag(f + kMaxWorld + rank, ep + 1ull);
            }
        }
        __threadfence();
        a.epochs[kEpReduce] = ep + 1ull;
    }
}

struct CtlArgs {
    StepCtl* ctl;
    const double* red_slots;             // [2][kMaxWorld]
    const unsigned long long* flags;     // [2][kMaxWorld]
    unsigned long long* epochs;
    int world, rank;
    double nglobal;
    int nrhs_inc;                        // RHS evaluations this step spent
};

constexpr int kSnapSlots = 32;       // ring capacity: two batches of kMaxBatch = 16 steps

__device__ void controller_body(CtlArgs& a);

// one thread: error norm -> accept/reject -> next dt (OrdinaryDiffEq's PI controller), then the snapshot decision
__global__ void controller_kernel(CtlArgs a)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    controller_body(a);
    StepCtl& c = *a.ctl;
    // device-side snapshot ring (the role of wsolve's FunctionCallingCallback, src/io.jl:51-58,128-130): every
    // snap_stride-th accepted step, and the completing one, is copied into the ring by the accept kernel
    c.snap_slot = -1;
    if (c.snap_stride > 0 && c.accepted && (c.naccept % c.snap_stride == 0 || c.done)) {
        c.snap_slot = (int)(c.nsnap % kSnapSlots);
        c.nsnap += 1;
    }
}

__device__ void controller_body(CtlArgs& a)
{
    StepCtl& c = *a.ctl;
    if (c.done) { c.accepted = 0; return; }
    const unsigned long long ep = *(volatile unsigned long long*)(a.epochs + kEpReduce);
    if (a.world > 1) wait_peers(a.flags + kMaxWorld, a.world, a.rank, ep, a.epochs + kEpError);
    if (*reinterpret_cast<volatile unsigned long long*>(a.epochs + kEpError)) {
        // a kernel of this step gave up waiting for a peer: its numbers are stale -- stop stepping, tell the host
        c.accepted = 0; c.done = 1; c.retcode = 3;
        return;
    }
    const size_t par = (size_t)((ep - 1ull) & 1ull);
    double tot = 0.0;
    for (int r = 0; r < a.world; ++r) tot += *(volatile const double*)(a.red_slots + par * kMaxWorld + r);
    const double eest = sqrt(tot / a.nglobal);
    c.eest = eest;
    c.nrhs += a.nrhs_inc;
    auto record = [&c]() {               // the accepted step joins the history grid of the multistep method
        c.hdt[3] = c.hdt[2]; c.hdt[2] = c.hdt[1]; c.hdt[1] = c.hdt[0]; c.hdt[0] = c.dt;
        if (c.nhist < 4) c.nhist += 1;
    };
    const double beta1 = 7.0 / 50.0, beta2 = 2.0 / 25.0, gamma = 0.9, qmin = 0.2, qmax = 10.0;
    if (c.fixed) {
        c.accepted = 1; c.naccept += 1; c.t += c.dt; c.dt_last = c.dt; record();
        if (c.t + c.dt > c.tstop) c.dt = c.tstop - c.t;
        if (c.t >= c.tstop - 4e-16 * fabs(c.tstop) || c.dt <= 0.0) c.done = 1;
        return;
    }
    if (!(eest == eest) || isinf(eest)) {           // NaN / Inf: treat as a failed step with maximal shrink
        c.accepted = 0; c.nreject += 1; c.dt *= qmin;
        if (c.dt < 1e-300) { c.done = 1; c.retcode = 2; }
        return;
    }
    const double q11 = pow(eest, beta1);
    double q = q11 / pow(c.qold, beta2);
    q = fmax(1.0 / qmax, fmin(1.0 / qmin, q / gamma));
    if (eest <= 1.0) {
        c.accepted = 1; c.naccept += 1;
        c.t += c.dt; c.dt_last = c.dt; record();
        c.qold = fmax(eest, 1e-4);
        double dtn = c.dt / q;
        if (dtn > c.dtmax) dtn = c.dtmax;
        if (c.t >= c.tstop - 4e-16 * fabs(c.tstop)) { c.done = 1; c.t = c.tstop; }
        else if (c.t + dtn > c.tstop) dtn = c.tstop - c.t;
        c.dt = dtn;
    } else {
        c.accepted = 0; c.nreject += 1;
        c.dt = c.dt / fmin(1.0 / qmin, q11 / gamma);
        if (c.dt <= fabs(c.t) * 2.2e-16) { c.done = 1; c.retcode = 2; }
    }
}

// on acceptance: u <- unew, k1 <- k7 (first-same-as-last)
// snapshot slot layout: u[n] | du[n] | t | accepted-step number
__device__ __forceinline__ void snap_store(const StepCtl* ctl, double* ring, size_t n, size_t i, double ui, double dui)
{
    const int slot = ctl->snap_slot;
    if (!ring || slot < 0) return;
    double* s = ring + (size_t)slot * (2 * n + 2);
    s[i] = ui;
    s[n + i] = dui;
    if (i == 0) { s[2 * n] = ctl->t; s[2 * n + 1] = (double)ctl->naccept; }
}

__global__ void __launch_bounds__(256)
accept_kernel(const StepCtl* __restrict__ ctl, size_t n, const double* __restrict__ unew,
              const double* __restrict__ k7, double* __restrict__ u, double* __restrict__ k1, double* ring)
{
    if (!ctl->accepted) return;      // (the controller clears `accepted` for steps enqueued past completion)
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const double a = unew[i], b = k7[i];
        u[i] = a; k1[i] = b;
        snap_store(ctl, ring, n, i, a, b);
    }
}

// on acceptance, multistep bookkeeping included: u <- unew and the derivative history shifts by one step
__global__ void __launch_bounds__(256)
accept_hist_kernel(const StepCtl* __restrict__ ctl, size_t n, const double* __restrict__ unew,
                   const double* __restrict__ fnew, double* __restrict__ u, double* __restrict__ f0,
                   double* __restrict__ h0, double* __restrict__ h1, double* __restrict__ h2, double* __restrict__ h3,
                   double* ring)
{
    if (!ctl->accepted) return;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double a = unew[i], b = fnew[i];
    u[i] = a;
    h3[i] = h2[i]; h2[i] = h1[i]; h1[i] = h0[i]; h0[i] = f0[i]; f0[i] = b;
    snap_store(ctl, ring, n, i, a, b);
}

// Weights of the variable-coefficient Adams formulas for the step [t_n, t_n + dt] on the grid of the last
// accepted steps: exact integrals of the Lagrange basis polynomials through the derivative samples
// (3-point Gauss-Legendre integrates the degree <= 5 basis exactly).  Node 0 is t_n + dt (the predicted
// derivative), nodes 1..5 are t_n, t_{n-1}, ..., t_{n-4}.
//   coef[0..3]   predictor (Adams-Bashforth on nodes 1-4, order 4)
//   coef[4..8]   corrector (Adams-Moulton on nodes 0-4, order 5): what the step advances with
//   coef[10..15] error estimate: order-6 formula (nodes 0-5) minus the corrector
// This is the Lagrange form of the divided-difference (g, phi, phi*) recurrences of Hairer, Norsett & Wanner
// III.5 that oracle/integrator.py follows -- two formulations of one formula.
constexpr int kCoefPred = 0, kCoefCorr = 4, kCoefErr = 10, kCoefLen = 16;

__global__ void abm_coef_kernel(const StepCtl* __restrict__ ctl, double* __restrict__ coef)
{
    __shared__ double w[15];
    if (ctl->done) return;
    const int t = threadIdx.x;
    const double dt = ctl->dt;
    double tau[6];
    tau[0] = dt; tau[1] = 0.0;
    tau[2] = -ctl->hdt[0]; tau[3] = tau[2] - ctl->hdt[1]; tau[4] = tau[3] - ctl->hdt[2]; tau[5] = tau[4] - ctl->hdt[3];
    if (t < 15) {
        int lo, hi, m;                                     // node set [lo, hi], basis node m
        if (t < 4) { lo = 1; hi = 4; m = 1 + t; }
        else if (t < 9) { lo = 0; hi = 4; m = t - 4; }
        else { lo = 0; hi = 5; m = t - 9; }
        double den = 1.0;
        for (int k = lo; k <= hi; ++k)
            if (k != m) den *= tau[m] - tau[k];
        const double gx[3] = {-0.7745966692414834, 0.0, 0.7745966692414834};
        const double gw[3] = {5.0 / 9.0, 8.0 / 9.0, 5.0 / 9.0};
        double acc = 0.0;
        for (int q = 0; q < 3; ++q) {
            const double x = 0.5 * dt * (1.0 + gx[q]);
            double num = 1.0;
            for (int k = lo; k <= hi; ++k)
                if (k != m) num *= x - tau[k];
            acc = fma(gw[q], num, acc);
        }
        w[t] = 0.5 * dt * acc / den;
    }
    __syncthreads();
    if (t < 4) coef[kCoefPred + t] = w[t];
    if (t >= 4 && t < 9) coef[kCoefCorr + (t - 4)] = w[t];
    if (t >= 9 && t < 15) coef[kCoefErr + (t - 9)] = w[t] - (t - 9 < 5 ? w[4 + (t - 9)] : 0.0);
}

static int launch_error_and_control(OqProblem* p, const Stages& ks, const double* wdev, int nrhs_inc)
{
    cudaStream_t st = p->stream;
    const size_t n = p->nstate;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    StepCtl* ctl = reinterpret_cast<StepCtl*>(p->ctl.p);
    ErrArgs ea{};
    ea.u = p->u.p; ea.unew = p->unew.p; ea.ks = ks; ea.wdev = wdev; ea.ctl = ctl; ea.n = n; ea.errpart = p->errpart.p;
    ea.epochs = p->epochs; ea.peers = comm_targets(p); ea.wl = p->wl;
    unsigned eblocks = blocks < 512 ? blocks : 512;
    error_kernel<<<eblocks, 256, 0, st>>>(ea);
    OQ_LAUNCHED();
    CtlArgs ca{ctl, p->red_slots, p->flags, p->epochs, p->world, p->rank, (double)p->nstate_global, nrhs_inc};
    controller_kernel<<<1, 32, 0, st>>>(ca);
    OQ_LAUNCHED();
    return 0;
}

static int launch_accept(OqProblem* p, bool with_history)
{
    const size_t n = p->nstate;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    const StepCtl* ctl = reinterpret_cast<const StepCtl*>(p->ctl.p);
    if (with_history)
        accept_hist_kernel<<<blocks, 256, 0, p->stream>>>(ctl, n, p->unew.p, p->k[6].p, p->u.p, p->k[0].p,
                                                          p->hist[0].p, p->hist[1].p, p->hist[2].p, p->hist[3].p,
                                                          p->snap_ring.p);
    else
        accept_kernel<<<blocks, 256, 0, p->stream>>>(ctl, n, p->unew.p, p->k[6].p, p->u.p, p->k[0].p, p->snap_ring.p);
    OQ_LAUNCHED();
    return 0;
}

// one predictor-evaluate-corrector-evaluate step of the order-5 Adams pair; f(t_n) is k[0], older samples hist[]
static int enqueue_abm_step(OqProblem* p)
{
    StepCtl* ctl = reinterpret_cast<StepCtl*>(p->ctl.p);
    double* coef = p->abm_coef.p;
    abm_coef_kernel<<<1, 32, 0, p->stream>>>(ctl, coef);
    OQ_LAUNCHED();
    StageSpec pred;                  // P: utmp = u + sum wp_m f_{n-m};  E: k[1] = f(utmp)
    pred.nk = 4; pred.u = p->u.p; pred.adev = coef + kCoefPred; pred.done = &ctl->done;
    pred.k[0] = p->k[0].p; pred.k[1] = p->hist[0].p; pred.k[2] = p->hist[1].p; pred.k[3] = p->hist[2].p;
    OQ_TRY(rhs_device(p, p->utmp.p, p->k[1].p, &pred));
    StageSpec corr;                  // C: unew = u + wc_0 f(utmp) + sum wc_m f_{n-m+1};  E: k[6] = f(unew)
    corr.nk = 5; corr.u = p->u.p; corr.adev = coef + kCoefCorr; corr.done = &ctl->done;
    corr.k[0] = p->k[1].p; corr.k[1] = p->k[0].p; corr.k[2] = p->hist[0].p; corr.k[3] = p->hist[1].p;
    corr.k[4] = p->hist[2].p;
    OQ_TRY(rhs_device(p, p->unew.p, p->k[6].p, &corr));
    Stages ks{};
    ks.k[0] = p->k[1].p; ks.k[1] = p->k[0].p;
    for (int j = 0; j < 4; ++j) ks.k[2 + j] = p->hist[j].p;
    ks.k[6] = p->k[0].p;             // unused
    OQ_TRY(launch_error_and_control(p, ks, coef + kCoefErr, 2));
    return launch_accept(p, true);
}

static int enqueue_step(OqProblem* p, bool with_history)
{
    StepCtl* ctl = reinterpret_cast<StepCtl*>(p->ctl.p);
    Stages ks;
    for (int j = 0; j < 7; ++j) ks.k[j] = p->k[j].p;
    for (int s = 1; s <= 6; ++s) {
        // stage combination y = u + dt Σ a_sj k_j is fused into the forcing kernel of the RHS evaluation
        double* y = s < 6 ? p->utmp.p : p->unew.p;
        StageSpec sp;
        sp.nk = s; sp.u = p->u.p; sp.dt = &ctl->dt; sp.done = &ctl->done;
        for (int j = 0; j < s; ++j) { sp.k[j] = p->k[j].p; sp.a[j] = hA[s][j]; }
        OQ_TRY(rhs_device(p, y, p->k[s].p, &sp));
    }
    OQ_TRY(launch_error_and_control(p, ks, nullptr, 6));
    return launch_accept(p, with_history);
}

}  // namespace oq

using namespace oq;

extern "C" int oq_solve(OqProblem* p, double t0, const OqSolveOptions* o, int64_t stride, OqSnapshotFn fn, void* user,
                        OqSolveStats* stats)
{
    OQ_CHECK(p && o, "NULL argument");
    OQ_CHECK(o->algorithm == OQ_ALG_TSIT5 || o->algorithm == OQ_ALG_VCABM5,
             "unknown algorithm (0 = Tsit5, 1 = VCABM5)");
    const bool multistep = o->algorithm == OQ_ALG_VCABM5;
    OQ_CHECK(o->tstop > t0, "tstop must be greater than t0");
    OQ_CHECK(o->fixed_dt || (o->reltol > 0 && o->abstol > 0), "tolerances must be positive");
    OQ_TRY(enter());
    if (stride < 1) stride = 1;
    comm_clear_error(p);
    StepCtl h{};
    h.t = t0; h.tstop = o->tstop;
    h.dtmax = o->dtmax > 0 ? o->dtmax : (o->tstop - t0);
    h.dt = o->dt0 > 0 ? o->dt0 : 1e-6 * (o->tstop - t0);
    if (h.dt > h.dtmax) h.dt = h.dtmax;
    if (t0 + h.dt > o->tstop) h.dt = o->tstop - t0;
    h.qold = 1e-4; h.reltol = o->reltol; h.abstol = o->abstol; h.fixed = o->fixed_dt;
    // Snapshot delivery.  Synchronous (default): the host copies the state and calls fn between batches, and fn's
    // return value stops the run at exactly that step.  Asynchronous (async_snapshots != 0, what wsolve uses): the
    // device copies every stride-th accepted state into a ring in HBM, a second stream drains the ring into
    // page-locked memory and fn runs on the host while the next batch of steps is already executing -- the
    // integration never waits for the callback or the disk; a stop request takes effect within one batch.
    const bool async = fn && o->async_snapshots != 0;
    h.snap_slot = -1; h.snap_stride = async ? stride : 0; h.nsnap = 0;
    const size_t slot_len = 2 * p->nstate + 2;
    struct HostRing {
        double* host = nullptr;
        cudaStream_t copy = nullptr;
        cudaEvent_t ev = nullptr;
        ~HostRing() { if (host) cudaFreeHost(host); if (copy) cudaStreamDestroy(copy); if (ev) cudaEventDestroy(ev); }
    } hr;
    if (async) {
        if (p->snap_ring.n != (size_t)kSnapSlots * slot_len) OQ_TRY(p->snap_ring.alloc((size_t)kSnapSlots * slot_len));
        OQ_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&hr.host), (size_t)kSnapSlots * slot_len * sizeof(double), cudaHostAllocDefault));
        OQ_CUDA(cudaStreamCreateWithFlags(&hr.copy, cudaStreamNonBlocking));
        OQ_CUDA(cudaEventCreateWithFlags(&hr.ev, cudaEventDisableTiming));
    } else {
        p->snap_ring.release();          // the accept kernels test the pointer
    }
    if (multistep) {
        for (int i = 0; i < 4; ++i)
            if (!p->hist[i].p) { OQ_TRY(p->hist[i].alloc(p->nstate + 2)); OQ_TRY(p->hist[i].zero()); }
        if (!p->abm_coef.p) { OQ_TRY(p->abm_coef.alloc(kCoefLen)); OQ_TRY(p->abm_coef.zero()); }
    }
    OQ_CUDA(cudaMemcpyAsync(p->ctl.p, &h, sizeof(h), cudaMemcpyHostToDevice, p->stream));
    // k1 = f(u0)
    OQ_TRY(rhs_device(p, p->u.p, p->k[0].p));
    OQ_CUDA(cudaStreamSynchronize(p->stream));

    // host mirrors for snapshots
    std::vector<std::vector<double>> hu(p->nparts), hdu(p->nparts);
    std::vector<const double*> pu(p->nparts), pdu(p->nparts);
    std::vector<double*> wu(p->nparts), wdu(p->nparts);
    for (int i = 0; i < p->nparts; ++i) {
        hu[i].resize(p->part_len[i] + 1); hdu[i].resize(p->part_len[i] + 1);
        pu[i] = wu[i] = hu[i].data(); pdu[i] = wdu[i] = hdu[i].data();
    }
    auto snapshot = [&](double t, int64_t step) -> int {
        if (!fn) return 0;
        if (oq_state_get(p, wu.data())) return -1;
        if (oq_state_get_du(p, wdu.data())) return -1;
        return fn(user, t, step, pu.data(), pdu.data());
    };
    int stop = snapshot(t0, 0);
    if (stop < 0) return 1;

    // capture one step into a CUDA graph (all step-dependent scalars live in device memory); the multistep
    // method has two: the Runge-Kutta step that builds its history, and the Adams step
    struct StepGraph {
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        int64_t launches = 0;
        ~StepGraph() { if (exec) cudaGraphExecDestroy(exec); if (graph) cudaGraphDestroy(graph); }
    } rk, abm;
    auto capture = [&](StepGraph& g, bool adams) -> int {
        OQ_CUDA(cudaStreamBeginCapture(p->stream, cudaStreamCaptureModeThreadLocal));
        const int64_t before = g_launches.load();
        const int r = adams ? enqueue_abm_step(p) : enqueue_step(p, multistep);
        g.launches = g_launches.load() - before;
        g_launches.fetch_sub(g.launches);   // capture enqueued nothing yet
        const cudaError_t e = cudaStreamEndCapture(p->stream, &g.graph);
        if (r) return r;
        OQ_CUDA(e);
        OQ_CUDA(cudaGraphInstantiate(&g.exec, g.graph, 0));
        return 0;
    };
    int rc = capture(rk, false);
    if (rc) return rc;
    if (multistep) { rc = capture(abm, true); if (rc) return rc; }
    cudaError_t ce = cudaSuccess;

    const int64_t maxiters = o->maxiters > 0 ? o->maxiters : 1000000;
    int64_t iters = 0;
    const int64_t kMaxBatch = 16;    // steps enqueued after completion exit at their first instruction (ctl.done)
    // asynchronous delivery: snapshots [pend_lo, pend_hi) are on their way to (or already in) the pinned ring
    long long pend_lo = 0, pend_hi = 0;
    auto deliver_pending = [&]() -> int {
        if (pend_hi == pend_lo) return 0;
        if (cudaStreamSynchronize(hr.copy) != cudaSuccess) return -1;
        int st = 0;
        for (long long q = pend_lo; q < pend_hi && st == 0; ++q) {
            const double* slot = hr.host + (size_t)(q % kSnapSlots) * slot_len;
            for (int i = 0; i < p->nparts; ++i) { pu[i] = slot + p->part_off[i]; pdu[i] = slot + p->nstate + p->part_off[i]; }
            st = fn(user, slot[2 * p->nstate], (int64_t)slot[2 * p->nstate + 1], pu.data(), pdu.data());
        }
        pend_lo = pend_hi;
        return st;
    };
    while (!stop && !h.done && iters < maxiters) {
        // Launch as many steps as can pass before the next snapshot is due (all decisions are taken on the
        // device; steps enqueued after completion are no-ops for the state), then read the control record once.
        // (the batch size must be a pure function of the control record: every rank of a multi-GPU run has to
        // enqueue exactly the same sequence of kernels, or the peers' epoch flags would never match)
        int64_t batch = (fn && !async) ? stride - (h.naccept % stride) : kMaxBatch;
        if (batch > kMaxBatch) batch = kMaxBatch;
        if (batch > maxiters - iters) batch = maxiters - iters;
        // the Adams pair needs four accepted steps of history; until then Tsit5 steps, one at a time
        const bool adams = multistep && h.nhist >= 4;
        if (multistep && !adams) batch = 1;
        if (batch < 1) batch = 1;
        const StepGraph& g = adams ? abm : rk;
        for (int64_t b = 0; b < batch && ce == cudaSuccess; ++b) {
            ce = cudaGraphLaunch(g.exec, p->stream);
            g_launches.fetch_add(g.launches);
        }
        if (ce != cudaSuccess) { rc = fail("cudaGraphLaunch: %s", cudaGetErrorString(ce)); break; }
        const int64_t acc_before = h.naccept;
        const long long snap_before = h.nsnap;
        ce = cudaMemcpyAsync(&h, p->ctl.p, sizeof(h), cudaMemcpyDeviceToHost, p->stream);
        if (async && ce == cudaSuccess) {
            ce = cudaEventRecord(hr.ev, p->stream);
            // the callbacks of the previous batch run on the host while this batch executes on the device
            const int st = deliver_pending();
            if (st < 0) { rc = fail("snapshot copy failed"); break; }
            if (st > 0) stop = st;
        }
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(p->stream);
        if (ce != cudaSuccess) { rc = fail("step failed: %s", cudaGetErrorString(ce)); break; }
        iters += batch;
        if (h.retcode == 3 || comm_check_error(p, "oq_solve")) {
            rc = fail("oq_solve: timed out waiting for a peer rank at t = %g (state not advanced past it; "
                      "OQ_PEER_TIMEOUT_S sets the limit)", h.t);
            break;
        }
        if (async) {
            // start draining what this batch wrote (the copy stream waits for the batch, not the other way round)
            if (h.nsnap > snap_before) {
                ce = cudaStreamWaitEvent(hr.copy, hr.ev, 0);
                for (long long q = snap_before; q < h.nsnap && ce == cudaSuccess; ++q) {
                    const size_t off = (size_t)(q % kSnapSlots) * slot_len;
                    ce = cudaMemcpyAsync(hr.host + off, p->snap_ring.p + off, slot_len * sizeof(double),
                                         cudaMemcpyDeviceToHost, hr.copy);
                }
                if (ce != cudaSuccess) { rc = fail("snapshot copy failed: %s", cudaGetErrorString(ce)); break; }
                pend_lo = snap_before; pend_hi = h.nsnap;
            }
        } else if (h.naccept > acc_before && (h.naccept % stride == 0 || h.done)) {
            stop = snapshot(h.t, h.naccept);
            if (stop < 0) { rc = 1; break; }
        }
    }
    if (async && !rc) {
        const int st = deliver_pending();
        if (st < 0) rc = fail("snapshot copy failed");
        else if (st > 0 && !stop) stop = st;
    }
    if (!rc && comm_check_error(p, "oq_solve")) rc = 1;
    if (stats) {
        stats->t = h.t; stats->dt_last = h.dt_last; stats->dt_next = h.dt;
        stats->naccept = h.naccept; stats->nreject = h.nreject; stats->nrhs = h.nrhs + 1;
        stats->retcode = h.retcode ? h.retcode : (h.done ? 0 : (stop > 0 ? 4 : (iters >= maxiters ? 1 : 0)));
    }
    return rc;
}
