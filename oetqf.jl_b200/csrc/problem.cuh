// problem.cuh -- the OqProblem handle: everything one rank needs to evaluate the ODE right-hand side
// of /root/reference/src/BEM/equation.jl:156-205 on its row shard, with the state resident in HBM.
#pragma once
#include "comm.cuh"
#include "common.cuh"

namespace oq {

enum ProblemKind : int { kFaultOnly = 0, kDilatancy = 1, kViscoelastic = 2 };

// One dense operand of the fused matvec: rows of `G` (row-major, leading dimension ld) times `x`.
struct MatOperand {
    const double* G = nullptr;
    size_t ld = 0;
    const double* x = nullptr;  // copy 0 of the forcing vector
    size_t x_stride = 0;        // distance to copy 1 (0: single copy)
    int cols = 0;
    int x_kind = 0;     // 0: x = v - vpl (columns are fault cells); 1: x = dϵ - dϵ0 (column p*ne + e is element e)
    int nseg = 0;       // number of column segments the row is split into
    int seg_len = 0;    // columns per segment (multiple of 512)
};

// Pointwise physics parameters on the device (local rows only).
struct FaultParams {
    const double *a, *b, *L, *sigma;                // [nfl]
    const double *tp, *epsd, *beta, *p0;            // dilatancy, [nfl] or null
    double eta, vpl, f0, v0;
};

struct MantleParams {
    const double *gamma, *npow;                     // [nlaws * nel] law-major
    int nlaws;
    double deps0[6];
};

// Views into one state-shaped vector (u, du or a stage) for this rank's rows.
struct StateView {
    double *v, *theta, *delta, *pr;                 // [nfl]   (pr: pore pressure, dilatancy only)
    double *eps, *sig;                              // [6*nel] at el + k*nel
};


}  // namespace oq

struct OqProblem {
    int kind = oq::kFaultOnly;
    int nx = 0, nxi = 0, nf = 0, ne = 0;           // global sizes
    int f0 = 0, f1 = 0, e0 = 0, e1 = 0;            // this rank's fault rows / mantle elements
    int nfl = 0, nel = 0;
    int gf11_form = OQ_GF11_DENSE;
    int nparts = 0;
    int part_len[5] = {0, 0, 0, 0, 0};
    size_t part_off[5] = {0, 0, 0, 0, 0};
    size_t nstate = 0;                              // local state length (sum of part_len)
    size_t nstate_global = 0;                       // global state length (for the RMS error norm)

    // borrowed matrices (must outlive the problem)
    const OqMatrix *g11 = nullptr, *g12 = nullptr, *g21 = nullptr, *g22 = nullptr;
    oq::DevBuf<double> st;                          // Toeplitz kernel [nx,nxi,nxi] for OQ_GF11_FFT

    // properties
    oq::DevBuf<double> props;                       // packed a,b,L,sigma,(tp,eps,beta,p0),(gamma,n)
    oq::FaultParams fp{};
    oq::MantleParams mp{};

    // forcing vectors (GLOBAL length, zero-padded to the matrices' leading dimension).  They live in the
    // peer-visible window so that other ranks can store their slices straight into them (comm.cu);
    // two copies alternate with the parity of the device-side evaluation counter, so a fast rank never
    // overwrites what a slow one is still reading.
    oq::DevBuf<double> window;
    oq::WindowLayout wl{};
    double* relv = nullptr;                         // copy 0 of v - vpl, [relv_len]; copy 1 follows
    double* reldeps = nullptr;                      // copy 0 of dϵ - dϵ0 at p*ne + e, [reldeps_len]
    double* red_slots = nullptr;                    // [2][kMaxWorld] partial sums of the step error norm
    unsigned long long* flags = nullptr;            // [2][kMaxWorld] arrival epochs: forcing, error norm
    unsigned long long* epochs = nullptr;           // local counters, see comm.cuh
    unsigned long long* err_host = nullptr;         // page-locked, device-mapped word: 1 after a peer-wait timeout
    // matvec scratch
    oq::DevBuf<double> partial_f, partial_m;        // [rows * nsegTotal]
    oq::DevBuf<unsigned> counters;                  // [row blocks fault + row blocks mantle]
    oq::DevBuf<unsigned long long> ticket;          // matvec pass counter + finished-CTA counter (traversal direction)
    oq::DevBuf<double> dtau0;                       // Toeplitz-form / class-form traction rate [nfl]
    oq::DevBuf<double> dsig0;                       // class-form stress rate [6*nel] when a dense mantle operand follows
    // FFT form (toeplitz_fft.cuh): transform length, local receiver-row range, spectrum and work arrays
    int fftN = 0, fj0 = 0, fnj = 0;
    oq::DevBuf<double> Ghat, Rhat, That, Wtw;      // Wtw: N/2 complex twiddles
    int nseg_f = 0, nseg_m = 0;
    oq::MatOperand opf[2], opm[2];

    // resident state + integrator storage: u, du(k1), k2..k7, utmp, unew
    oq::DevBuf<double> u, k[7], utmp, unew;
    oq::DevBuf<double> hist[4], abm_coef;           // multistep integrator: f(t_{n-1..n-4}), per-step weights
    oq::DevBuf<double> errpart;                     // per-block partial sums of the error norm
    oq::DevBuf<double> ctl;                         // device-side controller record (StepCtl)
    oq::DevBuf<double> snap_ring;                   // device-side snapshot ring of oq_solve (async_snapshots)

    cudaStream_t stream = nullptr;
    cudaGraphExec_t rhs_graph = nullptr;            // TWO resident RHS evaluations (u -> k1), captured once (the
                                                    // traversal direction of the matvec alternates between them)
    int64_t rhs_graph_launches = 0;                 // kernels per replay
    unsigned mv_seq = 0;                            // evaluations enqueued so far (parity = traversal direction)

    // optional per-launch timing of the matvec (bench.py's roofline line)
    bool prof_on = false;
    std::vector<cudaEvent_t> prof_ev;               // pairs (start, stop)
    size_t prof_used = 0;

    // multi-GPU
    int rank = 0, world = 1;
    oq::PeerWindow* peers = nullptr;

    ~OqProblem();
};

namespace oq {

// Runge-Kutta stage / Adams predictor-corrector combination fused into the forcing kernel:
//   y = u + dt * sum_{j<nk} a[j] k[j]
struct StageSpec {
    int nk = 0;
    const double* u = nullptr;
    const double* k[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    double a[6] = {0, 0, 0, 0, 0, 0};
    const double* adev = nullptr;    // device coefficients replacing a[] (the multistep weights change every step)
    const double* dt = nullptr;      // device scalar (the controller's current step); null: factor 1
    const int* done = nullptr;       // device flag: the integration is complete, the evaluation is skipped
};

StateView view_of(const OqProblem* p, double* base);
// one full RHS evaluation on the device: du = f(uin).  Enqueues on p->stream.  With a stage spec, uin is first
// filled with the stage state (each thread combines exactly the entries its forcing needs).
int rhs_device(OqProblem* p, const double* uin, double* du, const StageSpec* stage = nullptr);

}  // namespace oq
