/*
 * oetqf_b200.h -- C ABI of liboetqf_b200.so: the B200 (sm_100a) implementation of Oetqf.jl's two
 * data-parallel hot paths (Green's-function assembly; the ODE right-hand side and its integrator).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; oq_last_error() gives the message
 *     (thread-local).  Nothing throws or aborts across this boundary.
 *   - plain pointers and sizes only.  "host" pointers are ordinary process memory and are never
 *     retained after the call returns; handles (OqMatrix, OqProblem) own all device memory and are
 *     released by their *_destroy function.
 *   - host arrays use the reference's layouts (Julia column-major): fault fields [nx, nxi] at
 *     i + j*nx; mantle fields [ne, 6] at e + k*ne with k = xx,xy,xz,yy,yz,zz.
 *   - handles are not re-entrant; calls may come from any host thread (the device is re-selected
 *     on every call).  One process drives one GPU (one rank per GPU under torchrun / MPI / Distributed).
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails.
 *
 * Each entry point cites the reference interface it replaces (paths relative to the Oetqf.jl tree).
 */
#ifndef OETQF_B200_H
#define OETQF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OQ_ABI_VERSION 2

/* FaultType, src/BEM/GF.jl:3-5 */
enum { OQ_STRIKE_SLIP = 0, OQ_DIP_SLIP = 1 };

/* RectOkadaMesh, src/BEM/mesh.jl:5-21.  ax0/ax1[nx], axi0/axi1[nxi] are the cell edges
 * (mesh.ax[i][1], mesh.ax[i][2], mesh.aξ[j][1], mesh.aξ[j][2]). */
typedef struct OqFaultMesh {
    int32_t nx, nxi;
    const double *x, *ax0, *ax1;            /* [nx]  */
    const double *xi, *axi0, *axi1, *y, *z; /* [nxi] */
    double dx, dxi, dep, dip;               /* dip in degrees */
} OqFaultMesh;

/* BEMHex8Mesh, src/BEM/mesh.jl:58-72 (θ is stored by the reference but never used: GF.jl:219,281) */
typedef struct OqHex8Mesh {
    int32_t n;
    const double *cx, *cy, *cz, *qx, *qy, *qz, *dx, *dy, *dz; /* [n] */
} OqHex8Mesh;

/* QuadratureType, src/BEM/GF.jl:298,325-328: coords[3*nq] in [-1,1]^3, weights[nq] */
typedef struct OqQuadrature {
    int32_t nq;
    const double *coords, *weights;
} OqQuadrature;

/* ------------------------------------------------------------------ runtime */
int oq_abi_version(void);
const char *oq_last_error(void);
/* Select the CUDA device this process drives (default 0).  Fails if no sm_100 device is present. */
int oq_init(int device);
int oq_device_count(int *count);
/* Number of kernels of this library launched by the calling process so far (for bench accounting). */
int64_t oq_kernel_launch_count(void);
/* Measured fp64 FMA peak of the current device in flop/s (DFMA microbenchmark, CUDA-event timed). */
int oq_measure_fp64_peak(double *flops_per_s);
/* Measured HBM copy bandwidth (read+write bytes / s) over a buffer of `bytes` bytes. */
int oq_measure_hbm_copy(size_t bytes, double *bytes_per_s);
/* Page-lock and map a host array (cudaHostRegister) so that oq_rhs reads / writes it from the kernels themselves
 * instead of staging copies (the arrays the integrator hands to `ode`, src/BEM/equation.jl:156-205).  The array
 * must stay allocated until oq_host_unregister; memory from cudaHostAlloc / a pinned torch tensor needs neither. */
int oq_host_register(void *ptr, size_t bytes);
int oq_host_unregister(void *ptr);

/* ------------------------------------------------------------------ Green's functions, host output
 * Drop-in for the four `stress_greens_function` methods.  Output arrays are column-major exactly as
 * the Julia arrays returned by the reference.  kernel_ms (may be NULL) receives the device time of
 * the assembly kernel(s) alone (CUDA events). */

/* src/BEM/GF.jl:31-71.  out: fourier == 0 -> double[nx*nxi*nxi]; fourier != 0 -> interleaved
 * complex double[2*nx*nxi*nxi] (rfft of length 2nx-1 along strike of the even extension). */
int oq_gf_fault_fault(const OqFaultMesh *mf, double lambda, double mu, int ftype, int fourier,
                      int nrept, double buffer_ratio, double *out, double *kernel_ms);

/* src/BEM/GF.jl:123-174.  out: double[6*ne * nx*nxi] */
int oq_gf_fault_mantle(const OqFaultMesh *mf, const OqHex8Mesh *ma, const OqQuadrature *quad,
                       double lambda, double mu, int ftype, int nrept, double buffer_ratio,
                       double *out, double *kernel_ms);

/* src/BEM/GF.jl:194-227.  out: double[nx*nxi * 6*ne] */
int oq_gf_mantle_fault(const OqHex8Mesh *ma, const OqFaultMesh *mf, double lambda, double mu, int ftype,
                       double *out, double *kernel_ms);

/* src/BEM/GF.jl:250-296 (without the O(n^3) eigvals print of :291-294; the host side offers an Arnoldi
 * estimate of the largest real part built on oq_gemv instead).
 * out: double[6*ne * 6*ne] */
int oq_gf_mantle_mantle(const OqHex8Mesh *ma, const OqQuadrature *quad, double lambda, double mu,
                        double *out, double *kernel_ms);

/* Direct evaluation of the two kernels of GeoGreensFunctions.jl as the reference calls them
 * (GF.jl:49-54 / GF.jl:215-221), batched: n receivers against one source.  Gradient-only for dc3d:
 * out9[n*9] = (uxx,uyx,uzx,uxy,uyy,uzy,uxz,uyz,uzz), i.e. entries 4..12 of dc3d's 12-vector. */
int oq_dc3d_gradient(int n, const double *x, const double *y, const double *z, double alpha, double dep,
                     double dip, double al1, double al2, double aw1, double aw2, int ftype, double *out9);
/* out6[n*6] = stress (xx,xy,xz,yy,yz,zz) for eigenstrain eps[6] of one cuboid (θ = 0). */
int oq_stress_vol_hex8(int n, const double *x, const double *y, const double *z,
                       double qx, double qy, double qz, double dx, double dy, double dz,
                       const double *eps6, double mu, double nu, double *out6);

/* ------------------------------------------------------------------ device-resident matrices
 * A row shard of a Green's matrix living in HBM (row-major, leading dimension padded to 128 B).
 * Fault rows are sharded by contiguous vec index f = i + j*nx in [row_begin,row_end);
 * mantle rows are sharded by ELEMENT e in [row_begin,row_end) and carry all six components. */
typedef struct OqMatrix OqMatrix;

enum { OQ_ROWS_FAULT = 0, OQ_ROWS_MANTLE = 1 };

/* dense fault<-fault, expanded from the Toeplitz kernel: G[(i,j),(k,l)] = st[|i-k|,j,l]
 * (the construction of test/BEM/tests.jl:46-49).  rows [row_begin,row_end) of nx*nxi. */
int oq_matrix_fault_fault(const OqFaultMesh *mf, double lambda, double mu, int ftype, int nrept,
                          double buffer_ratio, int row_begin, int row_end, OqMatrix **out);
/* same, from a Toeplitz kernel st[nx,nxi,nxi] already on the host (e.g. irfft of a cached Fourier-form gf11) */
int oq_matrix_from_toeplitz(const double *st, int nx, int nxi, int row_begin, int row_end, OqMatrix **out);
/* mantle<-fault (the reference's gf12, GF.jl:123-174); rows = elements [e_begin,e_end) x 6 */
int oq_matrix_fault_mantle(const OqFaultMesh *mf, const OqHex8Mesh *ma, const OqQuadrature *quad,
                           double lambda, double mu, int ftype, int nrept, double buffer_ratio,
                           int e_begin, int e_end, OqMatrix **out);
/* fault<-mantle (gf21, GF.jl:194-227); rows = fault cells [row_begin,row_end) */
int oq_matrix_mantle_fault(const OqHex8Mesh *ma, const OqFaultMesh *mf, double lambda, double mu, int ftype,
                           int row_begin, int row_end, OqMatrix **out);
/* mantle<-mantle (gf22, GF.jl:250-296); rows = elements [e_begin,e_end) x 6 */
int oq_matrix_mantle_mantle(const OqHex8Mesh *ma, const OqQuadrature *quad, double lambda, double mu,
                            int e_begin, int e_end, OqMatrix **out);
/* The same three mantle operands kept in CLASS FORM instead of dense storage (csrc/classmat.cuh): the table of the
 * distinct 6x1 / 1x6 / 6x6 kernels of the matrix plus the maps from a (receiver, source) pair to its translation
 * class.  The handle is accepted wherever the dense one is -- oq_problem_create_viscoelastic (any mix of dense and
 * class-form operands), oq_gemv (the matvecmul! slot, pref.jl:15-21), oq_matrix_to_host / _rows_to_host (which expand
 * on request) -- and multiplies straight from the table: equation.jl:201-203 without streaming (6 N_e)^2 doubles.
 * The entries are those of the dense builders above, bit for bit (same table, same representatives).  Not in the
 * reference, which keeps these operands dense (GF.jl:123-296); it exploits the same invariance for the fault only
 * (GF.jl:31-71).  Returns an error (and no handle) when the mesh has no translation structure to exploit -- fewer
 * than 4 pairs per class -- or when one (y,z) slab of the table exceeds shared memory: keep the dense form then.
 * A class-form handle owns the scratch of its evaluation (the forcing vector in group order, the partial sums): one
 * evaluation at a time per handle (evaluations of one OqProblem are stream-ordered; do not share one class-form
 * operand between problems that run concurrently). */
int oq_matrix_fault_mantle_classes(const OqFaultMesh *mf, const OqHex8Mesh *ma, const OqQuadrature *quad, double lambda,
                                   double mu, int ftype, int nrept, double buffer_ratio, int e_begin, int e_end,
                                   OqMatrix **out);
int oq_matrix_mantle_fault_classes(const OqHex8Mesh *ma, const OqFaultMesh *mf, double lambda, double mu, int ftype,
                                   int row_begin, int row_end, OqMatrix **out);
int oq_matrix_mantle_mantle_classes(const OqHex8Mesh *ma, const OqQuadrature *quad, double lambda, double mu,
                                    int e_begin, int e_end, OqMatrix **out);
/* Host-only view of the plan behind oq_matrix_mantle_mantle_classes for the receivers [e_begin, e_end) (no device is
 * touched): out8 = { x classes, (y,z) classes, worthwhile (>= 4 pairs per class over the whole mesh), diagonal fast
 * path available (x class = function of the position difference), receiver runs of the diagonal kernel, x positions,
 * largest source group, (y,z) classes of the shard's receivers }. */
int oq_class_form_plan(const OqHex8Mesh *ma, int e_begin, int e_end, long long *out8);
/* Host-only self check of the sliding-window plan of a fault <-> mantle class operand (which = 1: mantle -> fault,
 * receivers = fault cells [begin, end); 2: fault -> mantle, receivers = elements [begin, end)): every (receiver, source)
 * pair reached through the plan's maps must get the class the general maps give it.  out6 = { plan found, residues,
 * runs, pairs checked, mismatches, receivers the plan does not reach }. */
int oq_class_window_check(const OqHex8Mesh *ma, const OqFaultMesh *mf, int which, int begin, int end, long long *out6);
/* form: 0 dense, 1 class form; device_bytes: HBM held by the operand (dense shard, or class table + maps) */
int oq_matrix_form(const OqMatrix *a, int *form, double *device_bytes);
/* Upload a user-supplied column-major m x n host matrix (e.g. one loaded from the reference's HDF5
 * cache, examples/otf-with-mantle.jl:39-56).  row_kind says how [row_begin,row_end) is interpreted. */
int oq_matrix_from_host(const double *a_colmajor, int m, int n, int row_kind, int row_begin, int row_end,
                        OqMatrix **out);
/* Download the shard as a column-major (local_rows x n) host array; local row order is
 * fault: f - row_begin ; mantle: (e - e_begin) + k*(e_end - e_begin). */
int oq_matrix_to_host(const OqMatrix *a, double *out_colmajor);
/* Download LOCAL rows [local_begin, local_end) of the shard as a ROW-major ((local_end-local_begin) x n) host
 * array (one strided device->host copy, no transposition): lets a caller inspect a window of a shard that is
 * itself too large to duplicate (parity checks of 100 GB-class matrices; row i of the output is what the
 * reference holds in st[row, :], GF.jl:123-296). */
int oq_matrix_rows_to_host(const OqMatrix *a, int local_begin, int local_end, double *out_rowmajor);
int oq_matrix_shape(const OqMatrix *a, int *local_rows, int *cols, int *global_rows);
/* device time (CUDA events) of the assembly kernel that filled this shard, in ms (0 for uploads) */
int oq_matrix_kernel_ms(const OqMatrix *a, double *ms);
/* How a hex8 shard (gf21, gf22) was assembled.  path: 0 = one thread per (receiver, source) pair, 1 = tiles of cells
 * sharing vertices, 2 = class tables -- the half-space is invariant under horizontal translation, so pairs with equal
 * (x_r - q_x, y_r - q_y, depths, sizes) share their 6 / 36 entries: the closed form of GF.jl:215-221 / :277-283 is
 * evaluated once per class (table_ms) and copied into the dense shard (expand_ms); the same invariance the reference
 * uses for gf11 (GF.jl:31-71).  -1: not a hex8 matrix.  OQ_HEX8 = pair | tile | classes forces a path. */
typedef struct OqAssemblyInfo {
    int path;
    int64_t pairs, unique_pairs;      /* pairs of the shard; closed-form evaluations actually made */
    double table_ms, expand_ms, kernel_ms;
} OqAssemblyInfo;
int oq_matrix_assembly_info(const OqMatrix *a, OqAssemblyInfo *info);
/* Host-only view of that class decomposition (no device needed): for n sample pairs (recv[k], src[k]) -- receivers are
 * mantle elements of [begin,end) when mf == NULL (gf22) or fault cells of [begin,end) (gf21) -- the receiver/source
 * whose coordinates stand for the pair in the x group and in the (y,z) group; counts[0..2] = classes of the x group,
 * of the (y,z) group (0, 0: the mesh has no exploitable structure) and pairs. */
int oq_hex8_pair_classes(const OqHex8Mesh *ma, const OqFaultMesh *mf, int begin, int end, int n, const int *recv,
                         const int *src, int *rep_recv_x, int *rep_src_x, int *rep_recv_yz, int *rep_src_yz,
                         long long *counts);
int oq_matrix_destroy(OqMatrix *a);

/* The `matvecmul!` backend slot, src/pref.jl:15-21 as used at src/BEM/equation.jl:201-203:
 * y = A*x (accumulate == 0, the 3-argument mul!) or y += A*x (accumulate != 0, the 5-argument form
 * with α = β = true).  x[n], y[local_rows] are host pointers. */
int oq_gemv(const OqMatrix *a, const double *x, double *y, int accumulate);

/* ------------------------------------------------------------------ properties (src/BEM/property.jl) */
/* RateStateQuasiDynamicProperty, property.jl:10-25; arrays [nx*nxi] */
typedef struct OqFaultProperty {
    const double *a, *b, *L, *sigma;
    double eta, vpl, f0, v0;
} OqFaultProperty;

/* PowerLawViscosityProperty / CompositePowerLawViscosityProperty, property.jl:34-48.
 * gamma, n: [nlaws*ne] (law-major); n holds "power - 1" as in the reference; deps0[6]. */
typedef struct OqMantleProperty {
    int32_t nlaws;
    const double *gamma, *n, *deps0;
} OqMantleProperty;

/* DilatancyProperty, property.jl:27-32; arrays [nx*nxi] */
typedef struct OqDilatancyProperty {
    const double *tp, *eps, *beta, *p0;
} OqDilatancyProperty;

/* ------------------------------------------------------------------ the ODE problem
 * Replaces `assemble` (src/BEM/equation.jl:81-154) and the in-place RHS `ode(du,u,p,t)`
 * (equation.jl:156-205).  State partitions, in the reference's ArrayPartition order:
 *   fault-only   : (v, θ, δ)            equation.jl:159-160
 *   dilatancy    : (v, θ, δ, 𝓅)         equation.jl:176-177
 *   viscoelastic : (v, θ, ϵ, σ, δ)      equation.jl:193-194
 */
typedef struct OqProblem OqProblem;

enum { OQ_GF11_DENSE = 0, OQ_GF11_FFT = 1 };

/* Fault-only problem.  Exactly one of g11 (dense shard, OQ_GF11_DENSE) or st_toeplitz (the real
 * [nx,nxi,nxi] kernel on the host, OQ_GF11_FFT: the reference's own algorithm, equation.jl:44-61,
 * evaluated on the device in its Toeplitz/Fourier form) is used.  dila may be NULL. */
int oq_problem_create_fault(int nx, int nxi, int gf11_form, const OqMatrix *g11, const double *st_toeplitz,
                            const OqFaultProperty *pf, const OqDilatancyProperty *dila, OqProblem **out);
/* Viscoelastic problem (fault + mantle).  g12: mantle<-fault, g21: fault<-mantle, g22: mantle<-mantle. */
int oq_problem_create_viscoelastic(int nx, int nxi, int ne, int gf11_form, const OqMatrix *g11,
                                   const double *st_toeplitz, const OqMatrix *g12, const OqMatrix *g21,
                                   const OqMatrix *g22, const OqFaultProperty *pf,
                                   const OqMantleProperty *pa, OqProblem **out);
int oq_problem_destroy(OqProblem *p);
/* number of state partitions and the length of each (local rows of this rank) */
int oq_problem_layout(const OqProblem *p, int *nparts, int *lengths /* [5] */);

/* "compat mode": the (du, u, p, t) call of OrdinaryDiffEq, host pointers in, host pointers out.
 * u_parts / du_parts: arrays of nparts host pointers in the partition order above.  Page-locked, mapped
 * partitions (cudaHostAlloc, oq_host_register) are read and written by the kernels directly over PCIe;
 * pageable ones are staged through device buffers.  Synchronous: du is complete on return. */
int oq_rhs(OqProblem *p, double t, const double *const *u_parts, double *const *du_parts);

/* "resident mode": state lives on the device; load / read it and evaluate the RHS in place. */
int oq_state_set(OqProblem *p, const double *const *u_parts);
int oq_state_get(const OqProblem *p, double *const *u_parts);
int oq_state_get_du(const OqProblem *p, double *const *du_parts); /* derivative at the current state */
/* nevals back-to-back device RHS evaluations on the resident state (for throughput measurement);
 * ms_total receives the CUDA-event time. */
int oq_rhs_resident(OqProblem *p, int nevals, double *ms_total);

/* Per-kernel timing of the dominant kernel (the fused matvec) inside the caller's timed region: when enabled,
 * every RHS evaluation brackets its matvec launch with CUDA events on the problem's stream.
 * oq_profile_read synchronises, returns the summed matvec time and launch count since the last read. */
int oq_profile_enable(OqProblem *p, int on);
int oq_profile_read(OqProblem *p, double *matvec_ms_total, int64_t *launches);
/* Algorithmic bytes one RHS evaluation streams from HBM on this rank (matrix shards + vectors). */
int oq_rhs_bytes(const OqProblem *p, double *bytes);

/* Adaptive integrator options (OrdinaryDiffEq semantics: examples/otf-with-mantle.jl:160-162). */
#define OQ_ALG_TSIT5 0
#define OQ_ALG_VCABM5 1 /* variable-coefficient Adams PECE, order 5, started with four Tsit5 steps */
typedef struct OqSolveOptions {
    double reltol, abstol, dt0, dtmax, tstop;
    int64_t maxiters;
    int32_t algorithm;      /* OQ_ALG_TSIT5 (test/tests.jl:11) or OQ_ALG_VCABM5 (examples/otf-with-mantle.jl:160) */
    int32_t fixed_dt;       /* != 0: take fixed steps of dt0 (no error control) */
    int32_t async_snapshots;/* != 0: snapshots go through a device-side ring drained by a second stream; fn runs while
                             * the next batch of steps executes (wsolve's mode, src/io.jl:51-58,128-130).  A stop request
                             * then takes effect within one batch (<= 16 steps) instead of at exactly that step. */
    int32_t reserved;
} OqSolveOptions;

typedef struct OqSolveStats {
    double t, dt_last, dt_next;
    int64_t naccept, nreject, nrhs;
    int32_t retcode;        /* 0 success (t == tstop), 1 maxiters, 2 dt underflow / unstable, 3 a peer rank did not
                             * deliver in time (multi-GPU; the call also returns non-zero), 4 terminated by the callback */
} OqSolveStats;

/* Snapshot callback: called on the host after every `stride`-th accepted step (and at t0), with the
 * state already copied to host buffers (the role of wsolve's FunctionCallingCallback, src/io.jl:51-58).
 * Return non-zero to stop the integration. */
typedef int (*OqSnapshotFn)(void *user, double t, int64_t step, const double *const *u_parts,
                            const double *const *du_parts);

/* Advance the resident state from t0 to opts->tstop on the device: RK stage combinations, the error
 * norm and the step-size controller run on the GPU; the host reads one small record per step. */
int oq_solve(OqProblem *p, double t0, const OqSolveOptions *opts, int64_t stride, OqSnapshotFn fn,
             void *user, OqSolveStats *stats);

/* ------------------------------------------------------------------ multi-GPU (one process per GPU)
 * Rows are sharded; every RHS all-gathers the two forcing vectors (v - vpl, dϵ - dϵ0) and every
 * step all-reduces one double.  The exchange runs over peer-mapped device memory (NVLink): each
 * rank exports a small window with oq_comm_export, the host runtime (torch.distributed, MPI, Julia
 * Distributed ...) all-gathers the opaque handles, and oq_comm_connect maps the peers' windows. */
#define OQ_COMM_HANDLE_BYTES 128
int oq_comm_export(OqProblem *p, int rank, int world, uint8_t handle[OQ_COMM_HANDLE_BYTES]);
int oq_comm_connect(OqProblem *p, const uint8_t *all_handles /* [world*OQ_COMM_HANDLE_BYTES] */);

#ifdef __cplusplus
}
#endif
#endif /* OETQF_B200_H */
