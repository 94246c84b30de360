// matvec_panel.cuh -- the fused RHS kernel: forcing vectors + peer all-gather + HBM-streaming matvec + pointwise
// physics in ONE launch per evaluation of /root/reference/src/BEM/equation.jl:156-205.
//
// Successor of matvec_stream.cuh (kept as a validation twin, OQ_MATVEC=stream), built for the small row shards of
// the 8-GPU configuration, where the evaluation is 40-50 us long and every microsecond outside the streaming loop
// shows:
//   * PROLOGUE in the kernel (was forcing_kernel, a second launch): the consumer warps of every CTA, idle until the
//     first forcing piece arrives, form their slice of v - vpl and dϵ - dϵ0 (with the integrator's stage combination fused in), stores it into
//     every rank's window over NVLink and joins a grid-wide count; the last CTA publishes the epoch to the peers.
//     The producer warp has its whole ring of matrix tiles in flight before any of this is awaited.
//   * PANEL traversal: a CTA walks its span of chunks in panels of up to kPnP row blocks, column by column, starting
//     with the columns THIS rank owns.  The forcing piece of a column is loaded once per panel (not once per chunk:
//     -20 % shared-memory fill traffic, one more ring stage), and a peer's flag is awaited only when the first of
//     its columns is reached -- by then the locally owned columns of the whole panel have been multiplied, which
//     hides the NVLink round trip of the exchange.
//   * the traversal direction (forwards / backwards through the row blocks, for L2 reuse between evaluations)
//     comes from the host, so a CTA of the NEXT evaluation that starts early under programmatic dependent launch
//     can fetch its first tiles before the previous kernel has finished.
// Roles per CTA as before: producer warp (bulk TMA into a kPnStages-deep ring + a double-buffered forcing piece),
// 8 consumer warps (LDS.128 + DFMA, kPnP x 4 running sums each), epilogue warp (prologue, fixed-order folds,
// cross-CTA merge of cut row blocks, friction law / stress-rate store).  Every sum is formed in an order that
// depends only on (world size, rank, grid size): results are bitwise repeatable.
#pragma once

namespace oq {

constexpr int kPnR = kStR;          // rows per row block
constexpr int kPnCH = kStCH;        // columns per chunk
constexpr int kPnStages = 6;        // ring stages of kPnR x kPnCH doubles (32 KB each)
constexpr int kPnPMax = 6;          // row blocks per panel (24 running sums per consumer thread); template parameter P <= kPnPMax
constexpr int kPnConsumers = 256;
constexpr int kPnCWarps = kPnConsumers / 32;
constexpr int kPnThreads = kPnConsumers + 64;
constexpr int kPnStageDoubles = kPnR * kPnCH;
constexpr size_t kPnSmemBytes = ((size_t)kPnStages * kPnStageDoubles + 2 * kPnCH) * sizeof(double) + 1024;

// the pointwise front end of an evaluation (what forcing_kernel does when it is a launch of its own)
struct Prologue {
    int enabled;           // 0: the forcing vectors were produced by a preceding forcing_kernel
    ForcingArgs fa;
};

struct PanelArgs {
    MatvecArgs mv;
    Prologue pro;
    ColOwners own;
    int reverse;           // walk the row blocks of a span backwards (alternates between evaluations)
    int cstart[2];         // first column chunk of a panel per row set: the chunk holding this rank's own columns
    int ne;                // mantle elements (column p*ne + e of the strain-rate operand belongs to element e)
    unsigned long long* timeline;   // debug (OQ_TIMELINE=file): [gridDim.x][32] globaltimer stamps, null: off
};

// debug timeline slots
enum : int { kTlStart = 0, kTlProdPre = 1, kTlProdWait = 2, kTlProdEp = 3, kTlPeer0 = 4 /* +rank, 16 */, kTlFirstX = 20,
             kTlProdEnd = 21, kTlProlBegin = 22, kTlProlStores = 23, kTlProlArrive = 24, kTlConsFirst = 25, kTlConsEnd = 26,
             kTlEpiEnd = 27, kTlPublished = 28 };
__device__ __forceinline__ void tl_stamp(const PanelArgs& A, int slot)
{
    if (A.timeline) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        A.timeline[(size_t)blockIdx.x * 32 + slot] = t;
    }
}

// segments [k0, k0 + ns) of a span (row-block parts in processing order) that form one panel
template <int kPnP>
struct Panel {
    int job, ns, cpr;
    int rb[kPnP], lo[kPnP], hi[kPnP];
    long long g0[kPnP], g1[kPnP];
    __device__ __forceinline__ int load(const MatvecArgs& a, const SpanWalk& w, int k0)
    {
        ns = 0; job = 0; cpr = 0;
        bool open = true;
#pragma unroll
        for (int s = 0; s < kPnP; ++s) {
            rb[s] = 0; lo[s] = 0; hi[s] = 0; g0[s] = 0; g1[s] = 0;
            if (open && k0 + s < w.nseg) {
                int jb, r, rem0, n;
                long long a0, a1;
                w.get(a, k0 + s, jb, r, rem0, n, a0, a1);
                if (s == 0) { job = jb; cpr = a.job[jb].chunks_per_rb; }
                if (jb == job) { rb[s] = r; lo[s] = rem0; hi[s] = rem0 + n; g0[s] = a0; g1[s] = a1; ns = s + 1; }
                else open = false;
            }
        }
        return k0 + ns;
    }
    __device__ __forceinline__ bool any(int c) const
    {
        bool f = false;
#pragma unroll
        for (int s = 0; s < kPnP; ++s) f |= (s < ns && lo[s] <= c && c < hi[s]);
        return f;
    }
};

// geometry of column chunk c of a row set: operand, first column, width (a multiple of 16 doubles)
__device__ __forceinline__ void panel_column(const MatvecJob& j, int c, int& osel, int& c0, int& ncol)
{
    osel = c < j.nch[0] ? 0 : 1;
    const MatOperand& op = j.op[osel];
    c0 = (osel ? c - j.nch[0] : c) * kPnCH;
    const int colsp = (op.cols + 15) & ~15;
    ncol = min(kPnCH, colsp - c0);
}

// one pass of the forcing front end over this CTA's slice of the local rows (executed by `nthr` threads)
__device__ __forceinline__ void prologue_slice(const ForcingArgs& a, int lane, int nthr, unsigned long long ep)
{
    const size_t par = (size_t)(ep & 1ull);
    const int world = a.peers.world;
    const bool staged = a.stage.nk > 0;
    const double dt = staged ? (a.stage.dt ? *a.stage.dt : 1.0) : 0.0;
    const double* __restrict__ adev = a.stage.adev;
    auto combine = [&](size_t idx) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < 6; ++j)
            if (j < a.stage.nk) acc = fma(adev ? adev[j] : a.stage.a[j], a.stage.k[j][idx], acc);
        return fma(dt, acc, a.stage.u[idx]);
    };
    const int G = gridDim.x, b = blockIdx.x;
    const int t0 = (int)(((long long)a.nfl * b) / G), t1 = (int)(((long long)a.nfl * (b + 1)) / G);
    for (int t = t0 + lane; t < t1; t += nthr) {
        // (the fault partitions of the stage state are written by the epilogue of the row, which needs them anyway)
        const double vt = staged ? combine(a.off_fault[0] + t) : a.v[t];
        const double rv = vt - a.vpl;                                        // equation.jl:38
        const size_t off = a.wl.off_relv + par * a.wl.relv_len + a.f0 + t;
        for (int r = 0; r < world; ++r) a.peers.base[r][off] = rv;           // local + NVLink peer stores
    }
    const int e0 = (int)(((long long)a.nel * b) / G), e1 = (int)(((long long)a.nel * (b + 1)) / G);
    for (int t = e0 + lane; t < e1; t += nthr) {
        const size_t n = a.nel;
        double s[6];
        if (staged) {
#pragma unroll
            for (int k = 0; k < 6; ++k) a.y[a.off_eps + t + k * n] = combine(a.off_eps + t + k * n);
#pragma unroll
            for (int k = 0; k < 6; ++k) { s[k] = combine(a.off_sig + t + k * n); a.y[a.off_sig + t + k * n] = s[k]; }
        } else {
#pragma unroll
            for (int k = 0; k < 6; ++k) s[k] = a.sig[t + k * n];
        }
        const double skk = (s[0] + s[3] + s[5]) / 3;                         // equation.jl:209
        const double sxx = s[0] - skk, syy = s[3] - skk, szz = s[5] - skk;
        const double tn = sqrt(sxx * sxx + syy * syy + szz * szz + 2 * (s[1] * s[1] + s[2] * s[2] + s[4] * s[4]));
        const double comp[6] = {sxx, s[1], s[2], syy, s[4], szz};
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
}

// the epilogue of a fault row with the stage combination of its own partitions fused in (the prologue only formed
// v for the exchange): y = u + dt sum a_j k_j for v, θ, δ (, 𝓅), stored, then the friction law on (y_v, y_θ)
__device__ __forceinline__ void fault_row_epilogue(const PanelArgs& A, int row, double dtau)
{
    const ForcingArgs& a = A.pro.fa;
    if (A.pro.enabled && a.stage.nk > 0) {
        const double dt = a.stage.dt ? *a.stage.dt : 1.0;
        const double* __restrict__ adev = a.stage.adev;
        for (int q = 0; q < a.n_fault_parts; ++q) {
            const size_t idx = a.off_fault[q] + row;
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j < 6; ++j)
                if (j < a.stage.nk) acc = fma(adev ? adev[j] : a.stage.a[j], a.stage.k[j][idx], acc);
            a.y[idx] = fma(dt, acc, a.stage.u[idx]);
        }
    }
    update_fault_row(A.mv.fe, row, dtau);
}

template <int kPnP>
__global__ void __launch_bounds__(kPnThreads, 1)
matvec_panel_kernel(const __grid_constant__ PanelArgs A)
{
    using Panel = oq::Panel<kPnP>;
    extern __shared__ __align__(128) double smem[];
    __shared__ __align__(8) uint64_t full_bar[kPnStages];
    __shared__ __align__(8) uint64_t empty_bar[kPnStages];
    __shared__ __align__(8) uint64_t xfull[2];
    __shared__ __align__(8) uint64_t xempty[2];
    __shared__ __align__(8) uint64_t red_full[2];
    __shared__ __align__(8) uint64_t red_empty[2];
    __shared__ __align__(8) uint64_t ep_bar;
    __shared__ double red[2][kPnCWarps][kPnPMax][kPnR];
    __shared__ unsigned long long ep_s;
    __shared__ int done_s;

    const MatvecArgs& args = A.mv;
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const long long total = args.total_chunks;
    const int grid = gridDim.x;
    const long long g_begin = span_begin(total, grid, blockIdx.x);
    const long long g_end = span_begin(total, grid, blockIdx.x + 1);
    double* xbuf = smem + (size_t)kPnStages * kPnStageDoubles;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kPnStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kPnCWarps);
        }
        mbar_init(&xfull[0], 1); mbar_init(&xfull[1], 1);
        mbar_init(&xempty[0], kPnCWarps); mbar_init(&xempty[1], kPnCWarps);
        mbar_init(&red_full[0], kPnCWarps); mbar_init(&red_full[1], kPnCWarps);
        mbar_init(&red_empty[0], 1); mbar_init(&red_empty[1], 1);
        mbar_init(&ep_bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) tl_stamp(A, kTlStart);
    // the next evaluation's CTAs may be scheduled as soon as this grid's CTAs leave their SMs
    pdl_launch_dependents();

    SpanWalk walk;
    const bool has_work = g_begin < g_end;
    if (has_work) walk.init(args, g_begin, g_end, A.reverse != 0);

    if (warp == kPnCWarps) {
        // ------------------------------------------------------------------ producer warp
        if (lane != 0) return;
        // matrix tile (panel pn, segment s, column c) -> ring stage
        auto issue_tile = [&](const Panel& pn, int s, int c, int stg) {
            const MatvecJob& j = args.job[pn.job];
            int osel, c0, ncol;
            panel_column(j, c, osel, c0, ncol);
            const MatOperand& op = j.op[osel];
            const unsigned bytes = (unsigned)(ncol * sizeof(double));
            double* dst = smem + (size_t)stg * kPnStageDoubles;
            mbar_arrive_expect_tx(&full_bar[stg], bytes * kPnR);
            const long long g = pn.g0[s] + c;
            const unsigned long long pol = args.keep_chunks < 0 ? kL2EvictNormal
                                           : (g - g_begin < args.keep_chunks ? kL2EvictLast : kL2EvictFirst);
#pragma unroll
            for (int r = 0; r < kPnR; ++r) {
                const int row = min(pn.rb[s] * kPnR + r, j.nrows - 1);
                tma_load_1d_hint(dst + r * kPnCH, op.G + (size_t)row * op.ld + c0, bytes, &full_bar[stg], pol);
            }
        };
        // 1. the first ring of matrix tiles does not depend on this evaluation's state: request it right away
        int npre = 0;
        if (has_work) {
            Panel pn;
            int k = 0;
            while (k < walk.nseg && npre < kPnStages) {
                const int kn = pn.load(args, walk, k);
                for (int ci = 0; ci < pn.cpr && npre < kPnStages; ++ci) {
                    int c = A.cstart[pn.job] + ci;
                    if (c >= pn.cpr) c -= pn.cpr;
#pragma unroll
                    for (int s = 0; s < kPnP; ++s)
                        if (s < pn.ns && pn.lo[s] <= c && c < pn.hi[s] && npre < kPnStages) { issue_tile(pn, s, c, npre); ++npre; }
                }
                k = kn;
            }
        }
        // 2. everything else reads what the predecessor kernel (and this grid's prologue) produced
        tl_stamp(A, kTlProdPre);
        pdl_wait();
        tl_stamp(A, kTlProdWait);
        mbar_wait(&ep_bar, 0);                        // the epilogue warp has read the epoch / the done flag
        tl_stamp(A, kTlProdEp);
        if (done_s || !has_work) {
            for (int s = 0; s < npre; ++s) mbar_wait(&full_bar[s], 0);   // never leave with bulk copies in flight
            return;
        }
        const unsigned long long ep = ep_s;           // publications before this evaluation
        const size_t par = (size_t)(ep & 1ull);
        // ranks whose forcing slice of this evaluation has arrived (a plain gemv has no exchange at all)
        unsigned confirmed = args.pw.epochs ? 0u : 0xffffffffu;
        const int world = A.own.world, self = args.pw.rank;
        auto need_rank = [&](int r) {
            if (confirmed & (1u << r)) return;
            if (r == self || world == 1) spin_until(args.pw.epochs + kEpForcing, ep + 1ull, false, args.pw.epochs);
            else {
                spin_until(args.pw.flags + r, ep + 1ull, true, args.pw.epochs);
                fence_proxy_async();                  // peer stores -> async-proxy (TMA) reads
            }
            confirmed |= 1u << r;
            tl_stamp(A, kTlPeer0 + (r & 15));
        };
        auto need_columns = [&](int x_is_strain, int c0, int ncol) {
            if (confirmed == 0xffffffffu) return;
            if (world == 1) { need_rank(0); return; }
            if (!x_is_strain) {
                for (int r = 0; r < world; ++r)
                    if (A.own.fb[r] < c0 + ncol && A.own.fb[r + 1] > c0) need_rank(r);
            } else {
                const int ne = A.ne;
                const int p0 = c0 / ne, p1 = (c0 + ncol - 1) / ne;
                if (p0 != p1) { for (int r = 0; r < world; ++r) need_rank(r); return; }
                const int ea = c0 - p0 * ne, eb = ea + ncol;
                for (int r = 0; r < world; ++r)
                    if (A.own.eb[r] < eb && A.own.eb[r + 1] > ea) need_rank(r);
            }
        };
        int stage = 0, issued = 0, xcount = 0;
        unsigned phase = 0;
        Panel pn;
        for (int k = 0; k < walk.nseg;) {
            const int kn = pn.load(args, walk, k);
            const MatvecJob& j = args.job[pn.job];
            for (int ci = 0; ci < pn.cpr; ++ci) {
                int c = A.cstart[pn.job] + ci;
                if (c >= pn.cpr) c -= pn.cpr;
                if (!pn.any(c)) continue;
                int osel, c0, ncol;
                panel_column(j, c, osel, c0, ncol);
                const MatOperand& op = j.op[osel];
                // forcing piece of this column: awaited from its owners, loaded once for the whole panel
                const int xs = xcount & 1;
                mbar_wait(&xempty[xs], ((xcount >> 1) & 1) ^ 1u);
                need_columns(op.x_kind, c0, ncol);
                const unsigned xbytes = (unsigned)(ncol * sizeof(double));
                mbar_arrive_expect_tx(&xfull[xs], xbytes);
                tma_load_1d_hint(xbuf + (size_t)xs * kPnCH, op.x + par * op.x_stride + c0, xbytes, &xfull[xs],
                                 args.keep_chunks < 0 ? kL2EvictNormal : kL2EvictLast);
                if (xcount == 0) tl_stamp(A, kTlFirstX);
                ++xcount;
#pragma unroll
                for (int s = 0; s < kPnP; ++s) {
                    if (s < pn.ns && pn.lo[s] <= c && c < pn.hi[s]) {
                        if (issued >= npre) {
                            mbar_wait(&empty_bar[stage], phase ^ 1u);
                            issue_tile(pn, s, c, stage);
                        }
                        ++issued;
                        if (++stage == kPnStages) { stage = 0; phase ^= 1u; }
                    }
                }
            }
            k = kn;
        }
        tl_stamp(A, kTlProdEnd);
        return;
    }

    if (warp == kPnCWarps + 1) {
        // ------------------------------------------------------------------ epilogue warp (prologue first)
        pdl_wait();                                   // state, stage slopes, epoch counters of the predecessor
        int done = 0;
        unsigned long long ep = 0ull;
        if (lane == 0) {
            done = (args.done && *reinterpret_cast<const volatile int*>(args.done)) ? 1 : 0;
            ep = args.pw.epochs ? *reinterpret_cast<volatile unsigned long long*>(args.pw.epochs + kEpForcing) : 0ull;
            // with a preceding forcing_kernel the publication is already counted: step back to "before it"
            if (!A.pro.enabled && args.pw.epochs) ep -= 1ull;
            ep_s = ep;
            done_s = done;
            mbar_arrive(&ep_bar);                     // release: producer and consumers may read ep_s / done_s
        }
        done = __shfl_sync(0xffffffffu, done, 0);
        ep = __shfl_sync(0xffffffffu, ep, 0);
        if (done) return;                             // integration already complete: identically on every rank
        if (!has_work) return;
        int buf = 0;
        unsigned rphase[2] = {0u, 0u};
        const int s_of = lane >> 2, r_of = lane & 3;  // lane -> (segment of the panel, row of the row block)
        Panel pn;
        for (int k = 0; k < walk.nseg;) {
            const int kn = pn.load(args, walk, k);
            const MatvecJob& j = args.job[pn.job];
            mbar_wait(&red_full[buf], rphase[buf]);   // all consumer warps have dropped the panel's partial sums
            rphase[buf] ^= 1u;
            const bool mine_seg = s_of < pn.ns && s_of < kPnP;
            double mine = 0.0;
            if (mine_seg) {
#pragma unroll
                for (int w = 0; w < kPnCWarps; ++w) mine += red[buf][w][s_of][r_of];
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&red_empty[buf]);          // consumers may reuse the slot
            buf ^= 1;
            // segment geometry of this lane (dynamic index into the panel: a handful of local loads per panel)
            int rb = 0;
            long long rb_g0 = 0, rb_g1 = 0;
#pragma unroll
            for (int s = 0; s < kPnP; ++s)
                if (s == s_of) { rb = pn.rb[s]; rb_g0 = pn.g0[s]; rb_g1 = pn.g1[s]; }
            const int myrow = rb * kPnR + r_of;
            const bool active = mine_seg && myrow < j.nrows;
            // contributors: only a row block cut by a span boundary has more than one
            int first = blockIdx.x, last = blockIdx.x;
            if (mine_seg && rb_g0 < g_begin) first = owner_of(total, grid, rb_g0);
            if (mine_seg && rb_g1 > g_end) last = owner_of(total, grid, rb_g1 - 1);
            const int ncontrib = last - first + 1;
            bool do_epilogue = mine_seg;
            // (every lane takes part in the votes / shuffles below; lanes without a segment carry ncontrib == 1)
            if (__any_sync(0xffffffffu, ncontrib > 1)) {
                if (ncontrib > 1 && active) j.partial[((size_t)myrow) * j.slots + ((int)blockIdx.x - first)] = mine;
                __threadfence();
                __syncwarp();
                unsigned prev = 0;
                if (ncontrib > 1 && r_of == 0 && mine_seg) {
                    prev = atomicAdd(&j.counters[rb], 1u);
                    if (prev == (unsigned)ncontrib - 1u) j.counters[rb] = 0u;   // re-arm for the next evaluation
                }
                prev = __shfl_sync(0xffffffffu, prev, lane & ~3);
                if (ncontrib > 1) {
                    do_epilogue = mine_seg && (prev == (unsigned)ncontrib - 1u);
                    if (do_epilogue) {
                        __threadfence();
                        if (active) {
                            mine = 0.0;
                            const double* pp = j.partial + (size_t)myrow * j.slots;
                            for (int q = 0; q < ncontrib; ++q) mine += ld_cg(pp + q);   // fixed order: deterministic
                        }
                    }
                }
            }
            if (do_epilogue && active) {
                if (j.y0) mine += j.y0[myrow];
                if (j.epilogue == kEpiFault) fault_row_epilogue(A, myrow, mine);
                else j.yout[myrow] = mine;
            }
            k = kn;
        }
        if (lane == 0) tl_stamp(A, kTlEpiEnd);
        return;
    }

    // ---------------------------------------------------------------------- consumer warps
    mbar_wait(&ep_bar, 0);
    if (done_s) return;
    if (A.pro.enabled) {
        pdl_wait();                                   // the prologue reads the predecessor's state and slopes
        // the forcing front end: idle until the first forcing piece arrives anyway, the 256 consumer threads form this
        // CTA's slice, then one of them joins the grid-wide count (the last CTA publishes the epoch)
        const unsigned long long ep = ep_s;
        if (tid == 0) tl_stamp(A, kTlProlBegin);
        prologue_slice(A.pro.fa, tid, kPnConsumers, ep);
        asm volatile("bar.sync 1, %0;" ::"n"(kPnConsumers) : "memory");
        if (tid == 0) {
            tl_stamp(A, kTlProlStores);
            const int world = A.pro.fa.peers.world;
            unsigned long long* epochs = A.pro.fa.epochs;
            if (world > 1) __threadfence_system(); else __threadfence();
            const unsigned long long prev = atomicAdd(&epochs[kEpBlocksF], 1ull);
            tl_stamp(A, kTlProlArrive);
            if (prev == (unsigned long long)gridDim.x - 1ull) {
                epochs[kEpBlocksF] = 0ull;
                if (world > 1) {
                    __threadfence_system();           // acquire the other CTAs' slices before telling the peers
                    for (int r = 0; r < world; ++r) {
                        if (r == A.pro.fa.peers.rank) continue;
                        unsigned long long* f =
                            reinterpret_cast<unsigned long long*>(A.pro.fa.peers.base[r] + A.pro.fa.wl.off_flags);
                        publish_flag(f + A.pro.fa.peers.rank, ep + 1ull);
                    }
                }
                __threadfence();
                asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(epochs + kEpForcing), "l"(ep + 1ull) : "memory");
                tl_stamp(A, kTlPublished);
            }
        }
    }
    if (!has_work) return;
    double acc[kPnP][kPnR];
#pragma unroll
    for (int s = 0; s < kPnP; ++s)
#pragma unroll
        for (int r = 0; r < kPnR; ++r) acc[s][r] = 0.0;
    int stage = 0, buf = 0, xcount = 0;
    unsigned phase = 0;
    unsigned ephase[2] = {0u, 0u};
    Panel pn;
    for (int k = 0; k < walk.nseg;) {
        const int kn = pn.load(args, walk, k);
        const MatvecJob& j = args.job[pn.job];
        for (int ci = 0; ci < pn.cpr; ++ci) {
            int c = A.cstart[pn.job] + ci;
            if (c >= pn.cpr) c -= pn.cpr;
            if (!pn.any(c)) continue;
            int osel, c0, ncol;
            panel_column(j, c, osel, c0, ncol);
            const int xs = xcount & 1;
            mbar_wait(&xfull[xs], (xcount >> 1) & 1u);
            if (xcount == 0 && tid == 0) tl_stamp(A, kTlConsFirst);
            // this thread's two double2 of the forcing piece stay in registers for every row block of the panel
            const double2* x2 = reinterpret_cast<const double2*>(xbuf + (size_t)xs * kPnCH);
            const bool in0 = 2 * tid < ncol, in1 = 2 * (tid + kPnConsumers) < ncol;
            const double2 xa = in0 ? x2[tid] : make_double2(0.0, 0.0);
            const double2 xb = in1 ? x2[tid + kPnConsumers] : make_double2(0.0, 0.0);
            __syncwarp();
            if (lane == 0) mbar_arrive(&xempty[xs]);              // the producer may refill the piece
            ++xcount;
#pragma unroll
            for (int s = 0; s < kPnP; ++s) {
                if (s < pn.ns && pn.lo[s] <= c && c < pn.hi[s]) {
                    mbar_wait(&full_bar[stage], phase);
                    const double2* s2 = reinterpret_cast<const double2*>(smem + (size_t)stage * kPnStageDoubles);
                    if (in0) {
#pragma unroll
                        for (int r = 0; r < kPnR; ++r) {
                            const double2 gv = s2[r * (kPnCH / 2) + tid];
                            acc[s][r] = fma(gv.x, xa.x, acc[s][r]);
                            acc[s][r] = fma(gv.y, xa.y, acc[s][r]);
                        }
                    }
                    if (in1) {
#pragma unroll
                        for (int r = 0; r < kPnR; ++r) {
                            const double2 gv = s2[r * (kPnCH / 2) + tid + kPnConsumers];
                            acc[s][r] = fma(gv.x, xb.x, acc[s][r]);
                            acc[s][r] = fma(gv.y, xb.y, acc[s][r]);
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty_bar[stage]);
                    if (++stage == kPnStages) { stage = 0; phase ^= 1u; }
                }
            }
        }
        // end of the panel: hand the partial sums of its row blocks to the epilogue warp and keep streaming
#pragma unroll
        for (int s = 0; s < kPnP; ++s) {
            if (s < pn.ns) {
#pragma unroll
                for (int r = 0; r < kPnR; ++r) {
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) acc[s][r] += __shfl_xor_sync(0xffffffffu, acc[s][r], off);
                }
            }
        }
        if (lane == 0) {
            mbar_wait(&red_empty[buf], ephase[buf] ^ 1u);         // slot free (always, except pathologically)
#pragma unroll
            for (int s = 0; s < kPnP; ++s)
#pragma unroll
                for (int r = 0; r < kPnR; ++r) red[buf][warp][s][r] = acc[s][r];
            mbar_arrive(&red_full[buf]);                          // release: the stores above are visible
        }
        ephase[buf] ^= 1u;
        buf ^= 1;
#pragma unroll
        for (int s = 0; s < kPnP; ++s)
#pragma unroll
            for (int r = 0; r < kPnR; ++r) acc[s][r] = 0.0;
        k = kn;
    }
    if (tid == 0) tl_stamp(A, kTlConsEnd);
}

}  // namespace oq
