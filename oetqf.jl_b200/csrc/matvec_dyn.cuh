// matvec_dyn.cuh -- the streaming matvec of matvec_stream.cuh with DYNAMIC work distribution.
//
// Same three roles per CTA (bulk-TMA producer warp, 8 consumer warps, epilogue warp) and the same shared-memory
// ring; what changes is who streams what.  ncu on the static split showed SMs finishing between 515k and 664k
// cycles of a 650k-cycle launch (sm__cycles_active min/avg/max, profiles/r01_matvec_stream_ncu_full.md): per-SM
// HBM throughput differs by +-13 %, so equal shares leave the fast SMs idle for the last tenth of the launch.
// Here the launch's work is cut into UNITS (one piece = `piece_len` consecutive chunks of one row block) and the
// producer warps draw units from a global ticket counter until none are left.
//
// Determinism: a unit's partial sums depend only on the unit (same thread-to-column map, same chunk order), a row
// block's pieces are folded in piece order by whichever CTA delivers the last one, so the result is bitwise
// independent of which SM streamed what.
//
// Ticket counter without a reset kernel: every CTA draws exactly one ticket past the end, so a launch consumes
// exactly U + G tickets and launch k owns the ticket range [k(U+G), (k+1)(U+G)); the first ticket a CTA draws
// tells it the base of the current launch.
#pragma once

namespace oq {

struct UnitRef {
    int job, rb, piece, c0, c1;        // chunks [c0, c1) of row block rb
};

__device__ __forceinline__ UnitRef decode_unit(const MatvecArgs& a, long long u)
{
    UnitRef r;
    r.job = (a.job[0].nrb > 0 && u < a.job[1].unit_begin) || a.job[1].nrb == 0 ? 0 : 1;
    const MatvecJob& j = a.job[r.job];
    const int loc = (int)(u - j.unit_begin);
    r.rb = loc / j.pieces_per_rb;
    r.piece = loc - r.rb * j.pieces_per_rb;
    r.c0 = r.piece * a.piece_len;
    r.c1 = min(r.c0 + a.piece_len, j.chunks_per_rb);
    return r;
}

__global__ void __launch_bounds__(kStThreads, 1)
matvec_dyn_kernel(const __grid_constant__ MatvecArgs args)
{
    extern __shared__ __align__(128) double smem[];
    __shared__ __align__(8) uint64_t full_bar[kStStages];
    __shared__ __align__(8) uint64_t empty_bar[kStStages];
    __shared__ __align__(8) uint64_t red_full[2];
    __shared__ __align__(8) uint64_t red_empty[2];
    __shared__ double red[2][kStCWarps][kStR];
    __shared__ int4 meta[kStStages];      // per ring stage: job (-1: no more work), row block, chunk, piece<<1 | last
    __shared__ int4 red_meta[2];          // per reduction slot: job (-1: end), row block, piece

    if (args.done && *reinterpret_cast<const volatile int*>(args.done)) return;   // integration already complete
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kStCWarps);
        }
        mbar_init(&red_full[0], kStCWarps); mbar_init(&red_full[1], kStCWarps);
        mbar_init(&red_empty[0], 1); mbar_init(&red_empty[1], 1);
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == kStCWarps) {
        // ------------------------------------------------------------------ producer warp
        if (lane != 0) return;
        const long long U = args.total_units;
        const unsigned long long period = (unsigned long long)U + gridDim.x;
        unsigned long long base = 0;
        bool have_base = false;
        auto draw = [&]() -> long long {               // next unit of this launch, or -1
            const unsigned long long t = atomicAdd(args.ticket, 1ull);
            if (!have_base) { base = (t / period) * period; have_base = true; }
            const long long u = (long long)(t - base);
            return u < U ? u : -1;
        };
        auto issue = [&](const UnitRef& un, int rem, int stg, bool matrix, bool vector, size_t par) {
            const MatvecJob& j = args.job[un.job];
            const int osel = rem < j.nch[0] ? 0 : 1;
            const MatOperand& op = j.op[osel];
            const int c0 = (osel ? rem - j.nch[0] : rem) * kStCH;
            const int ncol = min(kStCH, (int)op.ld - c0);
            const unsigned bytes = (unsigned)(ncol * sizeof(double));
            double* dst = smem + (size_t)stg * kStStageDoubles;
            if (matrix) {
                meta[stg] = make_int4(un.job, un.rb, rem, (un.piece << 1) | (rem + 1 == un.c1 ? 1 : 0));
                mbar_arrive_expect_tx(&full_bar[stg], bytes * (kStR + 1));     // release: meta is visible with the data
#pragma unroll
                for (int r = 0; r < kStR; ++r) {
                    const int row = min(un.rb * kStR + r, j.nrows - 1);
                    tma_load_1d(dst + r * kStCH, op.G + (size_t)row * op.ld + c0, bytes, &full_bar[stg]);
                }
            }
            if (vector) tma_load_1d(dst + kStR * kStCH, op.x + par * op.x_stride + c0, bytes, &full_bar[stg]);
        };
        // The matrix does not depend on this evaluation's forcing vector: the first ring of matrix pieces is
        // requested BEFORE waiting for the forcing kernel / the peers' publication.
        UnitRef un{};
        int rem = 0;
        bool more = false;
        auto advance = [&]() {                        // move (un, rem) to the next chunk to stream
            if (more && ++rem < un.c1) return;
            const long long u = draw();
            more = u >= 0;
            if (more) { un = decode_unit(args, u); rem = un.c0; }
        };
        UnitRef pre_un[kStStages];
        int pre_rem[kStStages];
        int npre = 0;
        advance();
        while (npre < kStStages && more) {
            pre_un[npre] = un; pre_rem[npre] = rem;
            issue(un, rem, npre, true, false, 0);
            ++npre;
            advance();
        }
        pdl_wait();                                   // the forcing kernel (predecessor) is complete from here on
        size_t par = 0;
        if (args.pw.epochs) {
            const unsigned long long ep = *(volatile unsigned long long*)(args.pw.epochs + kEpForcing);
            if (args.pw.world > 1) {
                wait_peers(args.pw.flags, args.pw.world, args.pw.rank, ep, args.pw.epochs + kEpError);
                fence_proxy_async();                  // peer stores -> async-proxy (TMA) reads
            }
            par = (size_t)((ep - 1ull) & 1ull);
        }
        for (int s = 0; s < npre; ++s) issue(pre_un[s], pre_rem[s], s, false, true, par);
        int stage = npre % kStStages;
        unsigned phase = npre >= kStStages ? 1u : 0u;
        while (more) {
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            issue(un, rem, stage, true, true, par);
            if (++stage == kStStages) { stage = 0; phase ^= 1u; }
            advance();
        }
        // no more work: tell the consumers through the ring
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        meta[stage] = make_int4(-1, 0, 0, 0);
        mbar_arrive(&full_bar[stage]);
        return;
    }

    if (warp == kStCWarps + 1) {
        // ------------------------------------------------------------------ epilogue warp
        pdl_wait();                                   // the physics reads the predecessor's state
        int buf = 0;
        unsigned rphase[2] = {0u, 0u};
        for (;;) {
            mbar_wait(&red_full[buf], rphase[buf]);   // all consumer warps have dropped their partial sums
            rphase[buf] ^= 1u;
            const int4 rm = red_meta[buf];
            if (rm.x < 0) break;
            const MatvecJob& j = args.job[rm.x];
            const int rb = rm.y, piece = rm.z;
            const int myrow = rb * kStR + lane;
            const bool active = lane < kStR && myrow < j.nrows;
            double mine = 0.0;
            if (lane < kStR) {
#pragma unroll
                for (int w = 0; w < kStCWarps; ++w) mine += red[buf][w][lane];
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&red_empty[buf]);          // consumers may reuse the slot
            buf ^= 1;
            const int ncontrib = j.pieces_per_rb;
            bool do_epilogue = true;
            if (ncontrib > 1) {
                if (active) j.partial[((size_t)myrow) * j.slots + piece] = mine;
                __threadfence();
                __syncwarp();
                unsigned prev = 0;
                if (lane == 0) {
                    prev = atomicAdd(&j.counters[rb], 1u);
                    if (prev == (unsigned)ncontrib - 1u) j.counters[rb] = 0u;   // re-arm for the next evaluation
                }
                prev = __shfl_sync(0xffffffffu, prev, 0);
                do_epilogue = (prev == (unsigned)ncontrib - 1u);
                if (do_epilogue) {
                    __threadfence();
                    if (active) {
                        mine = 0.0;
                        const double* pp = j.partial + (size_t)myrow * j.slots;
                        for (int q = 0; q < ncontrib; ++q) mine += ld_cg(pp + q);   // piece order: deterministic
                    }
                }
            }
            if (do_epilogue && active) {
                if (j.y0) mine += j.y0[myrow];
                if (j.epilogue == kEpiFault) update_fault_row(args.fe, myrow, mine);
                else j.yout[myrow] = mine;
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- consumer warps
    double acc[kStR];
#pragma unroll
    for (int r = 0; r < kStR; ++r) acc[r] = 0.0;
    int stage = 0, buf = 0;
    unsigned phase = 0;
    unsigned ephase[2] = {0u, 0u};

    for (;;) {
        mbar_wait(&full_bar[stage], phase);
        const int4 m = meta[stage];
        if (m.x < 0) break;
        const MatvecJob& j = args.job[m.x];
        const int osel = m.z < j.nch[0] ? 0 : 1;
        const MatOperand& op = j.op[osel];
        const int ncol = min(kStCH, (int)op.ld - (osel ? m.z - j.nch[0] : m.z) * kStCH);
        const double2* s2 = reinterpret_cast<const double2*>(smem + (size_t)stage * kStStageDoubles);
#pragma unroll
        for (int it = 0; it < kStCH / (2 * kStConsumers); ++it) {
            const int c2 = tid + it * kStConsumers;               // double2 index inside the chunk
            if (2 * c2 < ncol) {
                const double2 xv = s2[kStR * (kStCH / 2) + c2];
#pragma unroll
                for (int r = 0; r < kStR; ++r) {
                    const double2 gv = s2[r * (kStCH / 2) + c2];
                    acc[r] = fma(gv.x, xv.x, acc[r]);
                    acc[r] = fma(gv.y, xv.y, acc[r]);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
        if (++stage == kStStages) { stage = 0; phase ^= 1u; }
        if (m.w & 1) {
            // end of the piece: hand the partial sums to the epilogue warp and keep streaming
#pragma unroll
            for (int r = 0; r < kStR; ++r) {
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], off);
            }
            if (lane == 0) {
                mbar_wait(&red_empty[buf], ephase[buf] ^ 1u);     // slot free (always, except pathologically)
#pragma unroll
                for (int r = 0; r < kStR; ++r) red[buf][warp][r] = acc[r];
                if (warp == 0) red_meta[buf] = make_int4(m.x, m.y, m.w >> 1, 0);
                mbar_arrive(&red_full[buf]);                      // release: the stores above are visible
            }
            ephase[buf] ^= 1u;
            buf ^= 1;
#pragma unroll
            for (int r = 0; r < kStR; ++r) acc[r] = 0.0;
        }
    }
    // end of work: pass the end marker on to the epilogue warp
    if (lane == 0) {
        mbar_wait(&red_empty[buf], ephase[buf] ^ 1u);
        if (warp == 0) red_meta[buf] = make_int4(-1, 0, 0, 0);
        mbar_arrive(&red_full[buf]);
    }
}

}  // namespace oq
