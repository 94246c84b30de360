// comm.cu -- window allocation and CUDA-IPC peer mapping (see comm.cuh for the protocol).
#include "comm.cuh"

#include <cuda.h>

#include "common.cuh"
#include "problem.cuh"

namespace oq {

constexpr size_t kSegPad = 4096;   // one maximal matvec segment of slack behind each forcing vector

static WindowLayout make_layout(int nf, int ne)
{
    WindowLayout w;
    w.relv_len = round_up((size_t)nf, 16) + kSegPad;
    w.reldeps_len = round_up((size_t)6 * (ne > 0 ? ne : 1), 16) + kSegPad;
    size_t off = 0;
    w.off_relv = off;    off += 2 * w.relv_len;
    w.off_reldeps = off; off += 2 * w.reldeps_len;
    w.off_red = off;     off += 2 * kMaxWorld;
    w.off_flags = off;   off += 2 * kMaxWorld;
    w.off_epochs = off;  off += kEpCount;
    w.total = round_up(off, 32);
    return w;
}

int comm_alloc_window(OqProblem* p)
{
    p->wl = make_layout(p->nf, p->ne);
    OQ_TRY(p->window.alloc(p->wl.total));
    OQ_TRY(p->window.zero());
    double* b = p->window.p;
    p->relv = b + p->wl.off_relv;
    p->reldeps = b + p->wl.off_reldeps;
    p->red_slots = b + p->wl.off_red;
    p->flags = reinterpret_cast<unsigned long long*>(b + p->wl.off_flags);
    p->epochs = reinterpret_cast<unsigned long long*>(b + p->wl.off_epochs);
    // peer-wait policy: bounded spin (default 30 s; OQ_PEER_TIMEOUT_S=0 waits for ever, e.g. under a debugger or
    // with host callbacks that may stall one rank for long) and a host-visible error word
    double tmo = 30.0;
    if (const char* e = getenv("OQ_PEER_TIMEOUT_S")) tmo = atof(e);
    if (tmo < 0) tmo = 0;
    unsigned long long init[kEpCount] = {};
    init[kEpTimeoutNs] = (unsigned long long)(tmo * 1e9);
    if (!p->err_host) {
        if (cudaHostAlloc(reinterpret_cast<void**>(&p->err_host), sizeof(unsigned long long), cudaHostAllocMapped) ==
            cudaSuccess) {
            *p->err_host = 0;
            void* dev = nullptr;
            if (cudaHostGetDevicePointer(&dev, p->err_host, 0) == cudaSuccess) init[kEpHostErr] = (unsigned long long)dev;
        } else {
            cudaGetLastError();
            p->err_host = nullptr;
        }
    }
    OQ_CUDA(cudaMemcpy(p->epochs, init, sizeof(init), cudaMemcpyHostToDevice));
    return 0;
}

int comm_check_error(OqProblem* p, const char* where)
{
    if (p->err_host && *reinterpret_cast<volatile unsigned long long*>(p->err_host))
        return fail("%s: timed out waiting for a peer rank (results of this call are invalid; OQ_PEER_TIMEOUT_S sets the limit)", where);
    return 0;
}

void comm_clear_error(OqProblem* p)
{
    if (p->err_host && *reinterpret_cast<volatile unsigned long long*>(p->err_host)) {
        *p->err_host = 0;
        cudaMemsetAsync(p->epochs + kEpError, 0, sizeof(unsigned long long), p->stream);
    }
}

ColOwners comm_owners(const OqProblem* p)
{
    if (p->peers) return p->peers->own;
    ColOwners o;
    o.world = 1;
    o.fb[0] = 0; o.fb[1] = p->nf;
    o.eb[0] = 0; o.eb[1] = p->ne;
    return o;
}

PeerTargets comm_targets(const OqProblem* p)
{
    if (p->peers) return p->peers->t;
    PeerTargets t;
    t.world = 1; t.rank = 0;
    t.base[0] = p->window.p;
    return t;
}

void comm_release(OqProblem* p)
{
    if (!p->peers) {
        return;
    }
    for (int r = 0; r < p->peers->t.world; ++r)
        if (p->peers->opened[r] && p->peers->ipc_base[r]) cudaIpcCloseMemHandle(p->peers->ipc_base[r]);
    delete p->peers;
    p->peers = nullptr;
}

// what travels between ranks: the IPC handle of the window plus enough metadata to validate the mapping
struct HandleBlob {
    cudaIpcMemHandle_t ipc;       // 64 bytes
    int32_t rank, world;
    int32_t nf, ne, f0, f1, e0, e1;
    uint64_t window_doubles;
    uint64_t base_offset;         // offset of the window inside the exported allocation (cudaMalloc may suballocate)
    int32_t device;
    int32_t magic;
};
static_assert(sizeof(HandleBlob) <= OQ_COMM_HANDLE_BYTES, "handle blob too large");
constexpr int32_t kMagic = 0x4f513230;

}  // namespace oq

using namespace oq;

extern "C" {

int oq_comm_export(OqProblem* p, int rank, int world, uint8_t handle[OQ_COMM_HANDLE_BYTES])
{
    OQ_CHECK(p && handle, "NULL argument");
    OQ_CHECK(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, "bad rank %d / world %d", rank, world);
    OQ_TRY(enter());
    HandleBlob b{};
    OQ_CUDA(cudaIpcGetMemHandle(&b.ipc, p->window.p));
    // the driver entry point is resolved at run time so that the library itself does not link libcuda
    // (it must load on GPU-less build hosts)
    typedef CUresult (*GetRangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    OQ_CUDA(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qres));
    OQ_CHECK(fn && qres == cudaDriverEntryPointSuccess, "cuMemGetAddressRange is unavailable");
    CUdeviceptr abase = 0;
    size_t asize = 0;
    OQ_CHECK(reinterpret_cast<GetRangeFn>(fn)(&abase, &asize, (CUdeviceptr)p->window.p) == CUDA_SUCCESS,
             "cuMemGetAddressRange failed on the window");
    b.base_offset = (uint64_t)((CUdeviceptr)p->window.p - abase);
    b.rank = rank; b.world = world;
    b.nf = p->nf; b.ne = p->ne; b.f0 = p->f0; b.f1 = p->f1; b.e0 = p->e0; b.e1 = p->e1;
    b.window_doubles = p->wl.total;
    b.device = current_device();
    b.magic = kMagic;
    memset(handle, 0, OQ_COMM_HANDLE_BYTES);
    memcpy(handle, &b, sizeof(b));
    p->rank = rank; p->world = world;
    return 0;
}

int oq_comm_connect(OqProblem* p, const uint8_t* all)
{
    OQ_CHECK(p && all, "NULL argument");
    OQ_TRY(enter());
    const int world = p->world, rank = p->rank;
    OQ_CHECK(world >= 1, "call oq_comm_export first");
    comm_release(p);
    PeerWindow* pw = new PeerWindow();
    pw->t.world = world; pw->t.rank = rank;
    pw->own.world = world;
    int fcover = 0, ecover = 0;
    for (int r = 0; r < world; ++r) {
        HandleBlob b;
        memcpy(&b, all + (size_t)r * OQ_COMM_HANDLE_BYTES, sizeof(b));
        int rc = 0;
        if (b.magic != kMagic || b.rank != r || b.world != world) rc = fail("handle %d is malformed", r);
        else if (b.nf != p->nf || b.ne != p->ne || b.window_doubles != p->wl.total)
            rc = fail("rank %d was built for a different problem size", r);
        else if (b.f0 != fcover || b.e0 != ecover)
            rc = fail("row shards are not contiguous in rank order at rank %d", r);
        if (rc) { p->peers = pw; comm_release(p); return rc; }
        pw->own.fb[r] = b.f0; pw->own.fb[r + 1] = b.f1;
        pw->own.eb[r] = b.e0; pw->own.eb[r + 1] = b.e1;
        fcover = b.f1; ecover = b.e1;
        if (r == rank) { pw->t.base[r] = p->window.p; continue; }
        void* ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, b.ipc, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            p->peers = pw; comm_release(p);
            return fail("cudaIpcOpenMemHandle for rank %d failed: %s", r, cudaGetErrorString(e));
        }
        pw->ipc_base[r] = ptr;
        pw->t.base[r] = reinterpret_cast<double*>(static_cast<char*>(ptr) + b.base_offset);
        pw->opened[r] = true;
    }
    if (fcover != p->nf || ecover != p->ne) {
        p->peers = pw; comm_release(p);
        return fail("row shards do not cover the problem (fault %d/%d, mantle %d/%d)", fcover, p->nf, ecover, p->ne);
    }
    p->peers = pw;
    // a resident-RHS graph captured before the peers were mapped bakes single-rank targets into its kernels
    if (p->rhs_graph) { cudaGraphExecDestroy(p->rhs_graph); p->rhs_graph = nullptr; }
    return 0;
}

}  // extern "C"
