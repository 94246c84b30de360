// runtime.cu -- device selection, error string, launch accounting and the two roofline probes.
#include "common.cuh"

#include <cmath>

namespace oq {

std::string& last_error()
{
    thread_local std::string s;
    return s;
}

int fail(const char* fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    last_error() = buf;
    return 1;
}

std::atomic<int64_t> g_launches{0};
static std::atomic<int> g_device{-1};

int current_device() { return g_device.load(); }

static int select_device(int device)
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail("no CUDA device available (%s); liboetqf_b200 has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    OQ_CHECK(device >= 0 && device < count, "device %d out of range (%d devices)", device, count);
    cudaDeviceProp prop;
    OQ_CUDA(cudaGetDeviceProperties(&prop, device));
    OQ_CHECK(prop.major == 10, "device %d is sm_%d%d; this library is built for sm_100a (B200) only",
             device, prop.major, prop.minor);
    OQ_CUDA(cudaSetDevice(device));
    g_device.store(device);
    return 0;
}

int enter()
{
    int dev = g_device.load();
    if (dev < 0) return select_device(0);
    int cur = -1;
    OQ_CUDA(cudaGetDevice(&cur));
    if (cur != dev) OQ_CUDA(cudaSetDevice(dev));
    return 0;
}

void sincosd(double deg, double* s, double* c)
{
    double r = std::fmod(deg, 360.0);
    if (r < 0) r += 360.0;
    if (r == 0.0)   { *s = 0.0;  *c = 1.0;  return; }
    if (r == 90.0)  { *s = 1.0;  *c = 0.0;  return; }
    if (r == 180.0) { *s = 0.0;  *c = -1.0; return; }
    if (r == 270.0) { *s = -1.0; *c = 0.0;  return; }
    const long double a = (long double)deg * 3.14159265358979323846264338327950288L / 180.0L;
    *s = (double)sinl(a);
    *c = (double)cosl(a);
}

// ---- roofline probes ------------------------------------------------------------------------
// 8 independent DFMA chains per thread, 256-long unrolled inner loop: issue-bound on the fp64 pipe.
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double seed)
{
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    double a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.999999, c = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

__global__ void __launch_bounds__(256) copy_kernel(const double2* __restrict__ src, double2* __restrict__ dst, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = src[i];
}

}  // namespace oq

using namespace oq;

extern "C" {

int oq_abi_version(void) { return OQ_ABI_VERSION; }
const char* oq_last_error(void) { return last_error().c_str(); }
int oq_init(int device) { return select_device(device); }

int oq_device_count(int* count)
{
    OQ_CHECK(count, "count is NULL");
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) { *count = 0; return fail("cudaGetDeviceCount: %s", cudaGetErrorString(e)); }
    return 0;
}

int64_t oq_kernel_launch_count(void) { return g_launches.load(); }

int oq_measure_fp64_peak(double* flops_per_s)
{
    OQ_CHECK(flops_per_s, "flops_per_s is NULL");
    OQ_TRY(enter());
    cudaDeviceProp prop;
    OQ_CUDA(cudaGetDeviceProperties(&prop, current_device()));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 2000;
    DevBuf<double> out;
    OQ_TRY(out.alloc((size_t)blocks * threads));
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        EventTimer t;
        OQ_TRY(t.start());
        dfma_peak_kernel<<<blocks, threads>>>(out.p, iters, 1.0 + rep);
        OQ_LAUNCHED();
        double ms = 0;
        OQ_TRY(t.stop(&ms));
        const double flops = 2.0 * 8 * 32 * (double)iters * blocks * threads;
        if (rep > 0 && flops / (ms * 1e-3) > best) best = flops / (ms * 1e-3);
    }
    *flops_per_s = best;
    return 0;
}

int oq_measure_hbm_copy(size_t bytes, double* bytes_per_s)
{
    OQ_CHECK(bytes_per_s, "bytes_per_s is NULL");
    OQ_TRY(enter());
    const size_t n2 = bytes / sizeof(double2);
    OQ_CHECK(n2 > 0, "buffer too small");
    DevBuf<double2> a, b;
    OQ_TRY(a.alloc(n2));
    OQ_TRY(b.alloc(n2));
    OQ_TRY(a.zero());
    cudaDeviceProp prop;
    OQ_CUDA(cudaGetDeviceProperties(&prop, current_device()));
    double best = 0;
    for (int rep = 0; rep < 6; ++rep) {
        EventTimer t;
        OQ_TRY(t.start());
        copy_kernel<<<prop.multiProcessorCount * 16, 256>>>(a.p, b.p, n2);
        OQ_LAUNCHED();
        double ms = 0;
        OQ_TRY(t.stop(&ms));
        const double bw = 2.0 * n2 * sizeof(double2) / (ms * 1e-3);
        if (rep > 0 && bw > best) best = bw;
    }
    *bytes_per_s = best;
    return 0;
}

int oq_host_register(void* ptr, size_t bytes)
{
    OQ_CHECK(ptr && bytes > 0, "bad argument");
    OQ_TRY(enter());
    OQ_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterMapped | cudaHostRegisterPortable));
    return 0;
}

int oq_host_unregister(void* ptr)
{
    OQ_CHECK(ptr, "bad argument");
    OQ_TRY(enter());
    OQ_CUDA(cudaHostUnregister(ptr));
    return 0;
}

}  // extern "C"
