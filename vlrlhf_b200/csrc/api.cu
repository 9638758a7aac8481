// Library-wide plumbing: error string, SM count, launch counter, deterministic synthetic init.
#include <atomic>
#include <stdarg.h>

#include "common.cuh"

namespace vlb {

static thread_local char g_err[1024] = "";
static std::atomic<uint64_t> g_launches{0};

void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

template <typename T>
__global__ void init_uniform_kernel(T* dst, uint64_t n, uint32_t hseed, float scale, float shift) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t h = lowbias32((uint32_t)i ^ hseed);
        float u = (float)(h >> 8) * 5.9604644775390625e-08f;  // * 2^-24 (exact)
        u = __fsub_rn(__fmul_rn(u, 2.0f), 1.0f);
        const float w = __fadd_rn(__fmul_rn(u, scale), shift);
        if constexpr (sizeof(T) == 2) dst[i] = __float2bfloat16_rn(w);
        else dst[i] = w;
    }
}

__global__ void perturb_bf16_kernel(__nv_bfloat16* dst, const __nv_bfloat16* base, const __nv_bfloat16* other, uint64_t n,
                                    float alpha, float shift) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const float b = __bfloat162float(base[i]);
        const float o = __bfloat162float(other[i]);
        // same op order as oracle/restate.py::make_policy_and_ref (no FMA contraction)
        dst[i] = __float2bfloat16_rn(__fadd_rn(b, __fmul_rn(alpha, __fsub_rn(o, shift))));
    }
}

}  // namespace vlb

extern "C" int vlb200_abi_version(void) { return VLB200_ABI_VERSION; }
extern "C" const char* vlb200_last_error(void) { return vlb::g_err; }
extern "C" uint64_t vlb200_launch_count(void) { return vlb::g_launches.load(); }

extern "C" int vlb200_init_uniform(void* dst, int dtype, uint64_t n, uint32_t seed, float scale, float shift,
                                   void* stream) {
    using namespace vlb;
    VLB_REQUIRE(dst != nullptr || n == 0, "init_uniform: null dst");
    VLB_REQUIRE(n < (1ull << 32), "init_uniform: n must be < 2^32 per tensor");
    if (n == 0) return VLB200_OK;
    const uint32_t hseed = lowbias32(seed);
    const int threads = 256;
    const int blocks = (int)((n + threads - 1) / threads < (uint64_t)num_sms() * 16 ? (n + threads - 1) / threads
                                                                                     : (uint64_t)num_sms() * 16);
    if (dtype == VLB200_BF16)
        init_uniform_kernel<<<blocks, threads, 0, as_stream(stream)>>>((__nv_bfloat16*)dst, n, hseed, scale, shift);
    else if (dtype == VLB200_F32)
        init_uniform_kernel<<<blocks, threads, 0, as_stream(stream)>>>((float*)dst, n, hseed, scale, shift);
    else
        VLB_REQUIRE(false, "init_uniform: bad dtype %d", dtype);
    count_launch();
    VLB_LAUNCH_CHECK();
    return VLB200_OK;
}

extern "C" int vlb200_perturb_bf16(void* dst, const void* base, const void* other, uint64_t n, float alpha, float shift,
                                   void* stream) {
    using namespace vlb;
    if (n == 0) return VLB200_OK;
    VLB_REQUIRE(dst && base && other, "perturb: null pointer");
    const int threads = 256;
    const uint64_t want = (n + threads - 1) / threads;
    const int blocks = (int)(want < (uint64_t)num_sms() * 16 ? want : (uint64_t)num_sms() * 16);
    perturb_bf16_kernel<<<blocks, threads, 0, as_stream(stream)>>>((__nv_bfloat16*)dst, (const __nv_bfloat16*)base,
                                                                   (const __nv_bfloat16*)other, n, alpha, shift);
    count_launch();
    VLB_LAUNCH_CHECK();
    return VLB200_OK;
}
