// tcgen05.mma rate probe (diagnostic, not part of the library): cycles per MMA instruction for the operand flavours the
// attention kernels use.  One CTA per SM, one warp issues REPS x 8 MMAs (K = 16 each) back to back, then commits and waits.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I vlrlhf_b200/csrc tests/mma_rate_probe.cu -o tests/_bin/mma_rate_probe -lcuda
#include <cstdio>
#include <cstdlib>
#include "ptx.cuh"

using namespace vlb::ptx;
using vlb::smem_u32;

struct Case { int ts, n, b_mn, a_bytes_per_k, two_acc, ncta; const char* name; };

template <int N, bool TS, bool B_MN>
__global__ void __launch_bounds__(128, 1) probe(long long* out, int reps, int two_acc) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                  // [128 rows][128 dh] bf16, two 64-column chunks of [128][128 B]
    uint8_t* sB = smem + 32768;          // up to [256 rows][128 B] x 2 chunks
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (32768 + 65536) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tm = tmem_base_s;
    if (warp == 0) {
        constexpr uint32_t idesc = make_idesc_bf16_f32(128, N, false, B_MN);
        const uint64_t da = make_smem_desc_sw128(smem_u32(sA), 1024, 0);
        const uint64_t db = B_MN ? make_smem_desc_sw128(smem_u32(sB), 1024, 16384) : make_smem_desc_sw128(smem_u32(sB), 1024, 0);
        __syncwarp();
        long long t0 = clock64(), t1 = 0;
        if (elect_one_sync()) {
            for (int r = 0; r < reps; ++r) {
                const uint32_t d = tm + ((two_acc && (r & 1)) ? 256u : 0u);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint32_t off = ((k >> 2) * 16384 + (k & 3) * 32) >> 4;
                    const uint64_t boff = B_MN ? (uint64_t)(k * 128) : (uint64_t)off;
                    if (TS) umma_f16_ts(d, tm + 448 + k * 8, db + boff, idesc, (k != 0 || !two_acc) ? 1u : 0u);
                    else umma_f16_ss(d, da + off, db + boff, idesc, (k != 0 || !two_acc) ? 1u : 0u);
                }
            }
            t1 = clock64();
            umma_commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, 0, 1);
        long long t2 = clock64();
        t1 = __shfl_sync(0xffffffffu, t1, 0);   // (the elected lane is lane 0 for the full mask)
        if (lane == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) { tcgen05_fence_after(); tmem_dealloc(tm, 512); }
}

template <int N, bool TS, bool B_MN>
static void run(const char* name, int reps, int two_acc, int grid) {
    long long* out; cudaMalloc(&out, 16); cudaMemset(out, 0, 16);
    auto k = probe<N, TS, B_MN>;
    const int smem = 32768 + 65536 + 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int it = 0; it < 2; ++it) k<<<grid, 128, smem>>>(out, reps, two_acc);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    const double ideal = 128.0 * N / 256.0;
    printf("%-58s grid %3d: issue %7.1f  complete %7.1f cycles per MMA (tensor floor %5.1f)%s\n", name, grid, (double)h[0] / (reps * 8),
           (double)h[1] / (reps * 8), ideal, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(out);
}

int main() {
    const int reps = 64;
    for (int grid : {1, 148}) {
        run<64, false, false>("SS  M128 N64  B K-major, one accumulator", reps, 0, grid);
        run<64, false, false>("SS  M128 N64  B K-major, two accumulators (8 each)", reps, 1, grid);
        run<64, true, false>("TS  M128 N64  B K-major, two accumulators", reps, 1, grid);
        run<128, false, false>("SS  M128 N128 B K-major, two accumulators", reps, 1, grid);
        run<128, true, false>("TS  M128 N128 B K-major, one accumulator", reps, 0, grid);
        run<128, false, true>("SS  M128 N128 B MN-major, one accumulator", reps, 0, grid);
        run<128, true, true>("TS  M128 N128 B MN-major, one accumulator", reps, 0, grid);
        run<128, true, true>("TS  M128 N128 B MN-major, two accumulators", reps, 1, grid);
        run<256, false, false>("SS  M128 N256 B K-major, one accumulator", reps, 0, grid);
        run<256, true, false>("TS  M128 N256 B K-major, one accumulator", reps, 0, grid);
        run<256, true, true>("TS  M128 N256 B MN-major, one accumulator", reps, 0, grid);
    }
    return 0;
}
