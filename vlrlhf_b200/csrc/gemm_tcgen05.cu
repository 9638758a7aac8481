// bf16 GEMM on the 5th-gen tensor cores: TMA -> 128B-swizzled smem ring -> tcgen05.mma (TMEM
// accumulators, double buffered) -> tcgen05.ld epilogue (bias / activation / residual / accumulate).
//
// Persistent, warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM owner),
// warps 2-5 = epilogue (TMEM lane quadrant = warp_idx % 4).  One CTA per SM, grid = #SMs.
//
// Replaces every nn.Linear the reference reaches through transformers (see include/vlb200.h).
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>
#include <unordered_map>
#include <vector>

#include "ptx.cuh"

namespace vlb {
namespace gemm {

using namespace ptx;

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 192;
constexpr int A_TILE_BYTES = BLOCK_M * BLOCK_K * 2;

struct Params {
    int M, N, K;
    int num_m_blocks, num_n_blocks, num_tiles, num_k_blocks;
    int kb1;           // k-blocks [0, kb1) come from the first operand pair, [kb1, num_k_blocks) from the second (A2, B2)
    float alpha;       // the accumulator is scaled by alpha before bias / activation / residual
    // split-K (1-CTA kernel only): work item t = split * num_tiles + tile covers k-blocks [split * kb_per_split, ...) and
    // stores its raw fp32 partial tile to ws[split][M][N]; splitk_reduce_kernel sums the partials in a fixed order
    int splits, kb_per_split;
    float* ws;
    // SwiGLU pair kernel: B = [gate rows | up rows] ([2*ff, K]); D = act [M, ff]; G (optional) receives the bf16 gate|up
    int ff;
    __nv_bfloat16* G;
    long long ldg;
    void* D;
    long long ldd;
    int out_f32;
    const __nv_bfloat16* bias;
    int act;
    const void* residual;
    int residual_f32;
    long long ldr;
    int accumulate;
    int group;         // rasterisation: tiles of `group` M-blocks (or N-blocks) sweep the other dimension together
    int group_along_n;
};

__host__ __device__ __forceinline__ void tile_coords(const Params& p, int tile, int& m_blk, int& n_blk) {
    // grouped rasterisation: the CTAs running concurrently share a group of `p.group` blocks of one operand (kept hot
    // in L2) while the other operand streams past once per group; the host picks the orientation/size that minimises
    // the DRAM re-reads (r1 ncu: GROUP_M=8 re-streamed B 12.5x).
    if (!p.group_along_n) {
        const int tiles_per_group = p.group * p.num_n_blocks;
        const int g = tile / tiles_per_group;
        const int first_m = g * p.group;
        const int group_m = p.num_m_blocks - first_m < p.group ? p.num_m_blocks - first_m : p.group;
        const int in_group = tile - g * tiles_per_group;
        m_blk = first_m + in_group % group_m;
        n_blk = in_group / group_m;
    } else {
        const int tiles_per_group = p.group * p.num_m_blocks;
        const int g = tile / tiles_per_group;
        const int first_n = g * p.group;
        const int group_n = p.num_n_blocks - first_n < p.group ? p.num_n_blocks - first_n : p.group;
        const int in_group = tile - g * tiles_per_group;
        n_blk = first_n + in_group % group_n;
        m_blk = in_group / group_n;
    }
}

constexpr int ACT_SWIGLU_BWD = 100;   // internal epilogue mode of vlb200_gemm_swiglu_bwd_bf16 (not an activation of the public API)

__device__ __forceinline__ float apply_act(float x, int act) {
    if (act == VLB200_ACT_QUICK_GELU) return x / (1.0f + __expf(-1.702f * x));
    if (act == VLB200_ACT_GELU_ERF) return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
    return x;
}

// epilogue for 32 consecutive fp32 accumulator columns of one row: bias, activation, residual, accumulate, store
__device__ __forceinline__ void epilogue_store_32(const Params& p, int row, int col0, const uint32_t (&r)[32]) {
    if (row < p.M && col0 < p.N) {
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * p.alpha;
                    if (p.act == ACT_SWIGLU_BWD) {
                        // v = dact (the product dy Wd, never stored).  gate|up at G is read and overwritten IN PLACE by
                        // [dgate | dup]; act = silu(gate) * up goes to D when given.  dact is rounded to bf16 first and the
                        // arithmetic is that of swiglu_fwd_kernel / swiglu_bwd_kernel (elementwise.cu), so the result equals
                        // GEMM -> swiglu_fwd + swiglu_bwd bit for bit.
                        __nv_bfloat16* gp = p.G + (long long)row * p.ldg + col0;
                        __nv_bfloat16* ap = p.D != nullptr ? reinterpret_cast<__nv_bfloat16*>(p.D) + (long long)row * p.ldd + col0 : nullptr;
#pragma unroll
                        for (int j8 = 0; j8 < 4; ++j8) {
                            if (col0 + j8 * 8 < p.N) {
                                const uint4 gb = *reinterpret_cast<const uint4*>(gp + j8 * 8);
                                const uint4 ub = *reinterpret_cast<const uint4*>(gp + p.ff + j8 * 8);
                                const uint32_t gw[4] = {gb.x, gb.y, gb.z, gb.w}, uw[4] = {ub.x, ub.y, ub.z, ub.w};
                                uint32_t dgw[4], duw[4], aw[4];
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    const float2 g = unpack_bf16x2(gw[q]), u = unpack_bf16x2(uw[q]);
                                    const float2 d = unpack_bf16x2(pack_bf16x2(v[j8 * 8 + 2 * q], v[j8 * 8 + 2 * q + 1]));
                                    const float sx = 1.f / (1.f + __expf(-g.x)), sy = 1.f / (1.f + __expf(-g.y));
                                    dgw[q] = pack_bf16x2(d.x * u.x * sx * (1.f + g.x * (1.f - sx)), d.y * u.y * sy * (1.f + g.y * (1.f - sy)));
                                    duw[q] = pack_bf16x2(d.x * g.x * sx, d.y * g.y * sy);
                                    aw[q] = pack_bf16x2(g.x / (1.f + __expf(-g.x)) * u.x, g.y / (1.f + __expf(-g.y)) * u.y);
                                }
                                *reinterpret_cast<uint4*>(gp + j8 * 8) = make_uint4(dgw[0], dgw[1], dgw[2], dgw[3]);
                                *reinterpret_cast<uint4*>(gp + p.ff + j8 * 8) = make_uint4(duw[0], duw[1], duw[2], duw[3]);
                                if (ap != nullptr) *reinterpret_cast<uint4*>(ap + j8 * 8) = make_uint4(aw[0], aw[1], aw[2], aw[3]);
                            }
                        }
                        return;
                    }
                    if (p.bias != nullptr) {
#pragma unroll
                        for (int j8 = 0; j8 < 4; ++j8) {
                            if (col0 + j8 * 8 < p.N) {
                                const uint4 b = *reinterpret_cast<const uint4*>(p.bias + col0 + j8 * 8);
                                const uint32_t bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    const float2 f = unpack_bf16x2(bw[q]);
                                    v[j8 * 8 + 2 * q] += f.x;
                                    v[j8 * 8 + 2 * q + 1] += f.y;
                                }
                            }
                        }
                    }
                    if (p.act != VLB200_ACT_NONE) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], p.act);
                    }
                    if (p.residual != nullptr) {
                        if (p.residual_f32) {
                            const float* rp = reinterpret_cast<const float*>(p.residual) + (long long)row * p.ldr + col0;
#pragma unroll
                            for (int j4 = 0; j4 < 8; ++j4) {
                                if (col0 + j4 * 4 < p.N) {
                                    const float4 b = *reinterpret_cast<const float4*>(rp + j4 * 4);
                                    v[j4 * 4] += b.x; v[j4 * 4 + 1] += b.y; v[j4 * 4 + 2] += b.z; v[j4 * 4 + 3] += b.w;
                                }
                            }
                        } else {
                            const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(p.residual) + (long long)row * p.ldr + col0;
#pragma unroll
                            for (int j8 = 0; j8 < 4; ++j8) {
                                if (col0 + j8 * 8 < p.N) {
                                    const uint4 b = *reinterpret_cast<const uint4*>(rp + j8 * 8);
                                    const uint32_t bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                                    for (int q = 0; q < 4; ++q) {
                                        const float2 f = unpack_bf16x2(bw[q]);
                                        v[j8 * 8 + 2 * q] += f.x;
                                        v[j8 * 8 + 2 * q + 1] += f.y;
                                    }
                                }
                            }
                        }
                    }
                    if (p.out_f32) {
                        float* dp = reinterpret_cast<float*>(p.D) + (long long)row * p.ldd + col0;
#pragma unroll
                        for (int j4 = 0; j4 < 8; ++j4) {
                            if (col0 + j4 * 4 < p.N) {
                                float4 o = make_float4(v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]);
                                if (p.accumulate) {
                                    const float4 old = *reinterpret_cast<const float4*>(dp + j4 * 4);
                                    o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                                }
                                *reinterpret_cast<float4*>(dp + j4 * 4) = o;
                            }
                        }
                    } else {
                        __nv_bfloat16* dp = reinterpret_cast<__nv_bfloat16*>(p.D) + (long long)row * p.ldd + col0;
#pragma unroll
                        for (int j8 = 0; j8 < 4; ++j8) {
                            if (col0 + j8 * 8 < p.N) {
                                if (p.accumulate) {
                                    const uint4 b = *reinterpret_cast<const uint4*>(dp + j8 * 8);
                                    const uint32_t bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                                    for (int q = 0; q < 4; ++q) {
                                        const float2 f = unpack_bf16x2(bw[q]);
                                        v[j8 * 8 + 2 * q] += f.x;
                                        v[j8 * 8 + 2 * q + 1] += f.y;
                                    }
                                }
                                uint4 o;
                                o.x = pack_bf16x2(v[j8 * 8 + 0], v[j8 * 8 + 1]);
                                o.y = pack_bf16x2(v[j8 * 8 + 2], v[j8 * 8 + 3]);
                                o.z = pack_bf16x2(v[j8 * 8 + 4], v[j8 * 8 + 5]);
                                o.w = pack_bf16x2(v[j8 * 8 + 6], v[j8 * 8 + 7]);
                                *reinterpret_cast<uint4*>(dp + j8 * 8) = o;
                            }
                        }
                    }
    }
}

template <int BLOCK_N, int STAGES, bool A_KMAJOR, bool B_KMAJOR>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                 const __grid_constant__ CUtensorMap tma_a2, const __grid_constant__ CUtensorMap tma_b2, const Params p) {
    constexpr int B_TILE_BYTES = BLOCK_N * BLOCK_K * 2;
    constexpr uint32_t STAGE_TX_BYTES = A_TILE_BYTES + B_TILE_BYTES;
    constexpr int TMEM_COLS = 2 * BLOCK_N;  // two accumulator stages
    static_assert(TMEM_COLS == 512 || TMEM_COLS == 256 || TMEM_COLS == 128, "TMEM allocation must be a power of two");

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // manual 1024-byte alignment (SWIZZLE_128B atoms are 1024 B)
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * A_TILE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_b + STAGES * B_TILE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

    const int warp_idx = threadIdx.x >> 5;
    const int lane_idx = threadIdx.x & 31;

    if (warp_idx == 0 && lane_idx == 0) {
        prefetch_tensormap(&tma_a);
        prefetch_tensormap(&tma_b);
        if (p.kb1 < p.num_k_blocks) {
            prefetch_tensormap(&tma_a2);
            prefetch_tensormap(&tma_b2);
        }
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full_bar[a], 1);
            mbar_init(&tmem_empty_bar[a], 4);  // one arrive per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp_idx == 1) tmem_alloc(tmem_base_smem, TMEM_COLS);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;

    if (warp_idx == 0) {
        // ===================== TMA producer =====================
        if (lane_idx == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < p.num_tiles * p.splits; t += gridDim.x) {
                int m_blk, n_blk;
                const int tile = t % p.num_tiles, split = t / p.num_tiles;
                tile_coords(p, tile, m_blk, n_blk);
                const int m0 = m_blk * BLOCK_M, n0 = n_blk * BLOCK_N;
                const int kb_begin = split * p.kb_per_split, kb_end = min(p.num_k_blocks, kb_begin + p.kb_per_split);
                for (int kb = kb_begin; kb < kb_end; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1, 100 + stage);
                    mbar_arrive_expect_tx(&full_bar[stage], STAGE_TX_BYTES);
                    uint8_t* sa = smem_a + stage * A_TILE_BYTES;
                    uint8_t* sb = smem_b + stage * B_TILE_BYTES;
                    // second K segment (A2, B2): same tile format, its own tensor maps; a ragged segment end is
                    // zero-filled by TMA, so the segments need not be multiples of BLOCK_K
                    const bool seg2 = kb >= p.kb1;
                    const CUtensorMap* ma = seg2 ? &tma_a2 : &tma_a;
                    const CUtensorMap* mb = seg2 ? &tma_b2 : &tma_b;
                    const int k0 = (seg2 ? kb - p.kb1 : kb) * BLOCK_K;
                    if constexpr (A_KMAJOR) {
                        tma_load_2d(ma, &full_bar[stage], sa, k0, m0);  // box {64 k, 128 m}
                    } else {
#pragma unroll
                        for (int i = 0; i < BLOCK_M / 64; ++i)  // box {64 m, 64 k}
                            tma_load_2d(ma, &full_bar[stage], sa + i * (BLOCK_K * 128), m0 + i * 64, k0);
                    }
                    if constexpr (B_KMAJOR) {
                        tma_load_2d(mb, &full_bar[stage], sb, k0, n0);  // box {64 k, BLOCK_N n}
                    } else {
#pragma unroll
                        for (int i = 0; i < BLOCK_N / 64; ++i)  // box {64 n, 64 k}
                            tma_load_2d(mb, &full_bar[stage], sb + i * (BLOCK_K * 128), n0 + i * 64, k0);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp_idx == 1) {
        // ===================== MMA issuer =====================
        if (lane_idx == 0) {
            constexpr uint32_t idesc = make_idesc_bf16_f32(BLOCK_M, BLOCK_N, !A_KMAJOR, !B_KMAJOR);
            // K-major: SBO = 8 rows * 128 B; advance 32 B per UMMA_K.  MN-major: LBO = chunk pitch,
            // SBO = 8 k-rows * 128 B; advance 16 k-rows * 128 B per UMMA_K.
            constexpr uint32_t LBO = BLOCK_K * 128;
            constexpr uint32_t A_KADV = A_KMAJOR ? (UMMA_K * 2) >> 4 : (UMMA_K * 128) >> 4;
            constexpr uint32_t B_KADV = B_KMAJOR ? (UMMA_K * 2) >> 4 : (UMMA_K * 128) >> 4;
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int t = blockIdx.x; t < p.num_tiles * p.splits; t += gridDim.x) {
                const int split = t / p.num_tiles;
                const int kb_begin = split * p.kb_per_split, kb_end = min(p.num_k_blocks, kb_begin + p.kb_per_split);
                mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1, 200 + acc);
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
                for (int kb = kb_begin; kb < kb_end; ++kb) {
                    mbar_wait(&full_bar[stage], phase, 300 + stage);
                    tcgen05_fence_after();
                    const uint64_t a_desc =
                        make_smem_desc_sw128(smem_u32(smem_a + stage * A_TILE_BYTES), 1024, A_KMAJOR ? 0 : LBO);
                    const uint64_t b_desc =
                        make_smem_desc_sw128(smem_u32(smem_b + stage * B_TILE_BYTES), 1024, B_KMAJOR ? 0 : LBO);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        umma_f16_ss(tmem_d, a_desc + (uint64_t)(k * A_KADV), b_desc + (uint64_t)(k * B_KADV), idesc,
                                    ((kb - kb_begin) | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tmem_full_bar[acc]);  // accumulator complete -> epilogue
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue (4 warps) =====================
        const int quad = warp_idx & 3;  // TMEM lane quadrant this warp may touch
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int t = blockIdx.x; t < p.num_tiles * p.splits; t += gridDim.x) {
            int m_blk, n_blk;
            const int tile = t % p.num_tiles, split = t / p.num_tiles;
            tile_coords(p, tile, m_blk, n_blk);
            const int row = m_blk * BLOCK_M + quad * 32 + lane_idx;
            const int n0 = n_blk * BLOCK_N;
            mbar_wait(&tmem_full_bar[acc], acc_phase, 400 + acc);
            tcgen05_fence_after();
            const uint32_t taddr0 = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BLOCK_N;
#pragma unroll 1
            for (int c = 0; c < BLOCK_N / 32; ++c) {
                uint32_t r[32];
                tmem_ld_32x32(taddr0 + c * 32, r);
                tmem_ld_wait();
                if (p.splits == 1) {
                    epilogue_store_32(p, row, n0 + c * 32, r);
                } else if (row < p.M) {   // raw partial sums; N % 8 == 0 keeps every float4 inside the row
                    float* wp = p.ws + ((long long)split * p.M + row) * p.N + n0 + c * 32;
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4)
                        if (n0 + c * 32 + j4 * 4 < p.N)
                            *reinterpret_cast<float4*>(wp + j4 * 4) =
                                make_float4(__uint_as_float(r[j4 * 4]), __uint_as_float(r[j4 * 4 + 1]),
                                            __uint_as_float(r[j4 * 4 + 2]), __uint_as_float(r[j4 * 4 + 3]));
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane_idx == 0) mbar_arrive(&tmem_empty_bar[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp_idx == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// D = epilogue(alpha * sum_s ws[s]) for a split-K launch: partials summed in split order (deterministic), then the same
// residual / accumulate / output-dtype handling as the fused epilogue (bias and activation are not combined with split-K)
__global__ void splitk_reduce_kernel(const Params p) {
    const long long n4 = (long long)p.N / 4, total = (long long)p.M * n4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / n4;
        const int col = (int)(i - row * n4) * 4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s = 0; s < p.splits; ++s) {
            const float4 v = *reinterpret_cast<const float4*>(p.ws + ((long long)s * p.M + row) * p.N + col);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        float v[4] = {acc.x * p.alpha, acc.y * p.alpha, acc.z * p.alpha, acc.w * p.alpha};
        if (p.residual != nullptr) {
            if (p.residual_f32) {
                const float4 b = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.residual) + row * p.ldr + col);
                v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
            } else {
                const uint2 b = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p.residual) + row * p.ldr + col);
                const float2 f0 = unpack_bf16x2(b.x), f1 = unpack_bf16x2(b.y);
                v[0] += f0.x; v[1] += f0.y; v[2] += f1.x; v[3] += f1.y;
            }
        }
        if (p.out_f32) {
            float* dp = reinterpret_cast<float*>(p.D) + row * p.ldd + col;
            float4 o = make_float4(v[0], v[1], v[2], v[3]);
            if (p.accumulate) {
                const float4 old = *reinterpret_cast<const float4*>(dp);
                o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
            }
            *reinterpret_cast<float4*>(dp) = o;
        } else {
            __nv_bfloat16* dp = reinterpret_cast<__nv_bfloat16*>(p.D) + row * p.ldd + col;
            if (p.accumulate) {
                const uint2 b = *reinterpret_cast<const uint2*>(dp);
                const float2 f0 = unpack_bf16x2(b.x), f1 = unpack_bf16x2(b.y);
                v[0] += f0.x; v[1] += f0.y; v[2] += f1.x; v[3] += f1.y;
            }
            uint2 o;
            o.x = pack_bf16x2(v[0], v[1]);
            o.y = pack_bf16x2(v[2], v[3]);
            *reinterpret_cast<uint2*>(dp) = o;
        }
    }
}

// ------------------------------------------------------------------------------------------
// 2-CTA variant: a cluster of two CTAs (one TPC) computes a 256 x 256 tile with tcgen05.mma.cta_group::2.
// Each CTA loads its own 128 rows of A and HALF of the B tile (128 of the 256 N rows); the pair's MMA reads both
// halves, so B traffic per FLOP halves (r1 profile: the 1-CTA kernel is L2->SM bandwidth hungry at 85 FLOP/B).
//   * full[stage]   lives in the leader CTA (rank 0): 2 arrivals (leader arrive.expect_tx of both CTAs' bytes, peer remote
//                    arrive); both CTAs' TMA loads complete_tx on it (cta_group::2 loads with the peer bit cleared)
//   * empty[stage], tmem_full[acc]: one per CTA, signalled by a multicast tcgen05.commit from the leader's MMA thread
//   * tmem_empty[acc]: in the leader, 8 arrivals (4 epilogue warps of each CTA; the peer arrives remotely via mapa)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(const void* map, uint64_t* leader_bar_local_alias, void* smem_dst, int c0, int c1) {
    const uint32_t bar = smem_u32(leader_bar_local_alias) & 0xFEFFFFFFu;  // clear the peer bit: CTA 0's barrier
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar_local_alias, uint32_t cta) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(smem_u32(bar_local_alias)), "r"(cta)
        : "memory");
}
__device__ __forceinline__ void umma_f16_ss_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {  // arrives at this smem offset in BOTH CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

constexpr int PAIR_M = 256, PAIR_N = 256, HALF_N = 128;

// SwiGLU epilogue for 32 consecutive columns of one row: g, u are rounded to bf16 first and act is computed from the rounded
// values with the arithmetic of swiglu_fwd_kernel (elementwise.cu), so the fused result equals GEMM -> swiglu_fwd bit for
// bit and the backward's recompute from the saved gate|up stays consistent (modeling_llama.py:182-184).
__device__ __forceinline__ void epilogue_swiglu_32(const Params& p, int row, int col0, const uint32_t (&rg)[32],
                                                   const uint32_t (&ru)[32]) {
    if (row >= p.M || col0 >= p.ff) return;
    __nv_bfloat16* ap = reinterpret_cast<__nv_bfloat16*>(p.D) + (long long)row * p.ldd + col0;
    __nv_bfloat16* gp = p.G != nullptr ? p.G + (long long)row * p.ldg + col0 : nullptr;
#pragma unroll
    for (int j8 = 0; j8 < 4; ++j8) {
        if (col0 + j8 * 8 < p.ff) {
            uint4 gb, ub, ob;
            uint32_t* gw = reinterpret_cast<uint32_t*>(&gb);
            uint32_t* uw = reinterpret_cast<uint32_t*>(&ub);
            uint32_t* ow = reinterpret_cast<uint32_t*>(&ob);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                gw[q] = pack_bf16x2(__uint_as_float(rg[j8 * 8 + 2 * q]), __uint_as_float(rg[j8 * 8 + 2 * q + 1]));
                uw[q] = pack_bf16x2(__uint_as_float(ru[j8 * 8 + 2 * q]), __uint_as_float(ru[j8 * 8 + 2 * q + 1]));
                const float2 g = unpack_bf16x2(gw[q]), u = unpack_bf16x2(uw[q]);
                ow[q] = pack_bf16x2(g.x / (1.f + __expf(-g.x)) * u.x, g.y / (1.f + __expf(-g.y)) * u.y);
            }
            *reinterpret_cast<uint4*>(ap + j8 * 8) = ob;
            if (gp != nullptr) {
                *reinterpret_cast<uint4*>(gp + j8 * 8) = gb;
                *reinterpret_cast<uint4*>(gp + p.ff + j8 * 8) = ub;
            }
        }
    }
}

template <int STAGES, bool A_KMAJOR, bool B_KMAJOR, bool SWIGLU = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_2cta_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                      const __grid_constant__ CUtensorMap tma_a2, const __grid_constant__ CUtensorMap tma_b2, const Params p) {
    constexpr int B_TILE_BYTES = HALF_N * BLOCK_K * 2;
    constexpr uint32_t STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;  // per CTA
    constexpr int TMEM_COLS = 2 * PAIR_N;

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * A_TILE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_b + STAGES * B_TILE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

    const int warp_idx = threadIdx.x >> 5;
    const int lane_idx = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

    if (warp_idx == 0 && lane_idx == 0) {
        prefetch_tensormap(&tma_a);
        prefetch_tensormap(&tma_b);
        if (p.kb1 < p.num_k_blocks) {
            prefetch_tensormap(&tma_a2);
            prefetch_tensormap(&tma_b2);
        }
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 2);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full_bar[a], 1);
            mbar_init(&tmem_empty_bar[a], 8);
        }
        fence_barrier_init();
    }
    cluster_sync_all();  // barriers of both CTAs initialised before any remote arrive / multicast
    if (warp_idx == 1) tmem_alloc_2cta(tmem_base_smem, TMEM_COLS);
    tcgen05_fence_before();
    cluster_sync_all();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;

    if (warp_idx == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (lane_idx == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = pair; tile < p.num_tiles; tile += n_pairs) {
                int m_blk, n_blk;
                tile_coords(p, tile, m_blk, n_blk);
                // SWIGLU: the pair's 256 accumulator columns are 128 gate columns (CTA 0's half of B) next to the SAME 128 up
                // columns (CTA 1's half, p.ff rows further down the fused gate|up weight)
                const int m0 = m_blk * PAIR_M + (int)rank * BLOCK_M;
                const int n0 = SWIGLU ? n_blk * HALF_N + (int)rank * p.ff : n_blk * PAIR_N + (int)rank * HALF_N;
                for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1, 100 + stage);
                    if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
                    uint8_t* sa = smem_a + stage * A_TILE_BYTES;
                    uint8_t* sb = smem_b + stage * B_TILE_BYTES;
                    const bool seg2 = kb >= p.kb1;   // second K segment (A2, B2), see the 1-CTA kernel
                    const CUtensorMap* ma = seg2 ? &tma_a2 : &tma_a;
                    const CUtensorMap* mb = seg2 ? &tma_b2 : &tma_b;
                    const int k0 = (seg2 ? kb - p.kb1 : kb) * BLOCK_K;
                    if constexpr (A_KMAJOR) {
                        tma_load_2d_2sm(ma, &full_bar[stage], sa, k0, m0);
                    } else {
#pragma unroll
                        for (int i = 0; i < BLOCK_M / 64; ++i) tma_load_2d_2sm(ma, &full_bar[stage], sa + i * (BLOCK_K * 128), m0 + i * 64, k0);
                    }
                    if constexpr (B_KMAJOR) {
                        tma_load_2d_2sm(mb, &full_bar[stage], sb, k0, n0);
                    } else {
#pragma unroll
                        for (int i = 0; i < HALF_N / 64; ++i) tma_load_2d_2sm(mb, &full_bar[stage], sb + i * (BLOCK_K * 128), n0 + i * 64, k0);
                    }
                    if (!leader) mbar_arrive_remote(&full_bar[stage], 0);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp_idx == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader && lane_idx == 0) {
            constexpr uint32_t idesc = make_idesc_bf16_f32(PAIR_M, PAIR_N, !A_KMAJOR, !B_KMAJOR);
            constexpr uint32_t LBO = BLOCK_K * 128;
            constexpr uint32_t A_KADV = A_KMAJOR ? (UMMA_K * 2) >> 4 : (UMMA_K * 128) >> 4;
            constexpr uint32_t B_KADV = B_KMAJOR ? (UMMA_K * 2) >> 4 : (UMMA_K * 128) >> 4;
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = pair; tile < p.num_tiles; tile += n_pairs) {
                mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1, 200 + acc);
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + acc * PAIR_N;
                for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase, 300 + stage);
                    tcgen05_fence_after();
                    const uint64_t a_desc = make_smem_desc_sw128(smem_u32(smem_a + stage * A_TILE_BYTES), 1024, A_KMAJOR ? 0 : LBO);
                    const uint64_t b_desc = make_smem_desc_sw128(smem_u32(smem_b + stage * B_TILE_BYTES), 1024, B_KMAJOR ? 0 : LBO);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                        umma_f16_ss_2cta(tmem_d, a_desc + (uint64_t)(k * A_KADV), b_desc + (uint64_t)(k * B_KADV), idesc, (kb | k) != 0 ? 1u : 0u);
                    umma_commit_2cta(&empty_bar[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit_2cta(&tmem_full_bar[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue (4 warps per CTA; this CTA owns rows [m0 + rank*128, +128)) =====================
        const int quad = warp_idx & 3;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = pair; tile < p.num_tiles; tile += n_pairs) {
            int m_blk, n_blk;
            tile_coords(p, tile, m_blk, n_blk);
            const int row = m_blk * PAIR_M + (int)rank * BLOCK_M + quad * 32 + lane_idx;
            const int n0 = n_blk * PAIR_N;
            mbar_wait(&tmem_full_bar[acc], acc_phase, 400 + acc);
            tcgen05_fence_after();
            const uint32_t taddr0 = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * PAIR_N;
            if constexpr (SWIGLU) {
#pragma unroll 1
                for (int c = 0; c < HALF_N / 32; ++c) {
                    uint32_t rg[32], ru[32];
                    tmem_ld_32x32(taddr0 + c * 32, rg);
                    tmem_ld_32x32(taddr0 + HALF_N + c * 32, ru);
                    tmem_ld_wait();
                    epilogue_swiglu_32(p, row, n_blk * HALF_N + c * 32, rg, ru);
                }
            } else {
#pragma unroll 1
                for (int c = 0; c < PAIR_N / 32; ++c) {
                    uint32_t r[32];
                    tmem_ld_32x32(taddr0 + c * 32, r);
                    tmem_ld_wait();
                    epilogue_store_32(p, row, n0 + c * 32, r);
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane_idx == 0) {
                if (leader) mbar_arrive(&tmem_empty_bar[acc]);
                else mbar_arrive_remote(&tmem_empty_bar[acc], 0);
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tcgen05_fence_before();
    cluster_sync_all();  // the peer's smem/TMEM must stay alive until the leader's last MMA and commits have retired
    if (warp_idx == 1) {
        tcgen05_fence_after();
        tmem_dealloc_2cta(tmem_base, TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------
// host side: TMA descriptor cache + dispatch
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    });
    return fn;
}

struct MapKey {
    const void* ptr;
    uint64_t inner, outer, ld;
    uint32_t box_inner, box_outer;
    bool operator==(const MapKey& o) const {
        return ptr == o.ptr && inner == o.inner && outer == o.outer && ld == o.ld && box_inner == o.box_inner &&
               box_outer == o.box_outer;
    }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        size_t h = reinterpret_cast<size_t>(k.ptr);
        auto mix = [&](uint64_t v) { h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2); };
        mix(k.inner); mix(k.outer); mix(k.ld); mix(k.box_inner); mix(k.box_outer);
        return h;
    }
};

// 2-D bf16 tensor map, 128B swizzle; inner = contiguous dimension
int get_tensor_map(const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                   uint32_t box_outer, CUtensorMap* out) {
    static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
    static std::mutex mu;
    MapKey key{ptr, inner, outer, ld, box_inner, box_outer};
    {
        std::lock_guard<std::mutex> g(mu);
        auto it = cache.find(key);
        if (it != cache.end()) { *out = it->second; return VLB200_OK; }
    }
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) { set_last_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)"); return VLB200_ERR_CUDA; }
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMap m;
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_error("cuTensorMapEncodeTiled failed (%d): ptr=%p inner=%llu outer=%llu ld=%llu box=%ux%u", (int)r, ptr,
                       (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld, box_inner, box_outer);
        return VLB200_ERR_CUDA;
    }
    {
        std::lock_guard<std::mutex> g(mu);
        if (cache.size() > 65536) cache.clear();
        cache[key] = m;
    }
    *out = m;
    return VLB200_OK;
}

template <int BLOCK_N, int STAGES, bool A_KMAJOR, bool B_KMAJOR>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& ta2, const CUtensorMap& tb2,
                  const Params& p, cudaStream_t stream) {
    constexpr int smem_bytes = STAGES * (A_TILE_BYTES + BLOCK_N * BLOCK_K * 2) + 256 + 1024;
    auto kern = gemm_bf16_kernel<BLOCK_N, STAGES, A_KMAJOR, B_KMAJOR>;
    static bool configured = false;
    if (!configured) {
        VLB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        configured = true;
    }
    const int work = p.num_tiles * p.splits;
    const int grid = work < num_sms() ? work : num_sms();
    kern<<<grid, NUM_THREADS, smem_bytes, stream>>>(ta, tb, ta2, tb2, p);
    count_launch();
    VLB_LAUNCH_CHECK();
    if (p.splits > 1) {
        const long long total = (long long)p.M * (p.N / 4);
        const int blocks = (int)((total + 255) / 256 < 4LL * num_sms() ? (total + 255) / 256 : 4LL * num_sms());
        splitk_reduce_kernel<<<blocks, 256, 0, stream>>>(p);
        count_launch();
        VLB_LAUNCH_CHECK();
    }
    return VLB200_OK;
}

template <int STAGES, bool A_KMAJOR, bool B_KMAJOR, bool SWIGLU = false>
static int launch_2cta(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& ta2, const CUtensorMap& tb2,
                       const Params& p, cudaStream_t stream) {
    constexpr int smem_bytes = STAGES * (A_TILE_BYTES + HALF_N * BLOCK_K * 2) + 256 + 1024;
    auto kern = gemm_bf16_2cta_kernel<STAGES, A_KMAJOR, B_KMAJOR, SWIGLU>;
    static bool configured = false;
    if (!configured) {
        VLB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        configured = true;
    }
    const int max_pairs = num_sms() / 2;
    const int pairs = p.num_tiles < max_pairs ? p.num_tiles : max_pairs;
    kern<<<2 * pairs, NUM_THREADS, smem_bytes, stream>>>(ta, tb, ta2, tb2, p);
    count_launch();
    VLB_LAUNCH_CHECK();
    return VLB200_OK;
}

static int dispatch_2cta(bool ak, bool bk, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& ta2,
                         const CUtensorMap& tb2, const Params& p, cudaStream_t s) {
    if (ak && bk) return launch_2cta<6, true, true>(ta, tb, ta2, tb2, p, s);
    if (ak && !bk) return launch_2cta<6, true, false>(ta, tb, ta2, tb2, p, s);
    if (!ak && bk) return launch_2cta<6, false, true>(ta, tb, ta2, tb2, p, s);
    return launch_2cta<6, false, false>(ta, tb, ta2, tb2, p, s);
}

// split-K partials: one device buffer, grown on demand (warm-up only; a grow synchronises the device first because
// launches already queued may still be writing the old buffer)
static float* g_splitk_ws = nullptr;
static size_t g_splitk_ws_bytes = 0;
static int splitk_workspace(size_t bytes, float** out) {
    static std::mutex mu;
    std::lock_guard<std::mutex> g(mu);
    if (bytes > g_splitk_ws_bytes) {
        VLB_CHECK_CUDA(cudaDeviceSynchronize());
        if (g_splitk_ws) VLB_CHECK_CUDA(cudaFree(g_splitk_ws));
        g_splitk_ws = nullptr; g_splitk_ws_bytes = 0;
        const size_t want = bytes < (size_t(64) << 20) ? (size_t(64) << 20) : bytes;
        VLB_CHECK_CUDA(cudaMalloc(&g_splitk_ws, want));
        g_splitk_ws_bytes = want;
    }
    *out = g_splitk_ws;
    return VLB200_OK;
}

// Rasterisation: tiles of `group` blocks of one operand sweep the other dimension together, so that group (budget MB of
// operand panels) stays L2-resident while the other operand streams past once per group.  The orientation/size with the
// fewest modelled DRAM re-reads wins.  Budget: vlb200_set_gemm_raster_mb() > VLB200_RASTER_MB > 32 (r1d probe: 32 MB is the
// best single setting at the config-2 shapes).
static double g_raster_mb = -1.0;
static void choose_raster(Params& p, double M, double N, double Kt, int TM, int TN) {
    if (g_raster_mb <= 0.0) {
        const char* e = getenv("VLB200_RASTER_MB");
        g_raster_mb = e && atof(e) > 0.0 ? atof(e) : 32.0;
    }
    const double budget = g_raster_mb * 1024 * 1024;
    const double a_blk = TM * Kt * 2, b_blk = TN * Kt * 2;
    const double a_bytes = M * Kt * 2, b_bytes = N * Kt * 2;
    int gm = (int)(budget / a_blk); gm = gm < 4 ? 4 : gm; gm = gm > p.num_m_blocks ? p.num_m_blocks : gm;
    int gn = (int)(budget / b_blk); gn = gn < 2 ? 2 : gn; gn = gn > p.num_n_blocks ? p.num_n_blocks : gn;
    const double cost_m = a_bytes + b_bytes * ((p.num_m_blocks + gm - 1) / gm);
    const double cost_n = b_bytes + a_bytes * ((p.num_n_blocks + gn - 1) / gn);
    p.group_along_n = cost_n < cost_m;
    p.group = p.group_along_n ? gn : gm;
}
// ---- policy 1 (off by default; VLB200_RASTER_POLICY=model or vlb200_set_gemm_raster_policy(1)): pick (orientation, group) by
// replaying the tile schedule against an LRU model of L2 -- tests/raster_model.py is the Python twin and
// profiles/r1d_raster_model.md the fit.  `conc` tiles run at a time and walk K in lockstep (32 chunks here: the result hardly
// depend on the k granularity); per chunk a tile touches one slab of A and one of B, a finished tile streams its D through the
// (write-allocating) cache.  An effective capacity of ~60 MB -- half of the 126 MB, as if each L2 partition kept its own copy --
// reproduces the ncu DRAM reads of all five probed shapes at the default raster.  Returns the modelled DRAM read bytes.
static double lru_model_read_bytes(int num_m, int num_n, double Kt, int TM, int TN, double d_tile_bytes, int group, int along_n,
                                   double cap_bytes, int conc) {
    constexpr int NCH = 32;   // K chunks of the lockstep walk (64 k-blocks at K = 4096: the choice is stable from ~16 up)
    const double a_slab = TM * (Kt / NCH) * 2, b_slab = TN * (Kt / NCH) * 2;
    const int nA = num_m * NCH, nB = num_n * NCH, nD = num_m * num_n, n = nA + nB + nD;
    std::vector<int> prev(n + 1, -1), next(n + 1, -1);   // intrusive LRU list over dense keys; n = sentinel (head.next = oldest)
    std::vector<char> in(n, 0);
    const int H = n;
    prev[H] = next[H] = H;
    double used = 0.0, miss_bytes = 0.0;
    auto size_of = [&](int k) { return k < nA ? a_slab : (k < nA + nB ? b_slab : d_tile_bytes); };
    auto unlink = [&](int k) { next[prev[k]] = next[k]; prev[next[k]] = prev[k]; };
    auto push_new = [&](int k) { prev[k] = prev[H]; next[k] = H; next[prev[H]] = k; prev[H] = k; };
    auto touch = [&](int k, bool count) {
        if (in[k]) { unlink(k); push_new(k); return; }
        if (count) miss_bytes += size_of(k);
        in[k] = 1; used += size_of(k); push_new(k);
        while (used > cap_bytes && next[H] != H) {
            const int old = next[H];
            unlink(old); in[old] = 0; used -= size_of(old);
        }
    };
    Params q{};
    q.num_m_blocks = num_m; q.num_n_blocks = num_n; q.group = group; q.group_along_n = along_n;
    const int tiles = num_m * num_n;
    std::vector<int> wm(conc), wn(conc);
    for (int w0 = 0; w0 < tiles; w0 += conc) {
        const int cnt = tiles - w0 < conc ? tiles - w0 : conc;
        for (int i = 0; i < cnt; ++i) tile_coords(q, w0 + i, wm[i], wn[i]);
        for (int c = 0; c < NCH; ++c)
            for (int i = 0; i < cnt; ++i) { touch(wm[i] * NCH + c, true); touch(nA + wn[i] * NCH + c, true); }
        for (int i = 0; i < cnt; ++i) touch(nA + nB + wm[i] * num_n + wn[i], false);
    }
    return miss_bytes;
}

static int g_raster_policy = -1;   // -1: read VLB200_RASTER_POLICY on first use; 0: L2 budget (choose_raster); 1: LRU model
struct RasterPlan { int group, along_n; double model_bytes; };
static RasterPlan plan_raster_model(double M, double N, int num_m, int num_n, double Kt, int TM, int TN, double d_tile_bytes, int conc) {
    static const double cap_mb = [] { const char* e = getenv("VLB200_L2_MODEL_MB"); return e && atof(e) > 0.0 ? atof(e) : 60.0; }();
    static std::map<std::tuple<int, int, long long, int, int, long long, int>, RasterPlan> cache;   // host threads: one (Python GIL)
    const auto key = std::make_tuple(num_m, num_n, (long long)Kt, TM, TN, (long long)d_tile_bytes, conc);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    static const int cands[] = {1, 2, 3, 4, 5, 6, 8, 9, 10, 12, 16, 20, 24, 32, 43, 50, 64, 86, 1 << 30};
    RasterPlan best{1, 0, -1.0};
    for (int along_n = 0; along_n < 2; ++along_n) {
        const int lim = along_n ? num_n : num_m;
        int last = -1;
        for (int c : cands) {
            const int g = c < lim ? c : lim;
            if (g == last) break;
            last = g;
            // score = the worse of the fitted capacity and 0.85x of it: a resident group sized to the last megabyte falls off a
            // cliff (3x the traffic) if the real capacity is a little smaller, so such plans must not win on paper
            const double b_hi = lru_model_read_bytes(num_m, num_n, Kt, TM, TN, d_tile_bytes, g, along_n, cap_mb * 1024 * 1024, conc);
            const double b_lo = lru_model_read_bytes(num_m, num_n, Kt, TM, TN, d_tile_bytes, g, along_n, 0.85 * cap_mb * 1024 * 1024, conc);
            const double b = b_hi > b_lo ? b_hi : b_lo;
            if (best.model_bytes < 0.0 || b < best.model_bytes * 0.999) best = RasterPlan{g, along_n, b};
        }
    }
    // keep the measured budget rule unless the model's plan wins even at the pessimistic capacity: the rule's short-K plans
    // (a 32 MB resident group) are known to work on the hardware (ncu), and on paper they only lose by margins that depend on
    // the exact capacity; the long-K plans (square waves) win by 25 % at any capacity
    Params cur{};
    cur.num_m_blocks = num_m; cur.num_n_blocks = num_n;
    choose_raster(cur, M, N, Kt, TM, TN);
    const double cur_bytes = lru_model_read_bytes(num_m, num_n, Kt, TM, TN, d_tile_bytes, cur.group, cur.group_along_n,
                                                  cap_mb * 1024 * 1024, conc);
    if (!(best.model_bytes < 0.95 * cur_bytes)) best = RasterPlan{cur.group, cur.group_along_n, cur_bytes};
    cache[key] = best;
    return best;
}
static int raster_policy() {
    if (g_raster_policy < 0) {
        const char* e = getenv("VLB200_RASTER_POLICY");
        g_raster_policy = e && (strcmp(e, "model") == 0 || strcmp(e, "1") == 0) ? 1 : 0;
    }
    return g_raster_policy;
}
// the raster of one pair-kernel launch under the active policy (d_tile_bytes: what a finished tile writes)
static void choose_raster_pair(Params& p, double M, double N, double Kt, double d_tile_bytes) {
    if (raster_policy() == 1) {
        const RasterPlan r = plan_raster_model(M, N, p.num_m_blocks, p.num_n_blocks, Kt, PAIR_M, PAIR_N, d_tile_bytes, num_sms() / 2);
        p.group = r.group; p.group_along_n = r.along_n;
        return;
    }
    choose_raster(p, M, N, Kt, PAIR_M, PAIR_N);
}
static int g_gemm_mode = -1;  // -1: read VLB200_GEMM_2CTA on first use; 0: 1-CTA kernel; 1: 2-CTA pairs where the shape allows

template <int BLOCK_N, int STAGES>
static int dispatch_major(bool ak, bool bk, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& ta2,
                          const CUtensorMap& tb2, const Params& p, cudaStream_t s) {
    if (ak && bk) return launch<BLOCK_N, STAGES, true, true>(ta, tb, ta2, tb2, p, s);
    if (ak && !bk) return launch<BLOCK_N, STAGES, true, false>(ta, tb, ta2, tb2, p, s);
    if (!ak && bk) return launch<BLOCK_N, STAGES, false, true>(ta, tb, ta2, tb2, p, s);
    return launch<BLOCK_N, STAGES, false, false>(ta, tb, ta2, tb2, p, s);
}

}  // namespace gemm
}  // namespace vlb

// gate_up != nullptr: the SwiGLU-backward epilogue (vlb200_gemm_swiglu_bwd_bf16); D may then be null
static int gemm_ex_impl(const void* A, int lda, int a_kmajor, const void* B, int ldb, int b_kmajor,
                        const void* A2, int lda2, const void* B2, int ldb2, int K2, void* D, int ldd,
                        int out_dtype, int M, int N, int K, float alpha, const void* bias, int act,
                        const void* residual, int residual_dtype, int ldr, int accumulate, void* gate_up, long long ld_gu,
                        void* stream) {
    using namespace vlb;
    using namespace vlb::gemm;
    VLB_REQUIRE(A && B && (D || gate_up), "gemm: null pointer");
    VLB_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: bad shape M=%d N=%d K=%d", M, N, K);
    VLB_REQUIRE(N % 8 == 0, "gemm: N=%d must be a multiple of 8", N);
    VLB_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, "gemm: lda=%d / ldb=%d must be multiples of 8 elements (TMA 16 B strides)",
                lda, ldb);
    VLB_REQUIRE(ldd % 8 == 0 && (residual == nullptr || ldr % 8 == 0), "gemm: ldd/ldr must be multiples of 8");
    VLB_REQUIRE(residual == nullptr || residual_dtype == VLB200_BF16 || residual_dtype == VLB200_F32, "gemm: bad residual dtype");
    VLB_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(D) & 15) == 0,
                "gemm: pointers must be 16-byte aligned");
    VLB_REQUIRE(out_dtype == VLB200_BF16 || out_dtype == VLB200_F32, "gemm: bad out_dtype %d", out_dtype);
    VLB_REQUIRE(a_kmajor ? lda >= K : lda >= M, "gemm: lda too small");
    VLB_REQUIRE(b_kmajor ? ldb >= K : ldb >= N, "gemm: ldb too small");
    const bool dual = A2 != nullptr || B2 != nullptr || K2 > 0;
    if (dual) {
        VLB_REQUIRE(A2 && B2 && K2 > 0, "gemm: the second operand pair needs A2, B2 and K2 > 0");
        VLB_REQUIRE(lda2 % 8 == 0 && ldb2 % 8 == 0, "gemm: lda2=%d / ldb2=%d must be multiples of 8 elements", lda2, ldb2);
        VLB_REQUIRE((reinterpret_cast<uintptr_t>(A2) & 15) == 0 && (reinterpret_cast<uintptr_t>(B2) & 15) == 0,
                    "gemm: A2/B2 must be 16-byte aligned");
        VLB_REQUIRE(a_kmajor ? lda2 >= K2 : lda2 >= M, "gemm: lda2 too small");
        VLB_REQUIRE(b_kmajor ? ldb2 >= K2 : ldb2 >= N, "gemm: ldb2 too small");
    }
    if (g_gemm_mode < 0) {
        const char* e = getenv("VLB200_GEMM_2CTA");
        g_gemm_mode = (e != nullptr && e[0] == '0') ? 0 : 1;  // CTA pairs by default
    }
    // skinny outputs with a long contraction (LoRA weight gradients dB = dy^T ts, dA = dt^T x: a handful of tiles, K = all
    // tokens of the batch) would occupy a few SMs only: 128x128 tiles on the 1-CTA kernel, split along K so that every SM
    // streams a slice (r2 probe: 95 us at 1.1 TB/s for M 4096, N 128, K 12792 without it)
    const int kblocks_total = (K + BLOCK_K - 1) / BLOCK_K + (dual ? (K2 + BLOCK_K - 1) / BLOCK_K : 0);
    const long long tiles128 = (long long)((M + 127) / 128) * ((N + 127) / 128);
    static const bool splitk_on = [] { const char* e = getenv("VLB200_SPLITK"); return !(e && e[0] == '0'); }();
    const bool skinny = splitk_on && tiles128 <= num_sms() && kblocks_total >= 32 && bias == nullptr && act == VLB200_ACT_NONE &&
                        N % 4 == 0;
    int splits = 1, kb_per_split = kblocks_total;
    if (skinny) {
        splits = num_sms() / (int)tiles128;
        if (splits > kblocks_total / 8) splits = kblocks_total / 8;
        if (splits > 32) splits = 32;
        if (splits < 1) splits = 1;
        kb_per_split = (kblocks_total + splits - 1) / splits;
        splits = (kblocks_total + kb_per_split - 1) / kb_per_split;   // no empty split
    }
    const bool use_pair = !skinny && g_gemm_mode == 1 && M >= 256 && N >= 256;
    const bool big_n = !skinny && N > 128;
    const int BN = use_pair ? HALF_N : (big_n ? 256 : 128);  // rows of B fetched per TMA box

    CUtensorMap ta, tb, ta2, tb2;
    int rc;
    if (a_kmajor) rc = get_tensor_map(A, K, M, lda, BLOCK_K, BLOCK_M, &ta);
    else rc = get_tensor_map(A, M, K, lda, 64, BLOCK_K, &ta);
    if (rc) return rc;
    if (b_kmajor) rc = get_tensor_map(B, K, N, ldb, BLOCK_K, BN, &tb);
    else rc = get_tensor_map(B, N, K, ldb, 64, BLOCK_K, &tb);
    if (rc) return rc;
    if (dual) {
        if (a_kmajor) rc = get_tensor_map(A2, K2, M, lda2, BLOCK_K, BLOCK_M, &ta2);
        else rc = get_tensor_map(A2, M, K2, lda2, 64, BLOCK_K, &ta2);
        if (rc) return rc;
        if (b_kmajor) rc = get_tensor_map(B2, K2, N, ldb2, BLOCK_K, BN, &tb2);
        else rc = get_tensor_map(B2, N, K2, ldb2, 64, BLOCK_K, &tb2);
        if (rc) return rc;
    } else {
        ta2 = ta; tb2 = tb;  // never dereferenced (kb1 == num_k_blocks)
    }

    Params p;
    p.M = M; p.N = N; p.K = K;
    const int TM = use_pair ? PAIR_M : BLOCK_M, TN = use_pair ? PAIR_N : BN;
    p.num_m_blocks = (M + TM - 1) / TM;
    p.num_n_blocks = (N + TN - 1) / TN;
    p.num_tiles = p.num_m_blocks * p.num_n_blocks;
    p.kb1 = (K + BLOCK_K - 1) / BLOCK_K;
    p.num_k_blocks = p.kb1 + (dual ? (K2 + BLOCK_K - 1) / BLOCK_K : 0);
    p.alpha = alpha;
    p.splits = splits; p.kb_per_split = kb_per_split; p.ws = nullptr;
    p.ff = 0; p.G = nullptr; p.ldg = 0;
    if (gate_up != nullptr) { p.ff = N; p.G = reinterpret_cast<__nv_bfloat16*>(gate_up); p.ldg = ld_gu; }
    if (splits > 1) {
        rc = splitk_workspace((size_t)splits * M * N * sizeof(float), &p.ws);
        if (rc) return rc;
    }
    p.D = D; p.ldd = ldd; p.out_f32 = out_dtype == VLB200_F32;
    p.bias = reinterpret_cast<const __nv_bfloat16*>(bias);
    p.act = act;
    p.residual = residual;
    p.residual_f32 = residual_dtype == VLB200_F32;
    p.ldr = ldr;
    p.accumulate = accumulate;
    if (use_pair)
        vlb::gemm::choose_raster_pair(p, (double)M, (double)N, (double)K + (dual ? K2 : 0),
                                      (double)PAIR_M * PAIR_N * (out_dtype == VLB200_F32 ? 4 : 2));
    else
        vlb::gemm::choose_raster(p, (double)M, (double)N, (double)K + (dual ? K2 : 0), TM, TN);
    cudaStream_t s = as_stream(stream);
    if (use_pair) return dispatch_2cta(a_kmajor != 0, b_kmajor != 0, ta, tb, ta2, tb2, p, s);
    if (big_n) return dispatch_major<256, 4>(a_kmajor != 0, b_kmajor != 0, ta, tb, ta2, tb2, p, s);
    return dispatch_major<128, 6>(a_kmajor != 0, b_kmajor != 0, ta, tb, ta2, tb2, p, s);
}

extern "C" int vlb200_gemm_bf16_ex(const void* A, int lda, int a_kmajor, const void* B, int ldb, int b_kmajor,
                                   const void* A2, int lda2, const void* B2, int ldb2, int K2, void* D, int ldd,
                                   int out_dtype, int M, int N, int K, float alpha, const void* bias, int act,
                                   const void* residual, int residual_dtype, int ldr, int accumulate, void* stream) {
    VLB_REQUIRE(act == VLB200_ACT_NONE || act == VLB200_ACT_QUICK_GELU || act == VLB200_ACT_GELU_ERF, "gemm: unknown activation %d", act);
    return gemm_ex_impl(A, lda, a_kmajor, B, ldb, b_kmajor, A2, lda2, B2, ldb2, K2, D, ldd, out_dtype, M, N, K, alpha, bias, act,
                        residual, residual_dtype, ldr, accumulate, nullptr, 0, stream);
}

extern "C" int vlb200_gemm_swiglu_bwd_bf16(const void* dy, int ld_dy, const void* Wd, int ld_wd, const void* A2, int lda2,
                                           const void* B2, int ldb2, int K2, void* gate_up, int ld_gu, void* act, int ld_act,
                                           int M, int ff, int K, void* stream) {
    VLB_REQUIRE(dy && Wd && gate_up, "gemm_swiglu_bwd: null pointer");
    VLB_REQUIRE(ff % 8 == 0 && ld_gu % 8 == 0 && ld_gu >= 2 * ff && (act == nullptr || (ld_act % 8 == 0 && ld_act >= ff)),
                "gemm_swiglu_bwd: bad leading dimensions");
    VLB_REQUIRE((reinterpret_cast<uintptr_t>(gate_up) & 15) == 0 && (reinterpret_cast<uintptr_t>(act) & 15) == 0,
                "gemm_swiglu_bwd: gate_up / act must be 16-byte aligned");
    // dact[M, ff] = dy[M, K] Wd[K, ff] (Wd = down_proj.weight [d_model, ff] row-major, i.e. an MN-major B operand)
    return gemm_ex_impl(dy, ld_dy, 1, Wd, ld_wd, 0, A2, lda2, B2, ldb2, K2, act, act ? ld_act : 8, VLB200_BF16, M, ff, K, 1.0f, nullptr,
                        vlb::gemm::ACT_SWIGLU_BWD, nullptr, VLB200_BF16, 0, 0, gate_up, ld_gu, stream);
}

extern "C" int vlb200_gemm_bf16(const void* A, int lda, int a_kmajor, const void* B, int ldb, int b_kmajor, void* D,
                                int ldd, int out_dtype, int M, int N, int K, const void* bias, int act,
                                const void* residual, int residual_dtype, int ldr, int accumulate, void* stream) {
    return vlb200_gemm_bf16_ex(A, lda, a_kmajor, B, ldb, b_kmajor, nullptr, 0, nullptr, 0, 0, D, ldd, out_dtype, M, N, K, 1.0f,
                               bias, act, residual, residual_dtype, ldr, accumulate, stream);
}

extern "C" int vlb200_swiglu_fwd(const void* gate_up, int64_t ld_gu, void* act, int64_t ld_act, int rows, int ff, void* stream);

extern "C" int vlb200_gemm_swiglu_bf16(const void* A, int lda, const void* Wgu, int ldb, void* gu, int ld_gu, int write_gu,
                                       void* act, int ld_act, int M, int ff, int K, void* stream) {
    using namespace vlb;
    using namespace vlb::gemm;
    VLB_REQUIRE(A && Wgu && gu && act, "gemm_swiglu: null pointer");
    VLB_REQUIRE(M > 0 && ff > 0 && K > 0 && ff % 8 == 0, "gemm_swiglu: bad shape M=%d ff=%d K=%d", M, ff, K);
    VLB_REQUIRE(lda % 8 == 0 && ldb % 8 == 0 && ld_gu % 8 == 0 && ld_act % 8 == 0 && lda >= K && ldb >= K && ld_gu >= 2 * ff &&
                    ld_act >= ff, "gemm_swiglu: bad leading dimensions");
    if (g_gemm_mode < 0) {
        const char* e = getenv("VLB200_GEMM_2CTA");
        g_gemm_mode = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    static const bool fuse_on = [] { const char* e = getenv("VLB200_FUSE_SWIGLU"); return !(e && e[0] == '0'); }();
    if (!(fuse_on && g_gemm_mode == 1 && M >= 256 && ff % HALF_N == 0)) {
        // shapes the CTA-pair kernel does not take: plain GEMM into the gate|up buffer, then the elementwise kernel
        int rc = vlb200_gemm_bf16(A, lda, 1, Wgu, ldb, 1, gu, ld_gu, VLB200_BF16, M, 2 * ff, K, nullptr, VLB200_ACT_NONE, nullptr,
                                  VLB200_BF16, 0, 0, stream);
        if (rc) return rc;
        return vlb200_swiglu_fwd(gu, ld_gu, act, ld_act, M, ff, stream);
    }
    VLB_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(Wgu) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(gu) & 15) == 0 && (reinterpret_cast<uintptr_t>(act) & 15) == 0,
                "gemm_swiglu: pointers must be 16-byte aligned");
    CUtensorMap ta, tb;
    int rc = get_tensor_map(A, K, M, lda, BLOCK_K, BLOCK_M, &ta);
    if (rc) return rc;
    rc = get_tensor_map(Wgu, K, 2 * ff, ldb, BLOCK_K, HALF_N, &tb);
    if (rc) return rc;
    Params p;
    p.M = M; p.N = ff; p.K = K;
    p.num_m_blocks = (M + PAIR_M - 1) / PAIR_M;
    p.num_n_blocks = ff / HALF_N;
    p.num_tiles = p.num_m_blocks * p.num_n_blocks;
    p.kb1 = p.num_k_blocks = (K + BLOCK_K - 1) / BLOCK_K;
    p.alpha = 1.0f;
    p.splits = 1; p.kb_per_split = p.num_k_blocks; p.ws = nullptr;
    p.D = act; p.ldd = ld_act; p.out_f32 = 0;
    p.bias = nullptr; p.act = VLB200_ACT_NONE; p.residual = nullptr; p.residual_f32 = 0; p.ldr = 0; p.accumulate = 0;
    p.ff = ff; p.G = write_gu ? reinterpret_cast<__nv_bfloat16*>(gu) : nullptr; p.ldg = ld_gu;
    // a pair tile holds 256 rows of A and 128 gate + 128 up rows of B
    vlb::gemm::choose_raster_pair(p, (double)M, 2.0 * ff, (double)K, (double)PAIR_M * PAIR_N * (write_gu ? 3 : 1));
    return launch_2cta<6, true, true, true>(ta, tb, ta, tb, p, as_stream(stream));
}

extern "C" int vlb200_set_gemm_raster_mb(double mb) {
    vlb::gemm::g_raster_mb = mb;  // <= 0: back to VLB200_RASTER_MB / the default on the next launch
    return VLB200_OK;
}

extern "C" int vlb200_set_gemm_raster_policy(int policy) {
    vlb::gemm::g_raster_policy = policy == 1 ? 1 : (policy < 0 ? -1 : 0);   // < 0: back to VLB200_RASTER_POLICY / the default
    return VLB200_OK;
}

extern "C" int vlb200_gemm_plan_raster(int M, int N, int K, int out_bytes, int policy, int* group, int* along_n,
                                       double* model_read_bytes) {
    using namespace vlb::gemm;
    VLB_REQUIRE(M > 0 && N > 0 && K > 0 && (out_bytes == 2 || out_bytes == 4) && group && along_n, "gemm_plan_raster: bad arguments");
    Params p{};
    p.num_m_blocks = (M + PAIR_M - 1) / PAIR_M;
    p.num_n_blocks = (N + PAIR_N - 1) / PAIR_N;
    const double d_tile = (double)PAIR_M * PAIR_N * out_bytes;
    if (policy == 1) {
        const RasterPlan r = plan_raster_model((double)M, (double)N, p.num_m_blocks, p.num_n_blocks, (double)K, PAIR_M, PAIR_N, d_tile, 74);
        p.group = r.group; p.group_along_n = r.along_n;
    } else {
        choose_raster(p, (double)M, (double)N, (double)K, PAIR_M, PAIR_N);
    }
    *group = p.group; *along_n = p.group_along_n;
    if (model_read_bytes) {
        static const double cap_mb = [] { const char* e = getenv("VLB200_L2_MODEL_MB"); return e && atof(e) > 0.0 ? atof(e) : 60.0; }();
        *model_read_bytes = lru_model_read_bytes(p.num_m_blocks, p.num_n_blocks, (double)K, PAIR_M, PAIR_N, d_tile, p.group,
                                                 p.group_along_n, cap_mb * 1024 * 1024, 74);
    }
    return VLB200_OK;
}

extern "C" int vlb200_gemm_tile_coords(int num_m_blocks, int num_n_blocks, int group, int along_n, int tile, int* m_blk, int* n_blk) {
    VLB_REQUIRE(num_m_blocks > 0 && num_n_blocks > 0 && group > 0 && tile >= 0 && tile < num_m_blocks * num_n_blocks && m_blk && n_blk,
                "gemm_tile_coords: bad arguments");
    vlb::gemm::Params p{};
    p.num_m_blocks = num_m_blocks; p.num_n_blocks = num_n_blocks; p.group = group; p.group_along_n = along_n;
    vlb::gemm::tile_coords(p, tile, *m_blk, *n_blk);   // the function the kernels call
    return VLB200_OK;
}

extern "C" int vlb200_set_gemm_mode(int mode) {
    vlb::gemm::g_gemm_mode = mode ? 1 : 0;
    return VLB200_OK;
}
