// FlashAttention-style fused attention for both towers (K4: CLIP ViT non-causal, dh=64; K12: decoder causal
// + right-padding lengths, dh=128, MHA/GQA), forward and backward.
//
// Round-1 implementation: warp-level mma.sync (m16n8k16 bf16 -> fp32) with cp.async double buffering and
// XOR-swizzled shared memory.  This is the legacy tensor path -- correct and fused (no S/P in HBM), but not
// the tcgen05/TMEM design; DESIGN.md lists the tcgen05 rewrite as the next step for this kernel.
//
// Replaces: CLIPAttention eager/sdpa (modeling_clip.py:261-334), LlamaAttention (modeling_llama.py:199-290)
// and their autograd backward.  Softmax statistics in fp32, scores never leave the SM.
#include <algorithm>
#include <climits>

#include "common.cuh"

namespace vlb {
namespace attn {

constexpr int BM = 64;   // rows of the "outer" tile owned by a CTA (4 warps x 16 rows)
constexpr int BN = 64;   // rows of the streamed tile
constexpr int NTHREADS = 128;
constexpr float LOG2E_F = 1.4426950408889634f;
constexpr float LN2_F = 0.6931471805599453f;

struct Params {
    const __nv_bfloat16* q; const __nv_bfloat16* k; const __nv_bfloat16* v;  // row = b*S + t, head h at col h*DH
    long long ldq, ldk, ldv;
    __nv_bfloat16* o; long long ldo;
    float* lse;               // [B, H, S]
    const int* seqlens;       // [B] or null (= S)
    int B, S, H, KVH;
    float scale;
    // backward only
    const __nv_bfloat16* dout; long long lddo;
    const float* delta;       // [B, H, S]
    __nv_bfloat16* dq; __nv_bfloat16* dk; __nv_bfloat16* dv; long long lddq, lddk, lddv;
};

// ------------------------------------------------------------------ primitives
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;  // src-size 0 -> zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// [rows][DH] bf16 tile, 16-byte chunks XOR-swizzled by (row & 7)
template <int DH>
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) {
    return (uint32_t)((row * (DH / 8) + (chunk ^ (row & 7))) * 16);
}

// async copy of ROWS x DH from global (row stride ld) into a swizzled tile; rows >= valid_rows are zero-filled
template <int ROWS, int DH>
__device__ __forceinline__ void load_tile(uint32_t smem_base, const __nv_bfloat16* g, long long ld, int valid_rows) {
    constexpr int CH = DH / 8;
    for (int i = threadIdx.x; i < ROWS * CH; i += NTHREADS) {
        const int r = i / CH, c = i % CH;
        const bool ok = r < valid_rows;
        cp_async16(smem_base + tile_off<DH>(r, c), g + (ok ? (size_t)r * ld + c * 8 : 0), ok);
    }
}

// A fragment (m16 x k16) of rows [row0, row0+16), k-step ks from a swizzled [rows][DH] tile
template <int DH>
__device__ __forceinline__ void load_a(uint32_t base, int row0, int ks, uint32_t (&a)[4]) {
    const int l = threadIdx.x & 31;
    ldsm_x4(base + tile_off<DH>(row0 + (l & 15), ks * 2 + (l >> 4)), a[0], a[1], a[2], a[3]);
}
// B fragments for two adjacent n-tiles (n0..n0+15) x k16 where B[k][n] = T[n][k] (tile rows are n): non-transposed
template <int DH>
__device__ __forceinline__ void load_b_nk(uint32_t base, int n0, int ks, uint32_t (&b)[4]) {
    const int l = threadIdx.x & 31;
    ldsm_x4(base + tile_off<DH>(n0 + (l & 7) + ((l >> 4) << 3), ks * 2 + ((l >> 3) & 1)), b[0], b[1], b[2], b[3]);
}
// B fragments for k16 rows [k0, k0+16) x two adjacent n-tiles (chunks c, c+1) where B[k][n] = T[k][n]: transposed
template <int DH>
__device__ __forceinline__ void load_b_kn(uint32_t base, int k0, int chunk, uint32_t (&b)[4]) {
    const int l = threadIdx.x & 31;
    ldsm_x4_t(base + tile_off<DH>(k0 + (l & 7) + (((l >> 3) & 1) << 3), chunk + (l >> 4)), b[0], b[1], b[2], b[3]);
}

// ------------------------------------------------------------------ forward
template <int DH, bool CAUSAL>
__global__ void __launch_bounds__(NTHREADS)
attn_fwd_kernel(const Params p) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int TILE = BN * DH * 2;
    const uint32_t sQ = smem_u32(smem);
    const uint32_t sK = sQ + BM * DH * 2;       // 2 stages
    const uint32_t sV = sK + 2 * TILE;          // 2 stages
    const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int kvh = h / (p.H / p.KVH);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = qb * BM;
    const int kv_len = p.seqlens ? p.seqlens[b] : p.S;
    int kmax = kv_len;
    if (CAUSAL) kmax = min(kmax, q0 + BM);
    const int n_tiles = (kmax + BN - 1) / BN;
    const size_t row_base = (size_t)b * p.S;
    const __nv_bfloat16* gq = p.q + (row_base + q0) * p.ldq + (size_t)h * DH;
    const __nv_bfloat16* gk = p.k + row_base * p.ldk + (size_t)kvh * DH;
    const __nv_bfloat16* gv = p.v + row_base * p.ldv + (size_t)kvh * DH;

    load_tile<BM, DH>(sQ, gq, p.ldq, p.S - q0);
    load_tile<BN, DH>(sK, gk, p.ldk, p.S);
    load_tile<BN, DH>(sV, gv, p.ldv, p.S);
    cp_async_commit();

    float o[DH / 8][4];
#pragma unroll
    for (int i = 0; i < DH / 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    const float sl2 = p.scale * LOG2E_F;
    const int qrow0 = q0 + warp * 16 + (lane >> 2);  // this thread's rows: qrow0, qrow0 + 8

    for (int t = 0; t < n_tiles; ++t) {
        const int st = t & 1;
        if (t + 1 < n_tiles) {
            const int k1 = (t + 1) * BN;
            load_tile<BN, DH>(sK + (st ^ 1) * TILE, gk + (size_t)k1 * p.ldk, p.ldk, p.S - k1);
            load_tile<BN, DH>(sV + (st ^ 1) * TILE, gv + (size_t)k1 * p.ldv, p.ldv, p.S - k1);
        }
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();

        // S = Q K^T  (16 x 64 per warp)
        float s[BN / 8][4];
#pragma unroll
        for (int i = 0; i < BN / 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < DH / 16; ++ks) {
            uint32_t a[4];
            load_a<DH>(sQ, warp * 16, ks, a);
#pragma unroll
            for (int nt = 0; nt < BN / 16; ++nt) {
                uint32_t bb[4];
                load_b_nk<DH>(sK + st * TILE, nt * 16, ks, bb);
                mma16816(s[2 * nt], a, bb[0], bb[1]);
                mma16816(s[2 * nt + 1], a, bb[2], bb[3]);
            }
        }
        // mask + online softmax (log2 domain)
        const int k0 = t * BN;
        const bool need_mask = (k0 + BN > kv_len) || (CAUSAL && k0 + BN > q0 + warp * 16);
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nt = 0; nt < BN / 8; ++nt) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float x = s[nt][e] * sl2;
                if (need_mask) {
                    const int key = k0 + nt * 8 + (lane & 3) * 2 + (e & 1);
                    const int qr = qrow0 + (e >> 1) * 8;
                    if (key >= kv_len || (CAUSAL && key > qr)) x = -INFINITY;
                }
                s[nt][e] = x;
                mx[e >> 1] = fmaxf(mx[e >> 1], x);
            }
        }
        float corr[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
            const float m_new = fmaxf(m_run[r], mx[r]);
            // a fully masked row so far keeps m = -inf: use 0 as the reference point to avoid inf - inf
            const float m_use = m_new == -INFINITY ? 0.f : m_new;
            corr[r] = exp2f(m_run[r] - m_use);
            m_run[r] = m_new;
            l_run[r] *= corr[r];
            mx[r] = m_use;
        }
#pragma unroll
        for (int i = 0; i < DH / 8; ++i) { o[i][0] *= corr[0]; o[i][1] *= corr[0]; o[i][2] *= corr[1]; o[i][3] *= corr[1]; }
        uint32_t pa[BN / 16][4];
#pragma unroll
        for (int nt = 0; nt < BN / 8; ++nt) {
            const float p0 = exp2f(s[nt][0] - mx[0]), p1 = exp2f(s[nt][1] - mx[0]);
            const float p2 = exp2f(s[nt][2] - mx[1]), p3 = exp2f(s[nt][3] - mx[1]);
            l_run[0] += p0 + p1;
            l_run[1] += p2 + p3;
            pa[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(p0, p1);
            pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(p2, p3);
        }
        // O += P V
#pragma unroll
        for (int kk = 0; kk < BN / 16; ++kk) {
#pragma unroll
            for (int dc = 0; dc < DH / 16; ++dc) {
                uint32_t bb[4];
                load_b_kn<DH>(sV + st * TILE, kk * 16, dc * 2, bb);
                mma16816(o[2 * dc], pa[kk], bb[0], bb[1]);
                mma16816(o[2 * dc + 1], pa[kk], bb[2], bb[3]);
            }
        }
        __syncthreads();
    }
    cp_async_wait<0>();
    // epilogue: normalise, write O (bf16) and LSE (natural log of sum exp(scale * s))
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int qr = qrow0 + r * 8;
        if (qr < p.S) {
            const float inv = l_run[r] > 0.f ? 1.f / l_run[r] : 0.f;
            __nv_bfloat16* op = p.o + (row_base + qr) * p.ldo + (size_t)h * DH + (lane & 3) * 2;
#pragma unroll
            for (int i = 0; i < DH / 8; ++i)
                *reinterpret_cast<uint32_t*>(op + i * 8) = pack_bf16x2(o[i][2 * r] * inv, o[i][2 * r + 1] * inv);
            if ((lane & 3) == 0 && p.lse)
                p.lse[((size_t)b * p.H + h) * p.S + qr] = l_run[r] > 0.f ? (m_run[r] + log2f(l_run[r])) * LN2_F : -INFINITY;
        }
    }
}

// ------------------------------------------------------------------ backward: delta = rowsum(dO * O)
// row_starts (or null): first row of each sequence when the rows are ragged / packed; delta stays [B, H, S]
__global__ void attn_delta_kernel(const __nv_bfloat16* __restrict__ o, long long ldo, const __nv_bfloat16* __restrict__ dout,
                                  long long lddo, float* __restrict__ delta, const int* __restrict__ row_starts, size_t rows,
                                  int B, int S, int H, int DH) {
    const int warps_per_block = blockDim.x >> 5;
    const size_t total = rows * H;
    const int lane = threadIdx.x & 31;
    for (size_t w = blockIdx.x * (size_t)warps_per_block + (threadIdx.x >> 5); w < total; w += (size_t)gridDim.x * warps_per_block) {
        const int h = (int)(w % H);
        const size_t row = w / H;  // b*S + t, or row_starts[b] + t
        const __nv_bfloat16* op = o + row * ldo + (size_t)h * DH;
        const __nv_bfloat16* dp = dout + row * lddo + (size_t)h * DH;
        float acc = 0.f;
        for (int c = lane * 2; c < DH; c += 64) {
            const float2 a = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(op + c));
            const float2 d = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(dp + c));
            acc += a.x * d.x + a.y * d.y;
        }
        acc = warp_sum(acc);
        if (lane == 0) {
            size_t b = row / S, t = row % S;
            if (row_starts != nullptr) {
                b = 0;
                while (b + 1 < (size_t)B && (size_t)row_starts[b + 1] <= row) ++b;
                t = row - (size_t)row_starts[b];
            }
            if (t < (size_t)S) delta[(b * H + h) * S + t] = acc;
        }
    }
}

// ------------------------------------------------------------------ backward: dK, dV  (CTA = 64 keys of one kv head)
template <int DH, bool CAUSAL>
__global__ void __launch_bounds__(NTHREADS)
attn_bwd_dkdv_kernel(const Params p) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int TILE = 64 * DH * 2;
    const uint32_t sK = smem_u32(smem);
    const uint32_t sV = sK + TILE;
    const uint32_t sQ = sV + TILE;        // 2 stages
    const uint32_t sDO = sQ + 2 * TILE;   // 2 stages
    float* sLse = reinterpret_cast<float*>(smem + 6 * TILE);  // [2][64]
    float* sDelta = sLse + 2 * 64;                            // [2][64]
    const int kb = blockIdx.x, kvh = blockIdx.y, b = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int group = p.H / p.KVH;
    const int k0 = kb * 64;
    const int kv_len = p.seqlens ? p.seqlens[b] : p.S;
    const size_t row_base = (size_t)b * p.S;
    const float sl2 = p.scale * LOG2E_F;

    float dk[DH / 8][4], dv[DH / 8][4];
#pragma unroll
    for (int i = 0; i < DH / 8; ++i) { dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f; dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f; }

    if (k0 < kv_len) {
        load_tile<64, DH>(sK, p.k + (row_base + k0) * p.ldk + (size_t)kvh * DH, p.ldk, p.S - k0);
        load_tile<64, DH>(sV, p.v + (row_base + k0) * p.ldv + (size_t)kvh * DH, p.ldv, p.S - k0);
        // queries beyond the valid length have dO == 0: skip them
        const int qb_begin = CAUSAL ? kb : 0;
        const int qb_end = (kv_len + 63) / 64;
        const int n_q = max(qb_end - qb_begin, 0);
        const int n_iter = n_q * group;
        auto issue = [&](int it, int st) {
            const int hh = kvh * group + it / n_q;
            const int q0 = (qb_begin + it % n_q) * 64;
            load_tile<64, DH>(sQ + st * TILE, p.q + (row_base + q0) * p.ldq + (size_t)hh * DH, p.ldq, p.S - q0);
            load_tile<64, DH>(sDO + st * TILE, p.dout + (row_base + q0) * p.lddo + (size_t)hh * DH, p.lddo, p.S - q0);
            if (threadIdx.x < 64) {
                const int qr = q0 + threadIdx.x;
                const size_t si = ((size_t)b * p.H + hh) * p.S + qr;
                sLse[st * 64 + threadIdx.x] = qr < p.S ? p.lse[si] * LOG2E_F : 0.f;
                sDelta[st * 64 + threadIdx.x] = qr < p.S ? p.delta[si] : 0.f;
            }
        };
        if (n_iter > 0) issue(0, 0);
        cp_async_commit();
        for (int it = 0; it < n_iter; ++it) {
            const int st = it & 1;
            if (it + 1 < n_iter) issue(it + 1, st ^ 1);
            cp_async_commit();
            cp_async_wait<1>();
            __syncthreads();
            const int q0 = (qb_begin + it % n_q) * 64;
            const uint32_t cQ = sQ + st * TILE, cDO = sDO + st * TILE;
            const int key_r0 = k0 + warp * 16 + (lane >> 2);  // this thread's key rows: key_r0, key_r0 + 8
#pragma unroll
            for (int half = 0; half < 2; ++half) {  // 32 queries at a time (register budget)
                const int qh = half * 32;
                float st_[4][4], dp_[4][4];
#pragma unroll
                for (int i = 0; i < 4; ++i) { st_[i][0] = st_[i][1] = st_[i][2] = st_[i][3] = 0.f; dp_[i][0] = dp_[i][1] = dp_[i][2] = dp_[i][3] = 0.f; }
#pragma unroll
                for (int ks = 0; ks < DH / 16; ++ks) {
                    uint32_t ak[4], av[4];
                    load_a<DH>(sK, warp * 16, ks, ak);
                    load_a<DH>(sV, warp * 16, ks, av);
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt) {
                        uint32_t bq[4], bd[4];
                        load_b_nk<DH>(cQ, qh + nt * 16, ks, bq);
                        load_b_nk<DH>(cDO, qh + nt * 16, ks, bd);
                        mma16816(st_[2 * nt], ak, bq[0], bq[1]);
                        mma16816(st_[2 * nt + 1], ak, bq[2], bq[3]);
                        mma16816(dp_[2 * nt], av, bd[0], bd[1]);
                        mma16816(dp_[2 * nt + 1], av, bd[2], bd[3]);
                    }
                }
                uint32_t pa[2][4], dsa[2][4];
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    float pv[4], dsv[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int qi = qh + nt * 8 + (lane & 3) * 2 + (e & 1);  // query index inside the 64-tile
                        const int qg = q0 + qi;
                        const int key = key_r0 + (e >> 1) * 8;
                        const bool masked = key >= kv_len || qg >= p.S || (CAUSAL && key > qg);
                        const float pr = masked ? 0.f : exp2f(st_[nt][e] * sl2 - sLse[st * 64 + qi]);
                        pv[e] = pr;
                        dsv[e] = pr * (dp_[nt][e] - sDelta[st * 64 + qi]);
                    }
                    pa[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(pv[0], pv[1]);
                    pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(pv[2], pv[3]);
                    dsa[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(dsv[0], dsv[1]);
                    dsa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(dsv[2], dsv[3]);
                }
                // dV += P^T dO ; dK += dS^T Q   (k = queries)
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
#pragma unroll
                    for (int dc = 0; dc < DH / 16; ++dc) {
                        uint32_t bd[4], bq[4];
                        load_b_kn<DH>(cDO, qh + kk * 16, dc * 2, bd);
                        load_b_kn<DH>(cQ, qh + kk * 16, dc * 2, bq);
                        mma16816(dv[2 * dc], pa[kk], bd[0], bd[1]);
                        mma16816(dv[2 * dc + 1], pa[kk], bd[2], bd[3]);
                        mma16816(dk[2 * dc], dsa[kk], bq[0], bq[1]);
                        mma16816(dk[2 * dc + 1], dsa[kk], bq[2], bq[3]);
                    }
                }
            }
            __syncthreads();
        }
        cp_async_wait<0>();
    }
    // write dK (scaled) and dV; keys beyond kv_len get exact zeros
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int key = k0 + warp * 16 + (lane >> 2) + r * 8;
        if (key < p.S) {
            __nv_bfloat16* kp = p.dk + (row_base + key) * p.lddk + (size_t)kvh * DH + (lane & 3) * 2;
            __nv_bfloat16* vp = p.dv + (row_base + key) * p.lddv + (size_t)kvh * DH + (lane & 3) * 2;
#pragma unroll
            for (int i = 0; i < DH / 8; ++i) {
                *reinterpret_cast<uint32_t*>(kp + i * 8) = pack_bf16x2(dk[i][2 * r] * p.scale, dk[i][2 * r + 1] * p.scale);
                *reinterpret_cast<uint32_t*>(vp + i * 8) = pack_bf16x2(dv[i][2 * r], dv[i][2 * r + 1]);
            }
        }
    }
}

// ------------------------------------------------------------------ backward: dQ  (CTA = 64 queries of one head)
template <int DH, bool CAUSAL>
__global__ void __launch_bounds__(NTHREADS)
attn_bwd_dq_kernel(const Params p) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int TILE = 64 * DH * 2;
    const uint32_t sQ = smem_u32(smem);
    const uint32_t sDO = sQ + TILE;
    const uint32_t sK = sDO + TILE;      // 2 stages
    const uint32_t sV = sK + 2 * TILE;   // 2 stages
    const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int kvh = h / (p.H / p.KVH);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = qb * 64;
    const int kv_len = p.seqlens ? p.seqlens[b] : p.S;
    int kmax = kv_len;
    if (CAUSAL) kmax = min(kmax, q0 + 64);
    const int n_tiles = q0 < kv_len ? (kmax + 63) / 64 : 0;  // padded queries: dO == 0 -> dQ = 0
    const size_t row_base = (size_t)b * p.S;
    const float sl2 = p.scale * LOG2E_F;
    const __nv_bfloat16* gk = p.k + row_base * p.ldk + (size_t)kvh * DH;
    const __nv_bfloat16* gv = p.v + row_base * p.ldv + (size_t)kvh * DH;

    float dq[DH / 8][4];
#pragma unroll
    for (int i = 0; i < DH / 8; ++i) { dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f; }
    const int qrow0 = q0 + warp * 16 + (lane >> 2);
    float lse2[2], dlt[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int qr = qrow0 + r * 8;
        const size_t si = ((size_t)b * p.H + h) * p.S + qr;
        lse2[r] = qr < p.S ? p.lse[si] * LOG2E_F : 0.f;
        dlt[r] = qr < p.S ? p.delta[si] : 0.f;
    }
    if (n_tiles > 0) {
        load_tile<64, DH>(sQ, p.q + (row_base + q0) * p.ldq + (size_t)h * DH, p.ldq, p.S - q0);
        load_tile<64, DH>(sDO, p.dout + (row_base + q0) * p.lddo + (size_t)h * DH, p.lddo, p.S - q0);
        load_tile<64, DH>(sK, gk, p.ldk, p.S);
        load_tile<64, DH>(sV, gv, p.ldv, p.S);
    }
    cp_async_commit();
    for (int t = 0; t < n_tiles; ++t) {
        const int st = t & 1;
        if (t + 1 < n_tiles) {
            const int k1 = (t + 1) * 64;
            load_tile<64, DH>(sK + (st ^ 1) * TILE, gk + (size_t)k1 * p.ldk, p.ldk, p.S - k1);
            load_tile<64, DH>(sV + (st ^ 1) * TILE, gv + (size_t)k1 * p.ldv, p.ldv, p.S - k1);
        }
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const int k0 = t * 64;
        const uint32_t cK = sK + st * TILE, cV = sV + st * TILE;
#pragma unroll
        for (int half = 0; half < 2; ++half) {  // 32 keys at a time
            const int kh = half * 32;
            float s_[4][4], dp_[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { s_[i][0] = s_[i][1] = s_[i][2] = s_[i][3] = 0.f; dp_[i][0] = dp_[i][1] = dp_[i][2] = dp_[i][3] = 0.f; }
#pragma unroll
            for (int ks = 0; ks < DH / 16; ++ks) {
                uint32_t aq[4], ad[4];
                load_a<DH>(sQ, warp * 16, ks, aq);
                load_a<DH>(sDO, warp * 16, ks, ad);
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    uint32_t bk[4], bv[4];
                    load_b_nk<DH>(cK, kh + nt * 16, ks, bk);
                    load_b_nk<DH>(cV, kh + nt * 16, ks, bv);
                    mma16816(s_[2 * nt], aq, bk[0], bk[1]);
                    mma16816(s_[2 * nt + 1], aq, bk[2], bk[3]);
                    mma16816(dp_[2 * nt], ad, bv[0], bv[1]);
                    mma16816(dp_[2 * nt + 1], ad, bv[2], bv[3]);
                }
            }
            uint32_t dsa[2][4];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                float dsv[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int key = k0 + kh + nt * 8 + (lane & 3) * 2 + (e & 1);
                    const int qr = qrow0 + (e >> 1) * 8;
                    const bool masked = key >= kv_len || (CAUSAL && key > qr);
                    const float pr = masked ? 0.f : exp2f(s_[nt][e] * sl2 - lse2[e >> 1]);
                    dsv[e] = pr * (dp_[nt][e] - dlt[e >> 1]);
                }
                dsa[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(dsv[0], dsv[1]);
                dsa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(dsv[2], dsv[3]);
            }
            // dQ += dS K   (k = keys)
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
#pragma unroll
                for (int dc = 0; dc < DH / 16; ++dc) {
                    uint32_t bk[4];
                    load_b_kn<DH>(cK, kh + kk * 16, dc * 2, bk);
                    mma16816(dq[2 * dc], dsa[kk], bk[0], bk[1]);
                    mma16816(dq[2 * dc + 1], dsa[kk], bk[2], bk[3]);
                }
            }
        }
        __syncthreads();
    }
    cp_async_wait<0>();
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int qr = qrow0 + r * 8;
        if (qr < p.S) {
            __nv_bfloat16* qp = p.dq + (row_base + qr) * p.lddq + (size_t)h * DH + (lane & 3) * 2;
#pragma unroll
            for (int i = 0; i < DH / 8; ++i)
                *reinterpret_cast<uint32_t*>(qp + i * 8) = pack_bf16x2(dq[i][2 * r] * p.scale, dq[i][2 * r + 1] * p.scale);
        }
    }
}

template <typename K>
static int set_smem(K kern, int bytes) {
    VLB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    return VLB200_OK;
}

template <int DH>
static int launch_fwd(const Params& p, int causal, cudaStream_t s) {
    const int smem = (BM * DH + 4 * BN * DH) * 2;
    dim3 grid((p.S + BM - 1) / BM, p.H, p.B);
    if (causal) {
        if (int rc = set_smem(attn_fwd_kernel<DH, true>, smem)) return rc;
        attn_fwd_kernel<DH, true><<<grid, NTHREADS, smem, s>>>(p);
    } else {
        if (int rc = set_smem(attn_fwd_kernel<DH, false>, smem)) return rc;
        attn_fwd_kernel<DH, false><<<grid, NTHREADS, smem, s>>>(p);
    }
    count_launch();
    VLB_LAUNCH_CHECK();
    return VLB200_OK;
}

template <int DH>
static int launch_bwd(const Params& p, int causal, cudaStream_t s) {
    const int smem = 6 * 64 * DH * 2 + 4 * 64 * 4;
    dim3 gkv((p.S + 63) / 64, p.KVH, p.B), gq((p.S + 63) / 64, p.H, p.B);
    if (causal) {
        if (int rc = set_smem(attn_bwd_dkdv_kernel<DH, true>, smem)) return rc;
        if (int rc = set_smem(attn_bwd_dq_kernel<DH, true>, smem)) return rc;
        attn_bwd_dkdv_kernel<DH, true><<<gkv, NTHREADS, smem, s>>>(p);
        attn_bwd_dq_kernel<DH, true><<<gq, NTHREADS, smem, s>>>(p);
    } else {
        if (int rc = set_smem(attn_bwd_dkdv_kernel<DH, false>, smem)) return rc;
        if (int rc = set_smem(attn_bwd_dq_kernel<DH, false>, smem)) return rc;
        attn_bwd_dkdv_kernel<DH, false><<<gkv, NTHREADS, smem, s>>>(p);
        attn_bwd_dq_kernel<DH, false><<<gq, NTHREADS, smem, s>>>(p);
    }
    count_launch(2);
    VLB_LAUNCH_CHECK();
    return VLB200_OK;
}

}  // namespace attn
}  // namespace vlb

using namespace vlb;

static int check_attn_args(int B, int S, int H, int KVH, int DH, int64_t ldq, int64_t ldk, int64_t ldv) {
    VLB_REQUIRE(B > 0 && S > 0 && H > 0 && KVH > 0 && H % KVH == 0, "attention: bad B/S/H/KVH");
    VLB_REQUIRE(DH == 64 || DH == 128, "attention: head_dim %d unsupported (64 or 128)", DH);
    VLB_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0, "attention: row strides must be multiples of 8");
    return VLB200_OK;
}

extern "C" int vlb200_attn_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* out,
                               int64_t ldo, float* lse, const int* seqlens, int B, int S, int H, int KVH, int head_dim,
                               int causal, float scale, void* stream) {
    VLB_REQUIRE(q && k && v && out, "attn_fwd: null pointer");
    if (int rc = check_attn_args(B, S, H, KVH, head_dim, ldq, ldk, ldv)) return rc;
    attn::Params p{};
    p.q = (const __nv_bfloat16*)q; p.k = (const __nv_bfloat16*)k; p.v = (const __nv_bfloat16*)v;
    p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.o = (__nv_bfloat16*)out; p.ldo = ldo; p.lse = lse; p.seqlens = seqlens;
    p.B = B; p.S = S; p.H = H; p.KVH = KVH; p.scale = scale;
    if (head_dim == 64) return attn::launch_fwd<64>(p, causal, as_stream(stream));
    return attn::launch_fwd<128>(p, causal, as_stream(stream));
}

extern "C" int vlb200_attn_delta_varlen(const void* out, int64_t ldo, const void* dout, int64_t lddo, float* delta,
                                        const int* row_starts, int64_t total_rows, int B, int S, int H, int head_dim, void* stream) {
    VLB_REQUIRE(out && dout && delta, "attn_delta: null pointer");
    VLB_REQUIRE(row_starts == nullptr || total_rows > 0, "attn_delta: row_starts needs total_rows");
    const size_t rows = row_starts ? (size_t)total_rows : (size_t)B * S;
    const size_t nw = rows * H;
    const int blocks = (int)std::min<size_t>((nw + 7) / 8, (size_t)num_sms() * 16);
    attn::attn_delta_kernel<<<blocks, 256, 0, as_stream(stream)>>>((const __nv_bfloat16*)out, ldo, (const __nv_bfloat16*)dout, lddo,
                                                                  delta, row_starts, rows, B, S, H, head_dim);
    count_launch();
    VLB_LAUNCH_CHECK();
    return VLB200_OK;
}

extern "C" int vlb200_attn_delta(const void* out, int64_t ldo, const void* dout, int64_t lddo, float* delta, int B, int S, int H,
                                 int head_dim, void* stream) {
    return vlb200_attn_delta_varlen(out, ldo, dout, lddo, delta, nullptr, 0, B, S, H, head_dim, stream);
}

extern "C" int vlb200_attn_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                               const void* out, int64_t ldo, const void* dout, int64_t lddo, const float* lse, float* delta,
                               void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv, const int* seqlens,
                               int B, int S, int H, int KVH, int head_dim, int causal, float scale, void* stream) {
    VLB_REQUIRE(q && k && v && out && dout && lse && delta && dq && dk && dv, "attn_bwd: null pointer");
    if (int rc = check_attn_args(B, S, H, KVH, head_dim, ldq, ldk, ldv)) return rc;
    cudaStream_t s = as_stream(stream);
    const size_t nw = (size_t)B * S * H;
    const int blocks = (int)std::min<size_t>((nw + 7) / 8, (size_t)num_sms() * 16);
    attn::attn_delta_kernel<<<blocks, 256, 0, s>>>((const __nv_bfloat16*)out, ldo, (const __nv_bfloat16*)dout, lddo, delta, nullptr,
                                                   (size_t)B * S, B, S, H, head_dim);
    count_launch();
    VLB_LAUNCH_CHECK();
    attn::Params p{};
    p.q = (const __nv_bfloat16*)q; p.k = (const __nv_bfloat16*)k; p.v = (const __nv_bfloat16*)v;
    p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.lse = const_cast<float*>(lse); p.seqlens = seqlens;
    p.B = B; p.S = S; p.H = H; p.KVH = KVH; p.scale = scale;
    p.dout = (const __nv_bfloat16*)dout; p.lddo = lddo; p.delta = delta;
    p.dq = (__nv_bfloat16*)dq; p.dk = (__nv_bfloat16*)dk; p.dv = (__nv_bfloat16*)dv; p.lddq = lddq; p.lddk = lddk; p.lddv = lddv;
    if (head_dim == 64) return attn::launch_bwd<64>(p, causal, s);
    return attn::launch_bwd<128>(p, causal, s);
}
