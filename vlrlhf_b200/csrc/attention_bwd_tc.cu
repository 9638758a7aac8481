// FlashAttention backward on tcgen05 + TMEM + TMA (sm_100a).  Two launches of one templated kernel:
//
//   MODE 0 (dK, dV): a CTA keeps a 128-key tile (K_j, V_j) in smem and streams 64-query tiles (Q_i, dO_i):
//        S^T = K Q^T, dP^T = V dO^T            tcgen05.mma 128x64x16 (scores transposed: TMEM lane = key)
//        P^T = exp2(S^T*c - lse[q]),  dS^T = scale * P^T o (dP^T - delta[q])      (bf16, written back over the scores in TMEM)
//        dV += P^T dO,  dK += dS^T Q           tcgen05.mma 128xDHx16, B = streamed tile read MN-major
//   MODE 1 (dQ):     a CTA keeps a 128-query tile (Q_i, dO_i) and streams 64-key tiles (K_j, V_j):
//        S = Q K^T, dP = dO V^T;  dS = scale * P o (dP - delta[row]);  dQ += dS K
//
// dQ is produced by a second pass (7 tile-GEMMs instead of 5) so no atomics are needed and results are
// deterministic.  Persistent CTAs, warp 0 = TMA producer (stationary tiles + 4-stage ring of streamed tiles),
// warp 1 = MMA issuer, warps 2-9 = elementwise/epilogue (one stationary row per thread, two warps per TMEM lane quadrant).
// Score tiles are double buffered in TMEM; dK/dV (or dQ) accumulate in TMEM over the whole loop (GQA: over the group's heads).
//
// Replaces the autograd backward of LlamaAttention (modeling_llama.py:199-290).
#include <algorithm>
#include <type_traits>

#include "ptx.cuh"

namespace vlb {
namespace gemm {
int get_tensor_map(const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner, uint32_t box_outer,
                   CUtensorMap* out);
}
namespace attn_bwd_tc {

using namespace ptx;

constexpr int BX = 128;  // stationary rows
constexpr int BY = 64;   // streamed rows
constexpr int NST = 4;   // streamed-tile ring depth
constexpr int NTHREADS = 352;  // TMA warp, score-MMA warp, 8 elementwise warps (two per TMEM lane quadrant), accumulate-MMA warp
constexpr int NEW = 8;         // elementwise warps
constexpr float LOG2E_F = 1.4426950408889634f;

struct Params {
    const float* lse;    // [B, H, S] natural log
    const float* delta;  // [B, H, S]
    const int* seqlens;
    const int* row_starts;  // [B] or null: first row of each sequence (ragged / packed rows); null: b*S
    // shared-prefix attention (packed rows only; see vlb200_attn_fwd_tc_ctx): ctx[b] >= 0 names the sequence whose keys/values
    // every query of sequence b also sees; kids[2*b], kids[2*b+1] (or -1) are the sequences that name b as their context
    const int* ctx;
    const int* kids;
    __nv_bfloat16* out1; long long ld1;  // MODE 0: dV ; MODE 1: unused
    __nv_bfloat16* out2; long long ld2;  // MODE 0: dK ; MODE 1: dQ
    int B, S, H, KVH, causal;
    float scale;
    int n_xb, n_work;
};

// diagnostics (DBG bit 8): cycles one elementwise thread per CTA spends in each phase of its loop, summed over the grid
__device__ unsigned long long g_bwd_prof[32];   // [MODE][16]
// DBG bit 16: clock64 timestamps of CTA 0's first 64 tile iterations (one SM: the clocks of its warps are comparable):
// [MODE][event][tile]; events: 0 elementwise starts waiting for the scores, 1 scores seen, 2 E published (arrive issued),
// 3 MMA warp sees E, 4 accumulate MMAs issued, 5 MMA warp sees the streamed tile of the next score pair, 6 score MMAs issued
__device__ long long g_bwd_trace[2 * 32 * 64];   // events 8+w / 16+w: elementwise warp w publishes E / sees the scores
#define VLB_TRACE(ev, idx) do { if ((DBG & 16) && blockIdx.x == 0 && (idx) < 64 && lane_idx == 0) g_bwd_trace[(MODE * 32 + (ev)) * 64 + (idx)] = clock64(); } while (0)
#define VLB_PROF(i) do { if (DBG & 8) { const long long t_ = clock64(); prof[i] += (unsigned long long)(t_ - tp); tp = t_; } } while (0)

__device__ __forceinline__ void tmem_ld_32x32b(uint32_t taddr, uint32_t (&r)[32]) { tmem_ld_32x32(taddr, r); }
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int MODE>
__device__ __forceinline__ void item_coords(const Params& p, int w, int& b, int& hx, int& xb) {
    const int heads = MODE == 0 ? p.KVH : p.H;
    const int bh = w / p.n_xb;
    xb = MODE == 0 ? (w - bh * p.n_xb) : (p.n_xb - 1 - (w - bh * p.n_xb));  // heavy tiles first
    hx = bh % heads;
    b = bh / heads;
}
// Streamed tiles of one item: up to three SEGMENTS of 64-row tiles, the whole list repeated for `reps` heads (GQA group).
//   MODE 0 (stationary keys of sequence b):    seg 0 = b's own queries from the causal diagonal on,
//                                              seg 1/2 = ALL queries of the sequences that use b as their context (no causal mask)
//   MODE 1 (stationary queries of sequence b): seg 0 = ALL keys of b's context sequence (no causal mask),
//                                              seg 1 = b's own keys up to the causal diagonal
// A segment: tiles [yb, yb + n) of sequence `seq` (rows row0 + yt*64, `len` valid rows, statistics of `seq`).
struct Seg { int n, yb, row0, len, seq, causal; };
struct Item { Seg s[3]; int per_rep, reps, kv_len; };

template <int MODE>
__device__ __forceinline__ Item item_plan(const Params& p, int b, int xb) {
    Item it;
    int kv_len = p.seqlens ? p.seqlens[b] : p.S;
    kv_len = max(min(kv_len, p.S), 0);
    it.kv_len = kv_len;
    const int x0 = xb * BX;
    const int row0 = p.row_starts ? p.row_starts[b] : b * p.S;
#pragma unroll
    for (int i = 0; i < 3; ++i) it.s[i] = Seg{0, 0, 0, 0, b, 0};
    const bool live = x0 < kv_len;
    if (MODE == 0) {  // stationary keys [x0, x0+128); queries beyond kv_len have dO == 0
        it.reps = p.H / p.KVH;
        const int yb = p.causal ? x0 / BY : 0;
        const int ye = live ? (kv_len + BY - 1) / BY : yb;
        it.s[0] = Seg{max(ye - yb, 0), yb, row0, kv_len, b, p.causal};
        if (p.kids != nullptr && live) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int c = p.kids[2 * b + k];
                if (c >= 0) {
                    const int len = max(min(p.seqlens[c], p.S), 0);
                    it.s[1 + k] = Seg{(len + BY - 1) / BY, 0, p.row_starts[c], len, c, 0};
                }
            }
        }
    } else {          // stationary queries; keys up to the causal diagonal / kv_len
        it.reps = 1;
        int kmax = kv_len;
        if (p.causal) kmax = min(kmax, x0 + BX);
        const Seg self = Seg{live ? (kmax + BY - 1) / BY : 0, 0, row0, kv_len, b, p.causal};
        if (p.ctx != nullptr && live && p.ctx[b] >= 0) {
            const int c = p.ctx[b];
            const int len = max(min(p.seqlens[c], p.S), 0);
            it.s[0] = Seg{(len + BY - 1) / BY, 0, p.row_starts[c], len, c, 0};
            it.s[1] = self;
        } else {
            it.s[0] = self;
        }
    }
    it.per_rep = it.s[0].n + it.s[1].n + it.s[2].n;
    return it;
}
// tile u of one head repetition of an item -> (segment, tile index inside the sequence)
__device__ __forceinline__ void item_tile_u(const Item& it, int u, Seg& sg, int& yt) {
    if (u < it.s[0].n) { sg = it.s[0]; }
    else if (u < it.s[0].n + it.s[1].n) { u -= it.s[0].n; sg = it.s[1]; }
    else { u -= it.s[0].n + it.s[1].n; sg = it.s[2]; }
    yt = sg.yb + u;
}
// tile t of an item -> (segment, tile index inside the sequence, head repetition)
__device__ __forceinline__ void item_tile(const Item& it, int t, Seg& sg, int& yt, int& rep) {
    rep = t / it.per_rep;
    int u = t - rep * it.per_rep;
    if (u < it.s[0].n) { sg = it.s[0]; }
    else if (u < it.s[0].n + it.s[1].n) { u -= it.s[0].n; sg = it.s[1]; }
    else { u -= it.s[0].n + it.s[1].n; sg = it.s[2]; }
    yt = sg.yb + u;
}

// The elementwise results (P^T / dS^T, or dS) stay in TENSOR MEMORY: each elementwise warp writes its 32 bf16 columns, packed
// two per 32-bit column, over the first 16 columns of the 32 fp32 score columns it has just read (tcgen05.st), and the
// accumulate MMAs take their A operand from TMEM (TS mode, one K = 16 step = 8 columns): no E stores to / A reads from shared
// memory, no generic->async proxy fence.  A score buffer is recycled in MMA issue order (the score MMAs of tile t+2 are issued
// after the accumulate MMAs of tile t), not by a barrier.  The MMA warp runs CONVERGED (all lanes poll the barriers, votes
// make the branches warp-uniform) and one elected lane issues the MMAs: with the role under `if (lane == 0)` ptxas wraps every
// tcgen05.mma in an ELECT / R2UR / BRA.U.ANY waterfall, ~10 instructions per MMA -- at 20-24 MMAs of 32 tensor-cycles per
// 64-row tile the issue thread was the bound of the dQ pass (profiles/r2j_attn_phases.log: 1260 cycles per tile waiting for scores).
// DBG (diagnostics, wrong results, timing only), a bit mask: 1 = no MMAs are issued (barriers only), 2 = no exponentials,
// 4 = no streamed-tile loads, 8 = per-phase cycle counters
template <int DH, int MODE, int DBG = 0>
__global__ void __launch_bounds__(NTHREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tma_x1, const __grid_constant__ CUtensorMap tma_x2,
                   const __grid_constant__ CUtensorMap tma_y1, const __grid_constant__ CUtensorMap tma_y2, const Params p) {
    constexpr int NCH = DH / 64;
    constexpr int XCH = BX * 128;            // bytes of one [128 x 128 B] chunk
    constexpr int YCH = BY * 128;            // bytes of one [64 x 128 B] chunk
    constexpr int X_BYTES = NCH * XCH;       // one stationary tile
    constexpr int Y_BYTES = NCH * YCH;       // one streamed tile
    constexpr uint32_t TMEM_COLS = 512;
    constexpr uint32_t TM_T1 = 0, TM_T2 = 128, TM_A1 = 256, TM_A2 = 384;  // T buffers: +64 per stage
    // MODE 1 has one accumulator (dQ): the stationary Q and dO tiles live in the other accumulator's columns, two bf16 per
    // column, and the score MMAs run in TS mode.  An M = 128, N = 64 SS-mode MMA reads 6 KB of operands per 32 tensor cycles
    // -- 192 B/clk against the 128 B/clk shared-memory port (the dQ pass waited 900 cycles per tile for its scores); with A in
    // tensor memory only the 2 KB of the streamed tile are read.
    constexpr uint32_t TM_X1 = 256, TM_X2 = 320;

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sX1 = smem;
    uint8_t* sX2 = sX1 + X_BYTES;
    uint8_t* sY = sX2 + X_BYTES;                 // NST stages of (Y1, Y2)
    float* sStat = reinterpret_cast<float*>(sY + NST * 2 * Y_BYTES);  // MODE 0: [NST stages][lse2 x64 | delta x64] of the streamed queries
    uint8_t* sOut = reinterpret_cast<uint8_t*>(sStat + NST * 128);   // [NEW warps][32 rows][64 B]: epilogue staging (coalesced stores)
    uint64_t* bars = reinterpret_cast<uint64_t*>(sOut + NEW * 2048);
    uint64_t* x_full = bars + 0;
    uint64_t* x_empty = bars + 1;
    uint64_t* y_full = bars + 2;             // [NST]
    uint64_t* y_empty = bars + 2 + NST;      // [NST]
    uint64_t* t_full = bars + 2 + 2 * NST;   // [2]
    uint64_t* e_full = t_full + 2;           // [2]
    uint64_t* e_done = e_full + 2;           // [2]
    uint64_t* acc_free = e_done + 2;
    uint64_t* xt_full = acc_free + 1;        // MODE 1: stationary tiles copied to tensor memory by the 8 elementwise warps
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(xt_full + 1);

    const int warp_idx = threadIdx.x >> 5, lane_idx = threadIdx.x & 31;
    if (warp_idx == 0 && lane_idx == 0) {
        prefetch_tensormap(&tma_x1); prefetch_tensormap(&tma_x2); prefetch_tensormap(&tma_y1); prefetch_tensormap(&tma_y2);
        mbar_init(x_full, 1); mbar_init(x_empty, MODE == 1 ? NEW : 1); mbar_init(xt_full, NEW);
        for (int i = 0; i < NST; ++i) { mbar_init(&y_full[i], MODE == 0 ? 2 : 1); mbar_init(&y_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&e_full[i], NEW); mbar_init(&e_done[i], 1); }
        mbar_init(acc_free, NEW);
        fence_barrier_init();
    }
    if (warp_idx == 1) tmem_alloc(tmem_base_smem, TMEM_COLS);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;
    const int group = p.H / p.KVH;

    if (warp_idx == 0) {
        // ===================== TMA producer (lane 0) + per-column statistics of the streamed tiles (whole warp, MODE 0) =====
        uint32_t item = 0, yc = 0;
        for (int w = blockIdx.x; w < p.n_work; w += gridDim.x) {
            int b, hx, xb;
            item_coords<MODE>(p, w, b, hx, xb);
            const Item it = item_plan<MODE>(p, b, xb);
            const int n = __shfl_sync(0xffffffffu, it.per_rep * it.reps, 0);
            if (n == 0) continue;
            const int row0 = p.row_starts ? p.row_starts[b] : b * p.S;
            const int xcol = hx * DH;  // MODE 0: kv head; MODE 1: q head
            if (lane_idx == 0) {
                mbar_wait_relaxed(x_empty, (item & 1) ^ 1, 10);
                mbar_arrive_expect_tx(x_full, 2 * X_BYTES);
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    tma_load_2d(&tma_x1, x_full, sX1 + c * XCH, xcol + c * 64, row0 + xb * BX);
                    tma_load_2d(&tma_x2, x_full, sX2 + c * XCH, xcol + c * 64, row0 + xb * BX);
                }
            }
            // MODE 0: lane l holds log-sum-exp (log2 domain) and delta of queries l and 32 + l of the NEXT tile to publish (the
            // global loads fly while the warp waits for a free stage)
            float sl_[2] = {0.f, 0.f}, sd_[2] = {0.f, 0.f};
            auto load_stats = [&](int t) {
                if (MODE == 0 && t < n) {
                    Seg sg; int yt, rep;
                    item_tile(it, t, sg, yt, rep);
                    const long long sbase = ((long long)sg.seq * p.H + (hx * group + rep)) * p.S;
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        const int q = yt * BY + hh * 32 + lane_idx;
                        const bool ok = q < sg.len;
                        sl_[hh] = ok ? p.lse[sbase + q] : 0.f;
                        sd_[hh] = ok ? p.delta[sbase + q] : 0.f;
                    }
                }
            };
            load_stats(0);
            for (int t = 0; t < n; ++t, ++yc) {
                const int st = yc % NST;
                Seg sg; int yt, rep;
                item_tile(it, t, sg, yt, rep);
                const int ycol = (MODE == 0 ? (hx * group + rep) : (hx / group)) * DH;
                mbar_wait_relaxed(&y_empty[st], ((yc / NST) & 1) ^ 1, 20 + st);
                __syncwarp();
                if (lane_idx == 0) {
                    if (DBG & 4) {
                        mbar_arrive(&y_full[st]);
                    } else {
                        mbar_arrive_expect_tx(&y_full[st], 2 * Y_BYTES);
                        uint8_t* y1 = sY + st * 2 * Y_BYTES;
                        uint8_t* y2 = y1 + Y_BYTES;
#pragma unroll
                        for (int c = 0; c < NCH; ++c) {
                            tma_load_2d(&tma_y1, &y_full[st], y1 + c * YCH, ycol + c * 64, sg.row0 + yt * BY);
                            tma_load_2d(&tma_y2, &y_full[st], y2 + c * YCH, ycol + c * 64, sg.row0 + yt * BY);
                        }
                    }
                }
                if (MODE == 0) {
                    // statistics of the tile's 64 queries -> the stage's slot; the second arrival on y_full publishes them (the
                    // elementwise warps read them as broadcast float4s: one warp loads what eight warps used to load redundantly,
                    // 840 cycles per tile in profiles/r2k_attn.log)
                    const uint32_t slot_a = smem_u32(sStat + st * 128);
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        sts_f32(slot_a + ((hh * 32 + lane_idx) << 2), sl_[hh] * LOG2E_F);
                        sts_f32(slot_a + ((64 + hh * 32 + lane_idx) << 2), sd_[hh]);
                    }
                    __syncwarp();
                    if (lane_idx == 0) mbar_arrive(&y_full[st]);
                    load_stats(t + 1);
                }
            }
            ++item;
        }
    } else if (warp_idx == 1) {
        // ===================== score MMAs (whole warp converged; one elected lane issues) =====================
        // The MMAs are issued by TWO warps in a fixed order -- this one: S / dP of tile t once its streamed tile has landed and
        // the accumulate MMAs of tile t-2 have COMPLETED (they read the E operands that live in t's score buffer); warp 10:
        // the accumulate MMAs of tile t once its E operands are published.  One warp doing both spent ~1 700 of its ~2 400
        // cycles per tile in scalar control code between the two issue blocks (barrier polls, fences, descriptor set-up,
        // commits: profiles/r2o2_bwd_trace.log), which made it the bound of both passes.  The tile counter runs across items:
        // the first score tiles of the next item are issued while the elementwise warps write the previous item's epilogue.
        constexpr uint32_t idesc_t = make_idesc_bf16_f32(BX, BY, false, false);
        const uint64_t dx1 = make_smem_desc_sw128(smem_u32(sX1), 1024, 0), dx2 = make_smem_desc_sw128(smem_u32(sX2), 1024, 0);
        uint32_t item = 0, tc = 0;
        for (int w = blockIdx.x; w < p.n_work; w += gridDim.x) {
            int b, hx, xb;
            item_coords<MODE>(p, w, b, hx, xb);
            const Item it = item_plan<MODE>(p, b, xb);
            const int n = __shfl_sync(0xffffffffu, it.per_rep * it.reps, 0);
            if (n == 0) continue;
            if (MODE == 1) mbar_wait_relaxed(xt_full, item & 1, 31);
            else mbar_wait_relaxed(x_full, item & 1, 30);
            __syncwarp();
            for (int ts = 0; ts < n; ++ts, ++tc) {
                const uint32_t st = tc % NST, tb = tc & 1;
                mbar_wait_relaxed(&y_full[st], (tc / NST) & 1, 40);
                if (tc >= 2) mbar_wait_relaxed(&e_done[tb], ((tc - 2) >> 1) & 1, 44);
                __syncwarp();
                VLB_TRACE(5, tc);
                tcgen05_fence_after();
                const uint32_t y1 = smem_u32(sY + st * 2 * Y_BYTES), y2 = y1 + Y_BYTES;
                uint64_t dy1 = make_smem_desc_sw128(y1, 1024, 0), dy2 = make_smem_desc_sw128(y2, 1024, 0);
                uint64_t ax1 = dx1, ax2 = dx2;
                // (opaque copies: the per-k descriptors are derived HERE by immediate adds on the uniform datapath; hoisted out
                // of the loop they lived in vector registers and cost two R2UR each)
                asm volatile("" : "+l"(dy1), "+l"(dy2), "+l"(ax1), "+l"(ax2));
                if (elect_one_sync()) {
                    VLB_TRACE(26, tc);
                    if (!(DBG & 1)) {
#pragma unroll
                        for (int k = 0; k < DH / 16; ++k) {
                            const uint32_t xo = ((k >> 2) * XCH + (k & 3) * 32) >> 4, yo = ((k >> 2) * YCH + (k & 3) * 32) >> 4;
                            if (MODE == 1) umma_f16_ts(tmem_base + TM_T1 + tb * BY, tmem_base + TM_X1 + k * 8, dy1 + yo, idesc_t, k != 0);
                            else umma_f16_ss(tmem_base + TM_T1 + tb * BY, ax1 + xo, dy1 + yo, idesc_t, k != 0);
                        }
#pragma unroll
                        for (int k = 0; k < DH / 16; ++k) {
                            const uint32_t xo = ((k >> 2) * XCH + (k & 3) * 32) >> 4, yo = ((k >> 2) * YCH + (k & 3) * 32) >> 4;
                            if (MODE == 1) umma_f16_ts(tmem_base + TM_T2 + tb * BY, tmem_base + TM_X2 + k * 8, dy2 + yo, idesc_t, k != 0);
                            else umma_f16_ss(tmem_base + TM_T2 + tb * BY, ax2 + xo, dy2 + yo, idesc_t, k != 0);
                        }
                    }
                    VLB_TRACE(27, tc);
                    umma_commit(&t_full[tb]);
                    if (MODE == 0 && ts == n - 1) umma_commit(x_empty);
                }
                __syncwarp();
                VLB_TRACE(6, tc);
            }
            ++item;
        }
    } else if (warp_idx == 10) {
        // ===================== accumulate MMAs (whole warp converged; one elected lane issues) =====================
        constexpr uint32_t idesc_a = make_idesc_bf16_f32(BX, DH, false, true);  // B = streamed tile, MN-major
        uint32_t item = 0, ec = 0;
        for (int w = blockIdx.x; w < p.n_work; w += gridDim.x) {
            int b, hx, xb;
            item_coords<MODE>(p, w, b, hx, xb);
            const Item it = item_plan<MODE>(p, b, xb);
            const int n = __shfl_sync(0xffffffffu, it.per_rep * it.reps, 0);
            if (n == 0) continue;
            for (int ta = 0; ta < n; ++ta, ++ec) {
                const uint32_t st = ec % NST, eb = ec & 1;
                mbar_wait_relaxed(&e_full[eb], (ec >> 1) & 1, 41);
                if (ta == 0) mbar_wait_relaxed(acc_free, (item & 1) ^ 1, 60);   // the previous item's epilogue has read the accumulators
                __syncwarp();
                VLB_TRACE(3, ec);
                tcgen05_fence_after();
                const uint32_t y1 = smem_u32(sY + st * 2 * Y_BYTES), y2 = y1 + Y_BYTES;
                uint64_t by1 = make_smem_desc_sw128(y1, 1024, YCH), by2 = make_smem_desc_sw128(y2, 1024, YCH);
                asm volatile("" : "+l"(by1), "+l"(by2));
                if (elect_one_sync()) {
                    VLB_TRACE(24, ec);
                    if (!(DBG & 1)) {
#pragma unroll
                        for (int k = 0; k < BY / 16; ++k) {
                            const uint32_t acc = (ta != 0 || k != 0) ? 1u : 0u;
                            // streamed rows 16k..16k+15: written by elementwise half k/2 at columns half*32 + (k%2)*8 of the score buffer
                            const uint32_t ac = eb * BY + (k >> 1) * 32 + (k & 1) * 8;
                            // MODE 0: dV += P^T dO, dK += dS^T Q ; MODE 1: dQ += dS K
                            if (MODE == 0) umma_f16_ts(tmem_base + TM_A1, tmem_base + TM_T1 + ac, by2 + (uint64_t)(k * 128), idesc_a, acc);
                            umma_f16_ts(tmem_base + TM_A2, tmem_base + TM_T2 + ac, by1 + (uint64_t)(k * 128), idesc_a, acc);
                        }
                    }
                    VLB_TRACE(25, ec);
                    umma_commit(&e_done[eb]);
                    umma_commit(&y_empty[st]);
                }
                __syncwarp();
                VLB_TRACE(4, ec);
            }
            ++item;
        }
    } else {
        // ===================== elementwise + epilogue (8 warps: row = TMEM lane, two warps split the columns) ==========
        const int quad = warp_idx & 3;
        const int half = (warp_idx - 2) >> 2;  // 0: columns [0,32) of a score tile, 1: columns [32,64)
        const int r = quad * 32 + lane_idx;
        const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
        const float sl2 = p.scale * LOG2E_F;
        uint32_t tc = 0, xitem = 0;
        unsigned long long prof[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        long long tp = clock64();
        for (int w = blockIdx.x; w < p.n_work; w += gridDim.x) {
            int b, hx, xb;
            item_coords<MODE>(p, w, b, hx, xb);
            const Item it = item_plan<MODE>(p, b, xb);
            const int n = it.per_rep * it.reps;
            if (DBG & 8) { prof[8] += 1; prof[9] += n; }
            const int kv_len = it.kv_len;
            const int x0 = xb * BX;
            const int xrow = x0 + r;  // MODE 0: key index ; MODE 1: query index
            float row_lse2 = 0.f, row_delta = 0.f;
            // rows at or beyond kv_len are masked below (their tiles always take the masked path): their statistics are never
            // read -- with packed rows they were never written
            if (MODE == 1 && xrow < kv_len) {
                const long long si = ((long long)b * p.H + hx) * p.S + xrow;
                row_lse2 = p.lse[si] * LOG2E_F;
                row_delta = p.delta[si];
            }
            if (MODE == 1 && n > 0) {
                // stationary tiles: smem (TMA, 128B-swizzled) -> this thread's TMEM lane, two bf16 per column (the K-major A
                // operand of the score MMAs).  half 0 copies Q, half 1 copies dO.  Every MMA of the previous item has retired (its
                // epilogue waited for the last accumulate).
                mbar_wait(x_full, xitem & 1, 70);
                const uint32_t sx_a = smem_u32(half == 0 ? sX1 : sX2);
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        uint32_t qw[16];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const uint4 v = lds128(sx_a + c * XCH + r * 128 + (((hh * 4 + u) ^ (r & 7)) << 4));
                            qw[u * 4 + 0] = v.x; qw[u * 4 + 1] = v.y; qw[u * 4 + 2] = v.z; qw[u * 4 + 3] = v.w;
                        }
                        tmem_st_32x32_x16(tmem_base + lane_addr + (half == 0 ? TM_X1 : TM_X2) + c * 32 + hh * 16, qw);
                    }
                }
                tmem_st_wait_all();
                tcgen05_fence_before();
                __syncwarp();
                if (lane_idx == 0) { mbar_arrive(xt_full); mbar_arrive(x_empty); }
                ++xitem;
            }
            VLB_PROF(0);   // item start
            int tu = 0;   // tile index inside the current head repetition (no division per tile)
            for (int t = 0; t < n; ++t, ++tc) {
                const uint32_t tb = tc & 1;
                Seg sg; int yt;
                item_tile_u(it, tu, sg, yt);
                if (++tu == it.per_rep) tu = 0;
                const int y0 = yt * BY;
                const int ylen = sg.len;             // valid streamed rows of this tile's sequence
                const bool causal_t = sg.causal != 0;
                VLB_PROF(1);   // tile coordinates
                if (warp_idx == 2) VLB_TRACE(0, tc);
                mbar_wait(&t_full[tb], (tc >> 1) & 1, 80 + tb);
                tcgen05_fence_after();
                if (MODE == 0) mbar_wait(&y_full[tc % NST], (tc / NST) & 1, 75);   // acquires the producer's statistics (complete: the scores are)
                if (warp_idx == 2) VLB_TRACE(1, tc);
                VLB_TRACE(16 + warp_idx - 2, tc);
                VLB_PROF(2);   // wait for the score tiles
                if (DBG & 32) {   // handshake only: the tensor side alone
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane_idx == 0) mbar_arrive(&e_full[tb]);
                    continue;
                }
                uint32_t t1[32], t2[32];
                tmem_ld_32x32b(tmem_base + lane_addr + TM_T1 + tb * BY + half * 32, t1);
                tmem_ld_32x32b(tmem_base + lane_addr + TM_T2 + tb * BY + half * 32, t2);
                tmem_ld_wait();
                VLB_PROF(3);   // TMEM -> registers
                // mask only tiles that touch the causal diagonal or the end of the valid range
                const int ymax = y0 + BY - 1;
                bool need_mask;
                if (MODE == 0) need_mask = ymax >= ylen || x0 + BX > kv_len || (causal_t && x0 + BX - 1 > y0);
                else need_mask = ymax >= ylen || x0 + BX > kv_len || (causal_t && ymax > x0);
                // MODE 0: statistics of this warp's 32 columns (explicit ld.shared: through a pointer the compiler emitted
                // generic LD.E.128, whose latency sat at the head of every tile)
                const uint32_t cl = smem_u32(sStat) + (((tc % NST) * 128 + half * 32) << 2);
                const uint32_t cd = cl + 256;
                uint32_t e1[16], e2[16];
                // two straight-line copies of the tile body (masked / unmasked): a per-element `if (need_mask)` costs a
                // divergence region per element (seen in the r1 SASS: 33 BSSY/BSYNC pairs, 134 ISETP per 32 elements)
                auto tile_body = [&](auto masked_tag) {
                    constexpr bool MASKED = decltype(masked_tag)::value;
                    const uint64_t sl2p = pack_f32x2(sl2, sl2);
#pragma unroll
                    for (int c4 = 0; c4 < 32; c4 += 4) {
                        float l4[4], d4[4];
                        if (MODE == 0) {
                            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(l4[0]), "=f"(l4[1]), "=f"(l4[2]), "=f"(l4[3]) : "r"(cl + c4 * 4));
                            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(d4[0]), "=f"(d4[1]), "=f"(d4[2]), "=f"(d4[3]) : "r"(cd + c4 * 4));
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e) { l4[e] = row_lse2; d4[e] = row_delta; }
                        }
                        float pv[4], dv[4];
                        if (!MASKED) {
                            // two elements per instruction (FFMA2 / FADD2 / FMUL2): the phase costs (non-MUFU instructions) + (MUFU
                            // instructions), serialised (profiles/r2ah_experiment.log)
#pragma unroll
                            for (int e = 0; e < 4; e += 2) {
                                const uint64_t xp = fma_f32x2(pack_f32x2(__uint_as_float(t1[c4 + e]), __uint_as_float(t1[c4 + e + 1])), sl2p,
                                                              pack_f32x2(-l4[e], -l4[e + 1]));
                                float x0, x1;
                                unpack_f32x2(xp, x0, x1);
                                if (!(DBG & 2)) { x0 = ex2_approx(x0); x1 = ex2_approx(x1); }
                                pv[e] = x0; pv[e + 1] = x1;
                                const uint64_t dd = add_f32x2(pack_f32x2(__uint_as_float(t2[c4 + e]), __uint_as_float(t2[c4 + e + 1])),
                                                              pack_f32x2(-d4[e], -d4[e + 1]));
                                unpack_f32x2(mul_f32x2(pack_f32x2(x0, x1), dd), dv[e], dv[e + 1]);   // (x scale in the epilogue)
                            }
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int col = half * 32 + c4 + e;
                                float pr = fmaf(__uint_as_float(t1[c4 + e]), sl2, -l4[e]);
                                if (!(DBG & 2)) pr = ex2_approx(pr);
                                // stationary index xrow < kv_len, streamed index y0 + col < ylen; causal inside a sequence only
                                const int key = MODE == 0 ? xrow : y0 + col;
                                const int qi = MODE == 0 ? y0 + col : xrow;
                                if (!(xrow < kv_len && y0 + col < ylen && (!causal_t || key <= qi))) pr = 0.f;
                                pv[e] = pr;
                                dv[e] = pr * (__uint_as_float(t2[c4 + e]) - d4[e]);   // (x scale in the epilogue: dK / dQ are linear in dS)
                            }
                        }
                        e1[c4 >> 1] = pack_bf16x2(pv[0], pv[1]); e1[(c4 >> 1) + 1] = pack_bf16x2(pv[2], pv[3]);
                        e2[c4 >> 1] = pack_bf16x2(dv[0], dv[1]); e2[(c4 >> 1) + 1] = pack_bf16x2(dv[2], dv[3]);
                    }
                };
                if (need_mask) tile_body(std::true_type{});
                else tile_body(std::false_type{});
                VLB_PROF(4);   // elementwise
                // E operands over the score columns this warp has read (its own lanes, its own 32 columns)
                if (MODE == 0) tmem_st_32x32_x16(tmem_base + lane_addr + TM_T1 + tb * BY + half * 32, e1);
                tmem_st_32x32_x16(tmem_base + lane_addr + TM_T2 + tb * BY + half * 32, e2);
                tmem_st_wait_all();
                tcgen05_fence_before();
                __syncwarp();
                if (lane_idx == 0) mbar_arrive(&e_full[tb]);
                if (warp_idx == 2) VLB_TRACE(2, tc);
                VLB_TRACE(8 + warp_idx - 2, tc);
                VLB_PROF(5);   // E store, completion, fences, publish
            }
            // ---- epilogue: accumulators -> bf16 -> global (zeros when the item had no work); columns split by `half`
            if (n > 0) {  // the commit of the last tile's accumulate MMAs covers every earlier tcgen05 op of the MMA thread
                mbar_wait(&e_done[(tc - 1) & 1], ((tc - 1) >> 1) & 1, 95);
                tcgen05_fence_after();
            }
            // Each warp owns a [32 rows x 64 columns] block of every accumulator.  A thread holds one ROW of a 32-column chunk
            // (TMEM lane = row): storing it directly scatters every warp-wide 16-byte store over 32 cache lines (the epilogue cost
            // 8 230 cycles per item in the dK/dV pass, profiles/r2k_attn.log).  The chunk goes through a swizzled 2 KB staging
            // block of the warp instead and leaves as 64-byte row segments, 8 rows per store instruction.
            const int row_lim = p.row_starts ? kv_len : p.S;   // packed rows: the tile may run into the next sequence
            const long long grow0 = (p.row_starts ? (long long)p.row_starts[b] : (long long)b * p.S) + x0 + quad * 32;
            const uint32_t stg_a = smem_u32(sOut + (warp_idx - 2) * 2048);
#pragma unroll
            for (int a = (MODE == 0 ? 0 : 1); a < 2; ++a) {
                __nv_bfloat16* dst0 = (a == 0 ? p.out1 : p.out2) + (long long)hx * DH;
                const long long ldd = a == 0 ? p.ld1 : p.ld2;
                const float osc = a == 0 ? 1.f : p.scale;   // dK / dQ: the softmax scale folded out of dS
#pragma unroll
                for (int cc = 0; cc < DH / 64; ++cc) {
                    const int c = half * (DH / 64) + cc;
                    uint32_t acc[32];
                    if (n > 0) {
                        tmem_ld_32x32b(tmem_base + lane_addr + (a == 0 ? TM_A1 : TM_A2) + c * 32, acc);
                        tmem_ld_wait();
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) acc[i] = 0u;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        uint4 v;
                        v.x = pack_bf16x2(__uint_as_float(acc[u * 8 + 0]) * osc, __uint_as_float(acc[u * 8 + 1]) * osc);
                        v.y = pack_bf16x2(__uint_as_float(acc[u * 8 + 2]) * osc, __uint_as_float(acc[u * 8 + 3]) * osc);
                        v.z = pack_bf16x2(__uint_as_float(acc[u * 8 + 4]) * osc, __uint_as_float(acc[u * 8 + 5]) * osc);
                        v.w = pack_bf16x2(__uint_as_float(acc[u * 8 + 6]) * osc, __uint_as_float(acc[u * 8 + 7]) * osc);
                        sts128(stg_a + lane_idx * 64 + ((u ^ ((lane_idx >> 1) & 3)) << 4), v);
                    }
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int rr = i * 8 + (lane_idx >> 2), u = lane_idx & 3;
                        const uint4 v = lds128(stg_a + rr * 64 + ((u ^ ((rr >> 1) & 3)) << 4));
                        if (x0 + quad * 32 + rr < row_lim)
                            *reinterpret_cast<uint4*>(dst0 + (grow0 + rr) * ldd + c * 32 + u * 8) = v;
                    }
                    __syncwarp();
                }
            }
            if (n > 0) {
                tcgen05_fence_before();
                __syncwarp();
                if (lane_idx == 0) mbar_arrive(acc_free);
            }
            VLB_PROF(7);   // epilogue
        }
        if ((DBG & 8) && threadIdx.x == 64) {
            for (int i = 0; i < 10; ++i) atomicAdd(&g_bwd_prof[MODE * 16 + i], prof[i]);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp_idx == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

template <int DH, int MODE, int DBG = 0>
static int launch(const CUtensorMap& x1, const CUtensorMap& x2, const CUtensorMap& y1, const CUtensorMap& y2, const Params& p,
                  cudaStream_t s) {
    constexpr int NCH = DH / 64;
    constexpr int smem_bytes = 2 * NCH * BX * 128 + NST * 2 * NCH * BY * 128 + NST * 128 * 4 + NEW * 2048 + 256 + 1024;
    auto kern = attn_bwd_tc_kernel<DH, MODE, DBG>;
    static bool configured = false;
    if (!configured) {
        VLB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        configured = true;
    }
    const int grid = std::min(p.n_work, num_sms());
    kern<<<grid, NTHREADS, smem_bytes, s>>>(x1, x2, y1, y2, p);
    vlb::count_launch();
    VLB_LAUNCH_CHECK();
    return VLB200_OK;
}

// ------------------------------------------------------------------ backward: delta = rowsum(dO * O)
// row_starts (or null): first row of each sequence when the rows are ragged / packed; delta stays [B, H, S]
__global__ void attn_delta_kernel(const __nv_bfloat16* __restrict__ o, long long ldo, const __nv_bfloat16* __restrict__ dout,
                                  long long lddo, float* __restrict__ delta, const int* __restrict__ row_starts, uint32_t rows,
                                  int B, int S, int H, int DH) {
    // 16-byte loads: DH / 8 lanes cover one (row, head) -- 16 lanes at head_dim 128, 8 at 64 -- so a warp reduces 2 or 4 pairs
    // per pass (4-byte loads, one pair per warp, reached 29 % of the HBM bandwidth: profiles/r1b_ncu_summary.md); 32-bit index
    // arithmetic (with 64-bit divisions per pair the kernel was ALU-bound: 72 % ALU, profiles/r2ae_ncu_summary.md).
    const uint32_t lanes = (uint32_t)DH >> 3;   // lanes per (row, head)
    const uint32_t per_warp = 32u / lanes;
    const uint32_t warps_per_block = blockDim.x >> 5;
    const uint32_t total = rows * (uint32_t)H;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t sub = lane / lanes, li = lane % lanes;
    for (uint32_t w0 = (blockIdx.x * warps_per_block + (threadIdx.x >> 5)) * per_warp; w0 < total;
         w0 += gridDim.x * warps_per_block * per_warp) {
        const uint32_t w = w0 + sub;
        const bool live = w < total;
        const uint32_t row = live ? w / (uint32_t)H : 0u;   // b*S + t, or row_starts[b] + t
        const uint32_t h = live ? w - row * (uint32_t)H : 0u;
        float acc = 0.f;
        if (live) {
            const uint4 a = *reinterpret_cast<const uint4*>(o + (size_t)row * ldo + (size_t)h * DH + li * 8);
            const uint4 d = *reinterpret_cast<const uint4*>(dout + (size_t)row * lddo + (size_t)h * DH + li * 8);
            const uint32_t av[4] = {a.x, a.y, a.z, a.w}, dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 x = unpack_bf16x2(av[i]), y = unpack_bf16x2(dv[i]);
                acc += x.x * y.x + x.y * y.y;
            }
        }
        for (uint32_t off = lanes >> 1; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (live && li == 0) {
            uint32_t b = row / (uint32_t)S, t = row - b * (uint32_t)S;
            if (row_starts != nullptr) {
                b = 0;
                while (b + 1 < (uint32_t)B && (uint32_t)row_starts[b + 1] <= row) ++b;
                t = row - (uint32_t)row_starts[b];
            }
            if (t < (uint32_t)S) delta[((size_t)b * H + h) * S + t] = acc;
        }
    }
}

}  // namespace attn_bwd_tc
}  // namespace vlb

// diagnostics: read (and reset) the phase counters of the DBG-8 variants; not part of the ABI in include/vlb200.h
extern "C" int vlbdbg_attn_bwd_profile(unsigned long long* out32, int reset) {
    if (cudaMemcpyFromSymbol(out32, vlb::attn_bwd_tc::g_bwd_prof, 32 * sizeof(unsigned long long)) != cudaSuccess) return 1;
    if (reset) {
        unsigned long long z[32] = {0};
        if (cudaMemcpyToSymbol(vlb::attn_bwd_tc::g_bwd_prof, z, sizeof(z)) != cudaSuccess) return 1;
    }
    return 0;
}

extern "C" int vlbdbg_attn_bwd_trace(long long* out3072) {
    return cudaMemcpyFromSymbol(out3072, vlb::attn_bwd_tc::g_bwd_trace, 2 * 32 * 64 * sizeof(long long)) != cudaSuccess;
}

extern "C" int vlb200_attn_delta_varlen(const void* out, int64_t ldo, const void* dout, int64_t lddo, float* delta,
                                        const int* row_starts, int64_t total_rows, int B, int S, int H, int head_dim, void* stream) {
    VLB_REQUIRE(out && dout && delta, "attn_delta: null pointer");
    VLB_REQUIRE(head_dim >= 8 && head_dim <= 256 && (head_dim & (head_dim - 1)) == 0 && ldo % 8 == 0 && lddo % 8 == 0,
                "attn_delta: head_dim must be a power of two in [8, 256], row strides multiples of 8");
    VLB_REQUIRE(row_starts == nullptr || total_rows > 0, "attn_delta: row_starts needs total_rows");
    const size_t rows = row_starts ? (size_t)total_rows : (size_t)B * S;
    VLB_REQUIRE(rows * (size_t)H < (1ull << 31), "attn_delta: rows x heads must be below 2^31");
    const size_t nw = rows * H;   // (row, head) pairs; a warp takes 32 / (head_dim / 8) of them per pass
    const int blocks = (int)std::min<size_t>((nw * (head_dim / 8) / 32 + 7) / 8, (size_t)vlb::num_sms() * 16);
    vlb::attn_bwd_tc::attn_delta_kernel<<<blocks, 256, 0, vlb::as_stream(stream)>>>((const __nv_bfloat16*)out, ldo, (const __nv_bfloat16*)dout, lddo,
                                                                  delta, row_starts, (uint32_t)rows, B, S, H, head_dim);
    vlb::count_launch();
    VLB_LAUNCH_CHECK();
    return VLB200_OK;
}

extern "C" int vlb200_attn_delta(const void* out, int64_t ldo, const void* dout, int64_t lddo, float* delta, int B, int S, int H,
                                 int head_dim, void* stream) {
    return vlb200_attn_delta_varlen(out, ldo, dout, lddo, delta, nullptr, 0, B, S, H, head_dim, stream);
}

extern "C" int vlb200_attn_bwd_tc_ctx(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                      const void* out, int64_t ldo, const void* dout, int64_t lddo, const float* lse,
                                      float* delta, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv,
                                      const int* seqlens, const int* row_starts, const int* ctx, const int* kids,
                                      int64_t total_rows, int B, int S, int H, int KVH, int head_dim, int causal, float scale,
                                      void* stream) {
    using namespace vlb;
    using namespace vlb::attn_bwd_tc;
    VLB_REQUIRE(q && k && v && out && dout && lse && delta && dq && dk && dv, "attn_bwd_tc: null pointer");
    VLB_REQUIRE((ctx == nullptr) == (kids == nullptr), "attn_bwd_tc: ctx and kids come together");
    VLB_REQUIRE(ctx == nullptr || (row_starts != nullptr && causal), "attn_bwd_tc: context sequences need packed rows and causal attention");
    VLB_REQUIRE(row_starts == nullptr || (seqlens != nullptr && total_rows > 0), "attn_bwd_tc: row_starts needs seqlens and total_rows");
    VLB_REQUIRE(B > 0 && S > 0 && H > 0 && KVH > 0 && H % KVH == 0, "attn_bwd_tc: bad B/S/H/KVH");
    VLB_REQUIRE(head_dim == 64 || head_dim == 128, "attn_bwd_tc: head_dim %d unsupported (64 or 128)", head_dim);
    VLB_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && lddo % 8 == 0 && lddq % 8 == 0 && lddk % 8 == 0 && lddv % 8 == 0,
                "attn_bwd_tc: row strides must be multiples of 8");
    if (int rc = vlb200_attn_delta_varlen(out, ldo, dout, lddo, delta, row_starts, total_rows, B, S, H, head_dim, stream)) return rc;
    const uint64_t rows = row_starts ? (uint64_t)total_rows : (uint64_t)B * S;  // TMA zero-fills past the last row
    const uint64_t qcols = (uint64_t)H * head_dim, kcols = (uint64_t)KVH * head_dim;
    CUtensorMap xk, xv, yq, ydo, xq, xdo, yk, yv;
    int rc;
    if ((rc = gemm::get_tensor_map(k, kcols, rows, ldk, 64, BX, &xk))) return rc;
    if ((rc = gemm::get_tensor_map(v, kcols, rows, ldv, 64, BX, &xv))) return rc;
    if ((rc = gemm::get_tensor_map(q, qcols, rows, ldq, 64, BY, &yq))) return rc;
    if ((rc = gemm::get_tensor_map(dout, qcols, rows, lddo, 64, BY, &ydo))) return rc;
    if ((rc = gemm::get_tensor_map(q, qcols, rows, ldq, 64, BX, &xq))) return rc;
    if ((rc = gemm::get_tensor_map(dout, qcols, rows, lddo, 64, BX, &xdo))) return rc;
    if ((rc = gemm::get_tensor_map(k, kcols, rows, ldk, 64, BY, &yk))) return rc;
    if ((rc = gemm::get_tensor_map(v, kcols, rows, ldv, 64, BY, &yv))) return rc;
    Params p{};
    p.lse = lse; p.delta = delta; p.seqlens = seqlens; p.row_starts = row_starts; p.ctx = ctx; p.kids = kids;
    p.B = B; p.S = S; p.H = H; p.KVH = KVH; p.causal = causal; p.scale = scale;
    p.n_xb = (S + BX - 1) / BX;
    cudaStream_t s = as_stream(stream);
    // pass 1: dK, dV
    p.out1 = (__nv_bfloat16*)dv; p.ld1 = lddv; p.out2 = (__nv_bfloat16*)dk; p.ld2 = lddk;
    p.n_work = p.n_xb * KVH * B;
    static const int dbg = [] { const char* e = getenv("VLB200_ATTN_BWD_DBG"); return e ? atoi(e) : 0; }();
    if (dbg && head_dim == 128) {
        switch (dbg) {
#define VLB_DBG_CASE(D) case D: rc = launch<128, 0, D>(xk, xv, yq, ydo, p, s); if (rc) return rc; \
            p.out1 = nullptr; p.ld1 = 0; p.out2 = (__nv_bfloat16*)dq; p.ld2 = lddq; p.n_work = p.n_xb * H * B; \
            return launch<128, 1, D>(xq, xdo, yk, yv, p, s);
            VLB_DBG_CASE(8) VLB_DBG_CASE(16) VLB_DBG_CASE(32)
#undef VLB_DBG_CASE
            default: break;
        }
    }
    rc = head_dim == 64 ? launch<64, 0>(xk, xv, yq, ydo, p, s) : launch<128, 0>(xk, xv, yq, ydo, p, s);
    if (rc) return rc;
    // pass 2: dQ
    p.out1 = nullptr; p.ld1 = 0; p.out2 = (__nv_bfloat16*)dq; p.ld2 = lddq;
    p.n_work = p.n_xb * H * B;
    return head_dim == 64 ? launch<64, 1>(xq, xdo, yk, yv, p, s) : launch<128, 1>(xq, xdo, yk, yv, p, s);
}

extern "C" int vlb200_attn_bwd_tc_varlen(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                         const void* out, int64_t ldo, const void* dout, int64_t lddo, const float* lse,
                                         float* delta, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv,
                                         const int* seqlens, const int* row_starts, int64_t total_rows, int B, int S, int H,
                                         int KVH, int head_dim, int causal, float scale, void* stream) {
    return vlb200_attn_bwd_tc_ctx(q, ldq, k, ldk, v, ldv, out, ldo, dout, lddo, lse, delta, dq, lddq, dk, lddk, dv, lddv, seqlens,
                                  row_starts, nullptr, nullptr, total_rows, B, S, H, KVH, head_dim, causal, scale, stream);
}

extern "C" int vlb200_attn_bwd_tc(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                  const void* out, int64_t ldo, const void* dout, int64_t lddo, const float* lse, float* delta,
                                  void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv, const int* seqlens,
                                  int B, int S, int H, int KVH, int head_dim, int causal, float scale, void* stream) {
    return vlb200_attn_bwd_tc_varlen(q, ldq, k, ldk, v, ldv, out, ldo, dout, lddo, lse, delta, dq, lddq, dk, lddk, dv, lddv, seqlens,
                                     nullptr, 0, B, S, H, KVH, head_dim, causal, scale, stream);
}
