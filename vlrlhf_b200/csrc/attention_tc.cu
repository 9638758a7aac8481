// FlashAttention forward on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
//   S = Q K^T            tcgen05.mma 128x128x16, A = Q tile (smem, K-major), B = K tile (smem, K-major), D = TMEM
//   P = softmax-part(S)  4 softmax warps: thread r owns row r (TMEM lane r): tcgen05.ld, online max/sum in the
//                        log2 domain, bf16 P written to 128B-swizzled smem (K-major A operand)
//   O += P V             tcgen05.mma 128xDHx16, B = V tile (smem, MN-major), O accumulates in TMEM across KV tiles;
//                        rescaled lazily (only when the running max grows by > 2^8), normalised once at the end
//
// Persistent CTAs (grid = #SMs), warp 0 = TMA producer (Q tile + 2-stage K/V ring), warp 1 = MMA issuer,
// warps 2-5 = softmax/epilogue (TMEM lane quadrant = warp_idx % 4).  S is double buffered in TMEM so the
// QK^T of tile j+1 overlaps the softmax of tile j.
//
// Replaces LlamaAttention / CLIPAttention forward (modeling_llama.py:199-290, modeling_clip.py:261-334).
#include <algorithm>

#include "ptx.cuh"

namespace vlb {
namespace gemm {
int get_tensor_map(const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner, uint32_t box_outer,
                   CUtensorMap* out);
}
namespace attn_tc {

using namespace ptx;

constexpr int BM = 128, BN = 128;
constexpr int NTHREADS = 192;
constexpr float LOG2E_F = 1.4426950408889634f;
constexpr float LN2_F = 0.6931471805599453f;
constexpr float RESCALE_THRESHOLD = 8.0f;  // log2 units: P <= 2^8 before a lazy rescale of O

struct Params {
    __nv_bfloat16* o; long long ldo;
    float* lse;          // [B, H, S] or null
    const int* seqlens;  // [B] or null
    const int* row_starts;  // [B] or null: first row of each sequence (ragged / packed rows); null: sequence b starts at b*S
    const int* ctx;         // [B] or null: ctx[b] >= 0 names a sequence whose rows are extra keys/values visible to EVERY query of
                            // sequence b (the shared prompt+image prefix of a chosen/rejected pair); they precede b's own keys
    int B, S, H, KVH, causal;
    float scale;
    int n_qb, n_work;
};

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
          "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
          "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void work_coords(const Params& p, int w, int& b, int& h, int& qb) {
    const int bh = w / p.n_qb;
    qb = p.n_qb - 1 - (w - bh * p.n_qb);  // heavy (late) query tiles first within a head
    h = bh % p.H;
    b = bh / p.H;
}
__device__ __forceinline__ int num_kv_tiles(const Params& p, int b, int qb) {
    int kv_len = p.seqlens ? p.seqlens[b] : p.S;
    kv_len = max(kv_len, 1);
    int kmax = kv_len;
    if (p.causal) kmax = min(kmax, (qb + 1) * BM);
    return (kmax + BN - 1) / BN;
}
// KV tiles of one work item: n_ctx tiles of the context sequence (all keys visible), then n_self tiles of the sequence itself.
// skip: packed rows only -- a query tile wholly beyond its sequence is neither computed nor stored (padded layouts keep
// computing their padding rows: those are stored and must stay finite).
struct KvPlan { int n_ctx, n_self, ctx_row0, ctx_len; bool skip; };
__device__ __forceinline__ KvPlan kv_plan(const Params& p, int b, int qb) {
    KvPlan k;
    k.n_self = num_kv_tiles(p, b, qb);
    k.n_ctx = 0; k.ctx_row0 = 0; k.ctx_len = 0;
    k.skip = p.row_starts != nullptr && p.seqlens != nullptr && qb * BM >= p.seqlens[b];
    if (p.ctx != nullptr && !k.skip) {
        const int c = p.ctx[b];
        if (c >= 0) {
            k.ctx_len = max(min(p.seqlens[c], p.S), 0);
            k.ctx_row0 = p.row_starts[c];
            k.n_ctx = (k.ctx_len + BN - 1) / BN;
        }
    }
    return k;
}

// diagnostics (DBG bit 8): cycles one softmax thread per CTA spends in each phase of its loop, summed over the grid
__device__ unsigned long long g_fwd_prof[16];
#define VLB_PROF(i) do { if (DBG & 8) { const long long t_ = clock64(); prof[i] += (unsigned long long)(t_ - tp); tp = t_; } } while (0)

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// 2^x on the FMA / ALU pipes (no MUFU): n = round(x) by the 1.5 * 2^23 trick, 2^f on [-0.5, 0.5] by a degree-3 minimax
// polynomial (7.5e-5 relative: far inside bf16's 4e-3), n added to the exponent field.  x <= ~100; anything below -126
// (incl. -inf) gives ~1e-38.
__device__ __forceinline__ float ex2_poly(float x) {
    x = fmaxf(x, -126.f);
    const float t = x + 12582912.f;
    const float f = x - (t - 12582912.f);
    float p = fmaf(0.0551716685f, f, 0.2426111251f);
    p = fmaf(p, f, 0.6932609677f);
    p = fmaf(p, f, 0.9999280572f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
// (in the one-softmax-warp-per-scheduler kernels the polynomial for every fourth element -- Cody-Waite reduction + degree-3 minimax, 7.5e-5 relative --
// was measured SLOWER, 1390 vs 1202 cycles per tile for the exponential phase: one softmax warp per scheduler is bound by
// instruction issue, not by the MUFU: profiles/r2s_attn.log)

constexpr int NSTAGE = 3;  // K/V ring depth

// VAR 0: P goes through shared memory (SS-mode PV).  VAR 1: P stays in TENSOR MEMORY -- the softmax threads write bf16 P over
// the first 64 columns of the S buffer they just read (tcgen05.st) and PV runs with its A operand from TMEM (TS mode): 64 KB
// less shared-memory traffic per tile (32 KB of P stores + 32 KB of A reads out of 224 KB; the kernel is bound by the 128 B/clk
// shared-memory port, profiles/r2_attention_analysis.md).  VAR 2: Q is copied to TMEM once per query tile as well (S = Q K^T in
// TS mode): another 32 KB per tile.
template <int DH, int VAR>
__global__ void __launch_bounds__(NTHREADS, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                   const __grid_constant__ CUtensorMap tma_v, const Params p) {
    constexpr int NCH = DH / 64;                 // 64-column (128 B) chunks per row
    constexpr int CHUNK_BYTES = 128 * 128;       // [128 rows][128 B]
    constexpr int Q_BYTES = NCH * CHUNK_BYTES;
    constexpr int KV_BYTES = NCH * CHUNK_BYTES;  // one K or V tile
    constexpr int P_BYTES = 2 * CHUNK_BYTES;     // [128 q][128 keys] bf16
    constexpr bool TS = (VAR % 10) >= 1, QT = (VAR % 10) == 2;
    // diagnostics (wrong results, timing only), a bit mask: 1 = the MMA thread issues no MMAs (barriers only), 2 = the softmax
    // threads skip the exponentials, 4 = the producer loads no K/V tiles (stale shared memory)
    constexpr int DBG = VAR / 10;
    constexpr int KP_BYTES = TS ? KV_BYTES : (KV_BYTES > P_BYTES ? KV_BYTES : P_BYTES);  // K_j slot (VAR 0: reused for P_j once S_j has retired)
    constexpr int STAGE_BYTES = KP_BYTES + KV_BYTES;
    constexpr uint32_t TMEM_COLS = 512;
    constexpr uint32_t TM_S = 0, TM_O = 256, TM_Q = 384;  // S buffers at columns 0 / 128 (TS: P_j over the first 64 columns of S_j), O at 256, Q (QT) at 384

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sStage = sQ + Q_BYTES;  // NSTAGE x { KP slot, V slot }
    uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + NSTAGE * STAGE_BYTES);
    uint64_t* q_full = bars + 0;
    uint64_t* q_empty = bars + 1;
    uint64_t* kv_full = bars + 2;              // [NSTAGE]
    uint64_t* kv_empty = kv_full + NSTAGE;     // [NSTAGE]
    uint64_t* p_full = kv_empty + NSTAGE;      // [NSTAGE]
    uint64_t* s_full = p_full + NSTAGE;        // [2]
    uint64_t* s_empty = s_full + 2;            // [2]
    uint64_t* pv_done = s_empty + 2;
    uint64_t* o_free = pv_done + 1;
    uint64_t* qt_full = o_free + 1;            // QT: Q tile copied to tensor memory by the 4 softmax warps
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(qt_full + 1);

    const int warp_idx = threadIdx.x >> 5, lane_idx = threadIdx.x & 31;

    if (warp_idx == 0 && lane_idx == 0) {
        prefetch_tensormap(&tma_q); prefetch_tensormap(&tma_k); prefetch_tensormap(&tma_v);
        mbar_init(q_full, 1); mbar_init(q_empty, QT ? 4 : 1); mbar_init(qt_full, 4);
        for (int i = 0; i < NSTAGE; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); mbar_init(&p_full[i], 4); }
        for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 4); }
        mbar_init(pv_done, 1); mbar_init(o_free, 4);
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
            uint32_t item = 0, g = 0;  // g = global KV-tile counter (ring position)
            for (int w = blockIdx.x; w < p.n_work; w += gridDim.x) {
                int b, h, qb;
                work_coords(p, w, b, h, qb);
                const KvPlan plan = kv_plan(p, b, qb);
                if (plan.skip) continue;
                const int kvh = h / (p.H / p.KVH);
                const int n_tiles = plan.n_ctx + plan.n_self;
                const int row0 = p.row_starts ? p.row_starts[b] : b * p.S;
                mbar_wait(q_empty, (item & 1) ^ 1, 10);
                mbar_arrive_expect_tx(q_full, Q_BYTES);
#pragma unroll
                for (int c = 0; c < NCH; ++c) tma_load_2d(&tma_q, q_full, sQ + c * CHUNK_BYTES, h * DH + c * 64, row0 + qb * BM);
                for (int j = 0; j < n_tiles; ++j, ++g) {
                    const uint32_t st = g % NSTAGE;
                    mbar_wait(&kv_empty[st], ((g / NSTAGE) & 1) ^ 1, 20 + st);  // PV of the tile 3 back retired (K/P and V slots free)
                    if (DBG & 4) { mbar_arrive(&kv_full[st]); continue; }
                    mbar_arrive_expect_tx(&kv_full[st], 2 * KV_BYTES);
                    uint8_t* kp = sStage + st * STAGE_BYTES;
                    const int krow = j < plan.n_ctx ? plan.ctx_row0 + j * BN : row0 + (j - plan.n_ctx) * BN;
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {
                        tma_load_2d(&tma_k, &kv_full[st], kp + c * CHUNK_BYTES, kvh * DH + c * 64, krow);
                        tma_load_2d(&tma_v, &kv_full[st], kp + KP_BYTES + c * CHUNK_BYTES, kvh * DH + c * 64, krow);
                    }
                }
                ++item;
            }
        }
    } else if (warp_idx == 1) {
        // ===================== MMA issuer (whole warp converged; one elected lane issues: attention_bwd_tc.cu) ==========
        constexpr uint32_t idesc_s = make_idesc_bf16_f32(BM, BN, false, false);
        constexpr uint32_t idesc_o = make_idesc_bf16_f32(BM, DH, false, true);  // B = V is MN-major
        const uint64_t dq = make_smem_desc_sw128(smem_u32(sQ), 1024, 0);
        uint32_t item = 0, g0 = 0, sc = 0;
        for (int w = blockIdx.x; w < p.n_work; w += gridDim.x) {
            int b, h, qb;
            work_coords(p, w, b, h, qb);
            const KvPlan plan = kv_plan(p, b, qb);
            if (__shfl_sync(0xffffffffu, (int)plan.skip, 0)) continue;
            const int n_tiles = __shfl_sync(0xffffffffu, plan.n_ctx + plan.n_self, 0);
            if (QT) mbar_wait(qt_full, item & 1, 31);
            else mbar_wait(q_full, item & 1, 30);
            __syncwarp();
            tcgen05_fence_after();
            const uint32_t sc0 = sc;
            // Issue order: S_js as soon as its K tile and an S buffer are ready, else PV_jp once P_jp is published.  When only
            // one of the two can come next the warp BLOCKS on that barrier (mbarrier.try_wait suspends it; a spinning warp takes
            // issue slots from the softmax warp of its scheduler: attention_bwd_tc.cu, profiles/r2m_bwd_trace.log).
            int js = 0, jp = 0;
            auto issue_s = [&]() {
                const uint32_t g = g0 + js, st = g % NSTAGE, sb = sc & 1;
                tcgen05_fence_after();
                uint64_t dk = make_smem_desc_sw128(smem_u32(sStage + st * STAGE_BYTES), 1024, 0), aq = dq;
                asm volatile("" : "+l"(dk), "+l"(aq));   // opaque bases: per-k descriptors by immediate adds (attention_bwd_tc.cu)
                if (elect_one_sync()) {
                    if (!(DBG & 1)) {
#pragma unroll
                        for (int k = 0; k < DH / 16; ++k) {
                            const uint32_t off = ((k >> 2) * CHUNK_BYTES + (k & 3) * 32) >> 4;
                            if (QT) umma_f16_ts(tmem_base + TM_S + sb * BN, tmem_base + TM_Q + k * 8, dk + off, idesc_s, k != 0);
                            else umma_f16_ss(tmem_base + TM_S + sb * BN, aq + off, dk + off, idesc_s, k != 0);
                        }
                    }
                    umma_commit(&s_full[sb]);
                    if (!QT && js == n_tiles - 1) umma_commit(q_empty);  // Q tile free once the last S retires
                }
                __syncwarp();
                ++sc; ++js;
            };
            auto issue_pv = [&]() {
                const uint32_t g = g0 + jp, st = g % NSTAGE;
                if (jp == 0) { mbar_wait(o_free, (item & 1) ^ 1, 60); __syncwarp(); }  // epilogue of the previous item has read O
                tcgen05_fence_after();
                const uint32_t pbase = smem_u32(sStage + st * STAGE_BYTES), vbase = pbase + KP_BYTES;
                uint64_t dp = make_smem_desc_sw128(pbase, 1024, 0), dvv = make_smem_desc_sw128(vbase, 1024, CHUNK_BYTES);
                asm volatile("" : "+l"(dp), "+l"(dvv));
                if (elect_one_sync()) {
                    if (!(DBG & 1)) {
#pragma unroll
                        for (int k = 0; k < BN / 16; ++k) {
                            if (TS)
                                umma_f16_ts(tmem_base + TM_O, tmem_base + TM_S + ((sc0 + jp) & 1) * BN + k * 8,
                                            dvv + (uint64_t)(k * 128), idesc_o, (jp != 0 || k != 0) ? 1u : 0u);
                            else
                                umma_f16_ss(tmem_base + TM_O, dp + (uint64_t)(((k >> 2) * CHUNK_BYTES + (k & 3) * 32) >> 4),
                                            dvv + (uint64_t)(k * 128), idesc_o, (jp != 0 || k != 0) ? 1u : 0u);
                        }
                    }
                    umma_commit(pv_done);
                    umma_commit(&kv_empty[st]);
                }
                __syncwarp();
                ++jp;
            };
            long long t_spin = 0;
            while (jp < n_tiles) {
                // S buffer free: VAR 0 -- the softmax threads have read it (s_empty); TS -- P_j lives in S_j's buffer until PV_j
                // has consumed it: the MMAs of one thread execute in issue order, so S_{j+2} may be issued once PV_j has
                const bool s_wanted = js < n_tiles && (!TS || js - jp < 2);
                if (!s_wanted) {                 // PV_jp is the only thing that can come next
                    const uint32_t g = g0 + jp;
                    mbar_wait(&p_full[g % NSTAGE], (g / NSTAGE) & 1, 42);
                    __syncwarp();
                    issue_pv();
                    continue;
                }
                const uint32_t gs = g0 + js, gp = g0 + jp;
                bool s_ready = __all_sync(0xffffffffu, mbar_test_wait(&kv_full[gs % NSTAGE], (gs / NSTAGE) & 1));
                if (s_ready && !TS) s_ready = __all_sync(0xffffffffu, mbar_test_wait(&s_empty[sc & 1], ((sc >> 1) & 1) ^ 1));
                if (s_ready) { issue_s(); t_spin = 0; continue; }
                if (jp < js && __all_sync(0xffffffffu, mbar_test_wait(&p_full[gp % NSTAGE], (gp / NSTAGE) & 1))) { issue_pv(); t_spin = 0; continue; }
                if (jp == js && TS) {            // nothing to multiply by V yet: the K tile is the only thing to wait for
                    mbar_wait(&kv_full[gs % NSTAGE], (gs / NSTAGE) & 1, 43);
                    __syncwarp();
                    continue;
                }
                __nanosleep(32);
                if (t_spin == 0) t_spin = clock64();
                else if (clock64() - t_spin > VLB_WATCHDOG_CYCLES) {
                    if (lane_idx == 0) printf("[vlb200] attn_fwd_tc MMA watchdog: block %d js %d jp %d n %d\n", blockIdx.x, js, jp, n_tiles);
                    __trap();
                }
            }
            g0 += n_tiles;
            ++item;
        }
    } else {
        // ===================== softmax + epilogue (4 warps, one row per thread) =====================
        const int quad = warp_idx & 3;
        const int r = quad * 32 + lane_idx;  // row inside the Q tile == TMEM lane
        const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
        const float sl2 = p.scale * LOG2E_F;
        uint32_t sc = 0, g = 0, item = 0;  // g = global tile counter (== number of PVs issued for earlier tiles)
        unsigned long long prof[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        long long tp = clock64();
        for (int w = blockIdx.x; w < p.n_work; w += gridDim.x) {
            int b, h, qb;
            work_coords(p, w, b, h, qb);
            const KvPlan plan = kv_plan(p, b, qb);
            if (plan.skip) continue;
            const int n_tiles = plan.n_ctx + plan.n_self;
            if (DBG & 8) { prof[8] += 1; prof[9] += n_tiles; }
            if (QT) {
                // Q tile: smem (TMA, 128B-swizzled) -> this thread's TMEM lane, two bf16 per column (the K-major A operand of
                // S = Q K^T).  Every MMA of the previous item has retired (its epilogue waited for the last PV).
                mbar_wait(q_full, item & 1, 70);
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        uint32_t qw[16];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const uint4 v = *reinterpret_cast<const uint4*>(sQ + c * CHUNK_BYTES + r * 128 + (((hh * 4 + u) ^ (r & 7)) << 4));
                            qw[u * 4 + 0] = v.x; qw[u * 4 + 1] = v.y; qw[u * 4 + 2] = v.z; qw[u * 4 + 3] = v.w;
                        }
                        tmem_st_32x32_x16(tmem_base + lane_addr + TM_Q + c * 32 + hh * 16, qw);
                    }
                }
                tmem_st_wait_all();
                tcgen05_fence_before();
                __syncwarp();
                if (lane_idx == 0) { mbar_arrive(qt_full); mbar_arrive(q_empty); }
                ++item;
            }
            int kv_len = p.seqlens ? p.seqlens[b] : p.S;
            kv_len = max(kv_len, 1);
            const int qrow = qb * BM + r;
            float m_run = -INFINITY, l_run = 0.f;
            VLB_PROF(0);   // item start (plan, Q copy)
            for (int j = 0; j < n_tiles; ++j, ++sc, ++g) {
                const uint32_t sb = sc & 1, st = g % NSTAGE;
                mbar_wait(&s_full[sb], (sc >> 1) & 1, 80 + sb);
                tcgen05_fence_after();
                VLB_PROF(1);   // wait for S
                uint32_t sr[4][32];
#pragma unroll
                for (int c = 0; c < 4; ++c) tmem_ld_32x32(tmem_base + lane_addr + TM_S + sb * BN + c * 32, sr[c]);
                tmem_ld_wait();
                tcgen05_fence_before();
                __syncwarp();
                if (!TS && lane_idx == 0) mbar_arrive(&s_empty[sb]);  // S buffer is in registers now
                VLB_PROF(2);   // TMEM -> registers
                // context tiles: every key below ctx_len is visible; own tiles: keys below kv_len, up to the causal diagonal
                const bool is_ctx = j < plan.n_ctx;
                const int k0 = (is_ctx ? j : j - plan.n_ctx) * BN;
                const int vlim = (is_ctx ? plan.ctx_len : kv_len) - k0;              // keys [0, vlim) of the tile exist
                const int clim = (p.causal && !is_ctx) ? qrow - k0 : BN;             // keys [0, clim] of the tile are not in the future
                const bool need_mask = vlim < BN || (p.causal && !is_ctx && k0 + BN > qb * BM);
                // row max over 8 independent chains (one dependent chain of 64 3-input max instructions cost 538 cycles per
                // tile: profiles/r2j_attn_phases.log)
                if (need_mask) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const int kk = c * 32 + i;
                            if (kk >= vlim || kk > clim) sr[c][i] = 0xff800000u;  // -inf
                        }
                    }
                }
                // four independent chains of 3-input max (FMNMX3): 64 instructions, dependent depth 16
                float mxp[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    mxp[c] = fmaxf(__uint_as_float(sr[c][0]), __uint_as_float(sr[c][1]));
#pragma unroll
                    for (int i = 2; i < 32; i += 2) mxp[c] = fmaxf(fmaxf(mxp[c], __uint_as_float(sr[c][i])), __uint_as_float(sr[c][i + 1]));
                }
                float mx = fmaxf(fmaxf(mxp[0], mxp[1]), fmaxf(mxp[2], mxp[3]));
                mx *= sl2;  // scale > 0: max commutes with the scaling (log2 domain from here on)
                // lazy rescale: keep the stale max unless the new one exceeds it by more than 2^RESCALE_THRESHOLD
                float corr = 1.f;
                bool need = false;
                if (mx > m_run + RESCALE_THRESHOLD || m_run == -INFINITY) {
                    const float m_new = mx == -INFINITY ? m_run : mx;
                    if (m_run != -INFINITY && m_new != m_run) { corr = ex2_approx(m_run - m_new); need = true; }
                    m_run = m_new;
                }
                const float neg_m = m_run == -INFINITY ? 0.f : -m_run;
                VLB_PROF(3);   // mask + row max
                // P = exp2(s*c - m) (bf16) into the K slot of this stage (K_j is dead: S_j has retired), laid out as a
                // K-major 128B-swizzled A operand: chunk = 64 keys, row pitch 128 B
                uint8_t* sP = sStage + st * STAGE_BYTES;
                float rs0 = 0.f, rs1 = 0.f, rs2 = 0.f, rs3 = 0.f;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t wt[16];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        uint32_t w4[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            float p0 = fmaf(__uint_as_float(sr[c][u * 8 + 2 * e]), sl2, neg_m);
                            float p1 = fmaf(__uint_as_float(sr[c][u * 8 + 2 * e + 1]), sl2, neg_m);
                            if (!(DBG & 2)) {
                                p0 = ex2_approx(p0);
                                p1 = ex2_approx(p1);
                            }
                            w4[e] = pack_bf16x2(p0, p1);
                            if (e & 1) { rs2 += p0; rs3 += p1; }
                            else { rs0 += p0; rs1 += p1; }
                        }
                        if (TS) {   // P_j over S_j's first 64 columns: keys 2i | 2i+1 in column i of this thread's lane
                            wt[u * 4 + 0] = w4[0]; wt[u * 4 + 1] = w4[1]; wt[u * 4 + 2] = w4[2]; wt[u * 4 + 3] = w4[3];
                        } else {
                            const int unit = (c & 1) * 4 + u;  // 16-byte unit inside the 64-key chunk (c >> 1)
                            *reinterpret_cast<uint4*>(sP + (c >> 1) * CHUNK_BYTES + r * 128 + ((unit ^ (r & 7)) << 4)) =
                                make_uint4(w4[0], w4[1], w4[2], w4[3]);
                        }
                    }
                    if (TS) tmem_st_32x32_x16(tmem_base + lane_addr + TM_S + sb * BN + c * 16, wt);
                }
                l_run = l_run * corr + ((rs0 + rs1) + (rs2 + rs3));
                VLB_PROF(4);   // exponentials, row sum, pack, P store issue
                // publish P only after PV of the previous tile has retired: keeps the pv_done phase bookkeeping exact
                // (a waiter never runs two phases ahead) and orders the (rare) O correction before the next PV
                if (j > 0) mbar_wait(pv_done, (g - 1) & 1, 90);
                VLB_PROF(5);   // wait for the previous PV
                if (j > 0 && __any_sync(0xffffffffu, need)) {  // warp-uniform: tcgen05.ld/st are warp collectives
                    tcgen05_fence_after();
#pragma unroll
                    for (int c = 0; c < DH / 32; ++c) {
                        uint32_t orow[32];
                        tmem_ld_32x32(tmem_base + lane_addr + TM_O + c * 32, orow);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) orow[i] = __float_as_uint(__uint_as_float(orow[i]) * corr);
                        tmem_st_32x32(tmem_base + lane_addr + TM_O + c * 32, orow);
                    }
                    tmem_st_wait();
                }
                if (TS) tmem_st_wait_all();   // P_j is in tensor memory
                else fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor-core (async) proxy
                tcgen05_fence_before();
                __syncwarp();
                if (lane_idx == 0) mbar_arrive(&p_full[st]);
                VLB_PROF(6);   // O correction (rare), store completion, fences, publish
            }
            // ---- epilogue: wait for the last PV, normalise, store O and LSE
            mbar_wait(pv_done, (g - 1) & 1, 95);
            tcgen05_fence_after();
            const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
            // packed rows: the tile may run into the next sequence's rows, so only the attended prefix is stored
            const bool valid = qrow < (p.row_starts && p.seqlens ? min(p.seqlens[b], p.S) : p.S);
            const long long row0 = p.row_starts ? p.row_starts[b] : (long long)b * p.S;
            __nv_bfloat16* op = p.o + (row0 + qrow) * p.ldo + (long long)h * DH;
#pragma unroll
            for (int c = 0; c < DH / 32; ++c) {
                uint32_t orow[32];
                tmem_ld_32x32(tmem_base + lane_addr + TM_O + c * 32, orow);
                tmem_ld_wait();
                if (valid) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        uint4 v;
                        v.x = pack_bf16x2(__uint_as_float(orow[u * 8 + 0]) * inv, __uint_as_float(orow[u * 8 + 1]) * inv);
                        v.y = pack_bf16x2(__uint_as_float(orow[u * 8 + 2]) * inv, __uint_as_float(orow[u * 8 + 3]) * inv);
                        v.z = pack_bf16x2(__uint_as_float(orow[u * 8 + 4]) * inv, __uint_as_float(orow[u * 8 + 5]) * inv);
                        v.w = pack_bf16x2(__uint_as_float(orow[u * 8 + 6]) * inv, __uint_as_float(orow[u * 8 + 7]) * inv);
                        *reinterpret_cast<uint4*>(op + c * 32 + u * 8) = v;
                    }
                }
            }
            if (valid && p.lse) p.lse[((long long)b * p.H + h) * p.S + qrow] = l_run > 0.f ? (m_run + log2f(l_run)) * LN2_F : -INFINITY;
            tcgen05_fence_before();
            __syncwarp();
            if (lane_idx == 0) mbar_arrive(o_free);
            VLB_PROF(7);   // epilogue
        }
        if ((DBG & 8) && threadIdx.x == 64) {
            for (int i = 0; i < 10; ++i) atomicAdd(&g_fwd_prof[i], prof[i]);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp_idx == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------------------------------
// Second-generation forward (VLB200_ATTN_FWD_VARIANT=4).  Same arithmetic as VAR 1 (P in tensor memory, TS-mode PV); what
// changes is the plumbing around the softmax warps, whose serial per-tile chain bounds the kernel (profiles/r2s_attn.log:
// the per-item epilogue cost 4 000 cycles and the pipeline refill ~1 600 per query tile):
//   * K and V tiles travel through SEPARATE two-slot rings (a K slot is free as soon as S_j has retired, long before V_j's):
//     128 KB instead of 192 KB, which pays for a second Q buffer and an epilogue staging block;
//   * Q is double buffered and the tile counter runs ACROSS query tiles: the S warp issues S_0 / S_1 of the next query tile
//     while the softmax warps are still in the epilogue of the previous one;
//   * S = Q K^T and O += P V are issued by two converged warps in a fixed order (one elected lane each; attention_bwd_tc.cu);
//     an S buffer is recycled when the PV that read its P has COMPLETED (pv_done[buffer]);
//   * the epilogue leaves through a swizzled 2 KB staging block per warp as 64-byte row segments (8 rows per store
//     instruction instead of 32 scattered 16-byte pieces).
constexpr int NTHREADS2 = 224;  // TMA producer, S-MMA warp, 4 softmax warps, PV-MMA warp
constexpr int NKS = 2, NVS = 2;

template <int DH, int DBG>
__global__ void __launch_bounds__(NTHREADS2, 1)
attn_fwd_tc2_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                    const __grid_constant__ CUtensorMap tma_v, const Params p) {
    constexpr int NCH = DH / 64;
    constexpr int CHUNK_BYTES = 128 * 128;
    constexpr int T_BYTES = NCH * CHUNK_BYTES;   // one Q, K or V tile
    constexpr uint32_t TMEM_COLS = 512;
    // S buffers at columns 0 / 128, O at 256, bf16 P buffers (two keys per column) at 384 / 448.  P has its OWN columns: with
    // P_g written over S_g, S_{g+2} had to wait for the COMPLETION of PV_g (another warp issues it) and that chain -- publish,
    // PV warp wake-up, 8 MMAs, commit, S warp wake-up, 8 MMAs -- was as long as a softmax tile (200 cycles of waiting per tile)
    constexpr uint32_t TM_S = 0, TM_O = 256, TM_P = 384;

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                          // [2]
    uint8_t* sK = sQ + 2 * T_BYTES;              // [NKS]
    uint8_t* sV = sK + NKS * T_BYTES;            // [NVS]
    uint8_t* sOut = sV + NVS * T_BYTES;          // [4 warps][2][32 rows][64 B]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sOut + 4 * 4096);
    uint64_t* q_full = bars;                     // [2]
    uint64_t* q_empty = q_full + 2;              // [2]
    uint64_t* k_full = q_empty + 2;              // [NKS]
    uint64_t* k_empty = k_full + NKS;            // [NKS]
    uint64_t* v_full = k_empty + NKS;            // [NVS]
    uint64_t* v_empty = v_full + NVS;            // [NVS]
    uint64_t* s_full = v_empty + NVS;            // [2]
    uint64_t* s_empty = s_full + 2;              // [2]
    uint64_t* p_full = s_empty + 2;              // [2]
    uint64_t* pv_done = p_full + 2;              // [2]
    uint64_t* o_free = pv_done + 2;
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(o_free + 1);

    const int warp_idx = threadIdx.x >> 5, lane_idx = threadIdx.x & 31;
    if (warp_idx == 0 && lane_idx == 0) {
        prefetch_tensormap(&tma_q); prefetch_tensormap(&tma_k); prefetch_tensormap(&tma_v);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 4); mbar_init(&p_full[i], 4);
            mbar_init(&pv_done[i], 1);
        }
        for (int i = 0; i < NKS; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
        for (int i = 0; i < NVS; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
        mbar_init(o_free, 4);
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
            uint32_t item = 0, g = 0;  // g = global KV-tile counter
            for (int w = blockIdx.x; w < p.n_work; w += gridDim.x) {
                int b, h, qb;
                work_coords(p, w, b, h, qb);
                const KvPlan plan = kv_plan(p, b, qb);
                if (plan.skip) continue;
                const int kvh = h / (p.H / p.KVH);
                const int n_tiles = plan.n_ctx + plan.n_self;
                const int row0 = p.row_starts ? p.row_starts[b] : b * p.S;
                const uint32_t qs = item & 1;
                mbar_wait(&q_empty[qs], ((item >> 1) & 1) ^ 1, 10);
                mbar_arrive_expect_tx(&q_full[qs], T_BYTES);
#pragma unroll
                for (int c = 0; c < NCH; ++c) tma_load_2d(&tma_q, &q_full[qs], sQ + qs * T_BYTES + c * CHUNK_BYTES, h * DH + c * 64, row0 + qb * BM);
                for (int j = 0; j < n_tiles; ++j, ++g) {
                    const int krow = j < plan.n_ctx ? plan.ctx_row0 + j * BN : row0 + (j - plan.n_ctx) * BN;
                    const uint32_t ks = g % NKS, vs = g % NVS;
                    mbar_wait(&k_empty[ks], ((g / NKS) & 1) ^ 1, 20);   // S of the tile two back has retired
                    mbar_arrive_expect_tx(&k_full[ks], T_BYTES);
#pragma unroll
                    for (int c = 0; c < NCH; ++c) tma_load_2d(&tma_k, &k_full[ks], sK + ks * T_BYTES + c * CHUNK_BYTES, kvh * DH + c * 64, krow);
                    mbar_wait(&v_empty[vs], ((g / NVS) & 1) ^ 1, 21);   // PV of the tile two back has retired
                    mbar_arrive_expect_tx(&v_full[vs], T_BYTES);
#pragma unroll
                    for (int c = 0; c < NCH; ++c) tma_load_2d(&tma_v, &v_full[vs], sV + vs * T_BYTES + c * CHUNK_BYTES, kvh * DH + c * 64, krow);
                }
                ++item;
            }
        }
    } else if (warp_idx == 1) {
        // ===================== S = Q K^T (whole warp converged; one elected lane issues) =====================
        constexpr uint32_t idesc_s = make_idesc_bf16_f32(BM, BN, false, false);
        uint32_t item = 0, g = 0;
        for (int w = blockIdx.x; w < p.n_work; w += gridDim.x) {
            int b, h, qb;
            work_coords(p, w, b, h, qb);
            const KvPlan plan = kv_plan(p, b, qb);
            if (__shfl_sync(0xffffffffu, (int)plan.skip, 0)) continue;
            const int n_tiles = __shfl_sync(0xffffffffu, plan.n_ctx + plan.n_self, 0);
            const uint32_t qs = item & 1;
            mbar_wait(&q_full[qs], (item >> 1) & 1, 30);
            for (int j = 0; j < n_tiles; ++j, ++g) {
                const uint32_t sb = g & 1, ks = g % NKS;
                mbar_wait(&k_full[ks], (g / NKS) & 1, 31);
                mbar_wait(&s_empty[sb], ((g >> 1) & 1) ^ 1, 32);   // the softmax warps have S_{g-2} in registers
                __syncwarp();
                tcgen05_fence_after();
                uint64_t dq = make_smem_desc_sw128(smem_u32(sQ + qs * T_BYTES), 1024, 0);
                uint64_t dk = make_smem_desc_sw128(smem_u32(sK + ks * T_BYTES), 1024, 0);
                asm volatile("" : "+l"(dq), "+l"(dk));   // opaque bases: per-k descriptors by immediate adds on the uniform datapath
                if (elect_one_sync()) {
#pragma unroll
                    for (int k = 0; k < DH / 16; ++k) {
                        const uint32_t off = ((k >> 2) * CHUNK_BYTES + (k & 3) * 32) >> 4;
                        umma_f16_ss(tmem_base + TM_S + sb * BN, dq + off, dk + off, idesc_s, k != 0);
                    }
                    umma_commit(&s_full[sb]);
                    umma_commit(&k_empty[ks]);
                    if (j == n_tiles - 1) umma_commit(&q_empty[qs]);  // Q tile free once the last S retires
                }
                __syncwarp();
            }
            ++item;
        }
    } else if (warp_idx == 6) {
        // ===================== O += P V (whole warp converged; one elected lane issues) =====================
        constexpr uint32_t idesc_o = make_idesc_bf16_f32(BM, DH, false, true);  // B = V is MN-major
        uint32_t item = 0, g = 0;
        for (int w = blockIdx.x; w < p.n_work; w += gridDim.x) {
            int b, h, qb;
            work_coords(p, w, b, h, qb);
            const KvPlan plan = kv_plan(p, b, qb);
            if (__shfl_sync(0xffffffffu, (int)plan.skip, 0)) continue;
            const int n_tiles = __shfl_sync(0xffffffffu, plan.n_ctx + plan.n_self, 0);
            for (int j = 0; j < n_tiles; ++j, ++g) {
                const uint32_t sb = g & 1, vs = g % NVS;
                mbar_wait(&p_full[sb], (g >> 1) & 1, 40);
                mbar_wait(&v_full[vs], (g / NVS) & 1, 41);
                if (j == 0) mbar_wait(o_free, (item & 1) ^ 1, 42);   // the epilogue of the previous query tile has read O
                __syncwarp();
                tcgen05_fence_after();
                uint64_t dvv = make_smem_desc_sw128(smem_u32(sV + vs * T_BYTES), 1024, CHUNK_BYTES);
                asm volatile("" : "+l"(dvv));
                if (elect_one_sync()) {
#pragma unroll
                    for (int k = 0; k < BN / 16; ++k)
                        umma_f16_ts(tmem_base + TM_O, tmem_base + TM_P + sb * 64 + k * 8, dvv + (uint64_t)(k * 128), idesc_o,
                                    (j != 0 || k != 0) ? 1u : 0u);
                    umma_commit(&pv_done[sb]);
                    umma_commit(&v_empty[vs]);
                }
                __syncwarp();
            }
            ++item;
        }
    } else {
        // ===================== softmax + epilogue (4 warps, one row per thread) =====================
        const int quad = warp_idx & 3;
        const int r = quad * 32 + lane_idx;  // row inside the Q tile == TMEM lane
        const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
        const float sl2 = p.scale * LOG2E_F;
        const uint32_t stg_a = smem_u32(sOut + quad * 4096);
        uint32_t g = 0;
        unsigned long long prof[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        long long tp = clock64();
        for (int w = blockIdx.x; w < p.n_work; w += gridDim.x) {
            int b, h, qb;
            work_coords(p, w, b, h, qb);
            const KvPlan plan = kv_plan(p, b, qb);
            if (plan.skip) continue;
            const int n_tiles = plan.n_ctx + plan.n_self;
            if (DBG & 8) { prof[8] += 1; prof[9] += n_tiles; }
            int kv_len = p.seqlens ? p.seqlens[b] : p.S;
            kv_len = max(kv_len, 1);
            const int qrow = qb * BM + r;
            float m_run = -INFINITY, l_run = 0.f;
            VLB_PROF(0);   // item start
            for (int j = 0; j < n_tiles; ++j, ++g) {
                const uint32_t sb = g & 1;
                mbar_wait(&s_full[sb], (g >> 1) & 1, 80 + sb);
                tcgen05_fence_after();
                if ((DBG & 8) && j == 0) { const long long t_ = clock64(); prof[10] += (unsigned long long)(t_ - tp); tp = t_; }   // first tile of a query tile
                VLB_PROF(1);   // wait for S
                uint32_t sr[4][32];
#pragma unroll
                for (int c = 0; c < 4; ++c) tmem_ld_32x32(tmem_base + lane_addr + TM_S + sb * BN + c * 32, sr[c]);
                tmem_ld_wait();
                tcgen05_fence_before();
                __syncwarp();
                if (lane_idx == 0) mbar_arrive(&s_empty[sb]);   // S_g is in registers: the buffer may take S_{g+2}
                VLB_PROF(2);   // TMEM -> registers
                // context tiles: every key below ctx_len is visible; own tiles: keys below kv_len, up to the causal diagonal
                const bool is_ctx = j < plan.n_ctx;
                const int k0 = (is_ctx ? j : j - plan.n_ctx) * BN;
                const int vlim = (is_ctx ? plan.ctx_len : kv_len) - k0;              // keys [0, vlim) of the tile exist
                const int clim = (p.causal && !is_ctx) ? qrow - k0 : BN;             // keys [0, clim] of the tile are not in the future
                const bool need_mask = vlim < BN || (p.causal && !is_ctx && k0 + BN > qb * BM);
                if (need_mask) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const int kk = c * 32 + i;
                            if (kk >= vlim || kk > clim) sr[c][i] = 0xff800000u;  // -inf
                        }
                    }
                }
                // four independent chains of 3-input max (FMNMX3): 64 instructions, dependent depth 16
                float mxp[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    mxp[c] = fmaxf(__uint_as_float(sr[c][0]), __uint_as_float(sr[c][1]));
#pragma unroll
                    for (int i = 2; i < 32; i += 2) mxp[c] = fmaxf(fmaxf(mxp[c], __uint_as_float(sr[c][i])), __uint_as_float(sr[c][i + 1]));
                }
                float mx = fmaxf(fmaxf(mxp[0], mxp[1]), fmaxf(mxp[2], mxp[3]));
                mx *= sl2;  // scale > 0: max commutes with the scaling (log2 domain from here on)
                // lazy rescale: keep the stale max unless the new one exceeds it by more than 2^RESCALE_THRESHOLD
                float corr = 1.f;
                bool need = false;
                if (mx > m_run + RESCALE_THRESHOLD || m_run == -INFINITY) {
                    const float m_new = mx == -INFINITY ? m_run : mx;
                    if (m_run != -INFINITY && m_new != m_run) { corr = ex2_approx(m_run - m_new); need = true; }
                    m_run = m_new;
                }
                const float neg_m = m_run == -INFINITY ? 0.f : -m_run;
                VLB_PROF(3);   // mask + row max
                // P = exp2(s*c - m) (bf16) into P buffer g & 1: keys 2i | 2i+1 in column i of this thread's lane.  PV_{g-2}, which
                // read this buffer, has completed: the previous tile waited for PV_{g-2}... (pv_done below), the previous query
                // tile's epilogue for its last PV
                float rs0 = 0.f, rs1 = 0.f, rs2 = 0.f, rs3 = 0.f;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t wt[16];
#pragma unroll
                    for (int u = 0; u < 16; ++u) {
                        const float p0 = ex2_approx(fmaf(__uint_as_float(sr[c][2 * u]), sl2, neg_m));
                        const float p1 = ex2_approx(fmaf(__uint_as_float(sr[c][2 * u + 1]), sl2, neg_m));
                        wt[u] = pack_bf16x2(p0, p1);
                        if (u & 1) { rs2 += p0; rs3 += p1; }
                        else { rs0 += p0; rs1 += p1; }
                    }
                    tmem_st_32x32_x16(tmem_base + lane_addr + TM_P + sb * 64 + c * 16, wt);
                }
                l_run = l_run * corr + ((rs0 + rs1) + (rs2 + rs3));
                VLB_PROF(4);   // exponentials, row sum, pack, P store issue
                // the (rare) O correction needs the previous PV retired; waiting for it every tile also keeps the pv_done phase
                // bookkeeping exact (a waiter never falls two phases behind)
                if (j > 0) mbar_wait(&pv_done[(g - 1) & 1], ((g - 1) >> 1) & 1, 90);
                VLB_PROF(5);   // wait for the previous PV
                if (j > 0 && __any_sync(0xffffffffu, need)) {  // warp-uniform: tcgen05.ld/st are warp collectives
                    tcgen05_fence_after();
#pragma unroll
                    for (int c = 0; c < DH / 32; ++c) {
                        uint32_t orow[32];
                        tmem_ld_32x32(tmem_base + lane_addr + TM_O + c * 32, orow);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) orow[i] = __float_as_uint(__uint_as_float(orow[i]) * corr);
                        tmem_st_32x32(tmem_base + lane_addr + TM_O + c * 32, orow);
                    }
                }
                tmem_st_wait_all();   // P_g (and the corrected O) are in tensor memory
                tcgen05_fence_before();
                __syncwarp();
                if (lane_idx == 0) mbar_arrive(&p_full[sb]);
                VLB_PROF(6);   // O correction (rare), store completion, fences, publish
            }
            // ---- epilogue: wait for the last PV, normalise, store O (through the warp's staging block) and LSE
            mbar_wait(&pv_done[(g - 1) & 1], ((g - 1) >> 1) & 1, 95);
            tcgen05_fence_after();
            if (DBG & 8) { const long long t_ = clock64(); prof[11] += (unsigned long long)(t_ - tp); tp = t_; }   // wait for the last PV
            const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
            // packed rows: the tile may run into the next sequence's rows, so only the attended prefix is stored
            const int row_lim = (p.row_starts && p.seqlens) ? min(p.seqlens[b], p.S) : p.S;
            const long long row0 = p.row_starts ? p.row_starts[b] : (long long)b * p.S;
            __nv_bfloat16* obase = p.o + (row0 + qb * BM + quad * 32) * p.ldo + (long long)h * DH;
#pragma unroll
            for (int c2 = 0; c2 < DH / 64; ++c2) {   // two 32-column chunks per round (independent work for one warp's scheduler)
                uint32_t orow[2][32];
                tmem_ld_32x32(tmem_base + lane_addr + TM_O + c2 * 64, orow[0]);
                tmem_ld_32x32(tmem_base + lane_addr + TM_O + c2 * 64 + 32, orow[1]);
                tmem_ld_wait();
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        uint4 v;
                        v.x = pack_bf16x2(__uint_as_float(orow[hh][u * 8 + 0]) * inv, __uint_as_float(orow[hh][u * 8 + 1]) * inv);
                        v.y = pack_bf16x2(__uint_as_float(orow[hh][u * 8 + 2]) * inv, __uint_as_float(orow[hh][u * 8 + 3]) * inv);
                        v.z = pack_bf16x2(__uint_as_float(orow[hh][u * 8 + 4]) * inv, __uint_as_float(orow[hh][u * 8 + 5]) * inv);
                        v.w = pack_bf16x2(__uint_as_float(orow[hh][u * 8 + 6]) * inv, __uint_as_float(orow[hh][u * 8 + 7]) * inv);
                        sts128(stg_a + hh * 2048 + lane_idx * 64 + ((u ^ ((lane_idx >> 1) & 3)) << 4), v);
                    }
                }
                __syncwarp();
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int rr = i * 8 + (lane_idx >> 2), u = lane_idx & 3;
                        const uint4 v = lds128(stg_a + hh * 2048 + rr * 64 + ((u ^ ((rr >> 1) & 3)) << 4));
                        if (qb * BM + quad * 32 + rr < row_lim)
                            *reinterpret_cast<uint4*>(obase + (long long)rr * p.ldo + c2 * 64 + hh * 32 + u * 8) = v;
                    }
                }
                __syncwarp();
            }
            if (qrow < row_lim && p.lse) p.lse[((long long)b * p.H + h) * p.S + qrow] = l_run > 0.f ? (m_run + log2f(l_run)) * LN2_F : -INFINITY;
            tcgen05_fence_before();
            __syncwarp();
            if (lane_idx == 0) mbar_arrive(o_free);
            VLB_PROF(7);   // epilogue
        }
        if ((DBG & 8) && threadIdx.x == 64) {
            for (int i = 0; i < 12; ++i) atomicAdd(&g_fwd_prof[i], prof[i]);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp_idx == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

template <int DH, int DBG>
static int launch2(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const Params& p, cudaStream_t s) {
    constexpr int t_bytes = (DH / 64) * 128 * 128;
    constexpr int smem_bytes = (2 + NKS + NVS) * t_bytes + 4 * 4096 + 256 + 1024;
    auto kern = attn_fwd_tc2_kernel<DH, DBG>;
    static bool configured = false;
    if (!configured) {
        VLB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        configured = true;
    }
    const int grid = std::min(p.n_work, num_sms());
    kern<<<grid, NTHREADS2, smem_bytes, s>>>(tq, tk, tv, p);
    count_launch();
    VLB_LAUNCH_CHECK();
    return VLB200_OK;
}

// ------------------------------------------------------------------------------------------------------------------------
// Third-generation forward (VLB200_ATTN_FWD_VARIANT=5): the second generation with EIGHT softmax warps -- two per TMEM lane
// quadrant, each thread owning 64 of its row's 128 score columns.  One softmax warp per scheduler issues 452 instructions per
// tile at 2.6 cycles each (profiles/r2s_attn.log); two warps per scheduler interleave.  The pair exchanges its partial row
// maxima through shared memory behind a 64-thread named barrier (so both keep the SAME running max and take the same lazy
// rescale decisions), keeps partial row sums that meet only in the epilogue, and splits the O columns for the (rare)
// correction and for the epilogue.
// (second generation, for reference:) VLB200_ATTN_FWD_VARIANT=4.  Same arithmetic as VAR 1 (P in tensor memory, TS-mode PV); what
// changes is the plumbing around the softmax warps, whose serial per-tile chain bounds the kernel (profiles/r2s_attn.log:
// the per-item epilogue cost 4 000 cycles and the pipeline refill ~1 600 per query tile):
//   * K and V tiles travel through SEPARATE two-slot rings (a K slot is free as soon as S_j has retired, long before V_j's):
//     128 KB instead of 192 KB, which pays for a second Q buffer and an epilogue staging block;
//   * Q is double buffered and the tile counter runs ACROSS query tiles: the S warp issues S_0 / S_1 of the next query tile
//     while the softmax warps are still in the epilogue of the previous one;
//   * S = Q K^T and O += P V are issued by two converged warps in a fixed order (one elected lane each; attention_bwd_tc.cu);
//     an S buffer is recycled when the PV that read its P has COMPLETED (pv_done[buffer]);
//   * the epilogue leaves through a swizzled 2 KB staging block per warp as 64-byte row segments (8 rows per store
//     instruction instead of 32 scattered 16-byte pieces).
constexpr int NTHREADS3 = 352;  // TMA producer, S-MMA warp, 8 softmax warps, PV-MMA warp

template <int DH, int DBG>
__global__ void __launch_bounds__(NTHREADS3, 1)
attn_fwd_tc3_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                    const __grid_constant__ CUtensorMap tma_v, const Params p) {
    constexpr int NCH = DH / 64;
    constexpr int CHUNK_BYTES = 128 * 128;
    constexpr int T_BYTES = NCH * CHUNK_BYTES;   // one Q, K or V tile
    constexpr uint32_t TMEM_COLS = 512;
    // S buffers at columns 0 / 128, O at 256, bf16 P buffers (two keys per column) at 384 / 448.  P has its OWN columns: with
    // P_g written over S_g, S_{g+2} had to wait for the COMPLETION of PV_g (another warp issues it) and that chain -- publish,
    // PV warp wake-up, 8 MMAs, commit, S warp wake-up, 8 MMAs -- was as long as a softmax tile (200 cycles of waiting per tile)
    constexpr uint32_t TM_S = 0, TM_O = 256, TM_P = 384;

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                          // [2]
    uint8_t* sK = sQ + 2 * T_BYTES;              // [NKS]
    uint8_t* sV = sK + NKS * T_BYTES;            // [NVS]
    uint8_t* sOut = sV + NVS * T_BYTES;          // [8 warps][32 rows][64 B]
    float* sMx = reinterpret_cast<float*>(sOut + 8 * 2048);   // [2 tile parities][2 halves][128 rows]: partial row maxima
    float* sL = sMx + 2 * 2 * 128;                            // [2 halves][128 rows]: partial row sums (epilogue)
    uint64_t* bars = reinterpret_cast<uint64_t*>(sL + 2 * 128);
    uint64_t* q_full = bars;                     // [2]
    uint64_t* q_empty = q_full + 2;              // [2]
    uint64_t* k_full = q_empty + 2;              // [NKS]
    uint64_t* k_empty = k_full + NKS;            // [NKS]
    uint64_t* v_full = k_empty + NKS;            // [NVS]
    uint64_t* v_empty = v_full + NVS;            // [NVS]
    uint64_t* s_full = v_empty + NVS;            // [2]
    uint64_t* s_empty = s_full + 2;              // [2]
    uint64_t* p_full = s_empty + 2;              // [2]
    uint64_t* pv_done = p_full + 2;              // [2]
    uint64_t* o_free = pv_done + 2;
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(o_free + 1);

    const int warp_idx = threadIdx.x >> 5, lane_idx = threadIdx.x & 31;
    if (warp_idx == 0 && lane_idx == 0) {
        prefetch_tensormap(&tma_q); prefetch_tensormap(&tma_k); prefetch_tensormap(&tma_v);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 8); mbar_init(&p_full[i], 8);
            mbar_init(&pv_done[i], 1);
        }
        for (int i = 0; i < NKS; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
        for (int i = 0; i < NVS; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
        mbar_init(o_free, 8);
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
            uint32_t item = 0, g = 0;  // g = global KV-tile counter
            for (int w = blockIdx.x; w < p.n_work; w += gridDim.x) {
                int b, h, qb;
                work_coords(p, w, b, h, qb);
                const KvPlan plan = kv_plan(p, b, qb);
                if (plan.skip) continue;
                const int kvh = h / (p.H / p.KVH);
                const int n_tiles = plan.n_ctx + plan.n_self;
                const int row0 = p.row_starts ? p.row_starts[b] : b * p.S;
                const uint32_t qs = item & 1;
                mbar_wait_relaxed(&q_empty[qs], ((item >> 1) & 1) ^ 1, 10);
                mbar_arrive_expect_tx(&q_full[qs], T_BYTES);
#pragma unroll
                for (int c = 0; c < NCH; ++c) tma_load_2d(&tma_q, &q_full[qs], sQ + qs * T_BYTES + c * CHUNK_BYTES, h * DH + c * 64, row0 + qb * BM);
                for (int j = 0; j < n_tiles; ++j, ++g) {
                    const int krow = j < plan.n_ctx ? plan.ctx_row0 + j * BN : row0 + (j - plan.n_ctx) * BN;
                    const uint32_t ks = g % NKS, vs = g % NVS;
                    mbar_wait_relaxed(&k_empty[ks], ((g / NKS) & 1) ^ 1, 20);   // S of the tile two back has retired
                    mbar_arrive_expect_tx(&k_full[ks], T_BYTES);
#pragma unroll
                    for (int c = 0; c < NCH; ++c) tma_load_2d(&tma_k, &k_full[ks], sK + ks * T_BYTES + c * CHUNK_BYTES, kvh * DH + c * 64, krow);
                    mbar_wait_relaxed(&v_empty[vs], ((g / NVS) & 1) ^ 1, 21);   // PV of the tile two back has retired
                    mbar_arrive_expect_tx(&v_full[vs], T_BYTES);
#pragma unroll
                    for (int c = 0; c < NCH; ++c) tma_load_2d(&tma_v, &v_full[vs], sV + vs * T_BYTES + c * CHUNK_BYTES, kvh * DH + c * 64, krow);
                }
                ++item;
            }
        }
    } else if (warp_idx == 1) {
        // ===================== S = Q K^T (whole warp converged; one elected lane issues) =====================
        constexpr uint32_t idesc_s = make_idesc_bf16_f32(BM, BN, false, false);
        uint32_t item = 0, g = 0;
        for (int w = blockIdx.x; w < p.n_work; w += gridDim.x) {
            int b, h, qb;
            work_coords(p, w, b, h, qb);
            const KvPlan plan = kv_plan(p, b, qb);
            if (__shfl_sync(0xffffffffu, (int)plan.skip, 0)) continue;
            const int n_tiles = __shfl_sync(0xffffffffu, plan.n_ctx + plan.n_self, 0);
            const uint32_t qs = item & 1;
            mbar_wait_relaxed(&q_full[qs], (item >> 1) & 1, 30);
            for (int j = 0; j < n_tiles; ++j, ++g) {
                const uint32_t sb = g & 1, ks = g % NKS;
                mbar_wait_relaxed(&k_full[ks], (g / NKS) & 1, 31);
                mbar_wait_relaxed(&s_empty[sb], ((g >> 1) & 1) ^ 1, 32);   // the softmax warps have S_{g-2} in registers
                __syncwarp();
                tcgen05_fence_after();
                uint64_t dq = make_smem_desc_sw128(smem_u32(sQ + qs * T_BYTES), 1024, 0);
                uint64_t dk = make_smem_desc_sw128(smem_u32(sK + ks * T_BYTES), 1024, 0);
                asm volatile("" : "+l"(dq), "+l"(dk));   // opaque bases: per-k descriptors by immediate adds on the uniform datapath
                if (elect_one_sync()) {
#pragma unroll
                    for (int k = 0; k < DH / 16; ++k) {
                        const uint32_t off = ((k >> 2) * CHUNK_BYTES + (k & 3) * 32) >> 4;
                        umma_f16_ss(tmem_base + TM_S + sb * BN, dq + off, dk + off, idesc_s, k != 0);
                    }
                    umma_commit(&s_full[sb]);
                    umma_commit(&k_empty[ks]);
                    if (j == n_tiles - 1) umma_commit(&q_empty[qs]);  // Q tile free once the last S retires
                }
                __syncwarp();
            }
            ++item;
        }
    } else if (warp_idx == 10) {
        // ===================== O += P V (whole warp converged; one elected lane issues) =====================
        constexpr uint32_t idesc_o = make_idesc_bf16_f32(BM, DH, false, true);  // B = V is MN-major
        uint32_t item = 0, g = 0;
        for (int w = blockIdx.x; w < p.n_work; w += gridDim.x) {
            int b, h, qb;
            work_coords(p, w, b, h, qb);
            const KvPlan plan = kv_plan(p, b, qb);
            if (__shfl_sync(0xffffffffu, (int)plan.skip, 0)) continue;
            const int n_tiles = __shfl_sync(0xffffffffu, plan.n_ctx + plan.n_self, 0);
            for (int j = 0; j < n_tiles; ++j, ++g) {
                const uint32_t sb = g & 1, vs = g % NVS;
                mbar_wait_relaxed(&p_full[sb], (g >> 1) & 1, 40);
                mbar_wait_relaxed(&v_full[vs], (g / NVS) & 1, 41);
                if (j == 0) mbar_wait_relaxed(o_free, (item & 1) ^ 1, 42);   // the epilogue of the previous query tile has read O
                __syncwarp();
                tcgen05_fence_after();
                uint64_t dvv = make_smem_desc_sw128(smem_u32(sV + vs * T_BYTES), 1024, CHUNK_BYTES);
                asm volatile("" : "+l"(dvv));
                if (elect_one_sync()) {
#pragma unroll
                    for (int k = 0; k < BN / 16; ++k)
                        umma_f16_ts(tmem_base + TM_O, tmem_base + TM_P + sb * 64 + k * 8, dvv + (uint64_t)(k * 128), idesc_o,
                                    (j != 0 || k != 0) ? 1u : 0u);
                    umma_commit(&pv_done[sb]);
                    umma_commit(&v_empty[vs]);
                }
                __syncwarp();
            }
            ++item;
        }
    } else {
        // ===================== softmax + epilogue (8 warps: row = TMEM lane, two warps split the 128 score columns) ==========
        const int quad = warp_idx & 3;
        const int half = (warp_idx - 2) >> 2;   // 0: score columns [0, 64), 1: [64, 128)
        const int r = quad * 32 + lane_idx;     // row inside the Q tile == TMEM lane
        const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
        const float sl2 = p.scale * LOG2E_F;
        const uint32_t stg_a = smem_u32(sOut + (warp_idx - 2) * 2048);
        constexpr int OC = DH / 2;              // O columns of this warp: [half * OC, half * OC + OC)
        auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory"); };
        const uint32_t smx_a = smem_u32(sMx), sl_a = smem_u32(sL);
        uint32_t g = 0;
        unsigned long long prof[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        long long tp = clock64();
        for (int w = blockIdx.x; w < p.n_work; w += gridDim.x) {
            int b, h, qb;
            work_coords(p, w, b, h, qb);
            const KvPlan plan = kv_plan(p, b, qb);
            if (plan.skip) continue;
            const int n_tiles = plan.n_ctx + plan.n_self;
            if (DBG & 8) { prof[8] += 1; prof[9] += n_tiles; }
            int kv_len = p.seqlens ? p.seqlens[b] : p.S;
            kv_len = max(kv_len, 1);
            const int qrow = qb * BM + r;
            float m_run = -INFINITY, l_part = 0.f;   // l_part: this warp's 64 columns only
            VLB_PROF(0);   // item start
            for (int j = 0; j < n_tiles; ++j, ++g) {
                const uint32_t sb = g & 1;
                mbar_wait(&s_full[sb], (g >> 1) & 1, 80 + sb);
                tcgen05_fence_after();
                if ((DBG & 8) && j == 0) { const long long t_ = clock64(); prof[10] += (unsigned long long)(t_ - tp); tp = t_; }   // first tile of a query tile
                VLB_PROF(1);   // wait for S
                uint32_t sr[2][32];
#pragma unroll
                for (int c = 0; c < 2; ++c) tmem_ld_32x32(tmem_base + lane_addr + TM_S + sb * BN + half * 64 + c * 32, sr[c]);
                tmem_ld_wait();
                tcgen05_fence_before();
                __syncwarp();
                if (lane_idx == 0) mbar_arrive(&s_empty[sb]);   // this warp's half of S_g is in registers
                VLB_PROF(2);   // TMEM -> registers
                // context tiles: every key below ctx_len is visible; own tiles: keys below kv_len, up to the causal diagonal
                const bool is_ctx = j < plan.n_ctx;
                const int k0 = (is_ctx ? j : j - plan.n_ctx) * BN;
                const int vlim = (is_ctx ? plan.ctx_len : kv_len) - k0;              // keys [0, vlim) of the tile exist
                const int clim = (p.causal && !is_ctx) ? qrow - k0 : BN;             // keys [0, clim] of the tile are not in the future
                const bool need_mask = vlim < BN || (p.causal && !is_ctx && k0 + BN > qb * BM);
                if (need_mask) {
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const int kk = half * 64 + c * 32 + i;
                            if (kk >= vlim || kk > clim) sr[c][i] = 0xff800000u;  // -inf
                        }
                    }
                }
                float mxp[2];
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    mxp[c] = fmaxf(__uint_as_float(sr[c][0]), __uint_as_float(sr[c][1]));
#pragma unroll
                    for (int i = 2; i < 32; i += 2) mxp[c] = fmaxf(fmaxf(mxp[c], __uint_as_float(sr[c][i])), __uint_as_float(sr[c][i + 1]));
                }
                // the row maximum over all 128 columns: partial maxima meet through shared memory (slot by tile parity: the
                // partner reads slot g before it can pass the barrier of tile g + 1, so tile g + 2 may overwrite it)
                float mx = fmaxf(mxp[0], mxp[1]);
                sts_f32(smx_a + (((sb * 2 + half) * 128 + r) << 2), mx);
                pair_sync();
                mx = fmaxf(mx, lds_f32(smx_a + (((sb * 2 + (half ^ 1)) * 128 + r) << 2)));
                mx *= sl2;  // scale > 0: max commutes with the scaling (log2 domain from here on)
                // lazy rescale: keep the stale max unless the new one exceeds it by more than 2^RESCALE_THRESHOLD
                float corr = 1.f;
                bool need = false;
                if (mx > m_run + RESCALE_THRESHOLD || m_run == -INFINITY) {
                    const float m_new = mx == -INFINITY ? m_run : mx;
                    if (m_run != -INFINITY && m_new != m_run) { corr = ex2_approx(m_run - m_new); need = true; }
                    m_run = m_new;
                }
                const float neg_m = m_run == -INFINITY ? 0.f : -m_run;
                VLB_PROF(3);   // mask + row max
                // P = exp2(s*c - m) (bf16) into P buffer g & 1, this warp's 32 columns (keys 2i | 2i+1 of its half in column i)
                // two elements per instruction where the pipe allows it (FFMA2 for the scale + shift, FADD2 for the row sum): the
                // phase costs (non-MUFU instructions) + (MUFU instructions), serialised -- 399 + 732 cycles, profiles/r2ah
                const uint64_t sl2p = pack_f32x2(sl2, sl2), negp = pack_f32x2(neg_m, neg_m);
                uint64_t rsa = pack_f32x2(0.f, 0.f), rsb = rsa;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t wt[16];
#pragma unroll
                    for (int u = 0; u < 16; ++u) {
                        const uint64_t xp = fma_f32x2(pack_f32x2(__uint_as_float(sr[c][2 * u]), __uint_as_float(sr[c][2 * u + 1])), sl2p, negp);
                        float x0, x1;
                        unpack_f32x2(xp, x0, x1);
                        const float p0 = (DBG & 2) ? x0 : ex2_approx(x0);   // (DBG 2 / 16: timing diagnostics without ex2 / without the bf16 pack)
                        // DBG bit 4 (variant 45): every fourth exponential on the FMA pipe
                        const float p1 = (DBG & 2) ? x1 : ((DBG & 4) && (u & 1)) ? ex2_poly(x1) : ex2_approx(x1);
                        wt[u] = (DBG & 16) ? (__float_as_uint(p0) ^ (__float_as_uint(p1) >> 16)) : pack_bf16x2(p0, p1);
                        if (u & 1) rsb = add_f32x2(rsb, pack_f32x2(p0, p1));
                        else rsa = add_f32x2(rsa, pack_f32x2(p0, p1));
                    }
                    tmem_st_32x32_x16(tmem_base + lane_addr + TM_P + sb * 64 + half * 32 + c * 16, wt);
                }
                float rs0, rs1, rs2, rs3;
                unpack_f32x2(rsa, rs0, rs1);
                unpack_f32x2(rsb, rs2, rs3);
                l_part = l_part * corr + ((rs0 + rs1) + (rs2 + rs3));
                VLB_PROF(4);   // exponentials, row sum, pack, P store issue
                // the (rare) O correction needs the previous PV retired; waiting for it every tile also keeps the pv_done phase
                // bookkeeping exact (a waiter never falls two phases behind)
                if (j > 0) mbar_wait(&pv_done[(g - 1) & 1], ((g - 1) >> 1) & 1, 90);
                VLB_PROF(5);   // wait for the previous PV
                if (j > 0 && __any_sync(0xffffffffu, need)) {  // warp-uniform (and equal in both warps of the pair: same rows, same max)
                    tcgen05_fence_after();
#pragma unroll
                    for (int c = 0; c < OC / 32; ++c) {
                        uint32_t orow[32];
                        tmem_ld_32x32(tmem_base + lane_addr + TM_O + half * OC + c * 32, orow);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) orow[i] = __float_as_uint(__uint_as_float(orow[i]) * corr);
                        tmem_st_32x32(tmem_base + lane_addr + TM_O + half * OC + c * 32, orow);
                    }
                }
                tmem_st_wait_all();   // P_g (and the corrected O) are in tensor memory
                tcgen05_fence_before();
                __syncwarp();
                if (lane_idx == 0) mbar_arrive(&p_full[sb]);
                VLB_PROF(6);   // O correction (rare), store completion, fences, publish
            }
            // ---- epilogue: wait for the last PV, normalise, store this warp's O columns (through its staging block) and LSE
            // (everything that does not need O first: the last PV is still executing)
            sts_f32(sl_a + ((half * 128 + r) << 2), l_part);
            pair_sync();
            const float l_run = l_part + lds_f32(sl_a + (((half ^ 1) * 128 + r) << 2));
            pair_sync();   // (the partner has read this item's value before the next item overwrites it)
            const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
            // packed rows: the tile may run into the next sequence's rows, so only the attended prefix is stored
            const int row_lim = (p.row_starts && p.seqlens) ? min(p.seqlens[b], p.S) : p.S;
            const long long row0 = p.row_starts ? p.row_starts[b] : (long long)b * p.S;
            __nv_bfloat16* obase = p.o + (row0 + qb * BM + quad * 32) * p.ldo + (long long)h * DH + half * OC;
            if (half == 0 && qrow < row_lim && p.lse)
                p.lse[((long long)b * p.H + h) * p.S + qrow] = l_run > 0.f ? (m_run + log2f(l_run)) * LN2_F : -INFINITY;
            if (DBG & 8) { const long long t_ = clock64(); prof[12] += (unsigned long long)(t_ - tp); tp = t_; }   // row-sum exchange, 1/l, LSE, addresses
            mbar_wait(&pv_done[(g - 1) & 1], ((g - 1) >> 1) & 1, 95);
            tcgen05_fence_after();
            if (DBG & 8) { const long long t_ = clock64(); prof[11] += (unsigned long long)(t_ - tp); tp = t_; }   // wait for the last PV
#pragma unroll
            for (int c = 0; c < OC / 32; ++c) {
                uint32_t orow[32];
                tmem_ld_32x32(tmem_base + lane_addr + TM_O + half * OC + c * 32, orow);
                tmem_ld_wait();
                if (DBG & 8) { const long long t_ = clock64(); prof[14] += (unsigned long long)(t_ - tp); tp = t_; }   // O chunk: TMEM load
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    uint4 v;
                    v.x = pack_bf16x2(__uint_as_float(orow[u * 8 + 0]) * inv, __uint_as_float(orow[u * 8 + 1]) * inv);
                    v.y = pack_bf16x2(__uint_as_float(orow[u * 8 + 2]) * inv, __uint_as_float(orow[u * 8 + 3]) * inv);
                    v.z = pack_bf16x2(__uint_as_float(orow[u * 8 + 4]) * inv, __uint_as_float(orow[u * 8 + 5]) * inv);
                    v.w = pack_bf16x2(__uint_as_float(orow[u * 8 + 6]) * inv, __uint_as_float(orow[u * 8 + 7]) * inv);
                    sts128(stg_a + lane_idx * 64 + ((u ^ ((lane_idx >> 1) & 3)) << 4), v);
                }
                __syncwarp();
                if (DBG & 8) { const long long t_ = clock64(); prof[15] += (unsigned long long)(t_ - tp); tp = t_; }   // O chunk: scale, pack, staging store
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int rr = i * 8 + (lane_idx >> 2), u = lane_idx & 3;
                    const uint4 v = lds128(stg_a + rr * 64 + ((u ^ ((rr >> 1) & 3)) << 4));
                    if (qb * BM + quad * 32 + rr < row_lim) *reinterpret_cast<uint4*>(obase + (long long)rr * p.ldo + c * 32 + u * 8) = v;
                }
                __syncwarp();
                if (DBG & 8) { const long long t_ = clock64(); prof[13] += (unsigned long long)(t_ - tp); tp = t_; }   // O chunk: staging load, global store
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane_idx == 0) mbar_arrive(o_free);
            VLB_PROF(7);   // epilogue
        }
        if ((DBG & 8) && threadIdx.x == 64) {
            for (int i = 0; i < 16; ++i) atomicAdd(&g_fwd_prof[i], prof[i]);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp_idx == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

template <int DH, int DBG>
static int launch3(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const Params& p, cudaStream_t s) {
    constexpr int t_bytes = (DH / 64) * 128 * 128;
    constexpr int smem_bytes = (2 + NKS + NVS) * t_bytes + 8 * 2048 + (2 * 2 * 128 + 2 * 128) * 4 + 256 + 1024;
    auto kern = attn_fwd_tc3_kernel<DH, DBG>;
    static bool configured = false;
    if (!configured) {
        VLB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        configured = true;
    }
    const int grid = std::min(p.n_work, num_sms());
    kern<<<grid, NTHREADS3, smem_bytes, s>>>(tq, tk, tv, p);
    count_launch();
    VLB_LAUNCH_CHECK();
    return VLB200_OK;
}

template <int DH, int VAR>
static int launch(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const Params& p, cudaStream_t s) {
    constexpr int kv_bytes = (DH / 64) * 128 * 128;
    constexpr int kp_bytes = (VAR % 10) >= 1 ? kv_bytes : (kv_bytes > 2 * 128 * 128 ? kv_bytes : 2 * 128 * 128);
    constexpr int smem_bytes = kv_bytes + NSTAGE * (kp_bytes + kv_bytes) + 256 + 1024;
    auto kern = attn_fwd_tc_kernel<DH, VAR>;
    static bool configured = false;
    if (!configured) {
        VLB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        configured = true;
    }
    const int grid = std::min(p.n_work, num_sms());
    kern<<<grid, NTHREADS, smem_bytes, s>>>(tq, tk, tv, p);
    count_launch();
    VLB_LAUNCH_CHECK();
    return VLB200_OK;
}

}  // namespace attn_tc
}  // namespace vlb

static int& attn_fwd_variant() {
    static int v = [] { const char* e = getenv("VLB200_ATTN_FWD_VARIANT"); return e ? atoi(e) : 5; }();
    return v;
}
extern "C" int vlb200_set_attn_fwd_variant(int variant) {
    const int prev = attn_fwd_variant();
    if (variant >= 0) attn_fwd_variant() = variant;
    return prev;
}

// diagnostics: read (and reset) the phase counters of the DBG-8 variants; not part of the ABI in include/vlb200.h
extern "C" int vlbdbg_attn_fwd_profile(unsigned long long* out16, int reset) {
    if (cudaMemcpyFromSymbol(out16, vlb::attn_tc::g_fwd_prof, 16 * sizeof(unsigned long long)) != cudaSuccess) return 1;
    if (reset) {
        unsigned long long z[16] = {0};
        if (cudaMemcpyToSymbol(vlb::attn_tc::g_fwd_prof, z, sizeof(z)) != cudaSuccess) return 1;
    }
    return 0;
}

extern "C" int vlb200_attn_fwd_tc_ctx(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                      void* out, int64_t ldo, float* lse, const int* seqlens, const int* row_starts,
                                      const int* ctx, int64_t total_rows, int B, int S, int H, int KVH, int head_dim, int causal,
                                      float scale, void* stream) {
    using namespace vlb;
    VLB_REQUIRE(q && k && v && out, "attn_fwd_tc: null pointer");
    VLB_REQUIRE(ctx == nullptr || (row_starts != nullptr && causal), "attn_fwd_tc: context sequences need packed rows and causal attention");
    VLB_REQUIRE(row_starts == nullptr || (seqlens != nullptr && total_rows > 0), "attn_fwd_tc: row_starts needs seqlens and total_rows");
    VLB_REQUIRE(B > 0 && S > 0 && H > 0 && KVH > 0 && H % KVH == 0, "attn_fwd_tc: bad B/S/H/KVH");
    VLB_REQUIRE(head_dim == 64 || head_dim == 128, "attn_fwd_tc: head_dim %d unsupported (64 or 128)", head_dim);
    VLB_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0, "attn_fwd_tc: row strides must be multiples of 8");
    // ragged rows: the maps end at total_rows, so a tile that runs past the last sequence is zero-filled by TMA
    const uint64_t rows = row_starts ? (uint64_t)total_rows : (uint64_t)B * S;
    CUtensorMap tq, tk, tv;
    int rc;
    if ((rc = gemm::get_tensor_map(q, (uint64_t)H * head_dim, rows, ldq, 64, 128, &tq))) return rc;
    if ((rc = gemm::get_tensor_map(k, (uint64_t)KVH * head_dim, rows, ldk, 64, 128, &tk))) return rc;
    if ((rc = gemm::get_tensor_map(v, (uint64_t)KVH * head_dim, rows, ldv, 64, 128, &tv))) return rc;
    attn_tc::Params p{};
    p.o = (__nv_bfloat16*)out; p.ldo = ldo; p.lse = lse; p.seqlens = seqlens; p.row_starts = row_starts; p.ctx = ctx;
    p.B = B; p.S = S; p.H = H; p.KVH = KVH; p.causal = causal; p.scale = scale;
    p.n_qb = (S + attn_tc::BM - 1) / attn_tc::BM;
    p.n_work = p.n_qb * H * B;
    // VLB200_ATTN_FWD_VARIANT: 5 (default) = third-generation kernel (attn_fwd_tc3_kernel, eight softmax warps), 4 = second
    // generation (attn_fwd_tc2_kernel, four); first generation: 0 = P through shared memory, 1 = P in tensor memory, 2 = P and Q
    // in tensor memory; 45 = variant 5 with every fourth exponential as a polynomial; 80 / 81 / 84 / 85 / 125 = variants
    // 0 / 1 / 4 / 5 / 45 with per-phase cycle counters (tests/attn_phase_probe.py)
    const int variant = attn_fwd_variant();
    cudaStream_t st = as_stream(stream);
    if (head_dim == 64) {
        if (variant == 0) return attn_tc::launch<64, 0>(tq, tk, tv, p, st);
        if (variant == 1) return attn_tc::launch<64, 1>(tq, tk, tv, p, st);
        if (variant == 2) return attn_tc::launch<64, 2>(tq, tk, tv, p, st);
        if (variant == 4) return attn_tc::launch2<64, 0>(tq, tk, tv, p, st);
        return attn_tc::launch3<64, 0>(tq, tk, tv, p, st);
    }
    switch (variant) {
        case 0: return attn_tc::launch<128, 0>(tq, tk, tv, p, st);
        case 1: return attn_tc::launch<128, 1>(tq, tk, tv, p, st);
        case 2: return attn_tc::launch<128, 2>(tq, tk, tv, p, st);
        case 80: return attn_tc::launch<128, 80>(tq, tk, tv, p, st);
        case 81: return attn_tc::launch<128, 81>(tq, tk, tv, p, st);
        case 84: return attn_tc::launch2<128, 8>(tq, tk, tv, p, st);
        case 4: return attn_tc::launch2<128, 0>(tq, tk, tv, p, st);
        case 85: return attn_tc::launch3<128, 8>(tq, tk, tv, p, st);
        case 45: return attn_tc::launch3<128, 4>(tq, tk, tv, p, st);
        case 125: return attn_tc::launch3<128, 12>(tq, tk, tv, p, st);
        case 105: return attn_tc::launch3<128, 10>(tq, tk, tv, p, st);
        case 245: return attn_tc::launch3<128, 24>(tq, tk, tv, p, st);
        case 265: return attn_tc::launch3<128, 26>(tq, tk, tv, p, st);
        default: return attn_tc::launch3<128, 0>(tq, tk, tv, p, st);
    }
}

extern "C" int vlb200_attn_fwd_tc_varlen(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                         void* out, int64_t ldo, float* lse, const int* seqlens, const int* row_starts,
                                         int64_t total_rows, int B, int S, int H, int KVH, int head_dim, int causal, float scale,
                                         void* stream) {
    return vlb200_attn_fwd_tc_ctx(q, ldq, k, ldk, v, ldv, out, ldo, lse, seqlens, row_starts, nullptr, total_rows, B, S, H, KVH,
                                  head_dim, causal, scale, stream);
}

extern "C" int vlb200_attn_fwd_tc(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                  void* out, int64_t ldo, float* lse, const int* seqlens, int B, int S, int H, int KVH,
                                  int head_dim, int causal, float scale, void* stream) {
    return vlb200_attn_fwd_tc_varlen(q, ldq, k, ldk, v, ldv, out, ldo, lse, seqlens, nullptr, 0, B, S, H, KVH, head_dim, causal,
                                     scale, stream);
}
