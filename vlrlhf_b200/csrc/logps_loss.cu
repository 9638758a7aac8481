// K16 + K18: fused log-prob gather and preference loss.
//   vlb200_logps_fwd/bwd  <- VLDPOTrainer.get_batch_logps   (reference base/trainer.py:148-188)
//   vlb200_dpo_loss       <- VLDPOTrainer.dpo_loss          (reference base/trainer.py:244-301)
// HBM-bound: one CTA per logits row, 128-bit streaming loads, online log-sum-exp in fp32, warp-shuffle
// reductions; never materialises log_softmax.  Deterministic (no atomics).
#include "common.cuh"

namespace vlb {

constexpr int LOGPS_THREADS = 256;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

struct MS {  // running (max, sum of exp2((x - max) * log2e)) in the log2 domain
    float m, s;
};
__device__ __forceinline__ MS ms_combine(MS a, MS b) {
    const float m = fmaxf(a.m, b.m);
    if (m == -INFINITY) return MS{m, 0.f};
    return MS{m, a.s * exp2f(a.m - m) + b.s * exp2f(b.m - m)};
}

template <typename T>
struct Vec;  // 16-byte vector of T
template <>
struct Vec<__nv_bfloat16> {
    static constexpr int N = 8;
    __device__ static void load(const __nv_bfloat16* p, float (&f)[8]) {
        const uint4 u = ld_nc_v4(p);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 t = unpack_bf16x2(w[i]);
            f[2 * i] = t.x;
            f[2 * i + 1] = t.y;
        }
    }
};
template <>
struct Vec<float> {
    static constexpr int N = 4;
    __device__ static void load(const float* p, float (&f)[4]) {
        const uint4 u = ld_nc_v4(p);
        f[0] = __uint_as_float(u.x); f[1] = __uint_as_float(u.y); f[2] = __uint_as_float(u.z); f[3] = __uint_as_float(u.w);
    }
};

template <typename T>
__device__ __forceinline__ MS row_logsumexp2(const T* row, int V, bool vec_ok, float* smem /* >= 2*8 floats */) {
    constexpr int N = Vec<T>::N;
    MS acc{-INFINITY, 0.f};
    const int nvec = vec_ok ? V / N : 0;  // unaligned rows fall back to scalar loads
    // two independent 16-byte loads in flight per thread per iteration
    int i = threadIdx.x;
    for (; i + LOGPS_THREADS < nvec; i += 2 * LOGPS_THREADS) {
        float a[N], b[N];
        Vec<T>::load(row + (size_t)i * N, a);
        Vec<T>::load(row + (size_t)(i + LOGPS_THREADS) * N, b);
        float cm = a[0];
#pragma unroll
        for (int j = 1; j < N; ++j) cm = fmaxf(cm, a[j]);
#pragma unroll
        for (int j = 0; j < N; ++j) cm = fmaxf(cm, b[j]);
        cm *= LOG2E;
        if (cm > acc.m) { acc.s *= exp2f(acc.m - cm); acc.m = cm; }
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < N; ++j) s += exp2f(fmaf(a[j], LOG2E, -acc.m)) + exp2f(fmaf(b[j], LOG2E, -acc.m));
        acc.s += s;
    }
    for (; i < nvec; i += LOGPS_THREADS) {
        float a[N];
        Vec<T>::load(row + (size_t)i * N, a);
        float cm = a[0];
#pragma unroll
        for (int j = 1; j < N; ++j) cm = fmaxf(cm, a[j]);
        cm *= LOG2E;
        if (cm > acc.m) { acc.s *= exp2f(acc.m - cm); acc.m = cm; }
#pragma unroll
        for (int j = 0; j < N; ++j) acc.s += exp2f(fmaf(a[j], LOG2E, -acc.m));
    }
    for (int k = nvec * N + threadIdx.x; k < V; k += LOGPS_THREADS) {  // ragged tail (V % N)
        const float x = (float)row[k] * LOG2E;
        if (x > acc.m) { acc.s *= exp2f(acc.m - x); acc.m = x; }
        acc.s += exp2f(x - acc.m);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        MS other{__shfl_xor_sync(0xffffffffu, acc.m, o), __shfl_xor_sync(0xffffffffu, acc.s, o)};
        acc = ms_combine(acc, other);
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { smem[w] = acc.m; smem[8 + w] = acc.s; }
    __syncthreads();
    MS r{smem[0], smem[8]};
#pragma unroll
    for (int k = 1; k < LOGPS_THREADS / 32; ++k) r = ms_combine(r, MS{smem[k], smem[8 + k]});
    return r;
}

template <typename T>
__global__ void __launch_bounds__(LOGPS_THREADS)
logps_fwd_kernel(const T* __restrict__ logits, long long ld, const int64_t* __restrict__ target,
                 const uint8_t* __restrict__ weight, int V, int vec_ok, float* __restrict__ per_token, float* __restrict__ lse_out) {
    __shared__ float red[16];
    const int r = blockIdx.x;
    const int64_t t = target[r];
    const bool skip = t < 0 || (weight != nullptr && weight[r] == 0);
    if (skip) {  // row is never read
        if (threadIdx.x == 0) { per_token[r] = 0.f; if (lse_out) lse_out[r] = 0.f; }
        return;
    }
    const T* row = logits + (size_t)r * ld;
    const MS ms = row_logsumexp2<T>(row, V, vec_ok != 0, red);
    if (threadIdx.x == 0) {
        const float lse = (ms.m + log2f(ms.s)) * LN2;
        per_token[r] = (float)row[t] - lse;
        if (lse_out) lse_out[r] = lse;
    }
}

// one CTA per sequence: deterministic tree sum of that sequence's per-token log-probs
__global__ void __launch_bounds__(256)
logps_seq_reduce_kernel(const float* __restrict__ per_token, const int64_t* __restrict__ target,
                        const uint8_t* __restrict__ weight, int rows_per_seq, int average, float* __restrict__ logps) {
    __shared__ float ssum[8];
    __shared__ float scnt[8];
    const int s = blockIdx.x;
    float sum = 0.f, cnt = 0.f;
    for (int i = threadIdx.x; i < rows_per_seq; i += blockDim.x) {
        const int r = s * rows_per_seq + i;
        const bool on = target[r] >= 0 && (weight == nullptr || weight[r] != 0);
        if (on) { sum += per_token[r]; cnt += 1.f; }
    }
    sum = warp_sum(sum);
    cnt = warp_sum(cnt);
    if ((threadIdx.x & 31) == 0) { ssum[threadIdx.x >> 5] = sum; scnt[threadIdx.x >> 5] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, c = 0.f;
        for (int k = 0; k < 8; ++k) { a += ssum[k]; c += scnt[k]; }
        logps[s] = average ? a / c : a;  // 0/0 -> NaN exactly like the reference's sum/0 division
    }
}

template <typename T>
__global__ void __launch_bounds__(LOGPS_THREADS)
logps_bwd_kernel(const T* __restrict__ logits, long long ld, const int64_t* __restrict__ target,
                 const uint8_t* __restrict__ weight, const float* __restrict__ lse, const float* __restrict__ grad_logps,
                 const float* __restrict__ inv_count, int rows_per_seq, int V, int vec_ok,
                 __nv_bfloat16* __restrict__ dlogits, long long ldd) {
    const int r = blockIdx.x;
    const int64_t t = target[r];
    const bool skip = t < 0 || (weight != nullptr && weight[r] == 0);
    __nv_bfloat16* drow = dlogits + (size_t)r * ldd;
    constexpr int N = Vec<T>::N;
    const int nvec8 = vec_ok ? V / 8 : 0;
    if (skip) {
        const uint4 z = make_uint4(0, 0, 0, 0);
        for (int i = threadIdx.x; i < nvec8; i += LOGPS_THREADS) st_na_v4(drow + (size_t)i * 8, z);
        for (int k = nvec8 * 8 + threadIdx.x; k < V; k += LOGPS_THREADS) drow[k] = __float2bfloat16(0.f);
        return;
    }
    const int seq = r / rows_per_seq;
    float g = grad_logps[seq];
    if (inv_count) g *= inv_count[seq];
    const float l2 = lse[r] * LOG2E;
    const T* row = logits + (size_t)r * ld;
    // 8 outputs (one 16-byte bf16 store) per step; inputs are 1 (bf16) or 2 (f32) 16-byte loads
    for (int i = threadIdx.x; i < nvec8; i += LOGPS_THREADS) {
        float x[8];
        if constexpr (N == 8) {
            Vec<T>::load(row + (size_t)i * 8, x);
        } else {
            float a[4], b[4];
            Vec<T>::load(row + (size_t)i * 8, a);
            Vec<T>::load(row + (size_t)i * 8 + 4, b);
#pragma unroll
            for (int j = 0; j < 4; ++j) { x[j] = a[j]; x[4 + j] = b[j]; }
        }
        float d[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] = -g * exp2f(fmaf(x[j], LOG2E, -l2));
        const int base = i * 8;
        if (t >= base && t < base + 8) d[t - base] += g;
        uint4 o;
        o.x = pack_bf16x2(d[0], d[1]); o.y = pack_bf16x2(d[2], d[3]);
        o.z = pack_bf16x2(d[4], d[5]); o.w = pack_bf16x2(d[6], d[7]);
        st_na_v4(drow + (size_t)base, o);
    }
    for (int k = nvec8 * 8 + threadIdx.x; k < V; k += LOGPS_THREADS) {
        float d = -g * exp2f(fmaf((float)row[k], LOG2E, -l2));
        if (k == t) d += g;
        drow[k] = __float2bfloat16(d);
    }
}

__global__ void seq_inv_count_kernel(const int64_t* target, const uint8_t* weight, int rows_per_seq, float* inv_count) {
    __shared__ float scnt[8];
    const int s = blockIdx.x;
    float cnt = 0.f;
    for (int i = threadIdx.x; i < rows_per_seq; i += blockDim.x) {
        const int r = s * rows_per_seq + i;
        if (target[r] >= 0 && (weight == nullptr || weight[r] != 0)) cnt += 1.f;
    }
    cnt = warp_sum(cnt);
    if ((threadIdx.x & 31) == 0) scnt[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        float c = 0.f;
        for (int k = 0; k < 8; ++k) c += scnt[k];
        inv_count[s] = 1.f / c;
    }
}

// ------------------------------------------------------------------------------------------
// preference loss: one warp-sized problem (n_pairs is the per-GPU batch, 4..64); single CTA.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float logsigmoid(float x) { return fminf(x, 0.f) - log1pf(expf(-fabsf(x))); }
__device__ __forceinline__ float sigmoidf(float x) { return 1.f / (1.f + expf(-x)); }

__device__ float block_sum(float v, float* sm) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
    for (int k = 0; k < (blockDim.x >> 5); ++k) t += sm[k];
    return t;
}

__global__ void __launch_bounds__(256)
dpo_loss_kernel(const float* __restrict__ pol, const float* __restrict__ ref, int n, float beta, float ls, int loss_type,
                int reference_free, float loss_scale, float* __restrict__ losses, float* __restrict__ cr_out,
                float* __restrict__ rr_out, float* __restrict__ stats, float* __restrict__ grad) {
    __shared__ float sm[8];
    const float inv_n = 1.f / (float)n;
    float loss_sum = 0.f, acc_sum = 0.f, cr_sum = 0.f, rr_sum = 0.f;
    // kto_pair: batch-mean KL baselines (trainer.py:271-272) -- local batch, no detach in the reference
    float cKL_raw = 0.f, rKL_raw = 0.f;
    if (loss_type == VLB200_LOSS_KTO_PAIR) {
        float a = 0.f, b = 0.f;
        for (int i = threadIdx.x; i < n; i += blockDim.x) { a += pol[i] - ref[i]; b += pol[n + i] - ref[n + i]; }
        cKL_raw = block_sum(a, sm) * inv_n;
        rKL_raw = block_sum(b, sm) * inv_n;
    }
    const float cKL = fmaxf(cKL_raw, 0.f), rKL = fmaxf(rKL_raw, 0.f);
    float sa = 0.f, sz = 0.f;  // kto: sums of sigmoid' terms for the gradient through the KL baselines
    const int n_losses = loss_type == VLB200_LOSS_KTO_PAIR ? 2 * n : n;
    const float gscale = loss_scale / (float)n_losses;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float pc = pol[i], pr = pol[n + i], rc = ref[i], rr = ref[n + i];
        const float pi_lr = pc - pr;
        const float ref_lr = reference_free ? 0.f : rc - rr;
        const float d = pi_lr - ref_lr;
        float gc = 0.f, gr = 0.f;
        if (loss_type == VLB200_LOSS_SIGMOID || loss_type == VLB200_LOSS_DDPO) {
            const float l = -logsigmoid(beta * d) * (1.f - ls) - logsigmoid(-beta * d) * ls;
            losses[i] = l; loss_sum += l;
            const float dl = -beta * (1.f - ls) * sigmoidf(-beta * d) + beta * ls * sigmoidf(beta * d);
            gc = dl; gr = -dl;
        } else if (loss_type == VLB200_LOSS_HINGE) {
            const float l = fmaxf(1.f - beta * d, 0.f);
            losses[i] = l; loss_sum += l;
            const float dl = (1.f - beta * d) > 0.f ? -beta : 0.f;
            gc = dl; gr = -dl;
        } else if (loss_type == VLB200_LOSS_IPO) {
            const float e = d - 1.f / (2.f * beta);
            losses[i] = e * e; loss_sum += e * e;
            gc = 2.f * e; gr = -2.f * e;
        } else {  // kto_pair
            const float a = beta * ((pc - rc) - rKL);
            const float z = beta * (cKL - (pr - rr));
            const float s_a = sigmoidf(a), s_z = sigmoidf(z);
            losses[i] = 1.f - s_a; losses[n + i] = 1.f - s_z;
            loss_sum += (1.f - s_a) + (1.f - s_z);
            const float da = s_a * (1.f - s_a), dz = s_z * (1.f - s_z);
            sa += da; sz += dz;
            gc = -beta * da; gr = beta * dz;
        }
        const float cr = beta * (pc - rc), rw = beta * (pr - rr);
        cr_out[i] = cr; rr_out[i] = rw;
        cr_sum += cr; rr_sum += rw; acc_sum += cr > rw ? 1.f : 0.f;
        if (grad) { grad[i] = gc * gscale; grad[n + i] = gr * gscale; }
    }
    loss_sum = block_sum(loss_sum, sm);
    acc_sum = block_sum(acc_sum, sm);
    cr_sum = block_sum(cr_sum, sm);
    rr_sum = block_sum(rr_sum, sm);
    if (loss_type == VLB200_LOSS_KTO_PAIR && grad) {
        sa = block_sum(sa, sm);
        sz = block_sum(sz, sm);
        __syncthreads();
        // d losses_r / d cKL = -beta*sz_i ; d losses_c / d rKL = +beta*sa_i ; clamp gates the flow
        const float gc_kl = cKL_raw > 0.f ? -beta * sz * inv_n : 0.f;
        const float gr_kl = rKL_raw > 0.f ? beta * sa * inv_n : 0.f;
        for (int i = threadIdx.x; i < n; i += blockDim.x) { grad[i] += gc_kl * gscale; grad[n + i] += gr_kl * gscale; }
    }
    if (threadIdx.x == 0 && stats) {
        stats[0] = loss_sum / (float)n_losses;
        stats[1] = acc_sum * inv_n;
        stats[2] = cr_sum * inv_n;
        stats[3] = rr_sum * inv_n;
        stats[4] = (cr_sum - rr_sum) * inv_n;
        stats[5] = (float)n_losses;
    }
}

}  // namespace vlb

using namespace vlb;

extern "C" int vlb200_logps_fwd(const void* logits, int logits_dtype, int64_t ld_logits, const int64_t* target,
                                const uint8_t* weight, int rows, int rows_per_seq, int n_seq, int V, int average_log_prob,
                                float* per_token_logp, float* lse, float* logps, void* stream) {
    VLB_REQUIRE(logits && target && per_token_logp && logps, "logps_fwd: null pointer");
    VLB_REQUIRE(rows == rows_per_seq * n_seq && rows > 0,
                "Logits (batch and sequence length dim) and labels must have the same shape.");
    VLB_REQUIRE(V > 0 && ld_logits >= V, "logps_fwd: bad V/ld");
    const int align = logits_dtype == VLB200_BF16 ? 8 : 4;
    const int vec_ok = ld_logits % align == 0 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0;
    cudaStream_t s = as_stream(stream);
    if (logits_dtype == VLB200_BF16)
        logps_fwd_kernel<__nv_bfloat16><<<rows, LOGPS_THREADS, 0, s>>>((const __nv_bfloat16*)logits, ld_logits, target,
                                                                      weight, V, vec_ok, per_token_logp, lse);
    else if (logits_dtype == VLB200_F32)
        logps_fwd_kernel<float><<<rows, LOGPS_THREADS, 0, s>>>((const float*)logits, ld_logits, target, weight, V,
                                                              vec_ok, per_token_logp, lse);
    else
        VLB_REQUIRE(false, "logps_fwd: bad dtype %d", logits_dtype);
    VLB_LAUNCH_CHECK();
    logps_seq_reduce_kernel<<<n_seq, 256, 0, s>>>(per_token_logp, target, weight, rows_per_seq, average_log_prob, logps);
    count_launch(2);
    VLB_LAUNCH_CHECK();
    return VLB200_OK;
}

extern "C" int vlb200_logps_bwd(const void* logits, int logits_dtype, int64_t ld_logits, const int64_t* target,
                                const uint8_t* weight, const float* lse, const float* grad_logps, int rows,
                                int rows_per_seq, int n_seq, int V, int average_log_prob, void* dlogits,
                                int64_t ld_dlogits, void* stream) {
    VLB_REQUIRE(logits && target && lse && grad_logps && dlogits, "logps_bwd: null pointer");
    VLB_REQUIRE(rows == rows_per_seq * n_seq && rows > 0, "logps_bwd: rows != rows_per_seq * n_seq");
    const int in_align = logits_dtype == VLB200_BF16 ? 8 : 4;
    const int vec_ok = ld_dlogits % 8 == 0 && (reinterpret_cast<uintptr_t>(dlogits) & 15) == 0 &&
                       ld_logits % in_align == 0 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0;
    cudaStream_t s = as_stream(stream);
    float* inv_count = nullptr;
    if (average_log_prob) {
        // scratch lives in the (already consumed) lse tail is not available: use a tiny stream-ordered allocation
        VLB_CHECK_CUDA(cudaMallocAsync((void**)&inv_count, sizeof(float) * n_seq, s));
        seq_inv_count_kernel<<<n_seq, 256, 0, s>>>(target, weight, rows_per_seq, inv_count);
        count_launch();
    }
    if (logits_dtype == VLB200_BF16)
        logps_bwd_kernel<__nv_bfloat16><<<rows, LOGPS_THREADS, 0, s>>>(
            (const __nv_bfloat16*)logits, ld_logits, target, weight, lse, grad_logps, inv_count, rows_per_seq, V, vec_ok,
            (__nv_bfloat16*)dlogits, ld_dlogits);
    else if (logits_dtype == VLB200_F32)
        logps_bwd_kernel<float><<<rows, LOGPS_THREADS, 0, s>>>((const float*)logits, ld_logits, target, weight, lse,
                                                              grad_logps, inv_count, rows_per_seq, V, vec_ok,
                                                              (__nv_bfloat16*)dlogits, ld_dlogits);
    else
        VLB_REQUIRE(false, "logps_bwd: bad dtype %d", logits_dtype);
    count_launch();
    VLB_LAUNCH_CHECK();
    if (inv_count) VLB_CHECK_CUDA(cudaFreeAsync(inv_count, s));
    return VLB200_OK;
}

extern "C" int vlb200_dpo_loss(const float* policy_logps, const float* ref_logps, int n_pairs, float beta,
                               float label_smoothing, int loss_type, int reference_free, float loss_scale, float* losses,
                               float* chosen_rewards, float* rejected_rewards, float* stats, float* grad_policy_logps,
                               void* stream) {
    VLB_REQUIRE(policy_logps && ref_logps && losses && chosen_rewards && rejected_rewards, "dpo_loss: null pointer");
    VLB_REQUIRE(n_pairs > 0, "dpo_loss: n_pairs must be positive");
    VLB_REQUIRE(loss_type >= VLB200_LOSS_SIGMOID && loss_type <= VLB200_LOSS_DDPO,
                "Unknown loss type: %d. Should be one of ['sigmoid', 'hinge', 'ipo', 'kto_pair']", loss_type);
    dpo_loss_kernel<<<1, 256, 0, as_stream(stream)>>>(policy_logps, ref_logps, n_pairs, beta, label_smoothing, loss_type,
                                                      reference_free, loss_scale, losses, chosen_rewards,
                                                      rejected_rewards, stats, grad_policy_logps);
    count_launch();
    VLB_LAUNCH_CHECK();
    return VLB200_OK;
}
