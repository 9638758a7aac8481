// RMSNorm (Llama, modeling_llama.py:53-67) forward/backward and LayerNorm (CLIP, modeling_clip.py) forward.
// HBM-bound row kernels: 16-byte loads, fp32 statistics, one CTA (128 threads) per row in forward;
// backward is a persistent grid that also produces the weight gradient (two-stage, deterministic).
#include "common.cuh"

namespace vlb {

constexpr int NORM_THREADS = 128;
constexpr int NORM_MAX_VEC = 8;  // supports cols <= 128 * 8 * 8 = 8192

__device__ __forceinline__ float block_sum_128(float v, float* sm) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    const float t = sm[0] + sm[1] + sm[2] + sm[3];
    __syncthreads();
    return t;
}

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 t = unpack_bf16x2(w[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    uint4 o;
    o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]);
    o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
    return o;
}

// 8 consecutive elements as floats from a bf16 (16 B) or fp32 (32 B) row
template <typename XT>
__device__ __forceinline__ void load8(const XT* p, float (&f)[8]);
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&f)[8]) {
    unpack8(*reinterpret_cast<const uint4*>(p), f);
}
template <>
__device__ __forceinline__ void load8<float>(const float* p, float (&f)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// y = w * (x * rsqrt(mean(x^2) + eps))            (fp32 math, one rounding at the end; x may be bf16 or fp32)
template <typename XT>
__global__ void __launch_bounds__(NORM_THREADS)
rmsnorm_fwd_kernel(const XT* __restrict__ x, long long ldx, const __nv_bfloat16* __restrict__ w,
                   __nv_bfloat16* __restrict__ y, long long ldy, float* __restrict__ rstd_out, int cols, float eps) {
    __shared__ float sm[4];
    const int row = blockIdx.x;
    const int nvec = cols >> 3;
    const XT* xr = x + (size_t)row * ldx;
    float xf[NORM_MAX_VEC][8];
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < NORM_MAX_VEC; ++k) {
        const int i = threadIdx.x + k * NORM_THREADS;
        if (i < nvec) {
            load8<XT>(xr + (size_t)i * 8, xf[k]);
#pragma unroll
            for (int j = 0; j < 8; ++j) ss += xf[k][j] * xf[k][j];
        }
    }
    ss = block_sum_128(ss, sm);
    const float rstd = rsqrtf(ss / (float)cols + eps);
    if (threadIdx.x == 0 && rstd_out) rstd_out[row] = rstd;
    __nv_bfloat16* yr = y + (size_t)row * ldy;
#pragma unroll
    for (int k = 0; k < NORM_MAX_VEC; ++k) {
        const int i = threadIdx.x + k * NORM_THREADS;
        if (i < nvec) {
            float f[8], g[8];
            unpack8(*reinterpret_cast<const uint4*>(w + (size_t)i * 8), g);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = g[j] * (xf[k][j] * rstd);
            *reinterpret_cast<uint4*>(yr + (size_t)i * 8) = pack8(f);
        }
    }
}

// dx = rstd * (w*dy - xhat * mean(w*dy*xhat)) (+ dres);   dw_partial[cta] += dy * xhat
template <int VPT, typename XT>
__global__ void __launch_bounds__(NORM_THREADS)
rmsnorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const XT* __restrict__ x,
                   const __nv_bfloat16* __restrict__ w, const float* __restrict__ rstd_in,
                   const __nv_bfloat16* __restrict__ dres, __nv_bfloat16* __restrict__ dx, float* __restrict__ dw_partial,
                   int rows, int cols) {
    __shared__ float sm[4];
    const int nvec = cols >> 3;
    float dw[VPT][8];
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
#pragma unroll
        for (int j = 0; j < 8; ++j) dw[k][j] = 0.f;
    }
    for (int row = blockIdx.x; row < rows; row += gridDim.x) {
        const float rstd = rstd_in[row];
        const size_t off = (size_t)row * cols;
        // every global load of the row is issued before the block reduction: x, dy and the residual gradient are all in
        // flight together (the weight vector is re-read through L1 instead of living in 32 registers per thread)
        float xs[VPT][8];
        uint4 gv[VPT], rv[VPT];
#pragma unroll
        for (int k = 0; k < VPT; ++k) {
            const int i = threadIdx.x + k * NORM_THREADS;
            if (i < nvec) {
                load8<XT>(x + off + (size_t)i * 8, xs[k]);
                gv[k] = ld_nc_v4(dy + off + (size_t)i * 8);
                if (dres) rv[k] = ld_nc_v4(dres + off + (size_t)i * 8);
            }
        }
        float dot = 0.f;
#pragma unroll
        for (int k = 0; k < VPT; ++k) {
            const int i = threadIdx.x + k * NORM_THREADS;
            if (i < nvec) {
                float gf[8], wv[8];
                unpack8(gv[k], gf);
                unpack8(__ldg(reinterpret_cast<const uint4*>(w + (size_t)i * 8)), wv);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float xh = xs[k][j] * rstd;
                    dot += wv[j] * gf[j] * xh;
                    dw[k][j] += gf[j] * xh;
                }
            }
        }
        dot = block_sum_128(dot, sm) / (float)cols;
#pragma unroll
        for (int k = 0; k < VPT; ++k) {
            const int i = threadIdx.x + k * NORM_THREADS;
            if (i < nvec) {
                float gf[8], wv[8], o[8];
                unpack8(gv[k], gf);
                unpack8(__ldg(reinterpret_cast<const uint4*>(w + (size_t)i * 8)), wv);
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = rstd * (wv[j] * gf[j] - xs[k][j] * rstd * dot);
                if (dres) {
                    float rf[8];
                    unpack8(rv[k], rf);
#pragma unroll
                    for (int j = 0; j < 8; ++j) o[j] += rf[j];
                }
                *reinterpret_cast<uint4*>(dx + off + (size_t)i * 8) = pack8(o);
            }
        }
    }
    float* out = dw_partial + (size_t)blockIdx.x * cols;
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
        const int i = threadIdx.x + k * NORM_THREADS;
        if (i < nvec) {
#pragma unroll
            for (int j = 0; j < 8; ++j) out[(size_t)i * 8 + j] = dw[k][j];
        }
    }
}

// out[c] = sum_p partial[p, c]: 32 columns per CTA, the partials split over 8 warps (coalesced 128-byte rows, several
// loads in flight per lane), then a shared-memory reduction over the warps.  Deterministic: fixed summation order.
__device__ __forceinline__ float colsum_partials_body(const float* __restrict__ partial, int nparts, int cols, int& c) {
    __shared__ float sm[8][33];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    c = blockIdx.x * 32 + lane;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (c < cols) {
        int p = wid;
        for (; p + 24 < nparts; p += 32) {
            s0 += partial[(size_t)p * cols + c];
            s1 += partial[(size_t)(p + 8) * cols + c];
            s2 += partial[(size_t)(p + 16) * cols + c];
            s3 += partial[(size_t)(p + 24) * cols + c];
        }
        for (; p < nparts; p += 8) s0 += partial[(size_t)p * cols + c];
    }
    sm[wid][lane] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    float s = 0.f;
    if (wid == 0) {
#pragma unroll
        for (int w = 0; w < 8; ++w) s += sm[w][lane];
    }
    return s;
}
// out[c] (bf16) (+)= sum_p partial[p, c]
__global__ void __launch_bounds__(256) colsum_partials_kernel(const float* __restrict__ partial, int nparts, int cols,
                                                              __nv_bfloat16* __restrict__ out, int accumulate) {
    int c;
    float s = colsum_partials_body(partial, nparts, cols, c);
    if (threadIdx.x < 32 && c < cols) {
        if (accumulate) s += __bfloat162float(out[c]);
        out[c] = __float2bfloat16(s);
    }
}

// out[c] (f32) = sum_p partial[p, c]
__global__ void __launch_bounds__(256) colsum_partials_f32_kernel(const float* __restrict__ partial, int nparts, int cols,
                                                                  float* __restrict__ out) {
    int c;
    const float s = colsum_partials_body(partial, nparts, cols, c);
    if (threadIdx.x < 32 && c < cols) out[c] = s;
}
// out[0] = scale * sum_i a[i] * b[i]   (single CTA; n is a hidden size)
__global__ void __launch_bounds__(256) dot_f32_kernel(const float* __restrict__ a, const float* __restrict__ b, int n, float scale,
                                                      float* __restrict__ out) {
    __shared__ float sm[8];
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += a[i] * b[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int k = 0; k < 8; ++k) t += sm[k];
        out[0] = t * scale;
    }
}

// column sums of a bf16 matrix (bias gradients): partial[cta, c] = sum over the CTA's rows
__global__ void __launch_bounds__(256)
colsum_rows_kernel(const __nv_bfloat16* __restrict__ a, long long lda, int rows, int cols, float* __restrict__ partial) {
    const int c = blockIdx.y * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    float s = 0.f;
    for (int r = blockIdx.x; r < rows; r += gridDim.x) s += __bfloat162float(a[(size_t)r * lda + c]);
    partial[(size_t)blockIdx.x * cols + c] = s;
}

// LayerNorm forward (CLIP): y = (x - mean) * rsqrt(var + eps) * w + b
template <typename XT>
__global__ void __launch_bounds__(NORM_THREADS)
layernorm_fwd_kernel(const XT* __restrict__ x, long long ldx, const __nv_bfloat16* __restrict__ w,
                     const __nv_bfloat16* __restrict__ b, __nv_bfloat16* __restrict__ y, long long ldy, int cols,
                     float eps) {
    __shared__ float sm[4];
    const int row = blockIdx.x;
    const int nvec = cols >> 3;
    const XT* xr = x + (size_t)row * ldx;
    float xf[NORM_MAX_VEC][8];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < NORM_MAX_VEC; ++k) {
        const int i = threadIdx.x + k * NORM_THREADS;
        if (i < nvec) {
            load8<XT>(xr + (size_t)i * 8, xf[k]);
#pragma unroll
            for (int j = 0; j < 8; ++j) s += xf[k][j];
        }
    }
    const float mean = block_sum_128(s, sm) / (float)cols;
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < NORM_MAX_VEC; ++k) {
        const int i = threadIdx.x + k * NORM_THREADS;
        if (i < nvec) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { const float d = xf[k][j] - mean; v += d * d; }
        }
    }
    const float rstd = rsqrtf(block_sum_128(v, sm) / (float)cols + eps);
    __nv_bfloat16* yr = y + (size_t)row * ldy;
#pragma unroll
    for (int k = 0; k < NORM_MAX_VEC; ++k) {
        const int i = threadIdx.x + k * NORM_THREADS;
        if (i < nvec) {
            float f[8], g[8], h[8];
            unpack8(*reinterpret_cast<const uint4*>(w + (size_t)i * 8), g);
            unpack8(*reinterpret_cast<const uint4*>(b + (size_t)i * 8), h);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = (xf[k][j] - mean) * rstd * g[j] + h[j];
            *reinterpret_cast<uint4*>(yr + (size_t)i * 8) = pack8(f);
        }
    }
}

}  // namespace vlb

using namespace vlb;

static int check_cols(int cols, const char* who) {
    VLB_REQUIRE(cols > 0 && cols % 8 == 0 && cols <= NORM_THREADS * NORM_MAX_VEC * 8, "%s: cols=%d must be a multiple of 8 and <= %d",
                who, cols, NORM_THREADS * NORM_MAX_VEC * 8);
    return VLB200_OK;
}

extern "C" int vlb200_rmsnorm_fwd(const void* x, int x_dtype, int64_t ldx, const void* w, void* y, int64_t ldy, float* rstd,
                                  int rows, int cols, float eps, void* stream) {
    VLB_REQUIRE(x && w && y, "rmsnorm_fwd: null pointer");
    if (int rc = check_cols(cols, "rmsnorm_fwd")) return rc;
    VLB_REQUIRE(ldx % 8 == 0 && ldy % 8 == 0, "rmsnorm_fwd: row strides must be multiples of 8");
    if (rows <= 0) return VLB200_OK;
    VLB_REQUIRE(x_dtype == VLB200_BF16 || x_dtype == VLB200_F32, "rmsnorm_fwd: bad x dtype");
    if (x_dtype == VLB200_F32)
        rmsnorm_fwd_kernel<float><<<rows, NORM_THREADS, 0, as_stream(stream)>>>((const float*)x, ldx, (const __nv_bfloat16*)w,
                                                                               (__nv_bfloat16*)y, ldy, rstd, cols, eps);
    else
        rmsnorm_fwd_kernel<__nv_bfloat16><<<rows, NORM_THREADS, 0, as_stream(stream)>>>(
            (const __nv_bfloat16*)x, ldx, (const __nv_bfloat16*)w, (__nv_bfloat16*)y, ldy, rstd, cols, eps);
    count_launch();
    VLB_LAUNCH_CHECK();
    return VLB200_OK;
}

extern "C" int vlb200_norm_bwd_workspace_floats(int cols) { return 8 * num_sms() * cols; }

extern "C" int vlb200_rmsnorm_bwd(const void* dy, const void* x, int x_dtype, const void* w, const float* rstd, const void* dres,
                                  void* dx, void* dw, int dw_accumulate, float* workspace, int rows, int cols,
                                  void* stream) {
    VLB_REQUIRE(dy && x && w && rstd && dx && dw && workspace, "rmsnorm_bwd: null pointer");
    if (int rc = check_cols(cols, "rmsnorm_bwd")) return rc;
    if (rows <= 0) return VLB200_OK;
    const int grid = rows < 8 * num_sms() ? rows : 8 * num_sms();
    cudaStream_t s = as_stream(stream);
    const int vpt = ((cols >> 3) + NORM_THREADS - 1) / NORM_THREADS;
    VLB_REQUIRE(x_dtype == VLB200_BF16 || x_dtype == VLB200_F32, "rmsnorm_bwd: bad x dtype");
#define VLB_RMS_BWD(V)                                                                                                     \
    do {                                                                                                                   \
        if (x_dtype == VLB200_F32)                                                                                         \
            rmsnorm_bwd_kernel<V, float><<<grid, NORM_THREADS, 0, s>>>((const __nv_bfloat16*)dy, (const float*)x,           \
                                                                       (const __nv_bfloat16*)w, rstd, (const __nv_bfloat16*)dres, \
                                                                       (__nv_bfloat16*)dx, workspace, rows, cols);        \
        else                                                                                                               \
            rmsnorm_bwd_kernel<V, __nv_bfloat16><<<grid, NORM_THREADS, 0, s>>>(                                            \
                (const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, (const __nv_bfloat16*)w, rstd, (const __nv_bfloat16*)dres, \
                (__nv_bfloat16*)dx, workspace, rows, cols);                                                               \
    } while (0)
    if (vpt <= 1) VLB_RMS_BWD(1);
    else if (vpt <= 2) VLB_RMS_BWD(2);
    else if (vpt <= 4) VLB_RMS_BWD(4);
    else VLB_RMS_BWD(8);
#undef VLB_RMS_BWD
    VLB_LAUNCH_CHECK();
    colsum_partials_kernel<<<(cols + 31) / 32, 256, 0, s>>>(workspace, grid, cols, (__nv_bfloat16*)dw, dw_accumulate);
    count_launch(2);
    VLB_LAUNCH_CHECK();
    return VLB200_OK;
}

extern "C" int vlb200_colsum(const void* a, int64_t lda, int rows, int cols, void* out, int accumulate, float* workspace,
                             void* stream) {
    VLB_REQUIRE(a && out && workspace, "colsum: null pointer");
    VLB_REQUIRE(cols <= NORM_THREADS * NORM_MAX_VEC * 8 * 4, "colsum: cols too large for the workspace contract");
    if (rows <= 0 || cols <= 0) return VLB200_OK;
    const int gx = rows < 2 * num_sms() ? rows : 2 * num_sms();
    dim3 grid(gx, (cols + 255) / 256);
    cudaStream_t s = as_stream(stream);
    colsum_rows_kernel<<<grid, 256, 0, s>>>((const __nv_bfloat16*)a, lda, rows, cols, workspace);
    VLB_LAUNCH_CHECK();
    colsum_partials_kernel<<<(cols + 31) / 32, 256, 0, s>>>(workspace, gx, cols, (__nv_bfloat16*)out, accumulate);
    count_launch(2);
    VLB_LAUNCH_CHECK();
    return VLB200_OK;
}

extern "C" int vlb200_layernorm_fwd(const void* x, int x_dtype, int64_t ldx, const void* w, const void* b, void* y, int64_t ldy,
                                    int rows, int cols, float eps, void* stream) {
    VLB_REQUIRE(x && w && b && y, "layernorm_fwd: null pointer");
    if (int rc = check_cols(cols, "layernorm_fwd")) return rc;
    if (rows <= 0) return VLB200_OK;
    VLB_REQUIRE(x_dtype == VLB200_BF16 || x_dtype == VLB200_F32, "layernorm_fwd: bad x dtype");
    if (x_dtype == VLB200_F32)
        layernorm_fwd_kernel<float><<<rows, NORM_THREADS, 0, as_stream(stream)>>>(
            (const float*)x, ldx, (const __nv_bfloat16*)w, (const __nv_bfloat16*)b, (__nv_bfloat16*)y, ldy, cols, eps);
    else
        layernorm_fwd_kernel<__nv_bfloat16><<<rows, NORM_THREADS, 0, as_stream(stream)>>>(
            (const __nv_bfloat16*)x, ldx, (const __nv_bfloat16*)w, (const __nv_bfloat16*)b, (__nv_bfloat16*)y, ldy, cols, eps);
    count_launch();
    VLB_LAUNCH_CHECK();
    return VLB200_OK;
}

extern "C" int vlb200_colsum_f32(const void* a, int64_t lda, int rows, int cols, float* out, float* workspace, void* stream) {
    VLB_REQUIRE(a && out && workspace, "colsum_f32: null pointer");
    if (rows <= 0 || cols <= 0) return VLB200_OK;
    const int gx = rows < 2 * num_sms() ? rows : 2 * num_sms();
    dim3 grid(gx, (cols + 255) / 256);
    cudaStream_t s = as_stream(stream);
    colsum_rows_kernel<<<grid, 256, 0, s>>>((const __nv_bfloat16*)a, lda, rows, cols, workspace);
    VLB_LAUNCH_CHECK();
    colsum_partials_f32_kernel<<<(cols + 31) / 32, 256, 0, s>>>(workspace, gx, cols, out);
    count_launch(2);
    VLB_LAUNCH_CHECK();
    return VLB200_OK;
}

extern "C" int vlb200_dot_f32(const float* a, const float* b, int n, float scale, float* out, void* stream) {
    VLB_REQUIRE(a && b && out && n > 0, "dot_f32: bad arguments");
    dot_f32_kernel<<<1, 256, 0, as_stream(stream)>>>(a, b, n, scale, out);
    count_launch();
    VLB_LAUNCH_CHECK();
    return VLB200_OK;
}
