// Image side of the reference's DPO collator on the GPU: the CLIP preprocessing that
// `LlavaDPODataCollatorWithPadding.__call__` (models/Llava/__init__.py:435-443) runs on the CPU through
// transformers' CLIPImageProcessor -> Pillow `Image.resize(BICUBIC)` + numpy crop / rescale / normalize.
//
// Bit-exact with Pillow's 8-bit resampler (src/libImaging/Resample.c): the host builds the fixed-point (22 fractional
// bits) coefficient tables exactly as precompute_coeffs + normalize_coeffs_8bpc do (vlrlhf_b200/preprocess.py), and
// the two passes below reproduce ImagingResampleHorizontal_8bpc / ImagingResampleVertical_8bpc: int32 accumulate
// starting at 1<<21, arithmetic >>22, clip to [0,255], the horizontal result stored as uint8 before the vertical
// pass.  Only the pixels that survive the center crop are computed.  The vertical pass is fused with the crop, the
// float64 multiply by 1/255 stored as float32 (image_transforms.rescale), the float32 (x-mean)/std
// (image_transforms.normalize) and the HWC -> CHW transpose.
//
// HBM-bound integer/byte work; one image is ~1 MB in and 1.35 MB out, so the grid is sized by pixels.
#include "common.cuh"

namespace vlb {
namespace {

constexpr int PRECISION_BITS = 32 - 8 - 2;

__device__ __forceinline__ uint8_t clip8(int v) {
    v >>= PRECISION_BITS;  // arithmetic shift, as the C code's signed >>
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// tmp[r, x, c] = horizontal pass of input row (row0 + r) at resized column (left + x);  r < rows, x < crop_w
__global__ void resample_h_u8_kernel(const uint8_t* __restrict__ img, int W, const int* __restrict__ kh,
                                     const int* __restrict__ bh, int ksize, int row0, int rows, int left, int crop_w,
                                     int new_w, uint8_t* __restrict__ tmp) {
    const size_t total = (size_t)rows * crop_w;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % crop_w);
        const int r = (int)(i / crop_w);
        const int xx = left + x;
        if (xx < 0 || xx >= new_w) continue;  // window column outside the resized image: zero padding (vertical pass)
        const int xmin = bh[2 * xx], n = bh[2 * xx + 1];
        const int* k = kh + (size_t)xx * ksize;
        const uint8_t* p = img + ((size_t)(row0 + r) * W + xmin) * 3;
        int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
        for (int t = 0; t < n; ++t) {
            const int c = k[t];
            s0 += p[3 * t] * c;
            s1 += p[3 * t + 1] * c;
            s2 += p[3 * t + 2] * c;
        }
        uint8_t* o = tmp + i * 3;
        o[0] = clip8(s0); o[1] = clip8(s1); o[2] = clip8(s2);
    }
}

template <typename OT>
__device__ __forceinline__ OT to_out(float v);
template <> __device__ __forceinline__ float to_out<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 to_out<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// out[c, y, x] = normalize(rescale(vertical pass at resized row (top + y)))
template <typename OT>
__global__ void resample_v_norm_kernel(const uint8_t* __restrict__ tmp, const int* __restrict__ kv,
                                       const int* __restrict__ bv, int ksize, int row0, int top, int left, int new_h,
                                       int new_w, int crop_h, int crop_w, double rescale, float m0, float m1, float m2, float d0, float d1, float d2,
                                       OT* __restrict__ out) {
    const size_t total = (size_t)crop_h * crop_w;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % crop_w);
        const int y = (int)(i / crop_w);
        const int yy = top + y, xx = left + x;
        uint8_t u0 = 0, u1 = 0, u2 = 0;  // outside the resized image: the zero padding of the uint8 canvas
        if (yy >= 0 && yy < new_h && xx >= 0 && xx < new_w) {
            const int ymin = bv[2 * yy], n = bv[2 * yy + 1];
            const int* k = kv + (size_t)yy * ksize;
            const uint8_t* p = tmp + ((size_t)(ymin - row0) * crop_w + x) * 3;
            int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
            for (int t = 0; t < n; ++t) {
                const int c = k[t];
                const uint8_t* q = p + (size_t)t * crop_w * 3;
                s0 += q[0] * c;
                s1 += q[1] * c;
                s2 += q[2] * c;
            }
            u0 = clip8(s0); u1 = clip8(s1); u2 = clip8(s2);
        }
        // float64 multiply, float32 store; then float32 subtract and IEEE divide (no fast-math in this file)
        const float f0 = (float)((double)u0 * rescale);
        const float f1 = (float)((double)u1 * rescale);
        const float f2 = (float)((double)u2 * rescale);
        out[i] = to_out<OT>(__fdiv_rn(__fsub_rn(f0, m0), d0));
        out[total + i] = to_out<OT>(__fdiv_rn(__fsub_rn(f1, m1), d1));
        out[2 * total + i] = to_out<OT>(__fdiv_rn(__fsub_rn(f2, m2), d2));
    }
}

inline int grid_px(size_t n) {
    size_t b = (n + 255) / 256;
    const size_t cap = (size_t)num_sms() * 8;
    return (int)(b < cap ? (b ? b : 1) : cap);
}

}  // namespace
}  // namespace vlb

using namespace vlb;

extern "C" size_t vlb200_clip_preprocess_workspace_bytes(int in_h, int crop_w) { return (size_t)in_h * crop_w * 3; }

extern "C" int vlb200_clip_preprocess_u8(const uint8_t* image, int in_h, int in_w, const int* coef_h, const int* bounds_h,
                                         int ksize_h, const int* coef_v, const int* bounds_v, int ksize_v, int new_h,
                                         int new_w, int top, int left, int crop_h, int crop_w, int row0, int rows,
                                         uint8_t* workspace, size_t workspace_bytes, double rescale,
                                         const float* mean_std_host, void* out, int out_dtype, void* stream) {
    VLB_REQUIRE(image && coef_h && bounds_h && coef_v && bounds_v && workspace && mean_std_host && out,
                "clip_preprocess: null pointer");
    VLB_REQUIRE(in_h > 0 && in_w > 0 && ksize_h > 0 && ksize_v > 0, "clip_preprocess: bad sizes");
    VLB_REQUIRE(new_h > 0 && new_w > 0 && crop_h > 0 && crop_w > 0, "clip_preprocess: bad output geometry");
    VLB_REQUIRE(row0 >= 0 && rows >= 0 && row0 + rows <= in_h, "clip_preprocess: bad input row range");
    VLB_REQUIRE(workspace_bytes >= (size_t)rows * crop_w * 3, "clip_preprocess: workspace too small");
    VLB_REQUIRE(out_dtype == VLB200_F32 || out_dtype == VLB200_BF16, "clip_preprocess: bad out dtype");
    cudaStream_t s = as_stream(stream);
    if (rows > 0) {  // rows == 0: the window lies entirely in the padding
        resample_h_u8_kernel<<<grid_px((size_t)rows * crop_w), 256, 0, s>>>(image, in_w, coef_h, bounds_h, ksize_h, row0, rows,
                                                                           left, crop_w, new_w, workspace);
        VLB_LAUNCH_CHECK();
    }
    const float* m = mean_std_host;
    if (out_dtype == VLB200_F32)
        resample_v_norm_kernel<float><<<grid_px((size_t)crop_h * crop_w), 256, 0, s>>>(
            workspace, coef_v, bounds_v, ksize_v, row0, top, left, new_h, new_w, crop_h, crop_w, rescale, m[0], m[1], m[2], m[3],
            m[4], m[5], (float*)out);
    else
        resample_v_norm_kernel<__nv_bfloat16><<<grid_px((size_t)crop_h * crop_w), 256, 0, s>>>(
            workspace, coef_v, bounds_v, ksize_v, row0, top, left, new_h, new_w, crop_h, crop_w, rescale, m[0], m[1], m[2], m[3],
            m[4], m[5], (__nv_bfloat16*)out);
    count_launch(2);
    VLB_LAUNCH_CHECK();
    return VLB200_OK;
}
