// HBM-bound glue kernels of the DPO step: RoPE, SwiGLU, GELU, CLIP patch im2col, text/image merge
// (Llava/__init__.py:36-109), row gather/scatter, AdamW, gradient norm.  All use 16-byte accesses and
// grid-stride loops sized from the SM count.
#include <algorithm>
#include <climits>

#include "common.cuh"

namespace vlb {

__device__ __forceinline__ void unpack8e(const uint4& u, float (&f)[8]) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 t = unpack_bf16x2(w[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}
__device__ __forceinline__ uint4 pack8e(const float (&f)[8]) {
    uint4 o;
    o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]);
    o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
    return o;
}

static inline int grid_for(size_t work_items, int threads) {
    size_t b = (work_items + threads - 1) / threads;
    size_t cap = (size_t)num_sms() * 32;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// ---------------------------------------------------------------- RoPE (modeling_llama.py:146-168)
// rotate-half on the first n_rot_heads heads of every row of `qkv` (q heads then k heads), in place.
// cos/sin tables are [max_pos, dh/2] fp32 built on the host exactly like LlamaRotaryEmbedding.
__global__ void rope_kernel(__nv_bfloat16* __restrict__ qkv, long long ld, const int* __restrict__ pos,
                            const float* __restrict__ cos_t, const float* __restrict__ sin_t, int table_rows, int rows,
                            int n_rot_heads, int dh, float sign) {
    const int half = dh >> 1;
    const int chunks = half >> 3;  // 8 pairs per work item
    const size_t total = (size_t)rows * n_rot_heads * chunks;
    for (size_t w = blockIdx.x * (size_t)blockDim.x + threadIdx.x; w < total; w += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(w % chunks);
        const int h = (int)((w / chunks) % n_rot_heads);
        const int r = (int)(w / ((size_t)chunks * n_rot_heads));
        __nv_bfloat16* p = qkv + (size_t)r * ld + (size_t)h * dh + c * 8;
        const int ps = min(max(pos[r], 0), table_rows - 1);
        const float* cp = cos_t + (size_t)ps * half + c * 8;
        const float* sp = sin_t + (size_t)ps * half + c * 8;
        float x1[8], x2[8];
        unpack8e(*reinterpret_cast<const uint4*>(p), x1);
        unpack8e(*reinterpret_cast<const uint4*>(p + half), x2);
        float o1[8], o2[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float cs = cp[j], sn = sign * sp[j];
            o1[j] = x1[j] * cs - x2[j] * sn;
            o2[j] = x2[j] * cs + x1[j] * sn;
        }
        *reinterpret_cast<uint4*>(p) = pack8e(o1);
        *reinterpret_cast<uint4*>(p + half) = pack8e(o2);
    }
}

// ---------------------------------------------------------------- SwiGLU (modeling_llama.py:182-184)
__global__ void swiglu_fwd_kernel(const __nv_bfloat16* __restrict__ gu, long long ldgu, __nv_bfloat16* __restrict__ act,
                                  long long ldact, int rows, int ff) {
    const int chunks = ff >> 3;
    const size_t total = (size_t)rows * chunks;
    for (size_t w = blockIdx.x * (size_t)blockDim.x + threadIdx.x; w < total; w += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(w % chunks);
        const size_t r = w / chunks;
        float g[8], u[8], o[8];
        unpack8e(*reinterpret_cast<const uint4*>(gu + r * ldgu + c * 8), g);
        unpack8e(*reinterpret_cast<const uint4*>(gu + r * ldgu + ff + c * 8), u);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = g[j] / (1.f + __expf(-g[j])) * u[j];
        *reinterpret_cast<uint4*>(act + r * ldact + c * 8) = pack8e(o);
    }
}
__global__ void swiglu_bwd_kernel(const __nv_bfloat16* __restrict__ gu, long long ldgu,
                                  const __nv_bfloat16* __restrict__ dact, long long lddact,
                                  __nv_bfloat16* __restrict__ dgu, long long lddgu, int rows, int ff) {
    const int chunks = ff >> 3;
    const size_t total = (size_t)rows * chunks;
    for (size_t w = blockIdx.x * (size_t)blockDim.x + threadIdx.x; w < total; w += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(w % chunks);
        const size_t r = w / chunks;
        float g[8], u[8], d[8], dg[8], du[8];
        unpack8e(*reinterpret_cast<const uint4*>(gu + r * ldgu + c * 8), g);
        unpack8e(*reinterpret_cast<const uint4*>(gu + r * ldgu + ff + c * 8), u);
        unpack8e(*reinterpret_cast<const uint4*>(dact + r * lddact + c * 8), d);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float s = 1.f / (1.f + __expf(-g[j]));
            dg[j] = d[j] * u[j] * s * (1.f + g[j] * (1.f - s));
            du[j] = d[j] * g[j] * s;
        }
        *reinterpret_cast<uint4*>(dgu + r * lddgu + c * 8) = pack8e(dg);
        *reinterpret_cast<uint4*>(dgu + r * lddgu + ff + c * 8) = pack8e(du);
    }
}

// ---------------------------------------------------------------- GELU(erf) (projector, modeling_llava.py:87-107)
__global__ void gelu_fwd_kernel(const __nv_bfloat16* __restrict__ z, __nv_bfloat16* __restrict__ h, size_t n8) {
    for (size_t w = blockIdx.x * (size_t)blockDim.x + threadIdx.x; w < n8; w += (size_t)gridDim.x * blockDim.x) {
        float x[8];
        unpack8e(*reinterpret_cast<const uint4*>(z + w * 8), x);
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = 0.5f * x[j] * (1.f + erff(x[j] * 0.70710678118654752f));
        *reinterpret_cast<uint4*>(h + w * 8) = pack8e(x);
    }
}
__global__ void gelu_bwd_kernel(const __nv_bfloat16* __restrict__ z, const __nv_bfloat16* __restrict__ dh,
                                __nv_bfloat16* __restrict__ dz, size_t n8) {
    for (size_t w = blockIdx.x * (size_t)blockDim.x + threadIdx.x; w < n8; w += (size_t)gridDim.x * blockDim.x) {
        float x[8], d[8];
        unpack8e(*reinterpret_cast<const uint4*>(z + w * 8), x);
        unpack8e(*reinterpret_cast<const uint4*>(dh + w * 8), d);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float cdf = 0.5f * (1.f + erff(x[j] * 0.70710678118654752f));
            const float pdf = 0.3989422804014327f * __expf(-0.5f * x[j] * x[j]);
            d[j] *= cdf + x[j] * pdf;
        }
        *reinterpret_cast<uint4*>(dz + w * 8) = pack8e(d);
    }
}

// ---------------------------------------------------------------- CLIP patch im2col (modeling_clip.py:148-154,209)
// pixels [B,3,H,W] (f32 or bf16) -> patches [B*np, ldo] bf16, column = c*ps*ps + ky*ps + kx
template <typename T>
__global__ void im2col_kernel(const T* __restrict__ pix, __nv_bfloat16* __restrict__ out, long long ldo, int B, int H,
                              int W, int ps) {
    const int gw = W / ps, gh = H / ps;
    const int K = 3 * ps * ps;
    const size_t total = (size_t)B * gh * gw * K;
    for (size_t w = blockIdx.x * (size_t)blockDim.x + threadIdx.x; w < total; w += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(w % K);
        const size_t row = w / K;
        const int px = (int)(row % gw), py = (int)((row / gw) % gh), b = (int)(row / ((size_t)gw * gh));
        const int kx = k % ps, ky = (k / ps) % ps, c = k / (ps * ps);
        const float v = (float)pix[(((size_t)b * 3 + c) * H + (py * ps + ky)) * W + px * ps + kx];
        out[row * ldo + k] = __float2bfloat16(v);
    }
}
// x[b*(np+1), :] = cls + pos[0]
__global__ void cls_rows_kernel(__nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ cls,
                                const __nv_bfloat16* __restrict__ pos0, int B, int tokens_per_img, int d) {
    const int total = B * d;
    for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < total; w += gridDim.x * blockDim.x) {
        const int b = w / d, j = w % d;
        x[(size_t)b * tokens_per_img * d + j] = __float2bfloat16(__bfloat162float(cls[j]) + __bfloat162float(pos0[j]));
    }
}

// ---------------------------------------------------------------- strided row copy / gather / scatter (16 B)
// dst[g, r, :cols] = src[g*src_gs + (r + src_r0)*src_rs : ...]
__global__ void copy_rows_kernel(const __nv_bfloat16* __restrict__ src, long long src_gs, long long src_rs, int src_r0,
                                 __nv_bfloat16* __restrict__ dst, long long dst_gs, long long dst_rs, int groups,
                                 int rows_per_group, int cols) {
    const int chunks = cols >> 3;
    const size_t total = (size_t)groups * rows_per_group * chunks;
    for (size_t w = blockIdx.x * (size_t)blockDim.x + threadIdx.x; w < total; w += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(w % chunks);
        const int r = (int)((w / chunks) % rows_per_group);
        const int g = (int)(w / ((size_t)chunks * rows_per_group));
        *reinterpret_cast<uint4*>(dst + g * dst_gs + r * dst_rs + c * 8) =
            *reinterpret_cast<const uint4*>(src + g * src_gs + (r + src_r0) * src_rs + c * 8);
    }
}
// dst[i, :] = src[idx[i], :]   (idx < 0 -> zeros)
__global__ void gather_rows_kernel(const __nv_bfloat16* __restrict__ src, long long lds, const int* __restrict__ idx,
                                   __nv_bfloat16* __restrict__ dst, long long ldd, int n, int cols) {
    const int chunks = cols >> 3;
    const size_t total = (size_t)n * chunks;
    for (size_t w = blockIdx.x * (size_t)blockDim.x + threadIdx.x; w < total; w += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(w % chunks);
        const size_t i = w / chunks;
        const int s = idx[i];
        uint4 v = make_uint4(0, 0, 0, 0);
        if (s >= 0) v = *reinterpret_cast<const uint4*>(src + (size_t)s * lds + c * 8);
        *reinterpret_cast<uint4*>(dst + i * ldd + c * 8) = v;
    }
}
// dst[idx[i], :] = src[i, :]   (idx unique; idx < 0 skipped); dst must be pre-zeroed
__global__ void scatter_rows_kernel(const __nv_bfloat16* __restrict__ src, long long lds, const int* __restrict__ idx,
                                    __nv_bfloat16* __restrict__ dst, long long ldd, int n, int cols) {
    const int chunks = cols >> 3;
    const size_t total = (size_t)n * chunks;
    for (size_t w = blockIdx.x * (size_t)blockDim.x + threadIdx.x; w < total; w += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(w % chunks);
        const size_t i = w / chunks;
        const int s = idx[i];
        if (s >= 0) *reinterpret_cast<uint4*>(dst + (size_t)s * ldd + c * 8) = *reinterpret_cast<const uint4*>(src + i * lds + c * 8);
    }
}
// dst[idx[i], :] += scale * src[i, :]   (idx unique -> no atomics; idx < 0 skipped): the partial-LoRA update of
// InternLM-XComposer2 (build_mlp.py:194-203: res[im_mask] += Plora_B(Plora_A(x[im_mask])) * scaling)
template <typename DT>
__global__ void scatter_add_rows_kernel(const __nv_bfloat16* __restrict__ src, long long lds, const int* __restrict__ idx,
                                        DT* __restrict__ dst, long long ldd, int n, int cols, float scale) {
    const int chunks = cols >> 3;
    const size_t total = (size_t)n * chunks;
    for (size_t w = blockIdx.x * (size_t)blockDim.x + threadIdx.x; w < total; w += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(w % chunks);
        const size_t i = w / chunks;
        const int s = idx[i];
        if (s < 0) continue;
        float a[8];
        unpack8e(*reinterpret_cast<const uint4*>(src + i * lds + c * 8), a);
        DT* d = dst + (size_t)s * ldd + c * 8;
        if constexpr (sizeof(DT) == 2) {
            float b[8];
            unpack8e(*reinterpret_cast<const uint4*>(d), b);
#pragma unroll
            for (int j = 0; j < 8; ++j) b[j] += scale * a[j];
            *reinterpret_cast<uint4*>(d) = pack8e(b);
        } else {
            float4 lo = *reinterpret_cast<const float4*>(d), hi = *reinterpret_cast<const float4*>(d + 4);
            lo.x += scale * a[0]; lo.y += scale * a[1]; lo.z += scale * a[2]; lo.w += scale * a[3];
            hi.x += scale * a[4]; hi.y += scale * a[5]; hi.z += scale * a[6]; hi.w += scale * a[7];
            *reinterpret_cast<float4*>(d) = lo;
            *reinterpret_cast<float4*>(d + 4) = hi;
        }
    }
}

// ---------------------------------------------------------------- LLaVA merge index (Llava/__init__.py:36-109)
// One thread per sequence scans its L text tokens (integer work, ~L iterations).  Requires every sequence
// to hold the same number of <image> tokens (nb_image_pad == 0) -- the host checks that.
//   src_map[b,p]   >= 0: embed_tokens row; -1-k: image feature row k; INT_MIN: zero row (pad token / unused)
//   labels_m[b,p], mask_m[b,p], pos_ids[b,p] as in the reference; seqlen[b] = #attended positions (prefix)
//   img_pos[b, s*P+i] = merged position of patch i of image slot s
//   row_of_text[b, j-1] = flat merged row (b*S + p(j) - 1) whose logits predict text token j; target[b, j-1]
__global__ void merge_index_kernel(const int64_t* __restrict__ ids, const int64_t* __restrict__ amask,
                                   const int64_t* __restrict__ labels, int n_seq, int L, int S, int P, int n_img_batch,
                                   int imgs_per_seq, int image_token, int pad_token, int ignore_index,
                                   int* __restrict__ src_map, int64_t* __restrict__ labels_m, int* __restrict__ mask_m,
                                   int* __restrict__ pos_ids, int* __restrict__ seqlen, int* __restrict__ img_pos,
                                   int* __restrict__ row_of_text, int64_t* __restrict__ target, int* __restrict__ status) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_seq) return;
    const int64_t* id = ids + (size_t)b * L;
    const int64_t* am = amask + (size_t)b * L;
    const int64_t* lb = labels + (size_t)b * L;
    int* sm = src_map + (size_t)b * S;
    int64_t* lm = labels_m + (size_t)b * S;
    int* mm = mask_m + (size_t)b * S;
    for (int p = 0; p < S; ++p) { sm[p] = INT_MIN; lm[p] = ignore_index; mm[p] = 0; }
    int p = 0, slot = 0;
    const int img_base = (b % n_img_batch) * imgs_per_seq;
    for (int j = 0; j < L; ++j) {
        const int64_t t = id[j];
        if (t == image_token) {
            if (slot < imgs_per_seq && p + P <= S) {
                for (int i = 0; i < P; ++i) {
                    sm[p + i] = -1 - ((img_base + slot) * P + i);
                    mm[p + i] = 1;
                    img_pos[((size_t)b * imgs_per_seq + slot) * P + i] = p + i;
                }
            } else {
                atomicExch(status, 1);  // more image tokens than the batch layout allows
            }
            if (j >= 1) { row_of_text[(size_t)b * (L - 1) + j - 1] = b * S + p - 1; target[(size_t)b * (L - 1) + j - 1] = -100; }
            p += P;
            ++slot;
        } else {
            if (p < S) {
                sm[p] = (t == pad_token) ? INT_MIN : (int)t;  // pad rows are zeroed (:100-104)
                mm[p] = (int)am[j];
                lm[p] = lb[j];
            } else {
                atomicExch(status, 1);
            }
            if (j >= 1) {
                row_of_text[(size_t)b * (L - 1) + j - 1] = b * S + p - 1;
                target[(size_t)b * (L - 1) + j - 1] = lb[j] == ignore_index ? -100 : lb[j];
            }
            ++p;
        }
    }
    if (slot != imgs_per_seq || p != S) atomicExch(status, 2);  // ragged image counts: nb_image_pad != 0
    // position_ids = cumsum(mask) - 1, masked positions -> 1 (:98); seqlen = attended prefix length
    int run = 0, len = 0;
    bool prefix = true;
    for (int q = 0; q < S; ++q) {
        if (mm[q]) {
            pos_ids[(size_t)b * S + q] = run;
            ++run;
            if (!prefix) atomicExch(status, 3);  // attention mask is not a prefix (left padding): unsupported
            len = q + 1;
        } else {
            pos_ids[(size_t)b * S + q] = 1;
            prefix = false;
        }
    }
    seqlen[b] = len;
}

// out[r, :] = embed[src] | image_features[-1-src] | 0
template <typename OT>
__global__ void merge_embed_kernel(const int* __restrict__ src_map, const __nv_bfloat16* __restrict__ embed,
                                   const __nv_bfloat16* __restrict__ img, OT* __restrict__ out, int rows, int d) {
    const int chunks = d >> 3;
    const size_t total = (size_t)rows * chunks;
    for (size_t w = blockIdx.x * (size_t)blockDim.x + threadIdx.x; w < total; w += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(w % chunks);
        const size_t r = w / chunks;
        const int s = src_map[r];
        uint4 v = make_uint4(0, 0, 0, 0);
        if (s >= 0) v = *reinterpret_cast<const uint4*>(embed + (size_t)s * d + c * 8);
        else if (s != INT_MIN) v = *reinterpret_cast<const uint4*>(img + (size_t)(-1 - s) * d + c * 8);
        if constexpr (sizeof(OT) == 2) {
            *reinterpret_cast<uint4*>(out + r * d + c * 8) = v;
        } else {
            float f[8];
            unpack8e(v, f);
            *reinterpret_cast<float4*>(out + r * d + c * 8) = make_float4(f[0], f[1], f[2], f[3]);
            *reinterpret_cast<float4*>(out + r * d + c * 8 + 4) = make_float4(f[4], f[5], f[6], f[7]);
        }
    }
}
// text rows: dembed[id] += dx[r] (fp32 atomics: token ids repeat);  image rows handled by merge_img_bwd_kernel
__global__ void merge_embed_bwd_kernel(const int* __restrict__ src_map, const __nv_bfloat16* __restrict__ dx,
                                       float* __restrict__ dembed, int rows, int d) {
    const int chunks = d >> 3;
    const size_t total = (size_t)rows * chunks;
    for (size_t w = blockIdx.x * (size_t)blockDim.x + threadIdx.x; w < total; w += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(w % chunks);
        const size_t r = w / chunks;
        const int s = src_map[r];
        if (s < 0) continue;
        float f[8];
        unpack8e(*reinterpret_cast<const uint4*>(dx + r * d + c * 8), f);
        float* o = dembed + (size_t)s * d + c * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(o + j, f[j]);
    }
}
// dimg[k, :] = sum over the sequences that share image k (chosen b and rejected b + n_img_batch, ...)
__global__ void merge_img_bwd_kernel(const int* __restrict__ img_pos, const __nv_bfloat16* __restrict__ dx,
                                     __nv_bfloat16* __restrict__ dimg, int n_seq, int n_img_batch, int S, int feats_per_seq,
                                     int d) {
    const int chunks = d >> 3;
    const size_t total = (size_t)n_img_batch * feats_per_seq * chunks;
    for (size_t w = blockIdx.x * (size_t)blockDim.x + threadIdx.x; w < total; w += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(w % chunks);
        const int f = (int)((w / chunks) % feats_per_seq);
        const int bi = (int)(w / ((size_t)chunks * feats_per_seq));
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int b = bi; b < n_seq; b += n_img_batch) {
            const int p = img_pos[(size_t)b * feats_per_seq + f];
            if (p < 0) continue;   // shared-prefix rows: the image rows of a pair exist once (the rejected copy is dropped)
            float v[8];
            unpack8e(*reinterpret_cast<const uint4*>(dx + ((size_t)b * S + p) * d + c * 8), v);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += v[j];
        }
        *reinterpret_cast<uint4*>(dimg + ((size_t)bi * feats_per_seq + f) * d + c * 8) = pack8e(acc);
    }
}

// ---------------------------------------------------------------- packed (ragged) rows: drop the padding rows of a merge index
// The merge kernels above lay sequence b out at rows [b*S, (b+1)*S).  With right padding only the first len[b] rows of a
// sequence are attended; packing keeps those and stacks the sequences back to back (sequence b at row_starts[b],
// row_starts = exclusive prefix sum of len), so every row-wise kernel and GEMM runs over sum(len) rows instead of n_seq*S.
//   src_map_p / pos_p [row_starts[n_seq]]: the surviving rows of src_map / pos
//   row lists (row_of_text, img_rows: flat padded rows b*S + p) are rewritten in place to packed rows, -1 where p >= len[b]
//   (vlb200_gather_rows yields a zero row, vlb200_scatter_rows skips it);  img_pos (position p inside sequence b) becomes
//   the absolute packed row row_starts[b] + p (vlb200_llava_merge_bwd_rows with row_stride = 0).
__global__ void pack_rows_kernel(const int* __restrict__ src_map, const int* __restrict__ pos, const int* __restrict__ row_starts,
                                 int n_seq, int S, int* __restrict__ src_map_p, int* __restrict__ pos_p) {
    const size_t total = (size_t)n_seq * S;
    for (size_t r = blockIdx.x * (size_t)blockDim.x + threadIdx.x; r < total; r += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(r / S), p = (int)(r % S);
        const int start = row_starts[b];
        if (p < row_starts[b + 1] - start) {
            src_map_p[start + p] = src_map[r];
            pos_p[start + p] = pos[r];
        }
    }
}
__global__ void remap_rows_kernel(int* __restrict__ rows, size_t n, const int* __restrict__ row_starts, int n_seq, int S) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int r = rows[i];
        if (r < 0) continue;
        const int b = r / S, p = r % S;
        rows[i] = (b < n_seq && p < row_starts[b + 1] - row_starts[b]) ? row_starts[b] + p : -1;
    }
}
__global__ void abs_img_pos_kernel(int* __restrict__ img_pos, size_t n, int feats_per_seq, const int* __restrict__ row_starts) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        img_pos[i] += row_starts[i / feats_per_seq];
}
// ---- shared-prefix rows: the first pre[b] merged rows of the chosen and the rejected sequence of a pair
// are the same computation (same prompt tokens, same image, causal attention), so they are laid out ONCE.  Padded row (b, p):
//   p < pre[b]  -> pre_start[b] + p            (pre / pre_start are equal for the two sequences of a pair)
//   p < len[b]  -> suf_start[b] + (p - pre[b])
//   otherwise   -> dropped (-1)
__device__ __forceinline__ int shared_row(int b, int p, const int* pre, const int* pre_start, const int* suf_start, const int* len) {
    if (p >= len[b]) return -1;
    const int q = pre[b];
    return p < q ? pre_start[b] + p : suf_start[b] + (p - q);
}
__global__ void share_rows_kernel(const int* __restrict__ src_map, const int* __restrict__ pos, const int* __restrict__ pre,
                                  const int* __restrict__ pre_start, const int* __restrict__ suf_start, const int* __restrict__ len,
                                  int n_seq, int S, int* __restrict__ src_map_s, int* __restrict__ pos_s) {
    const size_t total = (size_t)n_seq * S;
    for (size_t r = blockIdx.x * (size_t)blockDim.x + threadIdx.x; r < total; r += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(r / S), p = (int)(r % S);
        const int dst = shared_row(b, p, pre, pre_start, suf_start, len);
        // a shared row is written by both sequences of the pair with the same values; the first half (chosen) is the writer
        if (dst >= 0 && (p >= pre[b] || b < n_seq / 2)) {
            src_map_s[dst] = src_map[r];
            pos_s[dst] = pos[r];
        }
    }
}
__global__ void share_remap_rows_kernel(int* __restrict__ rows, size_t n, const int* __restrict__ pre, const int* __restrict__ pre_start,
                                        const int* __restrict__ suf_start, const int* __restrict__ len, int n_seq, int S) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int r = rows[i];
        if (r < 0) continue;
        const int b = r / S, p = r % S;
        rows[i] = b < n_seq ? shared_row(b, p, pre, pre_start, suf_start, len) : -1;
    }
}
// img_pos (position inside sequence i / feats_per_seq) -> absolute shared row; the rejected copy of a row that lies in the
// shared prefix becomes -1 (vlb200_llava_merge_bwd_rows then counts the row once)
__global__ void share_img_pos_kernel(int* __restrict__ img_pos, size_t n, int feats_per_seq, const int* __restrict__ pre,
                                     const int* __restrict__ pre_start, const int* __restrict__ suf_start, const int* __restrict__ len,
                                     int n_seq) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(i / feats_per_seq), p = img_pos[i];
        if (p < 0) continue;
        img_pos[i] = (p < pre[b] && b >= n_seq / 2) ? -1 : shared_row(b, p, pre, pre_start, suf_start, len);
    }
}

// ---------------------------------------------------------------- LLaVA-Next merge index (LlavaNext/__init__.py:38-171)
// Same outputs as merge_index_kernel, but (i) image k contributes feat_off[k+1]-feat_off[k] packed feature rows
// (anyres: variable per image), (ii) tokens with attention_mask == 0 are never written (:96-99,124-127) and S is the
// longest VALID merged sequence (host-computed), (iii) pad-token embeddings are not zeroed.  Right padding only.
//   img_rows[rep*total_feats + k] = flat merged row holding packed feature row k in the rep-th sequence that
//   uses it (rep = b / n_img_batch: chosen, rejected).
__global__ void next_merge_index_kernel(const int64_t* __restrict__ ids, const int64_t* __restrict__ amask,
                                        const int64_t* __restrict__ labels, const int* __restrict__ feat_off, int n_seq,
                                        int L, int S, int n_img_batch, int imgs_per_seq, int total_feats, int image_token,
                                        int ignore_index, int* __restrict__ src_map, int64_t* __restrict__ labels_m,
                                        int* __restrict__ mask_m, int* __restrict__ pos_ids, int* __restrict__ seqlen,
                                        int* __restrict__ img_rows, int* __restrict__ row_of_text,
                                        int64_t* __restrict__ target, int* __restrict__ status) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_seq) return;
    const int64_t* id = ids + (size_t)b * L;
    const int64_t* am = amask + (size_t)b * L;
    const int64_t* lb = labels + (size_t)b * L;
    int* sm = src_map + (size_t)b * S;
    int64_t* lm = labels_m + (size_t)b * S;
    int* mm = mask_m + (size_t)b * S;
    for (int p = 0; p < S; ++p) { sm[p] = INT_MIN; lm[p] = ignore_index; mm[p] = 0; pos_ids[(size_t)b * S + p] = 1; }
    int p = 0, slot = 0;
    bool seen_masked = false;
    const int img_base = (b % n_img_batch) * imgs_per_seq;
    const int rep = b / n_img_batch;
    for (int j = 0; j < L; ++j) {
        const int64_t t = id[j];
        const size_t rj = (size_t)b * (L - 1) + j - 1;
        if (am[j] == 0) {
            seen_masked = true;
            if (t == image_token) atomicExch(status, 4);  // an <image> placeholder outside the attended prefix
            if (j >= 1) { row_of_text[rj] = b * S; target[rj] = -100; }
            continue;
        }
        if (seen_masked) atomicExch(status, 3);  // attended token after a masked one: not right padding
        if (t == image_token) {
            const int k0 = slot < imgs_per_seq ? feat_off[img_base + slot] : 0;
            const int F = slot < imgs_per_seq ? feat_off[img_base + slot + 1] - k0 : 0;
            if (slot < imgs_per_seq && p + F <= S) {
                for (int f = 0; f < F; ++f) {
                    sm[p + f] = -1 - (k0 + f);
                    mm[p + f] = 1;
                    img_rows[(size_t)rep * total_feats + k0 + f] = b * S + p + f;
                }
            } else {
                atomicExch(status, 1);
            }
            if (j >= 1) { row_of_text[rj] = b * S + (p > 0 ? p - 1 : 0); target[rj] = -100; }
            p += F;
            ++slot;
        } else {
            if (p < S) {
                sm[p] = (int)t;
                mm[p] = 1;
                lm[p] = lb[j];
            } else {
                atomicExch(status, 1);
            }
            if (j >= 1) {
                row_of_text[rj] = b * S + (p > 0 ? p - 1 : 0);
                target[rj] = (lb[j] == ignore_index || p == 0) ? -100 : lb[j];
            }
            ++p;
        }
    }
    if (slot != imgs_per_seq) atomicExch(status, 2);
    if (p > S) atomicExch(status, 1);
    const int len = p < S ? p : S;
    for (int q = 0; q < len; ++q) pos_ids[(size_t)b * S + q] = q;  // mask is a prefix: cumsum(mask) - 1
    seqlen[b] = len;
}
// dimg[k, :] = sum over reps of dx[img_rows[rep*total + k], :]
__global__ void next_merge_img_bwd_kernel(const int* __restrict__ img_rows, const __nv_bfloat16* __restrict__ dx,
                                          __nv_bfloat16* __restrict__ dimg, int total_feats, int reps, int d) {
    const int chunks = d >> 3;
    const size_t total = (size_t)total_feats * chunks;
    for (size_t w = blockIdx.x * (size_t)blockDim.x + threadIdx.x; w < total; w += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(w % chunks);
        const size_t k = w / chunks;
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int r = 0; r < reps; ++r) {
            const int row = img_rows[(size_t)r * total_feats + k];
            float v[8];
            unpack8e(*reinterpret_cast<const uint4*>(dx + (size_t)row * d + c * 8), v);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += v[j];
        }
        *reinterpret_cast<uint4*>(dimg + k * d + c * 8) = pack8e(acc);
    }
}

// ---------------------------------------------------------------- Qwen-VL image placement (modeling_qwen.py:524-528,614-621)
// The sequence already holds n_queries placeholder tokens between <img> (image_start_id) and </img> (+1): those
// positions read image feature rows, everything else (the two markers and padding included) reads its own embedding.
// S == L, labels pass through, position ids are arange(L) for every row (:573-580), seqlen = attended prefix.
__global__ void qwen_merge_index_kernel(const int64_t* __restrict__ ids, const int64_t* __restrict__ amask,
                                        const int64_t* __restrict__ labels, int n_seq, int L, int Q, int n_img_batch,
                                        int imgs_per_seq, int image_start, int ignore_index, int* __restrict__ src_map,
                                        int64_t* __restrict__ labels_m, int* __restrict__ mask_m, int* __restrict__ pos_ids,
                                        int* __restrict__ seqlen, int* __restrict__ row_of_text,
                                        int64_t* __restrict__ target, int* __restrict__ status) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_seq) return;
    const int64_t* id = ids + (size_t)b * L;
    const int64_t* am = amask + (size_t)b * L;
    const int64_t* lb = labels + (size_t)b * L;
    int* sm = src_map + (size_t)b * L;
    const int img_base = (b % n_img_batch) * imgs_per_seq;
    int slot = 0, open = -1, len = 0;
    bool prefix = true;
    for (int j = 0; j < L; ++j) {
        const int64_t t = id[j];
        sm[j] = (int)t;
        labels_m[(size_t)b * L + j] = lb[j];
        mask_m[(size_t)b * L + j] = (int)am[j];
        pos_ids[(size_t)b * L + j] = j;
        if (am[j]) { if (!prefix) atomicExch(status, 3); len = j + 1; } else prefix = false;
        if (j >= 1) {
            row_of_text[(size_t)b * (L - 1) + j - 1] = b * L + j - 1;
            target[(size_t)b * (L - 1) + j - 1] = lb[j] == ignore_index ? -100 : lb[j];
        }
        if (t == image_start) {
            if (open >= 0) atomicExch(status, 2);  // nested <img>
            open = j;
        } else if (t == image_start + 1) {
            if (open < 0 || j - open - 1 != Q || slot >= imgs_per_seq) {
                atomicExch(status, 2);
            } else {
                for (int q = 0; q < Q; ++q) sm[open + 1 + q] = -1 - ((img_base + slot) * Q + q);
            }
            open = -1;
            ++slot;
        }
    }
    if (open >= 0 || slot != imgs_per_seq) atomicExch(status, 2);
    seqlen[b] = len;
}

// ---------------------------------------------------------------- optimizer
__global__ void sumsq_partial_kernel(const __nv_bfloat16* __restrict__ x, size_t n8, float* __restrict__ partial) {
    __shared__ float sm[8];
    float s = 0.f;
    for (size_t w = blockIdx.x * (size_t)blockDim.x + threadIdx.x; w < n8; w += (size_t)gridDim.x * blockDim.x) {
        float f[8];
        unpack8e(ld_nc_v4(x + w * 8), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) s += f[j] * f[j];
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += sm[k];
        partial[blockIdx.x] = t;
    }
}
__global__ void sumsq_final_kernel(const float* __restrict__ partial, int n, float* __restrict__ out, int accumulate) {
    __shared__ float sm[8];
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += partial[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += sm[k];
        out[0] = accumulate ? out[0] + t : t;
    }
}

// torch.optim.AdamW semantics (decoupled weight decay), fp32 master + moments, bf16 params/grads.
// grad is pre-multiplied by grad_scale and by the clip coefficient min(1, max_norm / (sqrt(sumsq)*grad_scale + 1e-6)).
__global__ void adamw_kernel(__nv_bfloat16* __restrict__ p, const __nv_bfloat16* __restrict__ g, float* __restrict__ master,
                             float* __restrict__ m, float* __restrict__ v, size_t n8, float lr, float beta1, float beta2,
                             float eps, float wd, float bc1, float bc2_sqrt, float grad_scale,
                             const float* __restrict__ sumsq, float max_norm) {
    float gs = grad_scale;
    if (sumsq != nullptr && max_norm > 0.f) {
        const float norm = sqrtf(sumsq[0]) * grad_scale;
        gs *= fminf(1.f, max_norm / (norm + 1e-6f));
    }
    const float step = lr / bc1;
    const float decay = 1.f - lr * wd;
    for (size_t w = blockIdx.x * (size_t)blockDim.x + threadIdx.x; w < n8; w += (size_t)gridDim.x * blockDim.x) {
        float gf[8];
        unpack8e(ld_nc_v4(g + w * 8), gf);
        float pm[8], mm[8], vv[8];
        const uint4 a0 = ld_nc_v4(master + w * 8), a1 = ld_nc_v4(master + w * 8 + 4);
        const uint4 b0 = ld_nc_v4(m + w * 8), b1 = ld_nc_v4(m + w * 8 + 4);
        const uint4 c0 = ld_nc_v4(v + w * 8), c1 = ld_nc_v4(v + w * 8 + 4);
        const uint32_t aw[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const uint32_t bw[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        const uint32_t cw[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float gj = gf[j] * gs;
            mm[j] = beta1 * __uint_as_float(bw[j]) + (1.f - beta1) * gj;
            vv[j] = beta2 * __uint_as_float(cw[j]) + (1.f - beta2) * gj * gj;
            const float denom = sqrtf(vv[j]) / bc2_sqrt + eps;
            pm[j] = __uint_as_float(aw[j]) * decay - step * (mm[j] / denom);
        }
        st_na_v4(master + w * 8, make_uint4(__float_as_uint(pm[0]), __float_as_uint(pm[1]), __float_as_uint(pm[2]), __float_as_uint(pm[3])));
        st_na_v4(master + w * 8 + 4, make_uint4(__float_as_uint(pm[4]), __float_as_uint(pm[5]), __float_as_uint(pm[6]), __float_as_uint(pm[7])));
        st_na_v4(m + w * 8, make_uint4(__float_as_uint(mm[0]), __float_as_uint(mm[1]), __float_as_uint(mm[2]), __float_as_uint(mm[3])));
        st_na_v4(m + w * 8 + 4, make_uint4(__float_as_uint(mm[4]), __float_as_uint(mm[5]), __float_as_uint(mm[6]), __float_as_uint(mm[7])));
        st_na_v4(v + w * 8, make_uint4(__float_as_uint(vv[0]), __float_as_uint(vv[1]), __float_as_uint(vv[2]), __float_as_uint(vv[3])));
        st_na_v4(v + w * 8 + 4, make_uint4(__float_as_uint(vv[4]), __float_as_uint(vv[5]), __float_as_uint(vv[6]), __float_as_uint(vv[7])));
        st_na_v4(p + w * 8, pack8e(pm));
    }
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, size_t n, float scale) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = __float2bfloat16(src[i] * scale);
}
__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = __bfloat162float(src[i]);
}

}  // namespace vlb

using namespace vlb;
#define BF(p) reinterpret_cast<__nv_bfloat16*>(p)
#define CBF(p) reinterpret_cast<const __nv_bfloat16*>(p)

extern "C" int vlb200_rope(void* qkv, int64_t ld, const int* pos, const float* cos_table, const float* sin_table, int table_rows,
                           int rows, int n_rot_heads, int head_dim, int inverse, void* stream) {
    VLB_REQUIRE(qkv && pos && cos_table && sin_table && table_rows > 0, "rope: null pointer / empty table");
    VLB_REQUIRE(head_dim % 16 == 0 && ld % 8 == 0, "rope: head_dim must be a multiple of 16, ld of 8");
    if (rows <= 0 || n_rot_heads <= 0) return VLB200_OK;
    const size_t work = (size_t)rows * n_rot_heads * (head_dim / 16);
    rope_kernel<<<grid_for(work, 256), 256, 0, as_stream(stream)>>>(BF(qkv), ld, pos, cos_table, sin_table, table_rows, rows,
                                                                   n_rot_heads, head_dim, inverse ? -1.f : 1.f);
    count_launch();
    VLB_LAUNCH_CHECK();
    return VLB200_OK;
}

extern "C" int vlb200_swiglu_fwd(const void* gate_up, int64_t ld_gu, void* act, int64_t ld_act, int rows, int ff, void* stream) {
    VLB_REQUIRE(gate_up && act && ff % 8 == 0 && ld_gu % 8 == 0 && ld_act % 8 == 0, "swiglu_fwd: bad arguments");
    if (rows <= 0) return VLB200_OK;
    swiglu_fwd_kernel<<<grid_for((size_t)rows * (ff / 8), 256), 256, 0, as_stream(stream)>>>(CBF(gate_up), ld_gu, BF(act), ld_act, rows, ff);
    count_launch();
    VLB_LAUNCH_CHECK();
    return VLB200_OK;
}
extern "C" int vlb200_swiglu_bwd(const void* gate_up, int64_t ld_gu, const void* dact, int64_t ld_dact, void* dgate_up,
                                 int64_t ld_dgu, int rows, int ff, void* stream) {
    VLB_REQUIRE(gate_up && dact && dgate_up && ff % 8 == 0 && ld_gu % 8 == 0 && ld_dact % 8 == 0 && ld_dgu % 8 == 0, "swiglu_bwd: bad arguments");
    if (rows <= 0) return VLB200_OK;
    swiglu_bwd_kernel<<<grid_for((size_t)rows * (ff / 8), 256), 256, 0, as_stream(stream)>>>(CBF(gate_up), ld_gu, CBF(dact), ld_dact, BF(dgate_up), ld_dgu, rows, ff);
    count_launch();
    VLB_LAUNCH_CHECK();
    return VLB200_OK;
}
extern "C" int vlb200_gelu_fwd(const void* z, void* h, uint64_t n, void* stream) {
    VLB_REQUIRE(z && h && n % 8 == 0, "gelu_fwd: n must be a multiple of 8");
    if (n == 0) return VLB200_OK;
    gelu_fwd_kernel<<<grid_for(n / 8, 256), 256, 0, as_stream(stream)>>>(CBF(z), BF(h), n / 8);
    count_launch();
    VLB_LAUNCH_CHECK();
    return VLB200_OK;
}
extern "C" int vlb200_gelu_bwd(const void* z, const void* dh, void* dz, uint64_t n, void* stream) {
    VLB_REQUIRE(z && dh && dz && n % 8 == 0, "gelu_bwd: n must be a multiple of 8");
    if (n == 0) return VLB200_OK;
    gelu_bwd_kernel<<<grid_for(n / 8, 256), 256, 0, as_stream(stream)>>>(CBF(z), CBF(dh), BF(dz), n / 8);
    count_launch();
    VLB_LAUNCH_CHECK();
    return VLB200_OK;
}

extern "C" int vlb200_clip_im2col(const void* pixels, int pixel_dtype, void* patches, int64_t ld_out, int batch, int height,
                                  int width, int patch, void* stream) {
    VLB_REQUIRE(pixels && patches, "im2col: null pointer");
    VLB_REQUIRE(height % patch == 0 && width % patch == 0 && ld_out >= 3 * patch * patch, "im2col: bad geometry");
    const size_t work = (size_t)batch * (height / patch) * (width / patch) * 3 * patch * patch;
    if (work == 0) return VLB200_OK;
    if (pixel_dtype == VLB200_F32)
        im2col_kernel<float><<<grid_for(work, 256), 256, 0, as_stream(stream)>>>((const float*)pixels, BF(patches), ld_out, batch, height, width, patch);
    else if (pixel_dtype == VLB200_BF16)
        im2col_kernel<__nv_bfloat16><<<grid_for(work, 256), 256, 0, as_stream(stream)>>>(CBF(pixels), BF(patches), ld_out, batch, height, width, patch);
    else
        VLB_REQUIRE(false, "im2col: bad dtype");
    count_launch();
    VLB_LAUNCH_CHECK();
    return VLB200_OK;
}
extern "C" int vlb200_clip_cls_rows(void* x, const void* cls, const void* pos0, int batch, int tokens_per_img, int d, void* stream) {
    VLB_REQUIRE(x && cls && pos0, "cls_rows: null pointer");
    cls_rows_kernel<<<grid_for((size_t)batch * d, 256), 256, 0, as_stream(stream)>>>(BF(x), CBF(cls), CBF(pos0), batch, tokens_per_img, d);
    count_launch();
    VLB_LAUNCH_CHECK();
    return VLB200_OK;
}

extern "C" int vlb200_copy_rows(const void* src, int64_t src_group_stride, int64_t src_row_stride, int src_row0, void* dst,
                                int64_t dst_group_stride, int64_t dst_row_stride, int groups, int rows_per_group, int cols,
                                void* stream) {
    VLB_REQUIRE(src && dst && cols % 8 == 0 && src_row_stride % 8 == 0 && dst_row_stride % 8 == 0 &&
                    src_group_stride % 8 == 0 && dst_group_stride % 8 == 0, "copy_rows: alignment");
    const size_t work = (size_t)groups * rows_per_group * (cols / 8);
    if (work == 0) return VLB200_OK;
    copy_rows_kernel<<<grid_for(work, 256), 256, 0, as_stream(stream)>>>(CBF(src), src_group_stride, src_row_stride, src_row0, BF(dst), dst_group_stride, dst_row_stride, groups, rows_per_group, cols);
    count_launch();
    VLB_LAUNCH_CHECK();
    return VLB200_OK;
}
extern "C" int vlb200_gather_rows(const void* src, int64_t ld_src, const int* index, void* dst, int64_t ld_dst, int n, int cols, void* stream) {
    VLB_REQUIRE(src && index && dst && cols % 8 == 0 && ld_src % 8 == 0 && ld_dst % 8 == 0, "gather_rows: alignment");
    if (n <= 0) return VLB200_OK;
    gather_rows_kernel<<<grid_for((size_t)n * (cols / 8), 256), 256, 0, as_stream(stream)>>>(CBF(src), ld_src, index, BF(dst), ld_dst, n, cols);
    count_launch();
    VLB_LAUNCH_CHECK();
    return VLB200_OK;
}
extern "C" int vlb200_scatter_rows(const void* src, int64_t ld_src, const int* index, void* dst, int64_t ld_dst, int n, int cols, void* stream) {
    VLB_REQUIRE(src && index && dst && cols % 8 == 0 && ld_src % 8 == 0 && ld_dst % 8 == 0, "scatter_rows: alignment");
    if (n <= 0) return VLB200_OK;
    scatter_rows_kernel<<<grid_for((size_t)n * (cols / 8), 256), 256, 0, as_stream(stream)>>>(CBF(src), ld_src, index, BF(dst), ld_dst, n, cols);
    count_launch();
    VLB_LAUNCH_CHECK();
    return VLB200_OK;
}
extern "C" int vlb200_scatter_add_rows(const void* src, int64_t ld_src, const int* index, void* dst, int dst_dtype, int64_t ld_dst,
                                       int n, int cols, float scale, void* stream) {
    VLB_REQUIRE(src && index && dst && cols % 8 == 0 && ld_src % 8 == 0 && ld_dst % 8 == 0, "scatter_add_rows: alignment");
    VLB_REQUIRE(dst_dtype == VLB200_BF16 || dst_dtype == VLB200_F32, "scatter_add_rows: bad dst dtype");
    if (n <= 0) return VLB200_OK;
    if (dst_dtype == VLB200_F32)
        scatter_add_rows_kernel<float><<<grid_for((size_t)n * (cols / 8), 256), 256, 0, as_stream(stream)>>>(CBF(src), ld_src, index, (float*)dst, ld_dst, n, cols, scale);
    else
        scatter_add_rows_kernel<__nv_bfloat16><<<grid_for((size_t)n * (cols / 8), 256), 256, 0, as_stream(stream)>>>(CBF(src), ld_src, index, BF(dst), ld_dst, n, cols, scale);
    count_launch();
    VLB_LAUNCH_CHECK();
    return VLB200_OK;
}
extern "C" int vlb200_memset_zero(void* dst, uint64_t bytes, void* stream) {
    if (bytes == 0) return VLB200_OK;
    VLB_REQUIRE(dst, "memset_zero: null pointer");
    VLB_CHECK_CUDA(cudaMemsetAsync(dst, 0, bytes, as_stream(stream)));
    return VLB200_OK;
}

extern "C" int vlb200_pack_merge_rows(const int* src_map, const int* position_ids, const int* row_starts, int n_seq, int merged_len,
                                      int* src_map_packed, int* position_ids_packed, int* row_of_text, int64_t n_text_rows,
                                      int* img_pos, int64_t n_img_pos, int feats_per_seq, int* img_rows, int64_t n_img_rows,
                                      void* stream) {
    VLB_REQUIRE(src_map && position_ids && row_starts && src_map_packed && position_ids_packed && n_seq > 0 && merged_len > 0,
                "pack_merge_rows: bad arguments");
    VLB_REQUIRE(img_pos == nullptr || feats_per_seq > 0, "pack_merge_rows: img_pos needs feats_per_seq");
    cudaStream_t s = as_stream(stream);
    pack_rows_kernel<<<grid_for((size_t)n_seq * merged_len, 256), 256, 0, s>>>(src_map, position_ids, row_starts, n_seq, merged_len,
                                                                             src_map_packed, position_ids_packed);
    count_launch();
    if (row_of_text != nullptr && n_text_rows > 0) {
        remap_rows_kernel<<<grid_for((size_t)n_text_rows, 256), 256, 0, s>>>(row_of_text, (size_t)n_text_rows, row_starts, n_seq, merged_len);
        count_launch();
    }
    if (img_rows != nullptr && n_img_rows > 0) {
        remap_rows_kernel<<<grid_for((size_t)n_img_rows, 256), 256, 0, s>>>(img_rows, (size_t)n_img_rows, row_starts, n_seq, merged_len);
        count_launch();
    }
    if (img_pos != nullptr && n_img_pos > 0) {
        abs_img_pos_kernel<<<grid_for((size_t)n_img_pos, 256), 256, 0, s>>>(img_pos, (size_t)n_img_pos, feats_per_seq, row_starts);
        count_launch();
    }
    VLB_LAUNCH_CHECK();
    return VLB200_OK;
}
extern "C" int vlb200_share_prefix_rows(const int* src_map, const int* position_ids, const int* prefix_rows, const int* prefix_starts,
                                        const int* suffix_starts, const int* seq_lens, int n_seq, int merged_len, int64_t total_rows,
                                        int* src_map_shared, int* position_ids_shared, int* row_of_text, int64_t n_text_rows,
                                        int* img_pos, int64_t n_img_pos, int feats_per_seq, int* img_rows, int64_t n_img_rows,
                                        void* stream) {
    VLB_REQUIRE(src_map && position_ids && prefix_rows && prefix_starts && suffix_starts && seq_lens && src_map_shared &&
                    position_ids_shared && n_seq > 0 && n_seq % 2 == 0 && merged_len > 0 && total_rows > 0,
                "share_prefix_rows: bad arguments");
    VLB_REQUIRE(img_pos == nullptr || feats_per_seq > 0, "share_prefix_rows: img_pos needs feats_per_seq");
    cudaStream_t s = as_stream(stream);
    share_rows_kernel<<<grid_for((size_t)n_seq * merged_len, 256), 256, 0, s>>>(src_map, position_ids, prefix_rows, prefix_starts,
                                                                              suffix_starts, seq_lens, n_seq, merged_len,
                                                                              src_map_shared, position_ids_shared);
    count_launch();
    if (row_of_text != nullptr && n_text_rows > 0) {
        share_remap_rows_kernel<<<grid_for((size_t)n_text_rows, 256), 256, 0, s>>>(row_of_text, (size_t)n_text_rows, prefix_rows,
                                                                                 prefix_starts, suffix_starts, seq_lens, n_seq, merged_len);
        count_launch();
    }
    if (img_rows != nullptr && n_img_rows > 0) {
        share_remap_rows_kernel<<<grid_for((size_t)n_img_rows, 256), 256, 0, s>>>(img_rows, (size_t)n_img_rows, prefix_rows,
                                                                                prefix_starts, suffix_starts, seq_lens, n_seq, merged_len);
        count_launch();
    }
    if (img_pos != nullptr && n_img_pos > 0) {
        share_img_pos_kernel<<<grid_for((size_t)n_img_pos, 256), 256, 0, s>>>(img_pos, (size_t)n_img_pos, feats_per_seq, prefix_rows,
                                                                            prefix_starts, suffix_starts, seq_lens, n_seq);
        count_launch();
    }
    VLB_LAUNCH_CHECK();
    return VLB200_OK;
}
extern "C" int vlb200_llava_merge_index(const int64_t* input_ids, const int64_t* attention_mask, const int64_t* labels,
                                        int n_seq, int text_len, int merged_len, int n_patches, int n_img_batch,
                                        int imgs_per_seq, int image_token, int pad_token, int ignore_index, int* src_map,
                                        int64_t* labels_merged, int* mask_merged, int* position_ids, int* seqlens,
                                        int* img_pos, int* row_of_text, int64_t* target, int* status, void* stream) {
    VLB_REQUIRE(input_ids && attention_mask && labels && src_map && labels_merged && mask_merged && position_ids &&
                    seqlens && img_pos && row_of_text && target && status, "merge_index: null pointer");
    VLB_REQUIRE(merged_len == text_len + imgs_per_seq * (n_patches - 1), "merge_index: merged_len != text_len + imgs*(P-1)");
    VLB_REQUIRE(n_img_batch > 0 && n_seq % n_img_batch == 0, "merge_index: n_seq must be a multiple of the image batch");
    VLB_CHECK_CUDA(cudaMemsetAsync(status, 0, sizeof(int), as_stream(stream)));
    merge_index_kernel<<<(n_seq + 31) / 32, 32, 0, as_stream(stream)>>>(input_ids, attention_mask, labels, n_seq, text_len, merged_len, n_patches, n_img_batch, imgs_per_seq, image_token, pad_token, ignore_index, src_map, labels_merged, mask_merged, position_ids, seqlens, img_pos, row_of_text, target, status);
    count_launch();
    VLB_LAUNCH_CHECK();
    return VLB200_OK;
}
extern "C" int vlb200_llava_merge_embed(const int* src_map, const void* embed_tokens, const void* image_features, void* out,
                                        int out_dtype, int rows, int d, void* stream) {
    VLB_REQUIRE(src_map && embed_tokens && image_features && out && d % 8 == 0, "merge_embed: bad arguments");
    VLB_REQUIRE(out_dtype == VLB200_BF16 || out_dtype == VLB200_F32, "merge_embed: bad out dtype");
    if (out_dtype == VLB200_F32)
        merge_embed_kernel<float><<<grid_for((size_t)rows * (d / 8), 256), 256, 0, as_stream(stream)>>>(src_map, CBF(embed_tokens), CBF(image_features), (float*)out, rows, d);
    else
        merge_embed_kernel<__nv_bfloat16><<<grid_for((size_t)rows * (d / 8), 256), 256, 0, as_stream(stream)>>>(src_map, CBF(embed_tokens), CBF(image_features), BF(out), rows, d);
    count_launch();
    VLB_LAUNCH_CHECK();
    return VLB200_OK;
}
extern "C" int vlb200_llava_merge_bwd_rows(const int* src_map, const int* img_pos, const void* dx, float* dembed_f32,
                                           void* dimage_features, int64_t n_rows, int n_seq, int n_img_batch, int row_stride,
                                           int feats_per_seq, int d, void* stream) {
    VLB_REQUIRE(src_map && img_pos && dx && dembed_f32 && dimage_features && d % 8 == 0 && n_rows > 0, "merge_bwd: bad arguments");
    cudaStream_t s = as_stream(stream);
    merge_embed_bwd_kernel<<<grid_for((size_t)n_rows * (d / 8), 256), 256, 0, s>>>(src_map, CBF(dx), dembed_f32, (int)n_rows, d);
    VLB_LAUNCH_CHECK();
    merge_img_bwd_kernel<<<grid_for((size_t)n_img_batch * feats_per_seq * (d / 8), 256), 256, 0, s>>>(img_pos, CBF(dx), BF(dimage_features), n_seq, n_img_batch, row_stride, feats_per_seq, d);
    count_launch(2);
    VLB_LAUNCH_CHECK();
    return VLB200_OK;
}
extern "C" int vlb200_llava_merge_bwd(const int* src_map, const int* img_pos, const void* dx, float* dembed_f32,
                                      void* dimage_features, int n_seq, int n_img_batch, int merged_len, int feats_per_seq,
                                      int d, void* stream) {
    return vlb200_llava_merge_bwd_rows(src_map, img_pos, dx, dembed_f32, dimage_features, (int64_t)n_seq * merged_len, n_seq,
                                       n_img_batch, merged_len, feats_per_seq, d, stream);
}
extern "C" int vlb200_llavanext_merge_index(const int64_t* input_ids, const int64_t* attention_mask, const int64_t* labels,
                                            const int* feat_off, int n_seq, int text_len, int merged_len, int n_img_batch,
                                            int imgs_per_seq, int total_feats, int image_token, int ignore_index,
                                            int* src_map, int64_t* labels_merged, int* mask_merged, int* position_ids,
                                            int* seqlens, int* img_rows, int* row_of_text, int64_t* target, int* status,
                                            void* stream) {
    VLB_REQUIRE(input_ids && attention_mask && labels && feat_off && src_map && labels_merged && mask_merged &&
                    position_ids && seqlens && img_rows && row_of_text && target && status, "next_merge_index: null pointer");
    VLB_REQUIRE(n_img_batch > 0 && n_seq % n_img_batch == 0, "next_merge_index: n_seq must be a multiple of the image batch");
    VLB_REQUIRE(merged_len > 0 && text_len > 1 && total_feats > 0 && imgs_per_seq > 0, "next_merge_index: bad sizes");
    VLB_CHECK_CUDA(cudaMemsetAsync(status, 0, sizeof(int), as_stream(stream)));
    VLB_CHECK_CUDA(cudaMemsetAsync(img_rows, 0, sizeof(int) * (size_t)(n_seq / n_img_batch) * total_feats, as_stream(stream)));
    next_merge_index_kernel<<<(n_seq + 31) / 32, 32, 0, as_stream(stream)>>>(input_ids, attention_mask, labels, feat_off, n_seq, text_len, merged_len, n_img_batch, imgs_per_seq, total_feats, image_token, ignore_index, src_map, labels_merged, mask_merged, position_ids, seqlens, img_rows, row_of_text, target, status);
    count_launch();
    VLB_LAUNCH_CHECK();
    return VLB200_OK;
}
extern "C" int vlb200_llavanext_merge_bwd(const int* src_map, const int* img_rows, const void* dx, float* dembed_f32,
                                          void* dimage_features, int n_rows, int total_feats, int reps, int d,
                                          void* stream) {
    VLB_REQUIRE(src_map && img_rows && dx && dembed_f32 && dimage_features && d % 8 == 0 && reps > 0, "next_merge_bwd: bad arguments");
    cudaStream_t s = as_stream(stream);
    merge_embed_bwd_kernel<<<grid_for((size_t)n_rows * (d / 8), 256), 256, 0, s>>>(src_map, CBF(dx), dembed_f32, n_rows, d);
    VLB_LAUNCH_CHECK();
    next_merge_img_bwd_kernel<<<grid_for((size_t)total_feats * (d / 8), 256), 256, 0, s>>>(img_rows, CBF(dx), BF(dimage_features), total_feats, reps, d);
    count_launch(2);
    VLB_LAUNCH_CHECK();
    return VLB200_OK;
}
extern "C" int vlb200_qwen_merge_index(const int64_t* input_ids, const int64_t* attention_mask, const int64_t* labels, int n_seq,
                                       int text_len, int n_queries, int n_img_batch, int imgs_per_seq, int image_start_id,
                                       int ignore_index, int* src_map, int64_t* labels_out, int* mask_out, int* position_ids,
                                       int* seqlens, int* row_of_text, int64_t* target, int* status, void* stream) {
    VLB_REQUIRE(input_ids && attention_mask && labels && src_map && labels_out && mask_out && position_ids && seqlens &&
                    row_of_text && target && status, "qwen_merge_index: null pointer");
    VLB_REQUIRE(n_img_batch > 0 && n_seq % n_img_batch == 0 && text_len > 1 && n_queries > 0 && imgs_per_seq > 0,
                "qwen_merge_index: bad sizes");
    VLB_CHECK_CUDA(cudaMemsetAsync(status, 0, sizeof(int), as_stream(stream)));
    qwen_merge_index_kernel<<<(n_seq + 31) / 32, 32, 0, as_stream(stream)>>>(input_ids, attention_mask, labels, n_seq, text_len, n_queries, n_img_batch, imgs_per_seq, image_start_id, ignore_index, src_map, labels_out, mask_out, position_ids, seqlens, row_of_text, target, status);
    count_launch();
    VLB_LAUNCH_CHECK();
    return VLB200_OK;
}

extern "C" int vlb200_sumsq_bf16(const void* x, uint64_t n, float* workspace_1024, float* out, int accumulate, void* stream) {
    VLB_REQUIRE(x && workspace_1024 && out && n % 8 == 0, "sumsq: n must be a multiple of 8");
    cudaStream_t s = as_stream(stream);
    const int grid = (int)std::min<size_t>(1024, std::max<size_t>(1, (n / 8 + 255) / 256));
    sumsq_partial_kernel<<<grid, 256, 0, s>>>(CBF(x), n / 8, workspace_1024);
    VLB_LAUNCH_CHECK();
    sumsq_final_kernel<<<1, 256, 0, s>>>(workspace_1024, grid, out, accumulate);
    count_launch(2);
    VLB_LAUNCH_CHECK();
    return VLB200_OK;
}
extern "C" int vlb200_adamw(void* param_bf16, const void* grad_bf16, float* master, float* exp_avg, float* exp_avg_sq,
                            uint64_t n, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                            float grad_scale, const float* grad_sumsq, float max_grad_norm, void* stream) {
    VLB_REQUIRE(param_bf16 && grad_bf16 && master && exp_avg && exp_avg_sq, "adamw: null pointer");
    VLB_REQUIRE(n % 8 == 0 && step >= 1, "adamw: n must be a multiple of 8 and step >= 1");
    if (n == 0) return VLB200_OK;
    const float bc1 = 1.f - powf(beta1, (float)step);
    const float bc2s = sqrtf(1.f - powf(beta2, (float)step));
    adamw_kernel<<<grid_for(n / 8, 256), 256, 0, as_stream(stream)>>>(BF(param_bf16), CBF(grad_bf16), master, exp_avg, exp_avg_sq, n / 8, lr, beta1, beta2, eps, weight_decay, bc1, bc2s, grad_scale, grad_sumsq, max_grad_norm);
    count_launch();
    VLB_LAUNCH_CHECK();
    return VLB200_OK;
}
extern "C" int vlb200_cast_f32_to_bf16(const float* src, void* dst, uint64_t n, float scale, void* stream) {
    VLB_REQUIRE(src && dst, "cast: null pointer");
    if (n == 0) return VLB200_OK;
    cast_f32_bf16_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(src, BF(dst), n, scale);
    count_launch();
    VLB_LAUNCH_CHECK();
    return VLB200_OK;
}
extern "C" int vlb200_cast_bf16_to_f32(const void* src, float* dst, uint64_t n, void* stream) {
    VLB_REQUIRE(src && dst, "cast: null pointer");
    if (n == 0) return VLB200_OK;
    cast_bf16_f32_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(CBF(src), dst, n);
    count_launch();
    VLB_LAUNCH_CHECK();
    return VLB200_OK;
}
