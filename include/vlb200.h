/* libvlb200 -- C ABI of the B200-native VL-DPO hot path.
 *
 * The reference (TideDra/VL-RLHF) has no FFI of its own: its hot path is Python calling
 * transformers/trl/torch.  The entry points below are what a binding for that path needs:
 * plain device pointers + sizes + a cudaStream_t (as void*), no torch types.  Each block cites
 * the reference code it replaces (paths relative to the reference repo root, or the
 * transformers module the reference delegates to).  INTEGRATION.md shows the ctypes stub and
 * the ModelCoreMapper plug-in a maintainer would add.
 *
 * Conventions
 *   - every function returns 0 (VLB200_OK) or a VLB200_ERR_* code; vlb200_last_error() gives
 *     the message (thread-local).  Asynchronous CUDA faults surface at the next sync.
 *   - all pointers are DEVICE pointers unless the parameter name ends in _host.
 *   - work is enqueued on `stream` (cudaStream_t); nothing synchronises unless stated.
 *   - bf16 = __nv_bfloat16 storage; accumulation is always fp32.
 *   - no hidden allocations after the first call of a given shape (TMA descriptor cache only).
 */
#ifndef VLB200_H
#define VLB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VLB200_OK 0
#define VLB200_ERR_INVALID 1 /* bad argument (maps to Python ValueError, like base/trainer.py:157-158) */
#define VLB200_ERR_CUDA 2    /* CUDA runtime/driver failure (maps to RuntimeError) */
#define VLB200_ERR_UNSUPPORTED 3

#define VLB200_ABI_VERSION 1

int vlb200_abi_version(void);
const char* vlb200_last_error(void);
/* number of kernels this library has launched since load (bench.py's "gpu_launches") */
uint64_t vlb200_launch_count(void);
/* Which generation of the attention forward kernel vlb200_attn_fwd_tc* launches: 5 (default; environment VLB200_ATTN_FWD_VARIANT)
 * = eight softmax warps, 4 = four, 1 / 0 / 2 = first generation (P in tensor memory / through shared memory / P and Q in
 * tensor memory).  All give the same results up to accumulation order; returns the previous value, -1 = query only. */
int vlb200_set_attn_fwd_variant(int variant);

/* ---- dtype tags ---------------------------------------------------------------------- */
#define VLB200_BF16 0
#define VLB200_F32 1

/* ---- deterministic synthetic init (twin of oracle/restate.py::hash_uniform) -----------
 * dst[i] = bf16( shift + scale * (2 * (lowbias32(i ^ lowbias32(seed)) >> 8) * 2^-24 - 1) )   */
int vlb200_init_uniform(void* dst, int dtype, uint64_t n, uint32_t seed, float scale, float shift, void* stream);
/* dst = bf16( base + alpha * (other - shift) )  (reference-model perturbation, oracle make_policy_and_ref) */
int vlb200_perturb_bf16(void* dst, const void* base, const void* other, uint64_t n, float alpha, float shift,
                        void* stream);

/* ---- GEMM (tcgen05 / TMEM / TMA) --------------------------------------------------------
 * Replaces every nn.Linear / Conv2d-as-GEMM the reference reaches through transformers:
 * CLIP q/k/v/out/fc1/fc2 (modeling_clip.py), LlavaMultiModalProjector (modeling_llava.py:87-107),
 * Llama q/k/v/o/gate/up/down/lm_head (modeling_llama.py), and their autograd backward.
 *
 *   D[M,N] = epilogue( sum_k opA(A)[m,k] * opB(B)[n,k] )
 *   opA: a_kmajor=1 -> A is row-major [M,K] (lda >= K);  a_kmajor=0 -> A is row-major [K,M] (lda >= M)
 *   opB: b_kmajor=1 -> B is row-major [N,K] (ldb >= K)   (nn.Linear weight layout: y = x W^T)
 *        b_kmajor=0 -> B is row-major [K,N] (ldb >= N)
 *   epilogue: + bias[n] (bf16 or NULL), activation, + residual[m,n] (bf16 or fp32 per residual_dtype, ldr, or NULL),
 *             accumulate=1 adds the previous contents of D (gradient accumulation).
 *   out_dtype: VLB200_BF16 or VLB200_F32.   All leading dimensions in elements; pointers 16-byte
 *   aligned; lda/ldb multiples of 8.                                                         */
#define VLB200_ACT_NONE 0
#define VLB200_ACT_QUICK_GELU 1 /* x * sigmoid(1.702 x)   (CLIP MLP) */
#define VLB200_ACT_GELU_ERF 2   /* projector */
int vlb200_gemm_bf16(const void* A, int lda, int a_kmajor, const void* B, int ldb, int b_kmajor, void* D, int ldd,
                     int out_dtype, int M, int N, int K, const void* bias, int act, const void* residual, int residual_dtype,
                     int ldr, int accumulate, void* stream);

/* Same GEMM with (a) a second operand pair contracted into the same accumulator,
 *   D = epilogue( alpha * ( opA(A) opB(B)^T + opA(A2) opB(B2)^T ) ),   A2: [M,K2] / [K2,M], B2: [N,K2] / [K2,N]
 * in the majorness of A / B (A2 = B2 = NULL, K2 = 0: none), and (b) a scale on the accumulator.  This is the LoRA linear
 * of peft (`base(x) + lora_B(lora_A(x)) * scaling`; the reference trains with it, utils/auto_load.py:559-578) in ONE
 * launch: y = x W^T + ts B^T with ts = bf16(scaling * x A^T), and its input gradient dx = dy W + dt A.            */
int vlb200_gemm_bf16_ex(const void* A, int lda, int a_kmajor, const void* B, int ldb, int b_kmajor, const void* A2, int lda2,
                        const void* B2, int ldb2, int K2, void* D, int ldd, int out_dtype, int M, int N, int K, float alpha,
                        const void* bias, int act, const void* residual, int residual_dtype, int ldr, int accumulate,
                        void* stream);

/* Llama/Mistral MLP front half in one launch (modeling_llama.py:182-184): act[M,ff] = silu(A Wg^T) * (A Wu^T) with
 * Wgu = [Wg; Wu] row-major [2*ff, K] (nn.Linear layout).  The CTA-pair kernel puts 128 gate columns and the matching 128 up
 * columns into one accumulator (CTA 0 loads the gate rows of a tile, CTA 1 the up rows) and applies SwiGLU in the epilogue
 * on the bf16-rounded projections -- bit-identical to vlb200_gemm_bf16 into gu followed by vlb200_swiglu_fwd.  gu [M, 2*ff]
 * receives the bf16 gate|up projections when write_gu != 0 (needed by vlb200_swiglu_bwd) and is otherwise scratch that is
 * only touched for shapes the pair kernel does not take (M < 256 or ff % 128 != 0: GEMM + elementwise kernel).          */
int vlb200_gemm_swiglu_bf16(const void* A, int lda, const void* Wgu, int ldb, void* gu, int ld_gu, int write_gu, void* act,
                            int ld_act, int M, int ff, int K, void* stream);

/* mode 1: 256x256 tiles on CTA pairs (tcgen05.mma.cta_group::2, B tile split across the pair) where M,N >= 256;
 * mode 0: single-CTA 128x256 tiles.  Default 1; the environment variable VLB200_GEMM_2CTA=0 selects mode 0.   */
int vlb200_set_gemm_mode(int mode);

/* Tile rasterisation budget in MB: the group of operand panels the concurrently running tiles keep L2-resident while the
 * other operand streams past (default 32, or the environment variable VLB200_RASTER_MB; mb <= 0 restores that).  A tuning
 * knob: results do not depend on it.                                                                                 */
int vlb200_set_gemm_raster_mb(double mb);

/* Rasterisation policy of the CTA-pair kernel.  0 (default): the L2-budget rule above.  1 (or the environment variable
 * VLB200_RASTER_POLICY=model): per distinct launch shape, the (orientation, group) that reads the least DRAM under an LRU
 * model of L2 replaying the 74-pair tile schedule (effective capacity VLB200_L2_MODEL_MB, default 60; fitted to the ncu
 * DRAM reads of profiles/r1d_gemm_dram_traffic.json -- tests/raster_model.py, profiles/r1d_raster_model.md); cached per
 * shape.  < 0 restores the environment / default.  A tuning knob: results do not depend on it.
 * vlb200_gemm_plan_raster: host-only query (no GPU needed) of what a policy picks for an [M,K] x [N,K] launch writing
 * out_bytes (2|4) per element, and the DRAM read bytes the model predicts for that choice.                              */
int vlb200_set_gemm_raster_policy(int policy);
int vlb200_gemm_plan_raster(int M, int N, int K, int out_bytes, int policy, int* group, int* along_n, double* model_read_bytes);
/* Host-side evaluation of the kernels' tile rasterisation (tile index -> block coordinates) for a given group / orientation:
 * lets a CPU test check that every (group, orientation) the policies can pick visits each output tile exactly once.        */
int vlb200_gemm_tile_coords(int num_m_blocks, int num_n_blocks, int group, int along_n, int tile, int* m_blk, int* n_blk);

/* ---- log-prob gather (K16) -- base/trainer.py:148-188 VLDPOTrainer.get_batch_logps -------
 * logits: [rows, V] (dtype bf16|f32, row stride ld_logits elements).  Row r predicts target[r];
 * target[r] < 0 (label_pad) rows are skipped without being read.  rows = n_seq * rows_per_seq
 * (row r belongs to sequence r / rows_per_seq).
 * weight[r] (u8, optional) is the DDPO shared-token mask (trainer.py:169-184): 0 drops the row.
 * Outputs: per_token_logp[r] (f32, 0 for skipped rows), lse[r] (f32), logps[n_seq] (f32; sum, or
 * mean over counted rows when average_log_prob).  Deterministic (two-stage reduction).        */
int vlb200_logps_fwd(const void* logits, int logits_dtype, int64_t ld_logits, const int64_t* target,
                     const uint8_t* weight, int rows, int rows_per_seq, int n_seq, int V, int average_log_prob,
                     float* per_token_logp, float* lse, float* logps, void* stream);
/* dlogits[r, v] = g[seq(r)] * w[r] * (onehot(target[r]) - softmax(logits[r])) (bf16 out, may alias
 * bf16 logits).  Skipped rows are written as zeros.                                          */
int vlb200_logps_bwd(const void* logits, int logits_dtype, int64_t ld_logits, const int64_t* target,
                     const uint8_t* weight, const float* lse, const float* grad_logps, int rows, int rows_per_seq,
                     int n_seq, int V, int average_log_prob, void* dlogits, int64_t ld_dlogits, void* stream);

/* ---- preference loss (K18) -- base/trainer.py:244-301 VLDPOTrainer.dpo_loss -------------
 * loss_type: sigmoid / hinge / ipo / kto_pair / ddpo (ddpo == sigmoid on DDPO-masked logps).
 * Inputs policy_logps / ref_logps are [2*n_pairs] (chosen first, like concatenated_forward).
 * losses has n_pairs entries (2*n_pairs for kto_pair, chosen half then rejected half).
 * stats[0..5] = mean loss, reward accuracy, mean chosen reward, mean rejected reward, mean margin, n_losses
 * grad_policy_logps (optional) = d(mean(losses) * loss_scale)/d policy_logps.                */
#define VLB200_LOSS_SIGMOID 0
#define VLB200_LOSS_HINGE 1
#define VLB200_LOSS_IPO 2
#define VLB200_LOSS_KTO_PAIR 3
#define VLB200_LOSS_DDPO 4
int vlb200_dpo_loss(const float* policy_logps, const float* ref_logps, int n_pairs, float beta,
                    float label_smoothing, int loss_type, int reference_free, float loss_scale, float* losses,
                    float* chosen_rewards, float* rejected_rewards, float* stats, float* grad_policy_logps,
                    void* stream);

/* ---- norms ----------------------------------------------------------------------------
 * RMSNorm = LlamaRMSNorm (transformers modeling_llama.py:53-67); LayerNorm = CLIP pre/layer norms
 * (modeling_clip.py).  y,w,b bf16; x is bf16 or fp32 (x_dtype: the residual stream is kept in fp32); statistics fp32.
 * cols % 8 == 0, cols <= 8192.                                                                 */
int vlb200_rmsnorm_fwd(const void* x, int x_dtype, int64_t ldx, const void* w, void* y, int64_t ldy, float* rstd, int rows,
                       int cols, float eps, void* stream);
/* dx = d(rmsnorm)/dx (+ dres), dw (+)= sum_rows dy * xhat.  x/dy/dx/dres contiguous [rows, cols].
 * workspace: vlb200_norm_bwd_workspace_floats(cols) floats.                                    */
int vlb200_norm_bwd_workspace_floats(int cols);
int vlb200_rmsnorm_bwd(const void* dy, const void* x, int x_dtype, const void* w, const float* rstd, const void* dres, void* dx,
                       void* dw, int dw_accumulate, float* workspace, int rows, int cols, void* stream);
int vlb200_layernorm_fwd(const void* x, int x_dtype, int64_t ldx, const void* w, const void* b, void* y, int64_t ldy, int rows,
                         int cols, float eps, void* stream);
/* out[c] (bf16) (+)= sum_r a[r, c]  (bias gradients).  workspace as for rmsnorm_bwd.            */
int vlb200_colsum(const void* a, int64_t lda, int rows, int cols, void* out, int accumulate, float* workspace,
                  void* stream);

/* fp32 column sums and a fp32 dot product: TRL's `logits/chosen|rejected` metric is the mean of the full [B,S,V] logits
 * tensor, which equals dot(colsum(h_rows), colsum(W_lm)) / (B*S*V) -- no logits needed.          */
int vlb200_colsum_f32(const void* a, int64_t lda, int rows, int cols, float* out, float* workspace, void* stream);
int vlb200_dot_f32(const float* a, const float* b, int n, float scale, float* out, void* stream);

/* ---- RoPE / SwiGLU / GELU (modeling_llama.py:146-184, modeling_llava.py:87-107) -----------
 * rope: rotate-half in place on the first n_rot_heads heads (q heads then k heads) of each row of qkv;
 * pos[rows] int32 indexes cos/sin tables [table_rows, head_dim/2] fp32 (positions are clamped to the table: the host
 * grows the tables before a longer sequence is run, engine.ensure_rope_len); inverse=1 applies the transpose.   */
int vlb200_rope(void* qkv, int64_t ld, const int* pos, const float* cos_table, const float* sin_table, int table_rows,
                int rows, int n_rot_heads, int head_dim, int inverse, void* stream);
/* dact = dy Wd^T-free form: dact[M, ff] = dy[M, K] Wd[K, ff] (+ A2 B2, the LoRA term of the input gradient) with the SwiGLU
 * BACKWARD in the epilogue (modeling_llama.py:182-184 under autograd): the [M, ff] product never reaches HBM.  gate_up
 * [M, 2*ff] is read and overwritten IN PLACE with [dgate | dup]; act (optional, [M, ff]) receives silu(gate) * up recomputed
 * from gate_up -- what the down projection's weight gradient needs.  Bit-identical to vlb200_gemm_bf16_ex ->
 * vlb200_swiglu_fwd + vlb200_swiglu_bwd.  Wd is down_proj.weight [K = d_model, ff] row-major (an MN-major B operand).     */
int vlb200_gemm_swiglu_bwd_bf16(const void* dy, int ld_dy, const void* Wd, int ld_wd, const void* A2, int lda2,
                                const void* B2, int ldb2, int K2, void* gate_up, int ld_gu, void* act, int ld_act, int M,
                                int ff, int K, void* stream);
/* gate_up = [gate | up] per row (2*ff columns); act = silu(gate) * up                           */
int vlb200_swiglu_fwd(const void* gate_up, int64_t ld_gu, void* act, int64_t ld_act, int rows, int ff, void* stream);
int vlb200_swiglu_bwd(const void* gate_up, int64_t ld_gu, const void* dact, int64_t ld_dact, void* dgate_up,
                      int64_t ld_dgu, int rows, int ff, void* stream);
int vlb200_gelu_fwd(const void* z, void* h, uint64_t n, void* stream);
int vlb200_gelu_bwd(const void* z, const void* dh, void* dz, uint64_t n, void* stream);

/* ---- CLIP patch embedding helpers (modeling_clip.py:138-219) -------------------------------
 * im2col: pixels [B,3,H,W] (f32|bf16) -> patches [B*(H/p)*(W/p), ld_out] bf16, column = c*p*p + ky*p + kx.
 * cls_rows: x[b*tokens_per_img, :] = class_embedding + position_embedding[0].                      */
int vlb200_clip_im2col(const void* pixels, int pixel_dtype, void* patches, int64_t ld_out, int batch, int height,
                       int width, int patch, void* stream);
int vlb200_clip_cls_rows(void* x, const void* cls, const void* pos0, int batch, int tokens_per_img, int d, void* stream);

/* ---- row movers (bf16, 16-byte granularity) -------------------------------------------------- */
int vlb200_copy_rows(const void* src, int64_t src_group_stride, int64_t src_row_stride, int src_row0, void* dst,
                     int64_t dst_group_stride, int64_t dst_row_stride, int groups, int rows_per_group, int cols,
                     void* stream);
int vlb200_gather_rows(const void* src, int64_t ld_src, const int* index, void* dst, int64_t ld_dst, int n, int cols,
                       void* stream);
int vlb200_scatter_rows(const void* src, int64_t ld_src, const int* index, void* dst, int64_t ld_dst, int n, int cols,
                        void* stream);
/* dst[index[i], :] += scale * src[i, :] (index unique, < 0 skipped; dst bf16 or fp32): InternLM-XComposer2's partial LoRA
 * `res[im_mask] += Plora_B(Plora_A(x[im_mask])) * scaling` (models/InternLMXC2/build_mlp.py:194-203) on gathered image rows */
int vlb200_scatter_add_rows(const void* src, int64_t ld_src, const int* index, void* dst, int dst_dtype, int64_t ld_dst, int n,
                            int cols, float scale, void* stream);
int vlb200_memset_zero(void* dst, uint64_t bytes, void* stream);

/* ---- LLaVA text/image merge (K7+K8) -- models/Llava/__init__.py:36-109 ---------------------
 * merge_index: integer pass (one thread per sequence).  Outputs, all device:
 *   src_map[n_seq*S] (>=0 embed row, -1-k image feature row k, INT_MIN zero row), labels_merged[n_seq*S] i64,
 *   mask_merged / position_ids [n_seq*S] i32, seqlens[n_seq], img_pos[n_seq*imgs_per_seq*P],
 *   row_of_text[n_seq*(L-1)] (flat merged row whose logits predict text token j), target[n_seq*(L-1)] i64,
 *   status: 0 ok, 1 too many image tokens, 2 ragged image counts, 3 attention mask is not a prefix.
 * Sequence b uses images (b % n_img_batch)*imgs_per_seq ... (the reference duplicates pixel_values [v, v]).  */
int vlb200_llava_merge_index(const int64_t* input_ids, const int64_t* attention_mask, const int64_t* labels, int n_seq,
                             int text_len, int merged_len, int n_patches, int n_img_batch, int imgs_per_seq,
                             int image_token, int pad_token, int ignore_index, int* src_map, int64_t* labels_merged,
                             int* mask_merged, int* position_ids, int* seqlens, int* img_pos, int* row_of_text,
                             int64_t* target, int* status, void* stream);
int vlb200_llava_merge_embed(const int* src_map, const void* embed_tokens, const void* image_features, void* out,
                             int out_dtype, int rows, int d, void* stream);
/* dembed_f32[V,d] += text-row grads (fp32 atomics); dimage_features = sum over sequences sharing the image */
int vlb200_llava_merge_bwd(const int* src_map, const int* img_pos, const void* dx, float* dembed_f32,
                           void* dimage_features, int n_seq, int n_img_batch, int merged_len, int feats_per_seq, int d,
                           void* stream);
/* The same over n_rows rows of dx with sequence b's image position p at row b*row_stride + p; row_stride = 0: img_pos holds
 * absolute rows (packed rows, after vlb200_pack_merge_rows). */
int vlb200_llava_merge_bwd_rows(const int* src_map, const int* img_pos, const void* dx, float* dembed_f32,
                                void* dimage_features, int64_t n_rows, int n_seq, int n_img_batch, int row_stride,
                                int feats_per_seq, int d, void* stream);

/* Packed rows for the merge indices above: keep the first len[b] = row_starts[b+1] - row_starts[b] rows of
 * every sequence and stack the sequences back to back.  src_map_packed / position_ids_packed [row_starts[n_seq]] receive
 * the surviving rows; the row lists row_of_text [n_text_rows] and img_rows [n_img_rows] (flat padded rows b*merged_len + p)
 * are rewritten IN PLACE to packed rows, -1 where p >= len[b] (vlb200_gather_rows then yields a zero row and
 * vlb200_scatter_rows skips it); img_pos [n_img_pos] (position inside sequence i / feats_per_seq) becomes an absolute packed
 * row, to be consumed through vlb200_llava_merge_bwd_rows with row_stride = 0.  NULL lists are skipped.  Right padding only (the merge kernels enforce it).  */
int vlb200_pack_merge_rows(const int* src_map, const int* position_ids, const int* row_starts, int n_seq, int merged_len,
                           int* src_map_packed, int* position_ids_packed, int* row_of_text, int64_t n_text_rows,
                           int* img_pos, int64_t n_img_pos, int feats_per_seq, int* img_rows, int64_t n_img_rows,
                           void* stream);

/* Shared-prefix rows (base/trainer.py:124-146 concatenates chosen and rejected on the batch axis, so the
 * reference computes the prompt + image prefix of every pair twice per pass): sequence b (chosen b < n_seq/2, rejected of
 * the same pair at b + n_seq/2) keeps its first prefix_rows[b] merged rows in ONE copy per pair at prefix_starts[b] (both
 * arrays hold equal values for the two sequences of a pair) and its remaining seq_lens[b] - prefix_rows[b] rows at
 * suffix_starts[b].  Same outputs and in-place rewrites as vlb200_pack_merge_rows; additionally the rejected copy of an
 * img_pos entry that lies in the shared prefix becomes -1 (the row is counted once by vlb200_llava_merge_bwd_rows);
 * row_of_text of both sequences points at the shared row (the head gathers it twice, its gradient is scatter-ADDED).
 * Consumed with vlb200_attn_{fwd,bwd}_tc_ctx: attention sequences = {prefixes, chosen suffixes, rejected suffixes}.       */
int vlb200_share_prefix_rows(const int* src_map, const int* position_ids, const int* prefix_rows, const int* prefix_starts,
                             const int* suffix_starts, const int* seq_lens, int n_seq, int merged_len, int64_t total_rows,
                             int* src_map_shared, int* position_ids_shared, int* row_of_text, int64_t n_text_rows,
                             int* img_pos, int64_t n_img_pos, int feats_per_seq, int* img_rows, int64_t n_img_rows,
                             void* stream);

/* ---- LLaVA-Next text/image merge -- models/LlavaNext/__init__.py:38-171 (_merge_input_ids_with_image_features)
 * Differences from the LLaVA-1.5 merge above: image k contributes feat_off[k+1]-feat_off[k] PACKED feature rows
 * (anyres "spatial_unpad" + image_newline, modeling_llava_next.py pack_image_features; the row gather that builds
 * them is vlb200_gather_rows over the index vlrlhf_b200/host.py:anyres_pack_index computes), tokens whose
 * attention_mask is 0 are never written, merged_len is the longest valid merged sequence (host-computed:
 * (mask==1).sum() - n_image_tokens + sum(feature_lens), :83-87) and pad-token embeddings are kept.  Right padding
 * only (status 3 otherwise).  img_rows[rep*total_feats + k] = flat merged row of packed feature row k in the rep-th
 * sequence sharing the image (rep = seq / n_img_batch: chosen, rejected).  status: 0 ok, 1 overflow of merged_len,
 * 2 wrong number of <image> tokens (:75-79), 3 not right-padded, 4 masked <image> token. */
int vlb200_llavanext_merge_index(const int64_t* input_ids, const int64_t* attention_mask, const int64_t* labels,
                                 const int* feat_off, int n_seq, int text_len, int merged_len, int n_img_batch,
                                 int imgs_per_seq, int total_feats, int image_token, int ignore_index, int* src_map,
                                 int64_t* labels_merged, int* mask_merged, int* position_ids, int* seqlens, int* img_rows,
                                 int* row_of_text, int64_t* target, int* status, void* stream);
/* backward of the merge: dembed_f32[id] += dx[row] for text rows; dimage_features[k] = sum_rep dx[img_rows[rep, k]] */
int vlb200_llavanext_merge_bwd(const int* src_map, const int* img_rows, const void* dx, float* dembed_f32,
                               void* dimage_features, int n_rows, int total_feats, int reps, int d, void* stream);

/* ---- Qwen-VL image placement -- models/QwenVL/modeling_qwen.py:524-528 (span scan) and :614-621 (overwrite)
 * The text already holds n_queries placeholder tokens between <img> (image_start_id) and </img> (image_start_id+1);
 * those positions read image feature rows (src_map = -1 - (image*n_queries + q)), every other position its own
 * embedding row.  merged length == text_len, labels pass through (the model returns none, base/trainer.py:225-229),
 * position ids = arange (:573-580).  status: 0 ok, 2 malformed / wrong number of <img> spans, 3 not right-padded. */
int vlb200_qwen_merge_index(const int64_t* input_ids, const int64_t* attention_mask, const int64_t* labels, int n_seq,
                            int text_len, int n_queries, int n_img_batch, int imgs_per_seq, int image_start_id,
                            int ignore_index, int* src_map, int64_t* labels_out, int* mask_out, int* position_ids,
                            int* seqlens, int* row_of_text, int64_t* target, int* status, void* stream);

/* ---- attention (K4, K12) -- CLIPAttention (modeling_clip.py:261-334), LlamaAttention
 * (modeling_llama.py:199-290).  q/k/v/out rows are tokens (row = b*S + t), head h at column h*head_dim.
 * causal + key-padding via seqlens[B] (attended prefix length; NULL = S).  lse/delta: [B,H,S] f32.
 * head_dim 64 or 128; H % KVH == 0 (GQA).                                                          */
/* tcgen05/TMEM/TMA forward (S and O accumulate in TMEM, K/V tiles arrive by TMA) */
int vlb200_attn_fwd_tc(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* out,
                       int64_t ldo, float* lse, const int* seqlens, int B, int S, int H, int KVH, int head_dim, int causal,
                       float scale, void* stream);
/* delta[b,h,t] = sum_d dout[t,h,d] * out[t,h,d]  (softmax-backward row statistic) */
int vlb200_attn_delta(const void* out, int64_t ldo, const void* dout, int64_t lddo, float* delta, int B, int S, int H,
                      int head_dim, void* stream);
/* tcgen05/TMEM/TMA backward: pass 1 dK,dV (key tile stationary), pass 2 dQ */
int vlb200_attn_bwd_tc(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* out,
                       int64_t ldo, const void* dout, int64_t lddo, const float* lse, float* delta, void* dq, int64_t lddq,
                       void* dk, int64_t lddk, void* dv, int64_t lddv, const int* seqlens, int B, int S, int H, int KVH,
                       int head_dim, int causal, float scale, void* stream);

/* Ragged ("packed") rows -- SURVEY.md f-2, the var-len FlashAttention form: sequence b occupies rows
 * [row_starts[b], row_starts[b] + seqlens[b]) of q/k/v/out (total_rows rows in all) instead of [b*S, (b+1)*S), so the
 * padding the DPO collator adds (base/collator.py:44-60) costs no GEMM, norm or attention-store work.  S stays the upper
 * bound of seqlens and the per-sequence stride of lse/delta [B, H, S]; rows at or beyond seqlens[b] are neither read as
 * statistics nor written.  row_starts == NULL: identical to the functions above.                                     */
int vlb200_attn_fwd_tc_varlen(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* out,
                              int64_t ldo, float* lse, const int* seqlens, const int* row_starts, int64_t total_rows, int B,
                              int S, int H, int KVH, int head_dim, int causal, float scale, void* stream);
/* Context sequences ("tree" attention with one shared trunk): ctx[b] >= 0 names a sequence whose rows are additional keys /
 * values visible to EVERY query of sequence b, logically before b's own keys (b's own keys stay causal).  kids[2*b],
 * kids[2*b+1] (-1: none) list the sequences that name b as their context -- the backward of b's keys gathers their queries.
 * With {prefix, chosen suffix, rejected suffix} as three sequences per pair this computes exactly the attention of the two
 * full sequences while the shared prefix is stored, projected and attended once.  Packed rows + causal only.  lse / delta:
 * [B, H, S] indexed by (sequence, head, row inside the sequence).  ctx == NULL: identical to the _varlen entry points. */
int vlb200_attn_fwd_tc_ctx(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* out,
                           int64_t ldo, float* lse, const int* seqlens, const int* row_starts, const int* ctx,
                           int64_t total_rows, int B, int S, int H, int KVH, int head_dim, int causal, float scale, void* stream);
int vlb200_attn_bwd_tc_ctx(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* out,
                           int64_t ldo, const void* dout, int64_t lddo, const float* lse, float* delta, void* dq, int64_t lddq,
                           void* dk, int64_t lddk, void* dv, int64_t lddv, const int* seqlens, const int* row_starts,
                           const int* ctx, const int* kids, int64_t total_rows, int B, int S, int H, int KVH, int head_dim,
                           int causal, float scale, void* stream);
int vlb200_attn_delta_varlen(const void* out, int64_t ldo, const void* dout, int64_t lddo, float* delta, const int* row_starts,
                             int64_t total_rows, int B, int S, int H, int head_dim, void* stream);
int vlb200_attn_bwd_tc_varlen(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                              const void* out, int64_t ldo, const void* dout, int64_t lddo, const float* lse, float* delta,
                              void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv, const int* seqlens,
                              const int* row_starts, int64_t total_rows, int B, int S, int H, int KVH, int head_dim, int causal,
                              float scale, void* stream);

/* ---- optimizer (K24: torch.optim.AdamW semantics, HF Trainer max_grad_norm clipping) ---------
 * sumsq: out[0] (+)= sum x^2 (deterministic two-stage; workspace 1024 floats).
 * adamw: bf16 param/grad, fp32 master + moments; grads are multiplied by grad_scale and by the clip
 * coefficient min(1, max_grad_norm / (sqrt(grad_sumsq[0])*grad_scale + 1e-6)) when grad_sumsq != NULL.  */
int vlb200_sumsq_bf16(const void* x, uint64_t n, float* workspace_1024, float* out, int accumulate, void* stream);
int vlb200_adamw(void* param_bf16, const void* grad_bf16, float* master, float* exp_avg, float* exp_avg_sq, uint64_t n,
                 float lr, float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                 const float* grad_sumsq, float max_grad_norm, void* stream);
int vlb200_cast_f32_to_bf16(const float* src, void* dst, uint64_t n, float scale, void* stream);
int vlb200_cast_bf16_to_f32(const void* src, float* dst, uint64_t n, void* stream);

/* ---- input pipeline: CLIP image preprocessing of the DPO collator on the GPU -----------------
 * Replaces `self.processor.image_processor(images=imgs, return_tensors="pt")` in
 * LlavaDPODataCollatorWithPadding.__call__ (models/Llava/__init__.py:435-443) = transformers-4.41
 * CLIPImageProcessor.preprocess: Pillow Image.resize(BICUBIC) to shortest edge -> center crop -> x*(1/255) in float64
 * stored float32 -> (x-mean)/std in float32 -> CHW.  image: device uint8 [in_h, in_w, 3] RGB (the decoded file).
 * coef_h/bounds_h ([new_w, ksize_h] int32, [new_w, 2] int32 = first tap, tap count) and coef_v/bounds_v
 * ([new_h, ksize_v], [new_h, 2]) are Pillow's 22-bit fixed-point resampling tables (Resample.c precompute_coeffs +
 * normalize_coeffs_8bpc), device-resident, built by vlrlhf_b200/preprocess.py.  (top,left,crop_h,crop_w) is the output
 * window in resized-image coordinates: the center crop for CLIP, one 336x336 cell of the zero-padded anyres canvas for
 * LLaVA-Next (LlavaNextImageProcessor get_image_patches: top/left may be negative or reach past the resized image; pixels
 * outside it are the canvas' uint8 zeros).  Only input rows [row0, row0+rows) feed the window (rows may be 0).  workspace holds the
 * uint8 result of the horizontal pass (rows*crop_w*3 bytes).  mean_std_host: 6 floats on the HOST (mean RGB, std RGB).
 * out: [3, crop_h, crop_w] VLB200_F32 (what the reference's collator emits) or VLB200_BF16 (its rounding).
 * Bit-exact with Pillow + numpy for uint8 input. */
size_t vlb200_clip_preprocess_workspace_bytes(int in_h, int crop_w);
int vlb200_clip_preprocess_u8(const uint8_t* image, int in_h, int in_w, const int* coef_h, const int* bounds_h, int ksize_h,
                              const int* coef_v, const int* bounds_v, int ksize_v, int new_h, int new_w, int top, int left,
                              int crop_h, int crop_w, int row0, int rows, uint8_t* workspace, size_t workspace_bytes,
                              double rescale, const float* mean_std_host, void* out, int out_dtype, void* stream);

/* ---- DDPO token mask on the HOST (CPU integer work; no GPU, no launch) -------------------------------------------
 * The mask_shared_tokens branch of VLDPOTrainer.get_batch_logps (base/trainer.py:169-184) over
 * utils/diff_lib.get_diff_ids (utils/diff_lib.py:116-180), which calls Python difflib.SequenceMatcher
 * (autojunk on).  All pointers are HOST pointers.
 * host_matching_blocks: SequenceMatcher(None, a, b).get_matching_blocks() -> triples[3*t] = (i, j, size), the
 *   (len(a), len(b), 0) sentinel included; returns the block count, or -1 (see vlb200_last_error).
 * host_ddpo_row_weights: input_ids / attention_mask (NULL = LLaVA-1.5 merge: padded tokens stay in the sequence) /
 *   labels are the concatenated batch [n_seq, text_len] (chosen rows, then rejected rows); feat_len_per_seq[b] = image
 *   feature rows merged into sequence b; merged_len >= 0 pads every merged label sequence with ignore labels to that
 *   length (LLaVA-Next: LlavaNext/__init__.py:96-127), -1 = no padding.  weights[b, j-1] = 1 iff text token j of
 *   sequence b lies in a span modified on both sides (the tokens DDPO keeps). */
int vlb200_host_matching_blocks(const int64_t* a, int na, const int64_t* b, int nb, int* triples, int max_blocks);
int vlb200_host_ddpo_row_weights(const int64_t* input_ids, const int64_t* attention_mask, const int64_t* labels, int n_seq,
                                 int text_len, int image_token, const int* feat_len_per_seq, int merged_len,
                                 int64_t label_pad_token_id, int min_match_size, uint8_t* weights);

#ifdef __cplusplus
}
#endif
#endif /* VLB200_H */
