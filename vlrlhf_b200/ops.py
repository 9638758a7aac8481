"""Tensor-level wrappers over the C ABI.  torch is only the memory/stream plumbing here: every
function hands raw device pointers of CUDA tensors to libvlb200 on the current torch stream."""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from ._lib import check

BF16, F32 = 0, 1
ACT_NONE, ACT_QUICK_GELU, ACT_GELU_ERF = 0, 1, 2
LOSS_TYPES = {"sigmoid": 0, "hinge": 1, "ipo": 2, "kto_pair": 3, "ddpo": 4}

_L = _lib.load()  # raises ImportError when the extension is missing: no fallback


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    assert t.is_cuda, "vlb200 ops take CUDA tensors only (there is no CPU path)"
    return t.data_ptr()


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.bfloat16:
        return BF16
    if t.dtype == torch.float32:
        return F32
    raise ValueError(f"unsupported dtype {t.dtype}")


def _rowmajor_ld(t: torch.Tensor) -> int:
    assert t.dim() == 2 and t.stride(1) == 1, "expected a row-major 2-D view"
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


def launch_count() -> int:
    return int(_L.vlb200_launch_count())


def set_attn_fwd_variant(variant: int) -> int:
    """Select the attention forward kernel generation (include/vlb200.h); returns the previous one (-1: query only)."""
    return int(_L.vlb200_set_attn_fwd_variant(int(variant)))


def init_uniform_(t: torch.Tensor, seed: int, scale: float, shift: float = 0.0) -> torch.Tensor:
    assert t.is_contiguous()
    check(_L.vlb200_init_uniform(_ptr(t), _dt(t), t.numel(), seed & 0xFFFFFFFF, scale, shift, _stream()))
    return t


def perturb_(dst: torch.Tensor, base: torch.Tensor, other: torch.Tensor, alpha: float, shift: float) -> torch.Tensor:
    assert dst.dtype == base.dtype == other.dtype == torch.bfloat16
    assert dst.is_contiguous() and base.is_contiguous() and other.is_contiguous()
    check(_L.vlb200_perturb_bf16(_ptr(dst), _ptr(base), _ptr(other), dst.numel(), alpha, shift, _stream()))
    return dst


def gemm(a: torch.Tensor, b: torch.Tensor, *, a_kmajor: bool = True, b_kmajor: bool = True,
         out: Optional[torch.Tensor] = None, out_dtype: torch.dtype = torch.bfloat16,
         bias: Optional[torch.Tensor] = None, act: int = ACT_NONE, residual: Optional[torch.Tensor] = None,
         accumulate: bool = False, a2: Optional[torch.Tensor] = None, b2: Optional[torch.Tensor] = None,
         alpha: float = 1.0) -> torch.Tensor:
    """D[M,N] = epilogue(alpha * (opA(a) @ opB(b)^T + opA(a2) @ opB(b2)^T)).

    a: [M,K] if a_kmajor else [K,M];  b: [N,K] if b_kmajor (nn.Linear weight) else [K,N];  the optional second pair
    (a2: [M,K2] / [K2,M], b2: [N,K2] / [K2,N], same majorness) is contracted into the same accumulator -- the LoRA term
    of a peft linear, or of its input gradient, without a round trip through HBM."""
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    M, K = (a.shape[0], a.shape[1]) if a_kmajor else (a.shape[1], a.shape[0])
    N, Kb = (b.shape[0], b.shape[1]) if b_kmajor else (b.shape[1], b.shape[0])
    if K != Kb:
        raise ValueError(f"gemm: contraction mismatch {K} vs {Kb}")
    K2 = 0
    if (a2 is None) != (b2 is None):
        raise ValueError("gemm: a2 and b2 come together")
    if a2 is not None:
        assert a2.dtype == torch.bfloat16 and b2.dtype == torch.bfloat16
        M2, K2 = (a2.shape[0], a2.shape[1]) if a_kmajor else (a2.shape[1], a2.shape[0])
        N2, K2b = (b2.shape[0], b2.shape[1]) if b_kmajor else (b2.shape[1], b2.shape[0])
        if (M2, N2) != (M, N) or K2 != K2b:
            raise ValueError(f"gemm: second operand pair [{M2},{K2}] x [{N2},{K2b}] does not match M={M}, N={N}")
    if out is None:
        out = torch.empty(M, N, dtype=out_dtype, device=a.device)
    assert out.shape == (M, N)
    check(_L.vlb200_gemm_bf16_ex(_ptr(a), _rowmajor_ld(a), int(a_kmajor), _ptr(b), _rowmajor_ld(b), int(b_kmajor),
                                 _ptr(a2), _rowmajor_ld(a2) if a2 is not None else 0, _ptr(b2),
                                 _rowmajor_ld(b2) if b2 is not None else 0, K2,
                                 _ptr(out), _rowmajor_ld(out), _dt(out), M, N, K, float(alpha), _ptr(bias), act, _ptr(residual),
                                 _dt(residual) if residual is not None else BF16,
                                 _rowmajor_ld(residual) if residual is not None else 0, int(accumulate), _stream()))
    return out


def gemm_swiglu(a: torch.Tensor, wgu: torch.Tensor, gu: torch.Tensor, act: torch.Tensor, write_gu: bool = True) -> torch.Tensor:
    """act[M,ff] = silu(a Wg^T) * (a Wu^T), wgu = [Wg; Wu] ([2*ff, K]); gu [M, 2*ff] receives the bf16 gate|up projections
    when write_gu (the backward needs them) and is scratch otherwise.  Bit-identical to gemm(a, wgu, out=gu); swiglu_fwd(gu)."""
    assert a.dtype == wgu.dtype == gu.dtype == act.dtype == torch.bfloat16
    M, K = a.shape
    ff = wgu.shape[0] // 2
    if wgu.shape[1] != K or tuple(gu.shape) != (M, 2 * ff) or tuple(act.shape) != (M, ff):
        raise ValueError("gemm_swiglu: shape mismatch")
    check(_L.vlb200_gemm_swiglu_bf16(_ptr(a), _rowmajor_ld(a), _ptr(wgu), _rowmajor_ld(wgu), _ptr(gu), _rowmajor_ld(gu),
                                     int(write_gu), _ptr(act), _rowmajor_ld(act), M, ff, K, _stream()))
    return act


def gemm_swiglu_bwd(dy: torch.Tensor, wd: torch.Tensor, gu: torch.Tensor, act: Optional[torch.Tensor] = None,
                    a2: Optional[torch.Tensor] = None, b2: Optional[torch.Tensor] = None) -> torch.Tensor:
    """gu <- [dgate | dup] IN PLACE from dact = dy @ wd (+ a2 @ b2) without storing dact; act (optional) <- silu(gate) * up.
    dy [M, K], wd = down_proj.weight [K, ff] (row-major), a2 [M, K2] / b2 [K2, ff]: the LoRA term of the input gradient.
    Bit-identical to gemm(dy, wd, b_kmajor=False) -> swiglu_fwd + swiglu_bwd (include/vlb200.h)."""
    assert dy.dtype == wd.dtype == gu.dtype == torch.bfloat16
    M, K = dy.shape
    ff = wd.shape[1]
    if wd.shape[0] != K or tuple(gu.shape) != (M, 2 * ff) or (act is not None and tuple(act.shape) != (M, ff)):
        raise ValueError("gemm_swiglu_bwd: shape mismatch")
    if (a2 is None) != (b2 is None):
        raise ValueError("gemm_swiglu_bwd: a2 and b2 come together")
    K2 = 0
    if a2 is not None:
        K2 = a2.shape[1]
        if a2.shape[0] != M or tuple(b2.shape) != (K2, ff):
            raise ValueError("gemm_swiglu_bwd: second operand pair does not match")
    check(_L.vlb200_gemm_swiglu_bwd_bf16(_ptr(dy), _rowmajor_ld(dy), _ptr(wd), _rowmajor_ld(wd), _ptr(a2),
                                         _rowmajor_ld(a2) if a2 is not None else 0, _ptr(b2),
                                         _rowmajor_ld(b2) if b2 is not None else 0, K2, _ptr(gu), _rowmajor_ld(gu), _ptr(act),
                                         _rowmajor_ld(act) if act is not None else 0, M, ff, K, _stream()))
    return gu


def set_gemm_raster_mb(mb: float):
    """L2 budget (MB) of the tile rasterisation; <= 0 restores VLB200_RASTER_MB / the default.  Tuning only."""
    check(_L.vlb200_set_gemm_raster_mb(float(mb)))


def set_gemm_raster_policy(policy: int):
    """0: L2-budget raster (default); 1: the (orientation, group) an LRU model of L2 predicts to read least (tuning only,
    include/vlb200.h: vlb200_set_gemm_raster_policy); < 0 restores VLB200_RASTER_POLICY / the default."""
    check(_L.vlb200_set_gemm_raster_policy(int(policy)))


def set_gemm_mode(mode: int):
    """1: CTA-pair (cta_group::2) 256x256 tiles where the shape allows; 0: single-CTA 128x256 tiles."""
    check(_L.vlb200_set_gemm_mode(int(mode)))


def logps_fwd(logits: torch.Tensor, target: torch.Tensor, n_seq: int, weight: Optional[torch.Tensor] = None,
              average_log_prob: bool = False):
    """logits [rows, V] (bf16|f32); target [rows] int64 (<0 = skip).  -> (logps[n_seq], per_token[rows], lse[rows])"""
    rows, V = logits.shape
    if target.numel() != rows or rows % n_seq != 0:
        raise ValueError("Logits (batch and sequence length dim) and labels must have the same shape.")
    assert target.dtype == torch.int64 and target.is_contiguous()
    if weight is not None:
        assert weight.dtype == torch.uint8 and weight.is_contiguous() and weight.numel() == rows
    per_token = torch.empty(rows, dtype=torch.float32, device=logits.device)
    lse = torch.empty(rows, dtype=torch.float32, device=logits.device)
    logps = torch.empty(n_seq, dtype=torch.float32, device=logits.device)
    check(_L.vlb200_logps_fwd(_ptr(logits), _dt(logits), _rowmajor_ld(logits), _ptr(target), _ptr(weight), rows,
                              rows // n_seq, n_seq, V, int(average_log_prob), _ptr(per_token), _ptr(lse), _ptr(logps),
                              _stream()))
    return logps, per_token, lse


def logps_bwd(logits: torch.Tensor, target: torch.Tensor, n_seq: int, lse: torch.Tensor, grad_logps: torch.Tensor,
              weight: Optional[torch.Tensor] = None, average_log_prob: bool = False,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    rows, V = logits.shape
    if out is None:
        out = torch.empty(rows, V, dtype=torch.bfloat16, device=logits.device)
    assert grad_logps.dtype == torch.float32 and grad_logps.numel() == n_seq
    check(_L.vlb200_logps_bwd(_ptr(logits), _dt(logits), _rowmajor_ld(logits), _ptr(target), _ptr(weight), _ptr(lse),
                              _ptr(grad_logps), rows, rows // n_seq, n_seq, V, int(average_log_prob), _ptr(out),
                              _rowmajor_ld(out), _stream()))
    return out


def dpo_loss(policy_logps: torch.Tensor, ref_logps: torch.Tensor, beta: float, label_smoothing: float = 0.0,
             loss_type: str = "sigmoid", reference_free: bool = False, loss_scale: float = 1.0, want_grad: bool = True):
    """policy_logps/ref_logps: f32 [2*n_pairs] (chosen first).
    -> (losses, chosen_rewards, rejected_rewards, stats[6], grad_policy_logps|None)"""
    if loss_type not in LOSS_TYPES:
        raise ValueError(f"Unknown loss type: {loss_type}. Should be one of ['sigmoid', 'hinge', 'ipo', 'kto_pair']")
    n2 = policy_logps.numel()
    assert n2 % 2 == 0 and ref_logps.numel() == n2
    n = n2 // 2
    dev = policy_logps.device
    policy_logps = policy_logps.float().contiguous()
    ref_logps = ref_logps.float().contiguous()
    losses = torch.empty(2 * n if loss_type == "kto_pair" else n, dtype=torch.float32, device=dev)
    cr = torch.empty(n, dtype=torch.float32, device=dev)
    rr = torch.empty(n, dtype=torch.float32, device=dev)
    stats = torch.empty(6, dtype=torch.float32, device=dev)
    grad = torch.empty(n2, dtype=torch.float32, device=dev) if want_grad else None
    check(_L.vlb200_dpo_loss(_ptr(policy_logps), _ptr(ref_logps), n, beta, label_smoothing, LOSS_TYPES[loss_type],
                             int(reference_free), loss_scale, _ptr(losses), _ptr(cr), _ptr(rr), _ptr(stats), _ptr(grad),
                             _stream()))
    return losses, cr, rr, stats, grad


# ------------------------------------------------------------------------------------------
# norms / elementwise / movers
# ------------------------------------------------------------------------------------------
_ws_cache = {}


def _workspace(cols: int, device) -> torch.Tensor:
    n = int(_L.vlb200_norm_bwd_workspace_floats(cols))
    key = (str(device), n)
    if key not in _ws_cache:
        _ws_cache[key] = torch.empty(n, dtype=torch.float32, device=device)
    return _ws_cache[key]


def rmsnorm_fwd(x: torch.Tensor, w: torch.Tensor, eps: float, out: Optional[torch.Tensor] = None,
                rstd: Optional[torch.Tensor] = None):
    rows, cols = x.shape
    if out is None:
        out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    check(_L.vlb200_rmsnorm_fwd(_ptr(x), _dt(x), _rowmajor_ld(x), _ptr(w), _ptr(out), _rowmajor_ld(out), _ptr(rstd), rows, cols,
                                eps, _stream()))
    return out


def rmsnorm_bwd(dy: torch.Tensor, x: torch.Tensor, w: torch.Tensor, rstd: torch.Tensor, dw: torch.Tensor,
                dres: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None, dw_accumulate: bool = False):
    rows, cols = x.shape
    assert dy.is_contiguous() and x.is_contiguous() and (dres is None or dres.is_contiguous())
    if out is None:
        out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    check(_L.vlb200_rmsnorm_bwd(_ptr(dy), _ptr(x), _dt(x), _ptr(w), _ptr(rstd), _ptr(dres), _ptr(out), _ptr(dw),
                                int(dw_accumulate), _ptr(_workspace(cols, x.device)), rows, cols, _stream()))
    return out


def layernorm_fwd(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float, out: Optional[torch.Tensor] = None):
    rows, cols = x.shape
    if out is None:
        out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    check(_L.vlb200_layernorm_fwd(_ptr(x), _dt(x), _rowmajor_ld(x), _ptr(w), _ptr(b), _ptr(out), _rowmajor_ld(out), rows, cols,
                                  eps, _stream()))
    return out


def colsum(a: torch.Tensor, out: torch.Tensor, accumulate: bool = False):
    rows, cols = a.shape
    check(_L.vlb200_colsum(_ptr(a), _rowmajor_ld(a), rows, cols, _ptr(out), int(accumulate),
                           _ptr(_workspace(max(cols, 8), a.device)), _stream()))
    return out


def colsum_f32(a: torch.Tensor, out: torch.Tensor):
    rows, cols = a.shape
    assert out.dtype == torch.float32 and out.numel() == cols
    check(_L.vlb200_colsum_f32(_ptr(a), _rowmajor_ld(a), rows, cols, _ptr(out), _ptr(_workspace(max(cols, 8), a.device)),
                               _stream()))
    return out


def dot_f32(a: torch.Tensor, b: torch.Tensor, scale: float, out: torch.Tensor):
    assert a.dtype == torch.float32 and b.dtype == torch.float32 and a.numel() == b.numel()
    check(_L.vlb200_dot_f32(_ptr(a), _ptr(b), a.numel(), scale, _ptr(out), _stream()))
    return out


def rope_(qkv: torch.Tensor, pos: torch.Tensor, cos_t: torch.Tensor, sin_t: torch.Tensor, n_rot_heads: int, head_dim: int,
          inverse: bool = False):
    assert pos.dtype == torch.int32 and cos_t.dtype == torch.float32 and sin_t.dtype == torch.float32
    check(_L.vlb200_rope(_ptr(qkv), _rowmajor_ld(qkv), _ptr(pos), _ptr(cos_t), _ptr(sin_t), cos_t.shape[0], qkv.shape[0], n_rot_heads,
                         head_dim, int(inverse), _stream()))
    return qkv


def swiglu_fwd(gate_up: torch.Tensor, out: Optional[torch.Tensor] = None):
    rows, ff2 = gate_up.shape
    ff = ff2 // 2
    if out is None:
        out = torch.empty(rows, ff, dtype=torch.bfloat16, device=gate_up.device)
    check(_L.vlb200_swiglu_fwd(_ptr(gate_up), _rowmajor_ld(gate_up), _ptr(out), _rowmajor_ld(out), rows, ff, _stream()))
    return out


def swiglu_bwd(gate_up: torch.Tensor, dact: torch.Tensor, out: Optional[torch.Tensor] = None):
    rows, ff2 = gate_up.shape
    if out is None:
        out = torch.empty_like(gate_up)
    check(_L.vlb200_swiglu_bwd(_ptr(gate_up), _rowmajor_ld(gate_up), _ptr(dact), _rowmajor_ld(dact), _ptr(out),
                               _rowmajor_ld(out), rows, ff2 // 2, _stream()))
    return out


def gelu_fwd(z: torch.Tensor, out: Optional[torch.Tensor] = None):
    assert z.is_contiguous()
    if out is None:
        out = torch.empty_like(z)
    check(_L.vlb200_gelu_fwd(_ptr(z), _ptr(out), z.numel(), _stream()))
    return out


def gelu_bwd(z: torch.Tensor, dh: torch.Tensor, out: Optional[torch.Tensor] = None):
    assert z.is_contiguous() and dh.is_contiguous()
    if out is None:
        out = torch.empty_like(z)
    check(_L.vlb200_gelu_bwd(_ptr(z), _ptr(dh), _ptr(out), z.numel(), _stream()))
    return out


def clip_im2col(pixels: torch.Tensor, patch: int, out: torch.Tensor):
    B, C, H, W = pixels.shape
    assert C == 3 and pixels.is_contiguous()
    check(_L.vlb200_clip_im2col(_ptr(pixels), _dt(pixels), _ptr(out), _rowmajor_ld(out), B, H, W, patch, _stream()))
    return out


def clip_cls_rows_(x: torch.Tensor, cls: torch.Tensor, pos0: torch.Tensor, batch: int, tokens_per_img: int):
    check(_L.vlb200_clip_cls_rows(_ptr(x), _ptr(cls), _ptr(pos0), batch, tokens_per_img, x.shape[1], _stream()))
    return x


def copy_rows(src: torch.Tensor, src_group_stride: int, src_row_stride: int, src_row0: int, dst: torch.Tensor,
              dst_group_stride: int, dst_row_stride: int, groups: int, rows_per_group: int, cols: int):
    check(_L.vlb200_copy_rows(_ptr(src), src_group_stride, src_row_stride, src_row0, _ptr(dst), dst_group_stride,
                              dst_row_stride, groups, rows_per_group, cols, _stream()))
    return dst


def gather_rows(src: torch.Tensor, index: torch.Tensor, out: torch.Tensor):
    assert index.dtype == torch.int32
    check(_L.vlb200_gather_rows(_ptr(src), _rowmajor_ld(src), _ptr(index), _ptr(out), _rowmajor_ld(out), index.numel(),
                                src.shape[1], _stream()))
    return out


def scatter_rows(src: torch.Tensor, index: torch.Tensor, out: torch.Tensor):
    assert index.dtype == torch.int32
    check(_L.vlb200_scatter_rows(_ptr(src), _rowmajor_ld(src), _ptr(index), _ptr(out), _rowmajor_ld(out), index.numel(),
                                 src.shape[1], _stream()))
    return out


def scatter_add_rows(src: torch.Tensor, index: torch.Tensor, out: torch.Tensor, scale: float = 1.0):
    """out[index[i], :] += scale * src[i, :]  (unique rows; out bf16 or fp32, may be a column slice)"""
    assert index.dtype == torch.int32 and src.dtype == torch.bfloat16 and src.shape[0] == index.numel()
    check(_L.vlb200_scatter_add_rows(_ptr(src), _rowmajor_ld(src), _ptr(index), _ptr(out), _dt(out), _rowmajor_ld(out),
                                     index.numel(), src.shape[1], float(scale), _stream()))
    return out


def zero_(t: torch.Tensor):
    assert t.is_contiguous()
    check(_L.vlb200_memset_zero(_ptr(t), t.numel() * t.element_size(), _stream()))
    return t


# ------------------------------------------------------------------------------------------
# LLaVA merge
# ------------------------------------------------------------------------------------------
class MergeIndex:
    """Device-side result of the integer merge pass (Llava/__init__.py:36-109)."""
    __slots__ = ("src_map", "labels", "mask", "pos", "seqlens", "img_pos", "row_of_text", "target", "status",
                 "n_seq", "L", "S", "P", "n_img_batch", "imgs_per_seq", "total_feats", "reps",
                 "row_starts", "rows", "rows_chosen",
                 "att_starts", "att_lens", "att_ctx", "att_kids", "n_att", "range_chosen", "range_rejected", "shared_rows")

    # Row layout of the merged activations.  Padded (default): sequence b at rows [b*S, (b+1)*S).  Packed (pack_merge_rows):
    # sequence b at rows [row_starts[b], row_starts[b] + len[b]), `rows` = sum(len) in all, `rows_chosen` of them chosen.
    @property
    def packed(self) -> bool:
        return getattr(self, "row_starts", None) is not None

    @property
    def starts(self) -> Optional[torch.Tensor]:
        """[n_seq + 1] int32 first rows of the sequences when packed, else None."""
        return getattr(self, "row_starts", None)

    @property
    def T(self) -> int:
        """Rows of every merged activation matrix."""
        return self.rows if self.packed else self.n_seq * self.S

    @property
    def T_chosen(self) -> int:
        """Rows of the chosen half (the first n_seq/2 sequences)."""
        return self.rows_chosen if self.packed else (self.n_seq // 2) * self.S

    # Shared-prefix rows (share_prefix_rows): rows = [chosen suffixes | one prefix per pair | rejected suffixes]; attention runs
    # over n_att = 3 * n_pairs sequences in that order (a suffix names its pair's prefix as context).
    @property
    def shared(self) -> bool:
        return getattr(self, "att_ctx", None) is not None

    @property
    def chosen_rows(self):
        """(lo, hi): the contiguous rows that make up the chosen sequences (shared: chosen suffixes + prefixes)."""
        return self.range_chosen if self.shared else (0, self.T_chosen)

    @property
    def rejected_rows(self):
        """(lo, hi): the contiguous rows of the rejected sequences (shared: prefixes + rejected suffixes)."""
        return self.range_rejected if self.shared else (self.T_chosen, self.T)

    @property
    def n_attn_seq(self) -> int:
        return self.n_att if self.shared else self.n_seq

    def attn(self) -> dict:
        """Sequence description the attention kernels take: seqlens, B, row_starts, total_rows (+ ctx / kids when shared)."""
        if self.shared:
            return dict(seqlens=self.att_lens, B=self.n_att, S=self.S, row_starts=self.att_starts, total_rows=self.T,
                        ctx=self.att_ctx, kids=self.att_kids)
        return dict(seqlens=self.seqlens, B=self.n_seq, S=self.S, row_starts=self.starts, total_rows=self.T)

    @property
    def row_stride(self) -> int:
        """merged_len argument of the kernels that address rows as b*merged_len + p (0: positions are absolute rows)."""
        return 0 if self.packed else self.S


def llava_merge_index(input_ids: torch.Tensor, attention_mask: torch.Tensor, labels: torch.Tensor, n_patches: int,
                      n_img_batch: int, imgs_per_seq: int, image_token: int, pad_token: int, ignore_index: int = -100):
    n_seq, L = input_ids.shape
    S = L + imgs_per_seq * (n_patches - 1)
    dev = input_ids.device
    for t in (input_ids, attention_mask, labels):
        assert t.dtype == torch.int64 and t.is_contiguous() and t.shape == (n_seq, L)
    m = MergeIndex()
    m.n_seq, m.L, m.S, m.P, m.n_img_batch, m.imgs_per_seq = n_seq, L, S, n_patches, n_img_batch, imgs_per_seq
    m.src_map = torch.empty(n_seq * S, dtype=torch.int32, device=dev)
    m.labels = torch.empty(n_seq, S, dtype=torch.int64, device=dev)
    m.mask = torch.empty(n_seq, S, dtype=torch.int32, device=dev)
    m.pos = torch.empty(n_seq * S, dtype=torch.int32, device=dev)
    m.seqlens = torch.empty(n_seq, dtype=torch.int32, device=dev)
    m.img_pos = torch.zeros(n_seq * imgs_per_seq * n_patches, dtype=torch.int32, device=dev)
    m.row_of_text = torch.empty(n_seq * (L - 1), dtype=torch.int32, device=dev)
    m.target = torch.empty(n_seq * (L - 1), dtype=torch.int64, device=dev)
    m.status = torch.empty(1, dtype=torch.int32, device=dev)
    check(_L.vlb200_llava_merge_index(_ptr(input_ids), _ptr(attention_mask), _ptr(labels), n_seq, L, S, n_patches,
                                      n_img_batch, imgs_per_seq, image_token, pad_token, ignore_index, _ptr(m.src_map),
                                      _ptr(m.labels), _ptr(m.mask), _ptr(m.pos), _ptr(m.seqlens), _ptr(m.img_pos),
                                      _ptr(m.row_of_text), _ptr(m.target), _ptr(m.status), _stream()))
    return m


def llava_merge_embed(m: MergeIndex, embed_tokens: torch.Tensor, image_features: torch.Tensor, out: torch.Tensor):
    check(_L.vlb200_llava_merge_embed(_ptr(m.src_map), _ptr(embed_tokens), _ptr(image_features), _ptr(out), _dt(out),
                                      m.T, embed_tokens.shape[1], _stream()))
    return out


def llava_merge_bwd(m: MergeIndex, dx: torch.Tensor, dembed_f32: torch.Tensor, dimage_features: torch.Tensor):
    assert dembed_f32.dtype == torch.float32
    check(_L.vlb200_llava_merge_bwd_rows(_ptr(m.src_map), _ptr(m.img_pos), _ptr(dx), _ptr(dembed_f32), _ptr(dimage_features),
                                         m.T, m.n_seq, m.n_img_batch, m.row_stride, m.imgs_per_seq * m.P, dx.shape[1],
                                         _stream()))


def llavanext_merge_index(input_ids: torch.Tensor, attention_mask: torch.Tensor, labels: torch.Tensor,
                          feat_off: torch.Tensor, total_feats: int, merged_len: int, n_img_batch: int, imgs_per_seq: int,
                          image_token: int, ignore_index: int = -100):
    """Integer merge pass of LlavaNextForRL (LlavaNext/__init__.py:38-171): variable packed feature lengths
    (`feat_off` int32 prefix offsets, one entry per image + 1), masked tokens dropped, S = `merged_len`."""
    n_seq, L = input_ids.shape
    S = int(merged_len)
    dev = input_ids.device
    for t in (input_ids, attention_mask, labels):
        assert t.dtype == torch.int64 and t.is_contiguous() and t.shape == (n_seq, L)
    assert feat_off.dtype == torch.int32 and feat_off.numel() == n_img_batch * imgs_per_seq + 1
    m = MergeIndex()
    m.n_seq, m.L, m.S, m.P, m.n_img_batch, m.imgs_per_seq = n_seq, L, S, -1, n_img_batch, imgs_per_seq
    m.total_feats, m.reps = int(total_feats), n_seq // n_img_batch
    m.src_map = torch.empty(n_seq * S, dtype=torch.int32, device=dev)
    m.labels = torch.empty(n_seq, S, dtype=torch.int64, device=dev)
    m.mask = torch.empty(n_seq, S, dtype=torch.int32, device=dev)
    m.pos = torch.empty(n_seq * S, dtype=torch.int32, device=dev)
    m.seqlens = torch.empty(n_seq, dtype=torch.int32, device=dev)
    m.img_pos = torch.zeros(m.reps * m.total_feats, dtype=torch.int32, device=dev)  # img_rows
    m.row_of_text = torch.empty(n_seq * (L - 1), dtype=torch.int32, device=dev)
    m.target = torch.empty(n_seq * (L - 1), dtype=torch.int64, device=dev)
    m.status = torch.empty(1, dtype=torch.int32, device=dev)
    check(_L.vlb200_llavanext_merge_index(_ptr(input_ids), _ptr(attention_mask), _ptr(labels), _ptr(feat_off), n_seq, L, S,
                                          n_img_batch, imgs_per_seq, m.total_feats, image_token, ignore_index,
                                          _ptr(m.src_map), _ptr(m.labels), _ptr(m.mask), _ptr(m.pos), _ptr(m.seqlens),
                                          _ptr(m.img_pos), _ptr(m.row_of_text), _ptr(m.target), _ptr(m.status), _stream()))
    return m


def llavanext_merge_bwd(m: MergeIndex, dx: torch.Tensor, dembed_f32: torch.Tensor, dimage_features: torch.Tensor):
    assert dembed_f32.dtype == torch.float32 and dimage_features.shape[0] == m.total_feats
    check(_L.vlb200_llavanext_merge_bwd(_ptr(m.src_map), _ptr(m.img_pos), _ptr(dx), _ptr(dembed_f32),
                                        _ptr(dimage_features), m.T, m.total_feats, m.reps, dx.shape[1], _stream()))


def pack_merge_rows(m: MergeIndex, seq_lens) -> MergeIndex:
    """Drop the padding rows of a LLaVA-1.5 / LLaVA-Next / Qwen-VL merge index: `seq_lens` (host ints, one per sequence,
    == m.seqlens) gives the attended prefix of every sequence; afterwards sequence b lives at rows
    [row_starts[b], row_starts[b] + seq_lens[b]) and m.T = sum(seq_lens).  m.labels / m.mask keep the padded [n_seq, S] layout
    (they feed no kernel)."""
    lens = [int(x) for x in seq_lens]
    if len(lens) != m.n_seq or any(n < 0 or n > m.S for n in lens):
        raise ValueError(f"pack_merge_rows: {len(lens)} lengths for {m.n_seq} sequences of at most {m.S} rows")
    if m.packed:
        raise ValueError("pack_merge_rows: already packed")
    starts = [0]
    for n in lens:
        starts.append(starts[-1] + n)
    dev = m.src_map.device
    row_starts = torch.tensor(starts, dtype=torch.int32).to(dev, non_blocking=True)
    rows = max(starts[-1], 1)
    src_p = torch.empty(rows, dtype=torch.int32, device=dev)
    pos_p = torch.empty(rows, dtype=torch.int32, device=dev)
    next_layout = getattr(m, "total_feats", None) is not None and getattr(m, "reps", None) is not None
    img_pos, n_img_pos, feats, img_rows, n_img_rows = None, 0, 0, None, 0
    if getattr(m, "img_pos", None) is None:
        pass                 # Qwen-VL: the image rows are placeholder tokens of the text sequence (frozen tower): no row list
    elif next_layout:        # LLaVA-Next: img_pos holds flat merged rows
        img_rows, n_img_rows = m.img_pos, m.img_pos.numel()
    else:                    # LLaVA-1.5 / XC2: img_pos holds positions inside the sequence
        img_pos, n_img_pos, feats = m.img_pos, m.img_pos.numel(), m.imgs_per_seq * m.P
    check(_L.vlb200_pack_merge_rows(_ptr(m.src_map), _ptr(m.pos), _ptr(row_starts), m.n_seq, m.S, _ptr(src_p), _ptr(pos_p),
                                    _ptr(m.row_of_text), m.row_of_text.numel(), _ptr(img_pos), n_img_pos, feats,
                                    _ptr(img_rows), n_img_rows, _stream()))
    m.src_map, m.pos = src_p, pos_p
    m.row_starts, m.rows, m.rows_chosen = row_starts, rows, starts[m.n_seq // 2]
    return m


def share_prefix_rows(m: MergeIndex, seq_lens, prefix_rows) -> MergeIndex:
    """Lay the merged rows out with ONE copy of every pair's common prefix (include/vlb200.h: vlb200_share_prefix_rows).
    `seq_lens` (host ints, one per sequence == m.seqlens) and `prefix_rows` (host ints, one per PAIR: merged rows the chosen
    and the rejected sequence have in common; 0 = nothing shared) come from the host batch (host.shared_prefix_rows), so the step
    needs no device read-back.  Afterwards m.T = sum(seq_lens) - sum(prefix_rows); rows = [chosen suffixes | prefixes | rejected
    suffixes]; m.attn() describes 3 * n_pairs attention sequences."""
    lens = [int(x) for x in seq_lens]
    npair = m.n_seq // 2
    pre = [int(x) for x in prefix_rows]
    if len(lens) != m.n_seq or len(pre) != npair:
        raise ValueError(f"share_prefix_rows: {len(lens)} lengths / {len(pre)} prefixes for {m.n_seq} sequences")
    if m.packed:
        raise ValueError("share_prefix_rows: the index is already packed")
    for i in range(npair):
        if pre[i] < 0 or pre[i] > min(lens[i], lens[npair + i]):
            raise ValueError(f"share_prefix_rows: prefix {pre[i]} of pair {i} exceeds its sequences ({lens[i]}, {lens[npair + i]})")
    # rows: chosen suffixes, prefixes, rejected suffixes; attention sequences in the same order
    att_lens = [lens[i] - pre[i] for i in range(npair)] + pre + [lens[npair + i] - pre[i] for i in range(npair)]
    att_starts = [0]
    for n in att_lens:
        att_starts.append(att_starts[-1] + n)
    rows = max(att_starts[-1], 1)
    pre_start = [att_starts[npair + i] for i in range(npair)]
    suf_start = [att_starts[i] for i in range(npair)] + [att_starts[2 * npair + i] for i in range(npair)]
    ctx = [npair + i if pre[i] > 0 else -1 for i in range(npair)] + [-1] * npair + \
          [npair + i if pre[i] > 0 else -1 for i in range(npair)]
    kids = [-1] * (2 * 3 * npair)
    for i in range(npair):
        if pre[i] > 0:
            kids[2 * (npair + i)], kids[2 * (npair + i) + 1] = i, 2 * npair + i
    dev = m.src_map.device
    host = torch.tensor(pre + pre + pre_start + pre_start + suf_start + lens + att_starts + att_lens + ctx + kids, dtype=torch.int32)
    d = host.to(dev, non_blocking=True)
    n2, n3 = m.n_seq, 3 * npair
    o = 0
    pre_d = d[o:o + n2]; o += n2
    pst_d = d[o:o + n2]; o += n2
    sst_d = d[o:o + n2]; o += n2
    len_d = d[o:o + n2]; o += n2
    m.att_starts = d[o:o + n3 + 1]; o += n3 + 1
    m.att_lens = d[o:o + n3]; o += n3
    m.att_ctx = d[o:o + n3]; o += n3
    m.att_kids = d[o:o + 2 * n3]
    src_s = torch.empty(rows, dtype=torch.int32, device=dev)
    pos_s = torch.empty(rows, dtype=torch.int32, device=dev)
    next_layout = getattr(m, "total_feats", None) is not None and getattr(m, "reps", None) is not None
    img_pos, n_img_pos, feats, img_rows, n_img_rows = None, 0, 0, None, 0
    if getattr(m, "img_pos", None) is None:
        pass
    elif next_layout:
        raise ValueError("share_prefix_rows: LLaVA-Next image rows (variable feature lengths) are not supported yet")
    else:
        img_pos, n_img_pos, feats = m.img_pos, m.img_pos.numel(), m.imgs_per_seq * m.P
    check(_L.vlb200_share_prefix_rows(_ptr(m.src_map), _ptr(m.pos), _ptr(pre_d), _ptr(pst_d), _ptr(sst_d), _ptr(len_d), m.n_seq, m.S,
                                      rows, _ptr(src_s), _ptr(pos_s), _ptr(m.row_of_text), m.row_of_text.numel(), _ptr(img_pos),
                                      n_img_pos, feats, _ptr(img_rows), n_img_rows, _stream()))
    m.src_map, m.pos = src_s, pos_s
    m.row_starts, m.rows, m.rows_chosen = m.att_starts, rows, att_starts[2 * npair]
    m.n_att = n3
    m.range_chosen, m.range_rejected = (0, att_starts[2 * npair]), (att_starts[npair], att_starts[3 * npair])
    m.shared_rows = sum(pre)
    return m


def qwen_merge_index(input_ids: torch.Tensor, attention_mask: torch.Tensor, labels: torch.Tensor, n_queries: int,
                     n_img_batch: int, imgs_per_seq: int, image_start_id: int, ignore_index: int = -100):
    """Integer pass of QWenModel.forward's image placement (modeling_qwen.py:524-528,614-621); S == L."""
    n_seq, L = input_ids.shape
    dev = input_ids.device
    for t in (input_ids, attention_mask, labels):
        assert t.dtype == torch.int64 and t.is_contiguous() and t.shape == (n_seq, L)
    m = MergeIndex()
    m.n_seq, m.L, m.S, m.P, m.n_img_batch, m.imgs_per_seq = n_seq, L, L, n_queries, n_img_batch, imgs_per_seq
    m.total_feats, m.reps = n_img_batch * imgs_per_seq * n_queries, n_seq // n_img_batch
    m.src_map = torch.empty(n_seq * L, dtype=torch.int32, device=dev)
    m.labels = torch.empty(n_seq, L, dtype=torch.int64, device=dev)
    m.mask = torch.empty(n_seq, L, dtype=torch.int32, device=dev)
    m.pos = torch.empty(n_seq * L, dtype=torch.int32, device=dev)
    m.seqlens = torch.empty(n_seq, dtype=torch.int32, device=dev)
    m.img_pos = None
    m.row_of_text = torch.empty(n_seq * (L - 1), dtype=torch.int32, device=dev)
    m.target = torch.empty(n_seq * (L - 1), dtype=torch.int64, device=dev)
    m.status = torch.empty(1, dtype=torch.int32, device=dev)
    check(_L.vlb200_qwen_merge_index(_ptr(input_ids), _ptr(attention_mask), _ptr(labels), n_seq, L, n_queries, n_img_batch,
                                     imgs_per_seq, image_start_id, ignore_index, _ptr(m.src_map), _ptr(m.labels), _ptr(m.mask),
                                     _ptr(m.pos), _ptr(m.seqlens), _ptr(m.row_of_text), _ptr(m.target), _ptr(m.status),
                                     _stream()))
    return m


# ------------------------------------------------------------------------------------------
# attention
# ------------------------------------------------------------------------------------------
def attn_fwd_tc(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, out: torch.Tensor, lse: Optional[torch.Tensor],
                seqlens: Optional[torch.Tensor], B: int, S: int, H: int, KVH: int, head_dim: int, causal: bool,
                scale: float, row_starts: Optional[torch.Tensor] = None, total_rows: int = 0,
                ctx: Optional[torch.Tensor] = None, kids: Optional[torch.Tensor] = None):
    """tcgen05/TMEM/TMA forward; q/k/v/out: 2-D row-major views [B*S, >= heads*head_dim] (may be column slices of one qkv buffer).  row_starts ([B+1] int32) + total_rows: packed rows, sequence b
    at rows [row_starts[b], row_starts[b] + seqlens[b]) (include/vlb200.h: vlb200_attn_fwd_tc_varlen)."""
    if ctx is not None:
        assert row_starts is not None and ctx.dtype == torch.int32 and ctx.numel() == B
        check(_L.vlb200_attn_fwd_tc_ctx(_ptr(q), q.stride(0), _ptr(k), k.stride(0), _ptr(v), v.stride(0), _ptr(out),
                                        out.stride(0), _ptr(lse), _ptr(seqlens), _ptr(row_starts), _ptr(ctx), int(total_rows), B, S,
                                        H, KVH, head_dim, int(causal), scale, _stream()))
    elif row_starts is None:
        check(_L.vlb200_attn_fwd_tc(_ptr(q), q.stride(0), _ptr(k), k.stride(0), _ptr(v), v.stride(0), _ptr(out),
                                    out.stride(0), _ptr(lse), _ptr(seqlens), B, S, H, KVH, head_dim, int(causal), scale,
                                    _stream()))
    else:
        check(_L.vlb200_attn_fwd_tc_varlen(_ptr(q), q.stride(0), _ptr(k), k.stride(0), _ptr(v), v.stride(0), _ptr(out),
                                           out.stride(0), _ptr(lse), _ptr(seqlens), _ptr(row_starts), int(total_rows), B, S, H,
                                           KVH, head_dim, int(causal), scale, _stream()))
    return out


def attn_bwd_tc(q, k, v, out, dout, lse, delta, dq, dk, dv, seqlens, B, S, H, KVH, head_dim, causal, scale,
                row_starts: Optional[torch.Tensor] = None, total_rows: int = 0,
                ctx: Optional[torch.Tensor] = None, kids: Optional[torch.Tensor] = None):
    """tcgen05/TMEM/TMA backward; same layout as attn_fwd_tc.  row_starts/total_rows: packed rows (see attn_fwd_tc)."""
    if ctx is not None:
        assert row_starts is not None and kids is not None and ctx.numel() == B and kids.numel() == 2 * B
        check(_L.vlb200_attn_bwd_tc_ctx(_ptr(q), q.stride(0), _ptr(k), k.stride(0), _ptr(v), v.stride(0), _ptr(out),
                                        out.stride(0), _ptr(dout), dout.stride(0), _ptr(lse), _ptr(delta), _ptr(dq), dq.stride(0),
                                        _ptr(dk), dk.stride(0), _ptr(dv), dv.stride(0), _ptr(seqlens), _ptr(row_starts), _ptr(ctx),
                                        _ptr(kids), int(total_rows), B, S, H, KVH, head_dim, int(causal), scale, _stream()))
    elif row_starts is None:
        check(_L.vlb200_attn_bwd_tc(_ptr(q), q.stride(0), _ptr(k), k.stride(0), _ptr(v), v.stride(0), _ptr(out),
                                    out.stride(0), _ptr(dout), dout.stride(0), _ptr(lse), _ptr(delta), _ptr(dq), dq.stride(0),
                                    _ptr(dk), dk.stride(0), _ptr(dv), dv.stride(0), _ptr(seqlens), B, S, H, KVH, head_dim,
                                    int(causal), scale, _stream()))
    else:
        check(_L.vlb200_attn_bwd_tc_varlen(_ptr(q), q.stride(0), _ptr(k), k.stride(0), _ptr(v), v.stride(0), _ptr(out),
                                           out.stride(0), _ptr(dout), dout.stride(0), _ptr(lse), _ptr(delta), _ptr(dq),
                                           dq.stride(0), _ptr(dk), dk.stride(0), _ptr(dv), dv.stride(0), _ptr(seqlens),
                                           _ptr(row_starts), int(total_rows), B, S, H, KVH, head_dim, int(causal), scale,
                                           _stream()))


# ------------------------------------------------------------------------------------------
# optimizer
# ------------------------------------------------------------------------------------------
def sumsq(x: torch.Tensor, out: torch.Tensor, workspace: torch.Tensor, accumulate: bool = False):
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and workspace.numel() >= 1024
    check(_L.vlb200_sumsq_bf16(_ptr(x), x.numel(), _ptr(workspace), _ptr(out), int(accumulate), _stream()))
    return out


def adamw_(param: torch.Tensor, grad: torch.Tensor, master: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor,
           lr: float, beta1: float, beta2: float, eps: float, weight_decay: float, step: int, grad_scale: float = 1.0,
           grad_sumsq: Optional[torch.Tensor] = None, max_grad_norm: float = 0.0):
    n = param.numel()
    assert grad.numel() == n and master.numel() == n and exp_avg.numel() == n and exp_avg_sq.numel() == n
    check(_L.vlb200_adamw(_ptr(param), _ptr(grad), _ptr(master), _ptr(exp_avg), _ptr(exp_avg_sq), n, lr, beta1, beta2,
                          eps, weight_decay, step, grad_scale, _ptr(grad_sumsq), max_grad_norm, _stream()))


def cast_f32_to_bf16(src: torch.Tensor, dst: torch.Tensor, scale: float = 1.0):
    check(_L.vlb200_cast_f32_to_bf16(_ptr(src), _ptr(dst), src.numel(), scale, _stream()))
    return dst


def cast_bf16_to_f32(src: torch.Tensor, dst: torch.Tensor):
    check(_L.vlb200_cast_bf16_to_f32(_ptr(src), _ptr(dst), src.numel(), _stream()))
    return dst


# ------------------------------------------------------------------------------------------
# input pipeline: CLIP image preprocessing (Llava/__init__.py:435-443 -> CLIPImageProcessor)
# ------------------------------------------------------------------------------------------
def clip_preprocess_u8(image: torch.Tensor, coef_h: torch.Tensor, bounds_h: torch.Tensor, ksize_h: int, coef_v: torch.Tensor,
                       bounds_v: torch.Tensor, ksize_v: int, new_h: int, new_w: int, top: int, left: int, crop_h: int,
                       crop_w: int, row0: int, rows: int, workspace: torch.Tensor, rescale: float, mean_std_host,
                       out: torch.Tensor):
    """image uint8 [H, W, 3] (CUDA) -> out [3, crop_h, crop_w]; `mean_std_host` is a contiguous float32 numpy array
    of 6 values (mean RGB, std RGB) on the host."""
    assert image.dtype == torch.uint8 and image.is_contiguous() and image.dim() == 3 and image.shape[2] == 3
    assert workspace.dtype == torch.uint8 and out.is_contiguous() and tuple(out.shape) == (3, crop_h, crop_w)
    for t in (coef_h, bounds_h, coef_v, bounds_v):
        assert t.dtype == torch.int32 and t.is_contiguous()
    assert mean_std_host.dtype.name == "float32" and mean_std_host.size == 6 and mean_std_host.flags["C_CONTIGUOUS"]
    check(_L.vlb200_clip_preprocess_u8(_ptr(image), int(image.shape[0]), int(image.shape[1]), _ptr(coef_h), _ptr(bounds_h),
                                       ksize_h, _ptr(coef_v), _ptr(bounds_v), ksize_v, new_h, new_w, top, left, crop_h, crop_w,
                                       row0, rows, _ptr(workspace), workspace.numel(), float(rescale),
                                       mean_std_host.ctypes.data, _ptr(out), _dt(out), _stream()))
    return out
