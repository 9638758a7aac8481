"""TEST-ONLY stand-in for `vlrlhf_b200.ops` built from plain PyTorch CPU ops (fp32 math, bf16 storage).

Purpose: run the engine's host-side orchestration (buffer plumbing, GEMM orientations, backward formulas,
optimizer sequencing, data-parallel all-reduce over gloo) on the CPU-only build box, against the oracle.
It is never importable from the product package and never used on a GPU box."""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F

BF16, F32 = 0, 1
ACT_NONE, ACT_QUICK_GELU, ACT_GELU_ERF = 0, 1, 2
LOSS_TYPES = {"sigmoid": 0, "hinge": 1, "ipo": 2, "kto_pair": 3, "ddpo": 4}
_launches = [0]


def launch_count():
    return _launches[0]


def _c(n=1):
    _launches[0] += n


def _lowbias32(x):
    x = x.clone()
    x ^= x >> 16
    x = (x * 0x7FEB352D) & 0xFFFFFFFF
    x ^= x >> 15
    x = (x * 0x846CA68B) & 0xFFFFFFFF
    x ^= x >> 16
    return x


def init_uniform_(t, seed, scale, shift=0.0):
    from oracle.restate import hash_uniform
    t.copy_(hash_uniform(t.numel(), seed, scale, shift).to(t.dtype))
    _c()
    return t


def perturb_(dst, base, other, alpha, shift):
    dst.copy_((base.float() + alpha * (other.float() - shift)).to(torch.bfloat16))
    _c()
    return dst


def gemm(a, b, *, a_kmajor=True, b_kmajor=True, out=None, out_dtype=torch.bfloat16, bias=None, act=ACT_NONE,
         residual=None, accumulate=False, a2=None, b2=None, alpha=1.0):
    A = a.float() if a_kmajor else a.float().t()
    B = b.float() if b_kmajor else b.float().t()
    if A.shape[1] != B.shape[1]:
        raise ValueError("gemm: contraction mismatch")
    y = A @ B.t()
    if (a2 is None) != (b2 is None):
        raise ValueError("gemm: a2 and b2 come together")
    if a2 is not None:
        assert a2.dtype == torch.bfloat16 and b2.dtype == torch.bfloat16
        A2 = a2.float() if a_kmajor else a2.float().t()
        B2 = b2.float() if b_kmajor else b2.float().t()
        if A2.shape[0] != A.shape[0] or B2.shape[0] != B.shape[0] or A2.shape[1] != B2.shape[1]:
            raise ValueError("gemm: second operand pair does not match")
        y = y + A2 @ B2.t()
    y = y * alpha
    if bias is not None:
        y = y + bias.float()
    if act == ACT_QUICK_GELU:
        y = y * torch.sigmoid(1.702 * y)
    elif act == ACT_GELU_ERF:
        y = F.gelu(y)
    if residual is not None:
        y = y + residual.float()
    if out is None:
        out = torch.empty(y.shape, dtype=out_dtype)
    if accumulate:
        y = y + out.float()
    out.copy_(y.to(out.dtype))
    _c()
    return out


def gemm_swiglu(a, wgu, gu, act, write_gu=True):
    ff = wgu.shape[0] // 2
    if wgu.shape[1] != a.shape[1] or tuple(gu.shape) != (a.shape[0], 2 * ff) or tuple(act.shape) != (a.shape[0], ff):
        raise ValueError("gemm_swiglu: shape mismatch")
    y = (a.float() @ wgu.float().t()).to(torch.bfloat16)
    act.copy_((F.silu(y[:, :ff].float()) * y[:, ff:].float()).to(torch.bfloat16))
    if write_gu:
        gu.copy_(y)
    _c()
    return act


def logps_fwd(logits, target, n_seq, weight=None, average_log_prob=False):
    rows, V = logits.shape
    if target.numel() != rows or rows % n_seq != 0:
        raise ValueError("Logits (batch and sequence length dim) and labels must have the same shape.")
    on = target >= 0
    if weight is not None:
        on = on & (weight != 0)
    lse = torch.logsumexp(logits.float(), -1)
    pick = logits.float().gather(1, target.clamp(min=0)[:, None])[:, 0]
    per = torch.where(on, pick - lse, torch.zeros_like(lse))
    s = per.view(n_seq, -1).sum(-1)
    if average_log_prob:
        s = s / on.view(n_seq, -1).sum(-1)
    _c(2)
    return s, per, torch.where(on, lse, torch.zeros_like(lse))


def logps_bwd(logits, target, n_seq, lse, grad_logps, weight=None, average_log_prob=False, out=None):
    rows, V = logits.shape
    on = target >= 0
    if weight is not None:
        on = on & (weight != 0)
    g = grad_logps.repeat_interleave(rows // n_seq)
    if average_log_prob:
        g = g / on.view(n_seq, -1).sum(-1).repeat_interleave(rows // n_seq)
    p = torch.exp(logits.float() - lse[:, None])
    d = -g[:, None] * p
    d[torch.arange(rows), target.clamp(min=0)] += g
    d = d * on[:, None]
    if out is None:
        out = torch.empty(rows, V, dtype=torch.bfloat16)
    out.copy_(d.to(torch.bfloat16))
    _c()
    return out


def dpo_loss(policy_logps, ref_logps, beta, label_smoothing=0.0, loss_type="sigmoid", reference_free=False,
             loss_scale=1.0, want_grad=True):
    from oracle.restate import dpo_loss as oracle_loss  # test-only module: the oracle is allowed here
    if loss_type not in LOSS_TYPES:
        raise ValueError(f"Unknown loss type: {loss_type}. Should be one of ['sigmoid', 'hinge', 'ipo', 'kto_pair']")
    n = policy_logps.numel() // 2
    grad = None
    with torch.enable_grad():
        p = policy_logps.detach().float().clone().requires_grad_(True)
        losses, cr, rr = oracle_loss(p[:n], p[n:], ref_logps[:n].detach().float(), ref_logps[n:].detach().float(), beta,
                                     label_smoothing, loss_type, reference_free)
        if want_grad:
            (losses.mean() * loss_scale).backward()
            grad = p.grad
    stats = torch.stack([losses.mean().detach(), (cr > rr).float().mean(), cr.mean(), rr.mean(), (cr - rr).mean(),
                         torch.tensor(float(losses.numel()))])
    _c()
    return losses.detach(), cr, rr, stats, grad


def rmsnorm_fwd(x, w, eps, out=None, rstd=None):
    xf = x.float()
    r = torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    if rstd is not None:
        rstd.copy_(r[:, 0])
    if out is None:
        out = torch.empty_like(x)
    out.copy_((w.float() * (xf * r)).to(out.dtype))
    _c()
    return out


def rmsnorm_bwd(dy, x, w, rstd, dw, dres=None, out=None, dw_accumulate=False):
    xf, g, wf = x.float(), dy.float(), w.float()
    xh = xf * rstd[:, None]
    dot = (wf * g * xh).mean(-1, keepdim=True)
    dx = rstd[:, None] * (wf * g - xh * dot)
    if dres is not None:
        dx = dx + dres.float()
    dwv = (g * xh).sum(0)
    if dw_accumulate:
        dwv = dwv + dw.float()
    dw.copy_(dwv.to(dw.dtype))
    if out is None:
        out = torch.empty_like(x)
    out.copy_(dx.to(out.dtype))
    _c(2)
    return out


def layernorm_fwd(x, w, b, eps, out=None):
    if out is None:
        out = torch.empty_like(x)
    out.copy_(F.layer_norm(x.float(), (x.shape[1],), w.float(), b.float(), eps).to(out.dtype))
    _c()
    return out


def colsum(a, out, accumulate=False):
    s = a.float().sum(0)
    if accumulate:
        s = s + out.float()
    out.copy_(s.to(out.dtype))
    _c(2)
    return out


def rope_(qkv, pos, cos_t, sin_t, n_rot_heads, head_dim, inverse=False):
    T = qkv.shape[0]
    x = qkv[:, :n_rot_heads * head_dim].float().view(T, n_rot_heads, head_dim)
    c = torch.cat([cos_t, cos_t], -1)[pos.long()][:, None]
    s = torch.cat([sin_t, sin_t], -1)[pos.long()][:, None]
    if inverse:
        s = -s
    rot = torch.cat([-x[..., head_dim // 2:], x[..., :head_dim // 2]], -1)
    qkv[:, :n_rot_heads * head_dim] = (x * c + rot * s).reshape(T, -1).to(qkv.dtype)
    _c()
    return qkv


def swiglu_fwd(gate_up, out=None):
    ff = gate_up.shape[1] // 2
    y = F.silu(gate_up[:, :ff].float()) * gate_up[:, ff:].float()
    if out is None:
        out = torch.empty(y.shape, dtype=torch.bfloat16)
    out.copy_(y.to(out.dtype))
    _c()
    return out


def swiglu_bwd(gate_up, dact, out=None):
    ff = gate_up.shape[1] // 2
    g, u, d = gate_up[:, :ff].float(), gate_up[:, ff:].float(), dact.float()
    s = torch.sigmoid(g)
    res = torch.cat([d * u * s * (1 + g * (1 - s)), d * g * s], 1)
    if out is None:
        out = torch.empty_like(gate_up)
    out.copy_(res.to(out.dtype))
    _c()
    return out


def gelu_fwd(z, out=None):
    if out is None:
        out = torch.empty_like(z)
    out.copy_(F.gelu(z.float()).to(out.dtype))
    _c()
    return out


def gelu_bwd(z, dh, out=None):
    x = z.float()
    cdf = 0.5 * (1 + torch.erf(x / math.sqrt(2)))
    pdf = torch.exp(-0.5 * x * x) / math.sqrt(2 * math.pi)
    res = dh.float() * (cdf + x * pdf)
    if out is None:
        out = torch.empty_like(z)
    out.copy_(res.to(out.dtype))
    _c()
    return out


def clip_im2col(pixels, patch, out):
    B = pixels.shape[0]
    cols = F.unfold(pixels.float(), kernel_size=patch, stride=patch).transpose(1, 2).reshape(-1, 3 * patch * patch)
    out[:, :cols.shape[1]] = cols.to(out.dtype)
    _c()
    return out


def clip_cls_rows_(x, cls, pos0, batch, tokens_per_img):
    x.view(batch, tokens_per_img, -1)[:, 0] = (cls.float() + pos0.float()).to(x.dtype)
    _c()
    return x


def copy_rows(src, src_group_stride, src_row_stride, src_row0, dst, dst_group_stride, dst_row_stride, groups,
              rows_per_group, cols):
    s = torch.as_strided(src, (groups, rows_per_group, cols), (src_group_stride, src_row_stride, 1),
                         src.storage_offset() + src_row0 * src_row_stride)
    d = torch.as_strided(dst, (groups, rows_per_group, cols), (dst_group_stride, dst_row_stride, 1), dst.storage_offset())
    d.copy_(s)
    _c()
    return dst


def gather_rows(src, index, out):
    idx = index.long()
    out.copy_(torch.where((idx >= 0)[:, None], src[idx.clamp(min=0)], torch.zeros_like(out)))
    _c()
    return out


def scatter_rows(src, index, out):
    idx = index.long()
    ok = idx >= 0
    out[idx[ok]] = src[ok]
    _c()
    return out


def scatter_add_rows(src, index, out, scale=1.0):
    idx = index.long()
    ok = idx >= 0
    out[idx[ok]] = (out[idx[ok]].float() + scale * src[ok].float()).to(out.dtype)
    _c()
    return out


def zero_(t):
    t.zero_()
    return t


class MergeIndex:
    row_starts = None
    total_feats = None
    reps = None
    att_ctx = None

    @property
    def shared(self):
        return self.att_ctx is not None

    @property
    def chosen_rows(self):
        return self.range_chosen if self.shared else (0, self.T_chosen)

    @property
    def rejected_rows(self):
        return self.range_rejected if self.shared else (self.T_chosen, self.T)

    @property
    def n_attn_seq(self):
        return self.n_att if self.shared else self.n_seq

    def attn(self):
        if self.shared:
            return dict(seqlens=self.att_lens, B=self.n_att, S=self.S, row_starts=self.att_starts, total_rows=self.T,
                        ctx=self.att_ctx, kids=self.att_kids)
        return dict(seqlens=self.seqlens, B=self.n_seq, S=self.S, row_starts=self.starts, total_rows=self.T)

    @property
    def packed(self):
        return self.row_starts is not None

    @property
    def starts(self):
        return self.row_starts

    @property
    def T(self):
        return self.rows if self.packed else self.n_seq * self.S

    @property
    def T_chosen(self):
        return self.rows_chosen if self.packed else (self.n_seq // 2) * self.S

    @property
    def row_stride(self):
        return 0 if self.packed else self.S


def pack_merge_rows(m, seq_lens):
    """CPU mirror of vlb200_pack_merge_rows (csrc/elementwise.cu): keep the first seq_lens[b] rows of every sequence."""
    lens = [int(x) for x in seq_lens]
    if len(lens) != m.n_seq or any(n < 0 or n > m.S for n in lens):
        raise ValueError(f"pack_merge_rows: {len(lens)} lengths for {m.n_seq} sequences of at most {m.S} rows")
    if m.packed:
        raise ValueError("pack_merge_rows: already packed")
    starts = [0]
    for n in lens:
        starts.append(starts[-1] + n)
    S = m.S
    keep = torch.cat([torch.arange(b * S, b * S + n) for b, n in enumerate(lens)]) if starts[-1] else torch.zeros(0, dtype=torch.long)
    st = torch.tensor(starts, dtype=torch.long)
    ln = torch.tensor(lens, dtype=torch.long)

    def remap(rows):
        r = rows.long()
        b, p_ = r.clamp(min=0) // S, r.clamp(min=0) % S
        ok = (r >= 0) & (b < m.n_seq)
        bb = b.clamp(max=m.n_seq - 1)
        new = torch.where(ok & (p_ < ln[bb]), st[bb] + p_, torch.full_like(r, -1))
        return torch.where(r < 0, r, new).to(torch.int32)

    m.src_map = m.src_map.reshape(-1)[keep].contiguous()
    m.pos = m.pos.reshape(-1)[keep].contiguous()
    m.row_of_text = remap(m.row_of_text.reshape(-1))
    if getattr(m, "img_pos", None) is None:                 # Qwen-VL: no image row list
        pass
    elif m.total_feats is not None and m.reps is not None:  # LLaVA-Next: flat merged rows
        m.img_pos = remap(m.img_pos.reshape(-1))
    else:                                                   # LLaVA-1.5 / XC2: positions inside the sequence -> absolute rows
        feats = m.imgs_per_seq * m.P
        b = torch.arange(m.img_pos.numel()) // feats
        m.img_pos = (m.img_pos.reshape(-1).long() + st[b]).to(torch.int32)
    m.row_starts = torch.tensor(starts, dtype=torch.int32)
    m.rows, m.rows_chosen = max(starts[-1], 1), starts[m.n_seq // 2]
    _c(3)
    return m


def share_prefix_rows(m, seq_lens, prefix_rows):
    """CPU mirror of vlb200_share_prefix_rows (csrc/elementwise.cu) + ops.share_prefix_rows: one copy of every pair's common
    prefix; rows = [chosen suffixes | prefixes | rejected suffixes]; 3 * n_pairs attention sequences in that order."""
    lens = [int(x) for x in seq_lens]
    npair = m.n_seq // 2
    pre = [int(x) for x in prefix_rows]
    if len(lens) != m.n_seq or len(pre) != npair:
        raise ValueError("share_prefix_rows: bad lengths")
    if m.packed:
        raise ValueError("share_prefix_rows: the index is already packed")
    if getattr(m, "img_pos", None) is not None and m.total_feats is not None and m.reps is not None:   # (Qwen-VL: no row list)
        raise ValueError("share_prefix_rows: LLaVA-Next image rows (variable feature lengths) are not supported yet")
    for i in range(npair):
        if pre[i] < 0 or pre[i] > min(lens[i], lens[npair + i]):
            raise ValueError("share_prefix_rows: prefix exceeds its sequences")
    att_lens = [lens[i] - pre[i] for i in range(npair)] + pre + [lens[npair + i] - pre[i] for i in range(npair)]
    att_starts = [0]
    for n in att_lens:
        att_starts.append(att_starts[-1] + n)
    rows = max(att_starts[-1], 1)
    S = m.S

    def dest(b, p_):
        if p_ >= lens[b]:
            return -1
        i = b % npair
        if p_ < pre[i]:
            return att_starts[npair + i] + p_
        return att_starts[b if b < npair else 2 * npair + i] + (p_ - pre[i])

    src_old, pos_old = m.src_map.reshape(-1), m.pos.reshape(-1)
    src_new = torch.zeros(rows, dtype=torch.int32)
    pos_new = torch.zeros(rows, dtype=torch.int32)
    for b in range(m.n_seq):
        for p_ in range(lens[b]):
            r = dest(b, p_)
            if p_ >= pre[b % npair] or b < npair:
                src_new[r], pos_new[r] = src_old[b * S + p_], pos_old[b * S + p_]
            else:   # the rejected copy of a shared row is the same computation
                assert int(src_new[r]) == int(src_old[b * S + p_]) and int(pos_new[r]) == int(pos_old[b * S + p_])
    m.src_map, m.pos = src_new, pos_new
    rot = m.row_of_text.reshape(-1).clone()
    for i, r in enumerate(rot.tolist()):
        if r >= 0:
            rot[i] = dest(r // S, r % S) if r // S < m.n_seq else -1
    m.row_of_text = rot
    if getattr(m, "img_pos", None) is not None:
        feats = m.imgs_per_seq * m.P
        ip = m.img_pos.reshape(-1).clone()
        for i, p_ in enumerate(ip.tolist()):
            b = i // feats
            if p_ >= 0:
                ip[i] = -1 if (p_ < pre[b % npair] and b >= npair) else dest(b, p_)
        m.img_pos = ip
    ctx = [npair + i if pre[i] > 0 else -1 for i in range(npair)] + [-1] * npair + [npair + i if pre[i] > 0 else -1 for i in range(npair)]
    kids = [-1] * (6 * npair)
    for i in range(npair):
        if pre[i] > 0:
            kids[2 * (npair + i)], kids[2 * (npair + i) + 1] = i, 2 * npair + i
    m.att_starts = torch.tensor(att_starts, dtype=torch.int32)
    m.att_lens = torch.tensor(att_lens, dtype=torch.int32)
    m.att_ctx = torch.tensor(ctx, dtype=torch.int32)
    m.att_kids = torch.tensor(kids, dtype=torch.int32)
    m.row_starts, m.rows, m.rows_chosen = m.att_starts, rows, att_starts[2 * npair]
    m.n_att = 3 * npair
    m.range_chosen, m.range_rejected = (0, att_starts[2 * npair]), (att_starts[npair], att_starts[3 * npair])
    m.shared_rows = sum(pre)
    _c(3)
    return m


def llava_merge_index(input_ids, attention_mask, labels, n_patches, n_img_batch, imgs_per_seq, image_token, pad_token,
                      ignore_index=-100):
    INT_MIN = -(2 ** 31)
    n_seq, L = input_ids.shape
    P = n_patches
    S = L + imgs_per_seq * (P - 1)
    m = MergeIndex()
    m.n_seq, m.L, m.S, m.P, m.n_img_batch, m.imgs_per_seq = n_seq, L, S, P, n_img_batch, imgs_per_seq
    m.src_map = torch.full((n_seq * S,), INT_MIN, dtype=torch.int32)
    m.labels = torch.full((n_seq, S), ignore_index, dtype=torch.int64)
    m.mask = torch.zeros(n_seq, S, dtype=torch.int32)
    m.pos = torch.ones(n_seq * S, dtype=torch.int32)
    m.seqlens = torch.zeros(n_seq, dtype=torch.int32)
    m.img_pos = torch.zeros(n_seq * imgs_per_seq * P, dtype=torch.int32)
    m.row_of_text = torch.zeros(n_seq * (L - 1), dtype=torch.int32)
    m.target = torch.full((n_seq * (L - 1),), -100, dtype=torch.int64)
    m.status = torch.zeros(1, dtype=torch.int32)
    for b in range(n_seq):
        p = slot = 0
        base = (b % n_img_batch) * imgs_per_seq
        for j in range(L):
            t = int(input_ids[b, j])
            if t == image_token:
                if slot < imgs_per_seq and p + P <= S:
                    for i in range(P):
                        m.src_map[b * S + p + i] = -1 - ((base + slot) * P + i)
                        m.mask[b, p + i] = 1
                        m.img_pos[(b * imgs_per_seq + slot) * P + i] = p + i
                else:
                    m.status[0] = 1
                if j >= 1:
                    m.row_of_text[b * (L - 1) + j - 1] = b * S + p - 1
                p += P
                slot += 1
            else:
                if p < S:
                    m.src_map[b * S + p] = INT_MIN if t == pad_token else t
                    m.mask[b, p] = int(attention_mask[b, j])
                    m.labels[b, p] = int(labels[b, j])
                else:
                    m.status[0] = 1
                if j >= 1:
                    m.row_of_text[b * (L - 1) + j - 1] = b * S + p - 1
                    lv = int(labels[b, j])
                    m.target[b * (L - 1) + j - 1] = -100 if lv == ignore_index else lv
                p += 1
        if slot != imgs_per_seq or p != S:
            m.status[0] = 2
        run = 0
        prefix = True
        for q in range(S):
            if m.mask[b, q]:
                m.pos[b * S + q] = run
                run += 1
                if not prefix:
                    m.status[0] = 3
                m.seqlens[b] = q + 1
            else:
                prefix = False
    _c()
    return m


def llava_merge_embed(m, embed_tokens, image_features, out):
    INT_MIN = -(2 ** 31)
    s = m.src_map.long()
    out.zero_()
    t = s >= 0
    out[t] = embed_tokens[s[t]].to(out.dtype)
    im = (s < 0) & (s != INT_MIN)
    out[im] = image_features[-1 - s[im]].to(out.dtype)
    _c()
    return out


def llava_merge_bwd(m, dx, dembed_f32, dimage_features):
    INT_MIN = -(2 ** 31)
    s = m.src_map.long()
    t = s >= 0
    dembed_f32.index_add_(0, s[t], dx[t].float())
    im = (s < 0) & (s != INT_MIN)
    acc = torch.zeros(dimage_features.shape, dtype=torch.float32)
    acc.index_add_(0, -1 - s[im], dx[im].float())
    dimage_features.copy_(acc.to(dimage_features.dtype))
    _c(2)



def llavanext_merge_index(input_ids, attention_mask, labels, feat_off, total_feats, merged_len, n_img_batch,
                          imgs_per_seq, image_token, ignore_index=-100):
    INT_MIN = -(2 ** 31)
    n_seq, L = input_ids.shape
    S = int(merged_len)
    m = MergeIndex()
    m.n_seq, m.L, m.S, m.P, m.n_img_batch, m.imgs_per_seq = n_seq, L, S, -1, n_img_batch, imgs_per_seq
    m.total_feats, m.reps = int(total_feats), n_seq // n_img_batch
    m.src_map = torch.full((n_seq * S,), INT_MIN, dtype=torch.int32)
    m.labels = torch.full((n_seq, S), ignore_index, dtype=torch.int64)
    m.mask = torch.zeros(n_seq, S, dtype=torch.int32)
    m.pos = torch.ones(n_seq * S, dtype=torch.int32)
    m.seqlens = torch.zeros(n_seq, dtype=torch.int32)
    m.img_pos = torch.zeros(m.reps * m.total_feats, dtype=torch.int32)
    m.row_of_text = torch.zeros(n_seq * (L - 1), dtype=torch.int32)
    m.target = torch.full((n_seq * (L - 1),), -100, dtype=torch.int64)
    m.status = torch.zeros(1, dtype=torch.int32)
    off = feat_off.tolist()
    for b in range(n_seq):
        p = slot = 0
        seen_masked = False
        base = (b % n_img_batch) * imgs_per_seq
        rep = b // n_img_batch
        for j in range(L):
            t = int(input_ids[b, j])
            rj = b * (L - 1) + j - 1
            if int(attention_mask[b, j]) == 0:
                seen_masked = True
                if t == image_token:
                    m.status[0] = 4
                if j >= 1:
                    m.row_of_text[rj] = b * S
                continue
            if seen_masked:
                m.status[0] = 3
            if t == image_token:
                if slot < imgs_per_seq and p + off[base + slot + 1] - off[base + slot] <= S:
                    k0, F = off[base + slot], off[base + slot + 1] - off[base + slot]
                    for f in range(F):
                        m.src_map[b * S + p + f] = -1 - (k0 + f)
                        m.mask[b, p + f] = 1
                        m.img_pos[rep * m.total_feats + k0 + f] = b * S + p + f
                else:
                    F = 0
                    m.status[0] = 1
                if j >= 1:
                    m.row_of_text[rj] = b * S + max(p - 1, 0)
                p += F
                slot += 1
            else:
                if p < S:
                    m.src_map[b * S + p] = t
                    m.mask[b, p] = 1
                    m.labels[b, p] = int(labels[b, j])
                else:
                    m.status[0] = 1
                if j >= 1:
                    m.row_of_text[rj] = b * S + max(p - 1, 0)
                    lv = int(labels[b, j])
                    m.target[rj] = -100 if (lv == ignore_index or p == 0) else lv
                p += 1
        if slot != imgs_per_seq:
            m.status[0] = 2
        if p > S:
            m.status[0] = 1
        n = min(p, S)
        m.pos[b * S:b * S + n] = torch.arange(n, dtype=torch.int32)
        m.seqlens[b] = n
    _c()
    return m


def llavanext_merge_bwd(m, dx, dembed_f32, dimage_features):
    s = m.src_map.long()
    t = s >= 0
    dembed_f32.index_add_(0, s[t], dx[t].float())
    rows = m.img_pos.long().view(m.reps, m.total_feats)
    acc = torch.zeros(dimage_features.shape, dtype=torch.float32)
    for r in range(m.reps):
        acc += dx[rows[r]].float()
    dimage_features.copy_(acc.to(dimage_features.dtype))
    _c(2)


def qwen_merge_index(input_ids, attention_mask, labels, n_queries, n_img_batch, imgs_per_seq, image_start_id, ignore_index=-100):
    n_seq, L = input_ids.shape
    Q = n_queries
    m = MergeIndex()
    m.n_seq, m.L, m.S, m.P, m.n_img_batch, m.imgs_per_seq = n_seq, L, L, Q, n_img_batch, imgs_per_seq
    m.total_feats, m.reps = n_img_batch * imgs_per_seq * Q, n_seq // n_img_batch
    m.src_map = input_ids.to(torch.int32).reshape(-1).clone()
    m.labels = labels.clone()
    m.mask = attention_mask.to(torch.int32).clone()
    m.pos = torch.arange(L, dtype=torch.int32).repeat(n_seq)
    m.seqlens = torch.zeros(n_seq, dtype=torch.int32)
    m.img_pos = None
    m.row_of_text = (torch.arange(n_seq)[:, None] * L + torch.arange(L - 1)[None]).to(torch.int32).reshape(-1)
    tgt = labels[:, 1:].clone()
    tgt[tgt == ignore_index] = -100
    m.target = tgt.reshape(-1).contiguous()
    m.status = torch.zeros(1, dtype=torch.int32)
    for b in range(n_seq):
        slot, open_, prefix = 0, -1, True
        base = (b % n_img_batch) * imgs_per_seq
        for j in range(L):
            t = int(input_ids[b, j])
            if int(attention_mask[b, j]):
                if not prefix:
                    m.status[0] = 3
                m.seqlens[b] = j + 1
            else:
                prefix = False
            if t == image_start_id:
                if open_ >= 0:
                    m.status[0] = 2
                open_ = j
            elif t == image_start_id + 1:
                if open_ < 0 or j - open_ - 1 != Q or slot >= imgs_per_seq:
                    m.status[0] = 2
                else:
                    m.src_map[b * L + open_ + 1:b * L + j] = -1 - ((base + slot) * Q + torch.arange(Q, dtype=torch.int32))
                open_ = -1
                slot += 1
        if open_ >= 0 or slot != imgs_per_seq:
            m.status[0] = 2
    _c()
    return m

def _attn_ref(q, k, v, seqlens, B, S, H, KVH, dh, causal, scale):
    qf = q.float().reshape(B, S, H, dh)
    kf = k.float().reshape(B, S, KVH, dh).repeat_interleave(H // KVH, 2)
    vf = v.float().reshape(B, S, KVH, dh).repeat_interleave(H // KVH, 2)
    s = torch.einsum("bqhd,bkhd->bhqk", qf, kf) * scale
    mask = torch.zeros(B, 1, S, S, dtype=torch.bool)
    if causal:
        mask = mask | torch.ones(S, S, dtype=torch.bool).triu(1)[None, None]
    if seqlens is not None:
        mask = mask | (torch.arange(S)[None, :] >= seqlens[:, None].long())[:, None, None, :]
    s = s.masked_fill(mask, float("-inf"))
    return qf, kf, vf, s


def attn_fwd(q, k, v, out, lse, seqlens, B, S, H, KVH, head_dim, causal, scale):
    qf, kf, vf, s = _attn_ref(q, k, v, seqlens, B, S, H, KVH, head_dim, causal, scale)
    p = torch.softmax(s, -1)
    o = torch.einsum("bhqk,bkhd->bqhd", p, vf).reshape(B * S, H * head_dim)
    out.copy_(o.to(out.dtype))
    if lse is not None:
        lse.copy_(torch.logsumexp(s, -1))
    _c()
    return out


def attn_bwd(q, k, v, out, dout, lse, delta, dq, dk, dv, seqlens, B, S, H, KVH, head_dim, causal, scale):
    with torch.enable_grad():
        _attn_bwd_impl(q, k, v, dout, dq, dk, dv, seqlens, B, S, H, KVH, head_dim, causal, scale)
    _c(3)


def _attn_bwd_impl(q, k, v, dout, dq, dk, dv, seqlens, B, S, H, KVH, head_dim, causal, scale):
    qq = q.float().reshape(B, S, H, head_dim).clone().requires_grad_(True)
    kk = k.float().reshape(B, S, KVH, head_dim).clone().requires_grad_(True)
    vv = v.float().reshape(B, S, KVH, head_dim).clone().requires_grad_(True)
    kf = kk.repeat_interleave(H // KVH, 2)
    vf = vv.repeat_interleave(H // KVH, 2)
    s = torch.einsum("bqhd,bkhd->bhqk", qq, kf) * scale
    mask = torch.zeros(B, 1, S, S, dtype=torch.bool)
    if causal:
        mask = mask | torch.ones(S, S, dtype=torch.bool).triu(1)[None, None]
    if seqlens is not None:
        mask = mask | (torch.arange(S)[None, :] >= seqlens[:, None].long())[:, None, None, :]
    o = torch.einsum("bhqk,bkhd->bqhd", torch.softmax(s.masked_fill(mask, float("-inf")), -1), vf)
    (o * dout.float().reshape(B, S, H, head_dim)).sum().backward()
    dq.copy_(qq.grad.reshape(B * S, -1).to(dq.dtype))
    dk.copy_(kk.grad.reshape(B * S, -1).to(dk.dtype))
    dv.copy_(vv.grad.reshape(B * S, -1).to(dv.dtype))


def sumsq(x, out, workspace, accumulate=False):
    s = x.float().pow(2).sum()
    out[0] = out[0] + s if accumulate else s
    _c(2)
    return out


def adamw_(param, grad, master, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0,
           grad_sumsq=None, max_grad_norm=0.0):
    gs = grad_scale
    if grad_sumsq is not None and max_grad_norm > 0:
        norm = float(grad_sumsq[0].sqrt()) * grad_scale
        gs *= min(1.0, max_grad_norm / (norm + 1e-6))
    g = grad.float() * gs
    exp_avg.mul_(beta1).add_(g, alpha=1 - beta1)
    exp_avg_sq.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2s = math.sqrt(1 - beta2 ** step)
    master.mul_(1 - lr * weight_decay).addcdiv_(exp_avg, exp_avg_sq.sqrt() / bc2s + eps, value=-lr / bc1)
    param.copy_(master.to(param.dtype))
    _c()


def cast_f32_to_bf16(src, dst, scale=1.0):
    dst.copy_((src * scale).to(dst.dtype))
    _c()
    return dst


def cast_bf16_to_f32(src, dst):
    dst.copy_(src.float())
    _c()
    return dst


def _unpack_rows(t, row_starts, seqlens, B, S):
    """packed rows [sum(len), C] -> padded [B*S, C] (zeros in the padding rows)"""
    out = torch.zeros(B * S, t.shape[1], dtype=t.dtype)
    for b in range(B):
        n, r0 = int(seqlens[b]), int(row_starts[b])
        out[b * S:b * S + n] = t[r0:r0 + n]
    return out


def _pack_rows_into(dst, padded, row_starts, seqlens, B, S):
    for b in range(B):
        n, r0 = int(seqlens[b]), int(row_starts[b])
        dst[r0:r0 + n] = padded[b * S:b * S + n].to(dst.dtype)


def _attn_ctx_core(q, k, v, seqlens, row_starts, ctx, B, H, KVH, dh, scale):
    """Differentiable restatement of vlb200_attn_fwd_tc_ctx: sequence b's queries see ALL rows of sequence ctx[b] (if any), then
    their own rows causally.  q/k/v: fp32 [rows, heads*dh].  -> (out [rows, H*dh], [(b, lse [H, n])])"""
    out = torch.zeros(q.shape[0], H * dh, dtype=torch.float32)
    stats = []
    g = H // KVH
    for b in range(B):
        n, r0 = int(seqlens[b]), int(row_starts[b])
        if n == 0:
            stats.append(None)
            continue
        rows = torch.arange(r0, r0 + n)
        c = int(ctx[b])
        nc = 0
        if c >= 0:
            nc, c0 = int(seqlens[c]), int(row_starts[c])
            rows_k = torch.cat([torch.arange(c0, c0 + nc), rows])
        else:
            rows_k = rows
        qq = q[rows].view(n, H, dh).transpose(0, 1)                                 # [H, n, dh]
        kk = k[rows_k].view(-1, KVH, dh).transpose(0, 1).repeat_interleave(g, 0)     # [H, nk, dh]
        vv = v[rows_k].view(-1, KVH, dh).transpose(0, 1).repeat_interleave(g, 0)
        sc = qq @ kk.transpose(1, 2) * scale
        mask = torch.zeros(n, nc + n, dtype=torch.bool)
        mask[:, nc:] = torch.ones(n, n, dtype=torch.bool).triu(1)
        sc = sc.masked_fill(mask[None], float("-inf"))
        p_ = torch.softmax(sc, -1)
        out[rows] = (p_ @ vv).transpose(0, 1).reshape(n, H * dh)
        stats.append(torch.logsumexp(sc, -1))
    return out, stats


def attn_fwd_tc(q, k, v, out, lse, seqlens, B, S, H, KVH, head_dim, causal, scale, row_starts=None, total_rows=0, ctx=None,
                kids=None):
    """Same contract as the mma.sync kernels; row_starts: packed rows (only the attended prefix of every sequence is written,
    lse rows beyond it are left untouched -- as the CUDA kernel does)."""
    if ctx is not None:
        assert causal and row_starts is not None
        with torch.no_grad():
            o, stats = _attn_ctx_core(q.float(), k.float(), v.float(), seqlens, row_starts, ctx, B, H, KVH, head_dim, scale)
        out.copy_(o.to(out.dtype))
        if lse is not None:
            for b, st in enumerate(stats):
                if st is not None:
                    lse[b, :, :st.shape[1]] = st
        _c()
        return out
    if row_starts is None:
        return attn_fwd(q, k, v, out, lse, seqlens, B, S, H, KVH, head_dim, causal, scale)
    qp, kp, vp = (_unpack_rows(t, row_starts, seqlens, B, S) for t in (q, k, v))
    op = torch.zeros(B * S, H * head_dim, dtype=out.dtype)
    lp = torch.zeros(B, H, S) if lse is not None else None
    attn_fwd(qp, kp, vp, op, lp, seqlens, B, S, H, KVH, head_dim, causal, scale)
    _pack_rows_into(out, op, row_starts, seqlens, B, S)
    if lse is not None:
        for b in range(B):
            lse[b, :, :int(seqlens[b])] = lp[b, :, :int(seqlens[b])]
    return out


def attn_bwd_tc(q, k, v, out, dout, lse, delta, dq, dk, dv, seqlens, B, S, H, KVH, head_dim, causal, scale, row_starts=None,
                total_rows=0, ctx=None, kids=None):
    if ctx is not None:
        assert causal and row_starts is not None and kids is not None
        for b in range(B):   # kids must be the inverse of ctx (the kernel's dK/dV pass walks it)
            want = sorted(j for j in range(B) if int(ctx[j]) == b)
            assert sorted(int(x) for x in kids[2 * b:2 * b + 2] if int(x) >= 0) == want, (b, want)
        with torch.enable_grad():
            qf, kf, vf = (t.float().detach().clone().requires_grad_(True) for t in (q, k, v))
            o, _ = _attn_ctx_core(qf, kf, vf, seqlens, row_starts, ctx, B, H, KVH, head_dim, scale)
            (o * dout.float()).sum().backward()
        dq.copy_(qf.grad.to(dq.dtype)); dk.copy_(kf.grad.to(dk.dtype)); dv.copy_(vf.grad.to(dv.dtype))
        _c(3)
        return
    if row_starts is None:
        return attn_bwd(q, k, v, out, dout, lse, delta, dq, dk, dv, seqlens, B, S, H, KVH, head_dim, causal, scale)
    qp, kp, vp, dop = (_unpack_rows(t, row_starts, seqlens, B, S) for t in (q, k, v, dout))
    dqp, dkp, dvp = (torch.zeros(B * S, t.shape[1], dtype=t.dtype) for t in (dq, dk, dv))
    attn_bwd(qp, kp, vp, None, dop, lse, delta, dqp, dkp, dvp, seqlens, B, S, H, KVH, head_dim, causal, scale)
    for dst, src in ((dq, dqp), (dk, dkp), (dv, dvp)):
        _pack_rows_into(dst, src, row_starts, seqlens, B, S)


def gemm_swiglu_bwd(dy, wd, gu, act=None, a2=None, b2=None):
    ff = wd.shape[1]
    dact = gemm(dy, wd, b_kmajor=False, a2=a2, b2=b2)
    if act is not None:
        swiglu_fwd(gu, act)
    swiglu_bwd(gu, dact, out=gu)
    return gu


def colsum_f32(a, out):
    out.copy_(a.float().sum(0))
    _c(2)
    return out


def dot_f32(a, b, scale, out):
    out[0] = (a * b).sum() * scale
    _c()
    return out
