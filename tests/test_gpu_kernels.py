"""GPU numerics of the individual sm_100a kernels vs plain PyTorch fp32 references of the same op.
(bf16 storage, fp32 math: tolerances are bf16 rounding of the outputs.)"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    import vlrlhf_b200  # noqa: F401
    from vlrlhf_b200 import ops
    torch.backends.cuda.matmul.allow_tf32 = False
    return ops


def bf(t):
    return t.to(torch.bfloat16)


def close(got, want, rtol=1.6e-2, atol=1e-3):
    got, want = got.float(), want.float()
    err = (got - want).abs()
    tol = atol + rtol * want.abs()
    assert bool((err <= tol).all()), f"max err {err.max().item():.4g} (want absmax {want.abs().max().item():.4g}), " \
                                     f"bad {(err > tol).sum().item()}/{err.numel()}"


# ------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("M,N,K,ak,bk", [(128, 256, 64, 1, 1), (300, 320, 200, 1, 1), (257, 136, 72, 1, 1),
                                         (256, 512, 256, 1, 0), (256, 512, 256, 0, 1), (200, 264, 136, 0, 0),
                                         (576, 1024, 588, 1, 1), (1000, 32064, 128, 1, 1)])
def test_gemm_variants(ops, M, N, K, ak, bk):
    torch.manual_seed(M + N + K)
    lda = ((K if ak else M) + 7) // 8 * 8 + 8  # padded leading dimensions: views, not contiguous tensors
    ldb = ((K if bk else N) + 7) // 8 * 8 + 16
    a_full = bf(torch.randn((M if ak else K), lda, device="cuda") * 0.5)
    b_full = bf(torch.randn((N if bk else K), ldb, device="cuda") * 0.5)
    a = a_full[:, :(K if ak else M)]
    b = b_full[:, :(K if bk else N)]
    A = a.float() if ak else a.float().t()
    B = b.float() if bk else b.float().t()
    want = A @ B.t()
    got = ops.gemm(a, b, a_kmajor=bool(ak), b_kmajor=bool(bk), out_dtype=torch.float32)
    close(got, want, rtol=1e-3, atol=1e-3)
    got16 = ops.gemm(a, b, a_kmajor=bool(ak), b_kmajor=bool(bk))
    close(got16, want, rtol=1e-2, atol=1e-2)


def test_gemm_epilogues(ops):
    torch.manual_seed(1)
    M, N, K = 384, 768, 320
    a, b = bf(torch.randn(M, K, device="cuda") * 0.3), bf(torch.randn(N, K, device="cuda") * 0.3)
    bias, res = bf(torch.randn(N, device="cuda")), bf(torch.randn(M, N, device="cuda"))
    base = a.float() @ b.float().t()
    x = base + bias.float()
    close(ops.gemm(a, b, bias=bias, out_dtype=torch.float32), x, 1e-3, 1e-3)
    close(ops.gemm(a, b, bias=bias, act=ops.ACT_QUICK_GELU, out_dtype=torch.float32), x * torch.sigmoid(1.702 * x), 1e-3, 2e-3)
    close(ops.gemm(a, b, bias=bias, act=ops.ACT_GELU_ERF, out_dtype=torch.float32), F.gelu(x), 1e-3, 2e-3)
    close(ops.gemm(a, b, residual=res, out_dtype=torch.float32), base + res.float(), 1e-3, 1e-3)
    # in-place residual (out aliases residual) and accumulate
    r2 = res.clone()
    ops.gemm(a, b, out=r2, residual=r2)
    close(r2, base + res.float(), 1e-2, 2e-2)
    acc = torch.ones(M, N, dtype=torch.float32, device="cuda")
    ops.gemm(a, b, out=acc, accumulate=True)
    close(acc, base + 1.0, 1e-3, 1e-3)
    with pytest.raises(ValueError):
        ops.gemm(a, bf(torch.randn(N, K + 8, device="cuda")))


@pytest.mark.parametrize("M,N,K,K2,ak,bk", [(300, 320, 200, 48, 1, 1), (128, 128, 64, 16, 1, 1), (1000, 1024, 512, 128, 1, 1),
                                            (520, 768, 392, 96, 1, 0), (264, 512, 256, 64, 0, 1), (200, 264, 136, 40, 0, 0),
                                            (700, 96, 300, 16, 1, 0)])
def test_gemm_second_operand_pair_and_alpha(ops, M, N, K, K2, ak, bk):
    """D = alpha * (A B^T + A2 B2^T): the LoRA linear / its input gradient in one launch.  Ragged K and K2 (neither a
    multiple of the 64-wide k-block), column-slice operands, 1-CTA and CTA-pair tile shapes, every majorness."""
    torch.manual_seed(M + N + K + K2)
    dev = "cuda"

    def operand(rows_major, inner, outer, pad):   # a view with a padded leading dimension (a multiple of 8 elements)
        cols = inner if rows_major else outer
        full = bf(torch.randn(outer if rows_major else inner, (cols + 7) // 8 * 8 + pad, device=dev) * 0.5)
        return full[:, :cols]

    a, b = operand(ak, K, M, 8), operand(bk, K, N, 16)
    a2, b2 = operand(ak, K2, M, 24), operand(bk, K2, N, 8)
    f = lambda t, kmaj: t.float() if kmaj else t.float().t()  # noqa: E731
    want = f(a, ak) @ f(b, bk).t() + f(a2, ak) @ f(b2, bk).t()
    got = ops.gemm(a, b, a_kmajor=bool(ak), b_kmajor=bool(bk), a2=a2, b2=b2, out_dtype=torch.float32)
    close(got, want, rtol=1e-3, atol=1e-3)
    res = torch.randn(M, N, device=dev)
    got = ops.gemm(a, b, a_kmajor=bool(ak), b_kmajor=bool(bk), a2=a2, b2=b2, alpha=0.25, residual=res, out_dtype=torch.float32)
    close(got, 0.25 * want + res, rtol=1e-3, atol=1e-3)
    got16 = ops.gemm(a, b, a_kmajor=bool(ak), b_kmajor=bool(bk), a2=a2, b2=b2, alpha=2.0)
    close(got16, 2.0 * want, rtol=1e-2, atol=2e-2)
    close(ops.gemm(a, b, a_kmajor=bool(ak), b_kmajor=bool(bk), alpha=0.5, out_dtype=torch.float32), 0.5 * (f(a, ak) @ f(b, bk).t()),
          rtol=1e-3, atol=1e-3)
    with pytest.raises(ValueError):
        ops.gemm(a, b, a_kmajor=bool(ak), b_kmajor=bool(bk), a2=a2)


@pytest.mark.parametrize("M,N,K,K2,ak,bk", [(512, 128, 4096, 0, 0, 0), (128, 1024, 5000, 0, 0, 0), (384, 256, 2048, 64, 1, 1),
                                            (200, 96, 3000, 0, 1, 0), (256, 136, 2104, 0, 0, 1)])
def test_gemm_split_k_skinny_outputs(ops, M, N, K, K2, ak, bk):
    """Few output tiles, long contraction (the LoRA weight-gradient shapes): the 1-CTA kernel splits K over the SMs and a
    second kernel sums the fp32 partials in a fixed order -- same results as the unsplit path, bit-reproducible."""
    torch.manual_seed(M + N + K)
    dev = "cuda"
    f = lambda t, kmaj: t.float() if kmaj else t.float().t()  # noqa: E731
    a = bf(torch.randn((M, K) if ak else (K, M), device=dev) * 0.25)
    b = bf(torch.randn((N, K) if bk else (K, N), device=dev) * 0.25)
    want = f(a, ak) @ f(b, bk).t()
    kw = {}
    if K2:
        a2 = bf(torch.randn((M, K2) if ak else (K2, M), device=dev) * 0.25)
        b2 = bf(torch.randn((N, K2) if bk else (K2, N), device=dev) * 0.25)
        want = want + f(a2, ak) @ f(b2, bk).t()
        kw = dict(a2=a2, b2=b2)
    got = ops.gemm(a, b, a_kmajor=bool(ak), b_kmajor=bool(bk), out_dtype=torch.float32, **kw)
    close(got, want, rtol=1e-3, atol=2e-3)
    again = ops.gemm(a, b, a_kmajor=bool(ak), b_kmajor=bool(bk), out_dtype=torch.float32, **kw)
    assert torch.equal(got, again)                       # deterministic reduction order
    res = torch.randn(M, N, device=dev)
    acc = torch.full((M, N), 2.0, device=dev)
    ops.gemm(a, b, a_kmajor=bool(ak), b_kmajor=bool(bk), out=acc, alpha=0.5, residual=res, accumulate=True, **kw)
    close(acc, 0.5 * want + res + 2.0, rtol=1e-3, atol=2e-3)
    out16 = bf(torch.ones(M, N + 8, device=dev))[:, :N]   # bf16, strided destination, accumulate
    ops.gemm(a, b, a_kmajor=bool(ak), b_kmajor=bool(bk), out=out16, accumulate=True, **kw)
    close(out16, want + 1.0, rtol=1e-2, atol=3e-2)


@pytest.mark.parametrize("M,ff,K", [(600, 384, 320), (1000, 1408, 512), (256, 128, 64), (130, 256, 192), (512, 200, 128)])
def test_gemm_swiglu_epilogue_bit_equal_to_two_kernels(ops, M, ff, K):
    """act = silu(a Wg^T) * (a Wu^T) from the CTA-pair kernel's epilogue == GEMM into gate|up followed by the elementwise
    SwiGLU kernel, bit for bit (the last two shapes take the documented fallback: M < 256 / ff % 128 != 0)."""
    torch.manual_seed(M + ff + K)
    a = bf(torch.randn(M, K, device="cuda") * 0.5)
    wgu = bf(torch.randn(2 * ff, K, device="cuda") * 0.2)
    gu_ref = ops.gemm(a, wgu)
    act_ref = ops.swiglu_fwd(gu_ref)
    gu = torch.full((M, 2 * ff), 7.0, device="cuda", dtype=torch.bfloat16)
    act = torch.empty(M, ff, device="cuda", dtype=torch.bfloat16)
    ops.gemm_swiglu(a, wgu, gu, act, write_gu=True)
    assert torch.equal(act, act_ref) and torch.equal(gu, gu_ref)
    fused = M >= 256 and ff % 128 == 0
    gu.fill_(7.0)
    act.zero_()
    ops.gemm_swiglu(a, wgu, gu, act, write_gu=False)
    assert torch.equal(act, act_ref)
    if fused:
        assert bool((gu == 7.0).all())     # the projections never reached HBM
    want = F.silu(a.float() @ wgu[:ff].float().t()) * (a.float() @ wgu[ff:].float().t())
    close(act, want, rtol=2e-2, atol=2e-2)
    with pytest.raises(ValueError):
        ops.gemm_swiglu(a, wgu, gu, act[:, :-8])


# ------------------------------------------------------------------ norms
@pytest.mark.parametrize("rows,cols", [(37, 128), (1000, 4096), (5, 1024)])
def test_rmsnorm_fwd_bwd(ops, rows, cols):
    torch.manual_seed(rows)
    x = bf(torch.randn(rows, cols, device="cuda"))
    w = bf(1 + 0.1 * torch.randn(cols, device="cuda"))
    dy = bf(torch.randn(rows, cols, device="cuda"))
    dres = bf(torch.randn(rows, cols, device="cuda"))
    xf = x.float().requires_grad_(True)
    wf = w.float().requires_grad_(True)
    var = xf.pow(2).mean(-1, keepdim=True)
    y = wf * (xf * torch.rsqrt(var + 1e-5))
    rstd = torch.empty(rows, dtype=torch.float32, device="cuda")
    got = ops.rmsnorm_fwd(x, w, 1e-5, rstd=rstd)
    close(got, y.detach())
    close(rstd, torch.rsqrt(var + 1e-5).flatten().detach(), 1e-5, 1e-6)
    y.backward(dy.float())
    dw = torch.zeros(cols, dtype=torch.bfloat16, device="cuda")
    dx = ops.rmsnorm_bwd(dy, x, w, rstd, dw, dres=dres)
    close(dx, xf.grad + dres.float(), 1.6e-2, 2e-2)
    close(dw, wf.grad, 2e-2, 2e-2 * math.sqrt(rows))


def test_layernorm_and_colsum(ops):
    torch.manual_seed(0)
    x = bf(torch.randn(300, 1024, device="cuda") * 2 + 0.5)
    w, b = bf(1 + 0.1 * torch.randn(1024, device="cuda")), bf(0.1 * torch.randn(1024, device="cuda"))
    close(ops.layernorm_fwd(x, w, b, 1e-5), F.layer_norm(x.float(), (1024,), w.float(), b.float(), 1e-5))
    out = torch.zeros(1024, dtype=torch.bfloat16, device="cuda")
    ops.colsum(x, out)
    close(out, x.float().sum(0), 1e-2, 0.2)


# ------------------------------------------------------------------ rope / swiglu / gelu
def test_rope_roundtrip_and_reference(ops):
    torch.manual_seed(0)
    T, H, KV, dh = 50, 4, 2, 128
    qkv = bf(torch.randn(T, (H + 2 * KV) * dh, device="cuda"))
    pos = torch.randint(0, 2000, (T,), device="cuda", dtype=torch.int32)
    inv = 1.0 / (10000.0 ** (torch.arange(0, dh, 2, dtype=torch.int64).float() / dh))
    fr = torch.arange(2048, dtype=torch.float32)[:, None] * inv[None]
    cos_t, sin_t = fr.cos().cuda().contiguous(), fr.sin().cuda().contiguous()
    x = qkv.float().view(T, H + 2 * KV, dh)
    c = torch.cat([cos_t, cos_t], -1)[pos.long()][:, None]
    s = torch.cat([sin_t, sin_t], -1)[pos.long()][:, None]
    rot = torch.cat([-x[..., dh // 2:], x[..., :dh // 2]], -1)
    want = x.clone()
    want[:, :H + KV] = (x * c + rot * s)[:, :H + KV]
    got = qkv.clone()
    ops.rope_(got, pos, cos_t, sin_t, H + KV, dh)
    close(got.view(T, H + 2 * KV, dh), want)
    assert torch.equal(got[:, (H + KV) * dh:], qkv[:, (H + KV) * dh:])  # v untouched
    ops.rope_(got, pos, cos_t, sin_t, H + KV, dh, inverse=True)
    close(got, qkv, 2e-2, 2e-2)


def test_swiglu_gelu(ops):
    torch.manual_seed(0)
    T, ff = 77, 256
    gu = bf(torch.randn(T, 2 * ff, device="cuda") * 2)
    d = bf(torch.randn(T, ff, device="cuda"))
    g = gu[:, :ff].float().requires_grad_(True)
    u = gu[:, ff:].float().requires_grad_(True)
    y = F.silu(g) * u
    close(ops.swiglu_fwd(gu), y.detach())
    y.backward(d.float())
    close(ops.swiglu_bwd(gu, d), torch.cat([g.grad, u.grad], 1), 1.6e-2, 1e-2)
    z = bf(torch.randn(T, ff, device="cuda") * 2)
    zf = z.float().requires_grad_(True)
    h = F.gelu(zf)
    close(ops.gelu_fwd(z), h.detach())
    h.backward(d.float())
    close(ops.gelu_bwd(z, d), zf.grad, 1.6e-2, 1e-2)


# ------------------------------------------------------------------ attention
def ref_attention(q, k, v, causal, seqlens, scale):
    """q [B,S,H,dh], k/v [B,S,KV,dh] fp32 -> out, lse"""
    B, S, H, dh = q.shape
    KV = k.shape[2]
    kk = k.repeat_interleave(H // KV, dim=2)
    vv = v.repeat_interleave(H // KV, dim=2)
    s = torch.einsum("bqhd,bkhd->bhqk", q, kk) * scale
    mask = torch.zeros(B, 1, S, S, device=q.device, dtype=torch.bool)
    if causal:
        mask |= torch.ones(S, S, device=q.device, dtype=torch.bool).triu(1)[None, None]
    if seqlens is not None:
        mask |= (torch.arange(S, device=q.device)[None, :] >= seqlens[:, None].long())[:, None, None, :]
    s = s.masked_fill(mask, float("-inf"))
    p = torch.softmax(s, -1)
    return torch.einsum("bhqk,bkhd->bqhd", p, vv), torch.logsumexp(s, -1)


@pytest.mark.parametrize("B,S,H,KV,dh,causal,lens", [
    (2, 577, 4, 4, 64, False, None),          # ViT shape
    (2, 200, 2, 2, 128, True, [200, 131]),    # decoder MHA with right padding
    (3, 130, 4, 2, 128, True, [130, 64, 1]),  # GQA, ragged
    (1, 64, 2, 2, 64, True, None),
])
@pytest.mark.parametrize("fwd_variant", [5, 4, 1, 0])
def test_attention_fwd_bwd(ops, B, S, H, KV, dh, causal, lens, fwd_variant):
    """fwd_variant: every generation of the forward kernel the library ships (include/vlb200.h vlb200_set_attn_fwd_variant)."""
    prev = ops.set_attn_fwd_variant(fwd_variant)
    try:
        _attention_fwd_bwd(ops, B, S, H, KV, dh, causal, lens)
    finally:
        ops.set_attn_fwd_variant(prev)


def _attention_fwd_bwd(ops, B, S, H, KV, dh, causal, lens):
    fwd, bwd = ops.attn_fwd_tc, ops.attn_bwd_tc
    torch.manual_seed(S + H)
    dev = "cuda"
    ld = (H + 2 * KV) * dh
    qkv = bf(torch.randn(B * S, ld, device=dev))
    seqlens = torch.tensor(lens, device=dev, dtype=torch.int32) if lens is not None else None
    scale = 1.0 / math.sqrt(dh)
    q = qkv[:, :H * dh]
    k = qkv[:, H * dh:(H + KV) * dh]
    v = qkv[:, (H + KV) * dh:]
    out = torch.empty(B * S, H * dh, dtype=torch.bfloat16, device=dev)
    lse = torch.empty(B, H, S, dtype=torch.float32, device=dev)
    fwd(q, k, v, out, lse, seqlens, B, S, H, KV, dh, causal, scale)
    qf = q.float().view(B, S, H, dh).clone().requires_grad_(True)
    kf = k.float().view(B, S, KV, dh).clone().requires_grad_(True)
    vf = v.float().view(B, S, KV, dh).clone().requires_grad_(True)
    want, want_lse = ref_attention(qf, kf, vf, causal, seqlens, scale)
    valid = torch.ones(B, S, dtype=torch.bool, device=dev)
    if seqlens is not None:
        valid = torch.arange(S, device=dev)[None] < seqlens[:, None].long()
    got = out.view(B, S, H, dh).float()
    close(got[valid], want.detach()[valid], 1.6e-2, 1e-2)
    close(lse.permute(0, 2, 1)[valid], want_lse.detach().permute(0, 2, 1)[valid], 1e-3, 1e-3)
    assert torch.isfinite(got).all()  # padded query rows stay finite
    # backward: upstream gradient is zero on padded rows (as in the DPO step)
    dout = bf(torch.randn(B * S, H * dh, device=dev)) * valid.view(-1, 1)
    (want * dout.float().view(B, S, H, dh)).sum().backward()
    dqkv = torch.full_like(qkv, float("nan"))
    delta = torch.empty(B, H, S, dtype=torch.float32, device=dev)
    bwd(q, k, v, out, dout, lse, delta, dqkv[:, :H * dh], dqkv[:, H * dh:(H + KV) * dh], dqkv[:, (H + KV) * dh:],
        seqlens, B, S, H, KV, dh, causal, scale)
    assert torch.isfinite(dqkv).all()
    gq = dqkv[:, :H * dh].view(B, S, H, dh)
    gk = dqkv[:, H * dh:(H + KV) * dh].view(B, S, KV, dh)
    gv = dqkv[:, (H + KV) * dh:].view(B, S, KV, dh)
    for name, g, w in (("dq", gq, qf.grad), ("dk", gk, kf.grad), ("dv", gv, vf.grad)):
        rel = (g.float() - w).norm() / w.norm()
        assert rel < 2e-2, f"{name} rel l2 err {rel.item():.4g}"
        close(g, w, 3e-2, 3e-2)


# ------------------------------------------------------------------ merge / movers
def test_merge_index_embed_bwd(ops):
    from oracle import restate as R
    cfg = R.TINY
    batch = R.make_batch(cfg, 3, 20, 6, seed=5)
    cb = R.concatenated_inputs(batch)
    ids, am, lb = (cb[f"concatenated_{k}"] for k in ("input_ids", "attention_mask", "labels"))
    P, d = cfg.n_patches, cfg.hidden
    torch.manual_seed(0)
    emb = bf(torch.randn(cfg.vocab, d))
    img = bf(torch.randn(3 * P, d))
    img2 = torch.cat([img.view(3, P, d), img.view(3, P, d)], 0)  # the reference duplicates the images
    fe, fm, fl, fp, fmap = R.merge_input_ids_with_image_features(cfg, img2.float(), F.embedding(ids, emb.float()), ids, am, lb)
    m = ops.llava_merge_index(ids.cuda(), am.cuda(), lb.cuda(), P, 3, 1, cfg.image_token_index, cfg.pad_token_id)
    assert int(m.status.item()) == 0
    S = m.S
    assert torch.equal(m.labels.cpu(), fl)
    assert torch.equal(m.mask.cpu().long(), fm)
    assert torch.equal(m.pos.cpu().view(6, S).long(), fp)
    assert torch.equal(m.seqlens.cpu().long(), fm.sum(-1))
    assert torch.equal((m.src_map.cpu().view(6, S) < 0) & (m.src_map.cpu().view(6, S) > -(2 ** 31)), fmap)
    out = torch.empty(6 * S, d, dtype=torch.bfloat16, device="cuda")
    ops.llava_merge_embed(m, emb.cuda(), img.cuda(), out)
    assert torch.equal(out.cpu().float().view(6, S, d), fe)
    # targets / rows: shifted-label semantics of get_batch_logps on the merged sequence
    tgt = m.target.cpu().view(6, -1)
    rows = m.row_of_text.cpu().view(6, -1)
    shift = fl[:, 1:]
    for b in range(6):
        want_rows = (shift[b] != -100).nonzero().flatten() + b * S
        got_rows = rows[b][tgt[b] >= 0]
        assert torch.equal(got_rows.long(), want_rows)
        assert torch.equal(tgt[b][tgt[b] >= 0], shift[b][shift[b] != -100])
    # backward: embedding scatter-add and image-feature gradient (chosen + rejected share the image)
    dx = bf(torch.randn(6 * S, d))
    dembed = torch.zeros(cfg.vocab, d, dtype=torch.float32, device="cuda")
    dimg = torch.empty(3 * P, d, dtype=torch.bfloat16, device="cuda")
    ops.llava_merge_bwd(m, dx.cuda(), dembed, dimg)
    e = emb.float().requires_grad_(True)
    i2 = img.float().requires_grad_(True)
    fe2, *_ = R.merge_input_ids_with_image_features(cfg, torch.cat([i2.view(3, P, d)] * 2, 0), F.embedding(ids, e), ids, am, lb)
    (fe2.view(-1, d) * dx.float()).sum().backward()
    close(dembed.cpu(), e.grad, 1e-5, 1e-5)
    close(dimg.cpu(), i2.grad, 1e-2, 1e-2)
    # status codes: two image tokens in one sequence -> ragged
    bad = ids.clone()
    bad[0, 3] = cfg.image_token_index
    m2 = ops.llava_merge_index(bad.cuda(), am.cuda(), lb.cuda(), P, 3, 1, cfg.image_token_index, cfg.pad_token_id)
    assert int(m2.status.item()) != 0


def test_gather_scatter_copy_im2col(ops):
    torch.manual_seed(0)
    src = bf(torch.randn(50, 64, device="cuda"))
    idx = torch.tensor([3, -1, 49, 0, 7], device="cuda", dtype=torch.int32)
    out = torch.full((5, 64), 9.0, dtype=torch.bfloat16, device="cuda")
    ops.gather_rows(src, idx, out)
    want = src[idx.clamp(min=0).long()].clone()
    want[1] = 0
    assert torch.equal(out, want)
    dst = torch.zeros(50, 64, dtype=torch.bfloat16, device="cuda")
    ops.scatter_rows(out, idx, dst)
    assert torch.equal(dst[3], src[3]) and torch.equal(dst[49], src[49]) and float(dst[1].abs().sum()) == 0
    x = bf(torch.randn(3 * 5, 16, device="cuda"))
    y = torch.empty(3 * 4, 16, dtype=torch.bfloat16, device="cuda")
    ops.copy_rows(x, 5 * 16, 16, 1, y, 4 * 16, 16, 3, 4, 16)
    assert torch.equal(y.view(3, 4, 16), x.view(3, 5, 16)[:, 1:])
    pix = torch.randn(2, 3, 28, 28, device="cuda")
    pat = torch.zeros(2 * 4, 592, dtype=torch.bfloat16, device="cuda")
    ops.clip_im2col(pix, 14, pat)
    want = F.unfold(pix, kernel_size=14, stride=14).transpose(1, 2).reshape(8, 588)
    assert torch.equal(pat[:, :588], bf(want))


# ------------------------------------------------------------------ optimizer
def test_adamw_matches_torch(ops):
    torch.manual_seed(0)
    n = 4096 * 3
    p0 = torch.randn(n, device="cuda") * 0.05
    master = bf(p0).float()
    param = bf(p0).clone()
    m, v = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    ref_p = master.clone().requires_grad_(True)
    opt = torch.optim.AdamW([ref_p], lr=1e-3, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.01)
    ws, ss = torch.zeros(1024, device="cuda"), torch.zeros(1, device="cuda")
    for step in range(1, 4):
        g = bf(torch.randn(n, device="cuda") * (10.0 if step == 2 else 0.01))
        ops.sumsq(g, ss, ws)
        torch.testing.assert_close(ss[0], g.float().pow(2).sum(), rtol=1e-5, atol=1e-5)
        ref_p.grad = g.float().clone()
        torch.nn.utils.clip_grad_norm_([ref_p], 1.0)
        opt.step()
        ops.adamw_(param, g, master, m, v, 1e-3, 0.9, 0.98, 1e-6, 0.01, step, grad_scale=1.0, grad_sumsq=ss, max_grad_norm=1.0)
        torch.testing.assert_close(master, ref_p.detach(), rtol=2e-5, atol=1e-7)
        assert torch.equal(param, bf(master))


# ------------------------------------------------------------------ packed (ragged) rows, SURVEY.md f-2
@pytest.mark.parametrize("B,S,H,KV,dh,lens", [
    (3, 200, 2, 2, 128, [200, 131, 77]),      # tiles straddle sequence boundaries
    (4, 130, 4, 2, 128, [130, 64, 1, 129]),   # GQA, a one-token sequence
    (3, 300, 4, 4, 64, [256, 300, 128]),      # lengths on tile boundaries
    (2, 96, 2, 2, 64, [0, 96]),               # an empty sequence
])
@pytest.mark.parametrize("fwd_variant", [5, 4])
def test_attention_varlen_equals_padded(ops, B, S, H, KV, dh, lens, fwd_variant):
    prev = ops.set_attn_fwd_variant(fwd_variant)
    try:
        _attention_varlen_equals_padded(ops, B, S, H, KV, dh, lens)
    finally:
        ops.set_attn_fwd_variant(prev)


def _attention_varlen_equals_padded(ops, B, S, H, KV, dh, lens):
    """The var-len entry points (packed rows: sequence b at row_starts[b]) give, on every attended row, what the padded entry
    points give (same tiles relative to the sequence start, masked neighbours contribute exact zeros: at most an ulp apart),
    and write nothing else."""
    torch.manual_seed(S + H + dh)
    dev = "cuda"
    ld = (H + 2 * KV) * dh
    hq, hk = H * dh, KV * dh
    qkv = bf(torch.randn(B * S, ld, device=dev))
    seqlens = torch.tensor(lens, device=dev, dtype=torch.int32)
    scale = 1.0 / math.sqrt(dh)
    valid = (torch.arange(S, device=dev)[None] < seqlens[:, None].long()).view(-1)
    starts = [0]
    for n in lens:
        starts.append(starts[-1] + n)
    T = starts[-1]
    row_starts = torch.tensor(starts, device=dev, dtype=torch.int32)
    # padded run
    out = torch.empty(B * S, hq, dtype=torch.bfloat16, device=dev)
    lse = torch.empty(B, H, S, dtype=torch.float32, device=dev)
    ops.attn_fwd_tc(qkv[:, :hq], qkv[:, hq:hq + hk], qkv[:, hq + hk:], out, lse, seqlens, B, S, H, KV, dh, True, scale)
    dout = bf(torch.randn(B * S, hq, device=dev)) * valid.view(-1, 1)
    dqkv = torch.zeros_like(qkv)
    delta = torch.empty(B, H, S, dtype=torch.float32, device=dev)
    ops.attn_bwd_tc(qkv[:, :hq], qkv[:, hq:hq + hk], qkv[:, hq + hk:], out, dout, lse, delta, dqkv[:, :hq], dqkv[:, hq:hq + hk],
                    dqkv[:, hq + hk:], seqlens, B, S, H, KV, dh, True, scale)
    # packed run: NaN-poisoned outputs and statistics show any row that is written or read without belonging to a sequence
    qkv_p = qkv[valid].contiguous()
    dout_p = dout[valid].contiguous()
    out_p = torch.full((T, hq), float("nan"), dtype=torch.bfloat16, device=dev)
    lse_p = torch.full((B, H, S), float("nan"), dtype=torch.float32, device=dev)
    ops.attn_fwd_tc(qkv_p[:, :hq], qkv_p[:, hq:hq + hk], qkv_p[:, hq + hk:], out_p, lse_p, seqlens, B, S, H, KV, dh, True, scale,
                    row_starts=row_starts, total_rows=T)
    close(out_p, out[valid], 8e-3, 1e-3)
    vmask = (torch.arange(S, device=dev)[None] < seqlens[:, None].long())[:, None, :].expand(B, H, S)
    close(lse_p[vmask], lse[vmask], 1e-5, 1e-5)
    assert torch.isnan(lse_p[~vmask]).all()          # statistics of rows beyond a sequence are never written
    dqkv_p = torch.full_like(qkv_p, float("nan"))
    delta_p = torch.full((B, H, S), float("nan"), dtype=torch.float32, device=dev)
    ops.attn_bwd_tc(qkv_p[:, :hq], qkv_p[:, hq:hq + hk], qkv_p[:, hq + hk:], out_p, dout_p, lse_p, delta_p, dqkv_p[:, :hq],
                    dqkv_p[:, hq:hq + hk], dqkv_p[:, hq + hk:], seqlens, B, S, H, KV, dh, True, scale, row_starts=row_starts,
                    total_rows=T)
    assert torch.isfinite(dqkv_p).all()               # ... nor read
    close(delta_p[vmask], delta[vmask], 1e-5, 1e-5)
    close(dqkv_p, dqkv[valid], 8e-3, 1e-3)


def test_pack_merge_rows_equals_mirror(ops):
    """vlb200_pack_merge_rows == the plain-Python mirror (tests/mock_ops.py) for the LLaVA-1.5 and LLaVA-Next merge indices,
    and the packed merge (embedding rows, image-gradient rows) equals the padded one on the surviving rows."""
    from oracle import restate as R
    from tests import mock_ops
    from vlrlhf_b200 import host
    cfg = R.TINY
    batch = R.make_batch(cfg, 3, 20, 6, seed=5, ddpo_like=True)
    cb = R.concatenated_inputs(batch)
    ids, am, lb = (cb[f"concatenated_{k}"] for k in ("input_ids", "attention_mask", "labels"))
    am[4, 15:] = 0   # ragged right padding
    lb[4, 15:] = -100
    am[1, 18:] = 0
    lb[1, 18:] = -100
    P, d = cfg.n_patches, cfg.hidden
    lens = host.merged_seq_lens(ids, am, cfg.image_token_index, P)
    m_pad = ops.llava_merge_index(ids.cuda(), am.cuda(), lb.cuda(), P, 3, 1, cfg.image_token_index, cfg.pad_token_id)
    assert m_pad.seqlens.cpu().tolist() == lens
    m = ops.pack_merge_rows(ops.llava_merge_index(ids.cuda(), am.cuda(), lb.cuda(), P, 3, 1, cfg.image_token_index, cfg.pad_token_id), lens)
    want = mock_ops.pack_merge_rows(mock_ops.llava_merge_index(ids, am, lb, P, 3, 1, cfg.image_token_index, cfg.pad_token_id), lens)
    assert m.T == want.T == sum(lens) and m.T_chosen == want.T_chosen == sum(lens[:3]) and m.row_stride == 0
    for k in ("src_map", "pos", "row_of_text", "img_pos", "row_starts"):
        assert torch.equal(getattr(m, k).cpu().reshape(-1), getattr(want, k).reshape(-1)), k
    torch.manual_seed(0)
    emb, img = bf(torch.randn(cfg.vocab, d)).cuda(), bf(torch.randn(3 * P, d)).cuda()
    S = m_pad.S
    keep = torch.cat([torch.arange(b * S, b * S + n) for b, n in enumerate(lens)]).cuda()
    x_pad = torch.empty(6 * S, d, dtype=torch.float32, device="cuda")
    x = torch.empty(m.T, d, dtype=torch.float32, device="cuda")
    ops.llava_merge_embed(m_pad, emb, img, x_pad)
    ops.llava_merge_embed(m, emb, img, x)
    assert torch.equal(x, x_pad[keep])
    dx_pad = bf(torch.randn(6 * S, d)).cuda()
    dx_pad[torch.ones(6 * S, dtype=torch.bool, device="cuda").index_fill_(0, keep, False)] = 0  # padding rows carry no gradient
    de_pad, de = (torch.zeros(cfg.vocab, d, dtype=torch.float32, device="cuda") for _ in range(2))
    di_pad, di = (torch.empty(3 * P, d, dtype=torch.bfloat16, device="cuda") for _ in range(2))
    ops.llava_merge_bwd(m_pad, dx_pad, de_pad, di_pad)
    ops.llava_merge_bwd(m, dx_pad[keep].contiguous(), de, di)
    assert torch.equal(di, di_pad)
    close(de, de_pad, 1e-5, 1e-5)   # fp32 atomics: order differs
    with pytest.raises(ValueError):
        ops.pack_merge_rows(m, lens)  # already packed
    # LLaVA-Next layout: img_pos holds flat merged rows
    ncfg = R.TINY_NEXT
    sizes = [(28, 28), (20, 50), (60, 25)]
    nb = R.make_batch(ncfg, 3, 20, 6, seed=7, ddpo_like=True, image_sizes=sizes)
    ncb = R.concatenated_inputs(nb)
    nids, nam, nlb = (ncb[f"concatenated_{k}"] for k in ("input_ids", "attention_mask", "labels"))
    plan = host.anyres_pack_index(torch.tensor(sizes), ncfg.image_grid_pinpoints, ncfg.image_size, ncfg.patch_size)
    SN = host.next_merged_len(nids, nam, plan.feature_lens, ncfg.image_token_index)
    nlens = host.merged_seq_lens(nids, nam, ncfg.image_token_index, [plan.feature_lens[b % 3] for b in range(6)])
    args = (plan.feat_off, plan.total_feats, SN, 3, 1, ncfg.image_token_index)
    mn = ops.llavanext_merge_index(nids.cuda(), nam.cuda(), nlb.cuda(), plan.feat_off.cuda(), *args[1:])
    assert mn.seqlens.cpu().tolist() == nlens
    mn = ops.pack_merge_rows(mn, nlens)
    wn = mock_ops.pack_merge_rows(mock_ops.llavanext_merge_index(nids, nam, nlb, *args), nlens)
    for k in ("src_map", "pos", "row_of_text", "img_pos", "row_starts"):
        assert torch.equal(getattr(mn, k).cpu().reshape(-1), getattr(wn, k).reshape(-1)), k


@pytest.mark.parametrize("M,K,ff,r", [(600, 512, 1024, 0), (1000, 256, 768, 16), (130, 512, 256, 0)])
def test_gemm_swiglu_bwd_fused_is_bit_identical(ops, M, K, ff, r):
    """vlb200_gemm_swiglu_bwd_bf16 (SwiGLU backward + act recompute in the dgrad GEMM's epilogue) == gemm -> swiglu_fwd + swiglu_bwd,
    bit for bit, on CTA-pair, single-CTA and ragged-edge shapes, with and without the LoRA operand pair."""
    g = torch.Generator(device="cuda").manual_seed(M + ff)
    rn = lambda *s: (torch.randn(*s, device="cuda", generator=g) * 0.5).to(torch.bfloat16)  # noqa: E731
    dy, wd, gu = rn(M, K), rn(K, ff), rn(M, 2 * ff)
    a2, b2 = (rn(M, r), rn(r, ff)) if r else (None, None)
    dact = ops.gemm(dy, wd, b_kmajor=False, a2=a2, b2=b2)
    act_want = ops.swiglu_fwd(gu)
    dgu_want = ops.swiglu_bwd(gu, dact)
    gu1, act1 = gu.clone(), torch.full((M, ff), float("nan"), dtype=torch.bfloat16, device="cuda")
    ops.gemm_swiglu_bwd(dy, wd, gu1, act1, a2=a2, b2=b2)
    assert torch.equal(gu1, dgu_want) and torch.equal(act1, act_want)
    gu2 = gu.clone()
    ops.gemm_swiglu_bwd(dy, wd, gu2, None, a2=a2, b2=b2)          # act not wanted (checkpointed LoRA backward)
    assert torch.equal(gu2, dgu_want)
