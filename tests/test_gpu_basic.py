"""GPU parity: synthetic init, log-prob gather (K16) and preference loss (K18) vs the oracle / golden vectors.
All calls go through the C ABI (vlrlhf_b200.ops -> libvlb200.so)."""
import os

import numpy as np
import pytest
import torch

from oracle import restate as R

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def ops():
    import vlrlhf_b200  # noqa: F401
    from vlrlhf_b200 import ops
    return ops


def test_init_uniform_bit_exact(ops):
    for n, seed, scale, shift in [(1000, 5, 0.0346, 0.0), (1 << 20, 77, 0.1, 1.0), (12345, 0xDEADBEEF, 1.7320508, 0.0)]:
        want = R.hash_uniform(n, seed, scale, shift)
        got32 = ops.init_uniform_(torch.empty(n, dtype=torch.float32, device="cuda"), seed, scale, shift).cpu()
        assert torch.equal(got32, want)
        got16 = ops.init_uniform_(torch.empty(n, dtype=torch.bfloat16, device="cuda"), seed, scale, shift).cpu()
        assert torch.equal(got16, want.to(torch.bfloat16))


def test_perturb_bit_exact(ops):
    n = 50000
    base = R.bf16_round(R.hash_uniform(n, 1, 0.05, 1.0))
    other = R.bf16_round(R.hash_uniform(n, 2, 0.1, 1.0))
    want = R.bf16_round(base + 0.05 * (other - 1.0)).to(torch.bfloat16)
    got = ops.perturb_(torch.empty(n, dtype=torch.bfloat16, device="cuda"), base.to(torch.bfloat16).cuda(),
                       other.to(torch.bfloat16).cuda(), 0.05, 1.0).cpu()
    assert torch.equal(got, want)


def _shift(logits, labels):
    """reference shift (trainer.py:161-162): row (b,t) predicts labels[b,t+1]"""
    B2, S, V = logits.shape
    return logits[:, :-1, :].reshape(B2 * (S - 1), V).contiguous(), labels[:, 1:].reshape(-1).contiguous()


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_logps_fwd_golden(ops, tag):
    d = np.load(os.path.join(G, "g1_logps.npz"))
    logits, labels = torch.from_numpy(d[f"{tag}_logits"]), torch.from_numpy(d[f"{tag}_labels"])
    B2 = logits.shape[0]
    lg, tg = _shift(logits, labels)
    for avg, key in ((False, "sum"), (True, "avg")):
        got, per_tok, lse = ops.logps_fwd(lg.cuda(), tg.cuda(), B2, average_log_prob=avg)
        np.testing.assert_allclose(got.cpu().numpy(), d[f"{tag}_{key}"], rtol=2e-6, atol=2e-4)
    # bf16 logits, fp32 math: must match the reference run on the upcast bf16 logits
    got, _, _ = ops.logps_fwd(lg.to(torch.bfloat16).cuda(), tg.cuda(), B2)
    np.testing.assert_allclose(got.cpu().numpy(), d[f"{tag}_sum_bf16in_fp32math"], rtol=2e-6, atol=2e-4)
    # per-token values vs the oracle
    pt, mask = R.get_batch_logps(logits, labels, return_per_token=True)
    got, per_tok, lse = ops.logps_fwd(lg.cuda(), tg.cuda(), B2)
    np.testing.assert_allclose(per_tok.cpu().view(B2, -1).numpy(), (pt * mask).numpy(), rtol=1e-5, atol=1e-4)


def test_logps_ragged_vocab_and_mask(ops):
    g = torch.Generator().manual_seed(0)
    B2, S, V = 4, 9, 1003  # V not a multiple of the vector width; row stride padded to 1008
    logits = torch.randn(B2, S, V, generator=g) * 4
    labels = torch.randint(0, V, (B2, S), generator=g)
    labels[:, :3] = -100
    want_pt, mask = R.get_batch_logps(logits, labels, return_per_token=True)
    w = (torch.rand(B2, S - 1, generator=g) > 0.4)
    want = (want_pt * (mask & w)).sum(-1)
    buf = torch.zeros(B2 * (S - 1), 1008)
    lg, tg = _shift(logits, labels)
    buf[:, :V] = lg
    got, _, _ = ops.logps_fwd(buf.cuda()[:, :V], tg.cuda(), B2, weight=w.reshape(-1).to(torch.uint8).cuda())
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=1e-5, atol=1e-4)
    # empty sequence (all labels masked): sum -> 0 (reference: (x*mask).sum = 0)
    labels[1, :] = -100
    lg, tg = _shift(logits, labels)
    got, _, _ = ops.logps_fwd(lg.cuda(), tg.cuda(), B2)
    assert got[1].item() == 0.0
    with pytest.raises(ValueError):
        ops.logps_fwd(lg.cuda(), tg[:-1].cuda(), B2)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("avg", [False, True])
def test_logps_bwd_vs_autograd(ops, dtype, avg):
    g = torch.Generator().manual_seed(3)
    B2, S, V = 4, 12, 2056
    logits = (torch.randn(B2, S, V, generator=g) * 3).to(dtype).float()
    labels = torch.randint(0, V, (B2, S), generator=g)
    labels[:, :4] = -100
    labels[2, 9:] = -100
    gl = torch.randn(B2, generator=g)
    x = logits.clone().requires_grad_(True)
    R.get_batch_logps(x, labels, average_log_prob=avg).backward(gl)
    want = x.grad[:, :-1, :].reshape(-1, V)
    lg, tg = _shift(logits, labels)
    lgc = lg.to(dtype).cuda()
    _, _, lse = ops.logps_fwd(lgc, tg.cuda(), B2, average_log_prob=avg)
    got = ops.logps_bwd(lgc, tg.cuda(), B2, lse, gl.cuda(), average_log_prob=avg).float().cpu()
    # output is bf16: 2^-8 relative per element
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=1e-2, atol=1e-6)
    assert torch.all(got[tg < 0] == 0)


def test_dpo_loss_golden_and_grad(ops):
    d = np.load(os.path.join(G, "g2_loss.npz"))
    pc, pr, rc, rr = (torch.from_numpy(d[k]) for k in ("pc", "pr", "rc", "rr"))
    pol, ref = torch.cat([pc, pr]).cuda(), torch.cat([rc, rr]).cuda()
    for lt in ("sigmoid", "ddpo", "hinge", "ipo", "kto_pair"):
        for ls in (0.0, 0.1):
            for rf in (False, True):
                k = f"{lt}_ls{ls}_rf{int(rf)}"
                losses, cr, rj, stats, grad = ops.dpo_loss(pol, ref, 0.1, ls, lt, rf)
                np.testing.assert_allclose(losses.cpu().numpy(), d[k + "_losses"], rtol=2e-5, atol=2e-6)
                np.testing.assert_allclose(cr.cpu().numpy(), d[k + "_cr"], rtol=1e-6, atol=1e-6)
                np.testing.assert_allclose(rj.cpu().numpy(), d[k + "_rr"], rtol=1e-6, atol=1e-6)
                st = stats.cpu().numpy()
                np.testing.assert_allclose(st[0], d[k + "_losses"].mean(), rtol=2e-5)
                np.testing.assert_allclose(st[1], (d[k + "_cr"] > d[k + "_rr"]).mean(), rtol=1e-6)
                np.testing.assert_allclose(st[4], (d[k + "_cr"] - d[k + "_rr"]).mean(), rtol=1e-4, atol=1e-5)
                # gradient of mean(losses) wrt policy logps vs autograd through the oracle
                a, b = pc.clone().requires_grad_(True), pr.clone().requires_grad_(True)
                R.dpo_loss(a, b, rc, rr, 0.1, ls, lt, rf)[0].mean().backward()
                want = torch.cat([a.grad, b.grad]).numpy()
                np.testing.assert_allclose(grad.cpu().numpy(), want, rtol=2e-4, atol=1e-7)
    with pytest.raises(ValueError):
        ops.dpo_loss(pol, ref, 0.1, 0.0, "nope")
