"""World-size-2 data-parallel path on CPU (gloo): preference pairs shard across ranks, ONE all-reduce over the flat
bf16 gradient buffer, 1/world folded into AdamW.  Runs the engine over tests/mock_ops.py (no CUDA here)."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import importlib
    import vlrlhf_b200  # noqa: F401
    from tests import mock_ops
    sys.modules["vlrlhf_b200.ops"] = mock_ops
    engine = importlib.import_module("vlrlhf_b200.engine")
    from vlrlhf_b200 import config, host
    from oracle import restate as R
    torch.set_num_threads(2)
    os.environ["VLB200_SHARD_OPTIMIZER"] = "0"   # replicated AdamW first: the plain all-reduce path
    eng = engine.LlavaDPOEngine(config.TINY, config.TrainConfig(learning_rate=1e-3), device="cpu")
    assert not eng.shard_optimizer
    eng.init_synthetic(0)  # same weights on every rank
    batch = R.make_batch(R.TINY, 2, 24, 8, seed=100 + rank)  # rank-local pairs
    cb = host.concatenated_inputs(batch)
    a = eng.prepare_inputs(cb["concatenated_input_ids"], cb["concatenated_attention_mask"], cb["concatenated_labels"],
                           cb["concatenated_img_input_dict"]["pixel_values"])
    # local gradient (before the all-reduce) for the cross-check
    eng.overlap_allreduce = False
    pol, m, feats = eng.forward_logps(*a[:4], which="policy", save=True)
    ref, _, _ = eng.forward_logps(*a[:4], which="ref", save=False, feats=feats, m=m)
    _, _, _, _, grad = mock_ops.dpo_loss(pol, ref, 0.1)
    eng._backward(grad)
    local = eng.grads.clone()
    eng.allreduce_grads()
    summed = eng.grads.clone()
    eng.optimizer_step()
    # same step again from the same state with the per-layer buckets overlapped with backward: identical sums
    eng2 = engine.LlavaDPOEngine(config.TINY, config.TrainConfig(learning_rate=1e-3), device="cpu")
    eng2.init_synthetic(0)
    eng2.overlap_allreduce = True
    eng2.step(*a[:4], train=True)
    # optimizer sharded over the ranks (the default when world > 1): reduce-scatter, AdamW on the own slice, all-gather
    os.environ["VLB200_SHARD_OPTIMIZER"] = "1"
    eng3 = engine.LlavaDPOEngine(config.TINY, config.TrainConfig(learning_rate=1e-3), device="cpu")
    assert eng3.shard_optimizer and eng3.n_flat % (2 * engine.ALIGN) == 0
    assert eng3.master.numel() == eng3.n_flat // 2 and eng3.shard_lo == rank * eng3.n_flat // 2
    eng3.init_synthetic(0)
    eng3.step(*a[:4], train=True)
    # packed rows (TrainConfig.pack_sequences): every rank drops its own padding rows -- a different row count per rank, no
    # collective depends on it -- and lands on the same parameters as the padded sharded step
    eng4 = engine.LlavaDPOEngine(config.TINY, config.TrainConfig(learning_rate=1e-3, pack_sequences=True), device="cpu")
    eng4.init_synthetic(0)
    eng4.step(*a[:4], train=True)
    packed_rows = eng4._saved["m"].T
    # shared-prefix rows (TrainConfig.share_prefix) under the sharded step: every rank lays out its own pairs' common prefixes
    # once -- again a rank-local row count no collective depends on; the reduced gradient slice equals the padded step's up to
    # the accumulation order of the rows
    eng5 = engine.LlavaDPOEngine(config.TINY, config.TrainConfig(learning_rate=1e-3, share_prefix=True), device="cpu")
    eng5.init_synthetic(0)
    plan = eng5.host_row_plan(cb["concatenated_input_ids"], cb["concatenated_attention_mask"])
    eng5.step(*a[:4], train=True, **plan)
    shared_rows = eng5._saved["m"].T
    n = eng.params.numel()
    torch.save({"params_packed": eng4.params[:n].clone(), "packed_rows": packed_rows, "shared_rows": shared_rows,
                "params_shared": eng5.params[:n].clone(), "grads_shared_own": eng5.grads[eng5.shard_lo:eng5.shard_hi].clone(),
                "padded_rows": eng4._saved["m"].n_seq * eng4._saved["m"].S, "local": local, "summed": summed, "params": eng.params.clone(), "world": eng.world_size(),
                "sumsq": eng.grad_sumsq.clone(), "params_overlap": eng2.params.clone(), "grads_overlap": eng2.grads.clone(),
                "params_sharded": eng3.params[:n].clone(), "sumsq_sharded": eng3.grad_sumsq.clone(),
                "grads_sharded_own": eng3.grads[eng3.shard_lo:eng3.shard_hi].clone(), "shard": (eng3.shard_lo, eng3.shard_hi),
                "master_sharded": eng3.master.clone(), "master": eng.master.clone()},
               os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_data_parallel_step(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(tmp_path / "rank0.pt")
    r1 = torch.load(tmp_path / "rank1.pt")
    assert r0["world"] == 2
    # the all-reduced buffer is the sum of the two local gradients (bf16 rounding of the sum)
    want = (r0["local"].float() + r1["local"].float())
    got = r0["summed"].float()
    assert torch.equal(r0["summed"], r1["summed"])
    denom = want.abs().max().item()
    assert (got - want).abs().max().item() <= 0.01 * denom + 1e-6
    assert not torch.equal(r0["local"], r1["local"])  # different pairs on each rank
    # identical replicas after the step
    assert torch.equal(r0["params"], r1["params"])
    assert torch.equal(r0["sumsq"], r1["sumsq"])
    # bucketed + overlapped all-reduce gives the same reduced gradients and the same parameters
    assert torch.equal(r0["grads_overlap"], r0["summed"]) and torch.equal(r1["grads_overlap"], r0["summed"])
    assert torch.equal(r0["params_overlap"], r0["params"]) and torch.equal(r1["params_overlap"], r1["params"])
    # sharded optimizer: replicas identical after the all-gather, each rank's reduced slice == the all-reduced buffer,
    # and the update equals the replicated one (the gradient-norm scalar is summed in a different order: 1-ulp slack)
    assert torch.equal(r0["params_sharded"], r1["params_sharded"])
    n = r0["summed"].numel()
    for r in (r0, r1):
        lo, hi = r["shard"]
        hi_c = min(hi, n)
        assert torch.equal(r["grads_sharded_own"][:hi_c - lo], r0["summed"][lo:hi_c])
        torch.testing.assert_close(r["master_sharded"][:hi_c - lo], r["master"][lo:hi_c], rtol=1e-6, atol=1e-9)
    assert torch.equal(r0["sumsq_sharded"], r1["sumsq_sharded"])
    torch.testing.assert_close(r0["sumsq_sharded"], r0["sumsq"], rtol=1e-5, atol=0)
    same = (r0["params_sharded"] == r0["params"]).float().mean().item()
    assert same > 0.9999, same
    # packed rows: identical replicas, identical to the padded sharded step
    assert torch.equal(r0["params_packed"], r1["params_packed"]) and torch.equal(r0["params_packed"], r0["params_sharded"])
    assert r0["packed_rows"] < r0["padded_rows"] or r1["packed_rows"] < r1["padded_rows"]
    # shared-prefix rows: identical replicas after the all-gather; each rank's reduced gradient slice is the padded step's up to
    # the accumulation order (fewer rows: the common prefix of every pair once)
    assert torch.equal(r0["params_shared"], r1["params_shared"])
    assert r0["shared_rows"] < r0["packed_rows"] and r1["shared_rows"] < r1["packed_rows"]
    for r in (r0, r1):
        g_s, g_p = r["grads_shared_own"].float(), r["grads_sharded_own"].float()
        assert float((g_s - g_p).norm() / g_p.norm().clamp_min(1e-12)) < 2e-2


def _qwen_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import importlib
    import vlrlhf_b200  # noqa: F401
    from tests import mock_ops
    sys.modules["vlrlhf_b200.ops"] = mock_ops
    vlrlhf_b200.ops = mock_ops
    EQ = importlib.import_module("vlrlhf_b200.engine_qwen")
    from vlrlhf_b200 import config
    from oracle import qwen_restate as Q
    torch.set_num_threads(2)
    out = {}
    for shard in ("0", "1"):
        os.environ["VLB200_SHARD_OPTIMIZER"] = shard
        eng = EQ.QwenVLDPOEngine(config.TINY_QWEN, config.TrainConfig(learning_rate=1e-3), device="cpu")
        assert eng.shard_optimizer == (shard == "1")
        eng.init_synthetic(0)
        batch = Q.make_batch(Q.TINY_QWEN, 2, 48, 24, seed=100 + rank)  # rank-local pairs
        eng.train_step(batch, train=True)
        out[f"params{shard}"] = eng.params[: eng.layout.size].clone()
        out[f"base{shard}"] = eng.bparams.clone()
        if shard == "1":   # the same sharded step with shared-prefix rows (rank-local row plan from the host batch)
            out["grads_own"] = eng.grads[eng.shard_lo:eng.shard_hi].clone()
            eng_s = EQ.QwenVLDPOEngine(config.TINY_QWEN, config.TrainConfig(learning_rate=1e-3, share_prefix=True), device="cpu")
            eng_s.init_synthetic(0)
            eng_s.train_step(batch, train=True)
            assert eng_s._saved["m"].shared and eng_s._saved["m"].shared_rows > 0
            out["params_shared"] = eng_s.params[: eng_s.layout.size].clone()
            out["grads_shared_own"] = eng_s.grads[eng_s.shard_lo:eng_s.shard_hi].clone()
    torch.save(out, os.path.join(out_dir, f"qrank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_qwen_lora_step(tmp_path):
    """Qwen-VL + LoRA under data parallelism: only the adapter arena is reduced / sharded; replicas stay identical and the
    sharded update equals the replicated one."""
    port = _free_port()
    mp.spawn(_qwen_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / "qrank0.pt"), torch.load(tmp_path / "qrank1.pt")
    for s in ("0", "1"):
        assert torch.equal(r0[f"params{s}"], r1[f"params{s}"])
        assert torch.equal(r0[f"base{s}"], r1[f"base{s}"])
    assert (r0["params0"] == r0["params1"]).float().mean().item() > 0.9999
    # shared-prefix rows under the sharded step: identical replicas; the rank's reduced gradient slice is the padded step's up to
    # the accumulation order
    assert torch.equal(r0["params_shared"], r1["params_shared"])
    for r in (r0, r1):
        g_s, g_p = r["grads_shared_own"].float(), r["grads_own"].float()
        assert float((g_s - g_p).norm() / g_p.norm().clamp_min(1e-12)) < 2e-2
