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
    eng = engine.LlavaDPOEngine(config.TINY, config.TrainConfig(learning_rate=1e-3), device="cpu")
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
    torch.save({"local": local, "summed": summed, "params": eng.params.clone(), "world": eng.world_size(),
                "sumsq": eng.grad_sumsq.clone(), "params_overlap": eng2.params.clone(), "grads_overlap": eng2.grads.clone()},
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
