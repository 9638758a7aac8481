"""__graft_entry__.smoke(): one tiny DPO step (policy fwd, reference fwd, loss, backward, AdamW) through the
CUDA path on cuda:0, checked against the oracle (the oracle is the checker here, never the thing run)."""
from __future__ import annotations

import numpy as np
import torch


def run() -> None:
    from . import config, engine, host, ops
    from oracle import restate as R  # checker only
    cfg, rcfg = config.SMALL, R.SMALL
    eng = engine.LlavaDPOEngine(cfg, config.TrainConfig(learning_rate=1e-3))
    eng.init_synthetic(0)
    batch = R.make_batch(rcfg, 2, 96, 24, 0, ddpo_like=True)
    cb = host.concatenated_inputs(batch)
    ids, am, lb, px, _ = eng.prepare_inputs(cb["concatenated_input_ids"], cb["concatenated_attention_mask"],
                                            cb["concatenated_labels"], cb["concatenated_img_input_dict"]["pixel_values"])
    n0 = ops.launch_count()
    out = eng.step(ids, am, lb, px, train=True)
    torch.cuda.synchronize()
    wp, wr = R.make_policy_and_ref(rcfg, 0)
    with torch.no_grad():
        loss, _, aux = R.get_batch_loss_metrics(rcfg, wp, wr, batch)
    want = torch.cat([aux["policy_chosen_logps"], aux["policy_rejected_logps"]]).numpy()
    got = out.policy_logps.cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=1e-3)
    assert abs(float(out.stats[0]) - float(loss)) < 5e-2, (float(out.stats[0]), float(loss))
    eng.wait_optimizer()
    assert torch.isfinite(eng.master).all()
    print(f"smoke ok: logps {got.round(3).tolist()} loss {float(out.stats[0]):.5f} (oracle {float(loss):.5f}); "
          f"{ops.launch_count() - n0} vlb200 kernel launches")
