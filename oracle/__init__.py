"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference hot path (TideDra/VL-RLHF
`VLDPOTrainer.concatenated_forward -> get_batch_logps -> dpo_loss`, plus the
LLaVA-1.5 forward it drives).  Nothing in the product package
(`vlrlhf_b200/`) may import this package: only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs do, and there only as the checker.

Parity status: the reference ships NO golden vectors / tests, so
the restatement is pinned against outputs of the reference's own functions
executed in the build container (`oracle/ref_shim.py` imports them from
/root/reference/src with stub trl/peft/accelerate/deepspeed modules) -- the
generated vectors live in `tests/golden/` together with the script that made
them (`oracle/make_fixtures.py`).  The TRL-0.8.1 half of the path
(`concatenated_inputs`, `get_batch_loss_metrics`) is NOT on disk and is
restated from its published algorithm: that part is "parity unpinned".
"""
