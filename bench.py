#!/usr/bin/env python
"""bench.py -- preference-pairs/sec of one full DPO optimisation step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] ...    # the reference's PyTorch-CPU path (port)
    python bench.py --impl library   [--steps K]                   # stock transformers + SDPA + fused AdamW on the same GPU

Workload (BASELINE.json configs[1]): LLaVA-1.5-7B DPO, bf16, 4 pairs/GPU, text seq 1024 (+575 image
positions -> 1599 decoder tokens/sequence), 1x336-px image per pair, full fine-tune of projector + LLM,
frozen CLIP tower, separate frozen reference copy, AdamW + grad-norm clipping, synthetic data, seeded
random-init weights.  A step = policy fwd + reference fwd + loss + backward + grad all-reduce + AdamW.
Prints ONE JSON line (rank 0).  Timing: CUDA events on the launching stream, barrier + synchronize on both
sides, max over ranks; W >= 3 warm-up steps; every step streams > 100 GB of weights/activations/optimizer
state through HBM, far beyond the 126 MB L2, so no explicit L2 flush is needed ("l2": "inputs>>L2").
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "preference_pairs_per_sec"
UNIT = "pairs/s"
PAIRS_PER_GPU = 4
TEXT_LEN = 1024
PROMPT_LEN = 128
WORKLOAD = "LLaVA-1.5-7B DPO bf16 full-FT, 4 pairs/GPU, text 1024 (1599 merged), 1x336px image/pair"
# (bench_library.py and the CPU arm run the same workload string)
WORKLOAD_NEXT = ("LLaVA-Next-Mistral-7B DDPO bf16 full-FT (configs[3], side measurement), 4 pairs/GPU, text 1024, 1x336px image "
                 "-> 3 anyres crops -> 1176 packed image tokens (2199 merged), activation checkpointing")


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def _traffic():
    """(DRAM bytes per launch of the dominant kernel, provenance) from the committed `ncu --set full` capture (profiles/)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        return t["dram_bytes_read"] + t["dram_bytes_write"], {"algorithmic_bytes": t["algorithmic_bytes"], "source": t["source"]}
    except Exception:
        return None, None


def step_flops(cfg, n_pairs, S, rows_lm):
    """Algorithmic FLOPs of one step: policy fwd + 2x bwd + reference fwd; ViT once per pair."""
    d, ff, L = cfg.hidden, cfg.ff, cfg.layers
    p_layer = cfg.qkv_dim * d + d * cfg.heads * cfg.head_dim + 3 * d * ff
    per_seq = S * 2 * p_layer * L + L * 2 * S * S * d
    lm = rows_lm * 2 * d * cfg.vocab
    dv, Sv = cfg.v_hidden, cfg.n_patches + 1
    vit = cfg.v_used_layers * (Sv * 2 * (4 * dv * dv + 2 * dv * cfg.v_ff) + 4 * Sv * Sv * dv) + cfg.n_patches * 2 * cfg.patch_k * dv
    proj = cfg.n_patches * 2 * (dv * d + d * d)
    crops = getattr(cfg, "_bench_crops_per_image", 1)  # LLaVA-Next: the tower and projector run once per anyres crop
    return n_pairs * (2 * per_seq * 4 + crops * (vit + proj * 4)) + lm * 4


def executed_flops(cfg, n_pairs, S, rows_lm, plan):
    """FLOPs the step actually EXECUTES under a row plan (packed / shared-prefix rows): the linear layers see sum(rows), a
    sequence of n own rows with c context rows costs 4*d*(n*n/2 + n*c) per layer and pass in attention.  No plan: step_flops."""
    if not plan or "seq_lens" not in plan:
        return step_flops(cfg, n_pairs, S, rows_lm)
    d, ff, L = cfg.hidden, cfg.ff, cfg.layers
    lens = plan["seq_lens"]
    pre = plan.get("prefix_rows") or [0] * n_pairs
    p_layer = cfg.qkv_dim * d + d * cfg.heads * cfg.head_dim + 3 * d * ff
    rows, attn = 0, 0.0
    for i in range(n_pairs):
        p_, c, r = pre[i], lens[i] - pre[i], lens[n_pairs + i] - pre[i]
        rows += p_ + c + r
        attn += 4.0 * d * (p_ * p_ / 2 + c * c / 2 + c * p_ + r * r / 2 + r * p_)
    lin = rows * 2 * p_layer * L
    lm = rows_lm * 2 * d * cfg.vocab
    dv, Sv = cfg.v_hidden, cfg.n_patches + 1
    vit = cfg.v_used_layers * (Sv * 2 * (4 * dv * dv + 2 * dv * cfg.v_ff) + 4 * Sv * Sv * dv) + cfg.n_patches * 2 * cfg.patch_k * dv
    proj = cfg.n_patches * 2 * (dv * d + d * d)
    return (lin + L * attn) * 4 + n_pairs * (vit + proj * 4) + lm * 4


def plan_ratios(plan, n_pairs, S):
    """(rows executed / padded rows, attention work executed / padded) of a row plan (packed or shared-prefix rows)."""
    if not plan or "seq_lens" not in plan:
        return 1.0, 1.0, 1.0
    lens = plan["seq_lens"]
    pre = plan.get("prefix_rows") or [0] * n_pairs
    rows, attn = 0, 0.0
    for i in range(n_pairs):
        p_, c, r = pre[i], lens[i] - pre[i], lens[n_pairs + i] - pre[i]
        rows += p_ + c + r
        attn += p_ * p_ / 2 + c * c / 2 + c * p_ + r * r / 2 + r * p_
    shared_pairs = sum(1 for x in pre if x > 0)
    return rows / (2.0 * n_pairs * S), attn / (2.0 * n_pairs * S * S / 2), 1.0 - 0.5 * shared_pairs / n_pairs


class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms",
                                       "200", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); power.append(float(c[3]))
            except ValueError:
                continue
            for n, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(power))
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


# ----------------------------------------------------------------------------------------------
# CPU arm: the reference's PyTorch-CPU path (oracle port), bounded sample of the same workload
# ----------------------------------------------------------------------------------------------
_CPU_W = {}
_CPU_THREADS = []


def host_threads() -> int:
    """Threads for the CPU arm: the CPUs this process may run on (affinity mask, capped by the cgroup CPU quota), then the
    fastest of {n, n/2, n/4, ...} on a 4096^3 fp32 matmul -- a box that shows 128 logical CPUs ran the sample 3x SLOWER
    with 128 torch threads than this 8-core container (r1: 35.8 s vs 1.2 s for the ViT part)."""
    if _CPU_THREADS:
        return _CPU_THREADS[0]
    import torch
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(float(quota) / float(period))))
    except Exception:
        pass
    cands, c = [], n
    while c >= 4:
        cands.append(c)
        c //= 2
    best, best_t = n, float("inf")
    if len(cands) > 1:
        a = torch.randn(4096, 4096)
        for c in cands:
            torch.set_num_threads(c)
            a @ a
            t0 = time.perf_counter()
            a @ a
            dt = time.perf_counter() - t0
            if dt < best_t * 0.95:   # prefer more threads unless fewer are clearly faster
                best, best_t = c, dt
    _CPU_THREADS.append(best)
    return best


def cpu_sample_setup(threads: int):
    """7B-shape weights for the bounded CPU sample (ViT, projector, ONE decoder layer, final norm, lm_head); built once."""
    import torch
    from oracle import restate as R
    torch.set_num_threads(threads)
    if "w" not in _CPU_W:
        cfg = R.LLAVA15_7B
        one = R.LlavaCfg(**{**cfg.__dict__, "layers": 1})
        _CPU_W["one"] = one
        _CPU_W["w"] = R.make_weights(one, 0, [n for n, *_ in R.weight_specs(one)])
    return _CPU_W["one"], _CPU_W["w"]


def cpu_sample_seconds_per_pair(threads: int):
    """Bounded sample of the config-2 workload through the oracle port (the reference's PyTorch-CPU path): ONE sequence of
    1599 merged tokens (half a pair; times are doubled), with the 32 identical decoder layers sampled once:
    per pair = 4 x time(ViT+projector, 1 image)   [the reference runs the tower for chosen/rejected x policy/reference]
             + 2 x { 32 x [layer fwd+bwd (policy) + layer fwd (reference)]
                     + final norm/lm_head/get_batch_logps on the full [1,1599,32064] logits: fwd+bwd (policy) + fwd (reference) }"""
    import torch
    from oracle import restate as R
    one, w = cpu_sample_setup(threads)
    cfg = R.LLAVA15_7B
    S, d = TEXT_LEN - 1 + cfg.n_patches, cfg.hidden
    g = torch.Generator().manual_seed(0)
    t = {}
    with torch.no_grad():
        px = torch.randn(1, 3, cfg.image_size, cfg.image_size, generator=g)
        t0 = time.perf_counter()
        feats = R.clip_vision_features(cfg, w, px)[:, 1:]
        R.projector(cfg, w, feats)
        t["vit_proj_1img"] = time.perf_counter() - t0
    x = torch.randn(1, S, d, generator=g) * 0.02
    mask = torch.ones(1, S, dtype=torch.long)
    pos = torch.arange(S)[None]
    lw = {k: v.clone().requires_grad_(True) for k, v in w.items() if k.startswith("language_model.model.layers.0.")}
    wl = dict(w)
    wl.update(lw)
    xin = x.clone().requires_grad_(True)
    t0 = time.perf_counter()
    h = R.llama_decoder(one, wl, xin, mask, pos, return_hidden=True)  # 1 layer + final norm
    h.sum().backward()
    t["layer_fwd_bwd_1seq"] = time.perf_counter() - t0
    with torch.no_grad():
        t0 = time.perf_counter()
        R.llama_decoder(one, w, x, mask, pos, return_hidden=True)
        t["layer_fwd_1seq"] = time.perf_counter() - t0
    labels = torch.randint(3, 32000, (1, S), generator=g)
    labels[:, :PROMPT_LEN + cfg.n_patches - 1] = -100
    hw = w["language_model.lm_head.weight"].clone().requires_grad_(True)
    hh = h.detach().clone().requires_grad_(True)
    t0 = time.perf_counter()
    logits = torch.nn.functional.linear(hh, hw).float()
    R.get_batch_logps(logits, labels).sum().backward()
    t["head_fwd_bwd_1seq"] = time.perf_counter() - t0
    with torch.no_grad():
        t0 = time.perf_counter()
        R.get_batch_logps(torch.nn.functional.linear(hh, hw).float(), labels)
        t["head_fwd_1seq"] = time.perf_counter() - t0
    per_pair = 4 * t["vit_proj_1img"] + 2 * (cfg.layers * (t["layer_fwd_bwd_1seq"] + t["layer_fwd_1seq"]) + t["head_fwd_bwd_1seq"]
                                             + t["head_fwd_1seq"])
    return per_pair, t


CPU_SAMPLE_DESC = ("oracle port (torch fp32 CPU), 7B shapes, ONE sequence of 1599 merged tokens (half a pair, doubled): ViT+projector "
                   "on 1 image x4; one decoder layer fwd+bwd and fwd timed, scaled x32; full-vocab lm_head+get_batch_logps fwd+bwd and "
                   "fwd; optimizer and all-reduce not included")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_threads()
    cpu_sample_setup(cores)
    for _ in range(min(args.warmup, 1)):
        cpu_sample_seconds_per_pair(cores)
    ts = []
    for _ in range(max(1, args.steps)):
        s, parts = cpu_sample_seconds_per_pair(cores)
        ts.append(s)
    sec_per_pair = sum(ts) / len(ts)
    v = 1.0 / sec_per_pair
    line = {"metric": METRIC, "value": v, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec_per_pair * PAIRS_PER_GPU * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "CPU host cores only; extrapolated from a bounded sample"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "logical_cpus": os.cpu_count(), "kind": "port",
                             "sample": CPU_SAMPLE_DESC, "parts_s": {k: round(x, 3) for k, x in parts.items()}},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# CUDA arm
# ----------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        # with NCCL_DEBUG=VERSION|WARN (set on the GPU boxes) NCCL prints "NCCL version ..." on STDOUT, next to the one JSON
        # line this script owes its caller: send NCCL's own log to a per-process file instead (NCCL honours NCCL_DEBUG_FILE
        # only above the VERSION level, so a bare VERSION becomes WARN)
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", os.path.join(tempfile.gettempdir(), "vlb200_nccl.%h.%p.log"))
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import vlrlhf_b200  # noqa: F401
    from vlrlhf_b200 import config, engine, host, ops, synthetic
    cfg = {"7b": config.LLAVA15_7B, "small": config.SMALL, "tiny": config.TINY, "next7b": config.LLAVANEXT_MISTRAL_7B,
           "next_small": config.SMALL_NEXT, "qwen7b": config.QWEN_VL_CHAT, "qwen_small": config.SMALL_QWEN, "xc2_7b": config.XC2_VL_7B,
           "xc2_small": config.SMALL_XC2, "7b_lora": config.LLAVA15_7B_LORA, "next7b_lora": config.LLAVANEXT_MISTRAL_7B_LORA,
           "small_lora": config.SMALL_LORA, "next_small_lora": config.SMALL_NEXT_LORA}[args.model]
    text_len, prompt_len = (TEXT_LEN, PROMPT_LEN) if args.model in ("7b", "next7b") else (96, 24)
    is_next = cfg.family == "llava_next"
    if cfg.family == "llava_next" or args.pack:   # (LLaVA-Next: variable packed feature lengths, not built)
        args.share_prefix = False
    if cfg.family in ("qwen_vl", "xc2"):
        return run_b200_qwen(args, cfg, world, rank, local)
    if getattr(cfg, "lora_r", 0):
        return run_b200_lora(args, cfg, world, rank, local)
    loss_type = "ddpo" if (is_next and args.loss_type == "sigmoid") else args.loss_type  # configs[3] is DDPO
    # configs[3] (LLaVA-Next-Mistral-7B, S = 2199) keeps only the layer inputs for backward so that full-FT fits one GPU
    # built through the reference-facing wrapper (plugin.B200LlavaForRL: nn.Parameters that are views of the engine's arenas)
    from vlrlhf_b200 import plugin
    model = plugin.B200LlavaForRL(cfg, config.TrainConfig(loss_type=loss_type, activation_checkpointing=(args.model == "next7b"),
                                                          pack_sequences=args.pack, share_prefix=args.share_prefix))
    eng = model.engine
    eng.init_synthetic(0)  # same weights on every rank
    batch = synthetic.make_batch(cfg, PAIRS_PER_GPU, text_len, prompt_len, seed=1000 + rank, pin=True)  # rank-local pairs
    cb = host.concatenated_inputs(batch)
    ids_h, am_h, lb_h = cb["concatenated_input_ids"], cb["concatenated_attention_mask"], cb["concatenated_labels"]
    sizes_h = batch["img_input_dict"].get("image_sizes")
    wt_h = eng.ddpo_weights(ids_h, am_h, lb_h, sizes_h) if loss_type == "ddpo" else None
    dev_inputs = eng.prepare_inputs(ids_h, am_h, lb_h, batch["img_input_dict"]["pixel_values"], wt_h, sizes_h)
    S = dev_inputs[5].merged_len if is_next else text_len - 1 + cfg.n_patches
    if is_next:
        cfg._bench_crops_per_image = dev_inputs[5].crops[0]
    rows_lm = 2 * PAIRS_PER_GPU * (text_len - 1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps

    # --pack (side measurement, SURVEY f-2): the padding rows of the ragged synthetic batch are dropped from every kernel
    # --share-prefix: one copy of every pair's prompt + image prefix; same log-probs / losses / gradients
    plan = eng.host_row_plan(ids_h, am_h, sizes_h)
    seq_lens = plan.get("seq_lens")
    step_dev = lambda: eng.step(*dev_inputs, train=True, **plan)  # noqa: E731
    last = {}

    def step_e2e():
        last.update(eng.train_step(batch, train=True))

    for _ in range(max(3, args.warmup)):
        step_dev()
    sampler = ClockSampler(local) if rank == 0 else None
    n0 = ops.launch_count()
    nvtx = os.environ.get("VLB_NVTX") == "1"  # `ncu --nvtx --nvtx-include "vlb_step/"` profiles only the timed steps
    if nvtx:
        torch.cuda.nvtx.range_push("vlb_step")
    ms_dev = timed(step_dev, args.steps)
    if nvtx:
        torch.cuda.nvtx.range_pop()
    launches = (ops.launch_count() - n0)
    ms_e2e = timed(step_e2e, args.steps) if not args.skip_e2e else None
    clocks = sampler.stop() if sampler else None
    padded_layout = None
    if eng.tc.share_prefix and not args.skip_e2e:
        # the same step in the reference's padded [2B, S] layout (every prefix computed twice), on the same box, for reference
        eng.tc.share_prefix = False
        for _ in range(2):
            eng.step(*dev_inputs, train=True)
        ms_pad = timed(lambda: eng.step(*dev_inputs, train=True), min(args.steps, 5))
        eng.tc.share_prefix = True
        eng.step(*dev_inputs, train=True, **plan)
        padded_layout = {"value": PAIRS_PER_GPU * world / (ms_pad / 1e3), "unit": UNIT, "ms_per_step": ms_pad,
                         "rows_per_step": 2 * PAIRS_PER_GPU * S, "steps": min(args.steps, 5)}

    # dominant kernel live: the tcgen05 GEMM at its largest forward shape (gate_up: [T,d] x [2ff,d]^T)
    T = 2 * PAIRS_PER_GPU * S
    a = eng.buf("s.h", (T, cfg.hidden))
    wgu = eng.policy["L0.wgu"]
    out = eng.buf("s.gu" if eng.tc.activation_checkpointing else "a.gu.0", (T, 2 * cfg.ff))
    gemm_ms = timed(lambda: ops.gemm(a, wgu, out=out), 10)
    pk, pk_src = peaks()
    gemm_tf = 2.0 * T * 2 * cfg.ff * cfg.hidden / gemm_ms / 1e9
    traffic, traffic_detail = _traffic()
    flops = step_flops(cfg, PAIRS_PER_GPU, S, rows_lm)              # what the reference's padded [2B, S] formulation computes
    flops_exec = executed_flops(cfg, PAIRS_PER_GPU, S, rows_lm, plan)   # what this step executes (shared / packed rows)
    # second kernel family of the north_star ("fraction of the attention/GEMM roofline"): the fused causal attention of one
    # decoder layer, forward and backward, on random q|k|v at the step's shape; algorithmic FLOPs = causal half of QK^T + PV
    nseq, H, KV, dh = 2 * PAIRS_PER_GPU, cfg.heads, cfg.kv_heads, cfg.head_dim
    hd, kvd = H * dh, KV * dh
    qkv = (torch.randn(T, cfg.qkv_dim, device="cuda") * 0.5).to(torch.bfloat16)
    att, datt, dqkv = torch.empty(T, hd, dtype=torch.bfloat16, device="cuda"), (torch.randn(T, hd, device="cuda") * 0.1).to(torch.bfloat16), torch.empty_like(qkv)
    lse, delta = (torch.empty(nseq, H, S, dtype=torch.float32, device="cuda") for _ in range(2))
    full = torch.full((nseq,), S, dtype=torch.int32, device="cuda")
    sc = 1.0 / dh ** 0.5
    q_, k_, v_ = qkv[:, :hd], qkv[:, hd:hd + kvd], qkv[:, hd + kvd:]
    # kernels timed ALONE (burst peak as the denominator): untimed launches first -- the attention kernels are bound by the
    # softmax / elementwise warps' instruction issue, so their time follows the SM clock, which needs a few ms without the
    # GEMMs' power draw to leave the step's power-capped level (0.34 vs 0.26 ms measured for the same forward kernel)
    attn_f = lambda: ops.attn_fwd_tc(q_, k_, v_, att, lse, full, nseq, S, H, KV, dh, True, sc)  # noqa: E731
    attn_b = lambda: ops.attn_bwd_tc(q_, k_, v_, att, datt, lse, delta, dqkv[:, :hd], dqkv[:, hd:hd + kvd], dqkv[:, hd + kvd:],  # noqa: E731
                                     full, nseq, S, H, KV, dh, True, sc)
    def timed_alone(fn, untimed, n):
        """-> (ms per launch, median SM clock during the loop): `untimed` launches first (~0.1 s: the clock leaves the step's
        power-capped level), then n timed ones with the nvidia-smi sampler running (200 ms period)."""
        for _ in range(untimed):
            fn()
        smp = ClockSampler(local) if rank == 0 else None
        ms = timed(fn, n)
        clk = smp.stop() if smp else {}
        return ms, clk.get("sm_mhz")
    af_ms, af_mhz = timed_alone(attn_f, 300, 1500)
    ab_ms, ab_mhz = timed_alone(attn_b, 120, 600)
    attn_fl = 4.0 * S * S * dh * H * nseq / 2
    del qkv, att, datt, dqkv
    # the two long-K GEMM shapes that carry as much of the step as the forward shape: the gate|up input
    # gradient (K = 2*ff) and weight gradient (K = T, MN-major operands), timed alone like the headline kernel
    dgu = eng.buf("s.gu" if eng.tc.activation_checkpointing else "a.gu.0", (T, 2 * cfg.ff))
    dh2 = eng.buf("b.dxf", (T, cfg.hidden))
    dgrad_ms = timed(lambda: ops.gemm(dgu, wgu, b_kmajor=False, out=dh2), 10)
    gw = eng.g["L0.wgu"]
    wgrad_ms = timed(lambda: ops.gemm(dgu, a, a_kmajor=False, b_kmajor=False, out=gw), 10)
    gemm_fl = 2.0 * T * 2 * cfg.ff * cfg.hidden

    # the same step through the Trainer-side boundary (plugin.concatenated_forward x2 -> dpo_loss -> loss.backward() ->
    # B200FlatAdamW.step -> zero_grad), i.e. trl's get_batch_loss_metrics + HF training_step, host batch in, metrics out
    ms_plugin, plugin_last = None, {}
    trainer = opt = step_plugin = None
    if not args.skip_plugin and not args.skip_e2e:
        from types import SimpleNamespace
        eng.wait_optimizer()
        trainer = SimpleNamespace(loss_type=loss_type, beta=eng.tc.beta, label_smoothing=eng.tc.label_smoothing,
                                  reference_free=False, label_pad_token_id=-100, padding_value=0, is_encoder_decoder=False,
                                  ref_model=plugin.RefView(model), precompute_ref_log_probs=False)
        opt = model.flat_optimizer(lr=eng.tc.learning_rate, betas=(eng.tc.adam_beta1, eng.tc.adam_beta2), eps=eng.tc.adam_eps,
                                   weight_decay=eng.tc.weight_decay, max_grad_norm=eng.tc.max_grad_norm)

        def step_plugin():
            pc, pr, pcl, prl = plugin.concatenated_forward(trainer, model, batch)
            with torch.no_grad():
                rc, rr, _, _ = plugin.concatenated_forward(trainer, trainer.ref_model, batch)
            losses, cr, rj = plugin.dpo_loss(trainer, pc, pr, rc, rr)
            loss = losses.mean()
            loss.backward()
            opt.step()
            model.zero_grad()
            packed = torch.stack([loss.detach(), cr.mean(), rj.mean(), (cr > rj).float().mean(), pc.detach().mean(),
                                  pr.detach().mean(), pcl.detach().mean(), prl.detach().mean()]).cpu()   # the D2H read
            plugin_last.update(loss=float(packed[0]), **{"rewards/chosen": float(packed[1]), "rewards/rejected": float(packed[2]),
                               "rewards/accuracies": float(packed[3]), "logps/chosen": float(packed[4]),
                               "logps/rejected": float(packed[5]), "logits/chosen": float(packed[6]),
                               "logits/rejected": float(packed[7])})
        step_plugin()
        ms_plugin = timed(step_plugin, args.steps)
    if rank == 0:
        pairs = PAIRS_PER_GPU * world
        h2d = sum(int(t.numel() * t.element_size()) for t in (cb["concatenated_input_ids"], cb["concatenated_attention_mask"],
                                                             cb["concatenated_labels"], batch["img_input_dict"]["pixel_values"]))
        line = {
            "metric": METRIC, "value": pairs / (ms_dev / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD if args.model == "7b" else (
                           WORKLOAD_NEXT if args.model == "next7b" else f"{args.model} (dev config, NOT the benchmark)"),
                       "pairs_per_gpu": PAIRS_PER_GPU, "text_len": text_len, "merged_len": S, "loss_type": loss_type,
                       "activation_checkpointing": eng.tc.activation_checkpointing, "pack_sequences": eng.tc.pack_sequences,
                       "share_prefix": eng.tc.share_prefix,
                       "rows_per_step": (sum(seq_lens) - sum(plan.get("prefix_rows", [])) if seq_lens else 2 * PAIRS_PER_GPU * S),
                       "shared_prefix_rows_per_step": sum(plan.get("prefix_rows", [])),
                       "parallelism": (f"dp{world}: gradients reduce-scattered (NCCL), AdamW on each rank's 1/{world} slice (ZeRO-1), "
                                       f"parameters all-gathered; deferred to a side stream under the next reference pass")
                       if world > 1 else "dp1 (no collective)",
                       "optimizer": "AdamW fp32 master+moments, max_grad_norm 1.0",
                       "l2": "inputs>>L2 (each step streams >100 GB through HBM)",
                       "step_tflop_algorithmic": flops / 1e12, "step_tflop_executed": flops_exec / 1e12,
                       "step_tensor_util_of_sustained_peak": flops_exec / (ms_dev / 1e3) / 1e12 / pk["bf16_tflops_sustained"],
                       "step_tensor_util_note": "EXECUTED FLOPs / time / sustained cuBLAS peak (shared-prefix rows execute fewer FLOPs "
                                                "than the padded formulation in step_tflop_algorithmic)"},
            "e2e": {"value": (pairs / (ms_e2e / 1e3) if ms_e2e else None), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 11 * 4,
                    "ms_per_step": ms_e2e, "last_metrics": last},
            "padded_layout": padded_layout,
            "e2e_plugin": {"value": (pairs / (ms_plugin / 1e3) if ms_plugin else None), "unit": UNIT, "ms_per_step": ms_plugin,
                           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8 * 4, "last_metrics": plugin_last,
                           "path": "plugin.concatenated_forward(policy) + (RefView) -> plugin.dpo_loss -> losses.mean().backward() "
                                   "-> B200FlatAdamW.step() -> model.zero_grad(): trl get_batch_loss_metrics + HF training_step"},
            "gpu_launches": launches,
            "hbm_gb": {"peak_allocated": torch.cuda.max_memory_allocated() / 1e9, "peak_reserved": torch.cuda.max_memory_reserved() / 1e9,
                       "device_total": torch.cuda.get_device_properties(local).total_memory / 1e9},
            "roofline": {"bound": "tensor", "kernel": "gemm_bf16_2cta_kernel<6,K-major,K-major> (tcgen05 cta_group::2, 256x256 pair tiles, gate_up fwd shape)",
                         "achieved": gemm_tf, "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": gemm_tf / pk["bf16_tflops"],
                         "peak_source": f"{pk_src} (burst; kernel timed alone)", "traffic": traffic, "traffic_detail": traffic_detail},
            "roofline_gemm_longk": [
                {"bound": "tensor", "kernel": "gemm_bf16_2cta_kernel<6,K-major,MN-major> gate|up input gradient dh = dgu Wgu (M=T, N=d, K=2*ff)",
                 "achieved": gemm_fl / dgrad_ms / 1e9, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                 "frac": gemm_fl / dgrad_ms / 1e9 / pk["bf16_tflops"], "ms": dgrad_ms},
                {"bound": "tensor", "kernel": "gemm_bf16_2cta_kernel<6,MN-major,MN-major> gate|up weight gradient dW = dgu^T h (M=2*ff, N=d, K=T)",
                 "achieved": gemm_fl / wgrad_ms / 1e9, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                 "frac": gemm_fl / wgrad_ms / 1e9 / pk["bf16_tflops"], "ms": wgrad_ms}],
            "roofline_attention": [
                {"bound": "tensor", "kernel": f"attn_fwd_tc_kernel<{dh}> (tcgen05 causal FlashAttention forward, one decoder layer)",
                 "achieved": attn_fl / af_ms / 1e9, "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": attn_fl / af_ms / 1e9 / pk["bf16_tflops"],
                 "ms": af_ms, "sm_mhz": af_mhz},
                {"bound": "tensor", "kernel": f"attn_delta_kernel + attn_bwd_tc_kernel<{dh},dKdV> + <{dh},dQ> (backward of the same layer)",
                 "achieved": 2.5 * attn_fl / ab_ms / 1e9, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                 "frac": 2.5 * attn_fl / ab_ms / 1e9 / pk["bf16_tflops"], "ms": ab_ms, "sm_mhz": ab_mhz}],
            "clocks": clocks,
        }
        if world == 1 and args.model == "7b" and not args.no_library_baseline:
            # the stock library stack on the same GPU (bench_library.py): the engine's arenas are released first
            try:
                model = eng = dev_inputs = a = wgu = out = dgu = dh2 = gw = q_ = k_ = v_ = lse = delta = None   # noqa: F841
                trainer = opt = step_dev = step_e2e = step_plugin = None                                    # noqa: F841
                import gc
                gc.collect()
                torch.cuda.empty_cache()
                import bench_library
                from types import SimpleNamespace
                lib = bench_library.run_library(SimpleNamespace(steps=min(args.steps, 3), warmup=2))
                line["gpu_library_baseline"] = {"value": lib["value"], "unit": UNIT, "ms_per_step": lib["ms_per_step"],
                                                "e2e_value": lib["e2e"]["value"], "stack": lib["config"]["stack"],
                                                "versions": {k: lib["config"].get(k) for k in ("transformers", "torch")}}
            except Exception as e:  # a baseline that cannot run must not take the measurement down with it
                line["gpu_library_baseline"] = {"unavailable": f"{type(e).__name__}: {str(e)[:200]}"}
        if world == 1 and not args.no_cpu_baseline:
            cores = host_threads()
            sec, parts = cpu_sample_seconds_per_pair(cores)
            line["cpu_baseline"] = {"value": 1.0 / sec, "unit": UNIT, "cores": cores, "logical_cpus": os.cpu_count(), "kind": "port",
                                    "sample": CPU_SAMPLE_DESC,
                                    "parts_s": {k: round(x, 3) for k, x in parts.items()}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_b200_qwen(args, cfg, world, rank, local):
    """Side measurement for BASELINE.json configs[2]: Qwen-VL-Chat DPO, LoRA r=64 on the LM, frozen tower, 4 pairs/GPU,
    text 1024 (the 256 image tokens sit inside the text), one 448-px image per pair.  Same timing rules as run_b200."""
    import torch
    import torch.distributed as dist
    import vlrlhf_b200  # noqa: F401
    from vlrlhf_b200 import config, engine_qwen, engine_xc2, host, ops, synthetic
    full = args.model in ("qwen7b", "xc2_7b")
    is_xc2 = cfg.family == "xc2"
    if is_xc2:  # configs[4]: KTO-pair, text 1024 + 1225 image tokens (490 px) = 2248 merged
        text_len, prompt_len = (1024, 32) if full else (96, 24)
        loss_type = "kto_pair" if args.loss_type == "sigmoid" else args.loss_type
        eng = engine_xc2.XC2DPOEngine(cfg, config.TrainConfig(loss_type=loss_type, learning_rate=1e-5, weight_decay=0.1,
                                                              pack_sequences=args.pack, share_prefix=args.share_prefix))
        eng.init_synthetic(0)
        batch = synthetic.make_batch(cfg, PAIRS_PER_GPU, text_len, prompt_len, seed=1000 + rank, pin=True)
        args.loss_type = loss_type
    else:
        text_len, prompt_len = (1024, 320) if full else (128, 72)
        eng = engine_qwen.QwenVLDPOEngine(cfg, config.TrainConfig(loss_type=args.loss_type, learning_rate=1e-5, weight_decay=0.05,
                                                                  pack_sequences=args.pack, share_prefix=args.share_prefix))
        eng.init_synthetic(0)
        batch = synthetic.make_qwen_batch(cfg, PAIRS_PER_GPU, text_len, prompt_len, seed=1000 + rank, pin=True)
    cb = host.concatenated_inputs(batch)
    ids_h, am_h, lb_h = cb["concatenated_input_ids"], cb["concatenated_attention_mask"], cb["concatenated_labels"]
    wt_h = eng.ddpo_weights(ids_h, am_h, lb_h) if args.loss_type == "ddpo" else None
    dev_inputs = eng.prepare_inputs(ids_h, am_h, lb_h, batch["img_input_dict"]["pixel_values"], wt_h)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps

    last = {}
    plan = eng.host_row_plan(ids_h, am_h)   # {} for the padded layout; seq_lens (+ prefix_rows) for packed / shared-prefix rows
    step_dev = lambda: eng.step(*dev_inputs, train=True, **plan)  # noqa: E731
    step_e2e = lambda: last.update(eng.train_step(batch, train=True))  # noqa: E731
    for _ in range(max(3, args.warmup)):
        step_dev()
    sampler = ClockSampler(local) if rank == 0 else None
    n0 = ops.launch_count()
    ms_dev = timed(step_dev, args.steps)
    launches = ops.launch_count() - n0
    ms_e2e = timed(step_e2e, args.steps) if not args.skip_e2e else None
    clocks = sampler.stop() if sampler else None
    if is_xc2:
        return _finish_xc2(args, cfg, eng, world, rank, ms_dev, ms_e2e, launches, clocks, last, text_len, batch, ids_h, am_h, lb_h)
    # algorithmic FLOPs: LM forward x3 (policy fwd, reference fwd, dgrad-only backward) per sequence, adapters, tower once/pair
    d, ff, L, S, r = cfg.hidden, cfg.ff, cfg.layers, text_len, cfg.lora_r
    p_layer = 3 * d * d + d * d + 3 * d * ff
    lin_seq, attn_seq = S * 2 * p_layer * L, L * 2 * S * S * d
    per_seq = 3 * lin_seq + 4 * attn_seq  # linear: policy fwd + reference fwd + dgrad; attention: 2 fwd + a 2x backward
    lora_seq = S * 2 * L * r * (d + 3 * d + d + d + 2 * (d + ff))
    rows_lm = 2 * PAIRS_PER_GPU * (text_len - 1)
    w, P = cfg.v_width, cfg.n_patches
    vit = cfg.v_layers * (P * 2 * (4 * w * w + 2 * w * cfg.v_mlp) + 4 * P * P * w) + P * 2 * cfg.patch_k * w
    resampler = P * 2 * (w * d + 2 * d * d) + 4 * cfg.n_queries * P * d + cfg.n_queries * 2 * 2 * d * d
    flops = PAIRS_PER_GPU * (2 * (per_seq + lora_seq * 3) + vit + resampler) + rows_lm * 2 * d * cfg.vocab * 3
    rr, ra, _ = plan_ratios(plan, PAIRS_PER_GPU, S)   # shared-prefix / packed rows execute fewer FLOPs than the padded formulation
    flops_exec = PAIRS_PER_GPU * (2 * (3 * lin_seq * rr + 4 * attn_seq * ra + lora_seq * 3 * rr) + vit + resampler) + rows_lm * 2 * d * cfg.vocab * 3
    pk, pk_src = peaks()
    if rank == 0:
        pairs = PAIRS_PER_GPU * world
        h2d = sum(int(t.numel() * t.element_size()) for t in (ids_h, am_h, lb_h, batch["img_input_dict"]["pixel_values"]))
        line = {"metric": METRIC, "value": pairs / (ms_dev / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(3, args.warmup), "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": ("Qwen-VL-Chat DPO bf16 (configs[2], side measurement), LoRA r=64 alpha=16 on c_attn/attn.c_proj/"
                                        "w1/w2, frozen ViT-bigG+resampler, 4 pairs/GPU, text 1024 incl. 256 image tokens, 1x448px "
                                        "image/pair") if full else f"{args.model} (dev config, NOT the benchmark)",
                           "pairs_per_gpu": PAIRS_PER_GPU, "text_len": text_len, "merged_len": text_len, "loss_type": args.loss_type,
                           "pack_sequences": eng.tc.pack_sequences, "share_prefix": eng.tc.share_prefix,
                           "parallelism": f"dp{world}", "optimizer": "AdamW on the adapters only (fp32 master+moments)",
                           "step_tflop_algorithmic": flops / 1e12, "step_tflop_executed": flops_exec / 1e12,
                           "step_tensor_util_of_sustained_peak": flops_exec / (ms_dev / 1e3) / 1e12 / pk["bf16_tflops_sustained"],
                           "step_tensor_util_note": "EXECUTED FLOPs / time / sustained cuBLAS peak (shared-prefix rows execute fewer FLOPs "
                                                    "than the padded formulation, step_tflop_algorithmic)"},
                "e2e": {"value": (pairs / (ms_e2e / 1e3) if ms_e2e else None), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 11 * 4,
                        "ms_per_step": ms_e2e, "last_metrics": last},
                "gpu_launches": launches, "clocks": clocks}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_b200_lora(args, cfg, world, rank, local):
    """Side measurement: LLaVA-1.5-7B / LLaVA-Next-Mistral-7B trained the way the reference's launch scripts do it
    (scripts/dpo_llava.sh, dpo_llavanext.sh: LoRA r=128 alpha=256 on the seven decoder linears, frozen tower/projector,
    reference pass = adapters off), 4 pairs/GPU, text 1024, one 336-px image per pair.  Same timing rules as run_b200."""
    import torch
    import torch.distributed as dist
    import vlrlhf_b200  # noqa: F401
    from vlrlhf_b200 import config, engine_lora, host, ops, synthetic
    full = args.model in ("7b_lora", "next7b_lora")
    is_next = cfg.family == "llava_next"
    text_len, prompt_len = (TEXT_LEN, PROMPT_LEN) if full else (96, 24)
    loss_type = "ddpo" if (is_next and args.loss_type == "sigmoid") else args.loss_type  # configs[3] is DDPO
    eng = engine_lora.LlavaLoRADPOEngine(cfg, config.TrainConfig(loss_type=loss_type, learning_rate=1e-5,
                                                                 activation_checkpointing=args.checkpointing,
                                                                 pack_sequences=args.pack, share_prefix=args.share_prefix))
    eng.init_synthetic(0)
    batch = synthetic.make_batch(cfg, PAIRS_PER_GPU, text_len, prompt_len, seed=1000 + rank, pin=True)
    cb = host.concatenated_inputs(batch)
    ids_h, am_h, lb_h = cb["concatenated_input_ids"], cb["concatenated_attention_mask"], cb["concatenated_labels"]
    sizes_h = batch["img_input_dict"].get("image_sizes")
    wt_h = eng.ddpo_weights(ids_h, am_h, lb_h, sizes_h) if loss_type == "ddpo" else None
    dev_inputs = eng.prepare_inputs(ids_h, am_h, lb_h, batch["img_input_dict"]["pixel_values"], wt_h, sizes_h)
    S = dev_inputs[5].merged_len if is_next else text_len - 1 + cfg.n_patches
    crops = dev_inputs[5].crops[0] if is_next else 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps

    last = {}
    plan = eng.host_row_plan(ids_h, am_h, sizes_h)
    step_dev = lambda: eng.step(*dev_inputs, train=True, **plan)  # noqa: E731
    step_e2e = lambda: last.update(eng.train_step(batch, train=True))  # noqa: E731
    for _ in range(max(3, args.warmup)):
        step_dev()
    sampler = ClockSampler(local) if rank == 0 else None
    n0 = ops.launch_count()
    ms_dev = timed(step_dev, args.steps)
    launches = ops.launch_count() - n0
    ms_e2e = timed(step_e2e, args.steps) if not args.skip_e2e else None
    clocks = sampler.stop() if sampler else None
    # algorithmic FLOPs: base linears x3 (policy fwd, reference fwd, dgrad-only backward), attention 2 fwd + a 2x backward,
    # adapters x3 (fwd, dA/dB, dt/dx), tower once per crop, frozen projector once per pass
    d, ff, L, r = cfg.hidden, cfg.ff, cfg.layers, cfg.lora_r
    hd, kvd = cfg.heads * cfg.head_dim, cfg.kv_heads * cfg.head_dim
    p_layer = cfg.qkv_dim * d + d * hd + 3 * d * ff
    lin_seq, attn_seq = S * 2 * p_layer * L, L * 2 * S * S * d
    lora_seq = S * 2 * L * r * ((3 * d + cfg.qkv_dim) + (hd + d) + 2 * (d + ff) + (ff + d))
    per_seq = 3 * lin_seq + 4 * attn_seq + 3 * lora_seq
    rows_lm = 2 * PAIRS_PER_GPU * (text_len - 1)
    dv, Sv, P = cfg.v_hidden, cfg.n_patches + 1, cfg.n_patches
    vit = cfg.v_used_layers * (Sv * 2 * (4 * dv * dv + 2 * dv * cfg.v_ff) + 4 * Sv * Sv * dv) + P * 2 * cfg.patch_k * dv
    proj = P * 2 * (dv * d + d * d) * 2
    flops = PAIRS_PER_GPU * (2 * per_seq + crops * (vit + proj)) + rows_lm * 2 * d * cfg.vocab * 3
    # shared-prefix / packed rows execute fewer FLOPs than the padded formulation above: linear work scales with the rows, the
    # attention of a sequence with n own and c context rows with n*n/2 + n*c
    flops_exec = flops
    if "seq_lens" in plan:
        lens, pre = plan["seq_lens"], plan.get("prefix_rows") or [0] * PAIRS_PER_GPU
        rows, att = 0, 0.0
        for i in range(PAIRS_PER_GPU):
            p_, c_, r_ = pre[i], lens[i] - pre[i], lens[PAIRS_PER_GPU + i] - pre[i]
            rows += p_ + c_ + r_
            att += 4.0 * d * L * (p_ * p_ / 2 + c_ * c_ / 2 + c_ * p_ + r_ * r_ / 2 + r_ * p_)
        flops_exec = rows * 2 * p_layer * L * 3 + 4 * att + 3 * rows * 2 * L * r * ((3 * d + cfg.qkv_dim) + (hd + d) + 2 * (d + ff) + (ff + d)) \
            + PAIRS_PER_GPU * crops * (vit + proj) + rows_lm * 2 * d * cfg.vocab * 3
    pk, pk_src = peaks()
    if rank == 0:
        pairs = PAIRS_PER_GPU * world
        h2d = sum(int(t.numel() * t.element_size()) for t in (ids_h, am_h, lb_h, batch["img_input_dict"]["pixel_values"]))
        name = "LLaVA-Next-Mistral-7B DDPO" if is_next else "LLaVA-1.5-7B DPO"
        line = {"metric": METRIC, "value": pairs / (ms_dev / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(3, args.warmup), "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": (f"{name} bf16 with LoRA r={r} alpha={cfg.lora_alpha:g} on q/k/v/o/gate/up/down_proj (the reference "
                                        f"scripts' setting; side measurement), frozen CLIP-L/336 + projector, 4 pairs/GPU, text 1024 "
                                        f"({S} merged), 1x336px image/pair") if full else f"{args.model} (dev config, NOT the benchmark)",
                           "pairs_per_gpu": PAIRS_PER_GPU, "text_len": text_len, "merged_len": S, "loss_type": loss_type,
                           "activation_checkpointing": eng.tc.activation_checkpointing, "pack_sequences": eng.tc.pack_sequences,
                           "share_prefix": eng.tc.share_prefix, "shared_prefix_rows_per_step": sum(plan.get("prefix_rows", [])),
                           "parallelism": f"dp{world}", "optimizer": "AdamW on the adapters only (fp32 master+moments)",
                           "step_tflop_algorithmic": flops / 1e12, "step_tflop_executed": flops_exec / 1e12,
                           "step_tensor_util_of_sustained_peak": flops_exec / (ms_dev / 1e3) / 1e12 / pk["bf16_tflops_sustained"]},
                "e2e": {"value": (pairs / (ms_e2e / 1e3) if ms_e2e else None), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 11 * 4,
                        "ms_per_step": ms_e2e, "last_metrics": last},
                "gpu_launches": launches, "clocks": clocks}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _finish_xc2(args, cfg, eng, world, rank, ms_dev, ms_e2e, launches, clocks, last, text_len, batch, ids_h, am_h, lb_h):
    import torch.distributed as dist
    d, ff, L, r, pr, P = cfg.hidden, cfg.ff, cfg.layers, cfg.lora_r, cfg.plora_r, cfg.n_patches
    S = text_len - 1 + P
    hd, kvd = cfg.heads * cfg.head_dim, cfg.kv_heads * cfg.head_dim
    p_layer = cfg.qkv_dim * d + d * hd + 3 * d * ff
    lin_seq, attn_seq = S * 2 * p_layer * L, L * 2 * S * S * d
    lora_seq = S * 2 * L * r * ((d + cfg.qkv_dim) + (hd + d) + 2 * (d + ff) + (ff + d))
    plora_seq = P * 2 * L * pr * ((d + cfg.qkv_dim) + (hd + d) + 2 * (d + ff) + (ff + d))   # image rows only, frozen
    per_seq = 3 * (lin_seq + plora_seq) + 4 * attn_seq + 3 * lora_seq
    rows_lm = 2 * PAIRS_PER_GPU * (text_len - 1)
    dv, Sv = cfg.v_hidden, P + 1
    vit = cfg.v_used_layers * (Sv * 2 * (4 * dv * dv + 2 * dv * cfg.v_ff) + 4 * Sv * Sv * dv) + P * 2 * cfg.patch_k * dv
    proj = P * 2 * (dv * d + d * d) * 2
    flops = PAIRS_PER_GPU * (2 * per_seq + vit + proj) + rows_lm * 2 * d * cfg.vocab * 3
    plan = eng.host_row_plan(ids_h, am_h)
    rr, ra, ri = plan_ratios(plan, PAIRS_PER_GPU, S)   # ri: image rows (partial LoRA) executed / padded
    flops_exec = PAIRS_PER_GPU * (2 * (3 * (lin_seq * rr + plora_seq * ri) + 4 * attn_seq * ra + 3 * lora_seq * rr) + vit + proj) + rows_lm * 2 * d * cfg.vocab * 3
    pk, pk_src = peaks()
    if rank == 0:
        pairs = PAIRS_PER_GPU * world
        h2d = sum(int(t.numel() * t.element_size()) for t in (ids_h, am_h, lb_h, batch["img_input_dict"]["pixel_values"]))
        line = {"metric": METRIC, "value": pairs / (ms_dev / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(3, args.warmup), "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": ("InternLM-XComposer2-VL-7B KTO-pair bf16 (configs[4], side measurement), LoRA r=64 on wqkv/wo/"
                                        "w1/w2/w3 + frozen partial-LoRA r=256 on the image rows, frozen CLIP-L/490 + projector, "
                                        "4 pairs/GPU, text 1024 (2248 merged), 1x490px image/pair") if args.model == "xc2_7b"
                           else f"{args.model} (dev config, NOT the benchmark)",
                           "pairs_per_gpu": PAIRS_PER_GPU, "text_len": text_len, "merged_len": S, "loss_type": args.loss_type,
                           "pack_sequences": eng.tc.pack_sequences, "share_prefix": eng.tc.share_prefix,
                           "parallelism": f"dp{world}", "optimizer": "AdamW on the adapters only (fp32 master+moments)",
                           "step_tflop_algorithmic": flops / 1e12, "step_tflop_executed": flops_exec / 1e12,
                           "step_tensor_util_of_sustained_peak": flops_exec / (ms_dev / 1e3) / 1e12 / pk["bf16_tflops_sustained"],
                           "step_tensor_util_note": "EXECUTED FLOPs / time / sustained cuBLAS peak (shared-prefix rows execute fewer FLOPs "
                                                    "than the padded formulation, step_tflop_algorithmic)"},
                "e2e": {"value": (pairs / (ms_e2e / 1e3) if ms_e2e else None), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 11 * 4,
                        "ms_per_step": ms_e2e, "last_metrics": last},
                "gpu_launches": launches, "clocks": clocks}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "library"])
    ap.add_argument("--model", default="7b", choices=["7b", "small", "tiny", "next7b", "next_small", "qwen7b", "qwen_small", "xc2_7b", "xc2_small",
                                                   "7b_lora", "next7b_lora", "small_lora", "next_small_lora"],
                    help="7b = the benchmark (configs[1]); next7b = configs[3] LLaVA-Next-Mistral-7B DDPO, qwen7b = configs[2] Qwen-VL-Chat LoRA (side measurements)")
    ap.add_argument("--loss-type", dest="loss_type", default="sigmoid")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-library-baseline", action="store_true", help="skip the stock transformers+SDPA+AdamW arm on the same GPU")
    ap.add_argument("--skip-plugin", action="store_true", help="skip the e2e_plugin measurement (the Trainer-side boundary)")
    ap.add_argument("--checkpointing", action="store_true", help="activation checkpointing (the *_lora side measurements)")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs only")
    ap.add_argument("--no-share-prefix", dest="share_prefix", action="store_false",
                    help="LLaVA-1.5 family: keep the reference's padded [2B, S] row layout as the headline.  Default: "
                         "TrainConfig.share_prefix -- the prompt + image prefix the chosen and rejected sequence of a pair have in "
                         "common is laid out and computed ONCE (same log-probs, losses and gradients; parity tests "
                         "tests/test_gpu_share_prefix.py); the padded layout is then measured next to it (`padded_layout`)")
    ap.add_argument("--pack", action="store_true",
                    help="TrainConfig.pack_sequences (side measurement: padding rows dropped; the headline run keeps them, as the reference does)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "library":
        import bench_library
        line = bench_library.run_library(args)
        if line is not None:
            print(json.dumps(line), flush=True)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
