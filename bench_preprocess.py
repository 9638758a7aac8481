"""Side measurement for SURVEY.md §8 f-1 (input pipeline): CLIP preprocessing of decoded RGB images.

GPU arm : vlrlhf_b200.preprocess.ClipPreprocessor -- pinned uint8 host buffers -> H2D -> two libvlb200 kernels per
          image -> float32 [3,336,336] in HBM (what the engine consumes).  Timed with CUDA events incl. the H2D copy.
CPU arm : the reference collator's code path (models/Llava/__init__.py:435-443): transformers' PIL-backend
          CLIPImageProcessor (Pillow bicubic + numpy) on one host core, as a DataLoader worker runs it.
Prints one JSON line.  Not the headline bench (bench.py); the DPO step needs 4 images per GPU per step.
"""
import argparse
import json
import time

import numpy as np
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=64)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    import vlrlhf_b200  # noqa: F401
    from vlrlhf_b200 import ops, preprocess
    from oracle import image_restate as IR
    rs = np.random.RandomState(0)
    imgs = [rs.randint(0, 256, (args.height, args.width, 3), dtype=np.uint8) for _ in range(args.images)]
    pinned = [torch.from_numpy(im).pin_memory() for im in imgs]
    pre = preprocess.ClipPreprocessor()
    for _ in range(3):
        out = pre(pinned)
    torch.cuda.synchronize()
    n0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        out = pre(pinned)
    e1.record()
    torch.cuda.synchronize()
    gpu_ms = e0.elapsed_time(e1) / args.iters
    launches = (ops.launch_count() - n0) // args.iters
    # device-resident inputs: the kernels alone
    dev = [p.cuda() for p in pinned]
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.iters):
        out = pre(dev)
    e1.record()
    torch.cuda.synchronize()
    dev_ms = e0.elapsed_time(e1) / args.iters
    assert np.array_equal(out[0].cpu().numpy(), IR.clip_preprocess(imgs[0]))  # checked against the oracle
    in_bytes = args.height * args.width * 3
    out_bytes = 3 * 336 * 336 * 4
    line = {"metric": "clip_preprocess_images_per_sec", "unit": "images/s", "higher_is_better": True,
            "config": {"workload": f"{args.images} RGB uint8 images {args.height}x{args.width} -> float32 [3,336,336] "
                                   "(shortest edge 336 bicubic, center crop, rescale, normalize)"},
            "e2e": {"value": args.images / (gpu_ms * 1e-3), "unit": "images/s", "ms_per_batch": gpu_ms,
                    "h2d_bytes_per_image": in_bytes, "d2h_bytes_per_image": 0},
            "value": args.images / (dev_ms * 1e-3), "ms_per_batch_device_resident": dev_ms,
            "gpu_launches_per_batch": int(launches),
            "roofline": {"bound": "hbm", "unit": "GB/s", "achieved": args.images * (in_bytes + out_bytes) / (dev_ms * 1e-3) / 1e9,
                         "note": "2 launches per image; launch-latency bound at this batch size, not HBM bound"}}
    try:
        from PIL import Image
        from transformers.models.clip.image_processing_pil_clip import CLIPImageProcessorPil
        proc = CLIPImageProcessorPil(size={"shortest_edge": 336}, crop_size={"height": 336, "width": 336})
        pil = [Image.fromarray(im) for im in imgs[:16]]
        torch.set_num_threads(1)
        t = time.perf_counter()
        ref = proc(images=pil, return_tensors="pt")["pixel_values"]
        cpu_s = time.perf_counter() - t
        assert torch.equal(ref[0], out[0].cpu())
        line["cpu_baseline"] = {"value": len(pil) / cpu_s, "unit": "images/s", "cores": 1, "kind": "reference",
                                "sample": f"{len(pil)} images through transformers CLIPImageProcessorPil (Pillow + numpy)"}
    except ImportError as e:  # pragma: no cover
        line["cpu_baseline"] = {"unavailable": str(e)}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
