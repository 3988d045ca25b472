"""Kernel micro-benchmarks (development aid; the contract benchmark is bench.py)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mulactseg_b200 import acquisition as acq, ops, synth  # noqa: E402


def time_ms(fn, warmup=3, iters=10):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=16)
    ap.add_argument("--channels", type=int, default=20)
    ap.add_argument("--h", type=int, default=1024)
    ap.add_argument("--w", type=int, default=2048)
    ap.add_argument("--nseg", type=int, default=2048)
    args = ap.parse_args()
    dev = "cuda:0"
    n, c, h, w, s = args.images, args.channels, args.h, args.w, args.nseg
    res = {}
    for coherent in (0, 4):
        for dtype in (torch.float32, torch.bfloat16):
            logits = synth.logits(n, c, h, w, "cosine", seed=1, device=dev, coherent=coherent, dtype=dtype)
            for kind in ("jitter", "grid", "random"):
                spx = synth.superpixel_map(n, h, w, s, kind, seed=2, device=dev, dtype=torch.int32)
                for need_prob in (False, True):
                    stats = acq.RegionStats(n, s, c, dev, need_prob)
                    ms = time_ms(lambda: stats.add_batch(0, logits, spx, 0.1), iters=5 if kind == "random" else 10)
                    gb = n * h * w * (c * logits.element_size() + 4) / 1e9
                    key = f"stats coherent={coherent} {str(dtype)[6:]} map={kind} prob={int(need_prob)}"
                    res[key] = {"ms": round(ms, 3), "GBps": round(gb / ms * 1e3, 1)}
                    print(key, res[key], flush=True)
                    del stats
                if dtype == torch.bfloat16:
                    break
            del logits
    # epilogue + top-k over a 372-image shard
    nr = 372
    cls_sum = torch.rand((nr, s, c), device=dev)
    cls_cnt = torch.randint(0, 100, (nr, s, c), device=dev, dtype=torch.int32)
    w_ = torch.rand(c, device=dev)
    res["region_scores ms"] = round(time_ms(lambda: ops.region_scores(cls_sum, cls_cnt, w_)), 4)
    score, _, dom = ops.region_scores(cls_sum, cls_cnt, w_)
    mask = torch.ones((nr, s), dtype=torch.uint8, device=dev)
    rank = torch.arange(nr, dtype=torch.int32, device=dev)
    res["region_keys ms"] = round(time_ms(lambda: ops.region_keys(score, mask, rank)), 4)
    keys = ops.region_keys(score, mask, rank)
    res["topk(100001)+sort ms"] = round(time_ms(lambda: ops.topk_keys(keys, 100001, True)), 4)
    res["topk(100001) ms"] = round(time_ms(lambda: ops.topk_keys(keys, 100001, False)), 4)
    res["minmax ms"] = round(time_ms(lambda: ops.minmax_nonzero(score)), 4)
    print(json.dumps(res, indent=1))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/kbench.json", "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
