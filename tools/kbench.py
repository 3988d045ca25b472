"""Kernel micro-benchmarks (development aid; the contract benchmark is bench.py).

Times the fused scorer per launch for both data paths (TMA ring / LDG registers), batch sizes, logit
statistics and superpixel maps; inputs of one configuration are rotated over > 126 MB so that a launch
never finds its data in L2.  Writes gpurun_out/kbench.json.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mulactseg_b200 import acquisition as acq, ops, synth  # noqa: E402


def time_ms(fn, warmup=3, iters=10, join=None):
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    if join is not None:
        join()          # launches alternate over side streams: order the caller's stream after them
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--channels", type=int, default=19)
    ap.add_argument("--h", type=int, default=1024)
    ap.add_argument("--w", type=int, default=2048)
    ap.add_argument("--nseg", type=int, default=2048)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--profile", action="store_true", help="a handful of launches of the bench configuration (for ncu)")
    args = ap.parse_args()
    dev = "cuda:0"
    c, h, w, s = args.channels, args.h, args.w, args.nseg
    pool = 24                                    # resident images; launches rotate through them
    res = {}
    if args.profile:
        logits = synth.logits(pool, c, h, w, "cosine", seed=1, device=dev)
        spx = synth.superpixel_map(pool, h, w, s, "jitter", seed=2, device=dev, dtype=torch.int32)
        for need_prob in (True, False):
            stats = acq.RegionStats(pool, s, c, dev, need_prob, group_bytes=0)
            for i in range(6):
                j = (i % 6) * 4
                stats.add_batch(j, logits[j:j + 4], spx[j:j + 4], 0.1)
            torch.cuda.synchronize()
        return
    variants = [("tma", {}), ("tma_1lane", {"LANES": "1"}), ("ldg", {"MAS_SCORER_PATH": "ldg"})]
    if os.environ.get("KBENCH_EXTRA"):
        variants = [("tma", {}), ("tma_w4", {"MAS_SCORER_WARPS": "4"}), ("tma_w4_3lanes", {"MAS_SCORER_WARPS": "4", "LANES": "3"}),
                    ("tma_w2", {"MAS_SCORER_WARPS": "2"}), ("tma_w4_s3", {"MAS_SCORER_WARPS": "4", "MAS_SCORER_STAGES": "3"})]
    if not args.quick:
        variants += [("tma_s3_w6", {"MAS_SCORER_STAGES": "3", "MAS_SCORER_WARPS": "6"}),
                     ("tma_s2_w6", {"MAS_SCORER_WARPS": "6"})]
    for dtype in (torch.float32, torch.bfloat16):
        for coherent in (0, 4):
            logits = synth.logits(pool, c, h, w, "cosine", seed=1, device=dev, coherent=coherent, dtype=dtype)
            for kind in ("jitter", "random"):
                if kind == "random" and (coherent or dtype != torch.float32):
                    continue
                spx = synth.superpixel_map(pool, h, w, s, kind, seed=2, device=dev, dtype=torch.int32)
                for batch in (4, 12):
                    for need_prob in (False, True):
                        for name, env in variants:
                            for k in ("MAS_SCORER_PATH", "MAS_SCORER_STAGES", "MAS_SCORER_WARPS"):
                                os.environ.pop(k, None)
                            os.environ.update({k: v for k, v in env.items() if k != "LANES"})
                            stats = acq.RegionStats(pool, s, c, dev, need_prob, lanes=int(env.get("LANES", 2)), group_bytes=0)
                            nb = pool // batch

                            def run(i, stats=stats, batch=batch, nb=nb):
                                j = (i % nb) * batch
                                stats.add_batch(j, logits[j:j + batch], spx[j:j + batch], 0.1)

                            ms = time_ms(run, iters=6 if kind == "random" else 12, join=stats.join)
                            gb = batch * h * w * (c * logits.element_size() + 4) / 1e9
                            key = f"{name} {str(dtype)[6:]} coherent={coherent} map={kind} B={batch} prob={int(need_prob)}"
                            res[key] = {"ms": round(ms, 4), "GBps": round(gb / ms * 1e3, 1)}
                            print(key, res[key], flush=True)
                            del stats
            del logits
    for k in ("MAS_SCORER_PATH", "MAS_SCORER_STAGES", "MAS_SCORER_WARPS"):
        os.environ.pop(k, None)
    # epilogue + top-k over a 372-image shard
    nr = 372
    cls_sum = torch.rand((nr, s, c), device=dev)
    cls_cnt = torch.randint(0, 100, (nr, s, c), device=dev, dtype=torch.int32)
    w_ = torch.rand(c, device=dev)
    res["region_scores ms"] = round(time_ms(lambda i: ops.region_scores(cls_sum, cls_cnt, w_)), 4)
    score, _, dom = ops.region_scores(cls_sum, cls_cnt, w_)
    mask = torch.ones((nr, s), dtype=torch.uint8, device=dev)
    rank = torch.arange(nr, dtype=torch.int32, device=dev)
    res["region_keys ms"] = round(time_ms(lambda i: ops.region_keys(score, mask, rank)), 4)
    keys = ops.region_keys(score, mask, rank)
    res["topk(100001)+sort ms"] = round(time_ms(lambda i: ops.topk_keys(keys, 100001, True)), 4)
    res["topk(100001) ms"] = round(time_ms(lambda i: ops.topk_keys(keys, 100001, False)), 4)
    res["minmax ms"] = round(time_ms(lambda i: ops.minmax_nonzero(score)), 4)
    print(json.dumps(res, indent=1))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/kbench.json", "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
