"""Timings of the stage-1 loss step (BASELINE config 4) and the stage-2 prototype labeller (config 5) on one GPU.

Development / evidence aid (the contract benchmark is bench.py).  CUDA events, inputs larger than L2 or rotated;
prints one JSON object and writes gpurun_out/bench_stage.json.
"""
import argparse
import json
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mulactseg_b200 import labeller, losses, synth  # noqa: E402


def time_ms(fn, warmup=3, iters=10):
    torch.cuda.cudart().cudaProfilerStart()     # ncu --profile-from-start off skips the input generation
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


def peak():
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    return json.load(open(path))["hbm_gbs"] if os.path.exists(path) else 6650.0


def bench_losses(res, profile, only_rho=None, only_exact=None):
    dev = "cuda:0"
    n, c, h, w, nseg = 16, 20, 768, 768, 2048
    x = synth.logits(n, c, h, w, "cosine", seed=1, device=dev, coherent=4)
    spx = synth.pad_border(synth.superpixel_map(n, h, w, nseg, "jitter", seed=2, device=dev), nseg, 16)
    trg = synth.multihot_targets(n, nseg, c, seed=3, device=dev, p_ignore=0.0)
    args = types.SimpleNamespace(nseg=nseg, group_ce_temp=0.1, multi_ce_temp=0.1)
    for rho, exact in ((0.02, False), (0.2, False), (1.0, False), (1.0, True)):
        if (only_rho is not None and rho != only_rho) or (only_exact is not None and exact != bool(only_exact)):
            continue
        losses.EXACT_SOFTMAX = exact
        mask = synth.region_mask(spx, nseg, rho, seed=4)
        group, multi = losses.stage1_criterion(args, c - 1)
        xs = [x.clone().requires_grad_(True) for _ in range(3)]   # a fresh `preds` every step, like net(images)
        turn = [0]

        def fwd():
            turn[0] += 1
            xin = xs[turn[0] % len(xs)]
            xin.grad = None
            g = group(xin, trg, spx, mask)
            ce, mc = multi(xin, trg, spx, mask)
            return 16.0 * ce + 8.0 * mc + g

        def step():
            fwd().backward()

        if os.environ.get("MAS_PYPROF"):      # where does the host time of a step go?
            import cProfile, pstats
            for _ in range(20):
                step()
            torch.cuda.synchronize()
            pr = cProfile.Profile()
            pr.enable()
            for _ in range(200):
                step()
            torch.cuda.synchronize()
            pr.disable()
            pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
        with torch.no_grad():
            ms_f = time_ms(lambda: fwd(), iters=3 if profile else 10)
        ms_fb = time_ms(step, iters=3 if profile else 10)
        frac = float(mask.float().mean())
        P = h * w
        # algorithmic bytes: mask + int64 ids both directions, logits where selected (fwd + bwd), dense grad
        alg = n * P * ((1 + 8) * 2 + frac * c * 4 * 2 + c * 4)
        res[f"losses rho={rho} exact_softmax={int(exact)}"] = {"fwd_ms": round(ms_f, 4), "fwd_bwd_ms": round(ms_fb, 4), "selected_frac": round(frac, 4),
                                    "alg_GB": round(alg / 1e9, 3), "GBps": round(alg / ms_fb / 1e6, 1),
                                    "frac_of_measured_hbm": round(alg / ms_fb / 1e6 / peak(), 3)}
        print(f"losses rho={rho} exact_softmax={int(exact)}", res[f"losses rho={rho} exact_softmax={int(exact)}"], flush=True)


def bench_labeller(res, profile):
    dev = "cuda:0"
    for name, (h, w, nseg, c, rho) in {"cityscapes": (1024, 2048, 2048, 20, 0.08), "voc": (375, 500, 150, 21, 0.3)}.items():
        feats = synth.features(1, 256, h, w, seed=1, device=dev)
        logits = synth.logits(1, c, h, w, "normal", seed=2, device=dev, coherent=4)
        spx = synth.superpixel_map(1, h, w, nseg, "jitter", seed=3, device=dev)
        trg = synth.multihot_targets(1, nseg, c, seed=4, device=dev, p_ignore=0.0)
        mask = synth.region_mask(spx, nseg, rho, seed=5)
        ms = time_ms(lambda: labeller.pseudo_label_generation(None, feats, logits, trg, mask, spx, check=False),
                     iters=3 if profile else 10)
        out = labeller.pseudo_label_generation(None, feats, logits, trg, mask, spx)
        sel = float(mask.float().mean())
        labelled = float((out != 255).float().mean())
        P = h * w
        # feature columns of selected pixels (assign) + of every unselected pixel of a touched superpixel (propagate, >= once)
        chosen = torch.zeros(nseg, dtype=torch.bool, device=dev)
        chosen[spx[0][mask[0]]] = True
        near = torch.nn.functional.max_pool2d(chosen[spx[0]].float()[None, None], 3, 1, 1)[0, 0] > 0
        touched = torch.zeros(nseg, dtype=torch.bool, device=dev)
        touched[spx[0][near]] = True
        touched_frac = float(touched[spx[0]].float().mean())
        alg = P * (1 + 8) * 3 + sel * P * c * 4 + touched_frac * P * 256 * 4 + P
        res[f"labeller {name}"] = {"ms_per_image": round(ms, 4), "selected_frac": round(sel, 4), "touched_frac": round(touched_frac, 4),
                                   "labelled_frac": round(labelled, 4), "alg_GB": round(alg / 1e9, 4),
                                   "GBps": round(alg / ms / 1e6, 1), "frac_of_measured_hbm": round(alg / ms / 1e6 / peak(), 3)}
        print(f"labeller {name}", res[f"labeller {name}"], flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--profile", action="store_true", help="few iterations (for ncu)")
    ap.add_argument("--only", default="", choices=["", "losses", "labeller"])
    ap.add_argument("--rho", type=float, default=None, help="losses: only this labelled fraction")
    ap.add_argument("--exact", type=int, default=None, help="losses: only this softmax mode (0 fast / 1 exact)")
    args = ap.parse_args()
    res = {}
    if args.only in ("", "losses"):
        bench_losses(res, args.profile, args.rho, args.exact)
    if args.only in ("", "labeller"):
        bench_labeller(res, args.profile)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/bench_stage.json", "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
