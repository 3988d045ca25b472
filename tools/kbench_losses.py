"""Where the stage-1 loss step's time goes (development aid): warm CUDA-event times of the fused forward and backward
kernels called back to back through ops (no autograd, no torch glue) next to the whole training-step time."""
import json
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mulactseg_b200 import _lib, losses, ops, synth  # noqa: E402

DEV = "cuda:0"


def time_ms(fn, warmup=3, iters=20):
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
    n, c, h, w, nseg = 16, 20, 768, 768, 2048
    res = {}
    xs = [synth.logits(n, c, h, w, "cosine", seed=1 + i, device=DEV, coherent=4) for i in range(3)]
    spx = synth.pad_border(synth.superpixel_map(n, h, w, nseg, "jitter", seed=2, device=DEV), nseg, 16)
    trg = synth.multihot_targets(n, nseg, c, seed=3, device=DEV, p_ignore=0.0)
    flags = _lib.MAS_LOSS_CHOICE | _lib.MAS_LOSS_GROUP
    coef = torch.tensor([1e-4, 1e-4, 0.0, 1e-3], device=DEV)
    for rho in (0.02, 0.08, 0.2, 0.5, 1.0):
        mask = synth.region_mask(spx, nseg, rho, seed=4)
        info = ops.multihot_info(trg, c, _lib.MAS_GROUP_ONLYMULTI)
        turn = [0]

        tiles = ops.multihot_tiles(mask)
        use = [None]

        def fwd():
            turn[0] += 1
            return ops.multihot_loss_forward(xs[turn[0] % 3], spx, mask, info, nseg, 0.1, flags, use[0])

        acc, gmax = fwd()

        def bwd():
            turn[0] += 1
            return ops.multihot_loss_backward(xs[turn[0] % 3], spx, mask, info, gmax, coef, nseg, 0.1, flags, use[0])

        t_f0, t_b0 = time_ms(fwd), time_ms(bwd)          # every tile walked (no list)
        use[0] = tiles
        os.environ["MAS_LOSS_DENSE"] = "0"
        t_f, t_b = time_ms(fwd), time_ms(bwd)                       # active-tile list walk
        os.environ["MAS_LOSS_DENSE"] = "1"
        t_fd, t_bd = time_ms(fwd), time_ms(bwd)                     # dense TMA strip walk, forced
        del os.environ["MAS_LOSS_DENSE"]
        t_fa, t_ba = time_ms(fwd), time_ms(bwd)                     # the device picks
        n_groups = (tiles.numel() - 16 - 4) // 8
        active = int(tiles[16 + 4 * n_groups:].view(torch.int32)[n_groups])
        total_tiles = n * ((w + 31) // 32) * ((h + 7) // 8)
        t_scan = time_ms(lambda: ops.multihot_tiles(mask))
        zero = torch.empty_like(xs[0])
        t_z = time_ms(lambda: zero.zero_())
        args = types.SimpleNamespace(nseg=nseg, group_ce_temp=0.1, multi_ce_temp=0.1)
        group, multi = losses.stage1_criterion(args, c - 1)
        xg = [x.clone().requires_grad_(True) for x in xs]

        def step():
            turn[0] += 1
            xin = xg[turn[0] % 3]
            xin.grad = None
            g = group(xin, trg, spx, mask)
            ce, mc = multi(xin, trg, spx, mask)
            (16.0 * ce + 8.0 * mc + g).backward()

        t_s = time_ms(step)
        import time
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(50):
            step()
        host = (time.perf_counter() - t0) / 50 * 1e3        # host time to ENQUEUE a step (no sync inside)
        torch.cuda.synchronize()
        res[f"rho={rho}"] = {"fwd_all_tiles_ms": round(t_f0, 4), "bwd_all_tiles_ms": round(t_b0, 4), "tile_scan_ms": round(t_scan, 4),
                             "fwd_kernels_ms": round(t_f, 4), "bwd_kernel_ms": round(t_b, 4), "fwd_dense_ms": round(t_fd, 4),
                             "bwd_dense_ms": round(t_bd, 4), "fwd_auto_ms": round(t_fa, 4), "bwd_auto_ms": round(t_ba, 4),
                             "active_tile_frac": round(active / total_tiles, 4), "memset_grad_ms": round(t_z, 4),
                             "step_ms": round(t_s, 4), "host_enqueue_ms": round(host, 4)}
        print(f"rho={rho}", res[f"rho={rho}"], flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/kbench_losses.json", "w"), indent=1)


if __name__ == "__main__":
    main()
