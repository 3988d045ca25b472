"""Development probe: the dense loss kernels (csrc/losses_dense.cu) under their debug switches -- streaming floor,
L2 hint, ring depth, warps per CTA.  Prints ms for forward and backward at rho = 1 and 0.02."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mulactseg_b200 import _lib, ops, synth  # noqa: E402

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
    xs = [synth.logits(n, c, h, w, "cosine", seed=1 + i, device=DEV, coherent=4) for i in range(3)]
    spx = synth.pad_border(synth.superpixel_map(n, h, w, nseg, "jitter", seed=2, device=DEV), nseg, 16)
    trg = synth.multihot_targets(n, nseg, c, seed=3, device=DEV, p_ignore=0.0)
    flags = _lib.MAS_LOSS_CHOICE | _lib.MAS_LOSS_GROUP
    coef = torch.tensor([1e-4, 1e-4, 0.0, 1e-3], device=DEV)
    info = ops.multihot_info(trg, c, _lib.MAS_GROUP_ONLYMULTI)
    os.environ["MAS_LOSS_DENSE"] = "1"
    configs = [{}, {"MAS_LOSS_DENSE_DEBUG": "1"}, {"MAS_LOSS_DENSE_WARPS": "12"}, {"MAS_LOSS_DENSE_WARPS": "8"},
               {"MAS_LOSS_DENSE_BSTAGES": "2"}]
    if sys.argv[1:] == ["--one"]:      # one forward + one backward at rho = 1 (for ncu)
        mask = synth.region_mask(spx, nseg, 1.0, seed=4)
        tiles = ops.multihot_tiles(mask)
        for _ in range(2):
            _, gmax = ops.multihot_loss_forward(xs[0], spx, mask, info, nseg, 0.1, flags, tiles)
            ops.multihot_loss_backward(xs[1], spx, mask, info, gmax, coef, nseg, 0.1, flags, tiles)
        torch.cuda.synchronize()
        return
    configs += [dict(kv.split("=") for kv in arg.split(",")) for arg in sys.argv[1:]]
    for spx_t, tag in ((spx, "i64"),):
        for rho in (1.0, 0.02):
            mask = synth.region_mask(spx, nseg, rho, seed=4)
            tiles = ops.multihot_tiles(mask)
            for cfg in configs:
                for k in ("MAS_LOSS_DENSE_DEBUG", "MAS_LOSS_DENSE_STAGES", "MAS_LOSS_DENSE_WARPS", "MAS_LOSS_DENSE_BSTAGES"):
                    os.environ.pop(k, None)
                os.environ.update(cfg)
                turn = [0]

                def fwd():
                    turn[0] += 1
                    return ops.multihot_loss_forward(xs[turn[0] % 3], spx_t, mask, info, nseg, 0.1, flags, tiles)

                _, gmax = fwd()

                def bwd():
                    turn[0] += 1
                    return ops.multihot_loss_backward(xs[turn[0] % 3], spx_t, mask, info, gmax, coef, nseg, 0.1, flags, tiles)

                print(f"{tag} rho={rho} {cfg}: fwd {time_ms(fwd):.4f} ms  bwd {time_ms(bwd):.4f} ms", flush=True)


if __name__ == "__main__":
    main()
