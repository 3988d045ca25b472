"""Scorer data paths side by side (development / evidence aid; the contract benchmark is bench.py).

  * rows not 16-byte aligned (VOC crop 513x513x22): plain register path ("ldg") vs the abreast kernel;
  * VOC native 375x500x22: TMA ring vs abreast;
  * low-resolution entry: head logits 256x512 -> 1024x2048 interpolated inside the kernel, against
    the full-resolution TMA kernel alone and against F.interpolate + the full-resolution kernel (what the reference's
    model + selector do), fp32 and bf16.
CUDA events, inputs rotated over more than the 126 MB L2.  Writes gpurun_out/kbench_paths.json.
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mulactseg_b200 import acquisition as acq, synth  # noqa: E402

DEV = "cuda:0"


def time_ms(fn, warmup=3, iters=10):
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def full_res(res, name, c, h, w, nseg, n_img, paths, dtype=torch.float32):
    logits = synth.logits(n_img, c, h, w, "cosine", seed=1, device=DEV).to(dtype)
    spx = synth.superpixel_map(n_img, h, w, nseg, "jitter", seed=2, device=DEV, dtype=torch.int32)
    gb = n_img * (h * w * (c * logits.element_size() + 4) + nseg * c * 8) / 1e9
    for path in paths:
        if path == "default":
            os.environ.pop("MAS_SCORER_PATH", None)
        else:
            os.environ["MAS_SCORER_PATH"] = path
        stats = acq.RegionStats(n_img, nseg, c, DEV, need_prob=True, lanes=2, group_bytes=1536 << 20)

        def run(i):
            for j in range(0, n_img, 4):
                stats.add_batch(j, logits[j:j + 4], spx[j:j + 4], 0.1)
            stats.join()

        ms = time_ms(run)
        res[f"{name} path={path}"] = {"ms_per_round": round(ms, 4), "GBps": round(gb / ms * 1e3, 1), "images": n_img}
        print(f"{name} path={path}", res[f"{name} path={path}"], flush=True)
    os.environ.pop("MAS_SCORER_PATH", None)


def low_res(res, c, dtype):
    n_img, h, w, nseg = 96, 1024, 2048, 2048
    low = synth.logits(n_img, c, h // 4, w // 4, "cosine", seed=3, device=DEV).to(dtype)
    spx = synth.superpixel_map(n_img, h, w, nseg, "jitter", seed=4, device=DEV, dtype=torch.int32)
    stats = acq.RegionStats(n_img, nseg, c, DEV, need_prob=True, lanes=2)
    tag = f"lowres C={c} {str(dtype)[6:]}"

    def fused(i):
        for j in range(0, n_img, 4):
            stats.add_batch_lowres(j, low[j:j + 4], spx[j:j + 4], 0.1)
        stats.join()

    ms = time_ms(fused)
    res[f"{tag} fused (interpolate inside the scorer)"] = {"ms_per_image": round(ms / n_img, 5)}

    def reference_way(i):
        for j in range(0, n_img, 4):
            full = torch.nn.functional.interpolate(low[j:j + 4], size=(h, w), mode="bilinear", align_corners=False)
            stats.add_batch(j, full, spx[j:j + 4], 0.1)
        stats.join()

    ms2 = time_ms(reference_way)
    res[f"{tag} F.interpolate + full-resolution scorer"] = {"ms_per_image": round(ms2 / n_img, 5)}
    full = torch.nn.functional.interpolate(low[:24], size=(h, w), mode="bilinear", align_corners=False)

    def scorer_only(i):
        for j in range(0, 24, 4):
            stats.add_batch(j, full[j:j + 4], spx[j:j + 4], 0.1)
        stats.join()

    ms3 = time_ms(scorer_only)
    res[f"{tag} full-resolution scorer alone (HBM-bound)"] = {"ms_per_image": round(ms3 / 24, 5)}
    for k in list(res)[-3:]:
        print(k, res[k], flush=True)


def labeller_sources(res):
    """Prototype labeller, one Cityscapes-shaped image: feature sources side by side."""
    from mulactseg_b200 import labeller
    h, w, nseg, c, rho = 1024, 2048, 2048, 20, 0.08
    lows = [torch.nn.functional.normalize(torch.randn((1, 256, h // 4, w // 4), device=DEV,
                                                       generator=torch.Generator(device=DEV).manual_seed(i)), dim=1) for i in range(4)]
    logits = [synth.logits(1, c, h, w, "normal", seed=30 + i, device=DEV, coherent=4) for i in range(4)]
    spx = synth.superpixel_map(1, h, w, nseg, "jitter", seed=3, device=DEV)
    trg = synth.multihot_targets(1, nseg, c, seed=4, device=DEV, p_ignore=0.0)
    mask = synth.region_mask(spx, nseg, rho, seed=5)
    for name in ("full f32", "full bf16", "lowres f32", "lowres bf16", "F.interpolate + full f32"):
        if name.startswith("full"):
            feats = [torch.nn.functional.interpolate(x, size=(h, w), mode="bilinear", align_corners=False) for x in lows]
            if "bf16" in name:
                feats = [f.to(torch.bfloat16) for f in feats]
        else:
            feats = [x.to(torch.bfloat16) if "bf16" in name else x for x in lows]

        def run(i):
            f = feats[i % 4]
            if name.startswith("F.interpolate"):
                f = torch.nn.functional.interpolate(f, size=(h, w), mode="bilinear", align_corners=False)
            labeller.pseudo_label_generation(None, f, logits[i % 4], trg, mask, spx, check=False)

        ms = time_ms(run, iters=12)
        res[f"labeller cityscapes feats={name}"] = {"ms_per_image": round(ms, 4)}
        print(f"labeller cityscapes feats={name}", res[f"labeller cityscapes feats={name}"], flush=True)
        del feats


def main():
    res = {}
    if "labeller" in sys.argv[1:]:
        labeller_sources(res)
        print(json.dumps(res, indent=1))
        return
    if "voc513" in sys.argv[1:]:          # short: for ncu
        full_res(res, "voc_crop 513x513x22", 22, 513, 513, 150, 64, ["default"])
        return
    if "lowres" in sys.argv[1:]:
        low_res(res, 19, torch.float32)
        return
    full_res(res, "voc_crop 513x513x22", 22, 513, 513, 150, 512, ["ldg", "default"])
    for ahead in ("0", "1", "4"):
        os.environ["MAS_SCORER_AHEAD"] = ahead
        full_res(res, f"voc_crop 513x513x22 ahead={ahead}", 22, 513, 513, 150, 512, ["default"])
    os.environ.pop("MAS_SCORER_AHEAD", None)
    full_res(res, "voc_native 375x500x22", 22, 375, 500, 150, 512, ["default", "abreast", "ldg"])
    full_res(res, "cityscapes 1024x2048x19", 19, 1024, 2048, 2048, 48, ["default", "abreast"])
    full_res(res, "cityscapes bf16 1024x2048x19", 19, 1024, 2048, 2048, 48, ["default", "abreast"], torch.bfloat16)
    for c in (19, 20):
        for dtype in (torch.float32, torch.bfloat16):
            low_res(res, c, dtype)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/kbench_paths.json", "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
