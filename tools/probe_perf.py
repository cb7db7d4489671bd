"""Exploratory timing of the chaos-game kernel variants on one GPU (not the bench):
iterations/s of rfk_draw for several option sets and histogram sizes, CUDA-event timed."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import refrakt_b200 as r

FIX = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "fixtures")


def time_draw(flame, bins, W, H, passes, reps):
    flame.draw_to_bins(bins.data_ptr(), W * H, W, passes)  # warm
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        flame.draw_to_bins_async(bins.data_ptr(), W * H, W, passes)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    out = []
    compiler = r.FlameCompiler(os.path.join(FIX, "variations.yaml"))
    flame = r.Flame.load_flame(os.path.join(FIX, "electricsheep.247.11256.flam3"), compiler)
    P, TS = 2048 * 1024, 512
    r.set_sim_parameters(P, TS, 1024)
    variants = [
        ("default", {}),
        ("min_blocks_compiler_choice", dict(min_blocks=-1)),
        ("min_blocks4", dict(min_blocks=4)),
        ("min_blocks6", dict(min_blocks=6)),
        ("min_blocks7", dict(min_blocks=7)),
        ("min_blocks8", dict(min_blocks=8)),
        ("block128", dict(block_width=128)),
        ("block512", dict(block_width=512)),
        ("deal_period2", dict(deal_period=2)),
        ("deal_period4", dict(deal_period=4)),
        ("warp_aggregate", dict(warp_aggregate=1)),
        ("math0_libdevice", dict(math_mode=0)),
        ("math2_fast", dict(math_mode=2)),
        ("per_lane", dict(per_lane_xform=1)),
        ("deterministic", dict(deterministic=1)),
    ]
    sizes = [(1280, 720), (3840, 2160)]
    if "--big" in sys.argv:
        sizes.append((15360, 8640))
    for name, opt in variants:
        base = dict(math_mode=1, fmad=1, per_lane_xform=0, warp_aggregate=0, deterministic=0, count_xforms=0, min_blocks=0, block_width=256, deal_period=1)
        base.update(opt)
        t0 = time.time()
        flame.set_options(**base)
        flame.warmup(16, 1.2 / 60)
        info = flame.kernel_info("rfk_draw")
        for W, H in sizes:
            bins = torch.zeros(W * H * 4, dtype=torch.float32, device="cuda")
            passes = 128 if name != "per_lane" else 16
            ms = time_draw(flame, bins, W, H, passes, 3)
            before = flame.binned_total()
            flame.draw_to_bins(bins.data_ptr(), W * H, W, passes)
            frac = (flame.binned_total() - before) / (P * passes)
            rec = dict(variant=name, W=W, H=H, ms_per_call=ms, giter_s=P * passes / ms / 1e6, in_bounds=frac, **info, compile_s=time.time() - t0)
            print(json.dumps(rec), flush=True)
            out.append(rec)
            del bins
    # same-hardware baseline: the reference's structure (one iteration per launch, state through global memory)
    flame.set_options(math_mode=1, fmad=1, per_lane_xform=0, warp_aggregate=0, deterministic=0, count_xforms=0, min_blocks=0, block_width=256, deal_period=1)
    for W, H in sizes:
        bins = torch.zeros(W * H * 4, dtype=torch.float32, device="cuda")
        flame.reference_warmup(16, 1.2 / 60)
        flame.reference_draw_to_bins(bins.data_ptr(), W * H, W, 8)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        flame.reference_draw_to_bins(bins.data_ptr(), W * H, W, 128)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        rec = dict(variant="reference_pass_mode (flame.glsl structure in CUDA: 128 launches, state in HBM)", W=W, H=H, ms_per_call=ms, giter_s=P * 128 / ms / 1e6)
        print(json.dumps(rec), flush=True)
        out.append(rec)
        del bins
    # warm kernel (no histogram): the pure iteration rate
    flame.set_options(math_mode=1, fmad=1, per_lane_xform=0, warp_aggregate=0, deterministic=0, count_xforms=0, min_blocks=0, block_width=256, deal_period=1)
    torch.cuda.synchronize()
    t0 = time.time()
    flame.warmup(256, 1.2 / 60)
    dt = time.time() - t0
    print(json.dumps(dict(variant="warmup256", giter_s=P * 257 / dt / 1e9, s=dt)), flush=True)
    # density + tonemap
    for W, H in sizes:
        bins = torch.zeros(W * H * 4, dtype=torch.float32, device="cuda")
        flame.warmup(16, 1.2 / 60)
        for _ in range(4 if W < 2000 else 30):
            flame.draw_to_bins_async(bins.data_ptr(), W * H, W, 128)
        img = torch.empty(W * H * 4, dtype=torch.float32, device="cuda")
        u8 = torch.empty(W * H * 4, dtype=torch.uint8, device="cuda")
        p = flame.post_params()
        for label, args in (("de+tonemap f4", (bins.data_ptr(), img.data_ptr(), None)), ("de+tonemap u8", (bins.data_ptr(), None, u8.data_ptr()))):
            r.density_tonemap(*args, W, H, p)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(5):
                r.density_tonemap(*args, W, H, p)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            gb = W * H * (32 if "f4" in label else 20) / 1e9
            print(json.dumps(dict(variant=label, W=W, H=H, ms=ms, gbs=gb / ms * 1e3, nonzero=float((bins.view(-1, 4)[:, 3] > 0).float().mean()))), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/probe_perf.json", "w"), indent=1)


if __name__ == "__main__":
    main()
