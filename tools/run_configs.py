"""Runs the BASELINE.json configurations on the GPUs of one box and prints one JSON line per config.
  python tools/run_configs.py 1 2 3 5            (single GPU)
  torchrun --nproc-per-node N tools/run_configs.py 3 4 5   (N GPUs: histogram reduce / frame-parallel animation)
Config 4 renders `--frames` animation frames (default 16 per rank instead of 600 in total; the rate is per frame)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist

import refrakt_b200 as r
from refrakt_b200 import sharding

FIX = os.path.join(ROOT, "tests", "fixtures")
GENOME = os.path.join(FIX, "electricsheep.247.11256.flam3")
VARIATIONS = os.path.join(FIX, "variations.yaml")
P, TS = 2048 * 1024, 512
TSS = 1.2 / 60.0
HOT_BUDGET = 0


def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fn()
    e1.record()
    torch.cuda.synchronize()
    return out, e0.elapsed_time(e1)


def render_still(flame, W, H, target_binned=None, draw_calls=None, rank=0, world=1, downsample=False, hot_map=False):
    """warmup + draw + (reduce) + DE/tonemap; returns stage times in ms"""
    n = W * H
    bins = torch.zeros(n * 4, dtype=torch.float32, device="cuda")
    _, ms_warm = timed(lambda: flame.warmup(16, TSS))

    def draw():
        binned = calls = 0
        while (draw_calls is None or calls < draw_calls) and (target_binned is None or binned < target_binned):
            flame.draw_to_bins_async(bins.data_ptr(), n, W, 128)
            calls += 1
            if hot_map and calls == 1:  # kernel option l2_hints: classify the tiles once the first call has landed (inside the timed region)
                flame.build_hot_map(bins.data_ptr(), n, W, HOT_BUDGET)
            if target_binned is not None:
                binned = flame.binned_total()
        return flame.binned_total(), calls
    (binned, calls), ms_draw = timed(draw)
    _, ms_reduce = timed(lambda: sharding.reduce_histogram(bins, dst=0))
    ms_post = 0.0
    if rank == 0:
        image = torch.empty(n * 4, dtype=torch.float32, device="cuda")
        rgba8 = torch.empty(n * 4, dtype=torch.uint8, device="cuda")
        post = flame.post_params()
        small = torch.empty(n, dtype=torch.float32, device="cuda") if downsample else None  # allocated outside the timed region

        def post_fn():
            r.density_tonemap(bins.data_ptr(), image.data_ptr(), rgba8.data_ptr(), W, H, post)
            if downsample:
                r.downsample2x(image.data_ptr(), small.data_ptr(), W // 2, H // 2)
        _, ms_post = timed(post_fn)
    return dict(binned=int(binned), draw_calls=calls, iterations=P * (17 + 128 * calls), ms_warmup=ms_warm, ms_draw=ms_draw, ms_reduce=ms_reduce, ms_post=ms_post)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="+", type=int)
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--hot-budget-mb", type=float, default=0.0, help="config 3: hot-map budget (0 = the library default, half the L2)")
    ap.add_argument("--no-l2-hints", action="store_true", help="config 3 without the hot map / evict-first reductions")
    ap.add_argument("--staged", type=int, default=-1, help="config 3: kernel option staged_bins: -1 = automatic (the default: queues, regions of 2^22 bins), n = log2(bins per region), 0 = direct reductions with the L2 hot map")
    ap.add_argument("--draw-calls", type=int, default=256, help="config 3: draw calls per GPU")
    args = ap.parse_args()
    global HOT_BUDGET
    HOT_BUDGET = int(args.hot_budget_mb * 1e6)
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    r.lib().rfk_set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    compiler = r.FlameCompiler(VARIATIONS, overlay=r.OVERLAY_YAML)
    flame = r.Flame.load_flame(GENOME, compiler)
    r.set_sim_parameters(P, TS, 1024, seed=sharding.rank_seed(rank, P))

    def emit(cfg, desc, res, iters_all_ranks, ms_total):
        if world > 1:
            t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_total = float(t[0])
        if rank == 0:
            print(json.dumps(dict(config=cfg, desc=desc, n_gpus=world, iterations_per_s=iters_all_ranks / (ms_total * 1e-3), ms_total=ms_total, **res)), flush=True)

    for cfg in args.configs:
        if world > 1:
            dist.barrier()
        if cfg == 1:
            render_still(flame, 1280, 720, draw_calls=1)
            res = render_still(flame, 1280, 720, draw_calls=1, rank=rank, world=world)
            ms = res["ms_warmup"] + res["ms_draw"] + res["ms_reduce"] + res["ms_post"]
            emit(1, "shipped genome, 1280x720, 1 warmup + 128 draw passes (268 435 456 iterations) + DE + tonemap", res, res["iterations"] * world, ms)
        elif cfg == 2:
            render_still(flame, 3840, 2160, draw_calls=2)
            res = render_still(flame, 3840, 2160, target_binned=2000 * 3840 * 2160, rank=rank, world=world)
            ms = res["ms_warmup"] + res["ms_draw"] + res["ms_reduce"] + res["ms_post"]
            emit(2, "shipped genome, 3840x2160, 2000 spp, DE + tonemap (4K frame ms = ms_total)", res, res["iterations"] * world, ms)
        elif cfg == 3:
            hints = 0 if (args.no_l2_hints or args.staged != 0) else 1
            flame.set_options(l2_hints=hints, staged_bins=args.staged)
            render_still(flame, 15360, 8640, draw_calls=1)
            res = render_still(flame, 15360, 8640, draw_calls=args.draw_calls, rank=rank, world=world, downsample=True, hot_map=bool(hints))
            res["l2_hints"] = hints
            res["staged_bins"] = args.staged
            res["hot_budget_mb"] = args.hot_budget_mb
            flame.set_options(l2_hints=0, staged_bins=-1)
            flame.clear_hot_map()
            ms = res["ms_warmup"] + res["ms_draw"] + res["ms_reduce"] + res["ms_post"]
            emit(3, "shipped genome, 15360x8640 histogram (2x supersampled 7680x4320), %d draw calls per GPU, NCCL reduce, DE + tonemap + 2x2 box" % args.draw_calls, res, res["iterations"] * world, ms)
        elif cfg == 4:
            # frame-parallel animation: frame f rotated by 18 deg/s * f/60, per frame warmup(16) + 128 passes + DE + tonemap at 720p
            W, H = 1280, 720
            frames = sharding.frames_of_rank(args.frames * world, world, rank)
            bins = torch.zeros(W * H * 4, dtype=torch.float32, device="cuda")
            rgba8 = torch.empty(W * H * 4, dtype=torch.uint8, device="cuda")
            host = torch.empty(W * H * 4, dtype=torch.uint8).pin_memory()
            post = flame.post_params()

            def run():
                prev = 0
                for f in frames:
                    flame.rotate_xforms(18.0 * (f - prev) / 60.0)
                    prev = f
                    flame.warmup(16, TSS)
                    bins.zero_()
                    flame.draw_to_bins_async(bins.data_ptr(), W * H, W, 128)
                    r.density_tonemap(bins.data_ptr(), None, rgba8.data_ptr(), W, H, post)
                    host.copy_(rgba8, non_blocking=True)
                torch.cuda.synchronize()
            _, ms = timed(run)
            res = dict(frames_per_rank=len(frames), ms_per_frame=ms / len(frames), iterations=P * 145 * len(frames))
            emit(4, "animation, 1280x720, frames round-robin over ranks, each warmup(16) + 128 passes + DE + tonemap + read-back", res, res["iterations"] * world, ms)
            flame = r.Flame.load_flame(GENOME, compiler)  # undo the rotation
        elif cfg == 5:
            from conftest import stress_genome  # genome generator only

            class _Params:  # the generator needs each variation's parameter names: taken from the product's compiler
                def __init__(self, names): self.param = names

            class _Table:
                def __init__(self, comp): self.vars = {n: _Params(comp.get_parameters_for_variation(n)) for n in comp.variations()}
            vt = _Table(compiler)
            stress = r.Flame.load_flame_string(stress_genome(vt), compiler)
            assert stress is not None, r.Flame.last_error()
            render_still(stress, 3840, 2160, draw_calls=2)
            res = render_still(stress, 3840, 2160, draw_calls=64, rank=rank, world=world)
            ms = res["ms_warmup"] + res["ms_draw"] + res["ms_reduce"] + res["ms_post"]
            res["kernel_info"] = stress.kernel_info("rfk_draw")
            emit(5, "synthetic stress genome (12 xforms + final: julian/juliascope/trig/bipolar...), 3840x2160, 64 draw calls per GPU", res, res["iterations"] * world, ms)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
