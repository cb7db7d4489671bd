#!/usr/bin/env python
"""bench.py — chaos-game iterations/s on the 4K still of BASELINE.json configs[1].

One step = one full frame of the shipped genome (electricsheep.247.11256) at 3840x2160:
warmup (1 first-run + 16 passes), draw_to_bins calls of 128 passes until 2000 samples/pixel are
binned (the `accumulated < quality*W*H` rule of src/main.cpp:411), density estimation + tonemap.
P = 2 097 152 particles, 512 temporal samples (src/main.cpp:203).

  value : whole-job iterations/s with everything resident in HBM (genome parameters uploaded,
          histogram and image on the device), CUDA-event timed, max over ranks.
  e2e   : the same frame through the C-ABI host-buffer call rfk_render_frame: parameter upload,
          device allocation, all kernels and the read-back of the RGBA8 image into pinned host
          memory inside the timed region.
  roofline : rfk_draw (the dominant kernel), algorithmic bytes / its CUDA-event time.
  cpu_baseline / --impl reference : the oracle (a CPU port of the reference's GLSL path; the
          reference itself needs OpenGL and cannot run here) on the host cores, bounded sample.

N > 1 (torchrun): every rank renders the full 2000 spp from disjoint RNG seed ranges into a private
histogram (weak scaling); the histograms are summed onto rank 0 with one NCCL reduce before density
estimation.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FIX = os.path.join(ROOT, "tests", "fixtures")
GENOME = os.path.join(FIX, "electricsheep.247.11256.flam3")
VARIATIONS = os.path.join(FIX, "variations.yaml")

W, H = 3840, 2160
QUALITY = 2000
P, TS, NSHUF = 2048 * 1024, 512, 1024
WARMUP_PASSES, DRAW_PASSES = 16, 128
TSS_WIDTH = 1.2 / 60.0
METRIC = "chaos_game_iterations_per_second"
UNIT = "iterations/s"
WORKLOAD = "configs[1]: electricsheep.247.11256 still, 3840x2160, 2000 samples/pixel, P=2097152, TS=512, warmup 16 + draw 128 passes/call, density estimation + tonemap"


def draw_counters():
    """profiles/r01_rfk_draw.json (tools/summarize_profiles.py over the `ncu --set full` capture of rfk_draw in this very
    workload): DRAM bytes of one launch (dram__bytes_read.sum + dram__bytes_write.sum) and executed warp instructions per
    warp-iteration. None when the file is missing: the derived fields are then left out rather than guessed."""
    path = os.path.join(ROOT, "profiles", "r01_rfk_draw.json")
    return json.load(open(path)) if os.path.exists(path) else None


def roofline_probes():
    """profiles/r01_roofline_probes.json: measured FFMA issue rate and red.global.add.v4.f32 rates (tools/roofline_probes.cu)"""
    path = os.path.join(ROOT, "profiles", "r01_roofline_probes.json")
    return json.load(open(path)) if os.path.exists(path) else None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", d
    return 6650.0, "fallback (B200_PROFILING.md)", {}


class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region"""

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], threading.Event()
        self.thread = threading.Thread(target=self.run, daemon=True)

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.thread.join(timeout=6)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 2 + k and r[2 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def cpu_reference_run(steps, warmup, sample_passes=8, as_main=False):
    """The oracle (CPU port of the reference path) on the host cores: one warmup + one draw_to_bins call of
    `sample_passes` passes at the bench's P/TS/4K histogram + density estimation + tonemap per step."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import refrakt_oracle as ro

    vt = ro.VariationTable(VARIATIONS)
    orc = ro.Oracle(ro.load_flame(GENOME, vt), vt, native=True)
    cores = os.cpu_count() or 1
    threads = max(1, min(cores, orc.max_threads(), int(6e9 // (W * H * 16))))  # private histograms: 132.7 MB per thread
    orc.set_threads(threads)
    orc.set_sim_parameters(P, TS, 64)  # 64 shuffle buffers instead of 1024: seeding cost only, not timed
    iters_per_step = P * (1 + WARMUP_PASSES + sample_passes)
    times = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        orc.warmup(WARMUP_PASSES, TSS_WIDTH)
        bins = np.zeros((H, W, 4), dtype=np.float32)
        orc.draw_to_bins(bins, W, sample_passes)
        img = orc.tonemap(orc.density_estimate(bins, W, H))
        ro.to_rgba8(img)
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
    mean = sum(times) / len(times)
    sample = "per step: warmup(16) + one draw_to_bins of %d passes (P=%d, TS=%d) into the 3840x2160 histogram + density estimation + tonemap = %d iterations; not scaled" % (
        sample_passes, P, TS, iters_per_step)
    return {"value": iters_per_step / mean, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample, "seconds_per_step": mean}


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    base = cpu_reference_run(args.steps, args.warmup)
    line = {"metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": base["seconds_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "reference", "config": {"workload": WORKLOAD, "note": "CPU port of the reference's GLSL path (oracle/, OpenMP): bit-identical in output to the reference's own host code + shaders run over the software GL of oracle/softgl/, which is single-threaded and needs the reference's files, so it cannot be timed on this box (DESIGN.md sections 6-7)"},
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--quality", type=int, default=QUALITY, help=argparse.SUPPRESS)  # debugging only; the bench line uses 2000
    ap.add_argument("--no-cpu-baseline", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import refrakt_b200 as r
    from refrakt_b200 import sharding

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the render path has no CPU fallback (use --impl reference for the CPU port)")
    torch.cuda.set_device(local_rank)
    r.lib().rfk_set_device(local_rank)
    json_fd = 1
    if world > 1:
        # stdout carries exactly one JSON line: NCCL writes its version banner (and anything NCCL_DEBUG asks for) to file
        # descriptor 1 when its first communicator comes up, so for the length of the run descriptor 1 is stderr and the
        # line goes to a duplicate of the real stdout
        sys.stdout.flush()
        json_fd = os.dup(1)
        os.dup2(2, 1)
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    compiler = r.FlameCompiler(VARIATIONS)
    flame = r.Flame.load_flame(GENOME, compiler)
    if flame is None:
        raise SystemExit("genome failed to load: " + r.Flame.last_error())
    # rank g seeds particle slots [g*P, (g+1)*P): disjoint JSF32 streams
    r.set_sim_parameters(P, TS, NSHUF, seed=sharding.rank_seed(rank, P))

    nbins = W * H
    target = args.quality * nbins
    bins = torch.zeros(nbins * 4, dtype=torch.float32, device="cuda")
    image = torch.empty(nbins * 4, dtype=torch.float32, device="cuda")
    rgba8 = torch.empty(nbins * 4, dtype=torch.uint8, device="cuda")
    host_rgba8 = torch.empty(nbins * 4, dtype=torch.uint8).pin_memory()
    post = flame.post_params()
    draw_events, post_events = [], []

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def resident_step(record):
        """the frame with everything on the device; returns (iterations, binned, draw calls)"""
        flame.warmup(WARMUP_PASSES, TSS_WIDTH)
        bins.zero_()
        binned, calls = 0, 0
        while binned < target:
            if record:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            flame.draw_to_bins_async(bins.data_ptr(), nbins, W, DRAW_PASSES)
            if record:
                e1.record()
            total = flame.binned_total()  # blocking 8-byte read-back, as flame.cpp:329
            if record:
                draw_events.append((e0, e1, total - binned))
            binned = total
            calls += 1
        sharding.reduce_histogram(bins, dst=0)  # the one exchange step (NCCL, 132.7 MB per rank)
        if rank == 0:
            if record:
                p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                p0.record()
            r.density_tonemap(bins.data_ptr(), image.data_ptr(), rgba8.data_ptr(), W, H, post)
            if record:
                p1.record()
                post_events.append((p0, p1))
        return P * (1 + WARMUP_PASSES + DRAW_PASSES * calls), binned, calls

    def e2e_step():
        """the same frame through the host-buffer C-ABI call"""
        _, stats = flame.render_frame(W, H, target_binned=target, warmup_passes=WARMUP_PASSES, drawing_passes=DRAW_PASSES,
                                      tss_width=TSS_WIDTH, rgba8_out=host_rgba8.numpy().reshape(H, W, 4))
        return P * (1 + WARMUP_PASSES) + stats.iterations, stats

    for _ in range(args.warmup):
        resident_step(False)

    launches0 = r.kernel_launch_count()
    barrier()
    with ClockSampler(local_rank) as clocks:
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        iters = binned = calls = 0
        for _ in range(args.steps):
            i, b, c = resident_step(True)
            iters += i; binned += b; calls += c
        t1.record()
        barrier()
        ms = t0.elapsed_time(t1)
    launches = r.kernel_launch_count() - launches0

    draw_ms = [e0.elapsed_time(e1) for e0, e1, _ in draw_events]
    draw_binned = [b for _, _, b in draw_events]

    # e2e (every rank renders; rank 0's number is reported at N=1, whole-job aggregate at N>1)
    for _ in range(min(args.warmup, 1)):
        e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_iters = 0
    for _ in range(args.steps):
        i, stats = e2e_step()
        e2e_iters += i
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)

    stat = torch.tensor([ms, e2e_ms, float(iters), float(e2e_iters), float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = stat.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stat.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, e2e_ms = float(mx[0]), float(mx[1])
        iters, e2e_iters, launches = float(sm[2]), float(sm[3]), int(sm[4])

    if rank == 0:
        peak, peak_src, peaks = measured_peaks()
        counters = draw_counters()
        mean_draw_ms = sum(draw_ms) / len(draw_ms)
        mean_binned = sum(draw_binned) / len(draw_binned)
        # algorithmic bytes of one rfk_draw launch: 16 B per binned sample (one float4 reduction) + particle and
        # RNG state in and out once per particle (2 x (16 + 16) B); DESIGN.md "rfk_draw"
        alg_bytes = 16.0 * mean_binned + 64.0 * P
        achieved = alg_bytes / (mean_draw_ms * 1e-3) / 1e9
        sm_mhz = clocks.summary()["sm_mhz"] or peaks.get("sm_max_mhz", 1965.0)
        line = {
            "metric": METRIC, "value": iters / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "histogram_bytes": nbins * 16, "l2_note": "inputs larger than L2: 132.7 MB histogram + 64 MB particle/RNG state per launch",
                       "quality": args.quality, "draw_calls_per_step": calls / args.steps, "in_bounds_fraction": binned / max(1.0, calls * P * DRAW_PASSES),
                       "frame_ms": ms / args.steps},
            "e2e": {"value": e2e_iters / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 1024 * 4 + 256 * 16, "d2h_bytes_per_step": nbins * 4 + 8 * int(calls / args.steps),
                    "ms_per_step": e2e_ms / args.steps, "api": "rfk_render_frame (C ABI, host RGBA8 buffer, pinned)"},
            "gpu_launches": launches,
            "clocks": clocks.summary(),
            "roofline": {"kernel": "rfk_draw", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": counters["dram_bytes_per_launch"] if counters else None, "peak_source": peak_src, "launch_ms": mean_draw_ms, "launches_timed": len(draw_ms),
                         "algorithmic_bytes_per_launch": alg_bytes, "share_of_step": sum(draw_ms) / ms,
                         "binding_limit": "fp32/alu issue, not memory: see DESIGN.md and profiles/",
                         "iterations_per_s_kernel": P * DRAW_PASSES / (mean_draw_ms * 1e-3)},
        }
        if post_events:
            # density estimation + tonemap: 16 B/pixel histogram read + 16 B/pixel float image + 4 B/pixel RGBA8 written (32 B
            # for the float image alone); HBM-bound by design, instruction-bound as measured (profiles/r01_density_tonemap.md)
            post_ms = sum(a.elapsed_time(b) for a, b in post_events) / len(post_events)
            post_bytes = 36.0 * nbins
            line["roofline"]["post"] = {"kernel": "density_tonemap_kernel", "bound": "hbm", "launch_ms": post_ms, "algorithmic_bytes_per_launch": post_bytes,
                                        "achieved": post_bytes / (post_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": post_bytes / (post_ms * 1e-3) / 1e9 / peak,
                                        "share_of_step": post_ms * len(post_events) / ms}
        probes = roofline_probes()
        if probes and counters:
            kernel_iters_s = P * DRAW_PASSES / (mean_draw_ms * 1e-3)
            winst = kernel_iters_s / 32.0 * counters["warp_inst_per_unit"]
            line["roofline"]["issue"] = {"achieved_warp_inst_per_s": winst, "peak_measured_ffma_issue": probes["ffma_warp_inst_per_s"],
                                         "frac": winst / probes["ffma_warp_inst_per_s"], "warp_inst_per_warp_iteration": counters["warp_inst_per_unit"],
                                         "source": "ncu smsp__inst_executed.sum (profiles/r01_rfk_draw.json) x CUDA-event kernel rate; peak = tools/roofline_probes.cu"}
        if probes:
            red = probes["red_v4_f32"]["132.7MB_4K"]["v4_gred_per_s"]
            line["roofline"]["atomics"] = {"achieved_gred_per_s": mean_binned / (mean_draw_ms * 1e-3) / 1e9, "uniform_random_probe_gred_per_s": red,
                                           "note": "red.global.add.v4.f32 at random addresses over the same 132.7 MB footprint; a flame's hits are concentrated, so the kernel can exceed it"}
        if not args.no_cpu_baseline and world == 1:
            try:
                base = cpu_reference_run(1, 0)
                line["cpu_baseline"] = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
            except Exception as e:  # the bench line must still print
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: %s" % e}
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
