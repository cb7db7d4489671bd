#!/usr/bin/env python
"""bench.py — chaos-game iterations/s and frame time on the BASELINE.json configurations.

Default = configs[1]: ONE frame of the shipped genome (electricsheep.247.11256) at 3840x2160: warmup (1 first-run + 16
passes), draw_to_bins calls of 128 passes until 2000 samples/pixel are binned (the `accumulated < quality*W*H` rule of
src/main.cpp:411), density estimation + tonemap. P = 2 097 152 particles per GPU, 512 temporal samples (src/main.cpp:203).
One step = one frame.

  value : whole-job iterations/s with everything resident in HBM, CUDA-event timed, max over ranks.
  e2e   : the same frame through the C-ABI host-buffer call (rfk_render_frame; rfk_render_frame_sharded at N > 1):
          parameter upload, all kernels, the exchange between the GPUs and the read-back of the RGBA8 image into
          pinned host memory inside the timed region.
  roofline : rfk_draw, the dominant kernel — algorithmic FP32 flop / its CUDA-event time against the measured FFMA
          rate, with the issue-slot, L2-reduction and DRAM fractions and the density + tonemap kernel as flat keys.
  cpu_baseline / --impl reference : the oracle (a CPU port of the reference's GLSL path; the reference itself needs
          OpenGL and cannot run here) on all host cores, bounded sample of the same frame.

N > 1 (torchrun): STRONG scaling — the N GPUs render ONE frame together: disjoint particle streams, each rank 1/N of the
passes, histograms reduce-scattered over row slabs (peer memory over NVLink, NCCL as the fallback), density estimation
+ tonemap on H/N rows per rank, rows gathered on rank 0 (csrc/comm.cpp). `weak` holds the round-1 figure beside it:
every rank renders the full 2000 spp and the histograms are summed onto rank 0 with one NCCL reduce.

--config K selects another BASELINE configuration (1-5; see CONFIGS)."""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

P, TS, NSHUF = 2048 * 1024, 512, 1024
WARMUP_PASSES, DRAW_PASSES = 16, 128
TSS_WIDTH = 1.2 / 60.0
METRIC = "chaos_game_iterations_per_second"
UNIT = "iterations/s"
FLOP_PER_ITERATION_SHIPPED = 122.0  # algorithmic FP32 flop of one drawn iteration of the shipped genome

# BASELINE.json `configs`, numbered from 1. mode "quality": draw until quality * bins samples are binned; "calls": a fixed
# number of 128-pass draw calls per GPU; "animation": frames of warmup + one draw call + density/tonemap + read-back.
CONFIGS = {
    1: dict(genome="shipped", W=1280, H=720, ss=1, mode="calls", calls=1,
            workload="configs[0]: electricsheep.247.11256 still, 1280x720, 1 warmup + one draw_to_bins of 128 passes (268 435 456 iterations), density estimation + tonemap"),
    2: dict(genome="shipped", W=3840, H=2160, ss=1, mode="quality", quality=2000,
            workload="configs[1]: electricsheep.247.11256 still, 3840x2160, 2000 samples/pixel, P=2097152, TS=512, warmup 16 + draw 128 passes/call, density estimation + tonemap"),
    3: dict(genome="shipped", W=7680, H=4320, ss=2, mode="calls", calls=128,
            workload="configs[2]: electricsheep.247.11256, 7680x4320 image from a 2x supersampled 15360x8640 histogram (2.12 GB), 128 draw calls of 128 passes per GPU, histogram sum over the GPUs, density estimation + tonemap + spatial filter"),
    4: dict(genome="shipped", W=1280, H=720, ss=1, mode="animation", frames=600,
            workload="configs[3]: 600-frame animation of electricsheep.247.11256 (18 deg/s at 60 fps), 1280x720, per frame warmup 16 + one draw call of 128 passes + density estimation + tonemap + read-back, frames round-robin over the GPUs"),
    # a fixed number of draw calls, not a samples-per-pixel target: this genome's particles overflow to NaN one after the other
    # and never come back (the reference does not reset them: the badval line of flame.glsl:70 is commented out), so the
    # rate at which samples land decays and a 2000 spp target is never reached — in the reference either
    5: dict(genome="stress", W=3840, H=2160, ss=1, mode="calls", calls=64,
            workload="configs[4]: synthetic stress genome (12 xforms + final xform: julian / juliascope / trig / bipolar ..., numpy default_rng(247)), 3840x2160, 64 draw calls of 128 passes per GPU, density estimation + tonemap"),
}


def data_path(name):
    """the two data files the product reads: shipped under refrakt_b200/data/ (the same files sit under tests/fixtures/)"""
    p = os.path.join(ROOT, "refrakt_b200", "data", name)
    return p if os.path.exists(p) else os.path.join(ROOT, "tests", "fixtures", name)


GENOME = data_path("electricsheep.247.11256.flam3")
VARIATIONS = data_path("variations.yaml")


def static_config(cfg_id):
    """`config` of the JSON line: names the workload only, identical for both arms and every N"""
    c = CONFIGS[cfg_id]
    return {"workload": c["workload"], "config_id": cfg_id, "histogram": "%dx%d" % (c["W"] * c["ss"], c["H"] * c["ss"]),
            "histogram_bytes": c["W"] * c["ss"] * c["H"] * c["ss"] * 16, "particles_per_gpu": P, "temporal_samples": TS,
            "l2_note": "inputs larger than L2 (126 MB): %.1f MB histogram + 64 MB particle/RNG state per draw launch" % (c["W"] * c["ss"] * c["H"] * c["ss"] * 16 / 1e6)
            if c["W"] * c["ss"] * c["H"] * c["ss"] * 16 + 64 * P > 126e6 else "L2 flushed between timed steps by the 64 MB particle/RNG state + the histogram clear of the next step"}


def profile_json(name):
    path = os.path.join(ROOT, "profiles", name)
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
            self.stop.wait(0.25)

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


# ---------------------------------------------------------------------------------------------------------------------
# the CPU arm: the oracle (test infrastructure; bench.py may execute it for this leg only)
# ---------------------------------------------------------------------------------------------------------------------
def stress_xml(names_to_params):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import stress_genome

    class _Params:
        def __init__(self, names): self.param = names

    class _Table:
        def __init__(self, table): self.vars = {n: _Params(p) for n, p in table.items()}
    return stress_genome(_Table(names_to_params))


def cpu_reference_run(cfg_id, steps, warmup, sample_passes=24, one_thread_probe=True):
    """The oracle on the host cores, the SAME frame definition in a bounded sample: per step warmup(16) + `sample_passes`
    drawn passes of all P particles into the configuration's histogram (every thread a private histogram, merged in
    parallel), density estimation + tonemap of that histogram. The chaos game, the merge and the post step are timed
    separately; `value` is the chaos-game rate (warmup + drawn iterations over their time), and the frame time of the whole
    configuration is extrapolated from it. The OpenMP thread count is set explicitly (torchrun exports OMP_NUM_THREADS=1)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import refrakt_oracle as ro
    import refrakt_b200 as r

    c = CONFIGS[cfg_id]
    overlay = r.OVERLAY_YAML if c["genome"] == "stress" else None
    vt = ro.VariationTable(VARIATIONS, overlay=overlay)
    if c["genome"] == "stress":
        of = ro.load_flame_string(stress_xml({n: list(v.param) for n, v in vt.vars.items()}), vt)
    else:
        of = ro.load_flame(GENOME, vt)
    orc = ro.Oracle(of, vt, native=True)
    W, H = c["W"] * c["ss"], c["H"] * c["ss"]
    if W * H * 16 > 600e6:  # the 2.12 GB histogram of config 3 times the thread count does not fit a host: sample at 4K
        W, H = 3840, 2160
    cores = os.cpu_count() or 1
    threads = max(1, min(cores, int(24e9 // (W * H * 16))))  # private histograms: 132.7 MB per thread at 4K
    orc.set_threads(threads)
    orc.set_sim_parameters(P, TS, 64)  # 64 shuffle buffers instead of 1024: seeding cost only, not timed
    rows = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        orc.warmup(WARMUP_PASSES, TSS_WIDTH)
        orc.draw_accumulate(W, H, sample_passes)
        t1 = time.perf_counter()
        bins = np.zeros((H, W, 4), dtype=np.float32)
        orc.merge_private(bins)
        t2 = time.perf_counter()
        img = orc.tonemap(orc.density_estimate(bins, W, H))
        ro.to_rgba8(img)
        t3 = time.perf_counter()
        if s >= warmup:
            rows.append((t1 - t0, t2 - t1, t3 - t2))
    chaos_s = min(x[0] for x in rows)  # the best step: the host cores are shared with other tenants of the box
    merge_s = sum(x[1] for x in rows) / len(rows)
    post_s = sum(x[2] for x in rows) / len(rows)
    iters = P * (1 + WARMUP_PASSES + sample_passes)
    rate = iters / chaos_s
    one_thread = None
    if one_thread_probe and threads > 1:
        orc.set_threads(1)
        t0 = time.perf_counter()
        orc.warmup(1, TSS_WIDTH)
        orc.draw_accumulate(W, H, 2)
        one_thread = P * 4 / (time.perf_counter() - t0)
        orc.set_threads(threads)
    orc.release_private()
    sample = ("per step: warmup(16) + %d drawn passes of P=%d particles (TS=%d) into the %dx%d histogram = %d iterations on %d OpenMP threads, then merge of the private "
              "histograms, density estimation + tonemap; chaos game, merge and post timed separately (%d untimed + %d timed steps, chaos game: the fastest); not scaled") % (sample_passes, P, TS, W, H, iters, threads, warmup, steps)
    return {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample, "threads": threads, "host_cores": cores,
            "one_thread_value": one_thread, "speedup_over_one_thread": (rate / one_thread) if one_thread else None,
            "chaos_seconds_per_step": chaos_s, "merge_ms": merge_s * 1e3, "post_ms": post_s * 1e3, "iterations_per_step": iters}


def frame_iterations_estimate(cfg_id, in_bounds=None):
    """iterations of one frame of the configuration on one GPU (for the CPU arm's extrapolated frame time)"""
    c = CONFIGS[cfg_id]
    if c["mode"] == "calls":
        return P * (1 + WARMUP_PASSES + DRAW_PASSES * c["calls"])
    if c["mode"] == "animation":
        return P * (1 + WARMUP_PASSES + DRAW_PASSES) * c["frames"]
    frac = in_bounds if in_bounds else (0.814 if c["genome"] == "shipped" else 0.392)  # measured in-bounds fractions at 4K
    calls = -(-c["quality"] * c["W"] * c["H"] // int(P * DRAW_PASSES * frac))
    return P * (1 + WARMUP_PASSES + DRAW_PASSES * calls)


def run_reference_arm(args, rank):
    if rank != 0:
        return
    base = cpu_reference_run(args.config, args.steps, args.warmup)
    frame_iters = frame_iterations_estimate(args.config)
    frame_ms = frame_iters / base["value"] * 1e3 + base["merge_ms"] + base["post_ms"]
    step_ms = base["chaos_seconds_per_step"] * 1e3 + base["merge_ms"] + base["post_ms"]
    line = {"metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "reference", "config": static_config(args.config),
            "frame_ms_extrapolated": frame_ms,
            "note": "CPU port of the reference's GLSL path (oracle/, OpenMP, all host cores): bit-identical in output to the reference's own host code + shaders run over the software GL of oracle/softgl/, which is single-threaded and needs the reference's files, so it cannot be timed on this box (DESIGN.md sections 6-7). frame_ms_extrapolated = iterations of the whole frame / the measured chaos-game rate + merge + density/tonemap",
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample", "threads", "host_cores", "one_thread_value", "speedup_over_one_thread", "merge_ms", "post_ms")},
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# the GPU arm
# ---------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS), help="BASELINE.json configuration, numbered from 1 (default 2 = configs[1], the 4K still)")
    ap.add_argument("--quality", type=int, default=None, help=argparse.SUPPRESS)  # debugging only; the bench line uses the configuration's
    ap.add_argument("--no-cpu-baseline", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--no-weak", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--watchdog", type=int, default=1500, help=argparse.SUPPRESS)  # seconds until the Python stacks are dumped and the run is abandoned
    args = ap.parse_args()
    import faulthandler
    faulthandler.dump_traceback_later(args.watchdog, exit=True)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank)
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
        # the library's own communicator (C ABI): rank 0 makes the id, torch.distributed only carries it to the others
        box = [r.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        r.comm_init(box[0], rank, world)

    cfg = dict(CONFIGS[args.config])
    if args.quality is not None and cfg["mode"] == "quality":
        cfg["quality"] = args.quality
    OW, OH, ss = cfg["W"], cfg["H"], cfg["ss"]
    W, H = OW * ss, OH * ss
    nbins = W * H

    compiler = r.FlameCompiler(VARIATIONS, overlay=r.OVERLAY_YAML if cfg["genome"] == "stress" else None)
    if cfg["genome"] == "stress":
        flame = r.Flame.load_flame_string(stress_xml({n: compiler.get_parameters_for_variation(n) for n in compiler.variations()}), compiler)
    else:
        flame = r.Flame.load_flame(GENOME, compiler)
    if flame is None:
        raise SystemExit("genome failed to load: " + r.Flame.last_error())
    # rank g seeds particle slots [g*P, (g+1)*P): disjoint JSF32 streams
    r.set_sim_parameters(P, TS, NSHUF, seed=sharding.rank_seed(rank, P))
    post = flame.post_params()
    target = cfg["quality"] * nbins if cfg["mode"] == "quality" else 0
    calls_fixed = cfg.get("calls", 0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def event():
        return torch.cuda.Event(enable_timing=True)

    draw_events, post_events = [], []
    host_rgba8 = torch.empty(OW * OH * 4, dtype=torch.uint8).pin_memory()
    host_view = host_rgba8.numpy().reshape(OH, OW, 4)

    # ---- single GPU, everything resident: value ----
    if cfg["mode"] != "animation":
        bins = torch.zeros(nbins * 4, dtype=torch.float32, device="cuda")
        image = torch.empty(nbins * 4, dtype=torch.float32, device="cuda") if (ss > 1 or world == 1) else None
        rgba8 = torch.empty(OW * OH * 4, dtype=torch.uint8, device="cuda")
        small = torch.empty(OW * OH * 4, dtype=torch.float32, device="cuda") if ss > 1 else None

    def post_step(record):
        if record:
            p0, p1 = event(), event()
            p0.record()
        if ss == 1:
            r.density_tonemap(bins.data_ptr(), None, rgba8.data_ptr(), W, H, post)
        else:
            r.density_tonemap(bins.data_ptr(), image.data_ptr(), None, W, H, post)
            r.spatial_downsample(image.data_ptr(), small.data_ptr(), OW, OH, ss, 1.0)
        if record:
            p1.record()
            post_events.append((p0, p1))

    def resident_step(record, reduce_to_root=False):
        """one frame on this GPU with everything on the device; returns (iterations, binned, draw calls). The binned counter
        (the blocking 8-byte read-back of flame.cpp:329) is read after the first call and then once per batch of calls sized
        to stop just short of the target — the same calls the call-by-call loop makes, as rfk_render_frame does"""
        flame.warmup(WARMUP_PASSES, TSS_WIDTH)
        bins.zero_()
        binned, calls, per_call = 0, 0, 0
        while (target and binned < target) or (calls_fixed and calls < calls_fixed):
            if calls_fixed:
                batch = calls_fixed - calls
            elif per_call:
                batch = max(1, int((target - binned) / (per_call * 1.01)))
            else:
                batch = 1
            launched = []
            for _ in range(batch):
                if record:
                    e0, e1 = event(), event()
                    e0.record()
                flame.draw_to_bins_async(bins.data_ptr(), nbins, W, DRAW_PASSES)
                if record:
                    e1.record()
                    launched.append((e0, e1))
            total = flame.binned_total()
            per_call = (total - binned) / batch
            draw_events.extend((a, b, per_call) for a, b in launched)
            binned = total
            calls += batch
        if reduce_to_root and world > 1:
            r.comm_reduce_histogram(bins.data_ptr(), nbins, 0)  # the one exchange step of the weak-scaling run (ncclReduce)
        if rank == 0 or not reduce_to_root:
            post_step(record)
        return P * (1 + WARMUP_PASSES + DRAW_PASSES * calls), binned, calls

    def e2e_single():
        _, stats = flame.render_frame(OW, OH, target_binned=target, max_draw_calls=calls_fixed, warmup_passes=WARMUP_PASSES, drawing_passes=DRAW_PASSES,
                                      tss_width=TSS_WIDTH, rgba8_out=host_view, supersample=ss)
        return P * (1 + WARMUP_PASSES) + stats.iterations, stats

    def sharded_step():
        """one frame over all ranks through the C ABI, host RGBA8 buffer on rank 0"""
        _, _, st = flame.render_frame_sharded(OW, OH, target_binned=target, max_draw_calls=calls_fixed, warmup_passes=WARMUP_PASSES, drawing_passes=DRAW_PASSES,
                                              tss_width=TSS_WIDTH, supersample=ss, rgba8_out=host_view if rank == 0 else None)
        return st

    def animation_frames(frames, record):
        """BASELINE configs[3]: this rank's frames (f % world == rank), each warmup + one draw call + density/tonemap + read-back"""
        rotated = 0
        iters = 0
        for f in frames:
            while rotated < f:  # frame by frame: the rounding of the single-process animation (src/main.cpp:383-395)
                flame.rotate_xforms(18.0 / 60.0)
                rotated += 1
            _, stats = flame.render_frame(OW, OH, max_draw_calls=1, warmup_passes=WARMUP_PASSES, drawing_passes=DRAW_PASSES, tss_width=TSS_WIDTH, rgba8_out=host_view)
            iters += P * (1 + WARMUP_PASSES) + stats.iterations
        return iters

    detail = {}
    line_extra = {}
    launches0 = None

    if cfg["mode"] == "animation":
        frames = list(range(rank, cfg["frames"], world))
        probe = list(range(rank, min(cfg["frames"], 8 * world), world))
        n_xforms = flame.info().num_xforms
        loaded = [(i, flame.xform(i)) for i in list(range(n_xforms)) + ([-1] if flame.info().has_final_xform else [])]

        def rewind():  # frame 0 again: the affines as loaded (values only: no module is rebuilt)
            for i, x in loaded:
                flame.set_xform(i, x)
        for _ in range(max(1, min(args.warmup, 2))):
            animation_frames(probe, False)
            rewind()
        barrier()
        launches0 = r.kernel_launch_count()
        with ClockSampler(local_rank) as clocks:
            t0, t1 = event(), event()
            t0.record()
            iters = 0
            for _ in range(args.steps):
                iters += animation_frames(frames, True)
                rewind()
            t1.record()
            barrier()
            ms = t0.elapsed_time(t1)
        e2e_ms, e2e_iters = ms, iters  # every frame already goes through the host-buffer call
        detail.update(frames=cfg["frames"], frames_per_rank=len(frames), ms_per_frame_per_gpu=ms / args.steps / max(1, len(frames)))
        scaling = "strong"
        calls = binned = 0
    elif world == 1:
        for _ in range(args.warmup):
            resident_step(False)
        barrier()
        launches0 = r.kernel_launch_count()
        with ClockSampler(local_rank) as clocks:
            t0, t1 = event(), event()
            t0.record()
            iters = binned = calls = 0
            for _ in range(args.steps):
                i, b, c = resident_step(True)
                iters += i; binned += b; calls += c
            t1.record()
            barrier()
            ms = t0.elapsed_time(t1)
        launches_value = r.kernel_launch_count() - launches0
        for _ in range(min(args.warmup, 1)):
            e2e_single()
        barrier()
        e0, e1 = event(), event()
        e0.record()
        e2e_iters = 0
        for _ in range(args.steps):
            i, stats = e2e_single()
            e2e_iters += i
        e1.record()
        barrier()
        e2e_ms = e0.elapsed_time(e1)
        detail.update(draw_calls_per_step=calls / args.steps, in_bounds_fraction=binned / max(1.0, calls * P * DRAW_PASSES), e2e_draw_calls=stats.draw_calls)
        scaling = "strong"
    else:
        # ---- N GPUs, one frame together (strong scaling) ----
        for _ in range(args.warmup):
            sharded_step()
        barrier()
        launches0 = r.kernel_launch_count()
        dev_ms, its = [], 0
        stage = dict(warmup=0.0, draw=0.0, reduce=0.0, post=0.0, readback=0.0)
        with ClockSampler(local_rank) as clocks:
            t0, t1 = event(), event()
            t0.record()
            for _ in range(args.steps):
                st = sharded_step()
                its += st.iterations_global + world * P * (1 + WARMUP_PASSES)
                dev_ms.append(st.ms_warmup + st.ms_draw + st.ms_reduce + st.ms_post)
                for k in stage:
                    stage[k] += getattr(st, "ms_" + k) / args.steps
            t1.record()
            barrier()
            e2e_ms = t0.elapsed_time(t1)
        launches_timed_region = r.kernel_launch_count() - launches0
        ms = sum(dev_ms)  # device time of the frame without the read-back, this rank
        iters = e2e_iters = its / world  # summed over the ranks below
        binned, calls = st.binned_global / world, st.draw_calls
        detail.update(p2p=bool(st.p2p), passes_per_rank=int(st.passes), draw_calls_per_rank=int(st.draw_calls), rows_of_rank0=[int(st.y0), int(st.y1)],
                      stage_ms_rank0=stage, binned_global=int(st.binned_global))
        scaling = "weak" if cfg["mode"] == "calls" else "strong"
        # roofline probe of the draw kernel (outside the timed frames): a few 128-pass calls with events
        flame.warmup(WARMUP_PASSES, TSS_WIDTH)
        bins.zero_()
        before = flame.binned_total()
        for _ in range(8):
            e0, e1 = event(), event()
            e0.record()
            flame.draw_to_bins_async(bins.data_ptr(), nbins, W, DRAW_PASSES)
            e1.record()
            total = flame.binned_total()
            draw_events.append((e0, e1, total - before))
            before = total
        post_step(True) if ss == 1 else None
        # the weak-scaling figure of round 1 beside it: every rank the whole frame, one NCCL reduce onto rank 0
        if not args.no_weak and args.config == 2:
            resident_step(False, reduce_to_root=True)
            barrier()
            w0, w1 = event(), event()
            w0.record()
            wi = 0
            for _ in range(args.steps):
                i, _, _ = resident_step(False, reduce_to_root=True)
                wi += i
            w1.record()
            barrier()
            line_extra["weak"] = {"ms_per_step": w0.elapsed_time(w1) / args.steps, "iterations": wi}

    launches = launches_timed_region if (world > 1 and cfg["mode"] != "animation") else (launches_value if (world == 1 and cfg["mode"] != "animation") else r.kernel_launch_count() - launches0)
    torch.cuda.synchronize()
    draw_ms = [a.elapsed_time(b) for a, b, _ in draw_events]
    draw_binned = [b for _, _, b in draw_events]
    post_ms_list = [a.elapsed_time(b) for a, b in post_events]

    stat = torch.tensor([ms, e2e_ms, float(iters), float(e2e_iters), float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = stat.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stat.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, e2e_ms = float(mx[0]), float(mx[1])
        iters, e2e_iters, launches = float(sm[2]), float(sm[3]), int(sm[4])
        if "weak" in line_extra:
            w = torch.tensor([line_extra["weak"]["ms_per_step"], float(line_extra["weak"]["iterations"])], dtype=torch.float64, device="cuda")
            wm = w.clone(); dist.all_reduce(wm, op=dist.ReduceOp.MAX)
            ws = w.clone(); dist.all_reduce(ws, op=dist.ReduceOp.SUM)
            line_extra["weak"] = {"value": float(ws[1]) / (float(wm[0]) * args.steps * 1e-3), "unit": UNIT, "ms_per_step": float(wm[0]), "scaling": "weak",
                                  "what": "every rank renders the full 2000 spp into a private histogram; one ncclReduce onto rank 0; density estimation + tonemap there"}

    if rank == 0:
        hbm_peak, peak_src, peaks = measured_peaks()
        counters = profile_json("r02_rfk_draw.json") or profile_json("r02a_rfk_draw_specialised.json") or profile_json("r01_rfk_draw.json")
        probes = profile_json("r01_roofline_probes.json")
        ffma_tflops = probes["ffma_tflops"] if probes else 74.4
        line = {
            "metric": METRIC, "value": iters / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": static_config(args.config),
            "frame_ms": ms / args.steps, "frame_ms_e2e": e2e_ms / args.steps,
            "e2e": {"value": e2e_iters / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": (1024 * 4 + 256 * 16) * (cfg.get("frames", 1)),
                    "d2h_bytes_per_step": (OW * OH * 4) * cfg.get("frames", 1) + 8 * 4,
                    "ms_per_step": e2e_ms / args.steps,
                    "api": ("rfk_render_frame_sharded" if world > 1 and cfg["mode"] != "animation" else "rfk_render_frame") + " (C ABI, host RGBA8 buffer, pinned)"},
            "gpu_launches": launches,
            "clocks": clocks.summary(),
            "detail": detail,
        }
        line.update(line_extra)
        if world > 1 and cfg["mode"] != "animation":
            line["frame_ms_strong"] = e2e_ms / args.steps
        if draw_ms:
            mean_draw_ms = sum(draw_ms) / len(draw_ms)
            mean_binned = sum(draw_binned) / len(draw_binned)
            kernel_iters_s = P * DRAW_PASSES / (mean_draw_ms * 1e-3)
            staged = nbins * 16 >= (1 << 29)
            roof = {"kernel": "rfk_draw" + (" (+ stage_accumulate_kernel: region queues)" if staged else ""), "launch_ms": mean_draw_ms, "launches_timed": len(draw_ms),
                    "iterations_per_s_kernel": kernel_iters_s,
                    "share_of_step": (sum(draw_ms) / ms) if world == 1 else None}
            if cfg["genome"] == "shipped" and not staged:
                achieved = FLOP_PER_ITERATION_SHIPPED * kernel_iters_s / 1e12
                roof.update({"bound": "fp32", "achieved": achieved, "peak": ffma_tflops, "unit": "TFLOP/s", "frac": achieved / ffma_tflops,
                             "peak_source": "measured FFMA issue rate x 64 flop (tools/roofline_probes.cu, profiles/r01_roofline_probes.json)" if probes else "nominal 148 SM x 128 lanes x 2 x 1.965 GHz",
                             "flop_per_iteration": FLOP_PER_ITERATION_SHIPPED})
            else:
                roof.update({"bound": "hbm", "achieved": (16.0 * mean_binned + 64.0 * P) / (mean_draw_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                             "frac": (16.0 * mean_binned + 64.0 * P) / (mean_draw_ms * 1e-3) / 1e9 / hbm_peak, "peak_source": peak_src,
                             "note": "no flop census for this genome / the queued path: algorithmic bytes (16 B per binned sample + 64 B of state per particle) over the launch time"})
            roof["traffic"] = counters["dram_bytes_per_launch"] if (counters and cfg["genome"] == "shipped" and args.config == 2) else None
            roof["hbm_algorithmic_frac"] = (16.0 * mean_binned + 64.0 * P) / (mean_draw_ms * 1e-3) / 1e9 / hbm_peak
            if counters and cfg["genome"] == "shipped" and not staged:
                winst = kernel_iters_s / 32.0 * counters["warp_inst_per_unit"]
                roof["warp_inst_per_iteration"] = counters["warp_inst_per_unit"]
                roof["issue_frac"] = winst / (probes["ffma_warp_inst_per_s"] if probes else 1.163e12)
                roof["dram_frac"] = counters["dram_bytes_per_launch"] / (mean_draw_ms * 1e-3) / 1e9 / hbm_peak
                if counters.get("red_sectors_per_launch") and counters.get("red_sectors_pct_of_peak"):
                    peak_red = counters["red_sectors_per_launch"] / (counters["gpu_time_ms"] * 1e-3) / (counters["red_sectors_pct_of_peak"] / 100.0)
                    roof["l2_red_frac"] = mean_binned / (mean_draw_ms * 1e-3) / peak_red
                    roof["l2_red_per_s"] = mean_binned / (mean_draw_ms * 1e-3)
                roof["counters_source"] = counters.get("source")
            if post_ms_list:
                # density estimation + tonemap: 16 B/bin histogram read + 4 B/pixel RGBA8 written (+ 16 B/bin float image when supersampled)
                post_ms = sum(post_ms_list) / len(post_ms_list)
                post_bytes = 16.0 * nbins + 4.0 * OW * OH + (32.0 * nbins + 16.0 * OW * OH if ss > 1 else 0.0)
                roof.update({"post_kernel": "density_tonemap_kernel" + (" + spatial_downsample_kernel" if ss > 1 else ""), "post_bound": "hbm", "post_ms": post_ms,
                             "post_algorithmic_bytes": post_bytes, "post_gbs": post_bytes / (post_ms * 1e-3) / 1e9, "post_peak_gbs": hbm_peak,
                             "post_frac": post_bytes / (post_ms * 1e-3) / 1e9 / hbm_peak, "post_share_of_step": (post_ms * len(post_ms_list) / ms) if world == 1 else None})
            line["roofline"] = roof
        if not args.no_cpu_baseline and world == 1:
            try:
                base = cpu_reference_run(args.config, 2, 1, sample_passes=16)  # one untimed step: the private histograms are allocated and touched there
                line["cpu_baseline"] = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample", "threads", "one_thread_value", "speedup_over_one_thread", "merge_ms", "post_ms")}
            except Exception as e:  # the bench line must still print
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: %s" % e}
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        r.comm_destroy()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
