"""One rank of the multi-GPU tests (tests/test_sharded_gpu.py starts one process per GPU):
  python worker.py <rank> <world> <comm-file> <out-dir>
Writes <out-dir>/rank<r>.npz; the parent compares the ranks' files. NCCL id through a file, as the rfk_render CLI does."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import refrakt_b200 as r
from conftest import GENOME, VARIATIONS

TSS = 1.2 / 60.0


def exchange_id(path, rank):
    if rank == 0:
        uid = r.comm_unique_id()
        with open(path + ".tmp", "wb") as fh:
            fh.write(uid)
        os.rename(path + ".tmp", path)
        return uid
    for _ in range(3000):
        if os.path.exists(path) and os.path.getsize(path) == 128:
            return open(path, "rb").read()
        time.sleep(0.1)
    raise SystemExit("no NCCL id in " + path)


def main():
    rank, world, comm_file, out_dir = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
    assert r.lib().rfk_set_device(rank) == 0, r.Flame.last_error()
    r.comm_init(exchange_id(comm_file, rank), rank, world)
    assert r.comm_rank() == rank and r.comm_world() == world
    out = {}

    compiler = r.FlameCompiler(VARIATIONS)
    flame = r.Flame.load_flame(GENOME, compiler)
    P, TS, W, H = 256 * 16 * 32, 32, 320, 180

    # (a) deterministic mode: the NCCL sum of the per-rank histograms is rank0 + rank1 + ... bit for bit
    flame.set_options(deterministic=1)
    r.set_sim_parameters(P, TS, 64, seed=rank * P)  # rank g seeds slots [g * P, (g + 1) * P): disjoint streams
    out["rng_states"] = r.copy_rng_states(0, 4096)
    flame.warmup(16, TSS)
    bins = r.DeviceBuffer(W * H * 16)
    bins.zero_out()
    out["binned"] = np.array([flame.draw_to_bins(bins.ptr, W * H, W, 32)], dtype=np.int64)
    out["private_bins"] = bins.download(np.float32, (H, W, 4))
    r.comm_reduce_histogram(bins.ptr, W * H, -1)  # onto every rank
    out["allreduced_bins"] = bins.download(np.float32, (H, W, 4))
    bins.zero_out()
    flame.warmup(16, TSS)
    flame.draw_to_bins(bins.ptr, W * H, W, 32)
    out["private_bins_2"] = bins.download(np.float32, (H, W, 4))  # the RNG streams moved on: other samples than the first draw
    r.comm_reduce_histogram(bins.ptr, W * H, 0)  # onto rank 0
    r.comm_barrier()
    if rank == 0:
        out["reduced_bins"] = bins.download(np.float32, (H, W, 4))
    # the same kernels in every frame of (b): the value-specialised, paired build from the first warmup on (with the default
    # `specialize = 2` the first frame of a parameter set runs the generic build and later ones the specialised one, whose
    # xform picks come from other draws of the generators: other samples, and at 40 per pixel a visibly different image)
    flame.set_options(deterministic=0, specialize=1, pair_particles=2)

    # (b) one frame over all ranks, peer-memory path and NCCL path, with and without supersampling
    for tag, env in (("p2p", "1"), ("nccl", "0")):
        os.environ["RFK_COMM_P2P"] = env
        r.release_buffers()  # new buffers: the peers are mapped (or not) again
        r.set_sim_parameters(P, TS, 64, seed=rank * P)
        for ss in (1, 2):
            img8, imgf, st = flame.render_frame_sharded(W, H, target_binned=40 * W * H * ss * ss, drawing_passes=16, want_rgba8=True, want_image=True, supersample=ss)
            out["%s_ss%d_stats" % (tag, ss)] = np.array([st.iterations_global, st.binned_global, st.passes, st.draw_calls, st.p2p, st.y0, st.y1], dtype=np.int64)
            if rank == 0:
                out["%s_ss%d_rgba8" % (tag, ss)] = img8
                out["%s_ss%d_image" % (tag, ss)] = imgf
        # fixed work (no target): every rank exactly max_draw_calls x drawing_passes
        img8, _, st = flame.render_frame_sharded(W, H, max_draw_calls=3, drawing_passes=16)
        out[tag + "_fixed_stats"] = np.array([st.iterations_global, st.binned_global, st.passes, st.draw_calls], dtype=np.int64)
    os.environ.pop("RFK_COMM_P2P", None)

    # (c) frame-parallel animation (BASELINE configs[3]): frame f on rank f % world, seed by frame; deterministic kernels, so
    # the frames must not depend on which rank rendered them
    flame2 = r.Flame.load_flame(GENOME, compiler)
    flame2.set_options(deterministic=1)
    done = 0
    for f in range(4):
        if f % world != rank:
            continue
        while done < f:  # frame by frame, so that the rounding is that of the single-process animation
            flame2.rotate_xforms(18.0 / 60.0)
            done += 1
        r.set_sim_parameters(P, TS, 64, seed=1000 + f)
        img, _ = flame2.render_frame(W, H, max_draw_calls=2, drawing_passes=32)
        out["frame%d" % f] = img
    r.comm_barrier()
    r.comm_destroy()
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), **out)


if __name__ == "__main__":
    main()
