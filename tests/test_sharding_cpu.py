"""The N>1 host path on CPU: two gloo ranks render independent particle streams with the oracle (standing in
for the per-GPU kernels), sum their histograms with the same reduce the bench uses, and split frames."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GENOME, ROOT, VARIATIONS


def _worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import refrakt_oracle as ro
        from refrakt_b200 import sharding

        vt = ro.VariationTable(VARIATIONS)
        orc = ro.Oracle(ro.load_flame(GENOME, vt), vt)
        orc.set_threads(2)
        P, TS, W, H = 256 * 2 * 4, 4, 96, 54
        seed = sharding.rank_seed(rank, P)
        orc.set_sim_parameters(P, TS, 8, shuffle_seed=100 + rank, rng_seed=seed, pass_seed=7 + rank)
        assert np.array_equal(orc.rng_states(0, 1)[0], orc.jsf32_warmup(rank * P))  # disjoint seed ranges
        orc.warmup(8, 1.2 / 60)
        bins = np.zeros((H, W, 4), dtype=np.float32)
        binned = orc.draw_to_bins(bins, W, 16)
        np.save(os.path.join(tmpdir, "bins_%d.npy" % rank), bins)
        t = torch.from_numpy(bins.copy())
        sharding.reduce_histogram(t, dst=0)
        counts = sharding.gather_counts(binned)
        assert len(counts) == world and counts[rank] == binned
        if rank == 0:
            np.save(os.path.join(tmpdir, "reduced.npy"), t.numpy())
            np.save(os.path.join(tmpdir, "counts.npy"), np.array(counts))
        t2 = torch.from_numpy(bins.copy())
        sharding.allreduce_histogram(t2)
        np.save(os.path.join(tmpdir, "allreduced_%d.npy" % rank), t2.numpy())
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_histogram_reduce(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    b0, b1 = np.load(tmp_path / "bins_0.npy"), np.load(tmp_path / "bins_1.npy")
    assert not np.array_equal(b0, b1)  # independent streams
    reduced = np.load(tmp_path / "reduced.npy")
    assert np.array_equal(reduced, b0 + b1)
    assert np.array_equal(np.load(tmp_path / "allreduced_0.npy"), b0 + b1)
    assert np.array_equal(np.load(tmp_path / "allreduced_1.npy"), b0 + b1)
    counts = np.load(tmp_path / "counts.npy")
    assert counts.sum() == int(round(float(reduced[..., 3].sum())))


def test_partition_helpers():
    from refrakt_b200 import sharding
    for world in (1, 2, 4, 8):
        frames = [sharding.frames_of_rank(600, world, r) for r in range(world)]
        assert sorted(sum(frames, [])) == list(range(600))
        assert max(len(f) for f in frames) - min(len(f) for f in frames) <= 1
        shares = [sharding.rank_iteration_share(1000003, world, r) for r in range(world)]
        assert sum(shares) == 1000003 and max(shares) - min(shares) <= 1
        slabs = sharding.row_slabs(2160, world, 11)
        assert slabs[0][0] == 0 and slabs[-1][1] == 2160
        assert all(a[1] == b[0] for a, b in zip(slabs, slabs[1:]))
        assert all(s[2] == max(0, s[0] - 11) and s[3] == min(2160, s[1] + 11) for s in slabs)
        seeds = [sharding.rank_seed(r, 2097152) for r in range(world)]
        assert all(b - a == 2097152 for a, b in zip(seeds, seeds[1:]))
