"""The N > 1 host path on CPU, world size 2 over gloo: the oracle stands in for the per-GPU kernels and gloo for NCCL / peer
memory; the row-slab geometry is the C ABI's (rfk_comm_row_slab). What is checked is the LOGIC of rfk_render_frame_sharded:
disjoint seed ranges, the reduce-scatter over row slabs with an estimator-radius halo, density estimation + tonemap on every
rank's rows only, the gather on rank 0 — against the single-process frame of the summed histogram."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GENOME, ROOT, VARIATIONS

W, H, RADIUS = 96, 54, 11


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
        P, TS = 256 * 2 * 4, 4
        seed = sharding.rank_seed(rank, P)
        orc.set_sim_parameters(P, TS, 8, shuffle_seed=100 + rank, rng_seed=seed, pass_seed=7 + rank)
        assert np.array_equal(orc.rng_states(0, 1)[0], orc.jsf32_warmup(rank * P))  # disjoint seed ranges
        orc.warmup(8, 1.2 / 60)
        bins = np.zeros((H, W, 4), dtype=np.float32)
        binned = orc.draw_to_bins(bins, W, 16)
        np.save(os.path.join(tmpdir, "bins_%d.npy" % rank), bins)

        # the counters: one all-reduce (the barrier in front of the exchange)
        count = torch.tensor([binned], dtype=torch.int64)
        dist.all_reduce(count)

        # reduce-scatter over row slabs with halo: the source rows [src_y0, src_y1) of rank r are the histogram rows
        # [H - src_y1, H - src_y0); one reduce per destination rank, as comm::reduce_scatter_slabs_nccl
        slabs = [sharding.row_slab(H, RADIUS, r, world) for r in range(world)]
        mine = None
        for r, (y0, y1, s0, s1) in enumerate(slabs):
            part = torch.from_numpy(bins[H - s1:H - s0].copy())
            dist.reduce(part, dst=r)
            if r == rank:
                mine = part.numpy()
        y0, y1, s0, s1 = slabs[rank]
        # density estimation + tonemap on this rank's rows: rows outside the slab count as empty
        padded = np.zeros((H, W, 4), dtype=np.float32)
        padded[H - s1:H - s0] = mine
        rows = ro.to_rgba8(orc.tonemap(orc.density_estimate(padded, W, H)))[y0:y1]
        # gather on rank 0
        out = [torch.zeros((sl[1] - sl[0], W, 4), dtype=torch.uint8) for sl in slabs] if rank == 0 else None
        dist.gather(torch.from_numpy(np.ascontiguousarray(rows)), out, dst=0)
        if rank == 0:
            np.save(os.path.join(tmpdir, "image.npy"), np.concatenate([t.numpy() for t in out], axis=0))
            np.save(os.path.join(tmpdir, "count.npy"), count.numpy())
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_frame_logic(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import refrakt_oracle as ro
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    b0, b1 = np.load(tmp_path / "bins_0.npy"), np.load(tmp_path / "bins_1.npy")
    assert not np.array_equal(b0, b1)  # independent streams
    total = b0 + b1
    assert int(np.load(tmp_path / "count.npy")[0]) == int(round(float(total[..., 3].sum())))
    vt = ro.VariationTable(VARIATIONS)
    orc = ro.Oracle(ro.load_flame(GENOME, vt), vt)
    want = ro.to_rgba8(orc.tonemap(orc.density_estimate(total, W, H)))
    got = np.load(tmp_path / "image.npy")
    assert got.shape == want.shape and np.array_equal(got, want)  # slabs + halo lose nothing


def test_partition_rules(rfk):
    from refrakt_b200 import sharding
    for world in (1, 2, 3, 4, 8, 16):
        frames = [sharding.frames_of_rank(600, world, r) for r in range(world)]
        assert sorted(sum(frames, [])) == list(range(600))
        assert max(len(f) for f in frames) - min(len(f) for f in frames) <= 1
        for height, halo in ((2160, 11), (8640, 100), (5, 2), (720, 0)):
            slabs = [sharding.row_slab(height, halo, r, world) for r in range(world)]
            assert slabs[0][0] == 0 and slabs[-1][1] == height
            assert all(a[1] == b[0] for a, b in zip(slabs, slabs[1:]))
            sizes = [s[1] - s[0] for s in slabs]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
            for y0, y1, s0, s1 in slabs:
                assert (s0, s1) == ((max(0, y0 - halo), min(height, y1 + halo)) if y1 > y0 else (y0, y0))
        seeds = [sharding.rank_seed(r, 2097152) for r in range(world)]
        assert all(b - a == 2097152 for a, b in zip(seeds, seeds[1:]))
    with pytest.raises(rfk.RefraktError):
        rfk.comm_row_slab(100, 1, 2, 2)  # rank outside the world
    # the communicator entry points fail loudly without a communicator (and without a GPU)
    assert rfk.comm_world() == 1 and rfk.comm_rank() == 0 and not rfk.comm_p2p()
    with pytest.raises(rfk.RefraktError):
        rfk.comm_barrier()
    try:
        assert len(rfk.comm_unique_id()) == 128  # NCCL is loaded at run time; making an id needs no GPU
    except rfk.RefraktError as e:
        assert "libnccl" in str(e)  # a box without NCCL: loud, not silent
