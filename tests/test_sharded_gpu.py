"""GPU tests of the multi-GPU layer (include/refrakt_b200.h rfk_comm_*, csrc/comm.cpp): row slabs of the density estimation,
the sharded frame on one rank, and — where the box has two GPUs — two processes over NCCL / peer memory."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TSS = 1.2 / 60.0


def _gpu_count():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=20).stdout
        return sum(1 for l in out.splitlines() if l.startswith("GPU "))
    except Exception:
        return 0


@pytest.mark.parametrize("W,H,radius,min_,curve,world", [(200, 120, 11, 0, 0.6, 3), (97, 61, 5, 0, 0.4, 2), (64, 64, 11, 2, 0.6, 5), (130, 70, 0, 0, 0.6, 4),
                                                        (75, 50, 40, 0, 1.0, 8), (90, 33, 9, 0, 0.6, 16)])
def test_density_estimation_over_row_slabs_is_the_whole_image(gpu_ready, rfk, flame, W, H, radius, min_, curve, world):
    """every rank's rows from its slab of the histogram (own rows + estimator-radius halo) are bit for bit the rows of the
    whole-image launch: what makes the sharded frame independent of the number of GPUs"""
    from test_render_gpu import _run_post, _synthetic_bins
    bins = _synthetic_bins(W, H, seed=W * 3 + radius)
    p = flame.post_params()
    p.estimator_radius, p.estimator_min, p.estimator_curve = radius, min_, curve
    _, want, want8 = _run_post(rfk, bins, p, fused=True)
    halo = max(radius, min_)
    got = np.zeros_like(want)
    got8 = np.zeros_like(want8)
    covered = 0
    for rank in range(world):
        s = rfk.comm_row_slab(H, halo, rank, world)
        if s.y1 == s.y0:
            continue
        rows = np.ascontiguousarray(bins[H - s.src_y1: H - s.src_y0])  # source row cy is histogram row H - 1 - cy
        d_rows = rfk.DeviceBuffer(rows.nbytes)
        d_rows.upload(rows)
        n_out = s.y1 - s.y0
        d_out, d_u8 = rfk.DeviceBuffer(n_out * W * 16), rfk.DeviceBuffer(n_out * W * 4)
        rfk.density_tonemap_rows(d_rows.ptr, d_out.ptr, d_u8.ptr, W, H, p, s.y0, s.y1, s.src_y0, s.src_y1, s.y0)
        got[s.y0:s.y1] = d_out.download(np.float32, (n_out, W, 4))
        got8[s.y0:s.y1] = d_u8.download(np.uint8, (n_out, W, 4))
        covered += n_out
        for b in (d_rows, d_out, d_u8):
            b.free()
    assert covered == H
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)) and np.array_equal(got8, want8)
    with pytest.raises(rfk.RefraktError):  # a slab that does not cover the halo is refused
        s = rfk.comm_row_slab(H, 0, 0, 2)
        if halo == 0:
            raise rfk.RefraktError("no halo to miss")
        rfk.density_tonemap_rows(1, 1, None, W, H, p, s.y0, s.y1, s.src_y0, s.src_y1, s.y0)


@pytest.mark.parametrize("radius,min_,curve", [(11, 0, 0.0), (7, 0, -0.3), (3, 9, 0.6), (100, 0, 0.3)])
def test_density_estimation_radius_edge_cases(gpu_ready, rfk, flame, oracle, radius, min_, curve):
    """estimator_curve <= 0 (no radius thresholds: the kernel evaluates the reference's formula), estimator_min above the
    radius, and the maximum radius 100 (main.cpp:502), against the oracle"""
    from test_render_gpu import _run_post, _synthetic_bins
    W, H = (70, 48) if radius < 100 else (150, 110)
    bins = _synthetic_bins(W, H, seed=radius + 5 * min_)
    if radius == 100:
        bins[..., :] = 0
        rng = np.random.default_rng(4)
        for _ in range(40):  # a few isolated samples: radius 100 splats crossing tiles and borders
            y, x = int(rng.integers(0, H)), int(rng.integers(0, W))
            bins[y, x] = [1, 0.5, 0.25, 1]
        bins[H // 2, W // 2] = [3000, 2000, 1000, 5000]
    p = flame.post_params()
    p.estimator_radius, p.estimator_min, p.estimator_curve = radius, min_, curve
    de, out, u8 = _run_post(rfk, bins, p, fused=False)
    want_de = oracle.density_estimate(bins, W, H, radius, min_, curve)
    assert (np.abs(de - want_de) / np.maximum(1.0, np.abs(want_de))).max() <= 1e-4
    want = oracle.tonemap(want_de, scale_constant=p.scale_constant)
    assert np.abs(out - want).max() <= 1e-4
    p.estimator_min = 101
    with pytest.raises(rfk.RefraktError):  # was a host stack overflow (the radius table holds 100)
        _run_post(rfk, bins, p, fused=True)


def test_sharded_frame_on_one_rank_is_the_single_gpu_frame(gpu_ready, rfk, compiler):
    """rfk_render_frame_sharded with a communicator of one rank (NCCL data path) against rfk_render_frame: the same passes
    from the same seeds give the same image up to the order of the floating-point reductions; with a target the stopping
    rule holds without overshooting by a whole call"""
    from conftest import GENOME
    f = rfk.Flame.load_flame(GENOME, compiler)
    W, H, P, TS = 320, 180, 256 * 16 * 32, 32
    rfk.comm_init(rfk.comm_unique_id(), 0, 1)
    try:
        for ss in (1, 2):
            rfk.set_sim_parameters(P, TS, 64, seed=4)
            want, st0 = f.render_frame(W, H, max_draw_calls=3, drawing_passes=32, supersample=ss)
            rfk.set_sim_parameters(P, TS, 64, seed=4)
            got, gotf, st = f.render_frame_sharded(W, H, max_draw_calls=3, drawing_passes=32, supersample=ss, want_image=True)
            assert st.draw_calls == 3 and st.passes == 96 and st.iterations_global == st0.iterations and st.p2p == 0
            assert (st.y0, st.y1) == (0, H)
            assert abs(int(st.binned_global) - int(st0.binned)) <= 1e-4 * st0.binned
            diff = np.abs(got.astype(int) - want.astype(int))
            assert diff.max() <= 2 and (diff > 0).mean() < 0.02, (diff.max(), (diff > 0).mean())
            assert np.abs(np.rint(np.clip(gotf, 0, 1) * 255) - got).max() <= 1
        target = 60 * W * H
        _, _, st = f.render_frame_sharded(W, H, target_binned=target, drawing_passes=16)
        per_pass = st.binned_global / st.passes
        assert target <= st.binned_global <= target + 2.5 * per_pass, (st.binned_global, target, per_pass)
        with pytest.raises(rfk.RefraktError):
            f.render_frame_sharded(W, H)  # neither target nor call count
    finally:
        rfk.comm_destroy()
    assert rfk.comm_world() == 1
    with pytest.raises(rfk.RefraktError):
        f.render_frame_sharded(W, H, max_draw_calls=1)  # no communicator


@pytest.mark.skipif(_gpu_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_two_ranks_over_nccl_and_peer_memory(gpu_ready, rfk, compiler, tmp_path):
    """world size 2, one process per GPU: (a) in deterministic mode the reduced histogram is rank0 + rank1 bit for bit and the
    ranks' RNG streams are disjoint; (b) the sharded frame over peer memory, over NCCL and on one GPU agree; (c) frames of a
    frame-parallel animation do not depend on the rank that rendered them"""
    from conftest import GENOME
    world = 2
    comm_file = str(tmp_path / "nccl_id")
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mgpu", "worker.py")
    procs = [subprocess.Popen([sys.executable, worker, str(rank), str(world), comm_file, str(tmp_path)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for rank in range(world)]
    logs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=600)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        logs.append(out)
    assert all(p.returncode == 0 for p in procs), "\n".join(logs)
    r0, r1 = (np.load(str(tmp_path / ("rank%d.npz" % k))) for k in range(world))

    # (a)
    s0 = set(map(bytes, r0["rng_states"])); s1 = set(map(bytes, r1["rng_states"]))
    assert len(s0) == 4096 and not (s0 & s1)
    assert not np.array_equal(r0["private_bins"], r1["private_bins"])
    want = r0["private_bins"] + r1["private_bins"]  # binary32 addition, as NCCL's sum of two ranks
    assert np.array_equal(r0["allreduced_bins"].view(np.uint32), want.view(np.uint32))
    assert np.array_equal(r1["allreduced_bins"].view(np.uint32), want.view(np.uint32))
    want2 = r0["private_bins_2"] + r1["private_bins_2"]
    assert np.array_equal(r0["reduced_bins"].view(np.uint32), want2.view(np.uint32)) and not np.array_equal(want2, want)
    assert abs(float(want[..., 3].sum()) - float(r0["binned"][0] + r1["binned"][0])) <= 0.5

    # (b) the two data paths drew the same passes from the same seeds: same image up to reduction order
    W, H = 320, 180
    for ss in (1, 2):
        sp, sn = r0["p2p_ss%d_stats" % ss], r0["nccl_ss%d_stats" % ss]
        assert sp[4] == 1 and sn[4] == 0, "peer-memory path not taken: %s" % sp
        assert np.array_equal(sp[:4], r1["p2p_ss%d_stats" % ss][:4]) and sp[1] >= 40 * W * H * ss * ss
        assert (sp[5], sp[6]) == (0, H // 2) and tuple(r1["p2p_ss%d_stats" % ss][5:7]) == (H // 2, H)
        a, b = r0["p2p_ss%d_rgba8" % ss].astype(int), r0["nccl_ss%d_rgba8" % ss].astype(int)
        assert a.shape == (H, W, 4) and a[..., :3].max() > 100
        assert a[H // 2:].max() > 0  # rank 1's rows arrived
        if sp[2] == sn[2]:  # same number of passes (the in-bounds estimate is the same): the same samples
            assert np.abs(a - b).max() <= 2 and (np.abs(a - b) > 0).mean() < 0.02
        f8 = np.rint(np.clip(r0["p2p_ss%d_image" % ss], 0, 1) * 255)
        assert np.abs(f8 - a).max() <= 1
    assert np.array_equal(r0["p2p_fixed_stats"], r0["nccl_fixed_stats"]) and r0["p2p_fixed_stats"][2] == 48 and r0["p2p_fixed_stats"][3] == 3

    # the sharded frame against one GPU running density estimation on the summed histogram of the same two streams
    f = rfk.Flame.load_flame(GENOME, compiler)
    P, TS = 256 * 16 * 32, 32

    # (c) the same frames from one rank
    f.set_options(deterministic=1)
    for k in range(4):
        if k:
            f.rotate_xforms(18.0 / 60.0)
        rfk.set_sim_parameters(P, TS, 64, seed=1000 + k)
        img, _ = f.render_frame(W, H, max_draw_calls=2, drawing_passes=32)
        theirs = (r0 if k % 2 == 0 else r1)["frame%d" % k]
        assert np.array_equal(img, theirs), k
