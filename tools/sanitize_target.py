"""Small exercise of every kernel of the library, meant to be run under compute-sanitizer
(tools/sanitize.sh): warm-up, draw in every kernel mode, density estimation + tonemap, supersampled
frame, reference pass mode, seeding kernels. Sizes are tiny: the sanitizer tools slow kernels 10-100x."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import refrakt_b200 as r  # noqa: E402

FIX = os.path.join(ROOT, "refrakt_b200", "data")
compiler = r.FlameCompiler(os.path.join(FIX, "variations.yaml"))
flame = r.Flame.load_flame(os.path.join(FIX, "electricsheep.247.11256.flam3"), compiler)
assert flame is not None, r.Flame.last_error()

P, TS = 256 * 32, 16
r.set_sim_parameters(P, TS, 16, seed=3)
W, H = 96, 54
modes = [dict(specialize=0), dict(specialize=1, pair_particles=0), dict(specialize=1, pair_particles=2), dict(specialize=1, pair_particles=1),
         dict(warp_aggregate=1), dict(per_lane_xform=1), dict(deterministic=1), dict(count_xforms=1),
         dict(block_width=128), dict(block_width=512), dict(deal_period=4), dict(math_mode=0), dict(math_mode=2)]
base = flame.options()
defaults = {k: getattr(base, k) for k, _ in base._fields_}
for kw in modes:
    flame.set_options(**defaults)
    flame.set_options(**kw)
    flame.warmup(3, 1.2 / 60)
    img, stats = flame.render_frame(W, H, max_draw_calls=1, drawing_passes=4, warmup_passes=3)
    assert stats.binned > 0, kw
    print("mode", kw, "binned", stats.binned)
flame.set_options(**defaults)

# L2 hints: hot map built from a first call, then a hinted call
flame.set_options(l2_hints=1)
flame.warmup(3, 1.2 / 60)
hb = r.DeviceBuffer(W * H * 16)
hb.zero_out()
flame.draw_to_bins(hb.ptr, W * H, W, 2)
info = flame.build_hot_map(hb.ptr, W * H, W, 64 * 1024)
print("hot map", info.tiles_x, info.tiles_y, info.hot_tiles, "binned with hints", flame.draw_to_bins(hb.ptr, W * H, W, 2), int(flame.hot_map().sum()))
flame.clear_hot_map()
flame.set_options(**defaults)

# region queues (kernel option staged_bins): plenty of room, exhausted queues (one chunk per region), a wider CTA, a longer
# re-deal period; every case is checked against the direct path's binned count being positive and the density sum matching
for kw, max_bytes in ((dict(staged_bins=8), None), (dict(staged_bins=8), 21 * 4096), (dict(staged_bins=9, block_width=512), None),
                      (dict(staged_bins=10, deal_period=4), None), (dict(staged_bins=8, count_xforms=1), None)):
    if max_bytes:
        os.environ["RFK_STAGE_MAX_BYTES"] = str(max_bytes)
    else:
        os.environ.pop("RFK_STAGE_MAX_BYTES", None)
    flame.set_options(**defaults)
    flame.set_options(**kw)
    flame.warmup(3, 1.2 / 60)
    sb = r.DeviceBuffer(W * H * 16)
    sb.zero_out()
    n = flame.draw_to_bins(sb.ptr, W * H, W, 5) + flame.draw_to_bins(sb.ptr, W * H, W, 2)
    dens = sb.download(np.float32, (H, W, 4))[..., 3].astype(np.float64).sum()
    assert n > 0 and dens == n, (kw, n, dens)
    print("staged", kw, "queue bytes", max_bytes, "binned", n)
    sb.free()
os.environ.pop("RFK_STAGE_MAX_BYTES", None)
flame.set_options(**defaults)

# region queues inside a frame (rfk_render_frame, several draw calls)
for passes in (4, 40):
    flame.set_options(staged_bins=8)
    img, stats = flame.render_frame(W, H, max_draw_calls=3, drawing_passes=passes, warmup_passes=3)
    print("queued frame: passes", passes, "binned", stats.binned)
flame.set_options(**defaults)

img, stats = flame.render_frame(W, H, max_draw_calls=1, drawing_passes=4, warmup_passes=3, supersample=2, filter_radius=0.75)
print("supersampled frame", img.shape, stats.binned)

r.set_shuffle_buffers(count=8, seed=5)
flame.reference_warmup(3, 1.2 / 60, shuffle_ids=np.arange(8, dtype=np.uint32) % 8)
bins = r.DeviceBuffer(W * H * 16)
bins.zero_out()
n = flame.reference_draw_to_bins(bins.ptr, W * H, W, 3, shuffle_ids=np.arange(6, dtype=np.uint32) % 8)
print("reference pass mode binned", n)

# round 2: density estimation over row slabs (and a large radius), the sharded frame on a one-rank communicator (NCCL data path)
bins_h = np.zeros((H, W, 4), dtype=np.float32)
rng = np.random.default_rng(1)
bins_h[..., 3] = (rng.random((H, W)) < 0.2) * rng.integers(1, 60, (H, W))
bins_h[..., :3] = rng.random((H, W, 3)) * bins_h[..., 3:4]
post = flame.post_params()
for radius, world in ((11, 3), (40, 2), (0, 2)):
    post.estimator_radius = radius
    for rank in range(world):
        sl = r.comm_row_slab(H, radius, rank, world)
        rows = np.ascontiguousarray(bins_h[H - sl.src_y1: H - sl.src_y0])
        d_rows, d_out = r.DeviceBuffer(rows.nbytes), r.DeviceBuffer((sl.y1 - sl.y0) * W * 16)
        d_rows.upload(rows)
        r.density_tonemap_rows(d_rows.ptr, d_out.ptr, None, W, H, post, sl.y0, sl.y1, sl.src_y0, sl.src_y1, sl.y0)
        d_out.download(np.float32, (sl.y1 - sl.y0, W, 4))
        d_rows.free(); d_out.free()
    print("density over row slabs: radius", radius, "ranks", world)
try:
    r.comm_init(r.comm_unique_id(), 0, 1)
    for ss in (1, 2):
        img8, imgf, st = flame.render_frame_sharded(W, H, max_draw_calls=2, drawing_passes=4, warmup_passes=3, supersample=ss, want_image=True)
        print("sharded frame on one rank: supersample", ss, "binned", st.binned_global, "p2p", st.p2p)
    r.comm_destroy()
except r.RefraktError as e:
    print("sharded frame skipped:", e)

print("seed", r.seed_rng_states(1024, 0)[-1], "hammersley", r.make_sample_points(512)[-1], "shuffle", r.make_shuffle_buffers(512, 2, 1)[0, :4])
print("launches", r.kernel_launch_count())
