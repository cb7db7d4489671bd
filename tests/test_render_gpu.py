"""GPU tests of the render path: density estimation + tonemap against the oracle on the same
histogram, statistical parity of the chaos-game histogram, deterministic mode, the end-to-end frame."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TSS = 1.2 / 60.0


def _synthetic_bins(W, H, seed, dense_frac=0.3, sparse_frac=0.3):
    """float4 histogram with dense blobs (radius 0), sparse speckle (large radii) and empty areas"""
    rng = np.random.default_rng(seed)
    bins = np.zeros((H, W, 4), dtype=np.float32)
    yy, xx = np.mgrid[0:H, 0:W]
    blob = np.exp(-(((xx - W * 0.4) / (W * 0.15)) ** 2 + ((yy - H * 0.5) / (H * 0.2)) ** 2))
    dens = np.floor(blob * 3000 * dense_frac * rng.random((H, W))).astype(np.float32)
    speck = (rng.random((H, W)) < sparse_frac * 0.2).astype(np.float32) * rng.integers(1, 40, (H, W))
    d = dens + speck
    d[:, : W // 8] = 0
    col = rng.random((H, W, 3)).astype(np.float32)
    bins[..., :3] = col * d[..., None]
    bins[..., 3] = d
    # corners and edges populated: splats must clip at the image border
    bins[0, 0] = [1, 2, 3, 1]
    bins[H - 1, W - 1] = [2, 1, 1, 2]
    bins[0, W - 1] = [5, 5, 5, 60]
    return bins


def _run_post(rfk, bins, p, fused=True):
    H, W = bins.shape[:2]
    d_bins = rfk.DeviceBuffer(bins.nbytes)
    d_bins.upload(bins)
    d_img = rfk.DeviceBuffer(bins.nbytes)
    d_out = rfk.DeviceBuffer(bins.nbytes)
    d_u8 = rfk.DeviceBuffer(W * H * 4)
    if fused:
        rfk.density_tonemap(d_bins.ptr, d_out.ptr, d_u8.ptr, W, H, p)
        de = None
    else:
        rfk.density_estimate(d_bins.ptr, d_img.ptr, W, H, p)
        rfk.tonemap(d_img.ptr, d_out.ptr, d_u8.ptr, W, H, p)
        de = d_img.download(np.float32, (H, W, 4))
    out = d_out.download(np.float32, (H, W, 4))
    u8 = d_u8.download(np.uint8, (H, W, 4))
    for b in (d_bins, d_img, d_out, d_u8):
        b.free()
    return de, out, u8


@pytest.mark.parametrize("W,H,radius,min_,curve", [(200, 120, 11, 0, 0.6), (97, 61, 5, 0, 0.4), (64, 64, 11, 2, 0.6), (130, 70, 0, 0, 0.6), (75, 50, 20, 0, 1.0)])
def test_density_tonemap_matches_oracle(gpu_ready, rfk, flame, oracle, oracle_mod, W, H, radius, min_, curve):
    """same input histogram through the oracle (scatter form) and the GPU (gather form):
    float image within 1e-4 (relative to the pixel's own scale), 8-bit within 1 LSB"""
    bins = _synthetic_bins(W, H, seed=W * 7 + radius)
    p = flame.post_params()
    p.estimator_radius, p.estimator_min, p.estimator_curve = radius, min_, curve
    de, out, u8 = _run_post(rfk, bins, p, fused=False)
    want_de = oracle.density_estimate(bins, W, H, radius, min_, curve)
    scale = np.maximum(1.0, np.abs(want_de))
    assert (np.abs(de - want_de) / scale).max() <= 1e-4
    want = oracle.tonemap(want_de, scale_constant=p.scale_constant)
    assert np.abs(out - want).max() <= 1e-4
    assert np.abs(u8.astype(int) - oracle_mod.to_rgba8(want).astype(int)).max() <= 1
    _, fused_out, fused_u8 = _run_post(rfk, bins, p, fused=True)
    assert np.array_equal(fused_out, out) and np.array_equal(fused_u8, u8)
    assert (out[..., 3] == 1.0).all()
    empty = want_de[..., 3] == 0
    assert (out[empty][:, :3] == 0).all()


def test_density_shift_and_gain(gpu_ready, rfk, flame):
    """a single bin: radius-0 copy lands one pixel to the left; a radius-1 splat
    has gain ~2.33 (item 11)"""
    W, H = 33, 17
    p = flame.post_params()
    bins = np.zeros((H, W, 4), dtype=np.float32)
    bins[H - 1 - 5, 10] = [100, 200, 300, 1000]  # histogram row by = H-1-cy
    de, _, _ = _run_post(rfk, bins, p, fused=False)
    assert np.array_equal(de[5, 9], bins[H - 1 - 5, 10]) and np.count_nonzero(de[..., 3]) == 1
    p.estimator_radius, p.estimator_curve = 1, 0.0
    bins[...] = 0
    bins[H - 1 - 8, 20] = [1, 1, 1, 1]
    de, _, _ = _run_post(rfk, bins, p, fused=False)
    assert abs(de[..., 3].sum() - 2.326) < 0.01
    ys, xs = np.nonzero(de[..., 3])
    assert xs.min() >= 18 and xs.max() <= 20 and ys.min() >= 7 and ys.max() <= 9


def _pooled_density(bins, k=4):
    H, W = bins.shape[:2]
    d = bins[: H // k * k, : W // k * k, 3].astype(np.float64)
    return d.reshape(H // k, k, W // k, k).sum(axis=(1, 3))


def _norm_l1(a, b):
    return 0.5 * np.abs(a / a.sum() - b / b.sum()).sum()


def _gpu_histogram(rfk, flame, W, H, P, TS, passes, calls, seed, **options):
    flame.set_options(math_mode=1, fmad=1, per_lane_xform=0, warp_aggregate=0, deterministic=0, count_xforms=1, min_blocks=0, block_width=256, deal_period=1)
    if options:
        flame.set_options(**options)
    rfk.set_sim_parameters(P, TS, 64, seed=seed)
    flame.warmup(16, TSS)
    buf = rfk.DeviceBuffer(W * H * 16)
    buf.zero_out()
    binned = 0
    for _ in range(calls):
        binned += flame.draw_to_bins(buf.ptr, W * H, W, passes)
    bins = buf.download(np.float32, (H, W, 4))
    buf.free()
    return bins, binned


@pytest.fixture(scope="module")
def oracle_hist(oracle):
    """two independent oracle runs of the same small configuration"""
    W, H, P, TS = 320, 180, 256 * 16 * 32, 32
    out = []
    for rng_seed, shuf, pas in ((0, 0x5EED0000, 0x5EED0001), (P, 0x1234, 0x99)):
        oracle.set_sim_parameters(P, TS, 64, shuffle_seed=shuf, rng_seed=rng_seed, pass_seed=pas)
        oracle.warmup(16, TSS)
        bins = np.zeros((H, W, 4), dtype=np.float32)
        binned = oracle.draw_to_bins(bins, W, 64, count_xforms=True)
        out.append((bins, binned, oracle.xform_picks(10).astype(np.float64)))
    return (W, H, P, TS, 64), out


@pytest.mark.parametrize("mode", [dict(), dict(warp_aggregate=1), dict(per_lane_xform=1), dict(deterministic=1)])
def test_histogram_statistical_parity(gpu_ready, rfk, flame, oracle_hist, mode):
    """identical iteration counts, independent random streams (BASELINE.md §5 'histogram' row)"""
    (W, H, P, TS, passes), ((b1, n1, p1), (b2, n2, p2)) = oracle_hist
    bins, binned = _gpu_histogram(rfk, flame, W, H, P, TS, passes, 1, seed=12345, **mode)
    total = P * passes
    self_l1 = _norm_l1(_pooled_density(b1), _pooled_density(b2))
    l1 = _norm_l1(_pooled_density(bins), _pooled_density(b1))
    assert l1 <= max(0.02, 1.5 * self_l1), (l1, self_l1)
    assert abs(binned / total - n1 / total) <= 0.002 * (n1 / total) + 3e-4, (binned / total, n1 / total, n2 / total)
    assert abs(bins[..., 3].sum() - binned) <= 1e-3 * binned  # opacity 1: density channel counts samples
    picks = flame.xform_counts(10).astype(np.float64)
    assert picks.sum() == total
    weights = flame.copy_flame_data_to_buffer()[[0, 13, 30, 43, 59, 76, 92, 110, 125, 141]]
    draws = total if mode.get("per_lane_xform") else total / 32  # independent selections
    assert np.abs(picks / picks.sum() - weights).max() <= max(1e-3, 4.5 * np.sqrt(0.25 / draws))
    # colour: mean rgb per unit density agrees, over the image and per region of an 8 x 8 grid (where the region holds mass)
    c_gpu = bins[..., :3].sum(axis=(0, 1)) / bins[..., 3].sum()
    c_ref = b1[..., :3].sum(axis=(0, 1)) / b1[..., 3].sum()
    assert np.abs(c_gpu - c_ref).max() <= 0.01

    def regions(b, k=8):
        g = b[: H // k * k, : W // k * k].astype(np.float64).reshape(k, H // k, k, W // k, 4).sum(axis=(1, 3))
        return g[..., :3] / np.maximum(g[..., 3:4], 1e-9), g[..., 3]
    rg, _ = regions(bins)
    r1, m1 = regions(b1)
    r2, _ = regions(b2)
    heavy = m1 > 0.004 * m1.sum()
    assert heavy.sum() >= 8
    assert np.abs(rg - r1)[heavy].max() <= max(0.015, 2.0 * np.abs(r2 - r1)[heavy].max()), (np.abs(rg - r1)[heavy].max(), np.abs(r2 - r1)[heavy].max())


def test_deterministic_mode_is_bit_identical(gpu_ready, rfk, flame):
    W, H, P, TS = 256, 144, 256 * 8 * 16, 16
    a, na = _gpu_histogram(rfk, flame, W, H, P, TS, 48, 2, seed=9, deterministic=1)
    b, nb = _gpu_histogram(rfk, flame, W, H, P, TS, 48, 2, seed=9, deterministic=1)
    assert na == nb and np.array_equal(a, b)
    c, nc = _gpu_histogram(rfk, flame, W, H, P, TS, 48, 2, seed=10, deterministic=1)
    assert not np.array_equal(a, c)
    # float accumulation of the same samples differs only by rounding order
    d, nd = _gpu_histogram(rfk, flame, W, H, P, TS, 48, 2, seed=9, deterministic=0)
    assert nd == na
    assert np.abs(d[..., 3] - a[..., 3]).max() <= 1e-3 * max(1.0, a[..., 3].max())


def test_draw_requires_warmup_and_tracks_state(gpu_ready, rfk, compiler):
    from conftest import GENOME
    f = rfk.Flame.load_flame(GENOME, compiler)
    rfk.set_sim_parameters(256 * 4, 4, 8)
    assert f.needs_warmup()
    buf = rfk.DeviceBuffer(64 * 36 * 16)
    buf.zero_out()
    with pytest.raises(rfk.RefraktError):
        f.draw_to_bins(buf.ptr, 64 * 36, 64, 4)
    f.warmup(4, TSS)
    assert not f.needs_warmup()
    n = f.draw_to_bins(buf.ptr, 64 * 36, 64, 8)
    assert 0 < n <= 256 * 4 * 8 and f.binned_total() == n
    x = f.xform(0)
    x.weight *= 2
    f.set_xform(0, x)
    assert f.needs_warmup()  # the UI forces warmup after an edit (main.cpp:397-409)
    f.warmup(4, TSS)
    rfk.set_sim_parameters(256 * 8, 8, 8)  # invalidates live flames (flame.cpp:153-157)
    assert f.needs_warmup()
    with pytest.raises(rfk.RefraktError):
        rfk.set_sim_parameters(1000, 3, 8)  # particles per temporal sample not a multiple of 256
    buf.free()


def test_render_frame_end_to_end(gpu_ready, rfk, flame, oracle, oracle_mod):
    """host-buffer frame: image statistically matches the oracle's frame of the same sample count"""
    W, H, P, TS = 320, 180, 256 * 16 * 32, 32
    flame.set_options(math_mode=1, fmad=1, per_lane_xform=0, warp_aggregate=0, deterministic=0, count_xforms=0, min_blocks=0, block_width=256, deal_period=1)
    rfk.set_sim_parameters(P, TS, 64, seed=777)
    img = np.empty((H, W, 4), dtype=np.uint8)
    fimg = np.empty((H, W, 4), dtype=np.float32)
    _, stats = flame.render_frame(W, H, max_draw_calls=4, drawing_passes=64, rgba8_out=img, image_out=fimg)
    assert stats.draw_calls == 4 and stats.iterations == 4 * 64 * P and 0 < stats.binned <= stats.iterations
    assert np.array_equal(img, oracle_mod.to_rgba8(fimg))

    def oracle_frame(rng_seed, shuf):
        oracle.set_sim_parameters(P, TS, 64, shuffle_seed=shuf, rng_seed=rng_seed)
        oracle.warmup(16, TSS)
        bins = np.zeros((H, W, 4), dtype=np.float32)
        for _ in range(4):
            oracle.draw_to_bins(bins, W, 64)
        return oracle_mod.to_rgba8(oracle.tonemap(oracle.density_estimate(bins, W, H)))

    ref1, ref2 = oracle_frame(0, 1), oracle_frame(P, 2)

    def psnr(a, b):
        mse = np.mean((a[..., :3].astype(np.float64) - b[..., :3].astype(np.float64)) ** 2)
        return 10 * np.log10(255.0**2 / mse)

    self_psnr = psnr(ref1, ref2)
    got = psnr(img, ref1)
    assert got >= min(30.0, self_psnr - 1.0), (got, self_psnr)
    mae = np.mean(np.abs(img[..., :3].astype(np.float64) - ref1[..., :3].astype(np.float64)))
    self_mae = np.mean(np.abs(ref2[..., :3].astype(np.float64) - ref1[..., :3].astype(np.float64)))
    assert mae <= max(2.0, 1.25 * self_mae), (mae, self_mae)
    # the target-binned stopping rule of main.cpp:411
    _, stats2 = flame.render_frame(W, H, target_binned=3 * W * H, drawing_passes=8)
    assert stats2.binned >= 3 * W * H and stats2.binned - 3 * W * H < 8 * P


def test_cli_renders_animation_frames(gpu_ready, rfk, tmp_path):
    """rfk_render: headless stills / animation to PNG through the C ABI"""
    import json
    import os
    import subprocess
    from PIL import Image
    from conftest import GENOME, VARIATIONS
    cli = os.path.join(os.path.dirname(rfk.LIB_PATH), "rfk_render")
    out = str(tmp_path / "frame_%03d.png")
    r = subprocess.run([cli, "--genome", GENOME, "--variations", VARIATIONS, "--out", out, "--width", "320", "--height", "180", "--quality", "50",
                        "--particles", str(256 * 16 * 32), "--temporal-samples", "32", "--passes", "32", "--frames", "3", "--fps", "2"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr
    lines = [json.loads(l) for l in r.stdout.strip().split("\n")]
    assert len(lines) == 3 and all(l["binned"] >= 50 * 320 * 180 for l in lines)
    imgs = [np.array(Image.open(out % k)) for k in range(3)]
    assert all(i.shape == (180, 320, 4) and i[..., :3].max() > 100 and (i[..., 3] == 255).all() for i in imgs)
    assert np.abs(imgs[0].astype(int) - imgs[2].astype(int)).mean() > 0.5  # the xforms rotated between frames


def _linear_genome(n, final=False):
    rng = np.random.default_rng(n)
    xf = []
    for i in range(n):
        coefs = " ".join("%.4f" % v for v in (rng.normal(0, 0.45, 6)))
        xf.append('<xform weight="%.3f" color="%.3f" color_speed="0.5" animate="1" linear="1" spherical="%.2f" coefs="%s" opacity="1"/>' % (rng.uniform(0.1, 1), rng.random(), 0.05 * (i % 3), coefs))
    if final:
        xf.append('<finalxform color="0.5" color_speed="0.2" linear="1" coefs="0.9 0 0 0.9 0.05 0" opacity="1"/>')
    from conftest import GENOME_TEMPLATE
    return GENOME_TEMPLATE % "\n".join(xf)


@pytest.mark.parametrize("n,final", [(1, False), (2, True), (33, False), (34, True), (40, False)])
def test_xform_count_edge_cases(gpu_ready, rfk, compiler, vt, oracle_mod, n, final):
    """1 xform (the select template has no fall-through return), exactly 33, 34 and 40 xforms (the selection if-chain grows with the count)
    (if-chain fallback), with and without a final xform: selection bit-exact, single step within 1e-5, histogram mass conserved"""
    xml = _linear_genome(n, final)
    f = rfk.Flame.load_flame_string(xml, compiler)
    assert f is not None, rfk.Flame.last_error()
    of = oracle_mod.load_flame_string(xml, vt)
    orc = oracle_mod.Oracle(of, vt)
    ratio = np.concatenate([np.random.default_rng(1).random(50000).astype(np.float32), [0.0, 1.0]])
    got = f.select_xform(ratio)
    assert np.array_equal(got, orc.select_xform(ratio)) and got.min() >= 0 and got.max() == n - 1
    rng = np.random.default_rng(2)
    m = 20000
    xyz = np.concatenate([rng.normal(0, 1, (m, 2)), rng.random((m, 1))], axis=1).astype(np.float32)
    ids = rng.integers(-1 if final else 0, n, m).astype(np.int32)
    states = rng.integers(0, 2**32, (m, 4), dtype=np.uint64).astype(np.uint32)
    a, ra = f.single_step(xyz, ids, states)
    b, rb = orc.single_step(xyz, ids, states)
    assert np.array_equal(ra, rb)
    ok = np.isfinite(b).all(axis=1)
    err = np.linalg.norm(a[ok, :2] - b[ok, :2], axis=1) / np.maximum(1.0, np.linalg.norm(b[ok, :2], axis=1))
    assert (err <= 1e-5).mean() > 0.999
    f.set_options(count_xforms=1) if n <= 40 else None
    W, H, P, TS = 128, 96, 256 * 2 * 8, 8
    rfk.set_sim_parameters(P, TS, 8, seed=n)
    f.warmup(8, TSS)
    buf = rfk.DeviceBuffer(W * H * 16)
    buf.zero_out()
    binned = f.draw_to_bins(buf.ptr, W * H, W, 32)
    bins = buf.download(np.float32, (H, W, 4))
    buf.free()
    assert binned > 0 and abs(float(bins[..., 3].sum()) - binned) < 0.5
    picks = f.xform_counts(n).astype(np.float64)
    assert picks.sum() == P * 32
    w = f.copy_flame_data_to_buffer()[[m_["weight"] for m_ in of.buffer_map["xforms"]]]
    assert np.abs(picks / picks.sum() - w).max() < 0.05


def test_non_default_stream(gpu_ready, rfk, compiler):
    """rfk_set_stream: all work is ordered on the caller's stream"""
    import torch
    from conftest import GENOME
    f = rfk.Flame.load_flame(GENOME, compiler)
    s = torch.cuda.Stream()
    rfk.lib().rfk_set_stream(s.cuda_stream)
    try:
        rfk.set_sim_parameters(256 * 8 * 16, 16, 8, seed=4)
        W, H = 160, 90
        with torch.cuda.stream(s):
            bins = torch.zeros(W * H * 4, dtype=torch.float32, device="cuda")
        s.synchronize()
        f.warmup(8, TSS)
        n = f.draw_to_bins(bins.data_ptr(), W * H, W, 16)
        s.synchronize()
        assert n > 0 and int(round(float(bins.view(-1, 4)[:, 3].sum()))) == n
    finally:
        rfk.lib().rfk_set_stream(None)


@pytest.mark.parametrize("ss,radius", [(2, 1.0), (1, 1.0), (3, 0.7), (2, 0.0)])
def test_spatial_downsample_matches_numpy(gpu_ready, rfk, ss, radius):
    """flam3-style spatial filter + supersample reduction against a direct numpy evaluation of the
    same separable taps; a constant image stays constant (weights renormalised at the border); radius 0 at ss = 2 is
    the 2x2 box of rfk_downsample2x"""
    import torch
    W, H = 37, 23
    rng = np.random.default_rng(ss)
    src = rng.random((H * ss, W * ss, 4)).astype(np.float32)
    taps = rfk.spatial_filter_taps(ss, radius).astype(np.float64)
    n, off = len(taps), (len(taps) - ss) // 2
    assert abs(taps.sum() - 1) < 1e-6 and np.allclose(taps, taps[::-1]) and (n - ss) % 2 == 0
    want = np.zeros((H, W, 4))
    for y in range(H):
        for x in range(W):
            acc, ws = np.zeros(4), 0.0
            for j in range(n):
                sy = y * ss - off + j
                if not 0 <= sy < H * ss:
                    continue
                for i in range(n):
                    sx = x * ss - off + i
                    if 0 <= sx < W * ss:
                        acc += taps[i] * taps[j] * src[sy, sx]
                        ws += taps[i] * taps[j]
            want[y, x] = acc / ws
    d_in = torch.from_numpy(src).cuda()
    d_out = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
    rfk.spatial_downsample(d_in.data_ptr(), d_out.data_ptr(), W, H, ss, radius)
    got = d_out.cpu().numpy()
    assert np.abs(got - want).max() < 1e-5
    const = torch.full((H * ss, W * ss, 4), 0.75, dtype=torch.float32, device="cuda")
    rfk.spatial_downsample(const.data_ptr(), d_out.data_ptr(), W, H, ss, radius)
    assert float((d_out - 0.75).abs().max()) < 1e-6
    if ss == 2 and radius == 0.0:
        box = torch.empty_like(d_out)
        rfk.downsample2x(d_in.data_ptr(), box.data_ptr(), W, H)
        assert torch.allclose(box, torch.from_numpy(got).cuda(), atol=1e-6)


def test_render_frame_supersampled(gpu_ready, rfk, flame):
    """supersample 2: histogram at twice the image size, tonemapped image reduced with the spatial filter; the result is
    the same picture as the plain render (same genome, far more samples than noise)"""
    W, H, P, TS = 160, 90, 256 * 16 * 32, 32
    rfk.set_sim_parameters(P, TS, 64, seed=5)
    flame.set_options(math_mode=1, fmad=1, per_lane_xform=0, warp_aggregate=0, deterministic=0, count_xforms=0, min_blocks=0, block_width=256, deal_period=1)
    plain = np.empty((H, W, 4), dtype=np.float32)
    _, s1 = flame.render_frame(W, H, target_binned=400 * W * H, drawing_passes=32, image_out=plain)
    ss_img = np.empty((H, W, 4), dtype=np.float32)
    ss_u8 = np.empty((H, W, 4), dtype=np.uint8)
    _, s2 = flame.render_frame(W, H, target_binned=400 * W * H * 4, drawing_passes=32, image_out=ss_img, rgba8_out=ss_u8, supersample=2, filter_radius=0.5)
    assert s2.binned >= 4 * 400 * W * H and np.isfinite(ss_img).all() and (ss_img[..., 3] > 0.99).all()
    assert np.abs(ss_u8.astype(int) - np.rint(np.clip(ss_img, 0, 1) * 255).astype(int)).max() <= 0
    # same scene: the brightness depends on samples per bin, equal here, so means agree
    assert abs(float(plain[..., :3].mean()) - float(ss_img[..., :3].mean())) < 0.08
    corr = np.corrcoef(plain[..., :3].ravel(), ss_img[..., :3].ravel())[0, 1]
    assert corr > 0.6, corr


def test_l2_hints_change_nothing_but_the_cache_policy(gpu_ready, rfk, flame):
    """kernel option l2_hints + hot map: the density channel sums opacities of 1, exact in binary32 whatever the order, so it
    must be bit-identical with and without the hints; the map marks the densest tiles and stays inside its budget."""
    W, H, P, TS = 2048, 1152, 256 * 64, 16
    hists = []
    for hints in (0, 1):
        flame.set_options(math_mode=1, fmad=1, per_lane_xform=0, warp_aggregate=0, deterministic=0, count_xforms=0, min_blocks=0, block_width=256, deal_period=1, l2_hints=hints)
        rfk.set_sim_parameters(P, TS, 64, seed=9)
        flame.warmup(16, TSS)
        buf = rfk.DeviceBuffer(W * H * 16)
        buf.zero_out()
        n = flame.draw_to_bins(buf.ptr, W * H, W, 32)
        if hints:
            budget = 2 * 1024 * 1024  # 512 tiles of 4 KB
            info = flame.build_hot_map(buf.ptr, W * H, W, budget)
            assert (info.tiles_x, info.tiles_y) == (W // 16, H // 16) and 0 < info.hot_tiles <= budget // 4096
            first = buf.download(np.float32, (H, W, 4))
            tiles = first[..., 3].reshape(H // 16, 16, W // 16, 16).sum(axis=(1, 3))
            hot = flame.hot_map()[: tiles.size].reshape(tiles.shape)
            assert hot.sum() == info.hot_tiles
            assert tiles[hot].min() >= tiles[~hot].max() * 0.999  # the densest tiles, bucket granularity aside
            assert tiles[hot].sum() / tiles.sum() > 0.3           # a flame's hits are concentrated
        n += flame.draw_to_bins(buf.ptr, W * H, W, 32)
        hists.append((buf.download(np.float32, (H, W, 4)), n))
        buf.free()
    flame.clear_hot_map()
    flame.set_options(l2_hints=0)
    (a, na), (b, nb) = hists
    assert na == nb and np.array_equal(a[..., 3], b[..., 3])
    assert np.allclose(a[..., :3], b[..., :3], rtol=1e-5, atol=1e-5)


def test_staged_bins_give_the_same_histogram(gpu_ready, rfk, flame, monkeypatch):
    """kernel option staged_bins: samples queued per region and accumulated afterwards are the same samples — the density
    channel (sums of opacity 1, exact in binary32 in any order) is bit-identical to the direct path, the colour sums agree to
    rounding; also with queues so small (one chunk per region) that most samples overflow into the direct reduction."""
    W, H, P, TS = 2048, 1152, 256 * 64, 16
    hists = []
    #        staged_bins, queue bytes, block_width, deal_period
    cases = [(0, None, 256, 1), (16, None, 256, 1), (16, 36 * 4096, 256, 1), (17, None, 256, 2), (18, None, 512, 1), (21, None, 128, 1)]
    for staged, max_bytes, width, period in cases:
        if max_bytes:
            monkeypatch.setenv("RFK_STAGE_MAX_BYTES", str(max_bytes))
        else:
            monkeypatch.delenv("RFK_STAGE_MAX_BYTES", raising=False)
        flame.set_options(math_mode=1, fmad=1, per_lane_xform=0, warp_aggregate=0, deterministic=0, count_xforms=0, min_blocks=0,
                          block_width=width, deal_period=period, l2_hints=0, staged_bins=staged)
        rfk.set_sim_parameters(P, TS, 64, seed=9)
        flame.warmup(16, TSS)
        buf = rfk.DeviceBuffer(W * H * 16)
        buf.zero_out()
        n = flame.draw_to_bins(buf.ptr, W * H, W, 32)
        n += flame.draw_to_bins(buf.ptr, W * H, W, 8)   # a shorter call reuses the queues of the longer one
        hists.append((buf.download(np.float32, (H, W, 4)), n))
        buf.free()
    flame.set_options(staged_bins=-1, block_width=256, deal_period=1)
    for h, n in hists:  # every sample landed exactly once, whatever the kernel options
        assert n > 0.5 * P * 40 and h[..., 3].astype(np.float64).sum() == n
    # the same kernel options apart from staging (also with exhausted queues): the same samples
    (a, na) = hists[0]
    for b, nb in (hists[1], hists[2]):
        assert na == nb and np.array_equal(a[..., 3], b[..., 3])
        assert np.allclose(a[..., :3], b[..., :3], rtol=1e-5, atol=1e-5)


def test_staged_bins_reject_too_many_regions(gpu_ready, rfk, flame):
    flame.set_options(staged_bins=8)
    rfk.set_sim_parameters(256 * 64, 16, 64, seed=9)
    flame.warmup(4, TSS)
    buf = rfk.DeviceBuffer(2048 * 1152 * 16)
    buf.zero_out()
    with pytest.raises(Exception):
        flame.draw_to_bins(buf.ptr, 2048 * 1152, 2048, 8)
    flame.set_options(staged_bins=-1)
    buf.free()


def test_staging_survives_a_device_short_of_memory(gpu_ready, rfk, flame, monkeypatch):
    """the queues are halved until they fit (what overflows is reduced directly); with no room at all the automatic mode draws
    without staging (one launch per call) and an explicit staged_bins fails loudly. Allocation failures are simulated."""
    W = H = 8192
    P, TS = 256 * 64, 16
    out = []
    for fail_above in (None, 8 << 20, 0):
        if fail_above is None:
            monkeypatch.delenv("RFK_STAGE_FAIL_ABOVE_BYTES", raising=False)
        else:
            monkeypatch.setenv("RFK_STAGE_FAIL_ABOVE_BYTES", str(fail_above))
        flame.set_options(staged_bins=0)
        flame.set_options(staged_bins=-1)  # a fresh device state: the earlier verdict on memory is forgotten
        rfk.set_sim_parameters(P, TS, 64, seed=12)
        flame.warmup(16, TSS)
        buf = rfk.DeviceBuffer(W * H * 16)
        buf.zero_out()
        before = rfk.kernel_launch_count()
        n = flame.draw_to_bins(buf.ptr, W * H, W, 16)
        launches = rfk.kernel_launch_count() - before
        n += flame.draw_to_bins(buf.ptr, W * H, W, 16)
        out.append((buf.download(np.float32, (H, W, 4))[..., 3].copy(), n, launches))
        buf.free()
    assert [o[2] for o in out] == [2, 2, 1]
    for d, n, _ in out[1:]:
        assert n == out[0][1] and np.array_equal(d, out[0][0])
    flame.set_options(staged_bins=22)
    rfk.set_sim_parameters(P, TS, 64, seed=12)
    flame.warmup(4, TSS)
    buf = rfk.DeviceBuffer(W * H * 16)
    with pytest.raises(rfk.RefraktError):
        flame.draw_to_bins(buf.ptr, W * H, W, 4)
    buf.free()
    monkeypatch.delenv("RFK_STAGE_FAIL_ABOVE_BYTES")
    flame.set_options(staged_bins=-1)


def test_render_frame_of_a_supersampled_8k_wide_histogram_goes_through_the_queues(gpu_ready, rfk, flame):
    """the end-to-end entry point (rfk_render_frame, the call the CLI makes): a 4096 x 2048 image from a 2x supersampled
    histogram (8192 x 4096 bins, 512 MiB) is drawn through the region queues by default — two launches per draw call — and
    the float image equals the one rendered with staging off (same samples; colour sums differ by rounding order)"""
    W, H = 4096, 2048
    imgs, stats = [], []
    for staged in (-1, 0):
        flame.set_options(staged_bins=staged)
        rfk.set_sim_parameters(256 * 256, 16, 64, seed=21)
        img = np.empty((H, W, 4), dtype=np.float32)
        before = rfk.kernel_launch_count()
        _, st = flame.render_frame(W, H, max_draw_calls=3, drawing_passes=32, warmup_passes=8, image_out=img, supersample=2, filter_radius=0.5)
        imgs.append(img)
        stats.append((st.binned, st.draw_calls, rfk.kernel_launch_count() - before))
    flame.set_options(staged_bins=-1)
    assert stats[0][0] == stats[1][0] > 0 and stats[0][1] == stats[1][1] == 3
    assert stats[0][2] - stats[1][2] == 3            # one accumulation kernel per draw call
    assert np.isfinite(imgs[0]).all() and imgs[0][..., :3].max() > 0
    assert np.allclose(imgs[0], imgs[1], rtol=1e-4, atol=1e-5)


def test_staging_is_automatic_for_a_histogram_of_one_gibibyte(gpu_ready, rfk, flame):
    """staged_bins = -1 (the default): a histogram of 512 MiB or more (here 1 GiB) goes through the queues (one more launch per call: the
    accumulation kernel; the staged kernels are built on first use and see the parameters of the last warmup), a smaller
    one does not; the samples are the same as with staging off"""
    W = H = 8192
    P, TS = 256 * 64, 16
    out = []
    for staged in (-1, 0):
        flame.set_options(staged_bins=staged)
        rfk.set_sim_parameters(P, TS, 64, seed=11)
        flame.warmup(16, TSS)
        buf = rfk.DeviceBuffer(W * H * 16)
        buf.zero_out()
        before = rfk.kernel_launch_count()
        n = flame.draw_to_bins(buf.ptr, W * H, W, 16)
        launches = rfk.kernel_launch_count() - before
        flame.warmup(8, TSS)  # new parameters reach the staged module as well
        n += flame.draw_to_bins(buf.ptr, W * H, W, 16)
        small = rfk.DeviceBuffer(1024 * 1024 * 16)
        small.zero_out()
        before = rfk.kernel_launch_count()
        flame.draw_to_bins(small.ptr, 1024 * 1024, 1024, 4)
        out.append((buf.download(np.float32, (H, W, 4))[..., 3].copy(), n, launches, rfk.kernel_launch_count() - before))
        buf.free()
        small.free()
    flame.set_options(staged_bins=-1)
    (a, na, la, sa), (b, nb, lb, sb) = out
    assert (la, lb, sa, sb) == (2, 1, 1, 1)
    assert na == nb and na > 0 and np.array_equal(a, b)


def test_release_buffers_frees_and_the_library_recovers(gpu_ready, rfk, flame):
    """rfk_release_buffers: device memory owned by the library goes back to the driver; drawing without a new
    set_sim_parameters fails loudly; after it, the same frame renders again"""
    import torch
    W, H = 1024, 576
    rfk.set_sim_parameters(256 * 64, 16, 64, seed=3)
    flame.set_options(math_mode=1, fmad=1, per_lane_xform=0, warp_aggregate=0, deterministic=0, count_xforms=0, min_blocks=0, block_width=256, deal_period=1, l2_hints=0)
    img1, st1 = flame.render_frame(W, H, max_draw_calls=1, drawing_passes=16)
    free_before, _ = torch.cuda.mem_get_info()
    rfk.release_buffers()
    free_after, _ = torch.cuda.mem_get_info()
    assert free_after - free_before >= W * H * 16  # at least the histogram came back
    with pytest.raises(rfk.RefraktError):
        flame.warmup(4, TSS)
    rfk.set_sim_parameters(256 * 64, 16, 64, seed=3)
    img2, st2 = flame.render_frame(W, H, max_draw_calls=1, drawing_passes=16)
    assert st2.binned == st1.binned and np.abs(img1.astype(int) - img2.astype(int)).max() <= 1
