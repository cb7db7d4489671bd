"""GPU parity tests of the BASELINE.json configurations beyond the 4K still, against the oracle (not against another GPU
path): configs[2] a supersampled frame through the region queues, configs[3] animation frames (src/main.cpp:383-395),
configs[4] the stress genome at histogram and image level; and the value-specialised kernels (kernel option specialize)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TSS = 1.2 / 60.0


def _pooled(bins, k=4):
    d = bins[..., 3].astype(np.float64)
    H, W = d.shape
    return d[: H // k * k, : W // k * k].reshape(H // k, k, W // k, k).sum(axis=(1, 3))


def _norm_l1(a, b):
    return 0.5 * np.abs(a / a.sum() - b / b.sum()).sum()


def _psnr(a, b):
    mse = np.mean((a[..., :3].astype(np.float64) - b[..., :3].astype(np.float64)) ** 2)
    return 10 * np.log10(255.0 ** 2 / max(mse, 1e-12))


def _gpu_bins(rfk, flame, W, H, P, TS, passes, calls, seed, warm=16, **options):
    import torch
    if options:
        flame.set_options(**options)
    rfk.set_sim_parameters(P, TS, 64, seed=seed)
    flame.warmup(warm, TSS)
    bins = torch.zeros(W * H * 4, dtype=torch.float32, device="cuda")
    binned = 0
    for _ in range(calls):
        binned += flame.draw_to_bins(bins.data_ptr(), W * H, W, passes)
    return bins.view(H, W, 4).cpu().numpy(), binned


def _oracle_bins(orc, W, H, P, TS, passes, calls, rng_seed, shuf, warm=16):
    orc.set_sim_parameters(P, TS, 64, shuffle_seed=shuf, rng_seed=rng_seed)
    orc.warmup(warm, TSS)
    bins = np.zeros((H, W, 4), dtype=np.float32)
    binned = 0
    for _ in range(calls):
        binned += orc.draw_to_bins(bins, W, passes)
    return bins, binned


def _colour_by_region(bins, k=8):
    """mean palette colour per k x k block of blocks: rgb sums / density sums on a coarse grid (regions with mass only)"""
    H, W = bins.shape[:2]
    gh, gw = H // k, W // k
    b = bins[: gh * k, : gw * k].astype(np.float64).reshape(k, gh, k, gw, 4).sum(axis=(1, 3))
    return b[..., :3] / np.maximum(b[..., 3:4], 1e-9), b[..., 3]


@pytest.fixture(scope="module")
def stress(rfk, oracle_mod, overlay_compiler, overlay_vt):
    from conftest import stress_genome
    xml = stress_genome(overlay_vt)
    f = rfk.Flame.load_flame_string(xml, overlay_compiler)
    assert f is not None, rfk.Flame.last_error()
    orc = oracle_mod.Oracle(oracle_mod.load_flame_string(xml, overlay_vt), overlay_vt)
    return f, orc


def test_stress_genome_histogram_and_image_parity(gpu_ready, rfk, oracle_mod, stress):
    """BASELINE configs[4]: 12 xforms + final xform of divergent variations. A quarter of its particles overflow and never bin
    (the reference does not reset them); the rest must sample the oracle's measure: BASELINE.md §5 thresholds at small P."""
    f, orc = stress
    W, H, P, TS, passes, calls = 320, 180, 256 * 16 * 32, 32, 64, 2
    a, na = _oracle_bins(orc, W, H, P, TS, passes, calls, 0, 11)
    b, nb = _oracle_bins(orc, W, H, P, TS, passes, calls, P, 12)
    g, ng = _gpu_bins(rfk, f, W, H, P, TS, passes, calls, seed=2 * P)
    total = P * passes * calls
    assert na > 0.02 * total
    assert abs(ng / total - na / total) <= max(abs(nb - na) / total * 3, 0.01 * na / total + 3e-4), (ng / total, na / total, nb / total)
    self_l1 = _norm_l1(_pooled(a), _pooled(b))
    l1 = _norm_l1(_pooled(g), _pooled(a))
    assert l1 <= max(0.02, 1.5 * self_l1), (l1, self_l1)
    assert abs(g[..., 3].sum() - ng) <= 1e-3 * ng
    # colour per coarse region, where the region holds mass
    cg, mg = _colour_by_region(g)
    ca, ma = _colour_by_region(a)
    cb, _ = _colour_by_region(b)
    heavy = ma > 0.005 * ma.sum()
    assert heavy.sum() >= 4
    assert np.abs(cg - ca)[heavy].max() <= max(0.02, 2.0 * np.abs(cb - ca)[heavy].max())
    # image level: density estimation + tonemap of both histograms
    p = f.post_params()
    import torch
    d_bins = torch.from_numpy(g).cuda().contiguous()
    d_u8 = torch.empty(W * H * 4, dtype=torch.uint8, device="cuda")
    rfk.density_tonemap(d_bins.data_ptr(), None, d_u8.data_ptr(), W, H, p)
    img = d_u8.view(H, W, 4).cpu().numpy()
    ref_a = oracle_mod.to_rgba8(orc.tonemap(orc.density_estimate(a, W, H), scale_constant=p.scale_constant))
    ref_b = oracle_mod.to_rgba8(orc.tonemap(orc.density_estimate(b, W, H), scale_constant=p.scale_constant))
    self_psnr = _psnr(ref_a, ref_b)
    assert _psnr(img, ref_a) >= min(30.0, self_psnr - 1.0), (_psnr(img, ref_a), self_psnr)


@pytest.mark.parametrize("frame", [0, 17, 599])
def test_animation_frames_match_the_oracle(gpu_ready, rfk, compiler, vt, oracle_mod, frame):
    """BASELINE configs[3]: frame f has every animated xform rotated by 18 deg/s * f / 60 (src/main.cpp:224, :383-395, fixed
    dt = 1/60), then warmup + draw + density estimation + tonemap. The rotated parameters are bit-identical, the frame
    matches the oracle's frame of the same rotation statistically."""
    from conftest import GENOME
    f = rfk.Flame.load_flame(GENOME, compiler)
    of = oracle_mod.load_flame(GENOME, vt)
    for _ in range(frame):  # frame by frame, as the main loop accumulates the rotation (binary32 rounding included)
        f.rotate_xforms(18.0 / 60.0)
        for x in of.xforms + ([of.final_xform] if of.final_xform is not None else []):
            if x.rotation_frequency != 0:
                x.affine = list(oracle_mod.rotate_affine(x.affine, np.float32(18.0 / 60.0) * x.rotation_frequency))
    orc = oracle_mod.Oracle(of, vt)
    assert np.array_equal(f.copy_flame_data_to_buffer().view(np.uint32), orc.params().view(np.uint32))  # the rotated parameter block, bit for bit
    W, H, P, TS, passes, calls = 320, 180, 256 * 16 * 32, 32, 64, 2
    a, na = _oracle_bins(orc, W, H, P, TS, passes, calls, 0, 21)
    b, nb = _oracle_bins(orc, W, H, P, TS, passes, calls, P, 22)
    rfk.set_sim_parameters(P, TS, 64, seed=7 * P)
    img, stats = f.render_frame(W, H, max_draw_calls=calls, drawing_passes=passes)
    total = P * passes * calls
    assert abs(stats.binned / total - na / total) <= 0.004 * na / total + 3e-4 + abs(na - nb) / total
    p = f.post_params()
    ref_a = oracle_mod.to_rgba8(orc.tonemap(orc.density_estimate(a, W, H), scale_constant=p.scale_constant))
    ref_b = oracle_mod.to_rgba8(orc.tonemap(orc.density_estimate(b, W, H), scale_constant=p.scale_constant))
    self_psnr = _psnr(ref_a, ref_b)
    got = _psnr(img, ref_a)
    assert got >= min(30.0, self_psnr - 1.0), (frame, got, self_psnr)
    if frame:  # and the frame is not frame 0: the rotation moved the image
        f0 = rfk.Flame.load_flame(GENOME, compiler)
        img0, _ = f0.render_frame(W, H, max_draw_calls=calls, drawing_passes=passes)
        assert _psnr(img, img0) < got


def test_supersampled_frame_through_the_region_queues_matches_the_oracle(gpu_ready, rfk, flame, oracle):
    """BASELINE configs[2] in small: a 2x supersampled histogram drawn through the region queues (kernel option staged_bins)
    against the ORACLE's histogram of the same size — pooled density, in-bounds fraction, colour per region."""
    OW, OH, ss = 160, 90, 2
    W, H, P, TS, passes, calls = OW * ss, OH * ss, 256 * 16 * 32, 32, 64, 2
    a, na = _oracle_bins(oracle, W, H, P, TS, passes, calls, 0, 31)
    b, nb = _oracle_bins(oracle, W, H, P, TS, passes, calls, P, 32)
    try:
        g, ng = _gpu_bins(rfk, flame, W, H, P, TS, passes, calls, seed=3 * P, staged_bins=12)  # regions of 4096 bins: 15 queues
    finally:
        flame.set_options(staged_bins=-1)
    total = P * passes * calls
    assert abs(ng / total - na / total) <= 0.002 * na / total + 3e-4 + abs(na - nb) / total
    assert abs(g[..., 3].sum() - ng) <= 1e-3 * ng
    self_l1 = _norm_l1(_pooled(a), _pooled(b))
    assert _norm_l1(_pooled(g), _pooled(a)) <= max(0.02, 1.5 * self_l1)
    cg, _ = _colour_by_region(g)
    ca, ma = _colour_by_region(a)
    cb, _ = _colour_by_region(b)
    heavy = ma > 0.005 * ma.sum()
    assert np.abs(cg - ca)[heavy].max() <= max(0.02, 2.0 * np.abs(cb - ca)[heavy].max())


@pytest.mark.parametrize("xid", list(range(-1, 10)))
def test_specialised_single_step_matches_oracle(gpu_ready, rfk, compiler, oracle, xid):
    """kernel option specialize = 1: dispatch(v, xid) of the build that has the parameter values compiled in, same bar as the
    generic build (1e-5, RNG bit-exact, outliers explained by a <= 2-ulp nudge)"""
    from conftest import GENOME
    from test_parity_gpu import N_PER_XFORM, _inputs, _nudge, _rel_err
    f = rfk.Flame.load_flame(GENOME, compiler)
    f.set_options(specialize=1)
    xyz, states = _inputs(1000 + xid, N_PER_XFORM)
    ids = np.full(N_PER_XFORM, xid, dtype=np.int32)
    got, got_rng = f.single_step(xyz, ids, states)
    want, want_rng = oracle.single_step(xyz, ids, states)
    assert np.array_equal(got_rng, want_rng)
    finite = np.isfinite(want).all(axis=1)
    err = _rel_err(got, want)
    bad = finite & ~(err <= 1e-5)
    assert np.abs(got[finite, 2] - want[finite, 2]).max() <= 1e-5
    assert np.array_equal(got[finite, 3], want[finite, 3])
    assert bad.mean() <= 1e-3
    if bad.any():
        idx = np.nonzero(bad)[0]
        best = np.full(idx.size, np.inf)
        for k in (-2, -1, 1, 2):
            alt, _ = oracle.single_step(_nudge(xyz[idx], k), ids[idx], states[idx])
            best = np.minimum(best, _rel_err(got[idx], alt))
        assert (best <= 1e-5).all()


def test_specialised_kernels_draw_the_same_histogram(gpu_ready, rfk, compiler, oracle):
    """the value-specialised rfk_warm / rfk_draw against the generic build (same seeds: the same samples up to the order of
    the floating-point reductions) and against the oracle statistically; a changed value rebuilds them"""
    from conftest import GENOME
    f = rfk.Flame.load_flame(GENOME, compiler)
    W, H, P, TS, passes = 320, 180, 256 * 16 * 32, 32, 64
    g0, n0 = _gpu_bins(rfk, f, W, H, P, TS, passes, 1, seed=5, specialize=0)
    assert not f.uses_specialised()
    g1, n1 = _gpu_bins(rfk, f, W, H, P, TS, passes, 1, seed=5, specialize=1, pair_particles=0)
    assert f.uses_specialised()
    assert abs(n0 - n1) <= 1e-4 * n0
    assert _norm_l1(_pooled(g0), _pooled(g1)) <= 2e-3
    a, na = _oracle_bins(oracle, W, H, P, TS, passes, 1, 0, 41)
    b, nb = _oracle_bins(oracle, W, H, P, TS, passes, 1, P, 42)
    self_l1 = _norm_l1(_pooled(a), _pooled(b))
    assert _norm_l1(_pooled(g1), _pooled(a)) <= max(0.02, 1.5 * self_l1)
    # two particles per thread (the default of the specialised build): other picks for the same seeds, the same measure
    g2, n2 = _gpu_bins(rfk, f, W, H, P, TS, passes, 1, seed=5, specialize=1, pair_particles=2)
    assert f.uses_specialised() and f.pair_particles_state()[0] == 1
    total = P * passes
    assert abs(n2 / total - na / total) <= 0.002 * na / total + 3e-4 + abs(na - nb) / total
    assert abs(g2[..., 3].sum() - n2) <= 1e-3 * n2
    assert _norm_l1(_pooled(g2), _pooled(a)) <= max(0.02, 1.5 * self_l1)
    cg, _ = _colour_by_region(g2)
    ca, ma = _colour_by_region(a)
    cb, _ = _colour_by_region(b)
    heavy = ma > 0.005 * ma.sum()
    assert np.abs(cg - ca)[heavy].max() <= max(0.02, 2.0 * np.abs(cb - ca)[heavy].max())
    # pair_particles = 1: measured; whichever way it goes, the state and the two probe times are reported
    f.set_options(pair_particles=1)
    f.warmup(4, TSS)
    state, ms = f.pair_particles_state()
    assert state in (1, 2) and ms[0] > 0 and ms[1] > 0 and (state == 1) == (ms[1] < ms[0])
    # automatic mode: generic on the first warmup with these values, specialised from the second on; an edit goes back to generic
    f.set_options(specialize=2)
    x = f.xform(2)
    x.color = 0.25
    f.set_xform(2, x)
    f.warmup(4, TSS)
    assert not f.uses_specialised()
    f.warmup(4, TSS)
    assert f.uses_specialised()
    x.color = 0.75
    f.set_xform(2, x)
    f.warmup(4, TSS)
    assert not f.uses_specialised()
    # the specialised build really changes with the value: its source differs
    assert "0x1.8p-1f" in f.variant_source(False, True) or "0x1.8p-1" in f.variant_source(False, True)
