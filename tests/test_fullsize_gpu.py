"""Full-size (BASELINE configs 1/2 dimensions, P = 2 097 152, TS = 512) checks through size-independent
properties: mass conservation of the histogram, agreement of the in-bounds fraction with the oracle's,
bit-identical deterministic accumulation, linearity / shift structure of the density estimator, tonemap range."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

P, TS, TSS = 2048 * 1024, 512, 1.2 / 60.0


@pytest.fixture(scope="module")
def big(gpu_ready, rfk, compiler):
    from conftest import GENOME
    f = rfk.Flame.load_flame(GENOME, compiler)
    rfk.set_sim_parameters(P, TS, 1024, seed=0)
    return f


def _draw(rfk, f, W, H, calls, passes=128):
    f.warmup(16, TSS)
    bins = torch.zeros(W * H * 4, dtype=torch.float32, device="cuda")
    binned = 0
    for _ in range(calls):
        binned += f.draw_to_bins(bins.data_ptr(), W * H, W, passes)
    return bins, binned


@pytest.mark.parametrize("W,H", [(1280, 720), (3840, 2160)])
def test_histogram_mass_conservation_full_size(rfk, big, W, H):
    """every binned sample adds opacity 1 to the density channel and a palette colour to rgb: sum(density) == counter,
    rgb sums bounded by the palette range, nothing outside the histogram is touched"""
    guard = 4096
    buf = torch.zeros((W * H + 2 * guard) * 4, dtype=torch.float32, device="cuda")
    view = buf[guard * 4: (guard + W * H) * 4]
    big.warmup(16, TSS)
    binned = big.draw_to_bins(view.data_ptr(), W * H, W, 128) + big.draw_to_bins(view.data_ptr(), W * H, W, 128)
    b = view.view(-1, 4)
    assert int(b[:, 3].double().sum().item()) == binned == big.binned_total()
    assert abs(binned / (2 * 128 * P) - 0.8141) < 0.002  # the oracle's in-bounds fraction at 320x180 (resolution independent)
    assert float(buf[: guard * 4].abs().sum()) == 0.0 and float(buf[(guard + W * H) * 4:].abs().sum()) == 0.0
    pal = torch.from_numpy(big.palette()[:, :3]).cuda()
    ratio = b[:, :3].double().sum(0) / b[:, 3].double().sum()
    assert (ratio >= pal.min(0).values.double() - 1e-6).all() and (ratio <= pal.max(0).values.double() + 1e-6).all()
    assert torch.isfinite(b).all() and (b >= 0).all()


def test_deterministic_mode_full_size(rfk, big):
    W, H = 1280, 720
    big.set_options(deterministic=1)
    try:
        rfk.set_sim_parameters(P, TS, 1024, seed=3)
        a, na = _draw(rfk, big, W, H, 1)
        rfk.set_sim_parameters(P, TS, 1024, seed=3)
        b, nb = _draw(rfk, big, W, H, 1)
        assert na == nb and torch.equal(a, b)
    finally:
        big.set_options(deterministic=0)
        rfk.set_sim_parameters(P, TS, 1024, seed=0)


def test_density_estimator_structure_full_size(rfk, big):
    """at 4K on a real histogram: (a) radius 0 everywhere is the one-pixel left shift; (b) with a constant radius
    (curve 0) the estimator is linear: DE(2h) == 2 DE(h); (c) the tonemapped image is in [0, 1] with alpha 1 and is
    black exactly where the density image is empty"""
    W, H = 3840, 2160
    bins, binned = _draw(rfk, big, W, H, 4)
    p = big.post_params()
    img = torch.empty_like(bins)
    p.estimator_radius, p.estimator_min = 0, 0
    rfk.density_estimate(bins.data_ptr(), img.data_ptr(), W, H, p)
    h = bins.view(H, W, 4).flip(0)  # histogram row by = H-1-cy
    out = img.view(H, W, 4)
    assert torch.equal(out[:, :-1], h[:, 1:]) and float(out[:, -1].abs().sum()) == 0.0
    p.estimator_radius, p.estimator_min, p.estimator_curve = 3, 0, 0.0
    img2 = torch.empty_like(bins)
    rfk.density_estimate(bins.data_ptr(), img.data_ptr(), W, H, p)
    doubled = bins * 2.0
    rfk.density_estimate(doubled.data_ptr(), img2.data_ptr(), W, H, p)
    assert torch.allclose(img2, img * 2.0, rtol=1e-6, atol=1e-6)
    gain = float(img.view(-1, 4)[:, 3].double().sum() / bins.view(-1, 4)[:, 3].double().sum())
    assert abs(gain - 1.3629) < 0.02  # (1 + 0.5/r)^2-like gain of the radius-3 kernel, interior bins
    p = big.post_params()
    rfk.density_estimate(bins.data_ptr(), img.data_ptr(), W, H, p)
    tm = torch.empty_like(bins)
    u8 = torch.empty(W * H * 4, dtype=torch.uint8, device="cuda")
    rfk.tonemap(img.data_ptr(), tm.data_ptr(), u8.data_ptr(), W, H, p)
    t = tm.view(-1, 4)
    assert float(t.min()) >= 0.0 and float(t.max()) <= 1.0 and bool((t[:, 3] == 1.0).all())
    empty = img.view(-1, 4)[:, 3] == 0
    assert float(t[empty][:, :3].abs().sum()) == 0.0 and float(t[~empty][:, :3].max(1).values.mean()) > 0.01
    assert torch.equal(u8.view(-1, 4), torch.round(t * 255.0).to(torch.uint8))
    fused = torch.empty_like(bins)
    rfk.density_tonemap(bins.data_ptr(), fused.data_ptr(), None, W, H, p)
    assert torch.equal(fused, tm)


def test_xform_selection_frequencies_full_size(rfk, big):
    """268 M iterations: selection frequencies match the normalised weights (the read-out of main.cpp:595-611)"""
    big.set_options(count_xforms=1)
    try:
        W, H = 1280, 720
        bins, binned = _draw(rfk, big, W, H, 1)
        picks = big.xform_counts(10).astype(np.float64)
        assert picks.sum() == 128 * P
        weights = big.copy_flame_data_to_buffer()[[0, 13, 30, 43, 59, 76, 92, 110, 125, 141]]
        assert np.abs(picks / picks.sum() - weights).max() <= 1e-3
    finally:
        big.set_options(count_xforms=0)


def test_config1_histogram_and_image_parity_full_size(rfk, big, oracle, oracle_mod):
    """BASELINE configs[0] at its true size (1280x720, P = 2 097 152, TS = 512, warmup 16, one draw_to_bins of 128 passes =
    268 435 456 iterations): GPU against two independent oracle runs on the host cores — BASELINE.md §5 'histogram' and
    'final image' rows with their self-noise terms measured here"""
    W, H = 1280, 720
    refs = []
    for rng_seed, shuf, pas in ((0, 0x5EED0000, 0x5EED0001), (P, 0x77, 0x78)):
        oracle.set_sim_parameters(P, TS, 64, shuffle_seed=shuf, rng_seed=rng_seed, pass_seed=pas)
        oracle.warmup(16, TSS)
        bins = np.zeros((H, W, 4), dtype=np.float32)
        n = oracle.draw_to_bins(bins, W, 128)
        refs.append((bins, n))
    rfk.set_sim_parameters(P, TS, 1024, seed=2 * P)
    big.warmup(16, TSS)
    d_bins = torch.zeros(W * H * 4, dtype=torch.float32, device="cuda")
    binned = big.draw_to_bins(d_bins.data_ptr(), W * H, W, 128)
    got = d_bins.view(H, W, 4).cpu().numpy()

    def pooled(b, k=4):
        d = b[: H // k * k, : W // k * k, 3].astype(np.float64)
        return d.reshape(H // k, k, W // k, k).sum(axis=(1, 3))

    def l1(a, b):
        return 0.5 * np.abs(a / a.sum() - b / b.sum()).sum()

    self_l1 = l1(pooled(refs[0][0]), pooled(refs[1][0]))
    gpu_l1 = l1(pooled(got), pooled(refs[0][0]))
    assert gpu_l1 <= max(0.02, 1.5 * self_l1), (gpu_l1, self_l1)
    total = 128 * P
    assert abs(binned / total - refs[0][1] / total) <= 0.002 * refs[0][1] / total, (binned / total, refs[0][1] / total, refs[1][1] / total)

    def image(b):
        return oracle_mod.to_rgba8(oracle.tonemap(oracle.density_estimate(b, W, H)))

    ref_imgs = [image(refs[0][0]), image(refs[1][0])]
    p = big.post_params()
    u8 = torch.empty(W * H * 4, dtype=torch.uint8, device="cuda")
    rfk.density_tonemap(d_bins.data_ptr(), None, u8.data_ptr(), W, H, p)
    gpu_img = u8.view(H, W, 4).cpu().numpy()
    # the GPU post-processing of the ORACLE's histogram matches the oracle's own image to 1 LSB
    d_ref = torch.from_numpy(refs[0][0]).cuda()
    rfk.density_tonemap(d_ref.data_ptr(), None, u8.data_ptr(), W, H, p)
    assert np.abs(u8.view(H, W, 4).cpu().numpy().astype(int) - ref_imgs[0].astype(int)).max() <= 1

    def psnr(a, b):
        mse = np.mean((a[..., :3].astype(np.float64) - b[..., :3].astype(np.float64)) ** 2)
        return 10 * np.log10(255.0 ** 2 / mse)

    self_psnr, gpu_psnr = psnr(ref_imgs[0], ref_imgs[1]), psnr(gpu_img, ref_imgs[0])
    assert gpu_psnr >= min(30.0, self_psnr - 1.0), (gpu_psnr, self_psnr)
    mae = np.mean(np.abs(gpu_img[..., :3].astype(np.float64) - ref_imgs[0][..., :3].astype(np.float64)))
    self_mae = np.mean(np.abs(ref_imgs[1][..., :3].astype(np.float64) - ref_imgs[0][..., :3].astype(np.float64)))
    assert mae <= max(2.0, 1.25 * self_mae), (mae, self_mae)
    print("config 1 parity: L1 gpu %.4f self %.4f | in-bounds gpu %.5f ref %.5f %.5f | PSNR gpu %.2f self %.2f | MAE gpu %.3f self %.3f" % (
        gpu_l1, self_l1, binned / total, refs[0][1] / total, refs[1][1] / total, gpu_psnr, self_psnr, mae, self_mae))
