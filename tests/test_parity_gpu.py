"""GPU parity tests: the CUDA path, called through the C ABI, against the oracle on the same
seeded inputs. Thresholds are the ones BASELINE.md §5 states."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N_PER_XFORM = 100_000


def _inputs(seed, n, spread=1.0):
    rng = np.random.default_rng(seed)
    xyz = np.concatenate([rng.normal(0.0, spread, (n, 2)), rng.random((n, 1))], axis=1).astype(np.float32)
    states = rng.integers(0, 2**32, (n, 4), dtype=np.uint64).astype(np.uint32)
    return xyz, states


def _rel_err(got, want):
    d = np.linalg.norm(got[:, :2].astype(np.float64) - want[:, :2].astype(np.float64), axis=1)
    return d / np.maximum(1.0, np.linalg.norm(want[:, :2].astype(np.float64), axis=1))


def _nudge(xyz, k):
    """inputs moved by k ulp in x and y"""
    out = xyz.copy()
    for _ in range(abs(k)):
        out[:, :2] = np.nextafter(out[:, :2], np.float32(np.inf if k > 0 else -np.inf))
    return out


@pytest.mark.parametrize("xid", list(range(-1, 10)))
def test_single_step_matches_oracle(gpu_ready, flame, oracle, xid):
    """dispatch(v, xid): xy within 1e-5 relative, colour within 1e-5, opacity equal, RNG state bit-exact.
    >= 99.9 % of vectors pass outright; every outlier passes against the oracle evaluated at an input
    nudged by <= 2 ulp (discontinuous variations: rectangles, julia, julian)."""
    xyz, states = _inputs(1000 + xid, N_PER_XFORM)
    ids = np.full(N_PER_XFORM, xid, dtype=np.int32)
    got, got_rng = flame.single_step(xyz, ids, states)
    want, want_rng = oracle.single_step(xyz, ids, states)
    assert np.array_equal(got_rng, want_rng)
    finite = np.isfinite(want).all(axis=1)
    assert finite.mean() > 0.999
    assert np.array_equal(np.isfinite(got).all(axis=1), finite) or np.mean(np.isfinite(got).all(axis=1) == finite) > 0.9999
    err = _rel_err(got, want)
    bad = finite & ~(err <= 1e-5)
    assert np.abs(got[finite, 2] - want[finite, 2]).max() <= 1e-5
    assert np.array_equal(got[finite, 3], want[finite, 3])
    assert bad.mean() <= 1e-3, "xform %d: %.4f%% of vectors outside 1e-5" % (xid, 100 * bad.mean())
    if bad.any():
        idx = np.nonzero(bad)[0]
        best = np.full(idx.size, np.inf)
        for k in (-2, -1, 1, 2):
            alt, _ = oracle.single_step(_nudge(xyz[idx], k), ids[idx], states[idx])
            best = np.minimum(best, _rel_err(got[idx], alt))
        assert (best <= 1e-5).all(), "xform %d: %d outliers not explained by a 2-ulp nudge (worst %g)" % (xid, (best > 1e-5).sum(), best.max())


def test_single_step_first_run_colour(gpu_ready, flame, oracle):
    """first_run: the colour coordinate comes from randf() (variation_table.cpp:242)."""
    xyz, states = _inputs(77, 20000)
    ids = np.random.default_rng(5).integers(0, 10, 20000).astype(np.int32)
    got, got_rng = flame.single_step(xyz, ids, states, first_run=True)
    want, want_rng = oracle.single_step(xyz, ids, states, first_run=True)
    assert np.array_equal(got_rng, want_rng)
    finite = np.isfinite(want).all(axis=1)
    assert np.abs(got[finite, 2] - want[finite, 2]).max() <= 1e-5
    assert (_rel_err(got, want)[finite] <= 1e-5).mean() >= 0.999


def test_single_step_custom_params(gpu_ready, flame, oracle):
    """a caller-supplied fp[] block (one temporal sample's rotated affines) instead of the flame's own"""
    fp = oracle.animate(8, 1.2 / 60)[1]
    fp1024 = np.zeros(1024, dtype=np.float32)
    fp1024[: fp.size] = fp
    xyz, states = _inputs(31, 30000)
    ids = np.random.default_rng(6).integers(-1, 10, 30000).astype(np.int32)
    got, got_rng = flame.single_step(xyz, ids, states, fp=fp1024)
    want, want_rng = oracle.single_step(xyz, ids, states, fp=fp1024)
    assert np.array_equal(got_rng, want_rng)
    finite = np.isfinite(want).all(axis=1)
    assert (_rel_err(got, want)[finite] <= 1e-5).mean() >= 0.999


def test_single_step_empty(gpu_ready, flame):
    out, rng = flame.single_step(np.zeros((0, 3), np.float32), np.zeros(0, np.int32), np.zeros((0, 4), np.uint32))
    assert out.shape == (0, 4) and rng.shape == (0, 4)


def test_select_xform_bit_exact(gpu_ready, flame, oracle):
    """get_xform_id: cumulative `sum >= ratio` in binary32, last xform is the fall-through"""
    rng = np.random.default_rng(3)
    fp = oracle.params()
    cum = np.cumsum(fp[[0, 13, 30, 43, 59, 76, 92, 110, 125]].astype(np.float32), dtype=np.float32)
    edge = np.concatenate([cum, np.nextafter(cum, np.float32(0)), np.nextafter(cum, np.float32(2)), [0.0, 1.0]]).astype(np.float32)
    ratio = np.concatenate([rng.random(200000).astype(np.float32), edge])
    got = flame.select_xform(ratio)
    want = oracle.select_xform(ratio)
    assert np.array_equal(got, want)
    assert set(np.unique(got)) == set(range(10))


@pytest.mark.parametrize("W,H", [(1280, 720), (3840, 2160), (15360, 8640), (7, 5)])
def test_bucket_index_bit_exact(gpu_ready, flame, oracle, oracle_mod, W, H):
    """(x, y, w) -> bin index incl. bounds test, opacity test and row flip; palette index — bit-exact"""
    rng = np.random.default_rng(W)
    n = 300000
    xyzw = np.zeros((n, 4), dtype=np.float32)
    xyzw[:, :2] = rng.normal(0, 1.6, (n, 2))
    xyzw[:, 2] = rng.random(n)
    xyzw[:, 3] = rng.choice([1.0, 0.5, 0.0, -1.0], n, p=[0.7, 0.2, 0.05, 0.05])
    xyzw[:50, 0] = [np.nan, np.inf, -np.inf, 1e30, -1e30] * 10
    xyzw[50:60, 2] = [0.0, 1.0, 1.0 / 255, 254.5 / 255, 2.0, -0.5, 1e-8, 0.999999, 0.5, 0.25]
    ss = flame.screen_space_affine(W, H)
    assert np.array_equal(ss.view(np.uint32), oracle_mod.screen_space_affine(oracle.flame, W, H).view(np.uint32))
    gi, gp = flame.bucket_index(xyzw, ss, W, H)
    wi, wp = oracle.bucket_index(xyzw, ss, W, H)
    assert np.array_equal(gi, wi)
    assert np.array_equal(gp, wp)
    assert (gi >= 0).mean() > 0.2 and gi.max() < W * H and (gi[xyzw[:, 3] <= 0] == -1).all()


def test_rng_seeding_bit_exact(gpu_ready, rfk, oracle):
    """jsf32::warmup_ctx(state, slot) on the device (src/util.hpp:90-95); SURVEY Appendix B values"""
    states = rfk.seed_rng_states(4096, 0)
    assert [hex(v) for v in states[0]] == ["0x1b517aa6", "0xd3d55a3", "0x44d68d47", "0x7a484bc9"]
    assert [hex(v) for v in states[1]] == ["0x927aed26", "0x131fa903", "0x750a9db8", "0xa696f285"]
    for i in (0, 1, 2, 77, 4095):
        assert np.array_equal(states[i], oracle.jsf32_warmup(i))
    off = rfk.seed_rng_states(16, 2097151 - 3)
    assert [hex(v) for v in off[3]] == ["0xf6e7ac5c", "0xe9cb99e", "0x73daf56", "0x299cd81f"]
    rfk.set_sim_parameters(256 * 4, 4, 8, seed=5)
    assert np.array_equal(rfk.copy_rng_states(0, 8), np.stack([oracle.jsf32_warmup(5 + i) for i in range(8)]))


@pytest.mark.parametrize("count", [4096, 256, 1000, 777, 1])
def test_sample_points_bit_exact(gpu_ready, rfk, oracle, count):
    got = rfk.make_sample_points(count)
    want = oracle.make_sample_points(count)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_shuffle_buffers_are_permutations(gpu_ready, rfk):
    """every buffer is a permutation of [0, size); reproducible from the seed; buffers differ"""
    for size in (4096, 1000, 256, 3):
        a = rfk.make_shuffle_buffers(size, 16, seed=42)
        assert np.array_equal(np.sort(a, axis=1), np.tile(np.arange(size, dtype=np.uint32), (16, 1)))
        assert np.array_equal(a, rfk.make_shuffle_buffers(size, 16, seed=42))
        if size > 16:
            assert not np.array_equal(a[0], a[1])
            assert not np.array_equal(a, rfk.make_shuffle_buffers(size, 16, seed=43))
            # no fixed structure: displacement of a uniform permutation averages size/3
            disp = np.abs(a.astype(np.int64) - np.arange(size)).mean()
            assert 0.25 * size < disp < 0.42 * size


def test_animate_matches_oracle(gpu_ready, flame, oracle):
    """per-temporal-sample parameter blocks (animate.tpl.glsl): copies exact, rotated affines within 1e-6"""
    for ts, width in ((512, 1.2 / 60), (32, 0.5), (64, 0.0)):
        got = flame.animate(ts, width)
        want = oracle.animate(ts, width)
        assert got.shape == want.shape
        np.testing.assert_allclose(got, want, rtol=2e-6, atol=1e-7)
        rotated = set()
        for m in oracle.slots7:
            rotated.update(int(s) for s in m[:4])
        keep = [i for i in range(got.shape[1]) if i not in rotated]
        assert np.array_equal(got[:, keep], want[:, keep])
    assert np.array_equal(flame.animate(64, 0.0)[5], oracle.params()[:169])


def _parity_over_genome(rfk, oracle_mod, compiler, vt, xml, n_per_xform, seed):
    """single-step parity for every xform of a synthetic genome; vectors whose reference result is non-finite or
    huge (exp/cosh/tan blow-ups) are excluded from the 1e-5 statistic but must be non-finite/huge on both sides"""
    f = rfk.Flame.load_flame_string(xml, compiler)
    assert f is not None, rfk.Flame.last_error()
    of = oracle_mod.load_flame_string(xml, vt)
    orc = oracle_mod.Oracle(of, vt)
    ids = list(range(len(of.xforms))) + ([-1] if of.final_xform is not None else [])
    report = {}
    for xid in ids:
        xyz, states = _inputs(seed + 17 * xid, n_per_xform, spread=0.8)
        idv = np.full(n_per_xform, xid, dtype=np.int32)
        got, got_rng = f.single_step(xyz, idv, states)
        want, want_rng = orc.single_step(xyz, idv, states)
        assert np.array_equal(got_rng, want_rng), "xform %d: RNG state" % xid
        sane = np.isfinite(want).all(axis=1) & (np.abs(want[:, :2]).max(axis=1) < 1e4)
        assert sane.mean() > 0.5, "xform %d: test vectors mostly blow up" % xid
        err = _rel_err(got, want)
        bad = sane & ~(err <= 1e-5)
        if bad.any():
            # An outlier is explained (a) by a discontinuity: the oracle at an input nudged by <= 2 ulp agrees within 1e-5; or
            # (b) by its distance to a pole: the oracle ITSELF moves by at least half the GPU's deviation when its input moves by
            # <= 64 ulp (7.6e-6 relative, inside the 1e-5 of the contract) — tan(3x) next to pi/2 in popcorn, the denominators
            # of cross / cpow / edisc ...: there one ulp of an INTERMEDIATE (the affine's result absorbs smaller nudges of the
            # input) is amplified in both implementations alike, and neither result is the "right" one to 1e-5.
            idx = np.nonzero(bad)[0]
            best = np.full(idx.size, np.inf)
            for k in (-2, -1, 1, 2):
                alt, _ = orc.single_step(_nudge(xyz[idx], k), idv[idx], states[idx])
                best = np.minimum(best, _rel_err(got[idx], alt))
            open_cases = best > 1e-5
            sens = np.zeros(idx.size)
            if open_cases.any():
                sub = idx[open_cases]
                steep = np.zeros(sub.size)
                for k in (-64, -16, -4, 4, 16, 64):
                    alt, _ = orc.single_step(_nudge(xyz[sub], k), idv[sub], states[sub])
                    with np.errstate(invalid="ignore"):
                        steep = np.fmax(steep, _rel_err(alt, want[sub]))
                sens[open_cases] = steep
                open_cases[open_cases] = ~(err[sub] <= 2.0 * steep)
            unexplained = int(open_cases.sum())
            for j in np.nonzero(open_cases)[0][:4]:  # shown when the assertion on the count fails
                i = idx[j]
                print("unexplained outlier: xform %d input %r got %r want %r err %.3g nudge-best %.3g sensitivity %.3g" % (
                    xid, xyz[i].tolist(), got[i, :2].tolist(), want[i, :2].tolist(), err[i], best[j], sens[j]))
        else:
            unexplained = 0
        names = list((of.final_xform if xid == -1 else of.xforms[xid]).variations)
        report[xid] = (names, float(bad.mean()), int(unexplained))
    return report


# Variations with a pole inside the sampled domain (a denominator that crosses zero: sin^2/cos in arch, 1/(1 + e x/r) in
# conic, 1/(cosh - cos) in foci and coth, 1/cos in ngon, 1/(x^2 - y^2)-like terms in twintrian). Next to the pole a 1-ulp
# difference in an intermediate is amplified without bound, in the oracle as much as on the GPU. Only the variations the
# measured per-variation table (profiles/r01_variation_parity_mode1.json, 50 000 vectors each) shows above the 1e-3 rule
# are listed, each with its measured fraction of vectors outside 1e-5; an xform that mixes some of them may miss 1e-5 on
# 1.5 x the sum of their fractions (plus the 1e-3 of the rule). Every other variation — julian, juliascope, spherical, log,
# power ... included — is under the strict rule: >= 99.9 % of the vectors within 1e-5 outright and every outlier explained
# by a <= 2-ulp nudge of the input.
POLE_FRACTION = {"foci": 0.01856, "ngon": 0.01356, "arch": 0.00884, "coth": 0.00496, "conic": 0.00366, "twintrian": 0.00156}
# the strict rule: at most 0.1 % of the vectors outside 1e-5, and of those all but 0.05 % of the vectors (10 of 20 000) explained
# by a discontinuity or by the distance to a pole (_parity_over_genome). The remainder is the approximation error of math
# mode 1 itself where it is amplified: pow through lg2 / ex2 with large exponents (cpow), the SFU sine's absolute error under
# a logarithm (twintrian). The shipped genome has no such case: test_single_step_matches_oracle requires every outlier explained.
STRICT_OUTSIDE, STRICT_UNEXPLAINED_FRACTION = 1e-3, 5e-4


def _check_report(report, n_vectors=20000):
    for xid, (names, frac_bad, unexplained) in report.items():
        poles = [n for n in names if n in POLE_FRACTION]
        if poles:
            bound = STRICT_OUTSIDE + 1.5 * sum(POLE_FRACTION[n] for n in poles)
            assert frac_bad <= bound, (xid, names, frac_bad, bound)
        else:
            assert frac_bad <= STRICT_OUTSIDE, (xid, names, frac_bad)
            assert unexplained <= STRICT_UNEXPLAINED_FRACTION * n_vectors, (xid, names, unexplained)


@pytest.mark.parametrize("chunk", range(6))
def test_single_step_all_variations(gpu_ready, rfk, oracle_mod, compiler, vt, chunk):
    """every compile-clean variation of variations.yaml (68), three per xform, against the oracle"""
    from conftest import chunk_genome
    _check_report(_parity_over_genome(rfk, oracle_mod, compiler, vt, chunk_genome(chunk, vt), 20000, 500 + chunk))


def test_single_step_overlay_and_stress_genome(gpu_ready, rfk, oracle_mod, overlay_compiler, overlay_vt):
    """the eight corrected variations and the 12-xform stress genome of BASELINE configs[4]"""
    from conftest import BROKEN, chunk_genome, stress_genome
    for xml in (chunk_genome(99, overlay_vt, names=BROKEN), stress_genome(overlay_vt)):
        _check_report(_parity_over_genome(rfk, oracle_mod, overlay_compiler, overlay_vt, xml, 20000, 900))


def test_sim_cache_export_import_round_trip(gpu_ready, rfk, oracle, tmp_path):
    """the reference's on-disk cache layout (src/flame.cpp:112-148): rand_state/<P>/ holds the JSF32 states, shuffle/<PPT>/
    one permutation per file; states written by one run seed another"""
    P, TS = 256 * 4 * 2, 4
    rfk.set_sim_parameters(P, TS, 8, seed=0)
    rfk.export_sim_cache(str(tmp_path), shuffle_seed=3)
    gs = rfk.BufferGroup(str(tmp_path), "rand_state", str(P))
    names = gs.cached_buffers()
    assert len(names) == 1
    states = gs.read_buffer(names[0], np.uint32).reshape(P, 4)
    assert np.array_equal(states[5], oracle.jsf32_warmup(5)) and np.array_equal(states, rfk.copy_rng_states(0, P))
    gp = rfk.BufferGroup(str(tmp_path), "shuffle", str(P // TS))
    perms = [gp.read_buffer(n, np.uint32) for n in gp.cached_buffers()]
    assert len(perms) == 8 and all(np.array_equal(np.sort(p), np.arange(P // TS)) for p in perms)
    rfk.set_sim_parameters(P, TS, 8, seed=999)
    assert not np.array_equal(rfk.copy_rng_states(0, P), states)
    rfk.import_rng_states(str(tmp_path))
    assert np.array_equal(rfk.copy_rng_states(0, P), states)
    rfk.set_rng_states(states[:4][::-1].copy(), first=10)
    assert np.array_equal(rfk.copy_rng_states(10, 4), states[:4][::-1])


def test_reference_pass_mode_reproduces_the_oracle(gpu_ready, rfk, compiler, oracle):
    """flame.glsl:41-90 pass by pass (load RNG, first thread picks the workgroup's xform, shuffle-buffer gather / scatter,
    first-run jitter, dispatch, histogram update, save RNG) driven by the host loops of flame.cpp:228-330, with the oracle's
    own shuffle tables and pass ids. RNG consumption depends only on the picked xforms, so after warmup AND draw the RNG
    states must be bit-identical; particle buffers agree to rounding while orbits are short, histograms bin for bin
    except where a discontinuous variation flips."""
    from conftest import GENOME
    import torch
    f = rfk.Flame.load_flame(GENOME, compiler)
    P, TS, NSHUF, W, H = 256 * 2 * 4, 4, 8, 160, 90
    oracle.set_threads(1)
    oracle.set_sim_parameters(P, TS, NSHUF, shuffle_seed=11, rng_seed=0, pass_seed=5)
    oracle.id_log(clear=True)
    rfk.set_sim_parameters(P, TS, NSHUF, seed=0)
    rfk.set_shuffle_buffers(oracle.shuffle_table(NSHUF))

    # warmup: first-run pass + 2 passes
    oracle.warmup(2, 1.2 / 60)
    ids = oracle.id_log(clear=True)
    assert ids.shape == (3, 2)
    f.reference_warmup(2, 1.2 / 60, ids)
    assert np.array_equal(rfk.copy_rng_states(0, P), oracle.rng_states(0, P))
    got, want = f.copy_particles(P), oracle.particles()
    ok = np.isfinite(want).all(axis=1)
    err = np.linalg.norm(got[ok, :2] - want[ok, :2], axis=1) / np.maximum(1.0, np.linalg.norm(want[ok, :2], axis=1))
    assert (err <= 1e-4).mean() > 0.97, (err <= 1e-4).mean()   # three chained iterations: rounding differences start to grow
    assert np.abs(got[ok, 2] - want[ok, 2]).max() <= 1e-5 and (got[:, 3] == 0).all()

    # one drawn pass from the oracle's exact particle state
    bins_o = np.zeros((H, W, 4), dtype=np.float32)
    n_o = oracle.draw_to_bins(bins_o, W, 3)
    ids = oracle.id_log(clear=True)
    assert ids.shape == (3, 2)
    d_bins = torch.zeros(W * H * 4, dtype=torch.float32, device="cuda")
    n_g = f.reference_draw_to_bins(d_bins.data_ptr(), W * H, W, 3, ids)
    assert np.array_equal(rfk.copy_rng_states(0, P), oracle.rng_states(0, P))  # still bit-identical after 6 passes
    bins_g = d_bins.view(H, W, 4).cpu().numpy()
    assert abs(n_g - n_o) <= 0.01 * n_o and abs(float(bins_g[..., 3].sum()) - n_g) < 0.5
    assert np.abs(bins_g - bins_o).sum() / np.abs(bins_o).sum() < 0.25  # same samples up to chaotic drift of 4-6 chained steps
    # the product kernels continue from the same particle buffers (state layout is shared)
    assert not f.needs_warmup()
    assert f.draw_to_bins(d_bins.data_ptr(), W * H, W, 4) > 0
    oracle.set_threads(oracle.max_threads())


def test_reference_pass_mode_statistics(gpu_ready, rfk, compiler, oracle_mod):
    """the baseline mode and the product kernels sample the same measure"""
    from conftest import GENOME
    import torch
    f = rfk.Flame.load_flame(GENOME, compiler)
    P, TS, W, H = 256 * 16 * 32, 32, 320, 180
    rfk.set_sim_parameters(P, TS, 64, seed=21)
    hists = []
    for mode in ("reference", "product"):
        bins = torch.zeros(W * H * 4, dtype=torch.float32, device="cuda")
        if mode == "reference":
            f.reference_warmup(16, 1.2 / 60)
            n = f.reference_draw_to_bins(bins.data_ptr(), W * H, W, 64)
        else:
            f.warmup(16, 1.2 / 60)
            n = f.draw_to_bins(bins.data_ptr(), W * H, W, 64)
        hists.append((bins.view(H, W, 4).cpu().numpy(), n))
    (a, na), (b, nb) = hists
    assert abs(na - nb) / nb < 0.004

    def pooled(x):
        return x[:, :, 3].astype(np.float64).reshape(H // 4, 4, W // 4, 4).sum(axis=(1, 3))
    l1 = 0.5 * np.abs(pooled(a) / pooled(a).sum() - pooled(b) / pooled(b).sum()).sum()
    assert l1 < 0.03, l1


def _random_genome(seed, vt):
    """2-6 xforms of 1-4 random compile-clean variations each, random post affines, optional final xform"""
    from conftest import COMPILE_CLEAN, GENOME_TEMPLATE, xform_xml
    rng = np.random.default_rng(1000 + seed)
    xf = []
    for k in range(int(rng.integers(2, 7))):
        names = sorted(set(str(n) for n in rng.choice(COMPILE_CLEAN, int(rng.integers(1, 5)), replace=False)))
        if all("pre_xform" in vt.vars[n].flags for n in names):
            names.append("linear")
        x = xform_xml(names, vt, rng)
        if rng.random() < 0.3:
            x = x.replace(" opacity=", ' post="%s" opacity=' % " ".join("%.3f" % v for v in rng.normal(0, 0.5, 6)))
        xf.append(x)
    if rng.random() < 0.5:
        xf.append(xform_xml(["linear", str(rng.choice(["bubble", "spherical", "eyefish", "julia"]))], vt, rng, tag="finalxform"))
    return GENOME_TEMPLATE % "\n".join(xf)


@pytest.mark.parametrize("seed", range(6))
def test_random_genomes(gpu_ready, rfk, oracle_mod, compiler, vt, seed):
    """random genomes over the whole variation table: identical generated text, single-step parity (pole-aware), and a small
    render whose histogram conserves mass"""
    xml = _random_genome(seed, vt)
    report = _parity_over_genome(rfk, oracle_mod, compiler, vt, xml, 10000, 3000 + seed)
    _check_report(report, 10000)
    f = rfk.Flame.load_flame_string(xml, compiler)
    of = oracle_mod.load_flame_string(xml, vt)
    assert f.glsl_source() == oracle_mod.compile_flame_xforms(of, vt)
    P, TS, W, H = 256 * 2 * 8, 8, 128, 96
    rfk.set_sim_parameters(P, TS, 8, seed=seed)
    f.warmup(8, 1.2 / 60)
    buf = rfk.DeviceBuffer(W * H * 16)
    buf.zero_out()
    binned = f.draw_to_bins(buf.ptr, W * H, W, 32)
    bins = buf.download(np.float32, (H, W, 4))
    buf.free()
    assert np.isfinite(bins).all() and (bins >= 0).all() and abs(float(bins[..., 3].sum()) - binned) < 0.5
