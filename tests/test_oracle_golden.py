"""The oracle against every pin that exists for it (the reference has no tests of its own):
SURVEY Appendix B values, and the two reference functions that compile from /root/reference
(oracle/Makefile -> oracle/_ref/libref_pins.so)."""
import ctypes
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

PINS = os.path.join(ROOT, "oracle", "_ref", "libref_pins.so")


def test_jsf32_appendix_b(oracle):
    assert ["%08x" % v for v in oracle.jsf32_warmup(0)] == ["1b517aa6", "0d3d55a3", "44d68d47", "7a484bc9"]
    assert ["%08x" % v for v in oracle.jsf32_warmup(1)] == ["927aed26", "131fa903", "750a9db8", "a696f285"]
    assert ["%08x" % v for v in oracle.jsf32_warmup(2097151)] == ["f6e7ac5c", "0e9cb99e", "073daf56", "299cd81f"]


def test_hammersley_appendix_b(oracle):
    p = oracle.make_sample_points(4096)
    assert p[0].tolist() == [-1.0, -1.0, 0.0, 0.0]
    assert p[1, 0] == np.float32(-0.999511719) and p[1, 1] == 0.0
    assert p[2].tolist()[:2] == [np.float32(-0.999023438), -0.5] and p[3, 1] == 0.5
    assert p[4095, 0] == p[4095, 1] == np.float32(0.999511719)


def test_device_randf_semantics(oracle):
    """random.glsl:29-41: returns `a` after the update; randf = float(u) * 2^-32 in [0, 1]"""
    state = oracle.jsf32_warmup(7)
    vals, after = oracle.device_randf(state, 1000)
    assert vals.min() >= 0.0 and vals.max() <= 1.0 and 0.45 < vals.mean() < 0.55
    # first draw by hand
    a, b, c, d = (int(v) for v in state)
    rot = lambda x, k: ((x << k) | (x >> (32 - k))) & 0xFFFFFFFF
    e = (a - rot(b, 27)) & 0xFFFFFFFF
    a2 = b ^ rot(c, 17)
    assert vals[0] == np.float32(np.float32(a2) / np.float32(4294967295.0))


@pytest.mark.skipif(not os.path.exists(PINS), reason="oracle/_ref not built (needs /root/reference at build time)")
def test_restatement_matches_reference_compiled_pins(oracle):
    ref = ctypes.CDLL(PINS)
    for seed in (0, 1, 2, 12345, 2097151, 0xFFFFFFFF):
        out = np.zeros(4, dtype=np.uint32)
        ref.ref_jsf32_warmup(ctypes.c_uint32(seed), out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)))
        assert np.array_equal(out, oracle.jsf32_warmup(seed))
    for count in (4096, 256, 1000, 777, 2, 1):
        out = np.zeros((count, 4), dtype=np.float32)
        ref.ref_make_sample_points(ctypes.c_uint32(count), out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
        assert np.array_equal(out.view(np.uint32), oracle.make_sample_points(count).view(np.uint32))


def test_golden_single_step_vectors(oracle):
    """committed fixture (tests/golden/make_golden.py): guards the oracle itself against drift"""
    path = os.path.join(GOLDEN, "single_step_electricsheep.npz")
    g = np.load(path)
    out, rng = oracle.single_step(g["xyz"], g["xid"], g["rng_in"])
    assert np.array_equal(rng, g["rng_out"])
    np.testing.assert_allclose(out, g["out"], rtol=2e-6, atol=2e-7)


def test_golden_density_tonemap(oracle):
    g = np.load(os.path.join(GOLDEN, "density_tonemap_small.npz"))
    H, W = g["bins"].shape[:2]
    de = oracle.density_estimate(g["bins"], W, H, int(g["radius"]), int(g["min"]), float(g["curve"]))
    np.testing.assert_allclose(de, g["de"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(oracle.tonemap(de, scale_constant=1e-4), g["tonemapped"], rtol=1e-5, atol=1e-6)
    # Appendix D: kernel gain by radius for an isolated bin
    for r, gain in ((1, 2.326), (2, 1.600), (5, 1.214), (11, 1.093)):
        b = np.zeros((40, 40, 4), dtype=np.float32)
        b[20, 20] = 1.0
        out = oracle.density_estimate(b, 40, 40, r, 0, 0.0)
        assert abs(out[..., 3].sum() - gain) < 2e-3


def test_oracle_render_is_self_consistent(oracle):
    """sequential oracle: counter equals the density sum (opacity 1), in-bounds fraction ~0.81, warmup leaves no NaNs"""
    oracle.set_sim_parameters(256 * 2 * 8, 8, 16)
    oracle.warmup(16, 1.2 / 60)
    parts = oracle.particles()
    assert np.isfinite(parts).mean() > 0.999
    bins = np.zeros((90, 160, 4), dtype=np.float32)
    n = oracle.draw_to_bins(bins, 160, 32, count_xforms=True)
    assert n == int(round(float(bins[..., 3].sum())))
    assert 0.7 < n / (256 * 2 * 8 * 32) < 0.9
    assert oracle.xform_picks(10).sum() == 256 * 2 * 8 * 32


PINS_FLAME = os.path.join(ROOT, "oracle", "_ref", "libref_pins_flame.so")


@pytest.mark.skipif(not os.path.exists(PINS_FLAME), reason="oracle/_ref flame.hpp pins not built")
def test_affine_helpers_match_reference_compiled_flame_hpp(oracle, oracle_mod, rfk, flame):
    """rotate/scale/translate_affine and the screen-space affine of draw_to_bins, compiled from the reference's
    src/flame.hpp:97-128, against the oracle's and the product's restatements — bit-exact"""
    ref = ctypes.CDLL(PINS_FLAME)
    fp = ctypes.POINTER(ctypes.c_float)
    rng = np.random.default_rng(11)
    for _ in range(200):
        a = rng.normal(0, 2, 6).astype(np.float32)
        deg, s = np.float32(rng.uniform(-1000, 1000)), np.float32(rng.uniform(0.01, 500))
        t = rng.normal(0, 3, 2).astype(np.float32)
        out = np.zeros(6, dtype=np.float32)
        ref.ref_rotate_affine(a.ctypes.data_as(fp), ctypes.c_float(deg), out.ctypes.data_as(fp))
        assert np.array_equal(out.view(np.uint32), np.array(oracle_mod.rotate_affine(a, deg), dtype=np.float32).view(np.uint32))
        assert np.array_equal(out.view(np.uint32), rfk.rotate_affine(a, deg).view(np.uint32))
        ref.ref_scale_affine(a.ctypes.data_as(fp), ctypes.c_float(s), out.ctypes.data_as(fp))
        assert np.array_equal(out.view(np.uint32), np.array(oracle_mod.scale_affine(a, s), dtype=np.float32).view(np.uint32))
        assert np.array_equal(out.view(np.uint32), rfk.scale_affine(a, s).view(np.uint32))
        ref.ref_translate_affine(a.ctypes.data_as(fp), t.ctypes.data_as(fp), out.ctypes.data_as(fp))
        assert np.array_equal(out.view(np.uint32), np.array(oracle_mod.translate_affine(a, t), dtype=np.float32).view(np.uint32))
        assert np.array_equal(out.view(np.uint32), rfk.translate_affine(a, t).view(np.uint32))
    f = oracle.flame
    for W, H in ((1280, 720), (3840, 2160), (7680, 4320), (15360, 8640), (333, 77)):
        out = np.zeros(6, dtype=np.float32)
        ref.ref_screen_space_affine(ctypes.c_float(f.scale), ctypes.c_float(f.rotate), ctypes.c_float(f.center[0]), ctypes.c_float(f.center[1]),
                                    ctypes.c_uint(f.size[1]), ctypes.c_ulong(W), ctypes.c_ulong(H), out.ctypes.data_as(fp))
        assert np.array_equal(out.view(np.uint32), oracle_mod.screen_space_affine(f, W, H).view(np.uint32))
        assert np.array_equal(out.view(np.uint32), flame.screen_space_affine(W, H).view(np.uint32))


@pytest.mark.skipif(not os.path.exists(PINS), reason="oracle/_ref not built (needs /root/reference at build time)")
def test_macro_grammar_matches_reference_compiled_util_cpp(oracle_mod, vt, rfk):
    """replace_macro / find_macros (src/util.cpp:6-23, std::regex) against the oracle's restatement, on every snippet of
    variations.yaml and on adversarial strings (macro at the end of the text, adjacent macros, prefixes of longer names)"""
    ref = ctypes.CDLL(PINS)
    if not hasattr(ref, "ref_replace_macro"):
        pytest.skip("prebuilt pins predate the util.cpp wrappers")
    texts = [v.source for v in vt.vars.values()] + [v.result for v in vt.vars.values()] + list(vt.common.values())
    texts += ["$x", "$x$x $y", "a$x)", "$xy + $x_y + $x1 + $x ", "$$x $x$ ", "", "$weight *$v", "$c10*$c1 $c100 ", "$r;$r\n$r", "x$y.z$v,"]
    names = ["x", "y", "v", "weight", "r", "c10", "c1", "result", "julian_power", "a"]
    buf = ctypes.create_string_buffer(1 << 16)
    for t in texts:
        n = ref.ref_find_macros(t.encode(), buf, len(buf))
        assert n >= 0
        got = set(buf.value.decode().split("\n")) - {""}
        assert got == oracle_mod.find_macros(t), t
        for name in names:
            n = ref.ref_replace_macro(t.encode(), name.encode(), b"fp[7]", buf, len(buf))
            assert n >= 0 and buf.value.decode() == oracle_mod.replace_macro(t, name, "fp[7]"), (t, name)
            assert buf.value.decode() == rfk.replace_macro(t, name, "fp[7]"), (t, name)  # the product's own (csrc/textutil.cpp)
        assert rfk.find_macros(t) == oracle_mod.find_macros(t), t


@pytest.mark.skipif(not os.path.exists(PINS), reason="oracle/_ref not built (needs /root/reference at build time)")
def test_macro_grammar_property_based(oracle_mod, rfk):
    """hypothesis-generated strings over the macro alphabet: reference util.cpp == oracle == product"""
    from hypothesis import given, settings, strategies as st
    ref = ctypes.CDLL(PINS)
    if not hasattr(ref, "ref_replace_macro"):
        pytest.skip("prebuilt pins predate the util.cpp wrappers")
    buf = ctypes.create_string_buffer(1 << 14)
    alphabet = st.sampled_from(list("$$$xyvrc01_ab ()*;.\n"))

    @settings(max_examples=300, deadline=None)
    @given(st.text(alphabet, max_size=40), st.sampled_from(["x", "y", "v", "r", "c1", "c10", "a", "ab", "x_1"]), st.sampled_from(["fp[3]", "v.x", "", "$x"]))
    def check(text, name, value):
        n = ref.ref_replace_macro(text.encode(), name.encode(), value.encode(), buf, len(buf))
        assert n >= 0
        want = buf.value.decode()
        assert oracle_mod.replace_macro(text, name, value) == want
        assert rfk.replace_macro(text, name, value) == want
        n = ref.ref_find_macros(text.encode(), buf, len(buf))
        assert n >= 0
        found = set(buf.value.decode().split("\n")) - {""}
        assert oracle_mod.find_macros(text) == found and rfk.find_macros(text) == found

    check()


@pytest.mark.skipif(not os.path.exists(PINS), reason="oracle/_ref not built (needs /root/reference at build time)")
def test_cache_files_are_readable_by_the_reference_reader(rfk, tmp_path, monkeypatch):
    """files written by the product's buffer cache, read back by the reference's own buffer_group::cached_buffers /
    read_buffer (src/buffer_cache.hpp:27-50, compiled from the reference header)"""
    ref = ctypes.CDLL(PINS)
    if not hasattr(ref, "ref_cache_list"):
        pytest.skip("prebuilt pins predate the buffer_cache wrappers")
    ref.ref_cache_read_u32.restype = ctypes.c_long
    perm = np.random.default_rng(3).permutation(4096).astype(np.uint32)
    states = np.random.default_rng(4).integers(0, 2**32, (64, 4), dtype=np.uint64).astype(np.uint32)
    name_p = rfk.BufferGroup(str(tmp_path), "shuffle", "4096").write_buffer(perm)
    name_s = rfk.BufferGroup(str(tmp_path), "rand_state", "64").write_buffer(states)
    monkeypatch.chdir(tmp_path)  # the reference resolves "cache/" against the working directory
    buf = ctypes.create_string_buffer(4096)
    assert ref.ref_cache_list(b"shuffle", b"4096", buf, len(buf)) > 0 and buf.value.decode().split() == [name_p]
    out = np.zeros(4096, dtype=np.uint32)
    n = ref.ref_cache_read_u32(b"shuffle", b"4096", name_p.encode(), out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), 4096)
    assert n == 4096 and np.array_equal(out, perm)
    out = np.zeros(256, dtype=np.uint32)
    n = ref.ref_cache_read_u32(b"rand_state", b"64", name_s.encode(), out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), 256)
    assert n == 256 and np.array_equal(out.reshape(64, 4), states)
