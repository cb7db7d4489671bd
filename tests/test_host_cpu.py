"""CPU tests (no GPU): the C-ABI library loads and exports every symbol the header declares; genome
parsing, buffer map, parameter buffer and generated GLSL match the oracle and SURVEY Appendix A/B;
error behaviour of load_flame; NVRTC builds of the shipped genome and of every compile-clean variation."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import BROKEN, COMPILE_CLEAN, FIXTURES, GENOME, GENOME_TEMPLATE, ROOT, VARIATIONS, chunk_genome, stress_genome

def test_abi_exports_every_declared_symbol(rfk):
    header = open(os.path.join(ROOT, "include", "refrakt_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(rfk_[a-z0-9_]+)\s*\(", header))
    assert len(declared) > 60
    lib = rfk.lib()
    for name in sorted(declared):
        assert hasattr(lib, name), "library does not export " + name
    assert declared == set(rfk.SIGNATURES), declared ^ set(rfk.SIGNATURES)
    assert lib.rfk_abi_version() == 4
    out = subprocess.run(["nm", "-D", "--defined-only", rfk.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    exported = set(re.findall(r"\bT (rfk_[a-z0-9_]+)", out))
    assert declared <= exported


def test_library_has_no_libcuda_link_dependency(rfk):
    """loads on a box without a GPU driver; GPU entry points fail loudly instead of falling back"""
    out = subprocess.run(["ldd", rfk.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    assert "libcuda.so" not in out and "libnvrtc" in out


def test_variation_table(compiler, vt):
    names = compiler.variations()
    assert len(names) == 76 and names == sorted(names) and set(names) == set(vt.vars)
    assert set(COMPILE_CLEAN) | set(BROKEN) == set(names)
    for n in ("r", "rsq", "a", "phi", "sina", "cosa", "sinr", "cosr", "lr"):
        assert compiler.is_common(n)
    assert compiler.is_variation("julian") and not compiler.is_variation("julian_power")
    assert compiler.is_param("julian_power") and compiler.param_owner("julian_power") == "julian"
    assert compiler.get_parameters_for_variation("julian") == ["julian_power", "julian_dist"]  # document order
    assert compiler.get_parameters_for_variation("linear") == []
    assert compiler.get_parameters_for_variation("mobius") == vt.vars["mobius"].param
    assert not compiler.is_param("weight") and not compiler.is_variation("coefs")
    with pytest.raises(Exception):
        compiler.get_parameters_for_variation("no_such_variation")


def test_compiler_missing_file(rfk):
    with pytest.raises(rfk.RefraktError):
        rfk.FlameCompiler("/nonexistent/variations.yaml")


def test_genome_fields(flame, oracle):
    i, f = flame.info(), oracle.flame
    assert (i.num_xforms, i.has_final_xform, i.param_count) == (10, 1, 169)
    assert list(i.size) == [800, 592] == f.size
    assert np.float32(i.scale) == f.scale and np.float32(i.rotate) == f.rotate
    assert [np.float32(c) for c in i.center] == f.center
    assert (i.estimator_min, i.estimator_radius) == (0, 11)  # estimator_minimum is not read (flame.cpp:171)
    assert np.float32(i.estimator_curve) == np.float32(0.6) and i.gamma == 4.0 and i.vibrancy == 1.0
    assert np.float32(i.brightness) == np.float32(29.718)
    for k in list(range(10)) + [-1]:
        x, ox = flame.xform(k), (f.final_xform if k == -1 else f.xforms[k])
        assert [np.float32(v) for v in x.affine] == ox.affine
        assert bool(x.has_post) == (ox.post is not None)
        if ox.post is not None:
            assert [np.float32(v) for v in x.post] == ox.post
        assert flame.variations(k) == {n: float(v) for n, v in ox.variations.items()}
        assert flame.params_of(k) == {n: float(v) for n, v in ox.var_param.items()}
        assert (np.float32(x.weight), np.float32(x.color), np.float32(x.color_speed), np.float32(x.opacity)) == (ox.weight, ox.color, ox.color_speed, ox.opacity)
        assert np.float32(x.rotation_frequency) == ox.rotation_frequency
    assert flame.xform(-1).weight == 0.0 and flame.xform(-1).rotation_frequency == 0.0
    assert flame.xform(2).rotation_frequency == 1.0  # animate="0.265579" > 0
    pal = flame.palette()
    assert np.array_equal(pal, f.palette)
    assert pal[0].tolist() == [134 / 256.0, 181 / 256.0, 109 / 256.0, 1.0]  # stoi truncation (flame.cpp:217)


def test_buffer_map_matches_survey_appendix_a(flame, oracle):
    import json
    bm = json.loads(flame.buffer_map_json())
    assert bm == oracle.flame.buffer_map
    assert bm["size"] == 169
    starts = [x["meta"]["start"] for x in bm["xforms"]]
    assert starts == [0, 13, 30, 43, 59, 76, 92, 110, 125, 141] and bm["final_xform"]["meta"]["start"] == 153
    assert bm["xforms"][6]["post"] == list(range(99, 105)) and bm["xforms"][6]["variations"] == {"bubble": 105}
    assert bm["xforms"][4]["variations"] == {"julian": 66, "rectangles": 67}
    assert bm["xforms"][4]["param"] == {"julian_dist": 68, "julian_power": 69, "rectangles_x": 70, "rectangles_y": 71}
    assert bm["final_xform"]["variations"] == {"bubble": 160, "linear": 161, "rectangles": 162}
    assert [bm["xforms"][9][k] for k in ("color", "color_speed", "opacity", "rotation_frequency")] == [149, 150, 151, 152]


def test_parameter_buffer_bit_exact(flame, oracle):
    got, want = flame.copy_flame_data_to_buffer(), oracle.params()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    w = got[[0, 13, 30, 43, 59, 76, 92, 110, 125, 141]]
    assert abs(w.sum() - 1.0) < 1e-6 and abs(w[4] - 0.3959) < 1e-4  # normalised weights (flame.cpp:78-82)
    assert got[153] == 0.0  # final xform weight 0 / sum


def test_generated_glsl_is_the_reference_text(flame, oracle):
    """the product's compile_flame_xforms output equals the oracle's restatement character for character,
    and carries the landmarks of SURVEY Appendix A"""
    g = flame.glsl_source()
    assert g == oracle.glsl
    assert "case -1: {" in g and "default: {" in g and g.count("case ") == 10
    assert "float a = atan(v.x, v.y);\n\t\tfloat r = length(v.xy);" in g  # precalcs in dependency/alphabetical order
    assert "vec2 result = fp[7] *(randf() + randf() + randf() + randf() - 2.0) * sincos(randf() * 2.0 * PI).yx;" in g
    assert "v.x + fp[16] * sin(v.y/(fp[18] * fp[18] + EPS))," in g  # waves reads this xform's affine slots
    assert "return vec4(fp[148] * vec2(fma(fp[142], v.x, fma(fp[144], v.y, fp[146])), fma(fp[143], v.x, fma(fp[145], v.y, fp[147]))), mix(((first_run)? randf(): v.z), fp[149], fp[150]), fp[151]);" in g
    assert "result = vec2(fma(fp[99], result.x, fma(fp[101], result.y, fp[103])), fma(fp[100], result.x, fma(fp[102], result.y, fp[104])));" in g
    assert "if(sum >= ratio) return 8;\n\treturn 9;" in g


def test_cuda_dialect_rewrites(flame):
    c = flame.cuda_source()
    assert "2.0f * PI" in c and ".yx()" in c and "fp[163] == 0)" in c
    assert re.search(r"float _rf0 = randf\(\), _rf1 = randf\(\), _rf2 = randf\(\), _rf3 = randf\(\), _rf4 = randf\(\); vec2 result = rfk_cfp\[7\] \*\(_rf0 \+ _rf1 \+ _rf2 \+ _rf3 - 2.0f\) \* sincos\(_rf4 \* 2.0f \* PI\)\.yx\(\);", c)
    assert "((first_run)? randf(): v.z)" in c  # a conditional draw stays conditional
    assert not re.search(r"(?<![\w.])\d+\.\d+(?![\dfeE])", c.split("namespace rfk_glsl {\n#define randf()")[1].split("#undef randf")[0])
    assert "__constant__ int rfk_affine_slot[11] = {154, 1, 14, 31, 44, 60, 77, 93, 111, 126, 142};" in c


@pytest.mark.parametrize("W,H,hexes", [
    (1280, 720, "c2f6ebbd 4367c677 c367c677 c2f6ebbd 4420f3fb 43a2de80"),
    (3840, 2160, "c3b930ce 442dd4d9 c42dd4d9 c3b930ce 44f16df8 44744dc1"),
    (7680, 4320, "c43930ce 44add4d9 c4add4d9 c43930ce 45716df8 44f44dc1"),
    (15360, 8640, "c4b930ce 452dd4d9 c52dd4d9 c4b930ce 45f16df8 45744dc1")])
def test_screen_space_affine_golden(flame, oracle, oracle_mod, W, H, hexes):
    """SURVEY Appendix B (computed from the reference's flame.hpp:97-128 + flame.cpp:289-296)"""
    got = flame.screen_space_affine(W, H)
    assert ["%08x" % v for v in got.view(np.uint32)] == hexes.split()
    assert np.array_equal(got.view(np.uint32), oracle_mod.screen_space_affine(oracle.flame, W, H).view(np.uint32))


def test_affine_helpers_match_oracle(rfk, oracle_mod):
    rng = np.random.default_rng(1)
    for _ in range(50):
        a = rng.normal(0, 2, 6).astype(np.float32)
        deg, s, t = np.float32(rng.uniform(-720, 720)), np.float32(rng.uniform(0.1, 300)), rng.normal(0, 3, 2).astype(np.float32)
        assert np.array_equal(rfk.rotate_affine(a, deg).view(np.uint32), np.array(oracle_mod.rotate_affine(a, deg), dtype=np.float32).view(np.uint32))
        assert np.array_equal(rfk.scale_affine(a, s).view(np.uint32), np.array(oracle_mod.scale_affine(a, s), dtype=np.float32).view(np.uint32))
        assert np.array_equal(rfk.translate_affine(a, t).view(np.uint32), np.array(oracle_mod.translate_affine(a, t), dtype=np.float32).view(np.uint32))


def test_rotate_xforms_is_the_per_frame_animation_step(rfk, compiler, oracle_mod):
    f = rfk.Flame.load_flame(GENOME, compiler)
    before = [np.array(f.xform(k).affine, dtype=np.float32) for k in range(10)]
    f.rotate_xforms(18.0 / 60.0)  # DEGREES_PER_SECOND * dt, main.cpp:224, :383-395
    for k in range(10):
        want = np.array(oracle_mod.rotate_affine(before[k], np.float32(18.0 / 60.0) * np.float32(f.xform(k).rotation_frequency)), dtype=np.float32)
        assert np.array_equal(np.array(f.xform(k).affine, dtype=np.float32).view(np.uint32), want.view(np.uint32))
    assert f.needs_warmup()


def test_load_flame_failures(rfk, compiler, tmp_path):
    assert rfk.Flame.load_flame(str(tmp_path / "missing.flam3"), compiler) is None
    assert "cannot read" in rfk.Flame.last_error()
    bad = GENOME_TEMPLATE % '<xform weight="1" color="0" linear="1" not_a_variation="2" coefs="1 0 0 1 0 0" opacity="1"/>'
    assert rfk.Flame.load_flame_string(bad, compiler) is None
    assert "Unknown attribute not_a_variation" in rfk.Flame.last_error()  # flame.cpp:196
    assert rfk.Flame.load_flame_string("<flame><xform weight=1/></flame>", compiler) is None  # malformed XML
    assert rfk.Flame.load_flame_string(GENOME_TEMPLATE % "", compiler) is None  # no xforms
    # a variation parameter missing from the XML leaves a bare identifier behind: compile failure
    noparam = GENOME_TEMPLATE % '<xform weight="1" color="0" julian="1" coefs="1 0 0 1 0 0" opacity="1"/>'
    assert rfk.Flame.load_flame_string(noparam, compiler) is None
    assert "julian_power" in rfk.Flame.last_error()
    # one of the variations that do not compile in the reference's table either
    broken = GENOME_TEMPLATE % '<xform weight="1" color="0" oscope="1" coefs="1 0 0 1 0 0" opacity="1"/>'
    assert rfk.Flame.load_flame_string(broken, compiler) is None


def test_field_edits(rfk, compiler):
    f = rfk.Flame.load_flame(GENOME, compiler)
    x = f.xform(3)
    x.color_speed = 0.25
    f.set_xform(3, x)
    assert f.xform(3).color_speed == 0.25 and f.copy_flame_data_to_buffer()[56] == np.float32(0.25)
    f.set_variation(4, "julian", 0.5)
    f.set_param(4, "julian_power", 3.0)
    buf = f.copy_flame_data_to_buffer()
    assert buf[66] == 0.5 and buf[69] == 3.0
    with pytest.raises(rfk.RefraktError):
        f.set_variation(4, "swirl", 1.0)  # structure is fixed after load
    x = f.xform(6)
    x.has_post = 0
    with pytest.raises(rfk.RefraktError):
        f.set_xform(6, x)
    with pytest.raises(rfk.RefraktError):
        f.xform(10)
    i = f.info()
    i.estimator_radius = 500
    f.set_info(i)
    assert f.post_params().estimator_radius == 100  # main.cpp:502
    assert f.post_params().scale_constant == np.float32(1e-4)


def test_shipped_genome_cubin_is_sm100_with_vector_reductions(flame, tmp_path):
    """NVRTC builds the kernels for sm_100a without a GPU; the histogram update is one REDG.E.ADD.F32x4"""
    path = tmp_path / "k.cubin"
    path.write_bytes(flame.cubin())
    elf = subprocess.run(["cuobjdump", "-elf", str(path)], stdout=subprocess.PIPE, text=True).stdout
    assert "sm_100" in elf
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", "rfk_draw", str(path)], stdout=subprocess.PIPE, text=True).stdout
    assert "RED.E.ADD.F32x4" in sass.replace("REDG", "RED") or "REDG.E.ADD.F32x4" in sass
    assert "LDL" not in sass or sass.count("LDL") < 80  # no register spilling in the hot loop
    res = subprocess.run(["cuobjdump", "-res-usage", str(path)], stdout=subprocess.PIPE, text=True).stdout
    m = re.search(r"Function rfk_draw:\s*\n\s*REG:(\d+)", res)
    assert m and int(m.group(1)) <= 64


def test_staged_kernels_build_for_sm100_and_queue_their_samples(rfk, compiler, tmp_path):
    """kernel option staged_bins: rfk_draw appends 8-byte records with the streaming policy (STG.E.EF.64), numbers them with
    a shared-memory atomic per lane (no MATCH.ANY grouping); the direct reduction remains as the overflow path; no spills at
    2048 threads per SM. The option is validated: -1 (automatic, the default), 0, or 8..24 without the modes it excludes."""
    f = rfk.Flame.load_flame(GENOME, compiler)
    assert f.options().staged_bins == -1
    for bad in (dict(staged_bins=7), dict(staged_bins=25), dict(staged_bins=-2), dict(staged_bins=22, deterministic=1),
                dict(staged_bins=22, warp_aggregate=1), dict(staged_bins=22, l2_hints=1)):
        with pytest.raises(rfk.RefraktError):
            f.set_options(**bad)
    f.set_options(staged_bins=-1, deterministic=1)  # the automatic mode just stays off with these
    f.set_options(deterministic=0, staged_bins=22)
    path = tmp_path / "staged.cubin"
    path.write_bytes(f.cubin())
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", "rfk_draw", str(path)], stdout=subprocess.PIPE, text=True).stdout
    assert "STG.E.EF.64" in sass and "ATOMS.ADD" in sass and "MATCH" not in sass
    assert "REDG.E.ADD.F32x4" in sass  # queue exhausted: direct reduction
    res = subprocess.run(["cuobjdump", "-res-usage", str(path)], stdout=subprocess.PIPE, text=True).stdout
    m = re.search(r"Function rfk_draw:\s*\n\s*REG:(\d+) STACK:(\d+)", res)
    assert m and int(m.group(1)) <= 32 and int(m.group(2)) == 0


@pytest.mark.parametrize("chunk", range(6))
def test_every_compile_clean_variation_builds(rfk, compiler, vt, oracle_mod, chunk):
    """all 68 variations of SURVEY Appendix C that compile in the reference also compile here (CUDA dialect, NVRTC,
    sm_100a) and in the oracle (g++), with identical GLSL text"""
    xml = chunk_genome(chunk, vt)
    f = rfk.Flame.load_flame_string(xml, compiler)
    assert f is not None, rfk.Flame.last_error()
    assert len(f.cubin()) > 1000
    of = oracle_mod.load_flame_string(xml, vt)
    assert f.glsl_source() == oracle_mod.compile_flame_xforms(of, vt)
    assert np.array_equal(f.copy_flame_data_to_buffer().view(np.uint32), oracle_mod.copy_flame_data_to_buffer(of).view(np.uint32))
    oracle_mod.Oracle(of, vt)  # builds


def test_gpu_entry_points_fail_loudly_without_a_device(rfk, flame):
    """no CPU fallback: on a box without CUDA the run entry points raise instead of computing elsewhere"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(rfk.RefraktError):
        rfk.set_sim_parameters(256 * 4, 4, 8)
    with pytest.raises(rfk.RefraktError):
        flame.single_step(np.zeros((1, 3), np.float32), np.zeros(1, np.int32), np.zeros((1, 4), np.uint32))
    with pytest.raises(rfk.RefraktError):
        flame.render_frame(64, 64, max_draw_calls=1)


def test_overlay_fixes_the_eight_broken_variations(rfk, overlay_compiler, overlay_vt, oracle_mod):
    """refrakt_b200/data/variations_b200.yaml: with the overlay all 76 names compile on both sides (flam3 semantics;
    parity for these eight is unpinned — the reference's own text does not compile)"""
    assert len(overlay_compiler.variations()) == 76
    assert overlay_compiler.get_parameters_for_variation("oscope") == ["oscope_separation", "oscope_frequency", "oscope_amplitude", "oscope_damping"]
    xml = chunk_genome(99, overlay_vt, names=BROKEN)
    f = rfk.Flame.load_flame_string(xml, overlay_compiler)
    assert f is not None, rfk.Flame.last_error()
    of = oracle_mod.load_flame_string(xml, overlay_vt)
    assert f.glsl_source() == oracle_mod.compile_flame_xforms(of, overlay_vt)
    oracle_mod.Oracle(of, overlay_vt)


def test_stress_genome_builds(rfk, overlay_compiler, overlay_vt, oracle_mod):
    """BASELINE configs[4]: 12 xforms + final xform with divergent variations"""
    xml = stress_genome(overlay_vt)
    f = rfk.Flame.load_flame_string(xml, overlay_compiler)
    assert f is not None, rfk.Flame.last_error()
    i = f.info()
    assert (i.num_xforms, i.has_final_xform) == (12, 1)
    of = oracle_mod.load_flame_string(xml, overlay_vt)
    assert f.glsl_source() == oracle_mod.compile_flame_xforms(of, overlay_vt)
    assert np.array_equal(f.copy_flame_data_to_buffer().view(np.uint32), oracle_mod.copy_flame_data_to_buffer(of).view(np.uint32))


def test_png_writer_roundtrip(rfk, tmp_path):
    """rfk_write_png (the reference's stbi_write_png screenshot, main.cpp:590-593) decodes to the same pixels"""
    from PIL import Image
    rng = np.random.default_rng(0)
    for shape in ((37, 53, 4), (1, 1, 4), (64, 200, 4)):
        img = rng.integers(0, 256, shape, dtype=np.uint8)
        path = str(tmp_path / ("t%d.png" % shape[0]))
        rfk.write_png(path, img)
        assert np.array_equal(np.array(Image.open(path)), img)
    with pytest.raises(rfk.RefraktError):
        rfk.write_png(str(tmp_path / "no_such_dir" / "x.png"), img)


def test_cli_fails_loudly_without_a_device(rfk, tmp_path):
    import torch
    cli = os.path.join(os.path.dirname(rfk.LIB_PATH), "rfk_render")
    assert os.path.exists(cli)
    assert subprocess.run([cli], stdout=subprocess.PIPE, stderr=subprocess.PIPE).returncode == 2  # usage
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([cli, "--genome", GENOME, "--variations", VARIATIONS, "--out", str(tmp_path / "x.png")], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "rfk_render:" in r.stderr and not (tmp_path / "x.png").exists()


def test_buffer_cache_is_the_reference_file_format(rfk, tmp_path):
    """src/buffer_cache.cpp:7-21: <root>/cache/<type>/<group>/<32 hex digits>.bin = size_t byte count + payload;
    buffers are found by listing the directory (cached_buffers), so names only need to be unique"""
    import struct
    g = rfk.BufferGroup(str(tmp_path), "shuffle", "4096")
    assert g.cached_buffers() == []
    perm = np.random.default_rng(0).permutation(4096).astype(np.uint32)
    name = g.write_buffer(perm)
    assert len(name) == 32 and name == name.upper() and all(c in "0123456789ABCDEF" for c in name)
    raw = (tmp_path / "cache" / "shuffle" / "4096" / (name + ".bin")).read_bytes()
    assert struct.unpack("<Q", raw[:8])[0] == perm.nbytes and raw[8:] == perm.tobytes()
    assert np.array_equal(g.read_buffer(name, np.uint32), perm)
    assert g.write_buffer(perm) == name and g.cached_buffers() == [name]  # content-derived name
    try:  # ... and, where the system carries xxHash (this image does), the reference's own: XXH3_128bits, high64 then low64
        import ctypes.util
        import xxhash
        if ctypes.util.find_library("xxhash"):
            assert name == xxhash.xxh3_128_hexdigest(perm.tobytes()).upper()
            assert g.write_buffer(np.zeros(0, dtype=np.uint8)) == xxhash.xxh3_128_hexdigest(b"").upper()
            os.remove(str(tmp_path / "cache" / "shuffle" / "4096" / (xxhash.xxh3_128_hexdigest(b"").upper() + ".bin")))
    except ImportError:
        pass
    other = g.write_buffer(perm[::-1].copy())
    assert sorted(g.cached_buffers()) == sorted([name, other])
    # a file written the way the reference writes it is read back
    ref_style = tmp_path / "cache" / "rand_state" / "8"
    ref_style.mkdir(parents=True)
    states = np.arange(32, dtype=np.uint32)
    (ref_style / "00000000DEADBEEF00000000CAFEF00D.bin").write_bytes(struct.pack("<Q", states.nbytes) + states.tobytes())
    g2 = rfk.BufferGroup(str(tmp_path), "rand_state", "8")
    assert g2.cached_buffers() == ["00000000DEADBEEF00000000CAFEF00D"]
    assert np.array_equal(g2.read_buffer("00000000DEADBEEF00000000CAFEF00D", np.uint32), states)
    with pytest.raises(rfk.RefraktError):
        g2.read_buffer("missing")


def test_yaml_subset_reader_matches_pyyaml_on_edge_cases(rfk, oracle_mod, tmp_path):
    """the product's own YAML reader (csrc/yaml_lite.cpp, standing in for yaml-cpp) against PyYAML (the oracle's reader)
    on the constructs variations.yaml uses: literal blocks with inner indentation and trailing spaces, quoted scalars with
    ': ', flow collections, comments, an entry without a body"""
    text = '''# leading comment
common:
  r: length($v)   # trailing comment
  rsq: $x * $x + $y * ($y)

variations:
  first:
    # a comment between keys
    result: "vec2($x, ($y < 0)? 1: -1)"
  second:
    param:
      second_a: {}
      second_b: {}
    flags: [no_weight_mul]
    src: |-
      float t = $second_a;   
      if(t > 0.0) {
            t = -t;
      }

      t += $second_b;
    result: |-
      vec2(
        t * $x,
        $weight * $y
      )
  empty_one:

  third:
    flags: [pre_xform, other]
    result: 'vec2($x * 0.5, 0.0)'
'''
    path = tmp_path / "v.yaml"
    path.write_text(text)
    c = rfk.FlameCompiler(str(path))
    vt2 = oracle_mod.VariationTable(str(path))
    assert c.variations() == sorted(vt2.vars) == ["empty_one", "first", "second", "third"]
    assert c.get_parameters_for_variation("second") == ["second_a", "second_b"] and c.param_owner("second_b") == "second"
    assert c.is_common("r") and c.is_common("rsq") and not c.is_common("first")
    xml = GENOME_TEMPLATE % ('<xform weight="1" color="0" color_speed="0.5" first="0.5" second="0.25" second_a="1.5" second_b="-2" coefs="1 0 0 1 0 0" opacity="1"/>'
                             '<xform weight="1" color="1" color_speed="0.5" first="1" third="0.125" coefs="0.5 0 0 0.5 0.1 0" opacity="1"/>')
    f = rfk.Flame.load_flame_string(xml, c)
    assert f is not None, rfk.Flame.last_error()
    of = oracle_mod.load_flame_string(xml, vt2)
    assert f.glsl_source() == oracle_mod.compile_flame_xforms(of, vt2)  # literal blocks, trailing spaces and all
    assert "v.xy += fp[" in f.glsl_source()                            # the pre_xform flag of `third`
    assert np.array_equal(f.copy_flame_data_to_buffer().view(np.uint32), oracle_mod.copy_flame_data_to_buffer(of).view(np.uint32))


def test_xml_reader_edge_cases(rfk, compiler, oracle_mod, vt):
    """the product's own XML reader (csrc/xml_lite.cpp, standing in for pugixml): prolog, comments, entities, single quotes,
    nested <edit> history, attribute order, a <flames> wrapper; values parsed like pugixml / std::stof do"""
    body = """<?xml version="1.0"?>
<!-- a comment with <xform> inside -->
<flames name='pack'>
<flame name="a &amp; b" size='640 480' center="0.5 -0.25" scale="  100.5" rotate="1e1" brightness="4.0x" gamma="4" vibrancy="1"
   estimator_radius="9" estimator_curve=".4" unknown_flame_attr="ignored">
   <xform weight=".5" color="0.25" color_speed="0.5" animate="0" linear="1" coefs="1 0 0 1 1e-1 -.5" opacity="1"   />
   <xform weight="1.5" color="1" color_speed="0.5" spherical="0.5" linear="0.5" coefs="0.5 0.1 -0.1 0.5 0 0" post="1 0 0 1 0.25 0" opacity="0.5">
      <motion motion_frequency="1" linear="2"/>
   </xform>
   <color index="0" rgb="255.9 128 0"/>
   <color index="255" rgb="1 2 3"/>
   <edit nick="x"><edit action="nested &lt;stuff&gt;"/></edit>
</flame>
</flames>"""
    f = rfk.Flame.load_flame_string(body, compiler)
    assert f is not None, rfk.Flame.last_error()
    i = f.info()
    assert list(i.size) == [640, 480] and list(i.center) == [0.5, -0.25] and i.scale == 100.5 and i.rotate == 10.0
    assert i.brightness == 4.0 and i.num_xforms == 2 and not i.has_final_xform and i.estimator_curve == np.float32(0.4)
    x0, x1 = f.xform(0), f.xform(1)
    assert x0.rotation_frequency == 0.0 and x1.rotation_frequency == 0.0 and x1.has_post and list(x1.post) == [1, 0, 0, 1, 0.25, 0]
    assert list(x0.affine)[4:] == [np.float32(0.1), -0.5] and x1.opacity == 0.5
    assert f.variations(1) == {"linear": 0.5, "spherical": 0.5}  # std::map order
    pal = f.palette()
    assert pal[0].tolist() == [255 / 256.0, 0.5, 0.0, 1.0] and pal[255].tolist() == [1 / 256.0, 2 / 256.0, 3 / 256.0, 1.0] and pal[7].tolist() == [0, 0, 0, 0]
    of = oracle_mod.load_flame_string(body[body.index("<flames"):], vt)
    assert np.array_equal(f.copy_flame_data_to_buffer().view(np.uint32), oracle_mod.copy_flame_data_to_buffer(of).view(np.uint32))
    assert f.glsl_source() == oracle_mod.compile_flame_xforms(of, vt)
    for bad in ("", "<flame", "<flame><xform coefs='1 0 0 1 0 0' linear='1' weight='1' /></flam>", "<notflame/>",
                GENOME_TEMPLATE % '<xform weight="abc" linear="1" coefs="x y" opacity="1"/>'):
        got = rfk.Flame.load_flame_string(bad, compiler)
        assert got is None or isinstance(got, rfk.Flame)  # never crashes; malformed input is reported
    assert rfk.Flame.load_flame_string("<flame", compiler) is None and rfk.Flame.last_error() != ""
    assert rfk.Flame.load_flame_string(GENOME_TEMPLATE % '<xform weight="1" linear="1" coefs="x y" opacity="1"/>', compiler) is None


def _mutate(text, rng, inserts, n_edits):
    s = list(text)
    for _ in range(n_edits):
        i, op = rng.randrange(len(s)), rng.random()
        if op < 0.35:
            del s[i:i + rng.choice([1, 3, 10, 200])]
        elif op < 0.7:
            s.insert(i, rng.choice(inserts))
        else:
            s[i] = rng.choice(inserts)
    return "".join(s)


def test_genome_parser_survives_mutated_input(rfk, compiler):
    """load_flame returns NULL + a message or a usable flame, never crashes (seeded mutation fuzz of the shipped genome:
    deletions, stray markup, NUL and non-ASCII bytes, out-of-range numbers)"""
    import random
    from conftest import GENOME
    xml, rng = open(GENOME).read(), random.Random(1)
    inserts = ["<", ">", '"', "/", "=", " ", "&", "\x00", "é", "-1e999", "nan", "999999999999999999999", "<xform ", "/>"]
    loaded = rejected = 0
    for _ in range(600):
        f = rfk.Flame.load_flame_string(_mutate(xml, rng, inserts, rng.choice([1, 2, 5, 20])), compiler)
        if f is None:
            assert rfk.Flame.last_error()
            rejected += 1
        else:
            assert f.copy_flame_data_to_buffer().shape == (1024,) and "vec4 dispatch" in f.glsl_source()
            loaded += 1
    assert loaded > 50 and rejected > 50


def test_variation_table_reader_survives_mutated_input(rfk, tmp_path):
    """rfk_compiler_create on a damaged variations.yaml: an error or a table that still compiles the shipped genome's text"""
    import random
    from conftest import GENOME, VARIATIONS
    text, rng = open(VARIATIONS).read(), random.Random(2)
    inserts = [":", "-", "\n", "  ", "|", "$", '"', "'", "#", "\t", "{", "[", "x"]
    ok = bad = 0
    for k in range(120):
        p = tmp_path / ("v%d.yaml" % k)
        p.write_text(_mutate(text, rng, inserts, rng.choice([1, 3, 10])))
        try:
            c = rfk.FlameCompiler(str(p))
        except rfk.RefraktError:
            bad += 1
            continue
        ok += 1
        f = rfk.Flame.load_flame(GENOME, c)
        if f is not None:
            assert "vec4 dispatch" in f.glsl_source() and "rfk_draw" in f.cuda_source()
    assert ok > 10 and bad > 10


@pytest.mark.parametrize("block", [128, 256, 512])
def test_staging_ring_of_four_chunk_slots_never_aliases(block):
    """The protocol of rfk_draw's region queues (csrc/chaos_kernels.cuh), modelled for one region of one CTA: per iteration up
    to `block` samples draw consecutive numbers; the sample with number n mod 512 == 0 opens chunk n / 512 and publishes it in
    slot (n / 512) mod 4 BEFORE the iteration's barrier; every sample reads the slot of its chunk AFTER that barrier, possibly
    while the next iteration's openers are already writing (they only wait for the NEXT barrier). No reader may ever see a
    slot overwritten by a later chunk, for any sequence of per-iteration counts."""
    import random
    rng = random.Random(block)
    CH = 512
    for trial in range(200):
        counts = [rng.choice([0, 1, block // 3, block - 1, block]) if rng.random() < 0.5 else rng.randrange(block + 1) for _ in range(40)]
        slots = [None] * 4
        total = 0
        prev_reads = []                                      # (slot, chunk) pairs the previous iteration's samples still read
        for c in counts:
            numbers = range(total, total + c)
            opened = [n // CH for n in numbers if n % CH == 0]
            for chunk in opened:                             # openers of this iteration run concurrently with the previous readers
                for slot, want in prev_reads:
                    assert slot != chunk % 4 or want == chunk, (block, counts, chunk)
                slots[chunk % 4] = chunk
            reads = sorted({(n // CH % 4, n // CH) for n in numbers})
            for slot, want in reads:                         # after the barrier: every sample finds its own chunk
                assert slots[slot] == want, (block, counts, want)
            prev_reads = reads
            total += c


MOTION_GENOME = GENOME_TEMPLATE % """<xform weight="0.5" color="0.2" color_speed="0.5" animate="1" linear="0.7" julian="0.4" julian_power="3" julian_dist="1.1" coefs="0.8 0.1 -0.2 0.7 0.1 0.05" opacity="1">
  <motion motion_frequency="2" motion_function="sin" julian="0.25" julian_dist="-0.5"/>
  <motion motion_frequency="0.5" motion_function="triangle" color="0.3" weight="0.2"/>
 </xform>
 <xform weight="0.5" color="0.8" color_speed="0.5" animate="0" linear="1" coefs="0.5 0 0 0.5 -0.3 0.2" opacity="1">
  <motion motion_frequency="1" motion_function="hill" linear="-0.4" swirl="9"/>
 </xform>
 <finalxform color="0.5" color_speed="0" linear="1" coefs="1 0 0 1 0 0" opacity="1"><motion motion_frequency="3" motion_function="sin" opacity="-0.5"/></finalxform>"""


def test_motion_elements_are_parsed_and_applied(rfk, compiler, vt, oracle_mod):
    """<motion> children of an xform (motion_info, src/flame.hpp:15-19, :36; the parser src/flame.cpp:199-210 is commented
    out in the reference): parsed as that block intends, evaluated like flam3; product and oracle agree bit for bit"""
    f = rfk.Flame.load_flame_string(MOTION_GENOME, compiler)
    assert f is not None, rfk.Flame.last_error()
    of = oracle_mod.load_flame_string(MOTION_GENOME, vt)
    base = oracle_mod.load_flame_string(MOTION_GENOME, vt)
    assert f.motion(0) == {k: (float(a), b, float(c)) for k, (a, b, c) in of.xforms[0].motion.items()}
    assert f.motion(0)["julian"] == (2.0, "sin", 0.25) and f.motion(0)["weight"] == (0.5, "triangle", np.float32(0.2))
    assert f.motion(1).keys() == {"linear", "swirl"} and f.motion(-1) == {"opacity": (3.0, "sin", -0.5)}
    # a genome without <motion> has none, and the loaded values do not depend on the elements
    assert rfk.Flame.load_flame(GENOME, compiler).motion(0) == {}
    assert np.array_equal(f.copy_flame_data_to_buffer().view(np.uint32), oracle_mod.copy_flame_data_to_buffer(of).view(np.uint32))
    for name in ("sin", "triangle", "hill", "nonsense"):
        for x in (0.0, 0.1, 0.25, 0.5, 0.8, 1.0, -0.3, 7.625):
            got = rfk.lib().rfk_motion_function(name.encode(), x)
            assert np.float32(got) == oracle_mod.motion_function(name, x), (name, x)
    assert rfk.lib().rfk_motion_function(b"triangle", 0.25) == 1.0 and rfk.lib().rfk_motion_function(b"hill", 0.5) == 1.0
    for t in (0.0, 1.0 / 60, 0.125, 0.77, 3.5):
        n = f.apply_motion(t)
        assert n == oracle_mod.apply_motion(of, base, t) == 6  # swirl is not a variation of xform 1: ignored (structure is fixed)
        assert f.needs_warmup()
        assert np.array_equal(f.copy_flame_data_to_buffer().view(np.uint32), oracle_mod.copy_flame_data_to_buffer(of).view(np.uint32)), t
    assert abs(f.variations(0)["julian"] - (0.4 + 0.25 * np.sin(2 * np.pi * 2 * 3.5))) < 1e-6
    f.apply_motion(0.0)  # back to the loaded values: the base is kept aside, motion does not accumulate
    assert np.array_equal(f.copy_flame_data_to_buffer().view(np.uint32), oracle_mod.copy_flame_data_to_buffer(base).view(np.uint32))


def test_value_specialised_build_compiles_without_a_gpu(rfk, compiler, flame, overlay_compiler, overlay_vt):
    """kernel option `specialize`: the translation unit with every rfk_cfp[k] replaced by its value (NVRTC, sm_100a)"""
    src = flame.variant_source(False, True)
    body = src.split("#define randf() rfk_randf(rs)")[1].split("#undef randf")[0]
    assert "rfk_cfp[" not in re.sub(r"//[^\n]*", "", body) and "RFK_DIVC(" not in body  # no constant-bank reads left
    assert "#define RFK_BAKED 1" in src and "#define RFK_HOT_ONLY 1" in src
    assert "float sum = (0x1.44310ep-6f);" in body  # the first cumulative weight of get_xform_id(), exact
    assert "RFK_AFF(-1" not in body and "RFK_AFF(4, x)" in body  # the final xform does not rotate: literals; xform 4 does
    assert "if (rfk_pick == 4) {" in body and body.index("rfk_pick == 4") < body.index("rfk_pick == 5")  # heaviest xform first
    assert len(flame.variant_cubin(False, True)) > 10000
    assert "#define RFK_PAIRS 1" in flame.variant_source(False, 2) and "#define RFK_PAIRS 1" not in src  # two particles per thread
    assert src.count("#define RFK_EXPERIMENT") == 1 and "#define RFK_EXPERIMENT 0" in src  # timing builds only on request (environment)
    assert len(flame.variant_cubin(False, 2)) > 10000
    assert "#define RFK_STAGED_BINS 1" in flame.variant_source(True, False)
    # a genome whose reciprocals include infinities (1 / 0 for unused slots), and a uniform-weight genome (switch, not a chain)
    stress = rfk.Flame.load_flame_string(stress_genome(overlay_vt), overlay_compiler)
    assert stress is not None, rfk.Flame.last_error()
    s2 = stress.variant_source(False, True)
    assert "switch(xform)" in s2 and len(stress.variant_cubin(False, True)) > 10000
    # the table behind the build: the rotating affine coefficients are not part of it (one build per animation)
    f = rfk.Flame.load_flame(GENOME, compiler)
    before = f.variant_source(False, True)
    f.rotate_xforms(0.3)
    assert f.variant_source(False, True) == before
    f.set_variation(0, "julia", 0.5)
    assert f.variant_source(False, True) != before


def test_exr_writer_round_trip(rfk, tmp_path):
    """rfk_write_exr: scanline OpenEXR 2, uncompressed, FLOAT channels A B G R — read back here by the file layout"""
    import struct
    rng = np.random.default_rng(3)
    W, H = 37, 11
    img = rng.random((H, W, 4)).astype(np.float32) * 4 - 1
    path = str(tmp_path / "frame.exr")
    rfk.write_exr(path, img)
    raw = open(path, "rb").read()
    assert raw[:8] == bytes([0x76, 0x2F, 0x31, 0x01, 2, 0, 0, 0])
    pos, attrs = 8, {}
    while raw[pos] != 0:
        name_end = raw.index(b"\0", pos); type_end = raw.index(b"\0", name_end + 1)
        size = struct.unpack_from("<i", raw, type_end + 1)[0]
        attrs[raw[pos:name_end].decode()] = (raw[name_end + 1:type_end].decode(), raw[type_end + 5:type_end + 5 + size])
        pos = type_end + 5 + size
    pos += 1
    assert attrs["compression"] == ("compression", b"\0") and attrs["lineOrder"][1] == b"\0"
    assert struct.unpack("<4i", attrs["dataWindow"][1]) == (0, 0, W - 1, H - 1) == struct.unpack("<4i", attrs["displayWindow"][1])
    names = [c for c in attrs["channels"][1].split(b"\0") if len(c) == 1 and c.isalpha()]
    assert names == [b"A", b"B", b"G", b"R"]
    offsets = struct.unpack_from("<%dQ" % H, raw, pos)
    got = np.zeros_like(img)
    for y, off in enumerate(offsets):
        yy, nbytes = struct.unpack_from("<ii", raw, off)
        assert yy == y and nbytes == 16 * W
        planes = np.frombuffer(raw, dtype="<f4", count=4 * W, offset=off + 8).reshape(4, W)
        got[y, :, 3], got[y, :, 2], got[y, :, 1], got[y, :, 0] = planes
    assert np.array_equal(got.view(np.uint32), img.view(np.uint32))
    assert len(raw) == offsets[-1] + 8 + 16 * W
    with pytest.raises(rfk.RefraktError):
        rfk.write_exr(str(tmp_path / "no_such_dir" / "x.exr"), img)
