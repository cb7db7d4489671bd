"""Writes the fixtures that come from the REFERENCE'S OWN code, run in the build container (needs /root/reference):

    python tests/golden/make_reference_golden.py

oracle/Makefile compiles the reference's host sources (src/flame.cpp, src/variation_table.cpp, src/util.cpp,
src/hammersley.cpp, src/shuffle_buffers.cpp) unmodified into oracle/_ref/libref_host.so against stand-in third-party
headers and a software GL (oracle/softgl/) that executes the reference's GLSL text on the CPU; see oracle/ref_host.cpp for
what is and is not the reference's code. Two families of fixtures, for the shipped genome and six synthetic genomes that
together use every variation the reference's table compiles:

  reference_host_<name>.json.gz   load_flame only: parsed fields, buffer map, the fp[] upload, the generated
                                  get_xform_id() / dispatch() text, digest of the complete iterate shader, screen-space affine
  reference_histogram_electricsheep.npz   (only with --histogram, ~2 minutes) two independent runs of warmup(16) + 64 draw passes
                                  with 131072 particles into a 320 x 180 histogram: density channel, colour sums, binned counts
  reference_device_<name>.npz     set_sim_parameters + load_flame + warmup(2) + draw_to_bins(3 passes) + the density /
                                  tonemap sequence of main.cpp:490-535: the shuffle tables and per-pass shuffle ids the
                                  reference drew, RNG states, particle buffers, fp_inflated, bins, binned counter, density
                                  and tonemapped images

The reference's shader text itself is not stored (digest only). The GPU box has no reference; it uses these files."""
import gzip
import hashlib
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "oracle", "softgl"))
HERE = os.path.dirname(os.path.abspath(__file__))

P, TS, NSHUF, WARMUP, DRAW, TSS = 8192, 32, 8, 2, 3, 1.2 / 60
DIMS = {"electricsheep": (96, 56)}  # W, H; the rest 64 x 40 (tonemap.glsl runs W/8 x H/8 groups: multiples of 8)


def digest(a):
    """sha256 of the bit patterns, every NaN mapped to one pattern (payloads are not part of the contract)"""
    a = np.ascontiguousarray(a)
    if a.dtype == np.float32:
        a = np.where(np.isnan(a), np.float32(np.nan), a).astype(np.float32)
    return hashlib.sha256(a.tobytes()).hexdigest()


def with_full_palette(xml, shipped_xml):
    """the synthetic genomes define 3 palette rows; the reference leaves the other 253 uninitialised (garbage colours).
    For the device fixtures they get the shipped genome's 256 rows instead."""
    rows = [l for l in shipped_xml.split("\n") if "<color " in l]
    body = [l for l in xml.split("\n") if "<color " not in l]
    return "\n".join(body[:-1] + rows + body[-1:])


def host_fixture(h, name, xml):
    from glsl_to_cpp import expand_inja
    with tempfile.NamedTemporaryFile("w", suffix=".flam3", delete=False) as fh:
        fh.write(xml)
        path = fh.name
    d = h.describe(path, 1280, 720)
    os.unlink(path)
    if d["loaded"]:
        # inja's part (get_xform_id) is evaluated by the subset renderer of oracle/softgl/glsl_to_cpp.py
        d["compile_flame_xforms"] = expand_inja(d["compile_flame_xforms"]).replace("\r", "")
        shader = expand_inja(d.pop("iterate_shader")).replace("\r", "")
        d["iterate_shader_sha256"] = hashlib.sha256(shader.encode()).hexdigest()
        d["iterate_shader_contains_generated_text"] = d["compile_flame_xforms"] in shader
        d.pop("animate_shader", None)
    d["genome_xml"] = xml
    d["ss_affine_dims"] = [1280, 720]
    with gzip.open(os.path.join(HERE, "reference_host_%s.json.gz" % name), "wt") as out:
        json.dump(d, out, sort_keys=True)
    print("host  ", name, "loaded" if d["loaded"] else "rejected", d.get("iterate_shader_contains_generated_text"))


def device_fixture(h, name, xml, full):
    import ref_host
    W, H = DIMS.get(name, (64, 40))
    with tempfile.NamedTemporaryFile("w", suffix=".flam3", delete=False) as fh:
        fh.write(xml)
        path = fh.name
    r0 = h.run(path, P, TS, NSHUF, WARMUP, TSS, W, H, 0)
    assert r0["loaded"], name
    out = {"shuffle": h.buffer("shuffle", np.uint32, (NSHUF, P // TS)), "samples": h.buffer("samples", np.float32, (P // TS, 4)),
           "ids_warmup": np.stack([r0["shuf_buf_idx_in"], r0["shuf_buf_idx_out"]], 1).astype(np.uint32),
           "fp_inflated": h.buffer("fp_inflated", np.float32, (TS, -1))}
    rng_w, part_w = h.buffer("rand_states", np.uint32, (P, 4)), h.buffer("particles", np.float32, (P, 4))
    r1 = h.draw(W, H, DRAW)
    out["ids_draw"] = np.stack([r1["shuf_buf_idx_in"], r1["shuf_buf_idx_out"]], 1).astype(np.uint32)
    out["ss_affine"] = np.array([int(v, 16) for v in r1["ss_affine"]], dtype=np.uint32).view(np.float32)
    out["bins"] = h.buffer("bins", np.float32, (H, W, 4))
    rng_d, part_d = h.buffer("rand_states", np.uint32, (P, 4)), h.buffer("particles", np.float32, (P, 4))
    from refrakt_oracle import VariationTable, load_flame
    from conftest import VARIATIONS
    fl = load_flame(path, VariationTable(VARIATIONS))
    os.unlink(path)
    de, tm = h.post(fl.estimator_radius, fl.estimator_min, fl.estimator_curve, fl.gamma, fl.brightness, fl.vibrancy, 4.0)
    out["density"], out["tonemapped"] = de, tm
    meta = {"P": P, "TS": TS, "NSHUF": NSHUF, "warmup_passes": WARMUP, "draw_passes": DRAW, "tss_width": TSS, "W": W, "H": H,
            "binned": r1["binned"], "genome_xml": xml, "libm_probe": ref_host.libm_probe(),
            "sha256": {"rng_after_warmup": digest(rng_w), "particles_after_warmup": digest(part_w), "rng_after_draw": digest(rng_d),
                       "particles_after_draw": digest(part_d), "bins": digest(out["bins"]), "fp_inflated": digest(out["fp_inflated"])}}
    if full:
        out.update(rng_after_warmup=rng_w, particles_after_warmup=part_w, rng_after_draw=rng_d, particles_after_draw=part_d)
    else:  # a slice is enough to see WHERE a mismatch starts; the digests cover the rest
        out.update(rng_after_draw_head=rng_d[:512], particles_after_warmup_head=part_w[:512], particles_after_draw_head=part_d[:512])
    out["meta"] = np.frombuffer(json.dumps(meta, sort_keys=True).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "reference_device_%s.npz" % name), **out)
    print("device", name, "binned", r1["binned"], "finite particles %.4f" % np.isfinite(part_d).all(axis=1).mean())


def histogram_fixture(h, genome_path):
    """Two independent full-length runs (the reference reseeds tables and ids from the clock / random_device every time) at
    the configuration of tests/test_render_gpu.py::oracle_hist: what the production kernel is compared with statistically."""
    W, H, Ph, TSh, passes = 320, 180, 256 * 16 * 32, 32, 64
    runs = []
    for k in range(2):
        r = h.run(genome_path, Ph, TSh, 64, 16, TSS, W, H, passes)
        bins = h.buffer("bins", np.float32, (H, W, 4))
        runs.append((bins, r["binned"]))
        print("histogram run", k, "binned", r["binned"])
    (b1, n1), (b2, n2) = runs
    assert b1[..., 3].max() < 65536 and np.array_equal(b1[..., 3], np.rint(b1[..., 3]))
    np.savez_compressed(os.path.join(HERE, "reference_histogram_electricsheep.npz"),
                        density=b1[..., 3].astype(np.uint16), rgb_sum=b1[..., :3].sum(axis=(0, 1), dtype=np.float64), binned=np.int64(n1),
                        density_second_run=b2[..., 3].astype(np.uint16), binned_second_run=np.int64(n2),
                        config=np.array([W, H, Ph, TSh, passes, 16], dtype=np.int64))


def main():
    import ref_host
    import refrakt_oracle as ro
    from conftest import GENOME, VARIATIONS, chunk_genome
    assert ref_host.available(), "needs /root/reference and `make -C oracle`"
    h = ref_host.ReferenceHost()
    vt = ro.VariationTable(VARIATIONS)
    shipped = open(GENOME).read()
    cases = {"electricsheep": shipped}
    for chunk in range(6):
        cases["chunk%d" % chunk] = chunk_genome(chunk, vt)
    for name, xml in cases.items():
        host_fixture(h, name, xml)
        device_fixture(h, name, xml if name == "electricsheep" else with_full_palette(xml, shipped), full=(name == "electricsheep"))
    if "--histogram" in sys.argv:
        histogram_fixture(h, GENOME)
    host_fixture(h, "bad_attribute", cases["chunk0"].replace('opacity="1"/>', 'opacity="1" nonsense="3"/>', 1))


if __name__ == "__main__":
    main()
