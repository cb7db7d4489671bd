"""Writes tests/golden/reference_host_*.json.gz: outputs of the REFERENCE'S OWN host code (src/flame.cpp,
src/variation_table.cpp, src/util.cpp compiled from /root/reference by oracle/Makefile into oracle/_ref/libref_host.so,
see oracle/ref_host.cpp for what is and is not the reference's code) for the shipped genome and six synthetic genomes
that cover every compile-clean variation: parsed fields, buffer map, the fp[] upload, the generated dispatch() text (and
the digest of the complete iterate shader handed to glShaderSource), the screen-space affine. Run in the build container (needs
/root/reference): python tests/golden/make_reference_golden.py. The GPU box has no reference; it uses these fixtures."""
import ctypes
import gzip
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def describe(lib, genome_path, W, H):
    buf = ctypes.create_string_buffer(1 << 23)
    cwd = os.getcwd()
    os.chdir(REF)  # the reference reads variations.yaml and shaders/ relative to the working directory
    try:
        n = lib.ref_host_describe(genome_path.encode(), ctypes.c_ulong(W), ctypes.c_ulong(H), buf, len(buf))
    finally:
        os.chdir(cwd)
    assert n > 0
    return json.loads(buf.value.decode())


def main():
    import refrakt_oracle as ro
    from conftest import GENOME, VARIATIONS, chunk_genome
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_host.so"))
    lib.ref_host_describe.restype = ctypes.c_long
    vt = ro.VariationTable(VARIATIONS)
    cases = {"electricsheep": open(GENOME).read()}
    for chunk in range(6):
        cases["chunk%d" % chunk] = chunk_genome(chunk, vt)
    cases["bad_attribute"] = cases["chunk0"].replace('opacity="1"/>', 'opacity="1" nonsense="3"/>', 1)
    for name, xml in cases.items():
        with tempfile.NamedTemporaryFile("w", suffix=".flam3", delete=False) as fh:
            fh.write(xml)
            path = fh.name
        d = describe(lib, path, 1280, 720)
        os.unlink(path)
        if "iterate_shader" in d:  # the reference's shader text itself stays out of the repo: its digest is enough to detect drift
            import hashlib
            d["iterate_shader_sha256"] = hashlib.sha256(d["iterate_shader"].encode()).hexdigest()
            d["iterate_shader_contains_generated_text"] = d["compile_flame_xforms"] in d["iterate_shader"]
            del d["iterate_shader"]
        d["genome_xml"] = xml
        d["ss_affine_dims"] = [1280, 720]
        with gzip.open(os.path.join(HERE, "reference_host_%s.json.gz" % name), "wt") as out:
            json.dump(d, out, sort_keys=True)
        print(name, "loaded" if d["loaded"] else "rejected", d.get("iterate_shader_contains_generated_text"))


if __name__ == "__main__":
    main()
