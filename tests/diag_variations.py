"""Per-variation single-step parity table (GPU vs oracle), one single-variation genome per name. A test-side
diagnostic (it drives the oracle as the checker), kept under tests/ for that reason.
Usage: python tests/diag_variations.py [math_mode ...]  -> gpurun_out/variation_parity.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np

import refrakt_b200 as r
import refrakt_oracle as ro
from conftest import BROKEN, COMPILE_CLEAN, GENOME_TEMPLATE, VARIATIONS, xform_xml


def rel_err(got, want):
    d = np.linalg.norm(got[:, :2].astype(np.float64) - want[:, :2].astype(np.float64), axis=1)
    return d / np.maximum(1.0, np.linalg.norm(want[:, :2].astype(np.float64), axis=1))


def main():
    modes = [int(a) for a in sys.argv[1:]] or [1]
    compiler = r.FlameCompiler(VARIATIONS, overlay=r.OVERLAY_YAML)
    vt = ro.VariationTable(VARIATIONS, overlay=r.OVERLAY_YAML)
    n = 50000
    rows = []
    for name in COMPILE_CLEAN + BROKEN:
        rng = np.random.default_rng(abs(hash(name)) % 1000)
        names = [name] if "pre_xform" not in vt.vars[name].flags else ["linear", name]  # a pre_xform variation alone declares no result
        xml = GENOME_TEMPLATE % xform_xml(names, vt, np.random.default_rng(len(name)))
        of = ro.load_flame_string(xml, vt)
        orc = ro.Oracle(of, vt)
        f = r.Flame.load_flame_string(xml, compiler)
        xyz = np.concatenate([rng.normal(0, 0.8, (n, 2)), rng.random((n, 1))], axis=1).astype(np.float32)
        states = rng.integers(0, 2**32, (n, 4), dtype=np.uint64).astype(np.uint32)
        ids = np.zeros(n, dtype=np.int32)
        want, want_rng = orc.single_step(xyz, ids, states)
        sane = np.isfinite(want).all(axis=1) & (np.abs(want[:, :2]).max(axis=1) < 1e4)
        rec = {"variation": name, "sane_fraction": float(sane.mean())}
        for mode in modes:
            f.set_options(math_mode=mode)
            got, got_rng = f.single_step(xyz, ids, states)
            err = rel_err(got, want)
            bad = sane & ~(err <= 1e-5)
            small = sane & (np.abs(want[:, :2]).max(axis=1) < 10)
            rec["mode%d" % mode] = {"rng_exact": bool(np.array_equal(got_rng, want_rng)), "frac_bad": float(bad.mean()),
                                    "p999_err": float(np.quantile(err[sane], 0.999)) if sane.any() else None,
                                    "max_err_small_outputs": float(err[small].max()) if small.any() else None}
        rows.append(rec)
        print(json.dumps(rec), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "variation_parity.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
