"""Generates the committed golden fixtures from the oracle (run once in the build container:
python tests/golden/make_golden.py). The reference itself cannot run here, so these vectors
pin the ORACLE against drift; they are not outputs of the reference."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import refrakt_oracle as ro

FIX = os.path.join(ROOT, "tests", "fixtures")
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    vt = ro.VariationTable(os.path.join(FIX, "variations.yaml"))
    orc = ro.Oracle(ro.load_flame(os.path.join(FIX, "electricsheep.247.11256.flam3"), vt), vt)
    rng = np.random.default_rng(20261017)
    n = 2200
    xyz = np.concatenate([rng.normal(0, 1, (n, 2)), rng.random((n, 1))], axis=1).astype(np.float32)
    xid = np.repeat(np.arange(-1, 10), n // 11).astype(np.int32)
    states = rng.integers(0, 2**32, (n, 4), dtype=np.uint64).astype(np.uint32)
    out, rng_out = orc.single_step(xyz, xid, states)
    np.savez_compressed(os.path.join(HERE, "single_step_electricsheep.npz"), xyz=xyz, xid=xid, rng_in=states, out=out, rng_out=rng_out)

    H, W = 48, 64
    bins = np.zeros((H, W, 4), dtype=np.float32)
    d = (rng.random((H, W)) < 0.4) * rng.integers(1, 300, (H, W))
    bins[..., 3] = d
    bins[..., :3] = rng.random((H, W, 3)) * d[..., None]
    de = orc.density_estimate(bins, W, H, 9, 0, 0.5)
    np.savez_compressed(os.path.join(HERE, "density_tonemap_small.npz"), bins=bins, radius=9, min=0, curve=0.5, de=de, tonemapped=orc.tonemap(de, scale_constant=1e-4))


if __name__ == "__main__":
    main()
