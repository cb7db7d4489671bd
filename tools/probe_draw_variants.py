"""rfk_draw, generic build against the value-specialised one (kernel option `specialize`), on the shipped genome and the
stress genome of BASELINE configs[4] at 4K: ms per 128-pass call, iterations/s, registers, and the pooled L1 distance
between the two builds' histograms (same seeds). One JSON line per case. Exploratory timing, not the bench."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import refrakt_b200 as r

FIX = os.path.join(ROOT, "tests", "fixtures")
P, TS, W, H = 2048 * 1024, 512, 3840, 2160


PASSES = int(os.environ.get("PROBE_PASSES", "128"))  # iterations per draw call


def measure(flame, name, spec, calls=12, pairs=1):
    flame.set_options(specialize=spec, pair_particles=pairs)
    r.set_sim_parameters(P, TS, 1024, seed=0)
    flame.warmup(16, 1.2 / 60)
    bins = torch.zeros(W * H * 4, dtype=torch.float32, device="cuda")
    flame.draw_to_bins(bins.data_ptr(), W * H, W, PASSES)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(calls):
        flame.draw_to_bins_async(bins.data_ptr(), W * H, W, PASSES)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / calls
    binned = flame.binned_total()
    rec = dict(genome=name, specialize=spec, pair_particles=pairs, uses_specialised=flame.uses_specialised(), pair_state=flame.pair_particles_state(), ms_per_call=ms, passes=PASSES, giter_s=P * PASSES / ms / 1e6,
               in_bounds=binned / (P * PASSES * (calls + 1)), **flame.kernel_info("rfk_draw"))
    print(json.dumps(rec), flush=True)
    d = bins.view(H, W, 4)[..., 3].double().view(H // 8, 8, W // 8, 8).sum(dim=(1, 3))
    return (d / d.sum()).cpu().numpy()


def main():
    from conftest import stress_genome

    compiler = r.FlameCompiler(os.path.join(FIX, "variations.yaml"), overlay=r.OVERLAY_YAML)
    shipped = r.Flame.load_flame(os.path.join(FIX, "electricsheep.247.11256.flam3"), compiler)

    class _Params:
        def __init__(self, names): self.param = names

    class _Table:
        def __init__(self, comp): self.vars = {n: _Params(comp.get_parameters_for_variation(n)) for n in comp.variations()}
    stress = r.Flame.load_flame_string(stress_genome(_Table(compiler)), compiler)
    assert shipped is not None and stress is not None, r.Flame.last_error()
    for name, flame in (("electricsheep.247.11256", shipped), ("stress247", stress)):
        if "--pairs-only" in sys.argv:
            measure(flame, name, 1, pairs=2)
            continue
        a = measure(flame, name, 0)
        b = measure(flame, name, 1, pairs=0)
        c = measure(flame, name, 1, pairs=2)
        measure(flame, name, 1, pairs=1)  # measured choice
        print(json.dumps(dict(genome=name, pooled_l1_generic_vs_specialised=float(0.5 * np.abs(a - b).sum()),
                              pooled_l1_single_vs_pairs=float(0.5 * np.abs(b - c).sum()))), flush=True)


if __name__ == "__main__":
    main()
