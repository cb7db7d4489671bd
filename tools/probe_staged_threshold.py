"""Direct reductions against the region queues (kernel option staged_bins) at histogram sizes between 300 MB and 1 GB:
where the automatic mode should switch (profiles/r01_staged_threshold_probe.json). Run on a B200: python tools/probe_staged_threshold.py"""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import refrakt_b200 as r
import run_configs as rc
compiler = r.FlameCompiler(rc.VARIATIONS)
flame = r.Flame.load_flame(rc.GENOME, compiler)
r.set_sim_parameters(rc.P, rc.TS, 1024, seed=0)
for (W, H) in ((7680, 4320), (5760, 3240), (10240, 5760)):
    for staged in (0, 22, 21):
        flame.set_options(staged_bins=staged)
        rc.render_still(flame, W, H, draw_calls=1)
        res = rc.render_still(flame, W, H, draw_calls=32)
        print(json.dumps({"W": W, "H": H, "MB": W * H * 16 / 1e6, "staged": staged, "ms_per_call": res["ms_draw"] / 32}), flush=True)
