"""Throughput and register count of rfk_draw under __launch_bounds__(256, n) for several genomes (one GPU)."""
import os, sys, json
ROOT = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, numpy as np
import refrakt_b200 as r
from conftest import GENOME, VARIATIONS, chunk_genome, stress_genome, COMPILE_CLEAN
c = r.FlameCompiler(VARIATIONS, overlay=r.OVERLAY_YAML)
class VT:  # chunk_genome needs parameter names per variation
    def __init__(self): self.vars = {n: type("V", (), {"param": c.get_parameters_for_variation(n)})() for n in c.variations()}
vt = VT()
P, TS, W, H = 2048*1024, 512, 3840, 2160
r.set_sim_parameters(P, TS, 64)
bins = torch.zeros(W*H*4, dtype=torch.float32, device="cuda")
cases = {"shipped": open(GENOME).read(), "stress": stress_genome(vt)}
for k in range(6): cases["chunk%d" % k] = chunk_genome(k, vt)
for name, xml in cases.items():
    f = r.Flame.load_flame_string(xml, c)
    assert f is not None, (name, r.Flame.last_error())
    row = {"genome": name}
    for mb in (0, 6, 8):
        f.set_options(min_blocks=mb)
        f.warmup(4, 0.02)
        f.draw_to_bins(bins.data_ptr(), W*H, W, 32)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): f.draw_to_bins_async(bins.data_ptr(), W*H, W, 64)
        e1.record(); torch.cuda.synchronize()
        info = f.kernel_info()
        row["mb%d" % mb] = round(3*64*P/(e0.elapsed_time(e1)*1e-3)/1e9, 1)
        row["regs%d" % mb] = info["regs"]
    print(json.dumps(row), flush=True)
