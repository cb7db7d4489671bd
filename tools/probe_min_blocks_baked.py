"""Register budget of the value-specialised single-particle rfk_draw (kernel option min_blocks: 8 = 2048 resident threads / 32
registers ... 5 = 1280 threads / 48 registers) on the stress genome and the shipped genome at 4K: ms per 128-pass call.
Exploratory timing, not the bench. Result of round 2: 8 is the fastest on both (profiles/r02_probe_min_blocks_baked.jsonl)."""
import json, os, sys
ROOT="/root/repo" if os.path.exists("/root/repo/refrakt_b200") else os.getcwd()
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, refrakt_b200 as r
from conftest import stress_genome
FIX=os.path.join(ROOT,"tests","fixtures")
P,TS,W,H=2048*1024,512,3840,2160
compiler=r.FlameCompiler(os.path.join(FIX,"variations.yaml"), overlay=r.OVERLAY_YAML)
class _P:
    def __init__(s,n): s.param=n
class _T:
    def __init__(s,c): s.vars={n:_P(c.get_parameters_for_variation(n)) for n in c.variations()}
stress=r.Flame.load_flame_string(stress_genome(_T(compiler)), compiler)
shipped=r.Flame.load_flame(os.path.join(FIX,"electricsheep.247.11256.flam3"), compiler)
for name,fl,pairs in (("stress",stress,0),("shipped",shipped,0)):
    for mb in (8,7,6,5):
        fl.set_options(specialize=1, pair_particles=pairs, min_blocks=mb)
        r.set_sim_parameters(P,TS,1024,seed=0)
        fl.warmup(16,1.2/60)
        bins=torch.zeros(W*H*4,dtype=torch.float32,device="cuda")
        fl.draw_to_bins(bins.data_ptr(),W*H,W,128); torch.cuda.synchronize()
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(12): fl.draw_to_bins_async(bins.data_ptr(),W*H,W,128)
        e1.record(); torch.cuda.synchronize()
        print(json.dumps(dict(genome=name,min_blocks=mb,ms=e0.elapsed_time(e1)/12)),flush=True)
