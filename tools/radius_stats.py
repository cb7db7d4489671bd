"""Distribution of density-estimation radii over a rendered histogram (which radii dominate the DE kernel's work)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import refrakt_b200 as r
FIX = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "fixtures")
c = r.FlameCompiler(os.path.join(FIX, "variations.yaml"))
f = r.Flame.load_flame(os.path.join(FIX, "electricsheep.247.11256.flam3"), c)
r.set_sim_parameters(2048 * 1024, 512, 1024)
for (W, H, spp) in ((3840, 2160, 2000), (1280, 720, 120), (3840, 2160, 100)):
    bins = torch.zeros(W * H * 4, dtype=torch.float32, device="cuda")
    f.warmup(16, 1.2 / 60)
    while f.binned_total() < spp * W * H:
        f.draw_to_bins_async(bins.data_ptr(), W * H, W, 128)
    d = bins.view(-1, 4)[:, 3]
    nz = d > 0
    rad = torch.clamp((11.0 / d[nz].pow(0.6)).to(torch.int32), max=11)
    cnt = torch.bincount(rad, minlength=12).tolist()
    taps = sum(cnt[k] * (2 * k + 1) ** 2 for k in range(1, 12))
    print(json.dumps(dict(W=W, H=H, spp=spp, nonzero_fraction=float(nz.float().mean()), radius_counts=cnt, candidates=sum(cnt[1:]), square_taps=taps,
                          taps_by_radius=[cnt[k] * (2 * k + 1) ** 2 for k in range(12)])))
