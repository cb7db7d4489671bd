#!/bin/bash
# full single-GPU pass: all GPU tests, bench lines of every configuration, ncu of the density kernel
tag=${1:-r02e}
out=gpurun_out
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q > $out/pytest_gpu_$tag.log 2>&1
tail -12 $out/pytest_gpu_$tag.log | cut -c1-300
for cfg in 2 1 3 4 5; do
  timeout 900 python bench.py --config $cfg --steps 3 --warmup 3 --no-cpu-baseline > $out/bench_cfg${cfg}_$tag.json 2> $out/bench_cfg${cfg}_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("$out/bench_cfg${cfg}_$tag.json").read().strip().split("\n")[-1])
    r=d.get("roofline",{})
    print("cfg $cfg value %.4g e2e %.4g ms %.3f e2e_ms %.3f draw_ms %s post_ms %s frac %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], r.get("launch_ms"), r.get("post_ms"), r.get("frac")), d.get("detail"))
except Exception as e:
    print("cfg $cfg failed", e); print(open("$out/bench_cfg${cfg}_$tag.err").read()[-1500:])
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:density_tonemap --launch-skip 3 --launch-count 1 -f -o $out/prof_density_$tag \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/prof_density_$tag.log 2>&1
tail -2 $out/prof_density_$tag.log
