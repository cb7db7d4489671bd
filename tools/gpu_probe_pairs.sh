#!/bin/bash
# exploratory: generic / specialised / specialised with two particles per thread / measured choice, then the tests that touch them
tag=${1:-r02n}
out=gpurun_out
mkdir -p $out
timeout 600 python tools/probe_draw_variants.py > $out/probe_draw_$tag.jsonl 2> $out/probe_draw_$tag.err
cut -c1-330 $out/probe_draw_$tag.jsonl; tail -3 $out/probe_draw_$tag.err
timeout 1200 python -m pytest tests/test_configs_gpu.py tests/test_render_gpu.py tests/test_sharded_gpu.py tests/test_fullsize_gpu.py -m gpu -q > $out/pytest_gpu_$tag.log 2>&1
tail -6 $out/pytest_gpu_$tag.log | cut -c1-300
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
for cfg in 2 5 4; do
timeout 600 python bench.py --config $cfg --steps 3 --warmup 3 --no-cpu-baseline > $out/bench_cfg${cfg}_$tag.json 2> $out/bench_cfg${cfg}_$tag.err
python - <<PY
import json
d=json.loads(open("$out/bench_cfg${cfg}_$tag.json").read().strip().split("\n")[-1])
print("cfg $cfg value %.4g e2e %.4g ms %.3f e2e_ms %.3f draw_ms %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d.get("roofline",{}).get("launch_ms")))
PY
tail -2 $out/bench_cfg${cfg}_$tag.err
done
