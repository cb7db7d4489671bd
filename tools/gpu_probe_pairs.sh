#!/bin/bash
tag=${1:-r02o}
out=gpurun_out
mkdir -p $out
timeout 1200 python -m pytest tests/test_configs_gpu.py tests/test_render_gpu.py tests/test_sharded_gpu.py -m gpu -q > $out/pytest_gpu_$tag.log 2>&1
tail -6 $out/pytest_gpu_$tag.log | cut -c1-300
for cfg in 2 4 1; do
timeout 600 python bench.py --config $cfg --steps 3 --warmup 3 --no-cpu-baseline > $out/bench_cfg${cfg}_$tag.json 2> $out/bench_cfg${cfg}_$tag.err
python - <<PY
import json
d=json.loads(open("$out/bench_cfg${cfg}_$tag.json").read().strip().split("\n")[-1])
r=d.get("roofline",{})
print("cfg $cfg value %.4g e2e %.4g ms %.3f e2e_ms %.3f draw_ms %s post_ms %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], r.get("launch_ms"), r.get("post_ms")), d.get("detail"))
PY
tail -2 $out/bench_cfg${cfg}_$tag.err
done
