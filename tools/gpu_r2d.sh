#!/bin/bash
tag=${1:-r02d}
out=gpurun_out
mkdir -p $out
timeout 1200 python -m pytest tests/test_sharded_gpu.py tests/test_render_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x > $out/pytest_gpu_$tag.log 2>&1
tail -8 $out/pytest_gpu_$tag.log | cut -c1-300
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $out/bench_$tag.json 2> $out/bench_$tag.err
python - <<PY
import json
d=json.loads(open("$out/bench_$tag.json").read().strip().split("\n")[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"]); print({k:v for k,v in d["roofline"].items() if k.startswith("post")})
PY
tail -3 $out/bench_$tag.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:density_tonemap --launch-skip 3 --launch-count 1 -f -o $out/prof_density_$tag \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/prof_density_$tag.log 2>&1
tail -2 $out/prof_density_$tag.log
