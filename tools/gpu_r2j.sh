#!/bin/bash
tag=${1:-r02j}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_configs_gpu.py tests/test_render_gpu.py -m gpu -q > $out/pytest_gpu_$tag.log 2>&1
grep -n "unexplained outlier" $out/pytest_gpu_$tag.log | head -8 | cut -c1-300
tail -5 $out/pytest_gpu_$tag.log | cut -c1-300
nproc
( time python bench.py ) > $out/bench_default_$tag.json 2> $out/bench_default_$tag.err
python - <<PY
import json
d=json.loads(open("$out/bench_default_$tag.json").read().strip().split("\n")[-1])
r=d["roofline"]
print("default: value %.4g e2e %.4g ms %.3f e2e_ms %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"]["ms_per_step"]))
print({k:r[k] for k in ("bound","achieved","peak","frac","launch_ms","issue_frac","l2_red_frac","dram_frac","post_ms","post_frac","warp_inst_per_iteration") if k in r})
print(d.get("cpu_baseline"))
PY
tail -4 $out/bench_default_$tag.err
( time python bench.py --impl reference --steps 2 --warmup 1 ) > $out/bench_reference_$tag.json 2> $out/bench_reference_$tag.err
cat $out/bench_reference_$tag.json | cut -c1-1800; tail -4 $out/bench_reference_$tag.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rfk_draw --launch-skip 300 --launch-count 1 -f -o $out/prof_draw_$tag \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/prof_draw_$tag.log 2>&1
tail -1 $out/prof_draw_$tag.log
