#!/bin/bash
tag=${1:-r02i}
out=gpurun_out
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q > $out/pytest_gpu_$tag.log 2>&1
grep -n "unexplained outlier" $out/pytest_gpu_$tag.log | head -12 | cut -c1-400
tail -6 $out/pytest_gpu_$tag.log | cut -c1-300
one() {  # label, env, config
  env $2 timeout 600 python bench.py --config $3 --steps 3 --warmup 3 --no-cpu-baseline > $out/bench_$1_$tag.json 2> $out/bench_$1_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("$out/bench_$1_$tag.json").read().strip().split("\n")[-1])
    r=d.get("roofline",{})
    print("$1: value %.4g e2e %.4g ms %.3f e2e_ms %.3f draw_ms %s post_ms %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], r.get("launch_ms"), r.get("post_ms")), d.get("detail"))
except Exception as e:
    print("$1 failed", e); print(open("$out/bench_$1_$tag.err").read()[-1500:])
PY
}
one chain "RFK_DISPATCH=chain" 2
one switch "RFK_DISPATCH=switch" 2
one de6 "RFK_DE_MIN_BLOCKS=6" 2
one cfg4 "A=1" 4
one cfg1 "A=1" 1
one cfg5 "A=1" 5
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_$tag.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/bench_under_ncu_$tag.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rfk_draw --launch-skip 300 --launch-count 1 -f -o $out/prof_draw_$tag \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/prof_draw_$tag.log 2>&1
tail -2 $out/prof_draw_$tag.log
