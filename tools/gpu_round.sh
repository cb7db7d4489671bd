#!/bin/bash
# One gpurun call that refreshes the round's evidence: GPU tests, the bench line, the ncu launch list of the bench
# command and one `ncu --set full` capture each of rfk_draw and the density + tonemap kernel (outputs: gpurun_out/).
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh r01f'
tag=${1:-r01}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi_$tag.txt 2>&1
if [ -z "$SKIP_TESTS" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_gpu_$tag.log 2>&1
  echo "pytest exit $?" >> $out/pytest_gpu_$tag.log
  tail -3 $out/pytest_gpu_$tag.log
fi
timeout 600 python bench.py --steps 3 --warmup 3 > $out/bench_$tag.json 2> $out/bench_$tag.err
tail -c 600 $out/bench_$tag.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_$tag.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/bench_under_ncu_$tag.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rfk_draw --launch-skip 40 --launch-count 1 -f -o $out/prof_draw_$tag \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/prof_draw_$tag.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:density_tonemap --launch-skip 3 --launch-count 1 -f -o $out/prof_density_$tag \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/prof_density_$tag.log 2>&1
if [ -n "$WITH_CONFIGS" ]; then
  timeout 900 python tools/run_configs.py 1 2 3 5 > $out/configs_$tag.jsonl 2> $out/configs_$tag.err
  timeout 300 python tools/run_configs.py 3 --staged 0 >> $out/configs_$tag.jsonl 2>> $out/configs_$tag.err
  cut -c1-260 $out/configs_$tag.jsonl
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'rfk_draw|stage_accumulate' --launch-skip 2 --launch-count 2 -f -o $out/prof_config3_$tag \
    python tools/run_configs.py 3 --draw-calls 2 > $out/prof_config3_$tag.log 2>&1
fi
ls -la $out | tail -12
