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
ls -la $out | tail -12
