#!/bin/bash
# One gpurun call that refreshes the round's single-GPU evidence (outputs: gpurun_out/): all GPU tests, smoke(), the default
# bench line and one line per BASELINE configuration, the CPU arm, the ncu launch list of the bench command and one
# `ncu --set full` capture each of rfk_draw and the density + tonemap kernel.
#   gpurun --timeout 2400 -- 'bash tools/gpu_round.sh r02'
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi_$tag.txt 2>&1
if [ -z "$SKIP_TESTS" ]; then
  timeout 1500 python -m pytest tests -m gpu -q > $out/pytest_gpu_$tag.log 2>&1
  tail -4 $out/pytest_gpu_$tag.log | cut -c1-300
  timeout 300 python __graft_entry__.py --smoke > $out/smoke_$tag.log 2>&1; tail -2 $out/smoke_$tag.log
fi
timeout 900 python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err
tail -c 400 $out/bench_$tag.json; tail -2 $out/bench_$tag.err
for cfg in 1 3 4 5; do
  timeout 900 python bench.py --config $cfg --no-cpu-baseline > $out/bench_config${cfg}_$tag.json 2> $out/bench_config${cfg}_$tag.err
  python -c "
import json; d=json.loads(open('$out/bench_config${cfg}_$tag.json').read().strip().split('\n')[-1]); print('config $cfg: value %.4g e2e %.4g ms_per_step %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']), d.get('detail'))" || tail -5 $out/bench_config${cfg}_$tag.err
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_reference_$tag.json 2> $out/bench_reference_$tag.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_$tag.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/bench_under_ncu_$tag.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rfk_draw --launch-skip 300 --launch-count 1 -f -o $out/prof_draw_$tag \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/prof_draw_$tag.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:density_tonemap --launch-skip 3 --launch-count 1 -f -o $out/prof_density_$tag \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/prof_density_$tag.log 2>&1
ls -la $out | tail -6
