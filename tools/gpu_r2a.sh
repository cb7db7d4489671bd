#!/bin/bash
# round 2, first GPU call: the value-specialised rfk_draw (tests that touch it, probe, bench line, ncu opcode capture)
tag=${1:-r02a}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi_$tag.txt 2>&1
timeout 600 python tools/probe_draw_variants.py > $out/probe_draw_$tag.jsonl 2> $out/probe_draw_$tag.err
cat $out/probe_draw_$tag.jsonl; tail -3 $out/probe_draw_$tag.err
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_render_gpu.py -m gpu -x -q > $out/pytest_gpu_$tag.log 2>&1
tail -5 $out/pytest_gpu_$tag.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $out/bench_$tag.json 2> $out/bench_$tag.err
tail -c 1500 $out/bench_$tag.json; tail -3 $out/bench_$tag.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rfk_draw --launch-skip 300 --launch-count 1 -f -o $out/prof_draw_$tag \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/prof_draw_$tag.log 2>&1
tail -2 $out/prof_draw_$tag.log
ls -la $out | tail -8
