#!/bin/bash
# Timing experiments on the paired rfk_draw (RFK_EXPERIMENT, chaos_kernels.cuh): what the palette's bank conflicts and the
# reductions cost in the kernel whose L1TEX data pipe is 93 % busy. Then the usual single-GPU checks.
out=gpurun_out
mkdir -p $out
for e in 0 1 2 3 4; do
  echo "experiment $e"
  RFK_EXPERIMENT=$e timeout 300 python tools/probe_draw_variants.py --pairs-only 2>&1 | tail -n 2 | cut -c1-330
done | tee $out/probe_l1tex_r02.log
timeout 1500 python -m pytest tests -m gpu -q > $out/pytest_gpu_r02r.log 2>&1; tail -n 3 $out/pytest_gpu_r02r.log | cut -c1-300
timeout 300 python __graft_entry__.py --smoke > $out/smoke_r02r.log 2>&1; tail -n 2 $out/smoke_r02r.log
timeout 600 python bench.py > $out/bench_r02r.json 2> $out/bench_r02r.err; tail -c 600 $out/bench_r02r.json
timeout 600 python bench.py --config 4 --no-cpu-baseline > $out/bench_config4_r02r.json 2> $out/bench_config4_r02r.err; tail -c 900 $out/bench_config4_r02r.json
timeout 300 ./refrakt_b200/rfk_render --genome refrakt_b200/data/electricsheep.247.11256.flam3 --variations refrakt_b200/data/variations.yaml --width 3840 --height 2160 --quality 2000 \
     --out $out/cli_1gpu_%d.png --frames 3 > $out/cli_1gpu_r02r.jsonl 2> $out/cli_1gpu_r02r.err
cat $out/cli_1gpu_r02r.jsonl | cut -c1-400; rm -f $out/cli_1gpu_1.png $out/cli_1gpu_2.png
