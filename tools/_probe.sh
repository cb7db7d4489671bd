timeout 300 python -m pytest tests/test_render_gpu.py -x -q -m gpu -k "stag" 2>&1 | tail -3
timeout 300 python tools/run_configs.py 3 2>&1 | tail -1 | tee gpurun_out/config3_r01k.json | cut -c150-420
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'rfk_draw|stage_accumulate' --launch-skip 2 --launch-count 2 -f -o gpurun_out/prof_config3_r01k python tools/run_configs.py 3 --draw-calls 2 > gpurun_out/prof_config3_r01k.log 2>&1
bash tools/sanitize.sh 2>&1 | tail -4
