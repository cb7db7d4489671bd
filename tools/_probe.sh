timeout 600 python -m pytest tests -x -q -m gpu -k "variation or single_step or stress or parity" 2>&1 | tail -3
timeout 300 python tools/run_configs.py 5 2>&1 | tail -1 | cut -c150-420
timeout 500 ncu --set full --clock-control none --import-source on -k regex:rfk_draw --launch-skip 5 --launch-count 1 -f -o gpurun_out/prof_config5_r01j python tools/run_configs.py 5 > gpurun_out/prof_config5_r01j.log 2>&1
