mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_render_gpu.py -x -q -m gpu -k staged 2>&1 | tail -15 | tee gpurun_out/staged_test.log
timeout 200 python tools/run_configs.py 3 --staged 21 --draw-calls 32 2>&1 | tail -2 | tee gpurun_out/staged_cfg3.json
timeout 200 python tools/run_configs.py 3 --staged 20 --draw-calls 32 2>&1 | tail -2 | tee -a gpurun_out/staged_cfg3.json
timeout 200 python tools/run_configs.py 3 --draw-calls 32 2>&1 | tail -2 | tee -a gpurun_out/staged_cfg3.json
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v10.json 2>gpurun_out/bench_v10.err; tail -c 300 gpurun_out/bench_v10.json
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_atom.sum,lts__t_sectors_op_red.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:'rfk_draw|stage_acc' -c 4 --csv --log-file gpurun_out/staged_launches2.csv python tools/run_configs.py 3 --staged 21 --draw-calls 2 > gpurun_out/staged_ncu2.log 2>&1
