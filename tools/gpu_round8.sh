#!/bin/bash
# Eight-GPU evidence: the bench line and the BASELINE configs that shard (gpurun --gpus 8 -- 'bash tools/gpu_round8.sh r01h').
tag=${1:-r01}
out=gpurun_out
mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 8 --steps 3 --warmup 3 > $out/bench8_$tag.json 2> $out/bench8_$tag.err
tail -c 400 $out/bench8_$tag.json
timeout 900 $TR tools/run_configs.py 2 3 4 5 > $out/configs8_$tag.jsonl 2> $out/configs8_$tag.err
cut -c1-300 $out/configs8_$tag.jsonl
