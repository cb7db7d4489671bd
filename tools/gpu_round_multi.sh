#!/bin/bash
# Multi-GPU pass: the two-rank test, the sharded bench line of every configuration (peer-memory path, and the NCCL path for
# the 4K frame), the C++ CLI with one process per GPU.
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu_round_multi.sh 8 r02'
n=${1:-2}
tag=${2:-r02f}
out=gpurun_out
mkdir -p $out
nvidia-smi -L > $out/smi_$tag.txt
nvidia-smi topo -m >> $out/smi_$tag.txt 2>&1
if [ -z "$SKIP_TESTS" ]; then
  timeout 900 python -m pytest tests/test_sharded_gpu.py -m gpu -q -k "two_ranks" > $out/pytest_mgpu_$tag.log 2>&1
  tail -3 $out/pytest_mgpu_$tag.log | cut -c1-400
fi
run_bench() {  # config, extra args
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --config $1 $2 \
    > $out/bench_cfg$1_${n}gpu_$tag.json 2> $out/bench_cfg$1_${n}gpu_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("$out/bench_cfg$1_${n}gpu_$tag.json").read().strip().split("\n")[-1])
    print("cfg $1 N=$n value %.4g e2e %.4g frame_ms %.3f e2e_ms %.3f scaling %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["scaling"]), d.get("detail"), d.get("weak"))
except Exception as e:
    print("cfg $1 failed", e); print(open("$out/bench_cfg$1_${n}gpu_$tag.err").read()[-2500:])
PY
}
what=${WHAT:-"2 nccl 3 4 5 1 cli"}   # which parts to run
for w in $what; do
  case $w in
    2) run_bench 2 "--steps 5 --warmup 3" ;;
    nccl)
      NCCL_DEBUG=INFO RFK_COMM_P2P=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --config 2 --steps 3 --warmup 2 --no-weak \
          > $out/bench_cfg2_${n}gpu_nccl_$tag.json 2> $out/bench_cfg2_${n}gpu_nccl_$tag.err
      tail -c 700 $out/bench_cfg2_${n}gpu_nccl_$tag.json; grep -i "NVLS\|P2P/\|via" $out/bench_cfg2_${n}gpu_nccl_$tag.err | head -5 ;;
    3) run_bench 3 "--steps 2 --warmup 3" ;;
    4) run_bench 4 "--steps 2 --warmup 1" ;;
    5) run_bench 5 "--steps 3 --warmup 3" ;;
    1) run_bench 1 "--steps 5 --warmup 3" ;;
  esac
done
case " $what " in *" cli "*) ;; *) ls -la $out | tail -3; exit 0 ;; esac
# the C++ CLI, one process per GPU
rm -f /tmp/rfk_id
for r in $(seq 1 $((n-1))); do
  ./refrakt_b200/rfk_render --genome refrakt_b200/data/electricsheep.247.11256.flam3 --variations refrakt_b200/data/variations.yaml --width 3840 --height 2160 --quality 2000 \
     --out $out/cli_${n}gpu_%d.png --frames 3 --world $n --rank $r --comm-file /tmp/rfk_id > $out/cli_rank$r_$tag.log 2>&1 &
done
timeout 300 ./refrakt_b200/rfk_render --genome refrakt_b200/data/electricsheep.247.11256.flam3 --variations refrakt_b200/data/variations.yaml --width 3840 --height 2160 --quality 2000 \
     --out $out/cli_${n}gpu_%d.png --frames 3 --world $n --rank 0 --comm-file /tmp/rfk_id > $out/cli_${n}gpu_$tag.jsonl 2> $out/cli_${n}gpu_$tag.err
wait
cat $out/cli_${n}gpu_$tag.jsonl; tail -3 $out/cli_${n}gpu_$tag.err
rm -f $out/cli_${n}gpu_1.png $out/cli_${n}gpu_2.png
ls -la $out | tail -5
