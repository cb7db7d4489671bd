#!/bin/bash
n=${1:-8}
tag=${2:-r02h}
out=gpurun_out
mkdir -p $out
nvidia-smi -L > $out/smi_$tag.txt
run_bench() {
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --config $1 $2 \
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
run_bench 2 "--steps 5 --warmup 3"
RFK_COMM_P2P=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --config 2 --steps 5 --warmup 3 --no-weak \
    > $out/bench_cfg2_${n}gpu_nccl_$tag.json 2> $out/bench_cfg2_${n}gpu_nccl_$tag.err
python -c "
import json; d=json.loads(open('$out/bench_cfg2_${n}gpu_nccl_$tag.json').read().strip().split('\n')[-1]); print('NCCL path: frame_ms %.3f e2e_ms %.3f' % (d['ms_per_step'], d['e2e']['ms_per_step']), d['detail'])"
run_bench 3 "--steps 2 --warmup 3"
run_bench 4 "--steps 2 --warmup 1"
run_bench 5 "--steps 3 --warmup 3"
run_bench 1 "--steps 5 --warmup 3"
