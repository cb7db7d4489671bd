#!/bin/bash
n=${1:-2}
tag=${2:-r02g}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests/test_sharded_gpu.py "tests/test_parity_gpu.py::test_single_step_all_variations" "tests/test_parity_gpu.py::test_single_step_overlay_and_stress_genome" "tests/test_parity_gpu.py::test_random_genomes" -m gpu -q > $out/pytest_mgpu_$tag.log 2>&1
tail -12 $out/pytest_mgpu_$tag.log | cut -c1-400
run_bench() {
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
run_bench 3 "--steps 3 --warmup 3"
run_bench 4 "--steps 2 --warmup 2"
