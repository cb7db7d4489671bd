#!/bin/bash
# Host-side logic under AddressSanitizer + UndefinedBehaviorSanitizer (no GPU needed): builds a sanitized copy of
# librefrakt_b200.so in /tmp, runs tests/test_host_cpu.py (parsers, code generator, NVRTC builds, writers, the mutated-input
# tests) against it and restores the real library. Prints the number of sanitizer reports (0 expected).
set -e
cd "$(dirname "$0")/.."
bd=/tmp/rfk_asan_build
mkdir -p $bd
python - <<PY
import os, subprocess, sys
sys.path.insert(0, ".")
from refrakt_b200 import build as b
emb = b._embed(os.path.join(b.HERE, "build"))
common = [b.NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O1", "-g", "-std=c++20", "-Xcompiler", "-fPIC,-fsanitize=address,-fsanitize=undefined,-fno-omit-frame-pointer",
          "-I" + b.CSRC, "-I" + os.path.join(b.HERE, "..", "include")]
objs, procs = [], []
for src in b.HOST_SOURCES + b.DEVICE_SOURCES + [emb]:
    path = src if os.path.isabs(src) else os.path.join(b.CSRC, src)
    obj = os.path.join("$bd", os.path.basename(src) + ".o")
    objs.append(obj)
    procs.append(subprocess.Popen(common + ["-c", path, "-o", obj]))
assert all(p.wait() == 0 for p in procs)
subprocess.check_call([b.NVCC, "-shared", "-o", "$bd/librefrakt_b200.so"] + objs + ["-cudart", "static", "-lnvrtc", "-lz", "-ldl", "-Xcompiler", "-fsanitize=address,-fsanitize=undefined"])
PY
real=refrakt_b200/librefrakt_b200.so
cp $real $bd/real.so
trap 'cp $bd/real.so $real' EXIT
cp $bd/librefrakt_b200.so $real
LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" ASAN_OPTIONS=detect_leaks=0:halt_on_error=0 UBSAN_OPTIONS=print_stacktrace=1 \
  python -m pytest tests/test_host_cpu.py -q -s -m "not gpu" -p no:cacheprovider > $bd/run.log 2>&1 || true
tail -n 2 $bd/run.log
echo "sanitizer reports: $(grep -c 'AddressSanitizer\|runtime error' $bd/run.log)"
