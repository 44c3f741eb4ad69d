#!/bin/bash
# attention kernel: parity tests + timings of the shipped library and of the variant builds named on the command line
timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_backward_kernels_gpu.py -x -q -m gpu -k attention > gpurun_out/attn_tests.log 2>&1
grep -E "passed|failed|timed out|FAILED|assert" gpurun_out/attn_tests.log | head -8
run() { timeout 60 python tools/attn_one.py 10 2>&1 | tail -1; }
echo "== shipped"; run; B=32 T=16 run; B=16 T=16 N0=257 H=16 run; B=4 T=70 run
for v in "$@"; do
  lib=gava_clip_b200/lib/variants/libgava_$v.so
  [ -f $lib ] || continue
  echo "== variant $v"
  GAVA_LIB_PATH=$lib run
done
