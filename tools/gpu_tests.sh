#!/bin/bash
# run the GPU suite file by file with a per-file timeout (a hung kernel must not eat the gpurun budget)
# usage: tools/gpu_tests.sh [pytest -k expr]   -> gpurun_out/gputests.log
mkdir -p gpurun_out
: > gpurun_out/gputests.log
for f in "${@:-tests}"; do
  for t in $(ls $f/test_*cuda*.py $f 2>/dev/null | grep "\.py$" | sort -u); do
    echo "=== $t" >> gpurun_out/gputests.log
    timeout 300 python -m pytest $t -q -m gpu -x 2>&1 | tail -25 >> gpurun_out/gputests.log
    echo "rc=$?" >> gpurun_out/gputests.log
  done
done
grep -E "^===|passed|failed|error|rc=" gpurun_out/gputests.log | tail -60
