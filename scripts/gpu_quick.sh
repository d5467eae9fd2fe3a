#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/r2_tests.log 2>&1; tail -3 gpurun_out/r2_tests.log | cut -c1-300
for c in ZINC PATTERN; do
  for v in "FETA_COEFF_BRANCH_STREAM=1" "FETA_COEFF_BRANCH_STREAM=0"; do
    echo "== quick $c $v"; env $v timeout 300 python bench.py --quick --steps 30 --warmup 5 --config $c 2>gpurun_out/q.err | tail -1 | cut -c1-130; tail -2 gpurun_out/q.err
  done
done
