#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2_tests.log 2>&1; tail -40 gpurun_out/r2_tests.log
for tc in 0 1; do for c in ZINC PATTERN; do echo "== quick $c TC5=$tc"; FETA_LINEAR_TC5=$tc timeout 300 python bench.py --quick --steps 20 --warmup 5 --config $c 2>gpurun_out/q_${c}_${tc}.err | tail -1 | tee gpurun_out/q_${c}_${tc}.json | cut -c1-400; tail -3 gpurun_out/q_${c}_${tc}.err; done; done
