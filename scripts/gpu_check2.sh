#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2_tests.log 2>&1; tail -15 gpurun_out/r2_tests.log
for n in ZINC PATTERN; do
  echo "== layer microbench $n (tcgen05 linear)"; timeout 300 python scripts/layer_microbench.py $n 2>&1 | tail -3
  echo "== layer microbench $n (mma.sync linear)"; FETA_LINEAR_NO_TC5=1 timeout 300 python scripts/layer_microbench.py $n 2>&1 | tail -3
done
for tc in 0 1; do for c in ZINC PATTERN; do echo "== quick $c TC5=$tc"; FETA_LINEAR_TC5=$tc timeout 300 python bench.py --quick --steps 20 --warmup 5 --config $c 2>&1 | tail -1; done; done
