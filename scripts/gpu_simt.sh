#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_dense.py -m gpu -q -x --timeout 600 2>&1 | tail -2
for c in ZINC PATTERN; do timeout 280 python scripts/layer_microbench.py $c 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:v for k,v in d.items() if 'proj' in k or 'ffn' in k})"; done
for c in ZINC PATTERN; do for m in 1 0; do echo "== quick $c SIMT=$m"; FETA_LINEAR_SIMT=$m timeout 300 python bench.py --quick --steps 20 --warmup 5 --config $c 2>gpurun_out/q.err | tail -1 | cut -c1-120; tail -2 gpurun_out/q.err; done; done
