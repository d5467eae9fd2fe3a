#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_attention.py -m gpu -q -x --timeout 600 2>&1 | tail -5
for c in ZINC PATTERN CLUSTER MOLHIV; do timeout 200 python scripts/attn_microbench.py $c 2>&1 | tail -1; done
for c in ZINC PATTERN; do timeout 120 python scripts/builder_debug.py $c 2>&1 | tail -6; done
for c in ZINC PATTERN; do echo "== quick $c"; timeout 300 python bench.py --quick --steps 20 --warmup 5 --config $c 2>gpurun_out/q.err | tail -1 | cut -c1-160; tail -2 gpurun_out/q.err; done
