#!/bin/bash
( time timeout 1200 python bench.py > gpurun_out/r2c_bench_full.json 2> gpurun_out/r2c_bench_full.err ) 2>&1 | grep real; tail -3 gpurun_out/r2c_bench_full.err; cut -c1-400 gpurun_out/r2c_bench_full.json
for c in MOLHIV; do timeout 200 python scripts/attn_microbench.py $c 2>&1 | tail -1; done
