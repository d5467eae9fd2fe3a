#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/r2_tests.log 2>&1; tail -3 gpurun_out/r2_tests.log
( time timeout 900 python bench.py > gpurun_out/r2_bench_full.json 2> gpurun_out/r2_bench_full.err ) 2>&1 | grep real; tail -3 gpurun_out/r2_bench_full.err; cut -c1-200 gpurun_out/r2_bench_full.json
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err ) 2>&1 | grep real; cut -c1-300 gpurun_out/r2_bench_ref.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for c in CLUSTER MOLHIV MUTAG; do timeout 300 python bench.py --config $c --no-sweep --no-cpu-baseline --no-extra --no-builder > gpurun_out/r2_bench_$c.json 2>gpurun_out/q.err; cut -c1-150 gpurun_out/r2_bench_$c.json; done
