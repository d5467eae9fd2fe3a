#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/r2_tests.log 2>&1; tail -4 gpurun_out/r2_tests.log
( time timeout 900 python bench.py > gpurun_out/r2_bench_full.json 2> gpurun_out/r2_bench_full.err ) 2>&1 | grep real; tail -3 gpurun_out/r2_bench_full.err; cut -c1-300 gpurun_out/r2_bench_full.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
