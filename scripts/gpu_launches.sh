#!/bin/bash
# launch lists (ncu, cold-cache serialised: read SHARES) of the graph-replayed ZINC and PATTERN steps, current default build
for c in ZINC PATTERN; do
  timeout 300 python bench.py --quick --steps 2 --warmup 1 --pool 4 --no-extra --no-builder --config $c > gpurun_out/r2b_plain_$c.log 2>&1 || { echo "plain $c failed"; tail -5 gpurun_out/r2b_plain_$c.log; continue; }
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2b_launches_$c.csv \
     python bench.py --quick --steps 2 --warmup 1 --pool 4 --no-extra --no-builder --config $c > gpurun_out/r2b_ncu_$c.log 2>&1
  tail -1 gpurun_out/r2b_ncu_$c.log | cut -c1-200
done
