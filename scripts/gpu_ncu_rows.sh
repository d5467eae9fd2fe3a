#!/bin/bash
# ncu --set full of the matrix-free attention backward kernel at the ZINC and PATTERN shapes, + ZINC forward
for c in ZINC PATTERN; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:attn_rows_bwd -s 2 -c 1 -o gpurun_out/r2_attn_rows_bwd_$c -f \
     python scripts/attn_microbench.py $c > gpurun_out/ncu_rows_$c.log 2>&1; tail -1 gpurun_out/ncu_rows_$c.log | cut -c1-200
done
