#!/bin/bash
# Round-2 evidence for profiles/: launch lists of the steps and `ncu --set full` captures of the kernels the bench
# line's roofline objects name.  Every command ran to exit 0 WITHOUT ncu first (scripts/gpu_full.sh).
# Reports are reduced to their raw-metric CSV on the box (the .ncu-rep files exceed the 64 MiB return limit).
N="ncu --set full --import-source on --clock-control none -f"
raw() { ncu -i $1.ncu-rep --page raw --csv > $1.raw.csv 2>/dev/null; rm -f $1.ncu-rep; }
for F in 16 8; do
  timeout 600 $N -k regex:"cheb_fwd_lane|cheb_dtheta_lane" -c 3 -o gpurun_out/r2f_sweep_f${F} python bench.py --sweep-only --sweep-f $F --sweep-rows 3000000 > gpurun_out/ncu_sweep.log 2>&1; tail -1 gpurun_out/ncu_sweep.log | cut -c1-120
  raw gpurun_out/r2f_sweep_f${F}
done
for c in ZINC PATTERN; do
  timeout 600 $N -k regex:attn_rows_fwd -s 4 -c 1 -o gpurun_out/r2f_attn_rows_fwd_$c python scripts/attn_microbench.py $c > gpurun_out/ncu_rows.log 2>&1; tail -1 gpurun_out/ncu_rows.log | cut -c1-200
  raw gpurun_out/r2f_attn_rows_fwd_$c
  timeout 600 $N -k regex:attn_rows_bwd -s 2 -c 1 -o gpurun_out/r2f_attn_rows_bwd_$c python scripts/attn_microbench.py $c > gpurun_out/ncu_rows.log 2>&1
  raw gpurun_out/r2f_attn_rows_bwd_$c
  timeout 600 $N -k regex:linear_simt -s 8 -c 8 -o gpurun_out/r2f_linear_simt_$c python scripts/layer_microbench.py $c > gpurun_out/ncu_lin.log 2>&1; tail -1 gpurun_out/ncu_lin.log | cut -c1-200
  raw gpurun_out/r2f_linear_simt_$c
done
for c in ZINC PATTERN; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2f_launches_$c.csv \
     python bench.py --quick --steps 2 --warmup 1 --pool 4 --no-extra --no-builder --config $c > gpurun_out/r2f_ncu_$c.log 2>&1
  tail -1 gpurun_out/r2f_ncu_$c.log | cut -c1-100; grep -c . gpurun_out/r2f_launches_$c.csv
done
du -sh gpurun_out
