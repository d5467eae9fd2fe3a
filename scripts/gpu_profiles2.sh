#!/bin/bash
# second pass: the kernels the first pass (scripts/gpu_profiles.sh) did not reach -- Chebyshev dx / dTheta on the sweep,
# the projections' input-gradient kernel (selected by mangled template arguments).  Reports reduced to raw-metric CSV.
N="ncu --set full --import-source on --clock-control none -f --kernel-name-base mangled"
raw() { ncu -i $1.ncu-rep --page raw --csv > $1.raw.csv 2>/dev/null; rm -f $1.ncu-rep; }
for F in 16 8; do
  timeout 600 $N -k regex:"cheb_fwd_lane_kernelILi${F}ELi2ELi4ELb1E" -s 2 -c 1 -o gpurun_out/r2g_sweep_f${F}_dx python bench.py --sweep-only --sweep-f $F --sweep-rows 3000000 > gpurun_out/ncu_sweep.log 2>&1; raw gpurun_out/r2g_sweep_f${F}_dx
done
for c in ZINC PATTERN; do
  timeout 600 $N -k regex:"linear_simt_kernelILi1E" -s 4 -c 2 -o gpurun_out/r2g_linear_dx_$c python scripts/layer_microbench.py $c > gpurun_out/ncu_lin.log 2>&1; raw gpurun_out/r2g_linear_dx_$c
done
ls -la gpurun_out/r2g_*dx*
