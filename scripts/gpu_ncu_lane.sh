#!/bin/bash
timeout 600 ncu --set full --import-source on --clock-control none -k regex:cheb_fwd_lane -c 1 -o gpurun_out/r2_lane_f16_b python bench.py --sweep-only --sweep-f 16 --sweep-rows 1500000 > gpurun_out/ncu_lane.log 2>&1; tail -2 gpurun_out/ncu_lane.log
