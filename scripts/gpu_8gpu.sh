#!/bin/bash
run() { env FETA_COMM_SLICES=$2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 8 --quick --steps 30 --warmup 5 --config $1 2>gpurun_out/q8.err | tail -1 | cut -c1-130; }
echo "PATTERN slices 6"; run PATTERN 6
echo "PATTERN slices 3"; run PATTERN 3
echo "ZINC slices 6"; run ZINC 6
