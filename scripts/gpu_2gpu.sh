#!/bin/bash
for c in ZINC PATTERN; do
DDP2_CONFIG=$c timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 tests/ddp_two_rank_check.py 2>&1 | grep -v "OMP_NUM\|\*\*\*\*" | grep DDP2
done
