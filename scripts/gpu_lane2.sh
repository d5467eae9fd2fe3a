#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_cheb.py tests/test_gpu_reference_pin.py -m gpu -q -x --timeout 600 2>&1 | tail -3
for f in 16 8; do
  echo "== sweep F=$f"; timeout 300 python bench.py --sweep-only --sweep-f $f 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:(v['achieved'],v['frac'],v['ms_per_launch']) for k,v in d.items()})"
done
