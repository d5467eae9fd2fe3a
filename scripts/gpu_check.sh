#!/bin/bash
# One GPU call: the whole GPU suite, the default bench line, and the launch list of the graphed ZINC step.
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/r2_tests.log 2>&1; tail -4 gpurun_out/r2_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -3 gpurun_out/r2_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_bench.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "spread", "gpu_launches")})
    print("e2e", d["e2e"])
    print("extra", d["extra"])
    for f in ("F16", "F8"):
        print(f, {k: (v["frac"], v["ms_per_launch"]) for k, v in d["roofline_sweep"][f].items()})
    print("cpu", d["cpu_baseline"])
    print("roofline", d["roofline"]["frac"], d["roofline"]["us_per_launch"], d["roofline_cheb_in_step"]["us_per_launch"])
except Exception as e:
    print("bench parse failed", e)
PY
if [ "$1" == "launches" ]; then
python bench.py --quick --steps 2 --warmup 1 --pool 4 --no-extra > gpurun_out/plainq.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_launches_zinc.csv python bench.py --quick --steps 2 --warmup 1 --pool 4 --no-extra > gpurun_out/ncuq.log 2>&1; tail -2 gpurun_out/ncuq.log
fi
