#!/bin/bash
# one GPU round-trip: parity tests, then the bench; prints a short summary.  Usage: tools/gpu_check.sh <tag> [pytest -k expr]
tag=$1; kexpr=$2
mkdir -p gpurun_out
if [ -n "$kexpr" ]; then timeout 900 python -m pytest tests -q -m gpu -k "$kexpr" 2>&1 | tail -40 > gpurun_out/${tag}_pytest.log
else timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -60 > gpurun_out/${tag}_pytest.log; fi
timeout 400 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_bench.json"))
print(d["ms_per_step"], d["value"], d["grad_steps_per_sec"], d["e2e"]["value"])
print(d["roofline"]["per_kernel_us_per_step"])
print(d["breakdown"])
print({k:v for k,v in d["roofline"].items() if k in ("kernel","bound","achieved","peak","frac")})
print(d["cpu_baseline"])
PY
tail -25 gpurun_out/${tag}_pytest.log
