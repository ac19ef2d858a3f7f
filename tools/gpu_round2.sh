mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -30 > gpurun_out/r2b_pytest.log
tail -5 gpurun_out/r2b_pytest.log
if grep -q "failed" gpurun_out/r2b_pytest.log; then exit 1; fi
timeout 500 python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
timeout 300 python bench.py --workload mixed128 --steps 30 > gpurun_out/r2b_mixed128.json 2> gpurun_out/r2b_mixed128.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2b_bench.json"))
print(d["ms_per_step"], d["value"], d["grad_steps_per_sec"], d["e2e"]["value"])
print(d["roofline"]["per_kernel_us_per_step"]); print(d["breakdown"]); print(d["sub_records"])
PY
cat gpurun_out/r2b_mixed128.json | cut -c1-1500; tail -3 gpurun_out/r2b_mixed128.err
timeout 1000 python tools/learn_curve.py --n-envs 1 --iters 2000000 --eval-every 50000 --eval-episodes 10 --budget-s 840 --out gpurun_out/r2_learn_n1.jsonl > gpurun_out/r2_learn_n1.log 2>&1
tail -4 gpurun_out/r2_learn_n1.log | cut -c1-400
