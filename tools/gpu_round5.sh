mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -60 > gpurun_out/r2e_pytest.log
grep -E "^E  |passed|failed|^FAILED" gpurun_out/r2e_pytest.log | head -30
timeout 200 python tools/prof_act.py 3 --timeline > gpurun_out/r2e_timeline.txt 2>&1; tail -4 gpurun_out/r2e_timeline.txt | cut -c1-900
timeout 400 python bench.py --workload memory --steps 60 --no-cpu-baseline > gpurun_out/r2e_bench_memory.json 2> gpurun_out/r2e_bench_memory.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2e_bench_memory.json"))
print("memory", d["ms_per_step"], d["value"], d["grad_steps_per_sec"]); print(d["roofline"]["per_kernel_us_per_step"]); print(d["breakdown"])
PY
