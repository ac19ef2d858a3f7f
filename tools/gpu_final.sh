# final 1-GPU validation of the round: parity tests, smoke, the default bench line (+ reference arm), ncu evidence
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/r2_final_pytest.log; tail -3 gpurun_out/r2_final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1; tail -1 gpurun_out/r2_final_smoke.log
timeout 600 python bench.py --impl reference > gpurun_out/r2_final_bench_reference.json 2> gpurun_out/r2_final_bench_reference.err
timeout 600 python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_final_bench.json")); r=json.load(open("gpurun_out/r2_final_bench_reference.json"))
print("ours", d["ms_per_step"], d["value"], d["grad_steps_per_sec"], "e2e", d["e2e"]["value"], "ref", r["value"], r["cpu_baseline"]["cores"], r["cpu_baseline"]["one_thread_value"])
print(d["roofline"]["per_kernel_us_per_step"]); print(d["breakdown"]); print(d["sub_records"])
print({k:v for k,v in d["roofline"].items() if k in ("kernel","bound","achieved","peak","frac","avg_us_per_launch")})
PY
