# fold + bf16 attention + variants + wgrad reduce: tests, timeline, bench (chunk 64 vs 128)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/r2g_pytest.log
tail -15 gpurun_out/r2g_pytest.log
timeout 200 python tools/prof_act.py 3 --timeline > gpurun_out/r2g_timeline.txt 2>&1; tail -12 gpurun_out/r2g_timeline.txt
timeout 500 python bench.py --no-cpu-baseline > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2g_bench.json"))
print(d["ms_per_step"], d["value"], d["grad_steps_per_sec"], d["e2e"]["value"])
print(d["roofline"]["per_kernel_us_per_step"]); print(d["breakdown"]); print(d["sub_records"])
print({k:v for k,v in d["roofline"].items() if k in ("kernel","bound","achieved","peak","frac")})
PY
tail -3 gpurun_out/r2g_bench.err
if grep -q "failed" gpurun_out/r2g_pytest.log; then echo "TESTS FAILED: skipping the learning run"; exit 0; fi
timeout 1300 python tools/learn_curve.py --n-envs 4096 --iters 1300000 --eval-every 100000 --eval-episodes 4096 --buf-size 419430400 --record-every 32 --budget-s 900 --out gpurun_out/r2_learn_n4096_rec32.jsonl > gpurun_out/r2_learn_n4096_rec32.log 2>&1
python - <<PY
import json
for l in open("gpurun_out/r2_learn_n4096_rec32.jsonl"):
    d=json.loads(l); print(d["iteration"], round(d["success_rate"],3), round(d["mean_return"],3), round(d["episode_length"],1), round(d["td_error"],5), round(d["q_mean"],3))
PY
