# multi-GPU check: usage tools/gpu_multi.sh N   (run under gpurun --gpus N)
N=$1
mkdir -p gpurun_out
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests/test_p2p_gpu.py -q -m gpu 2>&1 | tail -4; fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 200 --warmup 5 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
tail -5 gpurun_out/r2_bench_${N}gpu.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_${N}gpu.json"))
print(d["n_gpus"], d["ms_per_step"], d["value"], d["grad_steps_per_sec"], "e2e", d["e2e"]["value"])
print("replicas_identical", d["replicas_identical"], d["exchange_check"]); print("nccl_arm", d["nccl_arm"]); print(d["config"]["grad_collective"])
print(d["roofline"]["per_kernel_us_per_step"])
PY
