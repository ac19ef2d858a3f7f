#!/bin/bash
# backward launch-shape sweep, then the network parity tests under the fastest setting (bounded)
mkdir -p gpurun_out
timeout 100 python tools/tune_bwd.py 300 > gpurun_out/tune_bwd.log 2>&1; echo "rc=$?" >> gpurun_out/tune_bwd.log
cat gpurun_out/tune_bwd.log | tail -16
if [ -f gpurun_out/best_bwd.env ]; then
  . gpurun_out/best_bwd.env; cat gpurun_out/best_bwd.env
  timeout 60 python -m pytest tests/test_net_gpu.py -m gpu -q > gpurun_out/tune_bwd_tests.log 2>&1; echo "rc=$?" >> gpurun_out/tune_bwd_tests.log
  tail -12 gpurun_out/tune_bwd_tests.log
fi
