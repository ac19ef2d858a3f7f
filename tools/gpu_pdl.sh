#!/bin/bash
# PDL A/B + parity with the knob on (bounded: the round's last GPU minutes)
mkdir -p gpurun_out
timeout 120 python tools/pdl_ab.py 400 > gpurun_out/pdl_ab.log 2>&1; echo "pdl_ab rc=$?" >> gpurun_out/pdl_ab.log
DTQN_B200_PDL=1 timeout 70 python -m pytest tests/test_net_gpu.py -m gpu -x -q > gpurun_out/pdl_tests_on.log 2>&1; echo "rc=$?" >> gpurun_out/pdl_tests_on.log
timeout 50 python -m pytest tests/test_net_gpu.py -m gpu -x -q > gpurun_out/pdl_tests_off.log 2>&1; echo "rc=$?" >> gpurun_out/pdl_tests_off.log
tail -8 gpurun_out/pdl_ab.log; tail -3 gpurun_out/pdl_tests_on.log; tail -3 gpurun_out/pdl_tests_off.log
