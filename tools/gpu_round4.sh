mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/r2d_pytest.log
tail -6 gpurun_out/r2d_pytest.log
timeout 200 python tools/prof_act.py 3 --timeline > gpurun_out/r2d_timeline.txt 2>&1; tail -4 gpurun_out/r2d_timeline.txt
timeout 300 python tools/tune_wgrad.py 2>&1 | tail -12 | tee gpurun_out/r2d_tune_wgrad.txt
