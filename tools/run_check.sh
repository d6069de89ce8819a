# all GPU tests and the smoke on one GPU, then the thread-per-stream sweep at c4's size
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/final_tests.log 2>&1; tail -2 gpurun_out/final_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/final_smoke.log 2>&1; tail -1 gpurun_out/final_smoke.log
timeout 300 python tools/tps_sweep.py 131072:8192 131072:1024 > gpurun_out/tps_final_sweep.log 2>&1
