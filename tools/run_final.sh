# round-end verification on one GPU: all GPU tests, smoke, the default bench line, an ncu capture of the thread-per-stream kernels
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/final_tests.log 2>&1; tail -2 gpurun_out/final_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/final_smoke.log 2>&1; tail -1 gpurun_out/final_smoke.log
timeout 1500 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -c 600 gpurun_out/final_bench.json
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name regex:k_tps -c 8 -o gpurun_out/r2_tps_final -f python tools/tps_profile.py > gpurun_out/tps_ncu.log 2>&1
tail -2 gpurun_out/tps_ncu.log
