# development run: the shared-memory thread-per-stream kernels (tests, then the block-count sweep)
timeout 900 python -m pytest tests -m gpu -x -q -k "many_streams or reference_formats" > gpurun_out/t_smem.log 2>&1
tail -3 gpurun_out/t_smem.log
timeout 300 python tools/tps_sweep.py 131072:8192 131072:1024 65536:256 16384:512 4096:256 > gpurun_out/tps_smem4.log 2>&1
