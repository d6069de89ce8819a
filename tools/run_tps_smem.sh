# development run: the shared-memory thread-per-stream kernels (tests, then the block-count sweep)
timeout 900 python -m pytest tests -m gpu -x -q -k "many_streams or reference_formats" > gpurun_out/t_smem.log 2>&1
tail -3 gpurun_out/t_smem.log
for v in dev devb; do echo "== $v"; FSE_B200_LIB=$PWD/tools/bin/lib$v.so timeout 300 python tools/tps_sweep.py 131072:8192 131072:1024 16384:512; done > gpurun_out/tps_smem5.log 2>&1
