# development run: thread-per-stream kernels, the release build against the FSE_DEV build under tools/bin
timeout 900 python -m pytest tests -m gpu -x -q -k "many_streams or reference_formats" > gpurun_out/t_smem.log 2>&1
tail -3 gpurun_out/t_smem.log
for v in rel dev; do echo "== $v"; if [ $v = dev ]; then export FSE_B200_LIB=$PWD/tools/bin/libdev.so; fi; timeout 300 python tools/tps_sweep.py 131072:8192 131072:1024 16384:512; done > gpurun_out/tps_smem6.log 2>&1
