# development run: the shared-memory thread-per-stream kernels (tests, then a sweep with the FSE_DEV build)
timeout 900 python -m pytest tests -m gpu -x -q -k "many_streams or reference_formats" > gpurun_out/t_smem.log 2>&1
tail -3 gpurun_out/t_smem.log
export FSE_B200_LIB=$PWD/tools/bin/libdev.so
for cfg in "4 8 1" "4 8 0" "5 4 1" "6 6 1"; do
  set -- $cfg
  echo "== TPS_LPW(dec)=$1 ENC_LPW=$2 COMPACT=$3"
  FSE_B200_TPS_LPW=$1 FSE_B200_TPS_ENC_LPW=$2 FSE_B200_TPS_COMPACT=$3 timeout 300 python tools/tps_sweep.py 131072:8192 131072:1024 16384:512
done > gpurun_out/tps_smem3.log 2>&1
