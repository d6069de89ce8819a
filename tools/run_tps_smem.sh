# development run: lanes per warp of the thread-per-stream kernels (FSE_DEV build under tools/bin)
export FSE_B200_LIB=$PWD/tools/bin/libdev.so
for cfg in "6 10" "7 13" "4 6" "3 5"; do
  set -- $cfg
  echo "== TPS_LPW(dec)=$1 ENC_LPW=$2"
  FSE_B200_TPS_LPW=$1 FSE_B200_TPS_ENC_LPW=$2 timeout 300 python tools/tps_sweep.py 131072:8192 131072:1024
done > gpurun_out/tps_smem7.log 2>&1
