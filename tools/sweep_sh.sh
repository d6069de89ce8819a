#!/bin/bash
# development sweep of the shared-table kernels (needs tools/bin/libfse_dev.so built with -DFSE_DEV)
export FSE_B200_LIB=tools/bin/libfse_dev.so
W=${1:-c5}
for cfg in "16 16 16 32" "16 32 8 32" "16 16 12 24" "8 32 16 32"; do
  set -- $cfg
  FSE_B200_SH_ROUNDS=$1 FSE_B200_SH_NSR=$2 FSE_B200_SH_WARPS=$3 FSE_B200_SHD_WARPS=$4 python bench.py --workload $W --steps 5 --no-e2e --no-cpu --no-configs --no-traffic 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$cfg', round(d['value'],1), {k: round(v,3) for k,v in d['kernel_ms_per_step'].items()})"
done
