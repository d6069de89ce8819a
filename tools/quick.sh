#!/bin/bash
# quick.sh "<bench args>" ... : one short device-resident run per argument string, prints value and kernel times
for a in "$@"; do
  python bench.py $a --steps 5 --no-e2e --no-cpu --no-configs --no-traffic 2>gpurun_out/quick.err | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$a', '| value', round(d['value'],1), 'enc', round(d['encode_GBps'],1), 'dec', round(d['decode_GBps'],1), 'ratio', round(d['compressed_ratio'],4), {k: round(v,3) for k,v in d['kernel_ms_per_step'].items()})
except Exception as e:
    print('$a', 'FAILED', e); print(open('gpurun_out/quick.err').read()[-1500:])"
done
