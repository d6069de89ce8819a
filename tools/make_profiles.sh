#!/bin/bash
# make_profiles.sh <rep.ncu-rep> <tag> <kernel_regex>... : text summaries of one ncu --set full capture (read here, no GPU):
# raw-page summary, and per kernel the shared-memory wavefronts per opcode, instruction mix and per-function shares.
# The library must be the build the capture ran (line_profile joins the cubin's line table with the capture).
set -e
rep=$1; tag=$2; shift 2
python tools/ncu_summary.py "$rep" > profiles/r2_ncu_full_${tag}.txt
: > profiles/r2_smem_wavefronts_${tag}.txt; : > profiles/r2_sass_mix_${tag}.txt; : > profiles/r2_function_profile_${tag}.txt
for k in "$@"; do
  ncu -i "$rep" --page source --csv --kernel-name "regex:$k" > /tmp/src_$k.csv 2>/dev/null
  nl=$(grep -c '^"Address"' /tmp/src_$k.csv)
  echo "# $k, $tag, per launch (mean of $nl captured launches); ncu --set full, source page" >> profiles/r2_smem_wavefronts_${tag}.txt
  python tools/smem_wavefronts.py /tmp/src_$k.csv $nl >> profiles/r2_smem_wavefronts_${tag}.txt
  echo "# instruction mix, $k, $tag, per launch" >> profiles/r2_sass_mix_${tag}.txt
  python tools/sass_mix.py /tmp/src_$k.csv $nl > /tmp/mix_$k.txt; head -28 /tmp/mix_$k.txt >> profiles/r2_sass_mix_${tag}.txt
  mangled=$(cuobjdump -elf entropy_coders_b200/libfse_b200.so 2>/dev/null | grep -o "\.text\._Z[A-Za-z0-9_]*$k[A-Za-z0-9_]*" | sort -u | head -1 | sed 's/^\.text\.//')
  [ -n "$mangled" ] && python tools/line_profile.py "$mangled" /tmp/src_$k.csv >> profiles/r2_function_profile_${tag}.txt
done
