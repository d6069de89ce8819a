"""Shared-memory wavefronts per SASS opcode of one kernel, from `ncu -i rep --page source --csv --kernel-name K`:
python tools/smem_wavefronts.py source.csv [launches_in_file]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
nl = int(sys.argv[2]) if len(sys.argv) > 2 else 1
hdr = next(r for r in rows if r and r[0] == 'Address')
isrc, iex = hdr.index('Source'), hdr.index('Instructions Executed')
iw, iwi = hdr.index('L1 Wavefronts Shared'), hdr.index('L1 Wavefronts Shared Ideal')
data = [r for r in rows if len(r) > iex and r[0].startswith('0x')]
tot = sum(int(r[iw]) for r in data)
print("shared-memory wavefronts per launch: %d; warp instructions per launch: %d" % (tot // nl, sum(int(r[iex]) for r in data) // nl))
agg = collections.defaultdict(lambda: [0, 0, 0])
for r in data:
    w = int(r[iw])
    if not w: continue
    toks = r[isrc].split(); op = toks[1] if toks[0].startswith('@') else toks[0]
    a = agg[op]; a[0] += w; a[1] += int(r[iwi]); a[2] += int(r[iex])
for op, (w, wi, n) in sorted(agg.items(), key=lambda x: -x[1][0]):
    print("%-10s wavefronts %11d (%4.1f%%)  ideal %11d  instructions %10d  wavefronts/instruction %.2f" % (op, w // nl, 100 * w / tot, wi // nl, n // nl, w / max(n, 1)))
