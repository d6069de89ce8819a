"""Instruction mix of a kernel from `ncu --page source --csv` output: python tools/sass_mix.py file.csv [copies]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
nl = int(sys.argv[2]) if len(sys.argv) > 2 else 1        # copies of the listing in the file (ncu prints the source page per view)
hdr = next(r for r in rows if r and r[0] == 'Address')
isrc, iex, ismp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
mix, smp = collections.Counter(), collections.Counter()
tot = 0
for r in rows:
    if len(r) <= iex or not r[0].startswith('0x'): continue
    toks = r[isrc].split()
    op = toks[1] if toks[0].startswith('@') else toks[0]
    parts = op.split('.')
    key = parts[0]
    if key in ('LDS', 'STS', 'LDG', 'STG', 'SHFL', 'IMAD', 'SHF', 'LOP3') and len(parts) > 1:
        key += '.' + parts[1]
    n = int(r[iex]); mix[key] += n; smp[key] += int(r[ismp]); tot += n
print('total warp instructions', tot // nl)
for op, n in mix.most_common(30):
    print('  %-14s %12d  %5.1f%%   samples %6d' % (op, n // nl, 100.0 * n / tot, smp[op] // nl))
