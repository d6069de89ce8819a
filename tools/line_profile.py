"""Instructions executed and stall samples per source function of one kernel: joins `nvdisasm -g -c` of the cubin
(line info, needs -lineinfo) with `ncu -i rep --page source --csv --kernel-name K` (per-instruction counters), in order.
python tools/line_profile.py kernel_mangled_name source.csv [unused] [library.so]"""
import csv, re, subprocess, sys, os, collections, tempfile
name, src_csv = sys.argv[1], sys.argv[2]
nl = int(sys.argv[3]) if len(sys.argv) > 3 else 2
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(sys.argv[4]) if len(sys.argv) > 4 else os.path.join(root, "entropy_coders_b200", "libfse_b200.so")], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
ins, cur, on = [], None, False
for l in dis:
    if l.startswith(".text.") and l.rstrip().endswith(":"):
        on = (l.strip() == ".text.%s:" % name)
        continue
    if not on: continue
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,6}\*/", l): ins.append(cur)
rows = list(csv.reader(open(src_csv)))
hdr = next(r for r in rows if r and r[0] == "Address")
iex, ismp = hdr.index("Instructions Executed"), hdr.index("# Samples")
data = [r for r in rows if len(r) > iex and r[0].startswith("0x")]
# the csv holds one instruction table per captured launch of the kernel: keep the last one
assert len(ins) and len(data) % len(ins) == 0, (len(data), len(ins))
data = data[-len(ins):]
# function ranges from the sources: "name(" at column 0..4 preceded by __device__/__global__
funcs = {}
for f in set(i[0] for i in ins if i):
    path = os.path.join(root, "entropy_coders_b200", "csrc", f)
    if not os.path.exists(path): continue
    marks = []
    for n, line in enumerate(open(path), 1):
        m = re.search(r"(?:__device__|__global__)[^;(]*?\b(\w+)\s*\(", line)
        if m and not line.strip().startswith("//"): marks.append((n, m.group(1)))
    funcs[f] = marks
def fn(file, line):
    best = "?"
    for n, nm in funcs.get(file, []):
        if n <= line: best = nm
    return best
agg = collections.defaultdict(lambda: [0, 0])
for i, r in zip(ins, data):
    key = fn(*i) if i else "?"
    agg[key][0] += int(r[iex]); agg[key][1] += int(r[ismp])
ti, ts = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
print("%s: %d warp instructions, %d samples" % (name, ti, ts))
for k, (a, b) in sorted(agg.items(), key=lambda x: -x[1][0]):
    print("  %-28s instructions %5.1f%%   samples %5.1f%%" % (k, 100.0 * a / ti, 100.0 * b / ts))
