"""Static issue stalls of a kernel's hottest loop, read from the control fields of the SASS (no GPU needed).
usage: python tools/sass_stalls.py <lib.so> <kernel name substring> [word stores per loop pass, default: largest loop] [rows]
Every 128-bit sm_100 instruction carries, in bits 105..125, the cycles the warp waits before its next instruction
(stall, 4 bits), the yield hint, the scoreboard it sets for its result / its operands (write / read barrier, 7 = none) and
the mask of scoreboards it waits for.  The sum of the stall fields over a loop body is the least time ONE warp needs for
a pass, before any wait for a variable-latency result (LDS, LDG)."""
import re
import subprocess
import sys


def instructions(lib, kernel):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.splitlines()
    on, lines = False, []
    for ln in txt:
        if "Function :" in ln:
            on = kernel in ln
        elif on:
            lines.append(ln)
    ins, i = [], 0
    while i < len(lines):
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* (0x[0-9a-f]{16}) \*/", lines[i])
        m2 = re.match(r"\s+/\* (0x[0-9a-f]{16}) \*/", lines[i + 1]) if m and i + 1 < len(lines) else None
        if m and m2:
            ctrl = (int(m2.group(1), 16) >> 41) & 0x7fffff
            ins.append(dict(addr=int(m.group(1), 16), text=m.group(2).strip(), stall=ctrl & 0xf, yld=(ctrl >> 4) & 1,
                            wbar=(ctrl >> 5) & 7, rbar=(ctrl >> 8) & 7, wait=(ctrl >> 11) & 0x3f))
            i += 2
        else:
            i += 1
    return ins


def loops(ins):
    for x in ins:
        m = re.search(r"BRA\s+(?:\w+,\s*)?0x([0-9a-f]+)", x["text"]) if "BRA" in x["text"] else None
        if m and int(m.group(1), 16) < x["addr"]:
            yield [y for y in ins if int(m.group(1), 16) <= y["addr"] <= x["addr"]]


if __name__ == "__main__":
    lib, kernel = sys.argv[1], sys.argv[2]
    nst = int(sys.argv[3]) if len(sys.argv) > 3 else None
    rows = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    ins = instructions(lib, kernel)
    cand = [b for b in loops(ins) if nst is None or sum(1 for y in b if "STG" in y["text"] and ".U8" not in y["text"]) == nst]
    body = max(cand, key=len)
    print("# %s: loop of %d instructions at 0x%x, stall fields add up to %d cycles per pass" % (kernel, len(body), body[0]["addr"], sum(y["stall"] for y in body)))
    hist = {}
    for y in body:
        hist[y["stall"]] = hist.get(y["stall"], 0) + 1
    print("# instructions by stall field:", " ".join("%d:%d" % kv for kv in sorted(hist.items())))
    for y in body[:rows]:
        print("%05x stall=%2d yield=%d wbar=%d rbar=%d wait=%02x  %s" % (y["addr"], y["stall"], y["yld"], y["wbar"], y["rbar"], y["wait"], y["text"]))
