import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print(round(d["value"], 1), "enc", round(d["encode_GBps"], 1), "dec", round(d["decode_GBps"], 1),
      {k: round(v, 4) for k, v in d["kernel_ms_per_step"].items()}, "e2e", d["e2e"]["value"] if d.get("e2e") else None)
