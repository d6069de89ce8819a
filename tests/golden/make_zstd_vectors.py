"""Writes tests/golden/zstd_weight_streams.json: two-state FSE streams WRITTEN BY libzstd (the FSE-compressed Huffman weights
of literal-only frames, tests/zstd_interop.py), each with the weights they decode to -- verified here by Huffman-decoding
the frame's literals back to the input -- and the normalised counts of their NCount header.  Needs libzstd; the fixture
lets the oracle and the GPU path be checked against libzstd's bytes where it is absent.
    python tests/golden/make_zstd_vectors.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O          # noqa: E402
import zstd_interop as ZI       # noqa: E402

CASES = [(1, 4000, 90, 0.93, 20), (2, 6000, 120, 0.95, 0), (3, 3000, 60, 0.90, 64), (4, 12000, 200, 0.97, 10),
         (5, 5000, 150, 0.985, 30), (6, 2500, 40, 0.85, 97), (7, 9000, 250, 0.99, 3), (8, 3500, 75, 0.92, 50)]

out = {"libzstd": ZI.zstd().ZSTD_versionString().decode(), "streams": []}
for seed, n, nsym, decay, base in CASES:
    src = ZI.skewed_bytes(seed, n, nsym, decay, base)
    info = ZI.parse_first_block(ZI.zstd_compress_literals_only(np.frombuffer(src, dtype=np.uint8)))
    if info is None or info["tree"][0] >= 128 or info["sequences"][:1] != b"\x00" or info["regen"] != len(src):
        continue
    blob = bytes(info["tree"][1:1 + info["tree"][0]])
    rc, nh, consumed = O.ncount_read(blob)
    weights = O.decompress_n_exhaust(blob, 2, 255)
    assert rc == 0 and ZI.huf_decode_literals(info, list(weights)) == src      # the weights are what libzstd encoded
    out["streams"].append({"case": [seed, n, nsym, decay, base], "blob": blob.hex(), "weights": weights.hex(),
                           "table_log": nh.log2, "header_bytes": consumed, "norm": list(nh.table[:nh.table_len])})
# two small frames with FSE-compressed sequence tables (tests/zstd_interop.py: decode_sequences)
out["sequence_frames"] = [{"case": [seed, size, level], "frame": ZI.zstd_compress(ZI.wordy_bytes(seed, size), level).hex()}
                          for seed, size, level in ((9, 9000, 3), (10, 14000, 5))]
with open(os.path.join(HERE, "zstd_weight_streams.json"), "w") as f:
    json.dump(out, f, indent=1)
print(len(out["streams"]), "streams from libzstd", out["libzstd"])
