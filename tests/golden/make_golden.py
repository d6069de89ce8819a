"""Regenerates tests/golden/kat.json.

The reference is a Rust crate and cannot be executed in this image (no cargo/rustc), and its own
tests hold no encoded-byte vectors.  These known-answer vectors are therefore produced by the C
oracle and accepted only if the independent mechanics model (oracle/pymodel.py) produces the
same bytes; KAT-A additionally matches the hand derivation in SURVEY.md Appendix C.

    python tests/golden/make_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]

import numpy as np  # noqa: E402
import oracle_lib as O  # noqa: E402
import pymodel as M  # noqa: E402


def exp_dist(log2):  # histogram.rs:621-638
    data, remaining, sym = [], 1 << log2, 0
    while True:
        data += [sym] * (remaining >> 1)
        remaining -= remaining >> 1
        sym += 1
        if remaining == 1:
            data.append(sym)
            break
    return bytes(data)


def one(name, src, table_log=0, note=""):
    src = bytes(src)
    ent = {"name": name, "note": note, "src_hex": src.hex() if len(src) <= 512 else None,
           "src_len": len(src), "table_log_req": table_log}
    h = O.histogram(src)
    if table_log == 0:
        rc, tl = O.optimal_log2(h)
        assert rc == 0
    else:
        tl = table_log
    rc, nh = O.normalize(h, tl)
    assert rc >= 0
    ent["log2"] = nh.log2
    ent["table_len"] = nh.table_len
    ent["norm"] = list(nh.table)[: nh.table_len]
    hdr, hbits = O.ncount_write(nh)
    ent["header_hex"] = hdr.hex()
    ent["header_bits"] = hbits
    et = O.enc_table(nh)
    dt = O.dec_table(nh)
    size = 1 << nh.log2
    if size <= 64:
        ent["spread"] = list(et.symbols)[:size]
        ent["enc_table"] = list(et.table)[:size]
        ent["symbol_tt"] = [[et.symbol_tt[i].bits, et.symbol_tt[i].find_state] for i in range(nh.table_len)]
        ent["dec_table"] = [[dt.table[i].new_state, dt.table[i].symbol, dt.table[i].num_bits] for i in range(size)]
    # python model must agree on norm/header/tables
    ph = M.Histogram(src).normalize(tl)
    assert ph.table[:256] == list(nh.table) and ph.log2 == nh.log2
    pv = M.Vec()
    assert ph.write(pv) == hbits and pv.bytes() == hdr
    pe = M.EncodeTable(ph)
    assert pe.table == list(et.table)[:size] and pe.symbols == list(et.symbols)[:size]
    pd = M.DecodeTable(ph)
    assert [x[0] for x in pd.table] == [dt.table[i].new_state for i in range(size)]
    ent["payload"] = {}
    for n_states in (1, 2, 4, 32, 64, 128):
        if len(src) < n_states:
            continue
        pay, pbits = O.encode_payload(et, src, n_states)
        ent["payload"][str(n_states)] = {"hex": pay.hex(), "bits": pbits}
    if table_log == 0:  # the literal reference drivers
        v = M.Vec()
        M.fse_compress(src, v)
        assert v.bytes() == hdr + bytes.fromhex(ent["payload"]["1"]["hex"]), name
        if len(src) >= 2:
            v = M.Vec()
            M.fse_compress2(src, v)
            assert v.bytes() == hdr + bytes.fromhex(ent["payload"]["2"]["hex"]), name
    return ent


def main():
    kats = []
    kats.append(one("KAT-A", [0, 0, 0, 0, 0, 0, 1, 1], note="SURVEY Appendix C, hand-derived"))
    kats.append(one("KAT-B", range(256), note="flat 256, fse.rs doc-test input"))
    kats.append(one("KAT-C", exp_dist(8), table_log=8, note="exp_dist log2=8, histogram.rs:621-638"))
    kats.append(one("KAT-D", O.generate("geo", 0xC0FFEE01, 1000).tobytes(), note="G_geo(0.2) first 1000 bytes, seed 0xC0FFEE01"))
    kats.append(one("KAT-E", O.generate("text", 0xC0FFEE02, 777).tobytes(), note="G_text first 777 bytes (odd length)"))
    kats.append(one("KAT-F", O.generate("few", 0xC0FFEE03, 513).tobytes(), table_log=9, note="G_few, tl=9, odd length"))
    kats.append(one("KAT-G", O.generate("uniform", 0xC0FFEE03, 2048).tobytes(), table_log=9,
                    note="uniform 2 KiB at tl=9"))
    gen = {k: O.generate(k, 0xC0FFEE00 + i, 32).tobytes().hex() for i, k in enumerate(["geo", "text", "few", "uniform"])}
    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump({"kats": kats, "generator_first32": gen}, f, indent=1)
    print("wrote", len(kats), "KATs")


if __name__ == "__main__":
    main()
