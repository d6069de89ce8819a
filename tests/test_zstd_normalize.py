"""the second normaliser, libzstd's FSE_normalizeCount.  No libzstd exports the function, so the oracle's
restatement is pinned here by vectors derived by hand from the published algorithm (each derivation is spelled out below)
and the GPU kernel is compared with the oracle on random histograms; tests/test_zstd_interop.py adds libzstd's own output
on real histograms (read back out of the frames it writes)."""
import numpy as np
import pytest

import oracle_lib as O


def znorm(counts, tl, low=True):
    c = list(counts) + [0] * (256 - len(counts))
    rc, n = O.normalize_zstd(O.hist_from_counts(c), tl, low)
    return rc, list(n.table)[:len(counts)], n.log2


def test_hand_derived_vectors():
    # (1) [6, 2] at table_log 5: step = 2^62 / 8 = 2^59, scale 57: 6 * 2^59 >> 57 = 24, 2 * 2^59 >> 57 = 8; 24 + 8 = 32:
    #     nothing left to distribute, 0 >= 24 >> 1 is false -> [24, 8]   (the same as the crate: SURVEY KAT-A)
    assert znorm([6, 2], 5) == (0, [24, 8], 5)
    # (2) [769, 1 x 255] at table_log 9 (total 1024, 256 symbols need FSE_minTableLog = 9): lowThreshold = 1024 >> 9 = 2, the
    #     255 ones -> lowProbCount (still = 512 - 255 = 257); symbol 0: floor(769 * 512 / 1024) = 384 -> still = -127;
    #     127 >= 384 >> 1 is false -> norm[0] = 384 - 127 = 257.
    assert znorm([769] + [1] * 255, 9, True) == (0, [257] + [-1] * 255, 9)
    assert znorm([769] + [1] * 255, 9, False) == (0, [257] + [1] * 255, 9)     # useLowProbCount = 0: +1 each
    # (3) FSE_normalizeM2.  [17 x 10, 5 x 6] at table_log 5 (total 200): lowThreshold = 6 -> the fives are low (-1, still 26);
    #     17 * 32 / 200 = 2.72: proba 2, remainder .72 > rtb[2] = 504333 / 2^20 = .481 -> 3 each, 30 in all -> still = -4;
    #     4 >= 3 >> 1 -> M2: lowOne = 600 >> 6 = 9; fives -1 (total left 170), seventeens unassigned; ToDistribute = 26;
    #     170 / 26 = 6 is not > 9; rStep = (26 * 2^57 + 2^56 - 1) / 170: every symbol advances the running total by
    #     2.65 * 2^57 from 0.5 * 2^57: 3.15 5.8 8.45 11.1 13.75 16.4 19.05 21.7 24.35 27 - eps -> floors 3 5 8 11 13 16 19 21 24 26
    #     -> weights 3 2 3 3 2 3 3 2 3 2 (sum 26).
    assert znorm([17] * 10 + [5] * 6, 5) == (0, [3, 2, 3, 3, 2, 3, 3, 2, 3, 2] + [-1] * 6, 5)
    # (4) one symbol holds everything: zstd's rle special case (returns 0, no table)
    assert znorm([0, 0, 9], 5)[0] == 3
    # (5) table_log limits: 0 = 11; > 12 tableLog_tooLarge; < 5 and < FSE_minTableLog GENERIC (oracle code -1)
    assert znorm([10, 20, 30, 40], 0)[2] == 11
    assert znorm([10, 20], 13)[0] == -3
    assert znorm([10, 20], 4)[0] == -1
    assert znorm(list(range(1, 201)), 7)[0] == -1            # 200 symbols need highbit(199) + 2 = 9 bits
    assert znorm([3] * 256, 8)[0] == -1                      # 256 symbols: FSE_minTableLog = min(10, 9) = 9
    # (6) an exact table: [2 x 16] at table_log 5 (total 32): proba 2 each, nothing left, 0 >= 2 >> 1 is false
    assert znorm([2] * 16, 5) == (0, [2] * 16, 5)


@pytest.mark.parametrize("low", [True, False])
def test_properties_on_random_histograms(low):
    rng = np.random.default_rng(7)
    done = 0
    for _ in range(400):
        nsym = int(rng.integers(2, 257))
        shape = rng.choice(["flat", "geo", "spiky"])
        if shape == "flat":
            c = rng.integers(0, 50, size=nsym)
        elif shape == "geo":
            c = (rng.geometric(0.02, size=nsym) * rng.integers(0, 2, size=nsym)).astype(np.int64)
        else:
            c = rng.integers(0, 3, size=nsym)
            c[rng.integers(0, nsym)] += int(rng.integers(100, 100000))
        if c.sum() < 2 or (c > 0).sum() < 2:
            continue
        c[-1] = max(c[-1], 1)
        for tl in (5, 8, 11, 12):
            rc, t, l2 = znorm(c.tolist(), tl, low)
            if rc < 0:
                continue
            assert sum(abs(x) for x in t) == 1 << l2            # the table is full
            assert all((x == 0) == (y == 0) for x, y in zip(t, c))   # zero <=> zero
            assert low or min(t) >= 0
            done += 1
    assert done > 300


@pytest.mark.gpu
def test_gpu_matches_the_oracle():
    import torch
    import entropy_coders_b200 as E
    ctx = E.Context(0)
    rng = np.random.default_rng(11)
    rows = []
    for _ in range(300):
        nsym = int(rng.integers(1, 257))
        c = np.zeros(256, dtype=np.int64)
        kind = rng.integers(0, 4)
        if kind == 0:
            c[:nsym] = rng.integers(0, 40, size=nsym)
        elif kind == 1:
            c[:nsym] = rng.geometric(0.01, size=nsym)
        elif kind == 2:
            c[:nsym] = rng.integers(0, 3, size=nsym)
            c[rng.integers(0, nsym)] += int(rng.integers(1000, 10 ** 7))
        else:
            c[rng.integers(0, 256)] = int(rng.integers(1, 1000))      # single symbol: rle
        rows.append(c)
    rows.append(np.array([769] + [1] * 255, dtype=np.int64))
    rows.append(np.array([17] * 10 + [5] * 6 + [0] * 240, dtype=np.int64))
    counts = torch.from_numpy(np.stack(rows)).to(ctx.device)
    for tl in (0, 5, 7, 9, 11, 12, 13):
        for low in (True, False):
            norm, log2, tlen, st = ctx.normalize_zstd(counts, tl, low)
            norm, st = norm.cpu().numpy(), st.cpu().numpy()
            for i, c in enumerate(rows):
                if c.sum() == 0:
                    assert st[i] == -8
                    continue
                rc, n = O.normalize_zstd(O.hist_from_counts(c.tolist()), tl, low)
                rc = {-1: -8}.get(rc, rc)                         # the oracle's PANIC code -> FSE_B200_ERR_PANIC
                assert st[i] == rc, (i, tl, low, st[i], rc)
                if rc >= 0:
                    assert list(norm[i]) == list(n.table), (i, tl, low)
    ctx.close()
