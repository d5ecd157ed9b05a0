"""Rank primitives of the oracle against a naive counter (test layer 1)."""
import numpy as np
import pytest

import oracle
from bwbble_b200 import load_bwt


@pytest.fixture(scope="module")
def case(small_case):
    ix = load_bwt(small_case["bwt"])
    orc = oracle.Oracle(small_case["bwt"])
    yield ix, orc
    orc.close()


def _positions(ix):
    n = ix.length
    s = ix.sa0_index
    edge = [0, 1, 126, 127, 128, 129, 255, 256, n - 2, n - 1, -1, s - 1, s, s + 1, (s // 128) * 128, (s // 128) * 128 + 127]
    rng = np.random.default_rng(1)
    return [p for p in edge if -1 <= p < n] + rng.integers(0, n, size=300).tolist()


def test_O_is_the_inclusive_rank_without_the_sentinel_row(case):
    ix, orc = case
    sym = ix.symbols()
    for c in range(1, 16):
        pref = np.cumsum(sym == c)
        for p in _positions(ix):
            exp = 0 if p == -1 else int(pref[p])
            assert orc.O(c, p) == exp, (c, p)


def test_O_alphabet_quirk_for_triallelic_codes(case):
    """O_alphabet never counts codes 5, 9, 11, 13 inside the block nor their checkpoint, but applies
    the checkpoint-symbol decrement (bwt.c:423-437,780); the i==-1 / i==length-1 shortcuts are exact."""
    ix, orc = case
    sym = ix.symbols()
    C = ix.C.astype(np.int64)
    n = ix.length
    for inc in (0, 1):
        for p in _positions(ix):
            got = orc.O_alphabet(p, inc).astype(np.int64)
            for j in range(1, 16):
                if p == n - 1:
                    exp = C[j + 1] + inc
                elif p == -1:
                    exp = C[j] + inc
                elif j in (5, 9, 11, 13):
                    exp = C[j] + inc - int(sym[(p // 128) * 128] == j)
                else:
                    exp = C[j] + inc + int((sym[: p + 1] == j).sum())
                assert got[j] == exp, (p, inc, j)
