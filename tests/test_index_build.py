"""Host index builder (bwbble_b200/csrc/index_build.cpp) against the reference's files and a naive SA."""
import gzip
import os

import numpy as np
import pytest

import golden_util as G
from bwbble_b200 import build_index, load_bwt

NT16 = {"$": 0, "T": 1, "K": 2, "G": 3, "S": 4, "B": 5, "Y": 6, "C": 7, "M": 8, "H": 9, "N": 10, "V": 11, "R": 12,
        "D": 13, "W": 14, "A": 15}
COMPL = [0, 15, 8, 7, 4, 11, 12, 3, 2, 13, 10, 5, 6, 9, 14, 1]


def test_builder_reproduces_reference_index_files(tmp_path):
    fa = str(tmp_path / "g.fa")
    open(fa, "wb").write(G.golden_bytes("g.fa"))
    build_index(fa)
    assert open(fa + ".bwt", "rb").read() == gzip.decompress(G.golden_bytes("g.fa.bwt.gz"))
    assert open(fa + ".ann", "rb").read() == G.golden_bytes("g.fa.ann")


def _naive_index(records):
    fwd = []
    for seq in records:
        fwd += [NT16.get(c, 10) for c in seq.upper()] + [0]
    text = fwd + [COMPL[c] for c in reversed(fwd)]
    n = len(text)
    sa = sorted(range(n + 1), key=lambda i: text[i:])        # shorter suffix first = sentinel smallest
    bwt = [0 if p == 0 else text[p - 1] for p in sa]
    return text, sa, bwt


@pytest.mark.parametrize("seed", range(6))
def test_builder_matches_naive_suffix_sort(tmp_path, seed):
    rng = np.random.default_rng(seed)
    alphabet = "ACGT" if seed % 2 else "ACGTNRYKMSWBDHV"
    records = ["".join(rng.choice(list(alphabet), size=int(rng.integers(1, 300)))) for _ in range(int(rng.integers(1, 4)))]
    if seed == 3:
        records = ["A" * 200, "ACAC" * 50]                                    # runs and periodic text
    fa = str(tmp_path / "t.fa")
    with open(fa, "w") as f:
        for i, r in enumerate(records):
            f.write(">r%d some description\n" % i)
            for k in range(0, len(r), 50):
                f.write((r[k:k + 50].lower() if (i + k) % 3 == 0 else r[k:k + 50]) + "\n")
    build_index(fa)
    ix = load_bwt(fa + ".bwt", load_sa=True)
    text, sa, bwt = _naive_index(records)
    assert ix.length == len(text) + 1
    assert ix.sa0_index == sa.index(0)
    assert ix.symbols().tolist() == bwt
    assert ix.SA.tolist() == [sa[i] for i in range(0, len(sa), 32)]
    counts = np.bincount([b for i, b in enumerate(bwt) if i != ix.sa0_index], minlength=16)
    assert ix.C.tolist() == [0] + np.cumsum(counts).tolist()
    # checkpoint rows are inclusive of row 128k and skip the sentinel row (bwt.c:280-291)
    O = ix.O.reshape(-1, 16)
    run = np.zeros(16, dtype=np.int64)
    for i, b in enumerate(bwt):
        if i != ix.sa0_index:
            run[b] += 1
        if i % 128 == 0:
            assert O[i // 128].tolist() == run.tolist()
    lines = open(fa + ".ann").read().splitlines()
    assert lines[0] == "%d\t%d" % (len(text) // 2, len(records))
    assert lines[1].startswith("r0 some description\t0\t")
