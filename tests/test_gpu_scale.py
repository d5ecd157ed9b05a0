"""Parity at BASELINE scale through size-independent properties (the oracle cannot run millions of reads):
chr21-scale multi-genome (116 M BWT rows, index built on the device), 10^6 reads of bench.py's workload.
 * batching invariance: one launch == four quarter launches, byte for byte (per-read records are
   self-delimiting, so the .aln stream of a batch is the concatenation of its parts);
 * scheduling invariance: K3b's heavy-first queue order and the K0b table do not change a byte;
 * anchor: a 2048-read prefix equals the CPU oracle, and the totals are plausible for the workload
   (almost every read sampled from the haplotype maps; hits are in input order)."""
import os
import sys

import numpy as np
import pytest

import oracle
from bwbble_b200 import Aligner, default_params
from bwbble_b200.aln import first_difference

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def chr21_case():
    sys.path.insert(0, ROOT)
    import bench
    with Aligner([0]) as al:
        fa = bench.prepare_index("chr21", 0, lambda: None, aligner=al)      # cached per box
    old = bench.WORKLOADS["chr21"]["batch"]
    bench.WORKLOADS["chr21"]["batch"] = 1 << 20
    try:
        reads = bench.make_batch("chr21", 0, 0, 1)
    finally:
        bench.WORKLOADS["chr21"]["batch"] = old
    return fa, reads


def test_chr21_scale_invariants(chr21_case):
    fa, reads = chr21_case
    p = default_params(n=5)
    with Aligner([0]) as al:
        al.load_index(fa + ".bwt")
        whole = al.align(reads.seq, reads.offsets, p)
        blob = whole.aln_bytes()
        counts = whole.counts()
        ids = whole.hits()["read_id"]
        assert (np.diff(ids.astype(np.int64)) >= 0).all() and int(ids.max()) < reads.n
        assert 0.97 < float((counts > 0).mean()) <= 1.0
        ctr = whole.counters()
        whole.close()
        # four quarter launches
        q = reads.n // 4
        parts = []
        for k in range(4):
            sub = reads.slice(k * q, reads.n if k == 3 else (k + 1) * q)
            parts.append(al.align(sub.seq, sub.offsets, p).aln_bytes())
        assert b"".join(parts) == blob
        # the oracle on a prefix
        sub = reads.slice(0, 2048)
        orc = oracle.Oracle(fa + ".bwt")
        exp, _ = orc.align(sub.seq, sub.offsets, p, threads=os.cpu_count() or 1)
        orc.close()
        got = al.align(sub.seq, sub.offsets, p).aln_bytes()
        assert got == exp, first_difference(got, exp)
        assert blob[:len(exp)] == exp
    # input-order queue and no k-mer table: same bytes, same amount of search
    with Aligner([0]) as al:
        al.set_option("heavy_first", 2)
        al.set_option("kmer_table", 2)
        al.load_index(fa + ".bwt")
        res = al.align(reads.seq, reads.offsets, p)
        assert res.aln_bytes() == blob
        c2 = res.counters()
        assert c2["pops"] == ctr["pops"] and c2["pushes"] == ctr["pushes"]
