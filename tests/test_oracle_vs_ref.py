"""Differential test of the CPU restatement against the UNMODIFIED reference binary (oracle/_ref/bwbble,
compiled from /root/reference by `make -C oracle ref`).  Runs wherever that binary exists or can be
built; on a box without the reference tree and without the prebuilt binary it is skipped and the
golden files (test_oracle_golden.py) carry the pin."""
import os
import subprocess

import pytest

import oracle
from bwbble_b200 import build_index, default_params, synth
from bwbble_b200.aln import first_difference
from bwbble_b200.params import params_to_cli

REF = oracle.ensure_ref_binary()
pytestmark = pytest.mark.skipif(not REF, reason="no reference binary (oracle/_ref/bwbble) and no /root/reference to build it")


@pytest.fixture(scope="module")
def case(tmp_path_factory):
    d = tmp_path_factory.mktemp("vsref")
    g = synth.make_genome(31, 120000, n_records=3, snp_rate=0.015, tri_frac=0.06, n_bubbles=40, n_frac=0.04,
                          n_repeat_copies=30, n_microsats=6, lowercase_frac=0.01)
    fa = str(d / "g.fa")
    g.write_fasta(fa)
    subprocess.run([REF, "index", fa], check=True, stdout=subprocess.DEVNULL)
    reads = synth.make_reads(g, 32, 500, 100, 3, indel_frac=0.2, n_base_frac=0.003, bubble_frac=0.1)
    fq = str(d / "r.fq")
    reads.write_fastq(fq)
    return {"dir": str(d), "fasta": fa, "fastq": fq, "reads": reads}


def test_product_index_builder_equals_reference_index(case, tmp_path):
    fa2 = str(tmp_path / "copy.fa")
    open(fa2, "wb").write(open(case["fasta"], "rb").read())
    build_index(fa2)
    assert open(fa2 + ".bwt", "rb").read() == open(case["fasta"] + ".bwt", "rb").read()
    assert open(fa2 + ".ann", "rb").read() == open(case["fasta"] + ".ann", "rb").read()


GRID = [dict(n=0), dict(n=2), dict(n=3), dict(n=5), dict(n=4, o=2, e=3, k=3, l=20), dict(n=3, M=2, O=5, E=2),
        dict(n=4, l=0), dict(n=3, t=3, k=1), dict(n=6, o=2, M=4, O=4, E=4), dict(n=2, is_multiref=0),
        dict(n=4, M=0), dict(n=4, E=0, o=2), dict(n=3, m=150),
        dict(n=3, M=10, O=30, E=10), dict(n=4, o=2, e=4, M=7, O=23, E=9)]      # more than 128 buckets in heap_init


@pytest.mark.parametrize("kw", GRID, ids=lambda k: "-".join("%s%d" % kv for kv in k.items()))
def test_oracle_aln_is_byte_identical_to_reference(case, kw, tmp_path):
    p = default_params(**kw)
    out = str(tmp_path / "ref.aln")
    subprocess.run([REF, "align", *params_to_cli(p), case["fasta"], case["fastq"], out], check=True,
                   stdout=subprocess.DEVNULL)
    exp = open(out, "rb").read()
    orc = oracle.Oracle(case["fasta"] + ".bwt")
    got, _ = orc.align(case["reads"].seq, case["reads"].offsets, p, threads=p.n_threads)
    orc.close()
    assert got == exp, first_difference(got, exp)


def test_oracle_150bp_gapped_reads_config5_shape(case, tmp_path):
    """BASELINE configs[4] in miniature: 150 bp reads, up to 4 differences drawn from substitutions and 1-3 bp indels,
    `-n 4 -o 1 -e 6`, serial and threaded drivers"""
    g_reads = synth.make_reads(case["genome"] if "genome" in case else _genome(case), 33, 300, 150, 3, indel_frac=0.5,
                               max_indel=3, n_base_frac=0.002, bubble_frac=0.1)
    fq = str(tmp_path / "r150.fq")
    g_reads.write_fastq(fq)
    for threads in (1, 4):
        p = default_params(n=4, o=1, e=6, t=threads)
        out = str(tmp_path / ("ref%d.aln" % threads))
        subprocess.run([REF, "align", *params_to_cli(p), case["fasta"], fq, out], check=True, stdout=subprocess.DEVNULL)
        exp = open(out, "rb").read()
        orc = oracle.Oracle(case["fasta"] + ".bwt")
        got, st = orc.align(g_reads.seq, g_reads.offsets, p, threads=threads)
        orc.close()
        assert got == exp, first_difference(got, exp)
        assert st["hits"] > 100


def _genome(case):
    return synth.make_genome(31, 120000, n_records=3, snp_rate=0.015, tri_frac=0.06, n_bubbles=40, n_frac=0.04,
                             n_repeat_copies=30, n_microsats=6, lowercase_frac=0.01)


def mixed_length_reads(genome, seed=77, n=400):
    """reads of 14..60 bases: about 40 % are no longer than the default seed (32) and consult whatever D_seed the
    previous longer read of their thread left behind (inexact_match.c:36,62-64,121,141-143)"""
    return synth.make_reads(genome, seed, n, 60, 2, n_base_frac=0.004, bubble_frac=0.1, ragged=(14, 60))


@pytest.mark.parametrize("kw", [dict(n=3), dict(n=3, t=3), dict(n=2, k=1, l=40)],      # (-P: tests/golden/aln_mixed_P_n3.aln)
                         ids=lambda k: "-".join("%s%d" % kv for kv in k.items()))
def test_oracle_short_reads_inherit_the_previous_seed_bounds(case, kw, tmp_path):
    """Q6: for reads with len <= seed_length the reference does not recompute D_seed; serial and OpenMP drivers differ
    in which stale array such a read sees.  The oracle follows both, byte for byte."""
    reads = mixed_length_reads(_genome(case))
    fq = str(tmp_path / "mixed.fq")
    reads.write_fastq(fq)
    p = default_params(**kw)
    fa = case["fasta"]
    out = str(tmp_path / "ref.aln")
    subprocess.run([REF, "align", *params_to_cli(p), fa, fq, out], check=True, stdout=subprocess.DEVNULL)
    exp = open(out, "rb").read()
    orc = oracle.Oracle(fa + ".bwt")
    got, _ = orc.align(reads.seq, reads.offsets, p, threads=p.n_threads)
    # the test bites: aligned one at a time (fresh, all-zero D_seed) some short read comes out differently
    alone = b"".join(orc.align(reads.read(r), [0, len(reads.read(r))], p)[0] for r in range(reads.n))
    orc.close()
    assert got == exp, first_difference(got, exp)
    lens = [len(reads.read(r)) for r in range(reads.n)]
    assert min(lens) <= 20 and sum(1 for x in lens if x <= p.seed_length) > 50
    assert alone != exp
