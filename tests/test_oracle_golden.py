"""The CPU restatement (oracle/) against the golden files produced by the real reference.
This is what pins the oracle on a box without /root/reference."""
import os

import numpy as np
import pytest

import golden_util as G
import oracle
from bwbble_b200 import default_params
from bwbble_b200.aln import first_difference, parse_aln
from bwbble_b200.fastx import read_fastq


@pytest.fixture(scope="module")
def case(tmp_path_factory):
    fa = G.materialise_index(tmp_path_factory.mktemp("golden"))
    reads = read_fastq(os.path.join(G.GOLDEN, "r.fq"))
    orc = oracle.Oracle(fa + ".bwt")
    yield fa, reads, orc
    orc.close()


@pytest.mark.parametrize("tag", sorted(G.grid()))
def test_oracle_reproduces_reference_aln(case, tag):
    fa, reads, orc = case
    kw = G.flags_to_kwargs(G.grid()[tag])
    threads = kw.pop("n_threads", 1)
    got, stats = orc.align(reads.seq, reads.offsets, default_params(**kw), threads=threads)
    exp = G.golden_bytes("aln_%s.aln" % tag)
    assert got == exp, "first difference (read, oracle, reference): %s" % (first_difference(got, exp),)
    assert stats["pops"] > 0


@pytest.mark.parametrize("tag", sorted(G.pgrid()))
def test_oracle_reproduces_reference_aln_with_precalc(case, tag):
    """-P: searches seeded from the 12-mer table, multi-genome and -S flavours"""
    fa, reads, orc = case
    kw = G.flags_to_kwargs(G.pgrid()[tag])
    got, stats = orc.align(reads.seq, reads.offsets, default_params(**kw))
    exp = G.golden_bytes("aln_%s.aln" % tag)
    assert got == exp, "first difference (read, oracle, reference): %s" % (first_difference(got, exp),)
    assert got != G.golden_bytes("aln_n3.aln")


@pytest.mark.parametrize("tag", sorted(G.mixed()))
def test_oracle_reproduces_reference_on_reads_shorter_than_the_seed(case, tag):
    """reads no longer than the seed consult the D_seed the previous longer read of their driver thread
    left behind -- serial driver, OpenMP driver (-t 3) and -P, as the reference wrote them for mixed.fq"""
    fa, _, orc = case
    reads = read_fastq(os.path.join(G.GOLDEN, "mixed.fq"))
    kw = G.flags_to_kwargs(G.mixed()[tag])
    threads = kw.pop("n_threads", 1)
    p = default_params(**kw)
    got, _ = orc.align(reads.seq, reads.offsets, p, threads=threads)
    exp = G.golden_bytes("aln_%s.aln" % tag)
    assert got == exp, "first difference (read, oracle, reference): %s" % (first_difference(got, exp),)
    lens = np.diff(reads.offsets.astype(np.int64))
    assert (lens <= p.seed_length).sum() > 50 and (lens > p.seed_length).sum() > 50


def test_mixed_length_goldens_depend_on_the_driver():
    """the fixture bites: serial and OpenMP drivers give different bytes for the same reads"""
    assert G.golden_bytes("aln_mixed_n3.aln") != G.golden_bytes("aln_mixed_n3_t3.aln")


def test_golden_fixture_exercises_the_hard_cases():
    """gapped hits (I and D), multiple hits per read, unmapped reads, reads with N, 36..150 bp."""
    hits = parse_aln(G.golden_bytes("aln_n4_o2_e3_k3_l20.aln"))
    states = {p & 3 for r in hits for h in r for p in h.pairs}
    assert states == {0, 1, 2}
    assert any(len(r) > 1 for r in hits) and any(len(r) == 0 for r in hits)
    reads = read_fastq(os.path.join(G.GOLDEN, "r.fq"))
    lens = [int(reads.offsets[i + 1] - reads.offsets[i]) for i in range(reads.n)]
    assert min(lens) <= 40 and max(lens) >= 140 and (reads.seq == 4).any()
    assert len(hits) == reads.n
    # tri-allelic codes (B, D, H, V) must be in the genome: they trigger quirk Q1 of O_alphabet
    fasta = G.golden_bytes("g.fa").decode()
    assert any(c in fasta for c in "BDHV")


def test_threaded_driver_equals_serial_on_golden():
    assert G.golden_bytes("aln_n3.aln") == G.golden_bytes("aln_n3_t4.aln")


@pytest.mark.parametrize("tag", ["sim_n0", "sim_n5"])
def test_oracle_on_the_reference_shipped_fastq(tmp_path, tag):
    """BASELINE configs[0]: test_data/sim_chr21_N100.fastq, the one input the reference ships, on the multi-genome its
    reads were planted into (tests/golden/make_golden.py --shipped-fastq): the reference's own .aln"""
    fa = G.materialise_index(tmp_path, "g21.fa")
    reads = read_fastq(os.path.join(G.GOLDEN, "sim_chr21_N100.fastq"))
    assert reads.n == 100
    orc = oracle.Oracle(fa + ".bwt")
    kw = G.flags_to_kwargs(G.MANIFEST["shipped"][tag])
    got, _ = orc.align(reads.seq, reads.offsets, default_params(**kw))
    orc.close()
    exp = G.golden_bytes("aln_%s.aln" % tag)
    assert got == exp, first_difference(got, exp)
    assert sum(1 for r in parse_aln(exp) if r) >= (90 if tag == "sim_n5" else 50)      # the reads do map
