"""SURVEY Q6 on the device (pytest -m gpu).  A read no longer than the seed (-l, default 32) gets no D_seed of its own
in the reference: it consults the array the previous longer read of its driver thread left behind -- the whole run for
the serial driver, the thread's static chunk of a 262144-read batch for the OpenMP one, and with -P a read skipped for
an N in its 12-mer leaves the array alone (inexact_match.c:36,50-64,115-143).  The device reproduces that (K3 computes
a short read's D_seed from its donor read, bwb_abi.cu seed_donors) and is held to the bytes the REFERENCE wrote for
tests/golden/mixed.fq (reads of 14..60 bases) and to the oracle on larger seeded cases."""
import os

import numpy as np
import pytest

import golden_util as G
import oracle
from bwbble_b200 import Aligner, default_params, synth
from bwbble_b200.aln import first_difference
from bwbble_b200.dist import seed_carry_read
from bwbble_b200.fastx import read_fastq

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["idx32", "idx64"])
def mixed(tmp_path_factory, request):
    d = tmp_path_factory.mktemp("mixed")
    fa = G.materialise_index(d)
    al = Aligner(heap_pool_mb=256)
    if request.param == "idx64":
        al.set_option("force_wide", 1)
    al.load_index(fa + ".bwt")
    reads = read_fastq(os.path.join(G.GOLDEN, "mixed.fq"))
    yield {"al": al, "fa": fa, "reads": reads, "dir": str(d)}
    al.close()


def _params(tag):
    kw = G.flags_to_kwargs(G.mixed()[tag])
    return default_params(**kw)


@pytest.mark.parametrize("tag", sorted(G.mixed()))
def test_one_call_equals_the_reference(mixed, tag):
    al, reads = mixed["al"], mixed["reads"]
    p = _params(tag)
    if p.use_precalc:
        al.build_precalc(True)
    got = al.align(reads.seq, reads.offsets, p).aln_bytes()
    exp = G.golden_bytes("aln_%s.aln" % tag)
    assert got == exp, "params %s: first difference (read, device, reference) %s" % (G.mixed()[tag], first_difference(got, exp))


def test_the_chain_runs_across_calls_of_a_run(mixed):
    """serial driver: option seed_carry (what bwb_align_fastq and the drop-in shim set) continues the chain over launches"""
    al, reads = mixed["al"], mixed["reads"]
    p = _params("mixed_n3")
    exp = G.golden_bytes("aln_mixed_n3.aln")
    for step in (1, 7, 64, 150):
        al.set_option("seed_carry", 1)                               # starts a run
        got = b""
        for lo in range(0, reads.n, step):
            sub = reads.slice(lo, min(lo + step, reads.n))
            got += al.align(sub.seq, sub.offsets, p).aln_bytes()
        al.set_option("seed_carry", 0)
        assert got == exp, "launches of %d reads: %s" % (step, first_difference(got, exp))
    # without the carry the launches are independent driver calls: a different (documented) result
    got = b"".join(al.align(reads.slice(lo, min(lo + 7, reads.n)).seq, reads.slice(lo, min(lo + 7, reads.n)).offsets, p).aln_bytes()
                   for lo in range(0, reads.n, 7))
    assert got != exp


def test_explicit_carry_for_a_shard(mixed):
    """what a process that aligns only reads [lo, n) of a run does (bwbble_b200/dist.py)"""
    al, reads = mixed["al"], mixed["reads"]
    p = _params("mixed_n3")
    exp = G.golden_bytes("aln_mixed_n3.aln")
    whole = b""
    for lo, hi in ((0, 133), (133, 266), (266, reads.n)):
        al.set_seed_carry(seed_carry_read(reads.seq, reads.offsets, lo, p.seed_length))
        sub = reads.slice(lo, hi)
        whole += al.align(sub.seq, sub.offsets, p).aln_bytes()
    al.set_option("seed_carry", 0)
    assert whole == exp, first_difference(whole, exp)


@pytest.mark.parametrize("batch", [0, 37])
@pytest.mark.parametrize("tag", ["mixed_n3", "mixed_n3_t3"])
def test_streaming_entry_point(mixed, tmp_path, tag, batch):
    """bwb_align_fastq: serial chain carried over its launches; with -t > 1 launches hold whole 262144-read batches"""
    al = mixed["al"]
    aln = str(tmp_path / "out.aln")
    n = al.align_fastq(os.path.join(G.GOLDEN, "mixed.fq"), aln, _params(tag), batch=batch)
    assert n == mixed["reads"].n
    got, exp = open(aln, "rb").read(), G.golden_bytes("aln_%s.aln" % tag)
    assert got == exp, first_difference(got, exp)


@pytest.mark.parametrize("threads", [1, 5])
def test_larger_seeded_case_against_the_oracle(small_case, threads):
    """20 000 reads of 12..70 bases with N bases, -l 32 and -l 40, serial and 5-thread chunking"""
    al = Aligner(heap_pool_mb=256)
    al.load_index(small_case["bwt"])
    orc = oracle.Oracle(small_case["bwt"])
    reads = synth.make_reads(small_case["genome"], 91, 20000, 70, 2, n_base_frac=0.004, ragged=(12, 70))
    for kw in (dict(n=2), dict(n=3, l=40, k=1)):
        p = default_params(t=threads, **kw)
        got = al.align(reads.seq, reads.offsets, p).aln_bytes()
        exp, _ = orc.align(reads.seq, reads.offsets, p, threads=threads)
        assert got == exp, "%s threads=%d: %s" % (kw, threads, first_difference(got, exp))
    orc.close()
    al.close()


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_donor_on_the_other_device(mixed):
    """in-process sharding: the first short reads of device 1's shard take their donor from device 0's"""
    reads = mixed["reads"]
    with Aligner([0, 1], heap_pool_mb=256) as al2:
        al2.load_index(mixed["fa"] + ".bwt")
        for tag in ("mixed_n3", "mixed_n3_t3"):
            got = al2.align(reads.seq, reads.offsets, _params(tag)).aln_bytes()
            exp = G.golden_bytes("aln_%s.aln" % tag)
            assert got == exp, first_difference(got, exp)
