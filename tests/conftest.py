import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def small_case(tmp_path_factory):
    """A ~60 kbp multi-genome (2 records, N run, SNPs incl. tri-allelic, bubbles, repeats) indexed
    with the product's host builder, plus mixed reads.  Shared by CPU and GPU tests."""
    from bwbble_b200 import synth, index
    d = tmp_path_factory.mktemp("small")
    g = synth.make_genome(7, 60000, n_records=2, n_bubbles=30, n_frac=0.04, n_repeat_copies=15, n_microsats=4,
                          lowercase_frac=0.01)
    fa = str(d / "g.fa")
    g.write_fasta(fa)
    index.build_index(fa)
    reads = synth.make_reads(g, 8, 600, 100, 3, indel_frac=0.15, n_base_frac=0.003)
    fq = str(d / "r.fq")
    reads.write_fastq(fq)
    return {"dir": str(d), "fasta": fa, "bwt": fa + ".bwt", "fastq": fq, "genome": g, "reads": reads}
