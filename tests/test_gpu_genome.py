"""BASELINE configs[3] for real: a >= 2^32-row index (synthetic 3.1 Gbp GRCh37-scale multi-genome, 6.9 G BWT rows)
built on the device by the chunked 64-bit suffix sorter (K7w), searched by the wide kernels, and compared byte for
byte with the UNMODIFIED reference binary (oracle/_ref/bwbble, uint64_t throughout: common.h:6, bwt.h:19-40,
bwt.c:348-372) on a 10^4-read prefix at -n 5.

Needs ~150 GB of device memory, ~60 GB of host memory, ~25 GB of scratch disk and about five minutes, so it only
runs when asked for:  BWBBLE_TEST_GENOME=1 python -m pytest tests/test_gpu_genome.py -m gpu
(the log of the round-2 run is kept in profiles/r02_test_gpu_genome.log)."""
import os
import subprocess
import sys

import pytest

from bwbble_b200 import Aligner, default_params
from bwbble_b200.aln import first_difference

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "bwbble")

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("BWBBLE_TEST_GENOME") != "1", reason="set BWBBLE_TEST_GENOME=1 (5 minutes, 150 GB of HBM)")]


def _free_device_gb():
    import torch
    free, _ = torch.cuda.mem_get_info(0)
    return free / 2**30


def test_genome_scale_index_and_search_equal_the_reference_binary(tmp_path):
    if _free_device_gb() < 150:
        pytest.skip("needs 150 GB of free device memory")
    if not os.path.exists(REF_BIN):
        pytest.skip("oracle/_ref/bwbble not built")
    sys.path.insert(0, ROOT)
    import bench
    from bwbble_b200 import synth
    import numpy as np
    with Aligner([0]) as al:
        fa = bench.prepare_index("genome", 0, lambda: None, aligner=al)          # K7w; cached per box
        hap = np.load(os.path.join(bench.CACHE, "genome", "hap.npy"), mmap_mode="r")
        reads = synth.make_reads(synth.Genome([], np.asarray(hap), [], 0), 20261017, 10_000, 100, 2, with_names=False, bubble_frac=0.0)
        al.load_index(fa + ".bwt")
        assert al.index_length() >= 1 << 32, "the point of this test is a >= 2^32-row index"
        res = al.align(reads.seq, reads.offsets, default_params(n=5))
        got = res.aln_bytes()
        ctr = res.counters()
    fq, out = str(tmp_path / "r.fq"), str(tmp_path / "ref.aln")
    reads.write_fastq(fq)
    subprocess.run([REF_BIN, "align", "-n", "5", "-t", str(os.cpu_count() or 1), fa, fq, out], check=True,
                   stdout=subprocess.DEVNULL, timeout=3000)
    exp = open(out, "rb").read()
    assert got == exp, first_difference(got, exp)
    assert ctr["max_heap"] > 1000 and ctr["pops"] > 10_000 * 1000          # a real search, not a trivial one
