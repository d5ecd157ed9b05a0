"""In-process multi-GPU sharding (bwb_create with several devices): the index is replicated, the reads of
one call are split in contiguous ranges, results come back in input order and are identical to the
single-device call.  Needs >= 2 visible GPUs (gpurun --gpus 2); skipped otherwise."""
import numpy as np
import pytest

from bwbble_b200 import Aligner, default_params

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("ndev", [2])
def test_multi_device_context_equals_single_device(small_case, ndev):
    reads = small_case["reads"]
    p = default_params(n=4)
    with Aligner([0], heap_pool_mb=512) as a1:
        a1.load_index(small_case["bwt"])
        one = a1.align(reads.seq, reads.offsets, p)
        exp, exp_counts = one.aln_bytes(), one.counts()
    with Aligner(list(range(ndev)), heap_pool_mb=512) as an:
        an.load_index(small_case["bwt"])
        res = an.align(reads.seq, reads.offsets, p)
        assert res.aln_bytes() == exp
        assert (res.counts() == exp_counts).all()
        ids = res.hits()["read_id"]
        assert (np.diff(ids.astype(np.int64)) >= 0).all() and ids.max() < reads.n      # global read ids, input order
        # an odd number of reads and a batch smaller than the device count
        sub = reads.slice(0, 1)
        assert an.align(sub.seq, sub.offsets, p).aln_bytes() == a1_bytes(small_case, sub, p)


def a1_bytes(case, sub, p):
    with Aligner([0], heap_pool_mb=512) as a1:
        a1.load_index(case["bwt"])
        return a1.align(sub.seq, sub.offsets, p).aln_bytes()
