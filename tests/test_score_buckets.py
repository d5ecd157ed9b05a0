"""K4 gives a bucket only to the scores an entry can have (bwb_score_buckets, include/bwbble_b200.h).  The map is
computed on the host from the parameters alone; here it is checked against what the reference algorithm (the oracle's
heap_push, inexact_match.c:548-591) really pushes, over the golden parameter grid -- no GPU needed."""
import ctypes as C
import os

import numpy as np
import pytest

import golden_util as G
import oracle
from bwbble_b200 import _lib, default_params
from bwbble_b200.fastx import read_fastq


def buckets(**kw):
    p = default_params(**kw)
    nb = (p.max_diff + 1) * p.mm_score + (p.max_gapo + 1) * p.gapo_score + (p.max_gape + 1) * p.gape_score
    out = (C.c_uint8 * max(nb, 1))()
    n = _lib.lib().bwb_score_buckets(C.byref(p), out, nb)
    return n, nb, np.frombuffer(bytes(out), dtype=np.uint8)[:nb]


def test_default_parameters_need_19_of_68_buckets():
    n, nb, m = buckets(n=5)
    assert (nb, n) == (68, 19)
    reach = np.nonzero(m != 0xff)[0]
    # o = 0: 3m; o = 1: 11 + 3m + 4e with m + e <= 4
    exp = sorted({3 * a for a in range(6)} | {11 + 3 * a + 4 * e for a in range(5) for e in range(5 - a)})
    assert list(reach) == exp
    assert list(m[reach]) == list(range(n))              # ascending scores <-> ascending buckets


def test_bucket_limit_is_on_reachable_scores():
    # 26 mismatches at the default penalties: 139 buckets in the reference's heap, 27 + 26*... reachable ones on the device
    n, nb, _ = buckets(n=26)
    assert nb > 128 and 0 < n <= 128
    n, nb, _ = buckets(n=60, o=3, e=20)
    assert n > 128                                        # bwb_align refuses this one (BWB_ERR_UNSUPPORTED)
    p = default_params(n=-1)
    assert _lib.lib().bwb_score_buckets(C.byref(p), None, 0) < 0


def test_degenerate_penalties():
    n, nb, m = buckets(n=0)
    assert n == 1 and m[0] == 0
    n, nb, m = buckets(n=4, M=0)                          # mismatches are free: classes share buckets
    assert m[0] == 0 and n >= 2
    n, nb, m = buckets(n=4, E=0, o=2)
    assert m[0] == 0 and m[11] != 0xff and m[22] != 0xff


@pytest.fixture(scope="module")
def case(tmp_path_factory):
    fa = G.materialise_index(tmp_path_factory.mktemp("golden"))
    reads = read_fastq(os.path.join(G.GOLDEN, "r.fq"))
    orc = oracle.Oracle(fa + ".bwt")
    yield reads, orc
    orc.close()


@pytest.mark.parametrize("tag", sorted(G.grid()) + sorted(G.pgrid()))
def test_every_score_the_reference_pushes_has_a_bucket(case, tag):
    reads, orc = case
    kw = G.flags_to_kwargs(dict(G.grid(), **G.pgrid())[tag])
    kw.pop("n_threads", None)
    n, nb, m = buckets(**kw)
    oracle.pushed_scores(reset=True)
    orc.align(reads.seq, reads.offsets, default_params(**kw))
    pushed = oracle.pushed_scores(reset=True)
    assert pushed and max(pushed) < nb
    missing = sorted(s for s in pushed if m[s] == 0xff)
    assert not missing, "scores pushed by the reference algorithm without a device bucket: %s" % missing
