"""SURVEY Q6, host side: which read's D_seed a read no longer than the seed consults (bwb_seed_donor_plan = the plan
the device path follows, bwb_abi.cu plan_seed_donors) against a literal simulation of the reference's two drivers
(inexact_match.c:25-168).  No GPU needed; the device side is tests/test_gpu_short_reads.py."""
import ctypes as C

import numpy as np
import pytest

from bwbble_b200 import _lib, default_params

READ_BATCH = 0x40000      # align.h:14


def reference_drivers(lens, skipped, seed_len, n_threads, use_precalc, carry):
    n = len(lens)
    out = np.full(n, -1, dtype=np.int64)

    def run(lo, hi, d_seed):
        for i in range(lo, hi):
            if use_precalc and skipped[i]:
                continue                                  # `continue` before calculate_d (:50-57, :129-136)
            if lens[i] > seed_len:
                d_seed = i                                # calculate_d(..., D_seed) (:62-64, :141-143)
            out[i] = d_seed
        return d_seed

    if n_threads <= 1:                                    # align_reads_inexact: one D_seed for all batches (:36)
        run(0, n, -2 if carry else -1)
    else:                                                 # align_reads_inexact_parallel (:103-149)
        for w0 in range(0, n, READ_BATCH):
            bs = min(READ_BATCH, n - w0)
            for tid in range(n_threads):
                run(w0 + tid * bs // n_threads, w0 + (tid + 1) * bs // n_threads, -1)      # calloc per thread (:121)
    return out


def plan(lens, first12_bad, p, carry):
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    seq = np.zeros(int(off[-1]) + 16, dtype=np.uint8)
    for r in np.nonzero(first12_bad & (lens >= 12))[0]:
        seq[int(off[r]) + 5] = 4                          # an N among the first 12 bases
    out = np.zeros(len(lens), dtype=np.int64)
    rc = _lib.lib().bwb_seed_donor_plan(C.byref(p), seq.ctypes.data, off.ctypes.data, len(lens), int(carry), out.ctypes.data)
    assert rc == 0
    return out


@pytest.mark.parametrize("n", [0, 1, 2, 5, 1000, READ_BATCH + 12345, 2 * READ_BATCH + 3])
@pytest.mark.parametrize("threads,use_p,carry", [(1, 0, 0), (1, 0, 1), (1, 1, 1), (2, 0, 0), (3, 1, 0), (7, 0, 0), (16, 0, 1)])
def test_plan_equals_the_reference_drivers(n, threads, use_p, carry):
    rng = np.random.default_rng(n * 31 + threads)
    lens = rng.integers(8, 60, size=n).astype(np.int64)
    if n > 100:
        lens[rng.integers(0, n, size=n // 3)] = 20        # stretches of short reads, also across chunk starts
        lens[:3] = 20
    bad = rng.random(n) < 0.1
    skipped = bad | (lens < 12)
    p = default_params(n=3, t=threads, use_precalc=use_p)
    got = plan(lens, bad, p, carry)
    exp = reference_drivers(lens, skipped, p.seed_length, threads, use_p, carry)
    assert (got == exp).all(), np.nonzero(got != exp)[0][:10]


def test_seed_length_variants():
    lens = np.array([40, 10, 33, 32, 50], dtype=np.int64)
    none = np.zeros(5, dtype=bool)
    assert list(plan(lens, none, default_params(n=2), 0)) == [0, 0, 2, 2, 4]
    assert list(plan(lens, none, default_params(n=2, l=45), 0)) == [-1, -1, -1, -1, 4]
    assert list(plan(lens, none, default_params(n=2, l=0), 0)) == [0, 1, 2, 3, 4]           # no seed: nothing consulted
