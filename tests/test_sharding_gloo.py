"""Multi-process path on CPU: world_size 2 and 3 over gloo.  The per-shard function is the CPU oracle
(the checker standing in for a GPU); what is under test is the host logic: contiguous sharding and the
ordered gather reproduce the single-process .aln stream byte for byte."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import numpy as np, torch.distributed as dist
import oracle
from bwbble_b200 import default_params, synth
from bwbble_b200.dist import align_sharded, shard_range
from bwbble_b200.fastx import read_fastq
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
reads = read_fastq(%(fq)r)
p = default_params(n=3)
orc = oracle.Oracle(%(bwt)r)
whole = align_sharded(lambda s, o: orc.align(s, o, p)[0], reads.seq, reads.offsets)
if rank == 0:
    exp, _ = orc.align(reads.seq, reads.offsets, p)
    assert whole == exp, "sharded stream differs from the single-process stream"
    covered = [shard_range(reads.n, r, world) for r in range(world)]
    assert covered[0][0] == 0 and covered[-1][1] == reads.n and all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
    print("OK", world, len(whole))
else:
    assert whole is None
dist.destroy_process_group()
"""


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_alignment_over_gloo(small_case, tmp_path, world):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT, "fq": small_case["fastq"], "bwt": small_case["bwt"]})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), str(script)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "OK %d" % world in r.stdout


def test_shard_ranges_are_the_reference_chunks():
    from bwbble_b200.dist import shard_range
    for n in (0, 1, 7, 262144, 1000003):
        for w in (1, 2, 3, 4, 8):
            rs = [shard_range(n, r, w) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            for r, (lo, hi) in enumerate(rs):
                assert lo == r * n // w and hi == (r + 1) * n // w       # inexact_match.c:115-116
