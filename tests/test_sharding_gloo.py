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


# Reads no longer than the seed: the shard's first short reads need the D_seed of the last longer read of the
# shards before it.  The oracle stands in for a GPU that was given the carry: the donor is aligned in front of the shard
# and its record dropped.
WORKER_MIXED = r"""
import os, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import numpy as np, torch.distributed as dist
import oracle
from bwbble_b200 import default_params
from bwbble_b200.aln import record_end
from bwbble_b200.dist import align_sharded, seed_carry_read
from bwbble_b200.fastx import read_fastq
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
reads = read_fastq(%(fq)r)
p = default_params(n=3)
orc = oracle.Oracle(%(bwt)r)
def fn(s, o, carry):
    if carry is None:
        return orc.align(s, o, p)[0]
    s2 = np.concatenate([carry, s]); o2 = np.concatenate([[0], np.asarray(o, dtype=np.uint64) + len(carry)]).astype(np.uint64)
    blob = orc.align(s2, o2, p)[0]
    return blob[record_end(blob, 0):]
whole = align_sharded(fn, reads.seq, reads.offsets, seed_length=p.seed_length)
naive = align_sharded(lambda s, o: orc.align(s, o, p)[0], reads.seq, reads.offsets)
if rank == 0:
    exp = open(%(exp)r, "rb").read()
    assert whole == exp, "sharded stream with carries differs from the reference's serial stream"
    print("OK", world, len(whole), "naive_equal=%%d" %% (naive == exp))
dist.destroy_process_group()
"""


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_alignment_of_mixed_lengths_carries_the_seed_donor(tmp_path, world):
    import golden_util as G
    fa = G.materialise_index(tmp_path)
    exp = tmp_path / "exp.aln"
    exp.write_bytes(G.golden_bytes("aln_mixed_n3.aln"))
    script = tmp_path / "worker.py"
    script.write_text(WORKER_MIXED % {"root": ROOT, "fq": os.path.join(G.GOLDEN, "mixed.fq"), "bwt": fa + ".bwt", "exp": str(exp)})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "OK %d" % world in r.stdout


def test_seed_carry_read_picks_the_last_longer_read():
    import numpy as np
    from bwbble_b200.dist import seed_carry_read
    lens = [40, 10, 50, 20, 12, 60]
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    seq = np.zeros(int(off[-1]), dtype=np.uint8)
    seq[int(off[2]) + 3] = 4                                  # an N among the first 12 bases of read 2
    assert seed_carry_read(seq, off, 0, 32) is None
    assert len(seed_carry_read(seq, off, 1, 32)) == 40
    assert len(seed_carry_read(seq, off, 5, 32)) == 50
    assert len(seed_carry_read(seq, off, 5, 32, use_precalc=True)) == 40       # -P skipped read 2: it left D_seed alone
    assert seed_carry_read(seq, off, 6, 64) is None and seed_carry_read(seq, off, 6, 0) is None


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
