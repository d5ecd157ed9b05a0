"""Multi-GPU: reads shard, the index is replicated, nothing is exchanged on the hot path.

One process per GPU (torchrun).  Rank r of W aligns the contiguous range shard_range(n, r, W) of the
batch -- the same static chunking the reference's OpenMP driver uses per thread
(inexact_match.c:115-116) -- and the per-shard `.aln` byte streams are concatenated in rank order, which
reproduces the single-GPU stream exactly (per-read records are self-delimiting, align.c:345-382).
The gather moves opaque byte blobs (a few tens of bytes per read) with torch.distributed: NCCL over
NVLink on the GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np


def shard_range(n_reads: int, rank: int, world: int) -> Tuple[int, int]:
    """[lo, hi) of rank `rank`: lo = rank*n/world (integer division), like chunk_start/chunk_end."""
    return rank * n_reads // world, (rank + 1) * n_reads // world


def gather_bytes(blob: bytes, dst: int = 0, device: Optional[str] = None) -> Optional[List[bytes]]:
    """Gather one byte blob per rank on `dst` (rank order).  Returns None on the other ranks."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    if device is None:
        device = "cuda" if dist.get_backend() == "nccl" else "cpu"
    sizes = torch.zeros(world, dtype=torch.int64, device=device)
    sizes[rank] = len(blob)
    dist.all_reduce(sizes)                                   # every rank learns every size
    mx = int(sizes.max().item())
    buf = torch.zeros(max(mx, 1), dtype=torch.uint8, device=device)
    if blob:
        buf[: len(blob)] = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(device)
    out = [torch.zeros_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, out, dst=dst)
    if rank != dst:
        return None
    return [bytes(out[r][: int(sizes[r])].cpu().numpy().tobytes()) for r in range(world)]


def seed_carry_read(seq: np.ndarray, offsets: np.ndarray, lo: int, seed_length: int,
                    use_precalc: bool = False) -> Optional[np.ndarray]:
    """The read whose D_seed the short reads at the start of shard [lo, ...) consult under the reference's SERIAL
    driver (inexact_match.c:36,62-64): the last read before `lo` that is longer than the seed -- and, with
    -P, was not skipped for an N among its first 12 bases (:50-57).  None if there is none.  Pass it to
    Aligner.set_seed_carry() before aligning the shard; the OpenMP driver (-t > 1) has no chain across 262144-read
    batches, so shards cut at multiples of that need nothing."""
    if seed_length <= 0 or lo <= 0:
        return None
    lens = np.diff(np.asarray(offsets[:lo + 1]).astype(np.int64))
    for r in np.nonzero(lens > seed_length)[0][::-1]:
        o = int(offsets[r])
        if use_precalc and (lens[r] < 12 or (np.asarray(seq[o:o + 12]) > 3).any()):
            continue
        return np.ascontiguousarray(seq[o:o + int(lens[r])], dtype=np.uint8)
    return None


def align_sharded(align_fn, seq: np.ndarray, offsets: np.ndarray, dst: int = 0, seed_length: int = 0,
                  use_precalc: bool = False) -> Optional[bytes]:
    """Every rank aligns its shard with `align_fn(seq_shard, offsets_shard) -> .aln bytes`; rank `dst`
    gets the whole batch's .aln stream in input order.  With seed_length > 0 the call is
    `align_fn(seq_shard, offsets_shard, carry)`, carry = seed_carry_read() of the shard (or None): what makes the
    sharded stream equal the serial reference's when reads no longer than the seed are present."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    n = len(offsets) - 1
    lo, hi = shard_range(n, rank, world)
    o = np.ascontiguousarray(offsets[lo:hi + 1])
    sub_seq = seq[int(o[0]):int(o[-1])]
    sub_off = (o - o[0]).astype(np.uint64)
    if seed_length > 0:
        blob = align_fn(sub_seq, sub_off, seed_carry_read(seq, offsets, lo, seed_length, use_precalc))
    else:
        blob = align_fn(sub_seq, sub_off)
    parts = gather_bytes(blob, dst)
    return None if parts is None else b"".join(parts)
