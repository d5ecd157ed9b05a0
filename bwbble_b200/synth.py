"""Synthetic multi-genome + read generator (seeded, numpy).

The reference ships no genome (test_data/*.fasta are missing from the tree), so every
parity test and bench workload is generated.  The construction mimics what the reference's
offline tools produce (mg-ref/comb.cpp:70-168 print_multigenome, :211-276 print_bubble):

  * a linear genome of uniform random bases, split into one or more FASTA records,
    optionally with long N runs (GRCh37-like telomere/centromere gaps), planted repeat
    families and microsatellites (they create the long interval lists / deep heaps);
  * SNP sites: the alternate allele is OR-ed into the IUPAC mask of the site; a fraction of
    sites get a second alternate allele (-> B/D/H/V, the codes `O_alphabet` mishandles,
    bwt.c:423-437, so fixtures must contain them);
  * indel "bubbles": extra FASTA records = 124 bp left flank + alternate allele + 124 bp
    right flank (mg-ref/comb.cpp:339), appended after the main records;
  * reads sampled from a random haplotype of the multi-genome (each SNP site picks one of
    its alleles), with substitutions, optional single indels and optional N bases, 50 %
    reverse-complemented.

Read bases are returned in the reference's nt4 code (io.h:113-130): A=0 G=1 C=2 T=3 N=4.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

# IUPAC letter for bitmask A=8 C=4 G=2 T=1 (mask 0 never occurs)
_MASK2IUPAC = np.frombuffer(b"?TGKCYSBAWRDMHVN", dtype=np.uint8)
# internal base index 0..3 = A,C,G,T
_BASE_MASK = np.array([8, 4, 2, 1], dtype=np.uint8)
_BASE2NT4 = np.array([0, 2, 1, 3], dtype=np.uint8)       # A,C,G,T -> nt4 (A0 G1 C2 T3)
_NT4_COMPL = np.array([3, 2, 1, 0, 4], dtype=np.uint8)   # io.h:111
_NT4_CHARS = np.frombuffer(b"AGCTN", dtype=np.uint8)


@dataclass
class Genome:
    """A generated multi-genome: FASTA records (name, IUPAC letters as uint8)."""
    records: List[Tuple[str, np.ndarray]]
    hap: np.ndarray            # one haplotype of the main sequence, internal base idx 0..3, 4 = N
    bubbles: List[np.ndarray]  # haplotype sequence of each bubble record (base idx)
    seed: int = 0

    def write_fasta(self, path: str, width: int = 60) -> None:
        with open(path, "wb") as f:
            for name, seq in self.records:
                f.write(b">" + name.encode() + b"\n")
                n = len(seq)
                full = (n // width) * width
                if full:
                    body = np.empty((n // width, width + 1), dtype=np.uint8)
                    body[:, :width] = seq[:full].reshape(-1, width)
                    body[:, width] = 10
                    f.write(body.tobytes())
                if full < n:
                    f.write(seq[full:].tobytes() + b"\n")

    @property
    def total_bases(self) -> int:
        return int(sum(len(s) for _, s in self.records))


@dataclass
class Reads:
    """Packed reads: nt4 codes, fixed or ragged lengths."""
    seq: np.ndarray        # uint8, concatenated nt4 codes
    offsets: np.ndarray    # uint32/uint64, len n+1
    names: Optional[List[str]] = None
    meta: dict = field(default_factory=dict)

    @property
    def n(self) -> int:
        return len(self.offsets) - 1

    def read(self, i: int) -> np.ndarray:
        return self.seq[self.offsets[i]:self.offsets[i + 1]]

    def slice(self, lo: int, hi: int) -> "Reads":
        o = self.offsets[lo:hi + 1]
        return Reads(self.seq[o[0]:o[-1]], (o - o[0]).astype(self.offsets.dtype),
                     None if self.names is None else self.names[lo:hi])

    def write_fastq(self, path: str, lo: int = 0, hi: Optional[int] = None) -> None:
        hi = self.n if hi is None else hi
        lens = np.diff(self.offsets[lo:hi + 1].astype(np.int64))
        if self.names is None and hi - lo > 1000 and len(lens) and (lens == lens[0]).all():
            # millions of equal-length reads: one fixed-width record matrix ("@r<10 digits>\n<seq>\n+\n<qual>\n")
            n, L = hi - lo, int(lens[0])
            rec = np.empty((n, 2 + 10 + 1 + L + 3 + L + 1), dtype=np.uint8)
            rec[:, 0] = ord("@"); rec[:, 1] = ord("r")
            ids = np.arange(lo, hi, dtype=np.int64)
            for d in range(10):
                rec[:, 11 - d] = 48 + (ids // 10 ** d) % 10
            rec[:, 12] = 10
            base = int(self.offsets[lo])
            rec[:, 13:13 + L] = _NT4_CHARS[self.seq[base:base + n * L].reshape(n, L)]
            rec[:, 13 + L] = 10; rec[:, 14 + L] = ord("+"); rec[:, 15 + L] = 10
            rec[:, 16 + L:16 + 2 * L] = ord("2")
            rec[:, 16 + 2 * L] = 10
            rec.tofile(path)
            return
        with open(path, "wb") as f:
            for i in range(lo, hi):
                s = _NT4_CHARS[self.read(i)].tobytes()
                name = self.names[i] if self.names is not None else "r%d" % i
                f.write(b"@" + name.encode() + b"\n" + s + b"\n+\n" + b"2" * len(s) + b"\n")


def make_genome(seed: int, n_bases: int, n_records: int = 1, snp_rate: float = 0.012,
                tri_frac: float = 0.03, n_bubbles: int = 0, n_frac: float = 0.0,
                n_repeat_copies: int = 0, repeat_len: int = 300, repeat_div: float = 0.10,
                n_microsats: int = 0, lowercase_frac: float = 0.0) -> Genome:
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 4, size=n_bases, dtype=np.uint8)

    # planted repeat family + microsatellites (long interval lists, deep heaps)
    if n_repeat_copies:
        elem = rng.integers(0, 4, size=repeat_len, dtype=np.uint8)
        starts = rng.integers(0, n_bases - repeat_len, size=n_repeat_copies)
        for s in starts:
            cp = elem.copy()
            mut = rng.random(repeat_len) < repeat_div
            cp[mut] = (cp[mut] + rng.integers(1, 4, size=int(mut.sum()))) % 4
            base[s:s + repeat_len] = cp
    for _ in range(n_microsats):
        unit = rng.integers(0, 4, size=int(rng.integers(1, 5)), dtype=np.uint8)
        ln = int(rng.integers(40, 400))
        s = int(rng.integers(0, n_bases - ln))
        base[s:s + ln] = np.resize(unit, ln)

    mask = _BASE_MASK[base].copy()
    is_n = np.zeros(n_bases, dtype=bool)
    if n_frac > 0:
        # one leading run (2/3 of the N budget) and one interior block (1/3), like GRCh37 chr21
        n_tot = int(n_bases * n_frac)
        lead = (2 * n_tot) // 3
        is_n[:lead] = True
        blk = n_tot - lead
        s = int(rng.integers(lead + n_bases // 10, n_bases - blk - 1))
        is_n[s:s + blk] = True

    # SNP sites
    n_snp = int(n_bases * snp_rate)
    hap = base.copy()
    if n_snp:
        sites = rng.choice(n_bases, size=n_snp, replace=False)
        alt = rng.integers(0, 4, size=n_snp, dtype=np.uint8)     # 1/4 are no-ops
        mask[sites] |= _BASE_MASK[alt]
        tri = sites[rng.random(n_snp) < tri_frac]
        alt2 = rng.integers(0, 4, size=len(tri), dtype=np.uint8)
        mask[tri] |= _BASE_MASK[alt2]
        # haplotype: half of the SNP sites carry the alternate allele
        take = rng.random(n_snp) < 0.5
        hap[sites[take]] = alt[take]
    mask[is_n] = 15
    hap[is_n] = 4

    letters = _MASK2IUPAC[mask]
    if lowercase_frac > 0:
        # soft-masked stretches; the host upper-cases them (io.c:252-254)
        n_low = max(1, int(n_bases * lowercase_frac / 500))
        for s in rng.integers(0, max(1, n_bases - 500), size=n_low):
            letters[s:s + 500] |= 0x20

    # split into records
    cuts = [0]
    if n_records > 1:
        inner = np.sort(rng.choice(np.arange(1, n_bases), size=n_records - 1, replace=False))
        cuts += [int(c) for c in inner]
    cuts.append(n_bases)
    records = [("chr%d synthetic seed=%d" % (i + 1, seed), letters[cuts[i]:cuts[i + 1]])
               for i in range(n_records)]

    # indel bubbles: flank(124) + alt allele + flank(124)   (mg-ref/comb.cpp:339)
    bubbles = []
    FL = 124
    if n_bubbles:
        ok_lo, ok_hi = FL + 1, n_bases - FL - 8
        ps = rng.integers(ok_lo, ok_hi, size=n_bubbles)
        for b, p in enumerate(ps):
            p = int(p)
            k = int(rng.integers(1, 6))
            if rng.random() < 0.5:   # insertion of k bases after p
                ins = rng.integers(0, 4, size=k, dtype=np.uint8)
                seq = np.concatenate([letters[p - FL:p], _MASK2IUPAC[_BASE_MASK[ins]],
                                      letters[p:p + FL]])
                hp = np.concatenate([hap[p - FL:p], ins, hap[p:p + FL]])
            else:                    # deletion of k bases at p
                seq = np.concatenate([letters[p - FL:p], letters[p + k:p + k + FL]])
                hp = np.concatenate([hap[p - FL:p], hap[p + k:p + k + FL]])
            records.append(("bubble%d chr %d" % (b, p - FL), seq & 0xDF))
            bubbles.append(hp)
    return Genome(records, hap, bubbles, seed)


def make_reads(genome: Genome, seed: int, n_reads: int, read_len: int = 100, max_sub: int = 2,
               indel_frac: float = 0.0, max_indel: int = 3, n_base_frac: float = 0.0,
               bubble_frac: float = 0.02, with_names: bool = True,
               ragged: Optional[Tuple[int, int]] = None) -> Reads:
    """Sample reads from the haplotype.  `ragged=(lo,hi)` draws lengths uniformly in [lo,hi]."""
    rng = np.random.default_rng(seed)
    hap = genome.hap
    n = len(hap)
    L = read_len if ragged is None else ragged[1]
    W = L + max_indel + 1
    # positions whose window holds no N (rejection sample)
    pos = np.empty(n_reads, dtype=np.int64)
    filled = 0
    isn = (hap == 4)
    csum = np.concatenate([[0], np.cumsum(isn, dtype=np.int64)])
    while filled < n_reads:
        cand = rng.integers(0, n - W, size=int((n_reads - filled) * 1.6) + 16)
        good = cand[(csum[cand + W] - csum[cand]) == 0]
        take = min(len(good), n_reads - filled)
        pos[filled:filled + take] = good[:take]
        filled += take
    # (n_reads, W) base idx; gathered in chunks: the int64 index matrix of a whole 8 M-read batch would be 7 GB
    win = np.empty((n_reads, W), dtype=hap.dtype)
    arW = np.arange(W)[None, :]
    for lo in range(0, n_reads, 1 << 19):
        win[lo:lo + (1 << 19)] = hap[pos[lo:lo + (1 << 19), None] + arW]

    # reads drawn from bubble haplotypes (exercise the bubble records)
    if genome.bubbles and bubble_frac > 0:
        nb = int(n_reads * bubble_frac)
        which = rng.integers(0, len(genome.bubbles), size=nb)
        rows = rng.choice(n_reads, size=nb, replace=False)
        for r, b in zip(rows, which):
            hb = genome.bubbles[b]
            if len(hb) >= W and not (hb == 4).any():
                s = int(rng.integers(0, len(hb) - W + 1))
                win[r] = hb[s:s + W]
                pos[r] = -1 - b

    # single indel per read (fraction indel_frac)
    has_indel = rng.random(n_reads) < indel_frac
    if not has_indel.any():
        idx = ins_mask = None                                  # identity gather: reads = win[:, :L]
    else:
        idx = np.broadcast_to(np.arange(L)[None, :], (n_reads, L)).copy()
        ins_mask = np.zeros((n_reads, L), dtype=bool)
    if has_indel.any():
        d = rng.integers(1, max_indel + 1, size=n_reads)
        o = rng.integers(8, max(9, L - 8 - max_indel), size=n_reads)
        is_del = rng.random(n_reads) < 0.5
        ar = np.arange(L)[None, :]
        sel_del = (has_indel & is_del)[:, None]
        sel_ins = (has_indel & ~is_del)[:, None]
        idx = np.where(sel_del, ar + (ar >= o[:, None]) * d[:, None], idx)
        shift = np.clip(ar - o[:, None], 0, d[:, None])
        idx = np.where(sel_ins, ar - shift, idx)
        ins_mask = sel_ins & (ar >= o[:, None]) & (ar < (o + d)[:, None])
    reads = win[:, :L].copy() if idx is None else np.take_along_axis(win, idx, axis=1)
    if ins_mask is not None and ins_mask.any():
        reads[ins_mask] = rng.integers(0, 4, size=int(ins_mask.sum()), dtype=np.uint8)

    # substitutions: k uniform in [0, max_sub]
    k = rng.integers(0, max_sub + 1, size=n_reads)
    for s in range(max_sub):
        rows = np.nonzero(k > s)[0]
        cols = rng.integers(0, L, size=len(rows))
        reads[rows, cols] = (reads[rows, cols] + rng.integers(1, 4, size=len(rows))) % 4

    nt4 = _BASE2NT4[reads]
    if n_base_frac > 0:
        nt4[rng.random(nt4.shape) < n_base_frac] = 4
    # strand
    rev = rng.random(n_reads) < 0.5
    rc = _NT4_COMPL[nt4[:, ::-1]]
    nt4 = np.where(rev[:, None], rc, nt4)

    if ragged is None:
        lens = np.full(n_reads, L, dtype=np.int64)
        seq = np.ascontiguousarray(nt4).reshape(-1)
    else:
        lens = rng.integers(ragged[0], ragged[1] + 1, size=n_reads)
        keep = np.arange(L)[None, :] < lens[:, None]
        seq = nt4[keep]
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    names = None
    if with_names:
        names = ["r%d_%d_%s_%d" % (i, int(pos[i]), "rc" if rev[i] else "nm", int(k[i]))
                 for i in range(n_reads)]
    return Reads(np.ascontiguousarray(seq, dtype=np.uint8), offsets, names,
                 {"seed": seed, "read_len": read_len, "max_sub": max_sub})
