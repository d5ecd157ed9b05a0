"""Host-side mirror of the reference's operator interface for the hot path.

Reference interface                      -> here
  load_bwt (bwt.c:90-125)                -> Aligner.load_index / Aligner.upload_index
  O / O_alphabet (bwt.c:348-438)         -> Aligner.occ / Aligner.occ_alphabet           (K1)
  exact_match (exact_match.c:58-60)      -> Aligner.exact_match                          (K2)
  calculate_d (inexact_match.c:171-254)  -> Aligner.calculate_d                          (K3)
  align_reads_inexact{,_parallel}        -> Aligner.align -> AlignResult.aln_bytes()     (K4+K5)
     (inexact_match.c:25-168)               (= the bytes alns2alnf_bin appends to the .aln file)
  align_reads (align.c:40-87)            -> align_reads(fasta, fastq, aln_out, params)

Everything runs through the C ABI of libbwbble_b200.so; there is no Python compute path.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import Hit, Params
from .index import BwtIndex, load_bwt
from .params import default_params


def _u8(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint8)


def _u64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint64)


class AlignResult:
    """Owner of a bwb_results*."""

    def __init__(self, aligner: "Aligner", handle: int):
        self._a = aligner
        self._h = C.c_void_p(handle)

    def close(self):
        if self._h and _lib is not None and _lib.lib is not None:
            _lib.lib().bwb_results_free(self._h)
            self._h = None

    __del__ = close

    def fetch(self):
        _lib.check(_lib.lib().bwb_results_fetch(self._h), self._a._ctx)
        return self

    @property
    def num_reads(self) -> int:
        return int(_lib.lib().bwb_results_num_reads(self._h))

    @property
    def num_hits(self) -> int:
        return int(_lib.lib().bwb_results_num_hits(self._h))

    def counts(self) -> np.ndarray:
        p = _lib.lib().bwb_results_counts(self._h)
        n = self.num_reads
        if not p or n == 0:
            return np.zeros(n, dtype=np.uint32)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint32)), shape=(n,)).copy()

    def hits(self) -> np.ndarray:
        """Structured array view of the bwb_hit records (copied)."""
        n = self.num_hits
        dt = np.dtype([("L", "<u8"), ("U", "<u8"), ("score", "<i4"), ("num_mm", "u1"), ("num_gapo", "u1"),
                       ("num_gape", "u1"), ("aln_length", "u1"), ("n_runs", "u1"), ("pad", "u1", 3),
                       ("read_id", "<u4"), ("runs", "u1", 16)])
        assert dt.itemsize == C.sizeof(Hit)
        p = _lib.lib().bwb_results_hits(self._h)
        if not p or n == 0:
            return np.zeros(0, dtype=dt)
        raw = C.string_at(p, n * dt.itemsize)
        return np.frombuffer(raw, dtype=dt).copy()

    def counters(self) -> dict:
        arr = (C.c_uint64 * 8)()
        _lib.check(_lib.lib().bwb_results_counters(self._h, C.byref(arr)), self._a._ctx)
        names = ["pops", "pushes", "exact_tails", "rank_queries", "max_heap", "max_list", "deferred_pass1", "deferred_pass2"]
        return {k: int(arr[i]) for i, k in enumerate(names)}

    @property
    def kernel_ms(self) -> float:
        return float(_lib.lib().bwb_results_kernel_ms(self._h))

    @property
    def k3_ms(self) -> float:
        return float(_lib.lib().bwb_results_k3_ms(self._h))

    def aln_bytes(self) -> bytes:
        buf = C.c_void_p()
        ln = C.c_uint64()
        _lib.check(_lib.lib().bwb_results_aln_bytes(self._h, C.byref(buf), C.byref(ln)), self._a._ctx)
        try:
            return C.string_at(buf, ln.value)
        finally:
            _lib.lib().bwb_free(buf)

    def locations(self) -> np.ndarray:
        """K6 output per read: ref_pos (SA of hit 0's L; 2^64-1 if unmapped), top1, top2."""
        dt = np.dtype([("ref_pos", "<u8"), ("top1", "<i4"), ("top2", "<i4")])
        p = _lib.lib().bwb_results_locations(self._h)
        n = self.num_reads
        if not p:
            raise _lib.BwbError(-4, "no sampled SA uploaded (Aligner.load_index(..., with_sa=True))")
        return np.frombuffer(C.string_at(p, n * dt.itemsize), dtype=dt).copy()

    def write_sam(self, sam_path: str, ann_path: str, names, seq, offsets, quals=None, max_mm: int = 6,
                  header: bool = True, append: bool = False):
        """`bwbble aln2sam -n max_mm` for these reads (align.c:494-652)."""
        seq, offsets = _u8(seq), _u64(offsets)
        n = len(offsets) - 1
        nm = (C.c_char_p * n)(*[x.encode() for x in names])
        ql = None
        if quals is not None:
            ql = (C.c_char_p * n)(*[x.encode() for x in quals])
        _lib.check(_lib.lib().bwb_results_write_sam(self._h, os.fsencode(ann_path), nm, seq.ctypes.data, offsets.ctypes.data,
                                                    ql, self._a.index.length, int(max_mm), os.fsencode(sam_path),
                                                    int(header), int(append)), self._a._ctx)

    def write_aln(self, path: str, append: bool = False):
        _lib.check(_lib.lib().bwb_results_write_aln(self._h, os.fsencode(path), int(append)), self._a._ctx)


class DeviceReads:
    """Owner of a bwb_reads* (reads resident in HBM)."""

    def __init__(self, aligner: "Aligner", handle: int, n: int):
        self._a = aligner
        self._h = C.c_void_p(handle)
        self.n = n

    def close(self):
        if self._h and _lib is not None and _lib.lib is not None:
            _lib.lib().bwb_reads_free(self._h)
            self._h = None

    __del__ = close


class Aligner:
    """A bwb_ctx: devices + replicated index + search scratch."""

    def __init__(self, devices: Optional[Sequence[int]] = None, **options):
        L = _lib.lib()
        if devices is None:
            self._ctx = L.bwb_create(None, 0)
        else:
            arr = (C.c_int * len(devices))(*devices)
            self._ctx = L.bwb_create(arr, len(devices))
        if not self._ctx:
            msg = L.bwb_last_error(None)
            raise _lib.BwbError(-2, msg.decode() if msg else "bwb_create failed")
        self._ctx = C.c_void_p(self._ctx)
        for k, v in options.items():
            self.set_option(k, v)
        self.index: Optional[BwtIndex] = None

    def close(self):
        if getattr(self, "_ctx", None) and _lib is not None and _lib.lib is not None:
            # device-resident read sets point into the context: release them first
            for ref in list(getattr(self, "_children", ())):
                obj = ref()
                if obj is not None:
                    obj.close()
            self._children = []
            _lib.lib().bwb_destroy(self._ctx)
            self._ctx = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_option(self, key: str, value: int):
        _lib.check(_lib.lib().bwb_set_option(self._ctx, key.encode(), int(value)), self._ctx)

    def set_stream(self, cuda_stream: int, dev_slot: int = 0):
        """Launch on a caller-owned stream, e.g. torch.cuda.current_stream().cuda_stream."""
        _lib.check(_lib.lib().bwb_set_stream(self._ctx, dev_slot, C.c_void_p(cuda_stream)), self._ctx)

    # ---- index -----------------------------------------------------------------------------
    def upload_index(self, ix: BwtIndex):
        C17 = _u64(ix.C)
        bwt = np.ascontiguousarray(ix.bwt, dtype=np.uint32)
        O = _u64(ix.O)
        _lib.check(_lib.lib().bwb_index_upload(self._ctx, ix.length, ix.sa0_index, C17.ctypes.data, bwt.ctypes.data,
                                               len(bwt), O.ctypes.data, len(O) // 16), self._ctx)
        self.index = ix

    def index_length(self) -> int:
        """BWT rows of the uploaded index (0 before an index is loaded)."""
        return int(self.index.length) if self.index is not None else 0

    def load_index(self, bwt_path: str, with_sa: bool = False):
        """load_bwt(path, loadSA) (bwt.c:90-125); with_sa also uploads the sampled SA, enabling K6."""
        ix = load_bwt(bwt_path, load_sa=with_sa)
        self.upload_index(ix)
        if with_sa:
            sa = _u64(ix.SA)
            _lib.check(_lib.lib().bwb_sa_upload(self._ctx, sa.ctypes.data, len(sa)), self._ctx)

    # ---- -P seed table (align.c:200-238) ---------------------------------------------------------
    def build_precalc(self, is_multiref: bool = True):
        """precalc_sa_intervals on the device (K0c): exact_match() of all 4^12 12-mers."""
        _lib.check(_lib.lib().bwb_precalc_build(self._ctx, int(bool(is_multiref))), self._ctx)

    def load_precalc(self, pre_path: str, is_multiref: bool = True):
        """load_precalc_sa_intervals(<fasta>.pre)"""
        _lib.check(_lib.lib().bwb_precalc_load_file(self._ctx, pre_path.encode(), int(bool(is_multiref))), self._ctx)

    def upload_precalc(self, sizes, intervals_lu, is_multiref: bool = True):
        sizes = np.ascontiguousarray(sizes, dtype=np.int32)
        lu = _u64(np.asarray(intervals_lu).reshape(-1))
        if len(sizes) != 1 << 24:
            raise ValueError("the seed table has 4^12 rows")
        _lib.check(_lib.lib().bwb_precalc_upload(self._ctx, sizes.ctypes.data, lu.ctypes.data, len(lu) // 2,
                                                 int(bool(is_multiref))), self._ctx)

    def write_precalc(self, pre_path: str):
        """store the table in the reference's .pre layout (align.c:144-152,217-222)"""
        _lib.check(_lib.lib().bwb_precalc_write(self._ctx, pre_path.encode()), self._ctx)

    def precalc_num_intervals(self) -> int:
        return int(_lib.lib().bwb_precalc_num_intervals(self._ctx))

    def precalc_row(self, row: int) -> np.ndarray:
        n = C.c_uint32()
        _lib.check(_lib.lib().bwb_precalc_row(self._ctx, row, None, 0, C.byref(n)), self._ctx)
        out = np.zeros((n.value, 2), dtype=np.uint64)
        if n.value:
            _lib.check(_lib.lib().bwb_precalc_row(self._ctx, row, out.ctypes.data, n.value, C.byref(n)), self._ctx)
        return out

    def download_blocks(self) -> np.ndarray:
        n = int(_lib.lib().bwb_index_num_blocks(self._ctx))
        out = np.zeros((n, 32), dtype=np.uint32)
        _lib.check(_lib.lib().bwb_index_download_blocks(self._ctx, out.ctypes.data), self._ctx)
        return out

    # ---- K1 ----------------------------------------------------------------------------------
    def occ(self, codes, pos) -> np.ndarray:
        codes, pos = _u8(codes), _u64(pos)
        out = np.zeros(len(pos), dtype=np.uint64)
        _lib.check(_lib.lib().bwb_occ(self._ctx, codes.ctypes.data, pos.ctypes.data, len(pos), out.ctypes.data), self._ctx)
        return out

    def occ_alphabet(self, pos, inc: int) -> np.ndarray:
        pos = _u64(pos)
        out = np.zeros((len(pos), 16), dtype=np.uint64)
        _lib.check(_lib.lib().bwb_occ_alphabet(self._ctx, pos.ctypes.data, len(pos), inc, out.ctypes.data), self._ctx)
        return out

    def occ_bench(self, n: int, seed: int = 1, mode: int = 0, iters: int = 10):
        ms = C.c_float()
        ck = C.c_uint64()
        _lib.check(_lib.lib().bwb_occ_bench(self._ctx, n, seed, mode, iters, C.byref(ms), C.byref(ck)), self._ctx)
        return float(ms.value), int(ck.value)

    # ---- K2 / K3 -------------------------------------------------------------------------------
    def exact_match(self, seq, offsets) -> List[np.ndarray]:
        seq, offsets = _u8(seq), _u64(offsets)
        n = len(offsets) - 1
        counts = np.zeros(n, dtype=np.uint32)
        buf = C.c_void_p()
        tot = C.c_uint64()
        _lib.check(_lib.lib().bwb_exact_match(self._ctx, seq.ctypes.data, offsets.ctypes.data, n, counts.ctypes.data,
                                              C.byref(buf), C.byref(tot)), self._ctx)
        try:
            flat = np.frombuffer(C.string_at(buf, tot.value * 16), dtype=np.uint64).reshape(-1, 2).copy()
        finally:
            _lib.lib().bwb_free(buf)
        ends = np.cumsum(counts)
        return [flat[e - c:e] for c, e in zip(counts, ends)]

    def calculate_d(self, seq, offsets, use_len: int = 0) -> List[np.ndarray]:
        seq, offsets = _u8(seq), _u64(offsets)
        n = len(offsets) - 1
        total = int(offsets[-1] - offsets[0])
        out = np.zeros(2 * (total + n), dtype=np.int32)
        _lib.check(_lib.lib().bwb_calculate_d(self._ctx, seq.ctypes.data, offsets.ctypes.data, n, use_len,
                                              out.ctypes.data), self._ctx)
        res = []
        base = int(offsets[0])
        for r in range(n):
            ln = int(offsets[r + 1] - offsets[r])
            dl = min(use_len, ln) if use_len > 0 else ln
            o = 2 * (int(offsets[r]) - base + r)
            res.append(out[o:o + 2 * (dl + 1)].reshape(-1, 2).copy())
        return res

    def lower_bounds(self, seq, offsets, seed_len: int):
        """K3 of the production engine -> (list of D arrays, list of D_seed arrays or None)."""
        seq, offsets = _u8(seq), _u64(offsets)
        n = len(offsets) - 1
        total = int(offsets[-1] - offsets[0])
        dm = np.zeros(2 * (total + n), dtype=np.int32)
        ds = np.zeros(2 * n * (seed_len + 1), dtype=np.int32) if seed_len else None
        _lib.check(_lib.lib().bwb_lower_bounds(self._ctx, seq.ctypes.data, offsets.ctypes.data, n, seed_len,
                                               dm.ctypes.data, ds.ctypes.data if seed_len else None), self._ctx)
        base = int(offsets[0])
        main = []
        for r in range(n):
            ln = int(offsets[r + 1] - offsets[r])
            o = 2 * (int(offsets[r]) - base + r)
            main.append(dm[o:o + 2 * (ln + 1)].reshape(-1, 2).copy())
        seed = None
        if seed_len:
            seed = [ds[2 * r * (seed_len + 1):2 * (r + 1) * (seed_len + 1)].reshape(-1, 2).copy() for r in range(n)]
        return main, seed

    # ---- K4 + K5 --------------------------------------------------------------------------------
    def align(self, seq, offsets, params: Optional[Params] = None) -> AlignResult:
        """Host buffers in, host-readable results out (H2D + kernels + D2H)."""
        seq, offsets = _u8(seq), _u64(offsets)
        params = params or default_params()
        h = C.c_void_p()
        _lib.check(_lib.lib().bwb_align(self._ctx, C.byref(params), seq.ctypes.data, offsets.ctypes.data,
                                        len(offsets) - 1, C.byref(h)), self._ctx)
        return AlignResult(self, h.value)

    def set_seed_carry(self, read_seq=None):
        """SURVEY Q6 across calls: the read (nt4 codes) whose D_seed short reads at the start of the next align() call
        inherit (dist.seed_carry_read); None clears it.  Later calls keep the chain going by themselves."""
        if read_seq is None or len(read_seq) == 0:
            _lib.check(_lib.lib().bwb_set_seed_carry(self._ctx, None, 0), self._ctx)
        else:
            rs = _u8(read_seq)
            _lib.check(_lib.lib().bwb_set_seed_carry(self._ctx, rs.ctypes.data, len(rs)), self._ctx)

    def align_fastq(self, fastq_path: str, aln_path: Optional[str], params: Optional[Params] = None, batch: int = 0,
                    sam_path: Optional[str] = None, ann_path: Optional[str] = None, max_mm: int = 6) -> int:
        """bwb_align_fastq on this context (index already loaded): FASTQ file -> .aln / SAM file, streamed."""
        params = params or default_params()
        n = _lib.lib().bwb_align_fastq(self._ctx, C.byref(params), os.fsencode(fastq_path),
                                       os.fsencode(aln_path) if aln_path else None,
                                       os.fsencode(sam_path) if sam_path else None,
                                       os.fsencode(ann_path) if ann_path else None,
                                       self.index_length(), int(max_mm), int(batch))
        if n < 0:
            _lib.check(int(n), self._ctx)
        return int(n)

    def upload_reads(self, seq, offsets) -> DeviceReads:
        seq, offsets = _u8(seq), _u64(offsets)
        h = C.c_void_p()
        _lib.check(_lib.lib().bwb_reads_upload(self._ctx, seq.ctypes.data, offsets.ctypes.data, len(offsets) - 1,
                                               C.byref(h)), self._ctx)
        dr = DeviceReads(self, h.value, len(offsets) - 1)
        if not hasattr(self, "_children"):
            self._children = []
        self._children = [r for r in self._children if r() is not None] + [weakref.ref(dr)]
        return dr

    def align_resident(self, reads: DeviceReads, params: Optional[Params] = None, fetch: bool = False) -> AlignResult:
        params = params or default_params()
        h = C.c_void_p()
        _lib.check(_lib.lib().bwb_align_resident(self._ctx, C.byref(params), reads._h, int(fetch), C.byref(h)), self._ctx)
        return AlignResult(self, h.value)


def align_reads(fasta_path: str, fastq_path: str, aln_path: str, params: Optional[Params] = None,
                devices: Optional[Sequence[int]] = None, batch: int = 0, sam_path: Optional[str] = None,
                max_mm: int = 6) -> int:
    """`bwbble align` (align_reads, align.c:40-87): loads <fasta>.bwt, streams the FASTQ through the native
    batch reader (fastq_stream.cpp) and writes the binary .aln file -- and, if sam_path is given, the
    SAM file `bwbble aln2sam -n max_mm` would write -- without holding all reads in memory."""
    params = params or default_params()
    with Aligner(devices) as al:
        al.load_index(fasta_path + ".bwt", with_sa=sam_path is not None)
        if params.use_precalc:                      # -P: <fasta>.pre is loaded, or made first (align.c:59-64)
            pre = fasta_path + ".pre"
            if os.path.exists(pre):
                al.load_precalc(pre, bool(params.is_multiref))
            else:
                al.build_precalc(bool(params.is_multiref))
                al.write_precalc(pre)
        n = _lib.lib().bwb_align_fastq(al._ctx, C.byref(params), os.fsencode(fastq_path), os.fsencode(aln_path),
                                       os.fsencode(sam_path) if sam_path else None,
                                       os.fsencode(fasta_path + ".ann") if sam_path else None,
                                       al.index.length, int(max_mm), int(batch))
        if n < 0:
            _lib.check(int(n), al._ctx)
    return int(n)


def alns2sam(fasta_path: str, fastq_path: str, sam_path: str, params: Optional[Params] = None, max_mm: int = 6,
             devices: Optional[Sequence[int]] = None) -> int:
    """`bwbble align` + `bwbble aln2sam -n max_mm` in one go, SAM straight from the device results
    (alns2sam, align.c:494-556, without the .aln round trip)."""
    from .fastx import read_fastq
    params = params or default_params()
    reads = read_fastq(fastq_path, with_quals=True)
    with Aligner(devices) as al:
        al.load_index(fasta_path + ".bwt", with_sa=True)
        res = al.align(reads.seq, reads.offsets, params)
        res.write_sam(sam_path, fasta_path + ".ann", reads.names, reads.seq, reads.offsets, reads.meta["quals"], max_mm)
        res.close()
    return reads.n
