"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every check goes through the C ABI of
libbwbble_b200.so and compares with the CPU restatement under oracle/ on the same seeded inputs.
Integer / byte / index work: the bar is bit-exact."""
import numpy as np
import pytest

import oracle
from bwbble_b200 import Aligner, default_params, load_bwt
from bwbble_b200.aln import first_difference

pytestmark = pytest.mark.gpu


import os

# the two A/B engines of round 1 (warp per read, 8 lanes per read) are only built with -DBWB_AB_ENGINES
_ENGINES = ["lane-idx32", "lane-idx64", "lane-idx32-recycle"] + (["group-idx32", "group-idx64", "warp-idx32", "warp-idx64"]
                                           if os.environ.get("BWBBLE_TEST_AB_ENGINES") else [])


@pytest.fixture(scope="module", params=_ENGINES)
def gpu_case(small_case, request):
    """lane = production engine (one read per lane: k_calc_d_g + k_search_l); group (8 lanes per read)
    and warp (k_align) are the A/B baselines the round-1 profiles compare against;
    idx32 = 32-bit SA coordinates + 16-byte heap entries (indexes < 2^32 rows, max_gapo <= 1);
    idx64 = the wide kernels (genome-scale format) forced onto the same small index;
    recycle = K4's slot recycling (the genome-scale / long-read configuration) forced on with 32-bit coordinates."""
    al = Aligner(heap_pool_mb=512, hits_per_read=256, list_cap=1024)
    if request.param.endswith("idx64"):
        al.set_option("force_wide", 1)           # (the wide format also recycles popped heap slots)
    if request.param.endswith("recycle"):
        al.set_option("recycle", 1)              # auto: only beyond 2^28 rows / 128-base reads / wide entries
    if not request.param.startswith("lane"):
        al.set_option("engine", {"warp": 1, "group": 2}[request.param.split("-")[0]])
    al.load_index(small_case["bwt"])
    orc = oracle.Oracle(small_case["bwt"])
    yield {"al": al, "orc": orc, **small_case}
    orc.close()
    al.close()


def _expected_blocks(ix):
    """numpy restatement of the K0 layout (bwb_device.cuh): exclusive counters + 4 bit planes."""
    sym = ix.symbols().astype(np.uint32)
    nb = (ix.length + 127) // 128
    pad = np.zeros(nb * 128, dtype=np.uint32)
    pad[: ix.length] = sym
    blk = pad.reshape(nb, 4, 32)
    out = np.zeros((nb, 32), dtype=np.uint32)
    O = ix.O.reshape(-1, 16).astype(np.int64)
    first = pad.reshape(nb, 128)[:, 0]
    cnt = O[:nb].copy()
    rows = np.arange(nb)
    adj = np.ones(nb, dtype=np.int64)
    adj[(first == 0) & (rows * 128 == ix.sa0_index)] = 0
    cnt[rows, first] -= adj
    out[:, :16] = cnt.astype(np.uint32)
    w = (1 << np.arange(32, dtype=np.uint64))[None, None, :]
    for k in range(4):
        out[:, 16 + 4 * k:20 + 4 * k] = (((blk >> k) & 1).astype(np.uint64) * w).sum(axis=2).astype(np.uint32)
    return out


def test_relayout_blocks(gpu_case):
    ix = load_bwt(gpu_case["bwt"])
    got = gpu_case["al"].download_blocks()
    exp = _expected_blocks(ix)
    assert got.shape == exp.shape
    bad = np.argwhere(got != exp)
    assert len(bad) == 0, "first differing (block, word): %s got %s exp %s" % (bad[0], got[tuple(bad[0])], exp[tuple(bad[0])])


def _positions(length, rng, n):
    edge = [0, 1, 126, 127, 128, 129, 255, 256, length - 2, length - 1, (1 << 64) - 1, length - 129, length - 128]
    pos = np.concatenate([np.array(edge, dtype=np.uint64), rng.integers(0, length, size=n, dtype=np.uint64)])
    return pos


def test_occ_matches_oracle(gpu_case):
    rng = np.random.default_rng(3)
    al, orc = gpu_case["al"], gpu_case["orc"]
    pos = _positions(orc.length, rng, 4000)
    codes = rng.integers(1, 16, size=len(pos), dtype=np.uint8)
    got = al.occ(codes, pos)
    exp = np.array([orc.O(int(c), int(p)) for c, p in zip(codes, pos)], dtype=np.uint64)
    bad = np.nonzero(got != exp)[0]
    assert len(bad) == 0, "O(%d,%d): got %d exp %d" % (codes[bad[0]], pos[bad[0]], got[bad[0]], exp[bad[0]])


def test_occ_sentinel_block(gpu_case):
    """rows around sa0_index (nibble 0 there) and every code."""
    al, orc = gpu_case["al"], gpu_case["orc"]
    sa0 = load_bwt(gpu_case["bwt"]).sa0_index
    pos = np.array([p for p in range(max(0, sa0 - 130), min(orc.length, sa0 + 130))], dtype=np.uint64)
    for c in range(1, 16):
        got = al.occ(np.full(len(pos), c, dtype=np.uint8), pos)
        exp = np.array([orc.O(c, int(p)) for p in pos], dtype=np.uint64)
        assert (got == exp).all(), "code %d" % c


@pytest.mark.parametrize("inc", [0, 1])
def test_occ_alphabet_matches_oracle_including_quirk(gpu_case, inc):
    rng = np.random.default_rng(4)
    al, orc = gpu_case["al"], gpu_case["orc"]
    pos = _positions(orc.length, rng, 3000)
    got = al.occ_alphabet(pos, inc)
    for q, p in enumerate(pos):
        exp = orc.O_alphabet(int(p), inc)
        assert (got[q, 1:] == exp[1:]).all(), "O_alphabet(%d,%d): got %s exp %s" % (p, inc, got[q], exp)


def test_exact_match_interval_lists(gpu_case):
    al, orc, reads = gpu_case["al"], gpu_case["orc"], gpu_case["reads"]
    got = al.exact_match(reads.seq, reads.offsets)
    nonempty = 0
    for r in range(reads.n):
        exp = orc.exact_match(reads.read(r))
        assert got[r].shape == exp.shape and (got[r] == exp).all(), "read %d: got %s exp %s" % (r, got[r], exp)
        nonempty += len(exp) > 0
    assert nonempty > 10


@pytest.fixture(scope="module")
def dense_case(tmp_path_factory):
    """1.5 Mbp genome with 4 % SNP sites: interval lists of several hundred entries (shared-memory
    spill + ordered merge across passes)."""
    from bwbble_b200 import synth, index
    d = tmp_path_factory.mktemp("dense")
    g = synth.make_genome(17, 1500000, snp_rate=0.04, tri_frac=0.1, n_bubbles=100)
    fa = str(d / "g.fa")
    g.write_fasta(fa)
    index.build_index(fa)
    al = Aligner(heap_pool_mb=512, hits_per_read=256, list_cap=2048)
    al.load_index(fa + ".bwt")
    orc = oracle.Oracle(fa + ".bwt")
    yield {"al": al, "orc": orc, "genome": g}
    orc.close()
    al.close()


def test_exact_match_short_prefixes_long_lists(dense_case):
    """short reads keep the search in the wide top of the tree: long lists, merges, smem spill."""
    from bwbble_b200 import synth
    al, orc = dense_case["al"], dense_case["orc"]
    reads = synth.make_reads(dense_case["genome"], 5, 300, 12, 0)
    rng = np.random.default_rng(5)
    lens = rng.integers(1, 13, size=300)
    seqs = [reads.read(i)[:l] for i, l in enumerate(lens)]
    seq = np.concatenate(seqs)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    got = al.exact_match(seq, off)
    longest = 0
    for r, s in enumerate(seqs):
        exp = orc.exact_match(s)
        assert got[r].shape == exp.shape and (got[r] == exp).all(), "read %d len %d" % (r, len(s))
        longest = max(longest, len(exp))
    assert longest > 64, "fixture does not reach the shared-memory spill (longest list %d)" % longest


def test_dense_genome_align(dense_case):
    """long interval lists inside calculate_d and the exact tails of the inexact search"""
    from bwbble_b200 import synth
    al, orc = dense_case["al"], dense_case["orc"]
    reads = synth.make_reads(dense_case["genome"], 6, 400, 100, 2, indel_frac=0.1)
    p = default_params(n=3)
    res = al.align(reads.seq, reads.offsets, p)
    got = res.aln_bytes()
    exp, st = orc.align(reads.seq, reads.offsets, p)
    assert got == exp, first_difference(got, exp)
    assert st["max_list"] > 32 and res.counters()["max_list"] == st["max_list"]


@pytest.mark.parametrize("use_len", [0, 32, 20])
def test_calculate_d(gpu_case, use_len):
    al, orc, reads = gpu_case["al"], gpu_case["orc"], gpu_case["reads"]
    got = al.calculate_d(reads.seq, reads.offsets, use_len)
    for r in range(reads.n):
        exp = orc.calculate_d(reads.read(r), use_len)
        assert got[r].shape == exp.shape and (got[r] == exp).all(), "read %d:\n got %s\n exp %s" % (r, got[r].T, exp.T)


@pytest.mark.parametrize("seed_len", [0, 32, 20])
def test_lower_bounds_k3(gpu_case, seed_len):
    al, orc, reads = gpu_case["al"], gpu_case["orc"], gpu_case["reads"]
    main, seed = al.lower_bounds(reads.seq, reads.offsets, seed_len)
    al.set_option("kmer_table", 2)                     # same arrays without the 10-mer table
    main2, seed2 = al.lower_bounds(reads.seq, reads.offsets, seed_len)
    al.set_option("kmer_table", 1)
    assert all((x == y).all() for x, y in zip(main, main2))
    assert seed is None or all((x == y).all() for x, y in zip(seed, seed2))
    for r in range(reads.n):
        rd = reads.read(r)
        exp = orc.calculate_d(rd)
        assert main[r].shape == exp.shape and (main[r] == exp).all(), "D read %d:\n got %s\n exp %s" % (r, main[r].T, exp.T)
        if seed_len:
            exps = orc.calculate_d(rd, seed_len) if len(rd) > seed_len else np.zeros((seed_len + 1, 2), dtype=np.int32)
            assert (seed[r] == exps).all(), "D_seed read %d:\n got %s\n exp %s" % (r, seed[r].T, exps.T)


GRID = [
    dict(n=0), dict(n=1), dict(n=2), dict(n=3), dict(n=5),
    dict(n=4, o=2, e=3, k=3, l=20), dict(n=3, M=2, O=5, E=2), dict(n=4, l=0), dict(n=3, k=1),
    dict(n=6, o=2, M=4, O=4, E=4), dict(n=3, o=0), dict(n=2, e=0), dict(n=4, m=200),
    dict(n=4, M=0, m=3000), dict(n=4, E=0, o=2), dict(n=3, M=11), dict(n=3, O=3, E=3),     # score classes sharing a bucket
    dict(n=3, M=10, O=30, E=10), dict(n=4, o=2, e=4, M=7, O=23, E=9),    # 170 / 149 buckets in the reference's heap (> 128), 6 / 20 reachable
]


@pytest.mark.parametrize("kw", GRID, ids=lambda k: "-".join("%s%d" % kv for kv in k.items()))
def test_align_aln_bytes_equal_oracle(gpu_case, kw):
    al, orc, reads = gpu_case["al"], gpu_case["orc"], gpu_case["reads"]
    p = default_params(**kw)
    res = al.align(reads.seq, reads.offsets, p)
    got = res.aln_bytes()
    exp, st = orc.align(reads.seq, reads.offsets, p)
    if got != exp:
        d = first_difference(got, exp)
        raise AssertionError("params %s: first differing read %s\n got %s\n exp %s" % (kw, d[0], d[1], d[2]))
    ctr = res.counters()
    assert ctr["pops"] == st["pops"] and ctr["pushes"] == st["pushes"], (ctr, st)
    assert ctr["exact_tails"] == st["exact_tail_calls"]
    assert ctr["max_heap"] == st["max_heap"]
    res.close()


def test_tiny_private_ranges_use_the_shared_chunk_pool(small_case):
    """heap_pool_mb=1 shrinks every private chunk range to the minimum, so the bucket heaps live
    on the shared lock-free pool (borrow at push, hand back at flush)."""
    al = Aligner(heap_pool_mb=1, hits_per_read=256, list_cap=1024)
    al.load_index(small_case["bwt"])
    orc = oracle.Oracle(small_case["bwt"])
    reads = small_case["reads"]
    p = default_params(n=5)
    for _ in range(2):          # second call re-uses the pool after the per-call reset
        got = al.align(reads.seq, reads.offsets, p).aln_bytes()
        exp, st = orc.align(reads.seq, reads.offsets, p)
        assert got == exp, first_difference(got, exp)
    assert st["max_heap"] > 1024
    orc.close()
    al.close()


def test_align_ragged_lengths_and_n_reads(gpu_case):
    from bwbble_b200 import synth
    al, orc = gpu_case["al"], gpu_case["orc"]
    reads = synth.make_reads(gpu_case["genome"], 11, 300, 150, 4, indel_frac=0.3, n_base_frac=0.01, ragged=(36, 150))
    p = default_params(n=4)
    got = al.align(reads.seq, reads.offsets, p).aln_bytes()
    exp, _ = orc.align(reads.seq, reads.offsets, p)
    assert got == exp, first_difference(got, exp)


def test_align_empty_batch_and_all_n_read(gpu_case):
    al, orc = gpu_case["al"], gpu_case["orc"]
    p = default_params(n=2)
    res = al.align(np.zeros(0, dtype=np.uint8), np.zeros(1, dtype=np.uint64), p)
    assert res.num_reads == 0 and res.aln_bytes() == b""
    seq = np.concatenate([np.full(50, 4, dtype=np.uint8), gpu_case["reads"].read(0)])
    off = np.array([0, 50, 50 + len(gpu_case["reads"].read(0))], dtype=np.uint64)
    got = al.align(seq, off, p).aln_bytes()
    exp, _ = orc.align(seq, off, p)
    assert got == exp


def test_hit_buffer_regrow_after_overflow(small_case):
    """with more hits than the output buffers hold, K5 must not touch memory; the host regrows and
    re-runs the shard.  hit_cap0=0 forces that path on the first attempt."""
    reads = small_case["reads"]
    p = default_params(n=3)
    orc = oracle.Oracle(small_case["bwt"])
    exp, _ = orc.align(reads.seq, reads.offsets, p)
    orc.close()
    for wide in (0, 1):
        with Aligner(heap_pool_mb=256) as al:
            if wide:
                al.set_option("force_wide", 1)
            al.set_option("hit_cap0", 0)
            al.load_index(small_case["bwt"], with_sa=True)          # K6 runs behind K5 as well
            got = al.align(reads.seq, reads.offsets, p).aln_bytes()
            assert got == exp, first_difference(got, exp)


def test_stale_resident_results_fail_loudly(gpu_case):
    """un-fetched results point into per-context buffers; a later launch invalidates them."""
    from bwbble_b200 import BwbError
    al, reads = gpu_case["al"], gpu_case["reads"]
    p = default_params(n=2)
    dr = al.upload_reads(reads.seq, reads.offsets)
    first = al.align_resident(dr, p, fetch=False)
    second = al.align_resident(dr, p, fetch=False)
    with pytest.raises(BwbError):
        first.fetch()
    assert second.fetch().aln_bytes() == al.align(reads.seq, reads.offsets, p).aln_bytes()


def test_resident_path_equals_host_path(gpu_case):
    al, reads = gpu_case["al"], gpu_case["reads"]
    p = default_params(n=3)
    a = al.align(reads.seq, reads.offsets, p).aln_bytes()
    dr = al.upload_reads(reads.seq, reads.offsets)
    res = al.align_resident(dr, p, fetch=False)
    b = res.fetch().aln_bytes()
    assert a == b


def test_unsupported_params_fail_loudly(gpu_case):
    from bwbble_b200 import BwbError
    al, reads = gpu_case["al"], gpu_case["reads"]
    for kw in (dict(n=2, use_precalc=1), dict(n=2, o=9), dict(n=140, M=9)):
        with pytest.raises(BwbError):
            al.align(reads.seq, reads.offsets, default_params(**kw))


def test_gpu_reproduces_the_reference_golden_files(tmp_path):
    """K0..K5 against the files the UNMODIFIED reference wrote (tests/golden/), no oracle involved."""
    import golden_util as G
    from bwbble_b200.fastx import read_fastq
    import os
    fa = G.materialise_index(tmp_path)
    reads = read_fastq(os.path.join(G.GOLDEN, "r.fq"))
    with Aligner(heap_pool_mb=512) as al:
        al.load_index(fa + ".bwt")
        for tag, flags in sorted(G.grid().items()):
            kw = G.flags_to_kwargs(flags)
            kw.pop("n_threads", None)
            got = al.align(reads.seq, reads.offsets, default_params(**kw)).aln_bytes()
            exp = G.golden_bytes("aln_%s.aln" % tag)
            assert got == exp, "%s: %s" % (tag, first_difference(got, exp))


@pytest.mark.parametrize("kw", [dict(n=0), dict(n=2), dict(n=4), dict(n=3, o=2, e=3, l=20, k=3), dict(n=4, M=0, m=3000)],
                         ids=lambda k: "-".join("%s%d" % kv for kv in k.items()))
def test_single_genome_mode_S(small_case, kw):
    """-S (is_multiref = 0): 4-code fan-out A,G,C,T (O_actg_alphabet bwt.c:440-463, exact_match_1to1_bounded
    exact_match.c:196-222, the !is_multiref branches of inexact_match.c) -- SURVEY 8f row 3."""
    reads = small_case["reads"]
    p = default_params(is_multiref=0, **kw)
    orc = oracle.Oracle(small_case["bwt"])
    exp, st = orc.align(reads.seq, reads.offsets, p)
    orc.close()
    for wide in (0, 1):
        with Aligner(heap_pool_mb=512) as al:
            if wide:
                al.set_option("force_wide", 1)
            al.load_index(small_case["bwt"])
            res = al.align(reads.seq, reads.offsets, p)
            got = res.aln_bytes()
            assert got == exp, first_difference(got, exp)
            ctr = res.counters()
            assert ctr["pops"] == st["pops"] and ctr["pushes"] == st["pushes"] and ctr["exact_tails"] == st["exact_tail_calls"]


@pytest.mark.parametrize("tag", ["n0", "n3", "n4_o2_e3_k3_l20"])
def test_device_locate_and_sam_equal_reference_aln2sam(tmp_path, tag):
    """K6 (SA locate + top1/top2) + the host SAM writer reproduce, byte for byte, the SAM file
    the reference's `aln2sam -n <n>` wrote for the same reads (tests/golden/sam_*.sam)."""
    import golden_util as G
    from bwbble_b200.fastx import read_fastq
    import os
    fa = G.materialise_index(tmp_path)
    reads = read_fastq(os.path.join(G.GOLDEN, "r.fq"), with_quals=True)
    kw = G.flags_to_kwargs(G.grid()[tag])
    p = default_params(**kw)
    for wide in (0, 1):
        with Aligner(heap_pool_mb=512) as al:
            if wide:
                al.set_option("force_wide", 1)
            al.load_index(fa + ".bwt", with_sa=True)
            res = al.align(reads.seq, reads.offsets, p)
            sam = str(tmp_path / ("out%d.sam" % wide))
            res.write_sam(sam, fa + ".ann", reads.names, reads.seq, reads.offsets, reads.meta["quals"], max_mm=kw["max_diff"])
            got, exp = open(sam, "rb").read(), G.golden_bytes("sam_%s.sam" % tag)
            if got != exp:
                gl, el = got.split(b"\n"), exp.split(b"\n")
                bad = [i for i in range(min(len(gl), len(el))) if gl[i] != el[i]][:3]
                raise AssertionError("SAM differs at lines %s:\n got %s\n exp %s" % (bad, [gl[i] for i in bad], [el[i] for i in bad]))
            loc = res.locations()
            assert ((loc["ref_pos"] == np.uint64(2**64 - 1)) == (res.counts() == 0)).all()


@pytest.mark.parametrize("batch", [0, 37])
def test_streaming_fastq_ingest_writes_reference_aln_and_sam(tmp_path, batch):
    """native batched FASTQ reader -> align -> append; any batch size gives the reference's files."""
    import golden_util as G
    from bwbble_b200 import align_reads
    import os
    fa = G.materialise_index(tmp_path)
    aln, sam = str(tmp_path / "o.aln"), str(tmp_path / "o.sam")
    n = align_reads(fa, os.path.join(G.GOLDEN, "r.fq"), aln, default_params(n=3), batch=batch, sam_path=sam, max_mm=3)
    assert n == 200
    assert open(aln, "rb").read() == G.golden_bytes("aln_n3.aln")
    assert open(sam, "rb").read() == G.golden_bytes("sam_n3.sam")


@pytest.mark.parametrize("tag", ["sim_n0", "sim_n5"])
def test_reference_shipped_fastq_through_the_streaming_entry_point(tmp_path, tag):
    """BASELINE configs[0]: the reference's own test_data/sim_chr21_N100.fastq through bwb_align_fastq (.aln + SAM)
    against what the unmodified reference wrote for it (index g21.fa: the reads' loci planted, IUPAC SNP sites)"""
    import golden_util as G
    from bwbble_b200 import align_reads
    import os
    fa = G.materialise_index(tmp_path, "g21.fa")
    kw = G.flags_to_kwargs(G.MANIFEST["shipped"][tag])
    aln, sam = str(tmp_path / "o.aln"), str(tmp_path / "o.sam")
    n = align_reads(fa, os.path.join(G.GOLDEN, "sim_chr21_N100.fastq"), aln, default_params(**kw), sam_path=sam,
                    max_mm=kw["max_diff"])
    assert n == 100
    assert open(aln, "rb").read() == G.golden_bytes("aln_%s.aln" % tag)
    assert open(sam, "rb").read() == G.golden_bytes("sam_%s.sam" % tag)


def test_kmer_table_on_dense_genome(dense_case):
    """10-mer table of calculate_d's top of tree: adoption of tabulated lists (dozens of intervals), restarts inside
    the window, N inside the window, reads shorter than the window -- D arrays equal the oracle's."""
    from bwbble_b200 import synth
    al, orc = dense_case["al"], dense_case["orc"]
    reads = synth.make_reads(dense_case["genome"], 9, 300, 60, 3, n_base_frac=0.02, ragged=(6, 60))
    main, seed = al.lower_bounds(reads.seq, reads.offsets, 20)
    for r in range(reads.n):
        rd = reads.read(r)
        exp = orc.calculate_d(rd)
        assert (main[r] == exp).all(), "D read %d (len %d):\n got %s\n exp %s" % (r, len(rd), main[r].T, exp.T)
        exps = orc.calculate_d(rd, 20) if len(rd) > 20 else np.zeros((21, 2), dtype=np.int32)
        assert (seed[r] == exps).all()


# ---- -P: 12-mer seed table ---------------------------------------------------------
@pytest.fixture(scope="module")
def precalc_case(tmp_path_factory):
    """golden index + the seed tables K0c builds for it, written in the reference's .pre layout"""
    import golden_util as G
    from bwbble_b200.fastx import read_fastq
    import os
    d = tmp_path_factory.mktemp("pre")
    fa = G.materialise_index(d)
    reads = read_fastq(os.path.join(G.GOLDEN, "r.fq"))
    pre = {}
    with Aligner(heap_pool_mb=512) as al:
        al.load_index(fa + ".bwt")
        for mode, multi in (("multi", True), ("single", False)):
            al.build_precalc(multi)
            pre[mode] = str(d / ("g.%s.pre" % mode))
            al.write_precalc(pre[mode])
            assert al.precalc_num_intervals() == G.MANIFEST["pre"][mode]["intervals"]
    return {"fa": fa, "reads": reads, "pre": pre}


@pytest.mark.parametrize("mode", ["multi", "single"])
def test_precalc_table_is_the_reference_pre_file(precalc_case, mode):
    """K0c + the .pre writer against the file precalc_sa_intervals() wrote (md5 pinned in the manifest: 68 MB)"""
    import golden_util as G
    import hashlib
    raw = open(precalc_case["pre"][mode], "rb").read()
    assert len(raw) == G.MANIFEST["pre"][mode]["bytes"]
    assert hashlib.md5(raw).hexdigest() == G.MANIFEST["pre"][mode]["md5"]


@pytest.mark.parametrize("wide", [0, 1])
def test_precalc_rows_equal_oracle(precalc_case, wide):
    """rows of the device table against the restated exact_match(), incl. the rows the reads use"""
    reads = precalc_case["reads"]
    orc = oracle.Oracle(precalc_case["fa"] + ".bwt")
    rng = np.random.default_rng(5)
    rows = set(int(x) for x in rng.integers(0, 1 << 24, 300))
    for r in range(reads.n):
        s = reads.seq[int(reads.offsets[r]):int(reads.offsets[r + 1])]
        if len(s) >= 12 and (s[:12] < 4).all():
            rows.add(int(sum((3 - int(s[j])) << (2 * j) for j in range(12))))
    try:
        for multi in (True, False):
            p = default_params(is_multiref=int(multi))
            with Aligner(heap_pool_mb=256) as al:
                if wide:
                    al.set_option("force_wide", 1)
                al.load_index(precalc_case["fa"] + ".bwt")
                al.build_precalc(multi)
                nonempty = 0
                for x in sorted(rows):
                    got, exp = al.precalc_row(x), orc.precalc_entry(x, p)
                    assert got.shape == exp.shape and (got == exp).all(), (multi, x, got, exp)
                    nonempty += len(exp) > 0
                assert nonempty > 50
    finally:
        orc.close()


def test_precalc_golden_files(precalc_case):
    """`bwbble align -P ...` of the unmodified reference (tests/golden/aln_P_*.aln, aln_SP_n3.aln): table loaded
    from the .pre file, 32- and 64-bit kernels"""
    import golden_util as G
    reads = precalc_case["reads"]
    for wide in (0, 1):
        for mode, multi in (("multi", 1), ("single", 0)):
            with Aligner(heap_pool_mb=512) as al:
                if wide:
                    al.set_option("force_wide", 1)
                al.load_index(precalc_case["fa"] + ".bwt")
                al.load_precalc(precalc_case["pre"][mode], bool(multi))
                for tag, flags in sorted(G.pgrid().items()):
                    kw = G.flags_to_kwargs(flags)
                    if kw.get("is_multiref", 1) != multi:
                        continue
                    got = al.align(reads.seq, reads.offsets, default_params(**kw)).aln_bytes()
                    exp = G.golden_bytes("aln_%s.aln" % tag)
                    assert got == exp, "%s wide=%d: %s" % (tag, wide, first_difference(got, exp))


@pytest.mark.parametrize("kw", [dict(n=2), dict(n=5), dict(n=4, o=2, e=4, l=24, k=3), dict(n=3, is_multiref=0)],
                         ids=lambda k: "-".join("%s%d" % kv for kv in k.items()))
def test_precalc_align_equals_oracle_with_counters(small_case, kw):
    """-P on the seeded synthetic case: .aln bytes and the pop/push counts of the restated search"""
    reads = small_case["reads"]
    p = default_params(use_precalc=1, **kw)
    orc = oracle.Oracle(small_case["bwt"])
    exp, st = orc.align(reads.seq, reads.offsets, p)
    orc.close()
    with Aligner(heap_pool_mb=512) as al:
        al.load_index(small_case["bwt"])
        al.build_precalc(bool(p.is_multiref))
        res = al.align(reads.seq, reads.offsets, p)
        got = res.aln_bytes()
        assert got == exp, first_difference(got, exp)
        ctr = res.counters()
        assert ctr["pops"] == st["pops"] and ctr["pushes"] == st["pushes"]
        dr = al.upload_reads(reads.seq, reads.offsets)
        assert al.align_resident(dr, p, fetch=True).aln_bytes() == exp


def test_precalc_misuse_fails_loudly(small_case):
    from bwbble_b200 import BwbError
    reads = small_case["reads"]
    with Aligner(heap_pool_mb=256) as al:
        al.load_index(small_case["bwt"])
        with pytest.raises(BwbError):                    # no table yet
            al.align(reads.seq, reads.offsets, default_params(n=2, use_precalc=1))
        al.build_precalc(True)
        with pytest.raises(BwbError):                    # table of the other mode
            al.align(reads.seq, reads.offsets, default_params(n=2, use_precalc=1, is_multiref=0))
    with Aligner(heap_pool_mb=256) as al:
        al.load_index(small_case["bwt"])
        al.build_precalc(True)
        al.load_index(small_case["bwt"])                 # a new index drops the table computed on the old one
        with pytest.raises(BwbError):
            al.align(reads.seq, reads.offsets, default_params(n=2, use_precalc=1))
    with Aligner(heap_pool_mb=256) as al:                # the round-1 A/B engines are not in the default build
        with pytest.raises(BwbError):
            al.set_option("engine", 1)


def test_heavy_first_queue_order_is_result_neutral(small_case):
    """K3b only changes the order in which K4 takes the reads: bytes and the pop/push totals stay the oracle's"""
    reads = small_case["reads"]
    p = default_params(n=4)
    orc = oracle.Oracle(small_case["bwt"])
    exp, st = orc.align(reads.seq, reads.offsets, p)
    orc.close()
    for opt in (1, 2):                                   # 1 = on (default), 2 = input order
        with Aligner(heap_pool_mb=512) as al:
            al.set_option("heavy_first", opt)
            al.load_index(small_case["bwt"])
            res = al.align(reads.seq, reads.offsets, p)
            got = res.aln_bytes()
            assert got == exp, (opt, first_difference(got, exp))
            ctr = res.counters()
            assert ctr["pops"] == st["pops"] and ctr["pushes"] == st["pushes"]


# ---- K7: index construction on the device (SURVEY 8f #4b) --------------------------------------------
def test_device_index_build_equals_the_reference_files(tmp_path):
    """`bwbble index g.fa` of the unmodified reference (tests/golden/g.fa.bwt.gz, .ann) rebuilt with the suffix
    sort on the GPU: byte-identical .bwt (header, C, packed BWT, checkpoints O, SA samples) and .ann"""
    import golden_util as G
    import hashlib
    from bwbble_b200 import index
    fa = str(tmp_path / "g.fa")
    open(fa, "wb").write(G.golden_bytes("g.fa"))
    with Aligner(heap_pool_mb=64) as al:
        index.build_index(fa, aligner=al)
        assert al.last_index_sort_rounds >= 2
    assert hashlib.md5(open(fa + ".bwt", "rb").read()).hexdigest() == G.MANIFEST["md5"]["g.fa.bwt"]
    assert open(fa + ".ann", "rb").read() == G.golden_bytes("g.fa.ann")


def test_device_index_build_equals_host_builder_on_deep_repeats(tmp_path):
    """long N runs, exact repeats and microsatellites: LCPs of tens of thousands, many doubling rounds"""
    from bwbble_b200 import index, synth
    g = synth.make_genome(77, 400_000, n_records=3, snp_rate=0.012, tri_frac=0.05, n_bubbles=60, n_frac=0.3,
                          n_repeat_copies=6, repeat_len=3000, n_microsats=4, lowercase_frac=0.01)
    a, b = str(tmp_path / "a.fa"), str(tmp_path / "b.fa")
    g.write_fasta(a)
    g.write_fasta(b)
    index.build_index(a)
    with Aligner(heap_pool_mb=64) as al:
        index.build_index(b, aligner=al)
        assert al.last_index_sort_rounds >= 10
        # and the device-built index maps reads like the host-built one
        reads = synth.make_reads(g, 78, 300, 100, 2, indel_frac=0.2)
        al.load_index(b + ".bwt")
        got = al.align(reads.seq, reads.offsets, default_params(n=3)).aln_bytes()
    assert open(a + ".bwt", "rb").read() == open(b + ".bwt", "rb").read()
    assert open(a + ".ann", "rb").read() == open(b + ".ann", "rb").read()
    orc = oracle.Oracle(a + ".bwt")
    exp, _ = orc.align(reads.seq, reads.offsets, default_params(n=3))
    orc.close()
    assert got == exp, first_difference(got, exp)


def test_arena_overflow_defers_reads_to_retry_passes(small_case):
    """thousands of reads in flight on the smallest arena: most of them run out of slots in the first pass, are
    deferred, and finish in the retry passes (fewer reads in flight, 8x / 64x the arena each) -- same bytes"""
    reads = small_case["reads"]
    rep = max(1, 80000 // reads.n)
    seq = np.tile(reads.seq, rep)
    total = int(reads.offsets[-1])
    offsets = np.concatenate([reads.offsets[:-1].astype(np.uint64) + np.uint64(k * total) for k in range(rep)]
                             + [np.array([rep * total], dtype=np.uint64)])
    p = default_params(n=5)
    orc = oracle.Oracle(small_case["bwt"])
    exp1, st = orc.align(reads.seq, reads.offsets, p)
    orc.close()
    with Aligner(heap_pool_mb=1, hits_per_read=64, throttle_pct=100) as al:      # clamps to the minimum: 512 slots per lane
        al.load_index(small_case["bwt"])                                         # (admission control off: let it overflow)
        res = al.align(seq, offsets, p)
        got = res.aln_bytes()
        assert got == exp1 * rep, first_difference(got, exp1 * rep)
        # deferred reads are searched twice: the pop total exceeds the oracle's iff something was deferred
        ctr = res.counters()
        assert ctr["pops"] > st["pops"] * rep and ctr["deferred_pass1"] > 0, "the arena never overflowed: the test does not bite"
        deferred_without_throttle = ctr["deferred_pass1"]
    # admission control (on by default together with slot recycling, i.e. for big indexes / long reads; forced here):
    # lanes wait for arena instead of starting reads that would overflow
    with Aligner(heap_pool_mb=1, hits_per_read=64, throttle_pct=60, recycle=1) as al:
        al.load_index(small_case["bwt"])
        res = al.align(seq, offsets, p)
        assert res.aln_bytes() == exp1 * rep
        assert res.counters()["deferred_pass1"] < deferred_without_throttle


def test_150bp_gapped_reads_config5_shape(small_case):
    """BASELINE configs[4] in miniature: 150 bp reads, substitutions + 1-3 bp indels, `-n 4 -o 1 -e 6`; 32- and 64-bit kernels"""
    from bwbble_b200 import synth
    reads = synth.make_reads(small_case["genome"], 91, 400, 150, 3, indel_frac=0.5, max_indel=3, n_base_frac=0.002)
    p = default_params(n=4, o=1, e=6)
    orc = oracle.Oracle(small_case["bwt"])
    exp, st = orc.align(reads.seq, reads.offsets, p)
    orc.close()
    for wide in (0, 1):
        with Aligner(heap_pool_mb=512) as al:
            if wide:
                al.set_option("force_wide", 1)
            al.load_index(small_case["bwt"])
            res = al.align(reads.seq, reads.offsets, p)
            got = res.aln_bytes()
            assert got == exp, first_difference(got, exp)
            ctr = res.counters()
            assert ctr["pops"] == st["pops"] and ctr["pushes"] == st["pushes"]


# ---- K7w: the chunked 64-bit suffix sorter (genome-scale builder) forced onto small genomes ---------------
@pytest.mark.parametrize("chunk", [2500, 6000, 1 << 20])
def test_wide_device_index_build_equals_the_reference_files(tmp_path, chunk):
    """the builder that indexes the 6.9 G-row genome (64-bit ranks, bounded sort chunks), run with chunks far
    smaller than the text: same bytes as the unmodified reference's `bwbble index g.fa`"""
    import golden_util as G
    import hashlib
    from bwbble_b200 import index
    fa = str(tmp_path / "g.fa")
    open(fa, "wb").write(G.golden_bytes("g.fa"))
    with Aligner(heap_pool_mb=64) as al:
        al.set_option("index_wide", 1)
        al.set_option("index_chunk", chunk)
        index.build_index(fa, aligner=al)
        assert al.last_index_sort_rounds >= 2
    assert hashlib.md5(open(fa + ".bwt", "rb").read()).hexdigest() == G.MANIFEST["md5"]["g.fa.bwt"]
    assert open(fa + ".ann", "rb").read() == G.golden_bytes("g.fa.ann")


def test_wide_device_index_build_on_deep_repeats(tmp_path):
    """N runs of 10^5 rows share their first symbols: groups far larger than most chunks would allow, 15+ rounds"""
    from bwbble_b200 import BwbError, index, synth
    g = synth.make_genome(77, 400_000, n_records=3, snp_rate=0.012, tri_frac=0.05, n_bubbles=60, n_frac=0.3,
                          n_repeat_copies=6, repeat_len=3000, n_microsats=4, lowercase_frac=0.01)
    a, b = str(tmp_path / "a.fa"), str(tmp_path / "b.fa")
    g.write_fasta(a)
    g.write_fasta(b)
    index.build_index(a)
    with Aligner(heap_pool_mb=64) as al:
        al.set_option("index_wide", 1)
        al.set_option("index_chunk", 4096)
        with pytest.raises(BwbError):                    # the N-run group does not fit a 4096-row chunk: loud, not wrong
            index.build_index(b, aligner=al)
        al.set_option("index_chunk", 300_000)
        index.build_index(b, aligner=al)
        assert al.last_index_sort_rounds >= 10
    assert open(a + ".bwt", "rb").read() == open(b + ".bwt", "rb").read()
    assert open(a + ".ann", "rb").read() == open(b + ".ann", "rb").read()
