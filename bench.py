#!/usr/bin/env python
"""bench.py -- reads/s of the BWBBLE hot path (100 bp reads, BWA-default diffs) on 1..N B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload chr21|small]

A "step" is one pass of the hot path (calculate_d + inexact_match for every read, K4+K5) over one
batch of synthetic reads.  Workload `chr21` = BASELINE.json configs[1]: a synthetic chr21-scale
multi-genome (48.1 Mbp, 27 % N runs, 1.2 % SNP sites of which 3 % tri-allelic, 40 k indel bubbles;
the real chr21 files are absent from the reference tree) and the 10 M-read set of 100 bp reads with
0-2 substitutions, consumed one batch per step; parameters `-n 5` + the reference defaults.

  value         whole-job reads/s with the reads already resident in HBM (kernels only)
  e2e           same metric through the public C-ABI call bwb_align() with HOST buffers: pinned
                H2D of the batch and D2H of every hit record inside the timed region
  roofline      K4's algorithmic bytes (rank queries the REFERENCE algorithm issues on these reads,
                counted by the instrumented oracle, x 128 B) / K4's CUDA-event duration, vs the
                measured HBM copy peak (MEASURED_PEAKS.json)
  cpu_baseline  the reference's own OpenMP CPU path (oracle/_ref/bwbble when present, else the
                oracle port) on this box's host cores, on a bounded sample of the same batch

Multi-GPU: one process per GPU (torchrun), index replicated, reads sharded contiguously, no
collective on the data path ("weak" scaling: every rank gets a full batch).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CACHE = os.environ.get("BWBBLE_B200_CACHE", "/tmp/bwbble_b200_cache")

_CHR21 = dict(seed=21, n_bases=48_100_000, n_records=1, snp_rate=0.012, tri_frac=0.03, n_bubbles=40_000, n_frac=0.27)
_GENOME = dict(seed=37, n_bases=3_100_000_000, n_records=24, snp_rate=0.012, tri_frac=0.03, n_bubbles=1_400_000, n_frac=0.05)
_G300 = dict(seed=37, n_bases=300_000_000, n_records=8, snp_rate=0.012, tri_frac=0.03, n_bubbles=130_000, n_frac=0.05)
WORKLOADS = {
    # name: index (cache key + genome kwargs), reads kwargs, alignment parameters, reads per step (per GPU), CPU sample
    # BASELINE configs[1] (the default; configs[0] is the 100-read CPU case of the same index)
    "chr21": dict(index="chr21", genome=_CHR21, reads=dict(read_len=100, max_sub=2), params=dict(n=5), batch=1 << 23,
                  cpu_sample=65536, params_str="-n 5 -k 2 -l 32 -o 1 -e 6 -M 3 -O 11 -E 4",
                  desc="synthetic chr21-scale multi-genome (48.1 Mbp, 27% N, 1.2% SNP, 40k bubbles); 10M x 100bp reads, 0-2 subs"),
    # BASELINE configs[2]: exact-match-only backward search (BWBBLE's own default -n 0), 0-edit reads
    "chr21-exact": dict(index="chr21", genome=_CHR21, reads=dict(read_len=100, max_sub=0), params=dict(n=0), batch=1 << 23,
                        cpu_sample=131072, params_str="-n 0 (exact-match only)",
                        desc="synthetic chr21-scale multi-genome (48.1 Mbp, 27% N, 1.2% SNP, 40k bubbles); 100bp reads, 0 edits, exact search"),
    # BASELINE configs[3]: GRCh37-scale multi-genome, >= 2^32 BWT rows (64-bit kernels, HBM-resident index)
    "genome": dict(index="genome", genome=_GENOME, reads=dict(read_len=100, max_sub=2), params=dict(n=5), batch=1 << 20,
                   cpu_sample=8192, params_str="-n 5 -k 2 -l 32 -o 1 -e 6 -M 3 -O 11 -E 4",
                   desc="synthetic 3.1 Gbp GRCh37-scale multi-genome (6.9 G BWT rows, 24 records, 5% N, 1.2% SNP, 1.4M bubbles); 100bp reads, 0-2 subs"),
    # BASELINE configs[4]: 150 bp reads, up to 4 differences with gaps, same genome-scale index
    "genome-150": dict(index="genome", genome=_GENOME, reads=dict(read_len=150, max_sub=3, indel_frac=0.4), params=dict(n=4, o=1, e=6),
                       batch=1 << 19, cpu_sample=4096, params_str="-n 4 -o 1 -e 6 (150 bp, gaps)",
                       desc="synthetic 3.1 Gbp GRCh37-scale multi-genome (6.9 G BWT rows); 150bp reads, 0-3 subs + one 1-3 bp indel in 40% of the reads"),
    # 600 M-row index (HBM-resident, ~5x L2): the regime of configs[3] at 1/11 of its size
    "g300": dict(index="g300", genome=_G300, reads=dict(read_len=100, max_sub=2), params=dict(n=5), batch=1 << 21,
                 cpu_sample=16384, params_str="-n 5 -k 2 -l 32 -o 1 -e 6 -M 3 -O 11 -E 4",
                 desc="synthetic 300 Mbp multi-genome (600 M BWT rows, 5% N, 1.2% SNP, 130k bubbles); 100bp reads, 0-2 subs"),
    "g300-150": dict(index="g300", genome=_G300, reads=dict(read_len=150, max_sub=3, indel_frac=0.4), params=dict(n=4, o=1, e=6),
                     batch=1 << 20, cpu_sample=8192, params_str="-n 4 -o 1 -e 6 (150 bp, gaps)",
                     desc="synthetic 300 Mbp multi-genome (600 M BWT rows); 150bp reads, 0-3 subs + one 1-3 bp indel in 40% of the reads"),
    "small": dict(index="small", genome=dict(seed=5, n_bases=2_000_000, n_records=2, snp_rate=0.012, tri_frac=0.03,
                                             n_bubbles=1000, n_frac=0.05),
                  reads=dict(read_len=100, max_sub=2), params=dict(n=5), batch=1 << 15, cpu_sample=2048,
                  params_str="-n 5", desc="2 Mbp synthetic multi-genome (CI-size)"),
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def prepare_index(workload: str, rank: int, barrier, aligner=None) -> str:
    """Generate the genome and build <fasta>.bwt once per box (cached under CACHE): with the device builder
    (K7) when an Aligner is given, else with the host builder -- the files are byte-identical."""
    from bwbble_b200 import index, synth
    d = os.path.join(CACHE, WORKLOADS[workload]["index"])
    fa = os.path.join(d, "g.fa")
    done = os.path.join(d, "DONE")
    if rank == 0 and not os.path.exists(done):
        os.makedirs(d, exist_ok=True)
        t = time.time()
        kw = dict(WORKLOADS[workload]["genome"])
        g = synth.make_genome(kw.pop("seed"), kw.pop("n_bases"), **kw)
        g.write_fasta(fa)
        np.save(os.path.join(d, "hap.npy"), g.hap)
        log("[bench] genome generated in %.1fs" % (time.time() - t))
        t = time.time()
        index.build_index(fa, aligner=aligner)
        dt = time.time() - t
        log("[bench] index built in %.1fs (%s)" % (dt, "device, K7" if aligner is not None else "host"))
        if aligner is not None:
            aligner.index_build_info = {"seconds_incl_fasta_parse_and_file_write": dt, "builder": "device (K7 / K7w)",
                                        "sort_rounds": getattr(aligner, "last_index_sort_rounds", None)}
        open(done, "w").close()
    barrier()
    return fa


def make_batch(workload: str, step: int, rank: int, world: int):
    """Batch `step` of the read set for this rank (seeded; identical on every run)."""
    from bwbble_b200 import synth
    w = WORKLOADS[workload]
    hap = np.load(os.path.join(CACHE, w["index"], "hap.npy"), mmap_mode="r")
    g = synth.Genome([], np.asarray(hap), [], 0)
    seed = 1_000_003 * (step + 1) + rank
    return synth.make_reads(g, seed, w["batch"], with_names=False, bubble_frac=0.0, **w["reads"])


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu: int):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag = gpu, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                for ln in out.strip().splitlines():
                    self.rows.append([x.strip() for x in ln.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_cpu_reference(fa: str, reads, n_sample: int, params: dict, threads: int, want_stats: bool, n_stats: int = 0):
    """Time the reference CPU implementation on reads[0:n_sample].  Returns dict(value, kind, ...).
    This is the ONE place bench.py executes anything under oracle/ (the cpu_baseline / reference arm)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle
    from bwbble_b200 import default_params
    sub = reads.slice(0, n_sample)
    out = {"cores": threads, "sample": "first %d reads of step-0 batch, -n %d, %d threads" % (n_sample, params["n"], threads)}
    stats = None
    p = default_params(**params)
    if want_stats:      # instrumented restatement: rank-query count Q of the reference algorithm (on a prefix)
        n_stats = n_stats or n_sample
        qs = reads.slice(0, n_stats)
        orc = oracle.Oracle(fa + ".bwt")
        t = time.time()
        aln_bytes, stats = orc.align(qs.seq, qs.offsets, p, threads=threads)
        port_s = time.time() - t
        orc.close()
        out["port_reads_per_s"] = n_stats / port_s
        out["port_aln_bytes"] = aln_bytes
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "bwbble")
    if os.path.exists(ref_bin):
        with tempfile.TemporaryDirectory() as d:
            fq = os.path.join(d, "s.fq")
            sub.write_fastq(fq)
            tiny = os.path.join(d, "t.fq")
            sub.write_fastq(tiny, 0, 4)
            cmd = [ref_bin, "align", "-n", str(params["n"]), "-t", str(threads)]
            for k, flag in (("o", "-o"), ("e", "-e"), ("k", "-k"), ("l", "-l")):
                if k in params:
                    cmd += [flag, str(params[k])]
            cmd.append(fa)
            t = time.time()
            subprocess.run(cmd + [tiny, os.path.join(d, "t.aln")], check=True, stdout=subprocess.DEVNULL)
            load_s = time.time() - t          # index + FASTQ load (subtract a 4-read run)
            t = time.time()
            subprocess.run(cmd + [fq, os.path.join(d, "s.aln")], check=True, stdout=subprocess.DEVNULL)
            full_s = time.time() - t
            import hashlib
            out["ref_aln_md5"] = hashlib.md5(open(os.path.join(d, "s.aln"), "rb").read()).hexdigest()
        out.update(kind="reference", value=n_sample / max(full_s - load_s, 1e-6), load_s=load_s)
    else:
        if not want_stats:
            orc = oracle.Oracle(fa + ".bwt")
            t = time.time()
            orc.align(sub.seq, sub.offsets, p, threads=threads)
            port_s = time.time() - t
            orc.close()
            out["port_reads_per_s"] = n_sample / port_s
        out.update(kind="port", value=out["port_reads_per_s"])
    out["unit"] = "reads/s"
    return out, stats


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="chr21", choices=list(WORKLOADS))
    ap.add_argument("--index-chunk", type=int, default=0, help="K7w: suffixes per sort chunk (0 = default 2^29)")
    ap.add_argument("--opt", action="append", default=[], help="key=value for bwb_set_option (experiments)")
    ap.add_argument("--cli-reads", type=int, default=-1, help="reads of the FASTQ-file-to-.aln-file leg (e2e_cli); 0 = skip, -1 = auto")
    ap.add_argument("--cpu-sample", type=int, default=0, help="reads in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--quick", action="store_true", help="A/B experiments: resident timing only, short JSON")
    ap.add_argument("--batch", type=int, default=0, help="override reads per step per GPU (profiling)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    w = WORKLOADS[args.workload]
    if args.batch:
        w["batch"] = args.batch
    PARAMS = w["params"]
    cores = host_cores()

    if args.impl == "reference":
        if rank != 0:
            return 0
        fa = prepare_index(args.workload, 0, lambda: None)
        reads = make_batch(args.workload, 0, 0, 1)
        # one step = a bounded sample of the workload (every OpenMP thread gets >= 4096 reads of the chr21 sample,
        # so static chunking does not under-report the CPU); distinct slices of the step-0 batch per step
        n_s = args.cpu_sample or w["cpu_sample"]
        times = []
        meta = None
        for s in range(args.warmup + args.steps):
            res, _ = run_cpu_reference(fa, reads.slice((s * n_s) % (reads.n - n_s), (s * n_s) % (reads.n - n_s) + n_s),
                                       n_s, PARAMS, cores, want_stats=False)
            meta = res
            if s >= args.warmup:
                times.append(n_s / res["value"])
        ms = 1e3 * float(np.mean(times))
        val = n_s / (ms / 1e3)
        line = {"impl": "reference", "metric": "reads/sec (100bp, BWA-default diffs)", "value": val, "unit": "reads/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64/u32 integer",
                "data": "synthetic",
                "config": {"workload": w["desc"], "params": w["params_str"], "reads_per_step": n_s},
                "cpu_baseline": {"value": val, "unit": "reads/s", "cores": cores, "kind": meta["kind"],
                                 "sample": "%d reads per step, %d host threads" % (n_s, cores)},
                "e2e": {"value": val, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return 0

    import torch
    import torch.distributed as dist
    from bwbble_b200 import Aligner, default_params
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)

    def barrier():
        if world > 1:
            dist.barrier()

    al = Aligner([local_rank])
    for kv in args.opt:
        k, v = kv.split("=")
        al.set_option(k, int(v))
    if args.index_chunk:
        al.set_option("index_chunk", args.index_chunk)
    fa = prepare_index(args.workload, rank, barrier, aligner=al)
    p = default_params(**PARAMS)
    n_steps = args.warmup + args.steps
    t0 = time.time()
    # at most 4 distinct seeded batches, taken round-robin (every timed step of the default run sees a different
    # one; a batch is ~0.9 GB of reads + its arena traffic, far beyond L2 either way)
    nd = min(n_steps, 4)
    batches = [make_batch(args.workload, s, rank, world) for s in range(nd)]
    # the batches live in pinned host memory (one copy): the e2e leg copies straight out of it
    pinned = []
    for b in batches:
        ts, to = torch.from_numpy(b.seq).pin_memory(), torch.from_numpy(b.offsets.view(np.int64)).pin_memory()
        b.seq, b.offsets = ts.numpy(), to.numpy().view(np.uint64)
        pinned.append((ts, to))
    log("[bench] rank %d: %d batches x %d reads generated in %.1fs" % (rank, nd, w["batch"], time.time() - t0))

    al.load_index(fa + ".bwt")
    stream = torch.cuda.current_stream()
    al.set_stream(stream.cuda_stream)

    # ---- value: reads resident in HBM ------------------------------------------------------
    dev_reads = [al.upload_reads(b.seq, b.offsets) for b in batches]
    kernel_ms, k3_ms, hits_total, ctr_sum = [], [], 0, {}
    for s in range(args.warmup):
        al.align_resident(dev_reads[s % nd], p, fetch=False).close()
    torch.cuda.synchronize()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for s in range(args.warmup, n_steps):
        r = al.align_resident(dev_reads[s % nd], p, fetch=False)
        kernel_ms.append(r.kernel_ms)
        k3_ms.append(r.k3_ms)
        hits_total += r.num_hits
        for k, v in r.counters().items():
            ctr_sum[k] = max(ctr_sum.get(k, 0), v) if k.startswith("max") else ctr_sum.get(k, 0) + v
        r.close()
    ev1.record(stream)
    torch.cuda.synchronize()
    barrier()
    dev_ms = ev0.elapsed_time(ev1)

    if args.quick:
        print(json.dumps({"reads_per_s": w["batch"] * args.steps / (dev_ms / 1e3), "kernel_ms": kernel_ms,
                          "lib": os.environ.get("BWBBLE_B200_LIB", "default"), "ctr": ctr_sum}), flush=True)
        return 0

    # ---- e2e: host buffers through bwb_align (pinned H2D + D2H of every hit) -------------------
    h2d = int(np.mean([b.seq.nbytes + b.offsets.nbytes for b in batches]))
    d2h_bytes = []
    al.align(batches[0].seq, batches[0].offsets, p).close()
    torch.cuda.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_host0 = time.time()
    e0.record(stream)
    for s in range(args.warmup, n_steps):
        r = al.align(batches[s % nd].seq, batches[s % nd].offsets, p)
        d2h_bytes.append(4 * r.num_reads + 48 * r.num_hits + 256)
        r.close()
    e1.record(stream)
    torch.cuda.synchronize()
    e2e_ms = max(e0.elapsed_time(e1), 1e3 * (time.time() - t_host0))
    barrier()
    sampler.stop_flag = True
    sampler.join(timeout=3)

    if world > 1:
        t = torch.tensor([dev_ms, e2e_ms, max(kernel_ms)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms = float(t[0]), float(t[1])
    reads_per_step = w["batch"] * world
    value = reads_per_step * args.steps / (dev_ms / 1e3)
    e2e = reads_per_step * args.steps / (e2e_ms / 1e3)

    # ---- N > 1: the gather the north star names -- per-shard .aln records to rank 0 over NCCL, in input order.
    # Timed on its own (serialisation + gather inside the region); and the sharded stream of one common read set
    # must equal the stream rank 0 produces alone (per-shard hash check).
    gathered = None
    if world > 1:
        import hashlib
        from bwbble_b200 import synth
        from bwbble_b200.dist import align_sharded, gather_bytes
        hap = np.load(os.path.join(CACHE, w["index"], "hap.npy"), mmap_mode="r")
        common = synth.make_reads(synth.Genome([], np.asarray(hap), [], 0), 4242, 1 << 17, with_names=False, bubble_frac=0.0, **w["reads"])
        whole = align_sharded(lambda sq, of: al.align(sq, of, p).aln_bytes(), common.seq, common.offsets)
        alone = al.align(common.seq, common.offsets, p).aln_bytes() if rank == 0 else None
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tg = time.time()
        g0.record(stream)
        gsteps = min(2, args.steps)
        gbytes = 0
        for s in range(gsteps):
            r = al.align(batches[s % nd].seq, batches[s % nd].offsets, p)
            blob = r.aln_bytes()
            r.close()
            parts = gather_bytes(blob)
            if rank == 0:
                gbytes += sum(len(x) for x in parts)
        g1.record(stream)
        torch.cuda.synchronize()
        barrier()
        g_ms = max(g0.elapsed_time(g1), 1e3 * (time.time() - tg))
        tt = torch.tensor([g_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        if rank == 0:
            gathered = {"value": w["batch"] * world * gsteps / (float(tt[0]) / 1e3), "unit": "reads/s", "steps": gsteps,
                        "aln_bytes_on_rank0_per_step": gbytes // max(gsteps, 1),
                        "what": "bwb_align + .aln serialisation on every rank + NCCL gather of the shard streams to rank 0 (input order)",
                        "sharded_stream_equals_single_gpu": whole == alone, "check_reads": common.n,
                        "check_md5": hashlib.md5(whole).hexdigest()}

    # ---- e2e_cli: the real entry points, FASTQ file in, .aln file out --------------------------
    cli = None
    if rank == 0 and world == 1 and args.cli_reads != 0:
        try:
            n_cli = args.cli_reads if args.cli_reads > 0 else min(w["batch"], 1 << 23)
            cli_batch = max(1 << 16, n_cli // 4)            # 4 launches: parse k+1 | device k | write k-1 overlap; 2 M-read launches keep K4's tail at ~8 %
            sub = batches[0].slice(0, min(n_cli, batches[0].n))
            d = os.path.join(CACHE, "cli")
            os.makedirs(d, exist_ok=True)
            fq, aln = os.path.join(d, "r.fq"), os.path.join(d, "out.aln")
            sub.write_fastq(fq)
            t = time.time()
            n_done = al.align_fastq(fq, aln, p, batch=cli_batch)
            dt = time.time() - t
            cli = {"value": n_done / dt, "unit": "reads/s", "reads": n_done, "seconds": dt, "reads_per_launch": cli_batch,
                   "what": "bwb_align_fastq: FASTQ file -> parse -> H2D -> K3/K4/K5 -> D2H -> serialise -> .aln file (3 threads)",
                   "fastq_bytes": os.path.getsize(fq), "aln_bytes": os.path.getsize(aln)}
            import hashlib
            cli["aln_md5_equals_bwb_align"] = (hashlib.md5(open(aln, "rb").read()).hexdigest()
                                               == hashlib.md5(al.align(sub.seq, sub.offsets, p).aln_bytes()).hexdigest())
            gpu_bin = os.path.join(ROOT, "oracle", "_ref", "bwbble_gpu")
            if os.path.exists(gpu_bin):      # the reference's own main()/fastq2reads()/alns2alnf_bin around the shim
                cmd = [gpu_bin, "align", "-n", str(PARAMS["n"])]
                for k, flag in (("o", "-o"), ("e", "-e")):
                    if k in PARAMS:
                        cmd += [flag, str(PARAMS[k])]
                out2 = os.path.join(d, "dropin.aln")
                al_free = True
                t = time.time()
                r = subprocess.run(cmd + [fa, fq, out2], capture_output=True, text=True)
                dt2 = time.time() - t
                cli["dropin_binary"] = {"value": n_done / dt2, "unit": "reads/s", "seconds": dt2, "rc": r.returncode,
                                        "what": "oracle/_ref/bwbble_gpu align (reference main.o/align.o/io.o + shim), wall clock incl. "
                                                "index load, fastq2reads of all reads, context creation",
                                        "aln_identical": r.returncode == 0 and open(out2, "rb").read() == open(aln, "rb").read()}
        except Exception as ex:
            cli = {"error": repr(ex)}

    if rank == 0:
        cpu, stats, n_s, n_q = None, None, 0, 0
        if not args.no_cpu and world == 1:
            n_s = args.cpu_sample or w["cpu_sample"]
            n_q = min(n_s, 16384)            # instrumented oracle (Q of the reference algorithm + parity bytes) on a prefix
            cpu, stats = run_cpu_reference(fa, batches[args.warmup % nd], n_s, PARAMS, cores, want_stats=True, n_stats=n_q)
        elif not args.no_cpu:
            # N > 1: no CPU timing (rank 0 at N=1 only), but Q of the reference algorithm is still
            # counted on a small sample so that the roofline can be reported
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import oracle
            n_q = 4096
            sub = batches[args.warmup % nd].slice(0, n_q)
            orc = oracle.Oracle(fa + ".bwt")
            _, stats = orc.align(sub.seq, sub.offsets, default_params(**PARAMS), threads=cores)
            orc.close()
        parity = None
        if cpu is not None and "port_aln_bytes" in cpu:
            # spot check at bench scale: the sample's .aln stream from the device == the oracle's
            sub = batches[args.warmup % nd].slice(0, n_q)
            got = al.align(sub.seq, sub.offsets, p).aln_bytes()
            parity = {"reads": n_q, "identical_to_oracle": got == cpu.pop("port_aln_bytes"), "aln_bytes": len(got)}
            if cpu.get("ref_aln_md5"):
                # and the whole CPU sample against the .aln file the unmodified reference binary just wrote
                import hashlib
                sub = batches[args.warmup % nd].slice(0, n_s)
                got = al.align(sub.seq, sub.offsets, p).aln_bytes()
                parity.update(reference_binary_reads=n_s, identical_to_reference_binary=hashlib.md5(got).hexdigest() == cpu["ref_aln_md5"])
        occ = {}
        try:
            nq = 1 << 26
            for mode, name in ((0, "O(c,i): 1 thread/query"), (1, "O_alphabet(i): 16 lanes/query")):
                ms, _ = al.occ_bench(nq, seed=7, mode=mode, iters=5)
                occ[name] = {"queries_per_s": nq / (ms / 1e3), "GBps_at_128B_per_query": nq * 128 / (ms / 1e3) / 1e9}
        except Exception as ex:          # never let the micro-benchmark break the headline line
            occ = {"error": str(ex)}
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        roof = None
        if stats:
            from bwbble_b200 import load_bwt
            props = torch.cuda.get_device_properties(local_rank)
            l2_bytes = int(getattr(props, "L2_cache_size", 0) or 126 * 1024 * 1024)
            index_bytes = ((al.index_length() + 127) // 128) * 128
            # the Occ gathers of an index that fits L2 never reach HBM: its ceiling is the box's measured random
            # 128-byte-gather rate on THIS index (K1, uniform random rows), not the HBM copy bandwidth
            gather_peak = max((v.get("GBps_at_128B_per_query", 0.0) for v in occ.values() if isinstance(v, dict)), default=0.0)
            bound = "l2" if index_bytes <= l2_bytes else "hbm"
            peak = gather_peak if (bound == "l2" and gather_peak > 0) else hbm_peak
            q_per_read = (stats["n_O"] + stats["n_Oalpha"]) / n_q
            k4_only = float(np.mean(kernel_ms))
            k_ms = k4_only + float(np.mean(k3_ms))      # the reference's Q spans calculate_d (K3) and inexact_match (K4)
            achieved = w["batch"] * q_per_read * 128 / (k_ms / 1e3) / 1e9
            physical = ctr_sum.get("rank_queries", 0) / args.steps * 128 / (k_ms / 1e3) / 1e9
            traffic, traffic_src, dram_frac = None, None, None
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[args.workload]["k_search_l"]
                traffic = tj["dram_bytes_per_read"] * w["batch"] / 1e9      # GB per launch, from the ncu capture of THIS workload
                traffic_src = tj["source"]
                dram_frac = traffic / (k4_only / 1e3) / hbm_peak
            except Exception:
                pass
            roof = {"bound": bound, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "peak_source": ("measured random 128-B gather rate of K1 on this index (occ_gather, same run): the index "
                                    "(%d MB) fits the %d MB L2" % (index_bytes >> 20, l2_bytes >> 20)) if bound == "l2" and gather_peak > 0
                                   else ("MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s"),
                    "frac_physical": (physical / gather_peak) if gather_peak > 0 else None,
                    "frac_of_hbm_copy_peak": achieved / hbm_peak, "hbm_copy_peak": hbm_peak, "gather_ceiling": gather_peak or None,
                    "traffic": traffic, "traffic_unit": "GB of DRAM read+write per launch", "traffic_source": traffic_src,
                    "dram_frac": dram_frac,
                    "algorithmic_gb_per_launch": w["batch"] * q_per_read * 128 / 1e9,
                    "kernel": "k_calc_d_g + k_search_l (K3 lower bounds + K4 search; K4 is the dominant launch)",
                    "kernel_ms_per_launch": k_ms, "k4_ms_per_launch": k4_only, "k3_ms_per_launch": float(np.mean(k3_ms)),
                    "kernel_share_of_step": k_ms / (dev_ms / args.steps),
                    "rank_queries_per_read_reference": q_per_read,
                    "bytes_per_query": 128,
                    "physical_block_loads_per_read": ctr_sum.get("rank_queries", 0) / (w["batch"] * args.steps),
                    "physical_gbs": physical,
                    "note": "achieved = the REFERENCE algorithm's rank queries on these reads (instrumented oracle, SURVEY 8d) x 128 B / "
                            "(K3 + K4 time).  frac divides by the ceiling of the regime the index is in (L2-resident: measured "
                            "gather rate on this index; else the measured HBM copy peak).  The device does fewer PHYSICAL block "
                            "loads than the reference issues queries (both interval ends and all 15 codes from one or two "
                            "blocks, K0b table for the top of calculate_d): frac_physical = those loads x 128 B / time / gather ceiling"}
        line = {"metric": "reads/sec (100bp, BWA-default diffs)" if w["reads"]["read_len"] == 100 and PARAMS.get("n") == 5
                          else "reads/sec (%dbp, %s)" % (w["reads"]["read_len"], w["params_str"]),
                "value": value, "unit": "reads/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64/u32 integer",
                "data": "synthetic",
                "config": {"workload": w["desc"], "workload_key": args.workload, "params": w["params_str"],
                           "reads_per_step": reads_per_step, "reads_per_step_per_gpu": w["batch"],
                           "l2": "inputs larger than L2: index %d MB + %d MB of reads + GBs of heap arena per step; "
                                 "%d distinct seeded batches round-robin, the timed steps all differ"
                                 % (al.index_length() >> 20, w["batch"] * (w["reads"]["read_len"] + 8) >> 20, nd),
                           "parallelism": "reads sharded x%d, index replicated, no collective" % world},
                "e2e": {"value": e2e, "unit": "reads/s", "h2d_bytes_per_step": h2d * world,
                        "d2h_bytes_per_step": int(np.mean(d2h_bytes)) * world},
                "gpu_launches": 8 * args.steps,      # K3, K3b hist + scatter, K4 + its 2 retry passes, K5 scan + emit
                "clocks": sampler.summary(),
                "roofline": roof, "cpu_baseline": None if cpu is None else
                {"value": cpu["value"], "unit": "reads/s", "cores": cpu["cores"], "kind": cpu["kind"], "sample": cpu["sample"],
                 "port_reads_per_s": cpu.get("port_reads_per_s")},
                "parity_check": parity, "occ_gather": occ, "e2e_gathered": gathered, "e2e_cli": cli,
                "index_build": getattr(al, "index_build_info", None),
                "counters_per_read": {k: (v if k.startswith("max") else v / (w["batch"] * args.steps)) for k, v in ctr_sum.items()},
                "hits_per_read": hits_total / (w["batch"] * args.steps)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
