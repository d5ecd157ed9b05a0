"""A/B helper (not part of the product): one resident batch through a chosen engine.
usage: python scratch/eng.py <engine 0|1|2> <batch> [blocks_per_sm]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from bwbble_b200 import Aligner, default_params

eng = int(sys.argv[1])
batch = int(sys.argv[2])
bps = int(sys.argv[3]) if len(sys.argv) > 3 else 0
use_p = len(sys.argv) > 4 and sys.argv[4] == 'P'
import time
bench.WORKLOADS['chr21']['batch'] = batch
fa = bench.prepare_index('chr21', 0, lambda: None)
b = bench.make_batch('chr21', 0, 0, 1)
al = Aligner([0])
if eng:
    al.set_option('engine', eng)
if bps:
    al.set_option('blocks_per_sm', bps)
opts = os.environ.get('ENG_OPTS', '')
for kv in filter(None, opts.split(',')):
    k, v = kv.split('=')
    al.set_option(k, int(v))
al.load_index(fa + '.bwt')
dr = al.upload_reads(b.seq, b.offsets)
p = default_params(n=5, use_precalc=int(use_p))
pre_s = 0.0
if use_p:
    t0 = time.time()
    al.build_precalc(True)
    pre_s = time.time() - t0
for it in range(2):
    r = al.align_resident(dr, p, fetch=False)
    ms, k3, c = r.kernel_ms, r.k3_ms, r.counters()
    r.close()
print("RESULT " + json.dumps({'lib': os.environ.get('BWBBLE_B200_LIB', 'default').split('/')[-1], 'engine': eng, 'opts': opts,
                              'batch': batch, 'bps': bps, 'k4_ms': ms, 'k3_ms': k3, 'precalc_build_s': pre_s, 'pre_intervals': al.precalc_num_intervals() if use_p else 0,
                              'reads_per_s': batch / (ms + k3) * 1e3, 'pops': c['pops']}))
