"""scratch: chr21-scale index, host builder vs K7 (device), byte comparison + timings"""
import os, sys, time, shutil, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from bwbble_b200 import Aligner, index
wl = sys.argv[1] if len(sys.argv) > 1 else 'chr21'
t0 = time.time()
fa = bench.prepare_index(wl, 0, lambda: None)          # host builder (cached per box)
t_host = time.time() - t0
fb = fa + ".dev.fa"
shutil.copy(fa, fb)
with Aligner([0]) as al:
    t0 = time.time()
    index.build_index(fb, aligner=al)
    t_dev = time.time() - t0
    rounds = al.last_index_sort_rounds
same = open(fa + ".bwt", "rb").read() == open(fb + ".bwt", "rb").read()
print("RESULT " + json.dumps({"workload": wl, "host_total_s_incl_genome": t_host, "device_s": t_dev, "rounds": rounds,
                              "identical": same, "bwt_bytes": os.path.getsize(fa + ".bwt")}))
