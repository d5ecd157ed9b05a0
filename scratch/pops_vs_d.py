"""scratch analysis: per-read pops of the restated search against K3's whole-read lower bound D[len-1]."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import bench, oracle
from bwbble_b200 import default_params
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
bench.WORKLOADS['chr21']['batch'] = 1 << 16
fa = bench.prepare_index('chr21', 0, lambda: None)
b = bench.make_batch('chr21', 0, 0, 1)
orc = oracle.Oracle(fa + '.bwt')
p = default_params(n=5)
rows = []
for r in range(n):
    s = b.seq[int(b.offsets[r]):int(b.offsets[r + 1])]
    off = np.array([0, len(s)], dtype=np.uint64)
    _, st = orc.align(s, off, p)
    D = orc.calculate_d(s, 0, p)
    rows.append((int(D[len(s) - 1, 0]), int(st['pops']), int(st['hits']), int((s > 3).sum())))
a = np.array(rows)
np.save('/tmp/pops_vs_d.npy', a)
for d in sorted(set(a[:, 0])):
    m = a[a[:, 0] == d]
    print("D=%d n=%d pops mean %.0f median %.0f max %d  unmapped %d" % (d, len(m), m[:, 1].mean(), np.median(m[:, 1]), m[:, 1].max(), (m[:, 2] == 0).sum()))
print("total pops", a[:, 1].sum(), "top 1%% share %.2f" % (np.sort(a[:, 1])[-max(1, n // 100):].sum() / a[:, 1].sum()))
