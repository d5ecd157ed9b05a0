import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import bench, oracle
from bwbble_b200 import default_params
a = np.load('/tmp/pops_vs_d.npy')
n = len(a)
bench.WORKLOADS['chr21']['batch'] = 1 << 16
fa = bench.prepare_index('chr21', 0, lambda: None)
b = bench.make_batch('chr21', 0, 0, 1)
orc = oracle.Oracle(fa + '.bwt')
p = default_params(n=5)
feat = []
for r in range(n):
    s = b.seq[int(b.offsets[r]):int(b.offsets[r + 1])]
    D = orc.calculate_d(s, 0, p)
    Ds = orc.calculate_d(s, 32, p)
    # position of the first / last increment of D
    z = D[:len(s), 0]
    inc = np.nonzero(np.diff(np.concatenate([[0], z])))[0]
    feat.append((int(Ds[31, 0]), int(inc[0]) if len(inc) else -1, int(inc[-1]) if len(inc) else -1))
f = np.array(feat)
np.save('/tmp/pops_feat.npy', f)
for d in range(3):
    for ds in range(3):
        m = a[(a[:, 0] == d) & (f[:, 0] == ds)]
        if len(m): print("D=%d Ds=%d n=%d pops mean %.0f median %.0f" % (d, ds, len(m), m[:, 1].mean(), np.median(m[:, 1])))
mapped = a[:, 2] > 0
for d in (1, 2):
    sel = (a[:, 0] == d) & mapped
    for lo, hi in ((0, 20), (20, 40), (40, 60), (60, 80), (80, 100)):
        m = a[sel & (f[:, 1] >= lo) & (f[:, 1] < hi)]
        if len(m): print("D=%d first-inc in [%d,%d) n=%d pops mean %.0f" % (d, lo, hi, len(m), m[:, 1].mean()))
