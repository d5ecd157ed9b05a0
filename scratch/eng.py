import sys, json, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from bwbble_b200 import Aligner, default_params
eng=int(sys.argv[1]); batch=int(sys.argv[2]); bps=int(sys.argv[3]) if len(sys.argv)>3 else 0
bench.WORKLOADS['chr21']['batch']=batch
fa=bench.prepare_index('chr21',0,lambda:None)
b=bench.make_batch('chr21',0,0,1)
al=Aligner([0]); 
if eng: al.set_option('engine',eng)
if bps: al.set_option('blocks_per_sm',bps)
al.load_index(fa+'.bwt')
dr=al.upload_reads(b.seq,b.offsets)
p=default_params(n=5)
for it in range(2):
    r=al.align_resident(dr,p,fetch=False); ms=r.kernel_ms; c=r.counters(); r.close()
print("RESULT "+json.dumps({'lib':os.environ.get('BWBBLE_B200_LIB','default').split('/')[-1],'engine':eng,'batch':batch,'bps':bps,'kernel_ms':ms,'reads_per_s':batch/ms*1e3,'pops':c['pops']}))
