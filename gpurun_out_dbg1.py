import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from helpers import *
from minrms_b200 import SearchOrCuda
for cfg in ('c1','c2'):
    z=np.load('/root/repo/tests/golden/ref_%s.npz'%cfg); g={k:z[k] for k in z.files}
    s=SearchOrCuda(g['A'],g['B'],settings_from_golden(g))
    s1=SearchOrCuda(g['A'],g['B'],settings_from_golden(g)); s1.set_option('dp_kernel',1)
    bad=0
    for k,o in enumerate(g['sample_ids']):
        r=s.eval_one(g['sample_orient'][k]); r1=s1.eval_one(g['sample_orient'][k])
        nl=int(g['sample_n_last'][k])
        if nl==0: continue
        d=np.nonzero(bits(r['cost'][:nl])!=bits(g['sample_cost'][k][:nl]))[0]
        if len(d):
            bad+=1
            print(cfg,'o',o,'nl',nl,'mismatch n=',(d+1)[:10],'got',r['cost'][d[:5]],'exp',g['sample_cost'][k][d[:5]], 'general', r1['cost'][d[:5]])
        v=g['sample_valid'][k].astype(bool)
        if not np.array_equal(bits(r['ssd'][v]),bits(g['sample_ssd'][k][v])): print(cfg,'o',o,'ssd mismatch')
    print(cfg,'bad',bad)
