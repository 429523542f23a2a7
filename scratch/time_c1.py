import sys, time, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from helpers import *
from minrms_b200 import SearchOrCuda, OrientationGenerator
for cfg in ('c1','c2'):
    z=np.load('/root/repo/tests/golden/ref_%s.npz'%cfg); g={k:z[k] for k in z.files}
    s=SearchOrCuda(g['A'],g['B'],settings_from_golden(g))
    t0=time.perf_counter(); og=OrientationGenerator(s); t1=time.perf_counter()
    for rep in range(3):
        s.reset_best(); s.stats(reset=True)
        t2=time.perf_counter(); s.Search(og); t3=time.perf_counter()
        st=s.stats()
        t4=time.perf_counter(); ev=s.Refine(); t5=time.perf_counter()
    print(cfg,'orient gen %.1f ms; search %.2f ms (nw %.2f dp %.2f tb %.2f) orientations %d dp %d; refine %.1f ms (%d evals, %d batches, %d launched)'%((t1-t0)*1e3,(t3-t2)*1e3,st['nw_kernel_ms'],st['dp_kernel_ms'],st['tb_kernel_ms'],st['orientations'],st['orientations_dp'],(t5-t4)*1e3,ev,s.stats()['refine_batches'],s.stats()['refine_evals_launched']))
