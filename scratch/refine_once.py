import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from helpers import *
from minrms_b200 import SearchOrCuda, OrientationGenerator
z=np.load('/root/repo/tests/golden/ref_c1.npz'); g={k:z[k] for k in z.files}
s=SearchOrCuda(g['A'],g['B'],settings_from_golden(g))
og=OrientationGenerator(s); s.Search(og); print('evals', s.Refine())
