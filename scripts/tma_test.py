import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, mesh_reconstruction_b200 as mr
from mesh_reconstruction_b200 import synth
sc=synth.make_scene(160,128,2,step=0.1)
a,b=sc.frame(0),sc.frame(1)
lib=mr.load_library(); lib.mr_set_vr_impl(1)
f=mr.calculateFlow(a,b); print('ok',np.abs(f).max())
