import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from srcfinder_b200 import ColumnwiseMF, synth
cube = synth.make_cube(640, 12, seed=51, bad_pixels=True)
active=[351,422]; ab = synth.load_ch4_library()[350:422,2]
L,B,S = cube.shape
def snap(eng):
    return dict(mask=eng.mask(), n=eng.nvalid(), mu=eng.mu(), ai=eng.alpha_index(), mf=eng.mf(), w=eng.weights())
def cmp(a,b,tag):
    for k in a:
        same = np.array_equal(a[k], b[k], equal_nan=True)
        d = 0 if same else np.nanmax(np.abs(a[k].astype(float)-b[k].astype(float)))
        print(tag, k, 'same' if same else 'DIFF max %g'%d, flush=True)
with ColumnwiseMF(L,B,S,active,ab) as eng:
    t=time.time(); eng.upload(cube); eng.run(); r1 = snap(eng); print('upload run', time.time()-t)
    eng.run(); r2 = snap(eng); cmp(r1,r2,'rerun')
    slab = torch.from_numpy(np.ascontiguousarray(cube[:,350:422,:])).cuda(); torch.cuda.synchronize()
    eng.bind_device(slab.data_ptr()); eng.run(); r3 = snap(eng); cmp(r1,r3,'compact-bind')
    dev = torch.from_numpy(cube).cuda(); torch.cuda.synchronize()
    eng.bind_device(dev.data_ptr()+350*S*4, line_pitch=B*S, band_pitch=S); eng.run(); r4=snap(eng); cmp(r1,r4,'full-bind')
    print('n', r1['n'], r4['n'])
    print('mu0', r1['mu'][0,:4], r4['mu'][0,:4])
