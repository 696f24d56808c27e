"""Golden vectors of the occupancy post-step: outputs of the REFERENCE's marching cubes (oracle/_ref/libmcubes_ref.so, built by
oracle/Makefile from /root/reference/convocc/src/utils/libmcubes) and of Generator3D.extract_mesh's arithmetic on top of it, on
the seeded volumes of oracle/cases.py.  Run in the build container:  make -C oracle && python oracle/make_golden_mcubes.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cases, mcubes_oracle as mo  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')
d = {}
for name, shape, iso in (('a', (20, 17, 23), 0.1), ('b', (9, 9, 9), -0.25), ('c', (1, 6, 7), 0.0)):
    vol = cases.mcubes_volume(shape, seed=len(d))
    ref = mo.ref_marching_cubes(vol.numpy().astype(np.float64), iso)
    assert ref is not None, "build oracle/_ref first (make -C oracle)"
    mine = mo.marching_cubes(vol.numpy(), iso)
    assert np.array_equal(ref[0], mine[0]) and np.array_equal(ref[1], mine[1])
    d[name] = {'shape': shape, 'iso': iso, 'vol_sum': float(vol.double().sum()), 'vertices': torch.from_numpy(ref[0]),
               'triangles': torch.from_numpy(ref[1])}
    print(name, shape, ref[0].shape, ref[1].shape)
# extract_mesh: logit grid 24^3, threshold 0.2, padding 0.1 (the reference's generation settings)
vol = cases.mcubes_volume((24, 24, 24), seed=9, noise=0.5, scale=1.5)
v, t = mo.extract_mesh(vol.numpy(), 0.2, 0.1, mc=mo.ref_marching_cubes)
v2, t2 = mo.extract_mesh(vol.numpy(), 0.2, 0.1)
assert np.array_equal(v, v2) and np.array_equal(t, t2)
d['mesh'] = {'shape': (24, 24, 24), 'vol_sum': float(vol.double().sum()), 'vertices': torch.from_numpy(v), 'triangles': torch.from_numpy(t)}
print('mesh', v.shape, t.shape)
torch.save(d, os.path.join(OUT, 'mcubes.pt'))
print('wrote', os.path.join(OUT, 'mcubes.pt'), os.path.getsize(os.path.join(OUT, 'mcubes.pt')), 'bytes')
