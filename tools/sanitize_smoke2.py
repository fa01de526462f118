"""compute-sanitizer coverage of the section-8(f) kernels: ICLM, NR2D1, Strain 2D/3D (incl. the k-nearest fallback), EpipolarSearch."""
import sys; sys.path.insert(0, '.')
import numpy as np, opencorr_b200 as ob
from opencorr_b200 import synth
ref, tar = synth.speckle_pair_2d(200, 180); xy = synth.grid_2d(30, 30, 6, 5, 22, 25)
e = ob.Engine(0)
e.set_images_2d(ref, tar)
for r in (16, 9):
    q = ob.make_poi2d(xy); e.fftcc2d(q, r, r); e.icgn2d_prepare()
    a = q.copy(); e.iclm2d(1, a, r, r, 0.001, 10); b = q.copy(); e.iclm2d(2, b, r, r, 0.001, 10)
    e.nr2d_prepare(); c = q.copy(); e.nr2d1(c, r, r, 0.001, 10)
    print('r', r, 'iclm1', (a[:, 16] > 0).sum(), 'iclm2', (b[:, 16] > 0).sum(), 'nr', (c[:, 16] > 0).sum())
e.strain(c, 30.0, 5); print('strain2d', (c[:, 20] != 0).sum())
rng = np.random.default_rng(0)
s = ob.make_poi2d(rng.uniform(0, 500, (300, 2)).astype(np.float32)); s[:, 2] = 0.01 * s[:, 0]; s[:, 16] = 0.95; s[3, 0] = np.nan
e.strain(s, 10.0, 5); print('strain2d knn', (s[:, 20] != 0).sum())
p3 = ob.make_poi3d(rng.uniform(0, 100, (2000, 3)).astype(np.float32)); p3[:, 3] = 0.02 * p3[:, 1]; p3[:, 18] = rng.uniform(0.8, 1, 2000)
e.strain(p3, 15.0, 5, 0.9, 2); print('strain3d', (p3[:, 22] != 0).sum())
fm = np.array([[0, 0, 0], [0, 0, -1], [0, 1, 0]], np.float32)
q = ob.make_poi2d(xy); e.icgn2d_prepare(); e.epipolar_search2d(q, fm, [0, 0, 1], [0, 0, 0], 20, 4, 10, 10, 0.05, 5); print('epipolar', (q[:, 16] > 0.5).sum())
e.close(); print('done')
# register-codelet FFT-CC kernels (2D N=20, 40; 3D N=12, 20) and the one-warp-per-POI IC-GN path (queue larger than the resident slots)
import numpy as np, opencorr_b200 as ob
from opencorr_b200 import synth
e = ob.Engine(0)
ref, tar = synth.speckle_pair_2d(400, 300); e.set_images_2d(ref, tar)
for r in (10, 20):
    q = ob.make_poi2d(synth.grid_2d(40, 40, 13, 11, 24, 20)); e.fftcc2d(q, r, r); print('fftcc2d_reg r', r, (q[:, 16] > 0.5).sum())
big = ob.make_poi2d(synth.grid_2d(30, 30, 60, 40, 5, 6)); e.fftcc2d(big, 16, 16); e.icgn2d_prepare(); e.icgn2d1(big, 16, 16, 0.001, 10); print('icgn2d1 wpp1', (big[:, 16] > 0.9).sum(), len(big))
r3, t3 = synth.speckle_pair_3d(64, 60, 56); e.set_images_3d(r3, t3)
for r in (6, 10):
    q = ob.make_poi3d(np.array([[32, 30, 28], [30, 28, 27]], np.float32)); e.fftcc3d(q, r, r, r); print('fftcc3d_reg r', r, q[:, 18])
e.close(); print('done2')
