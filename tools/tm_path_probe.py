import sys, time
t0 = time.time()
sys.path.insert(0, ".")
import numpy as np
import opencorr_b200 as ob
from opencorr_b200 import synth
from oracle.oracle import Oracle2D
sys.path.insert(0, "tests")
from test_gpu_sentinel import _discs
ref, tar = synth.speckle_pair_2d(704, 704)
holes = _discs((704, 704), 60, 4, 14, 3)
ref, tar = np.where(holes, 0, ref).astype(np.float32), np.where(holes, 0, tar).astype(np.float32)
xy = synth.grid_2d(40, 40, 70, 70, 9, 9)  # 4900 POIs: the Tensor-Memory variant (>= 2368)
q = ob.make_poi2d(xy)
o = Oracle2D(ref, tar)
o.fftcc2d(q, 16, 16)
qg, qc = q.copy(), q.copy()
eng = ob.Engine(0)
eng.set_images_2d(ref, tar)
eng.icgn2d_prepare()
eng.icgn2d1(qg, 16, 16, 0.001, 10)
o.icgn2d1(qc, 16, 16, 0.001, 10)
a, b = qg[:, 16], qc[:, 16]
print("TM sentinel: n", len(q), "oracle -3:", int((b == -3).sum()), "mismatch:", int(((a == -3) != (b == -3)).sum()))
ok = (a >= 0) & (b >= 0) & (qg[:, 17] == qc[:, 17])
print("  same-iteration POIs", int(ok.sum()), "max |du,dv|", float(np.abs(qg[ok][:, [2, 8]] - qc[ok][:, [2, 8]]).max()), "max dZNCC", float(np.abs(a[ok] - b[ok]).max()))
# slow path at TM size: 12 % stretch (samples leave the staged tile)
ref, _ = synth.speckle_pair_2d(704, 704)
yy, xx = np.mgrid[0:704, 0:704].astype(np.float32)
o_ref = Oracle2D(ref, ref); o_ref.prepare()
src = np.stack([(352 + (xx - 352) / 1.12).ravel(), (352 + (yy - 352) / 1.12).ravel()], 1)
tar = np.clip(o_ref.bicubic(src), 0, 255).reshape(704, 704).astype(np.float32)
xy = synth.grid_2d(100, 100, 56, 56, 9, 9)  # 3136 POIs
q = ob.make_poi2d(xy)
q[:, 2] = (xy[:, 0] - 352) * 0.12; q[:, 8] = (xy[:, 1] - 352) * 0.12; q[:, 3] = 0.12; q[:, 10] = 0.12
qg, qc = q.copy(), q.copy()
eng.set_images_2d(ref, tar); eng.icgn2d_prepare(); eng.icgn2d1(qg, 16, 16, 0.001, 10)
Oracle2D(ref, tar).icgn2d1(qc, 16, 16, 0.001, 10)
ok = (qg[:, 16] >= 0) & (qc[:, 16] >= 0) & (qg[:, 17] == qc[:, 17])
print("TM slow path: n", len(q), "codes equal", bool(np.array_equal(qg[:, 16] < 0, qc[:, 16] < 0)), "same-iteration", int(ok.sum()),
      "max |du,dv|", float(np.abs(qg[ok][:, [2, 8]] - qc[ok][:, [2, 8]]).max()), "max dZNCC", float(np.abs(qg[ok, 16] - qc[ok, 16]).max()), "t", round(time.time() - t0, 1))
