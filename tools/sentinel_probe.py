"""Development probe (GPU box): agreement of the -3 'negative interpolated sample' rule between the CUDA path and the
oracle on black-background patterns, with a breakdown of any mismatch.  Not a test; tests/test_gpu_sentinel.py asserts."""
import os
import sys
import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import opencorr_b200 as ob
from opencorr_b200 import synth
from oracle.oracle import Oracle2D, Oracle3D


def patterns_2d():
    ref, tar = synth.speckle_pair_2d(512, 512)
    yield "shift24x1.1", np.clip(ref - 24.0, 0, 255).astype(np.float32) * 1.1, np.clip(tar - 24.0, 0, 255).astype(np.float32) * 1.1
    r0, t0 = synth.speckle_pair_2d(512, 512, background=0.0)
    yield "background0", r0, t0
    yield "threshold60", np.where(ref < 60, 0, ref).astype(np.float32), np.where(tar < 60, 0, tar).astype(np.float32)
    r1, t1 = synth.speckle_pair_2d(512, 512, background=0.0, rho=3.5, seed=7)
    yield "background0_rho3.5", r1, t1
    r2, t2 = synth.speckle_pair_2d(2048, 512, background=2.0, seed=11)
    yield "background2_wide", r2, t2


def run_2d(eng, order):
    for name, ref, tar in patterns_2d():
        h, w = ref.shape
        xy = synth.grid_2d(40, 40, (w - 80) // 9, (h - 80) // 9, 9, 9)
        q = ob.make_poi2d(xy)
        o = Oracle2D(ref, tar)
        o.fftcc2d(q, 16, 16)
        qg, qc = q.copy(), q.copy()
        eng.set_images_2d(ref, tar)
        eng.icgn2d_prepare()
        (eng.icgn2d1 if order == 1 else eng.icgn2d2)(qg, 16, 16, 0.001, 10)
        (o.icgn2d1 if order == 1 else o.icgn2d2)(qc, 16, 16, 0.001, 10)
        a, b = qg[:, 16], qc[:, 16]
        mism = np.where((a == -3) != (b == -3))[0]
        print("2D order %d %-20s n=%d oracle -3: %d  gpu -3: %d  mismatches: %d" % (order, name, len(q), (b == -3).sum(), (a == -3).sum(), len(mism)))
        for i in mism[:8]:
            print("    poi %d (%.0f,%.0f) gpu zncc %.6f it %.0f | oracle zncc %.6f it %.0f" % (i, xy[i, 0], xy[i, 1], a[i], qg[i, 17], b[i], qc[i, 17]))


def run_3d(eng):
    for name, bg, thr in (("background0", 0.0, None), ("threshold50", 24.0, 50.0), ("background1", 1.0, None)):
        ref, tar = synth.speckle_pair_3d(96, 88, 80, background=bg)
        if thr is not None:
            ref, tar = np.where(ref < thr, 0, ref).astype(np.float32), np.where(tar < thr, 0, tar).astype(np.float32)
        xyz = synth.grid_3d(24, 24, 24, 8, 7, 6, 6, 6, 6)
        q = ob.make_poi3d(xyz)
        o = Oracle3D(ref, tar)
        o.fftcc3d(q, 8, 8, 8)
        qg, qc = q.copy(), q.copy()
        eng.set_images_3d(ref, tar)
        eng.icgn3d_prepare()
        eng.icgn3d1(qg, 8, 8, 8, 0.001, 20)
        o.icgn3d1(qc, 8, 8, 8, 0.001, 20)
        a, b = qg[:, 18], qc[:, 18]
        mism = np.where((a == -3) != (b == -3))[0]
        print("3D %-20s n=%d oracle -3: %d  gpu -3: %d  mismatches: %d" % (name, len(q), (b == -3).sum(), (a == -3).sum(), len(mism)))
        for i in mism[:8]:
            print("    poi %d gpu zncc %.6f it %.0f | oracle zncc %.6f it %.0f" % (i, a[i], qg[i, 19], b[i], qc[i, 19]))


if __name__ == "__main__":
    eng = ob.Engine(0)
    run_2d(eng, 1)
    run_2d(eng, 2)
    run_3d(eng)
