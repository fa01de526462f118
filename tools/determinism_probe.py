"""Development probe (GPU box): is a small-queue ICGN2D1 run (two warps per POI) bit-reproducible from run to run?"""
import os
import sys
import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import opencorr_b200 as ob
from opencorr_b200 import synth
from oracle.oracle import Oracle2D

ref, tar = synth.speckle_pair_2d(320, 300)
xy = synth.grid_2d(40, 40, 10, 9, 24, 24)
seed = ob.make_poi2d(xy)
Oracle2D(ref, tar).fftcc2d(seed, 16, 16)
eng = ob.Engine(0)
eng.set_images_2d(ref, tar)
eng.icgn2d_prepare()
outs = []
for i in range(12):
    q = seed.copy()
    if i % 3 == 2:  # re-upload in between, like the interleaving test
        eng.set_images_2d(ref, tar)
        eng.icgn2d_prepare()
    eng.icgn2d1(q, 16, 16, 0.001, 10)
    outs.append(q)
distinct = []
for q in outs:
    if not any(np.array_equal(q, d) for d in distinct):
        distinct.append(q)
print("WPP env", os.environ.get("OCB_ICGN2D_WPP"), "lib", os.environ.get("OCB_LIB_PATH", "default"), "-> distinct results:", len(distinct))
if len(distinct) > 1:
    d = np.abs(distinct[0] - distinct[1])
    print("  max diff %.3g at column %d, POIs differing: %d" % (d.max(), int(np.argmax(d.max(0))), int((d.max(1) > 0).sum())))
