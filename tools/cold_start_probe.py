"""Development probe (GPU box): where does the first call's time go?  Fresh process, C-ABI through ctypes."""
import ctypes
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
t0 = time.perf_counter()
import numpy as np
from opencorr_b200 import _capi, synth
import opencorr_b200 as ob
t_import = time.perf_counter() - t0

ref, tar = synth.speckle_pair_2d(512, 512)
xy = synth.grid_2d(40, 40, 30, 30, 14, 14)
q = ob.make_poi2d(xy)
lib = _capi.load()


def T(label, f):
    t = time.perf_counter()
    r = f()
    print("%-28s %8.1f ms" % (label, 1e3 * (time.perf_counter() - t)))
    return r


vp = lambda a: ctypes.c_void_p(a.ctypes.data)
print("import numpy+package          %8.1f ms" % (1e3 * t_import))
T("ocb_device_count", lambda: lib.ocb_device_count())
ctx = T("ocb_create(0)", lambda: lib.ocb_create(0))
T("set_images_2d (512^2)", lambda: (lib.ocb_set_images_2d(ctx, vp(ref), vp(tar), 512, 512, 0), lib.ocb_sync(ctx)))
T("fftcc2d first", lambda: lib.ocb_fftcc2d(ctx, vp(q), len(q), 16, 16))
T("fftcc2d second", lambda: lib.ocb_fftcc2d(ctx, vp(q), len(q), 16, 16))
lib.ocb_icgn2d_prepare(ctx)
T("icgn2d1 first", lambda: lib.ocb_icgn2d1(ctx, vp(q), len(q), 16, 16, ctypes.c_float(1e-3), ctypes.c_float(10)))
T("icgn2d1 second", lambda: lib.ocb_icgn2d1(ctx, vp(q), len(q), 16, 16, ctypes.c_float(1e-3), ctypes.c_float(10)))
T("icgn2d2 first (r=16)", lambda: lib.ocb_icgn2d2(ctx, vp(q), len(q), 16, 16, ctypes.c_float(1e-3), ctypes.c_float(10)))
