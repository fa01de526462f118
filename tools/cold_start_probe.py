"""Where does a fresh process's first call go?  C ABI through ctypes; prints one JSON object (bench.py's "cold_start" record).
The interpreter / numpy start-up is not part of it: the clock starts at the first library call."""
import ctypes
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from opencorr_b200 import _capi, synth
import opencorr_b200 as ob

ref, tar = synth.speckle_pair_2d(512, 512)
xy = synth.grid_2d(40, 40, 30, 30, 14, 14)
q = ob.make_poi2d(xy)
lib = _capi.load()
vp = lambda a: ctypes.c_void_p(a.ctypes.data)
out = {}


def T(label, f):
    t = time.perf_counter()
    r = f()
    out[label] = round(time.perf_counter() - t, 6)
    return r


t_all = time.perf_counter()
T("device_count_s", lambda: lib.ocb_device_count())  # cuInit
ctx = T("create_s", lambda: lib.ocb_create(int(os.environ.get("OCB_PROBE_DEVICE", "0"))))  # primary context, stream, first cudaMalloc
T("set_images_s", lambda: (lib.ocb_set_images_2d(ctx, vp(ref), vp(tar), 512, 512, 0), lib.ocb_sync(ctx)))
T("fftcc2d_first_s", lambda: lib.ocb_fftcc2d(ctx, vp(q), len(q), 16, 16))  # module load of the first kernel
T("fftcc2d_second_s", lambda: lib.ocb_fftcc2d(ctx, vp(q), len(q), 16, 16))
lib.ocb_icgn2d_prepare(ctx)
T("icgn2d1_first_s", lambda: lib.ocb_icgn2d1(ctx, vp(q), len(q), 16, 16, ctypes.c_float(1e-3), ctypes.c_float(10)))
T("icgn2d1_second_s", lambda: lib.ocb_icgn2d1(ctx, vp(q), len(q), 16, 16, ctypes.c_float(1e-3), ctypes.c_float(10)))
out["first_result_s"] = round(out["device_count_s"] + out["create_s"] + out["set_images_s"] + out["fftcc2d_first_s"] + out["icgn2d1_first_s"], 6)
print(json.dumps(out))
