"""Development probe (GPU box): cost of cudaHostRegister / cudaHostUnregister on a POI-queue-sized buffer, and of a
pageable vs registered 5 MB round trip."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from opencorr_b200 import _capi
import opencorr_b200 as ob
lib = _capi.load()
eng = ob.Engine(0)
for mb in (1, 5, 50):
    a = np.zeros(mb * 250000, np.float32)
    a[:] = 1.0
    p = ctypes.c_void_p(a.ctypes.data)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); rc1 = lib.ocb_host_register(p, a.nbytes); t1 = time.perf_counter(); rc2 = lib.ocb_host_unregister(p); t2 = time.perf_counter()
        ts.append((t1 - t0, t2 - t1))
    print("%3d MB: register %.3f ms, unregister %.3f ms (min of 5; rc %d %d)" % (mb, 1e3 * min(t[0] for t in ts), 1e3 * min(t[1] for t in ts), rc1, rc2))
