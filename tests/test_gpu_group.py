"""GROUP contexts (one process, several devices: ocb_create(-1) / ocb_create_multi): host-queue calls shard the caller's array
over the member devices inside the C ABI and must give records bit-identical to a single-device context -- what
FFTCC2D::compute(queue) / ICGN2D1::compute(queue) (reference src/oc_fftcc.cpp:277-285, src/oc_icgn.cpp:343-351) do with OpenMP
threads.  On a one-GPU box the group has a single member (the plumbing still runs); with more GPUs every device takes a share."""
import ctypes

import numpy as np
import pytest

import opencorr_b200 as ob
from opencorr_b200 import _capi, synth

pytestmark = pytest.mark.gpu


def _devices():
    return list(range(_capi.load().ocb_device_count()))


def test_group_2d_matches_single_device(engine):
    ref, tar = synth.speckle_pair_2d(1536, 1024)
    xy = synth.grid_2d(40, 40, 208, 135, 7, 7)  # 28 080 POIs: enough for up to three devices (8192 per device minimum)
    grp = ob.Engine(_devices())
    assert grp.member_count == len(_devices())
    res = []
    for eng in (engine, grp):
        q = ob.make_poi2d(xy)
        eng.set_images_2d(ref, tar)
        eng.fftcc2d(q, 16, 16)
        eng.icgn2d_prepare()
        eng.icgn2d1(q, 16, 16, 0.001, 10)
        q2 = q.copy()
        q2[:, 16] = 0
        eng.icgn2d2(q2, 16, 16, 0.001, 10)
        res.append((q, q2))
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    assert (res[0][0][:, 16] > 0.9).all()
    # offsets and the self-adaptive mode are sliced together with the queue
    off = np.random.default_rng(3).uniform(-3, 3, (len(xy), 2)).astype(np.float32)
    out = []
    for eng in (engine, grp):
        q = res[0][0].copy()
        q[:, 16] = 0
        q[:, 23:25] = np.array([[12, 12], [16, 16], [10, 18]], np.float32)[np.arange(len(q)) % 3]
        eng.icgn2d_ex(1, q, 16, 16, 0.001, 10, center_offsets=off, self_adaptive=True)
        out.append(q)
    # self-adaptive radii make one launch per radius group; a group too small to fill a device runs with two warps per POI, which
    # splits the sums differently: codes and iteration counts identical, values equal to the last few bits
    assert np.array_equal(out[0][:, 16] < 0, out[1][:, 16] < 0) and np.array_equal(out[0][:, 17], out[1][:, 17])
    assert np.abs(out[0] - out[1]).max() < 2e-5
    # the device-pointer entry points are refused on a group
    with pytest.raises(ob.OpenCorrB200Error):
        grp.icgn2d1_dev(0, 0, 16, 16, 0.001, 10)
    grp.close()


def test_group_3d_and_u8_match_single_device(engine):
    ref, tar = synth.speckle_pair_3d(96, 88, 80)
    xyz = synth.grid_3d(20, 20, 20, 9, 8, 7, 6, 6, 6)  # 504 POIs
    grp = ob.Engine(_devices())
    res = []
    for eng, cast in ((engine, np.float32), (grp, np.uint8)):
        q = ob.make_poi3d(xyz)
        eng.set_images_3d(ref.astype(cast), tar.astype(cast))
        eng.fftcc3d(q, 8, 8, 8)
        eng.icgn3d_prepare()
        eng.icgn3d1(q, 8, 8, 8, 0.001, 20)
        res.append(q)
    assert np.array_equal(res[0], res[1])
    assert (res[0][:, 18] > 0.9).mean() > 0.99
    grp.close()


def test_create_all_and_members():
    lib = _capi.load()
    n = lib.ocb_device_count()
    ctx = lib.ocb_create(-1)
    assert ctx
    assert lib.ocb_member_count(ctx) == n
    assert lib.ocb_member(ctx, 0)
    assert not lib.ocb_member(ctx, n)
    lib.ocb_destroy(ctx)
    two = (ctypes.c_int * 2)(0, 0)
    assert not lib.ocb_create_multi(two, 2)  # a device listed twice
    assert b"twice" in lib.ocb_last_error(None)


def test_host_register_round_trip(engine):
    """A page-locked queue (ocb_host_register) is not copied: the kernels read and write the caller's records in place through
    PCIe, and FFT-CC chunks start as soon as the image rows they need have arrived (banded upload).  Same records as with a
    pageable queue, for a queue long enough for the chunked path and for a short one."""
    lib = _capi.load()
    ref, tar = synth.speckle_pair_2d(1280, 1024)
    for nx, ny in ((170, 120), (40, 30)):  # 20 400 POIs (chunked, banded) and 1 200 POIs (one launch)
        xy = synth.grid_2d(40, 40, nx, ny, 7, 7)
        out = []
        for pin in (False, True):
            q = ob.make_poi2d(xy)
            if pin:
                assert lib.ocb_host_register(ctypes.c_void_p(q.ctypes.data), q.nbytes) == 0
            try:
                engine.set_images_2d(ref, tar)
                engine.fftcc2d(q, 16, 16)
                engine.icgn2d_prepare()
                engine.icgn2d1(q, 16, 16, 0.001, 10)
                q2 = q.copy()
                q2[:, 16] = 0
                if pin:
                    assert lib.ocb_host_register(ctypes.c_void_p(q2.ctypes.data), q2.nbytes) == 0
                try:
                    engine.icgn2d2(q2, 16, 16, 0.001, 10)
                finally:
                    if pin:
                        assert lib.ocb_host_unregister(ctypes.c_void_p(q2.ctypes.data)) == 0
            finally:
                if pin:
                    assert lib.ocb_host_unregister(ctypes.c_void_p(q.ctypes.data)) == 0
            out.append((q.copy(), q2.copy()))
        assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
        assert (out[0][0][:, 16] > 0.9).all()
