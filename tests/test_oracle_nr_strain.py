"""CPU: the oracle's NR2D1 and Strain restatements pinned to the result tables the reference ships
(examples/2d_dic/oht_cfrp_4_fftcc_nr1_r16.csv, examples/dvc/Torus_def_sift_icgn1_r16.csv; sub-sampled
into tests/golden by make_golden.py)."""
import numpy as np
import pytest

from oracle import oracle
from oracle.oracle import Oracle2D
from opencorr_b200 import make_poi2d
import util


@pytest.mark.parametrize("exact", [0, 1])
def test_nr2d1_golden_table(exact):
    ref, tar = util.oht_cfrp_pair()
    tab = util.oht_cfrp_nr_golden()["table"]
    q = make_poi2d(tab[:, 0:2])
    o = Oracle2D(ref, tar)
    o.fftcc2d(q, 16, 16, exact=exact)
    o.nr2d1(q, 16, 16, 0.001, 10, exact=exact)
    assert np.array_equal(q[:, 14:16], tab[:, 4:6])  # FFT-CC guess kept in u0, v0
    # the shipped table predates the -4 code: rows that hit the iteration limit keep a ZNCC there
    conv = tab[:, 7] < 10
    same_it = q[:, 17] == tab[:, 7]
    assert (same_it | ~conv).mean() > 0.995
    ok = conv & same_it
    assert ok.sum() > 0.9 * len(tab)
    # NR converges linearly on the poorly correlated POIs around the specimen's hole: rounding differences are
    # amplified there, so the tight bound is on the well correlated POIs
    good = ok & (tab[:, 6] >= 0.9)
    d = np.abs(q[:, [2, 8]] - tab[:, [2, 3]]).max(1)
    assert d[good].max() < 5e-5, d[good].max()
    assert d[ok].max() < 1e-3, d[ok].max()
    assert np.abs(q[good, 16] - tab[good, 6]).max() < 2e-6
    assert np.abs(q[good, 18] - tab[good, 8]).max() < 1e-4
    # current source: not converged -> -4
    nc = q[:, 17] >= 10
    assert np.all(q[nc & (q[:, 18] >= 0.001), 16] == -4)


def test_nr2d1_sentinels():
    ref, tar = util.oht_cfrp_pair()
    h, w = ref.shape
    q = make_poi2d(np.array([[5, 5], [w - 3, 100], [100, 100], [120, 120], [140, 140]], np.float32))
    q[3, 16] = -2.0      # incoming negative ZNCC below -1 is kept (src/oc_nr.cpp:170)
    q[4, 2] = np.nan     # NaN guess
    q[4, 14] = 1.5
    Oracle2D(ref, tar).nr2d1(q, 16, 16, 0.001, 10)
    assert q[0, 16] == -1 and q[1, 16] == -1       # border guard writes -1, not -3
    assert q[2, 16] > 0.9 or q[2, 16] == -4
    assert q[3, 16] == -2
    assert q[4, 16] == -5 and q[4, 2] == 1.5       # NaN -> -5, u restored from u0 (:319-324)


@pytest.mark.parametrize("exact", [0, 1])
def test_strain2d_golden_band(exact):
    q, gold, check = util.strain_band_queue()
    oracle.strain(q, 20.0, 5, 0.9, 1, exact=exact)
    good = check & (q[:, 16] >= 0.9)
    assert good.sum() > 6000
    d = np.abs(q[good, 20:23] - gold[good]).max()
    assert d < (2e-7 if exact else 6e-7), d
    # the shipped table was written before Strain::compute(queue) started skipping POIs below the ZNCC threshold
    # (src/oc_strain.cpp:244-248): the current source leaves those untouched
    bad = q[:, 16] < 0.9
    assert bad.sum() > 500 and np.all(q[bad, 20:23] == 0)


@pytest.mark.parametrize("exact", [0, 1])
def test_strain3d_golden_crop(exact):
    q, gold, check = util.torus_queue()
    oracle.strain(q, 30.0, 5, 0.9, 1, exact=exact)
    good = check & (q[:, 18] >= 0.9)
    assert good.sum() > 1000
    d = np.abs(q[good, 22:28] - gold[good]).max()
    assert d < (2e-6 if exact else 5e-6), d


def test_strain_knn_fallback_green_and_threshold():
    rng = np.random.default_rng(5)
    n = 400
    xy = rng.uniform(0, 1000, (n, 2)).astype(np.float32)   # sparse: most POIs have < 5 neighbours within 20 px
    q = make_poi2d(xy)
    q[:, 2] = 0.01 * xy[:, 0] + 0.002 * xy[:, 1]           # u = 0.01 x + 0.002 y
    q[:, 8] = -0.003 * xy[:, 0] + 0.02 * xy[:, 1]
    q[:, 16] = 0.95
    q[::7, 16] = 0.5
    a = q.copy()
    oracle.strain(a, 20.0, 5, 0.9, 1)
    done = np.any(a[:, 20:23] != 0, axis=1)
    assert done.sum() > 50 and not done[::7].any()
    assert np.abs(a[done, 20] - 0.01).max() < 1e-5 and np.abs(a[done, 21] - 0.02).max() < 1e-5
    assert np.abs(a[done, 22] - 0.5 * (0.002 - 0.003)).max() < 1e-5
    b = q.copy()
    oracle.strain(b, 20.0, 5, 0.9, 2)                       # Green strain (src/oc_strain.cpp:229-235)
    ux, uy, vx, vy = 0.01, 0.002, -0.003, 0.02
    assert np.abs(b[done, 20] - (ux + 0.5 * (ux * ux + vx * vx))).max() < 1e-5
    assert np.abs(b[done, 22] - 0.5 * (uy + vx + uy * ux + vy * vx)).max() < 1e-5


@pytest.mark.parametrize("legacy", [1, 0])
def test_epipolar_search_golden_crop(legacy):
    """EpipolarSearch -> ICGN2D2 (reference examples/test_3d_reconstruction_epipolar.cpp) vs the shipped
    'Step18 00,00-0005_1_reconstruction_epipolar.csv' on the 60-POI crop.  legacy=1 drops the '-4 = not converged' rule
    the table predates (on the full grid ~26 % of the POIs need it to reproduce the table, see oracle/oc_oracle.cpp);
    in this block the current source finds the same matches."""
    v1, v2, fm, tab = util.step18_epipolar_fixture()
    p = util.STEP18_EPIPOLAR
    oracle.set_legacy_no_minus4(legacy)
    try:
        q = make_poi2d(tab[:, 0:2])
        o = Oracle2D(v1, v2)
        o.epipolar_search(q, fm, p["parallax_x"], p["parallax_y"], p["search_radius"], p["search_step"], p["rx"], p["ry"], p["conv"], p["stop"])
        assert np.all(q[:, 16] > 0.9)
        assert np.all(q[:, 14] == np.round(q[:, 14])) and np.all(q[:, 15] == np.round(q[:, 15]))  # winning candidate: integer offsets
        o.icgn2d2(q, 9, 9, 0.001, 10)
    finally:
        oracle.set_legacy_no_minus4(0)
    assert np.abs(q[:, 0] + q[:, 2] - tab[:, 3]).max() < 2e-4   # r2_x (8 significant digits printed at ~1100 px)
    assert np.abs(q[:, 1] + q[:, 8] - tab[:, 4]).max() < 2e-4
    conv = q[:, 16] != -4    # current source: a refinement that hits the iteration limit loses its ZNCC to the -4 code
    assert conv.mean() > 0.9 and (legacy == 0 or conv.all())
    assert np.abs(q[conv, 16] - tab[conv, 2]).max() < 2e-6


@pytest.mark.parametrize("exact", [0, 1])
def test_strain_stereo_golden_crop(exact):
    """Strain on POI2DS records (stereo DIC, reference src/oc_strain.cpp:252-371) vs the strains shipped in
    examples/3d_dic/GT4-0273_0_epipolar_sift_r16.csv.  The fit is poorly conditioned there (~20 neighbours on a nearly
    flat patch, 3D coordinates of ~400 mm): the reference's own float32 QR carries ~1e-4 of noise, which bounds the
    agreement of either flavour with the table."""
    q, gold, check = util.gt4_stereo_queue()
    oracle.strain(q, 20.0, 5, 0.9, 1, exact=exact)
    good = check & np.all(q[:, 5:8] >= 0.9, axis=1)
    assert good.sum() > 1500
    d = np.abs(q[good, 20:26] - gold[good]).max(1)
    assert np.median(d) < (2e-5 if exact else 6e-5) and d.max() < 1e-3, (np.median(d), d.max())
    bad = ~np.all(q[:, 5:8] >= 0.9, axis=1)
    assert np.all(q[bad, 20:26] == 0)
