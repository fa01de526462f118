"""Unit-level checks of the CPU oracle's building blocks against independent NumPy formulations
(reference citations are in oracle/oc_oracle.cpp)."""
import numpy as np
import pytest

from oracle.oracle import Oracle2D, Oracle3D
from opencorr_b200 import make_poi2d, make_poi3d, synth


@pytest.fixture(scope="module")
def pair2d():
    return synth.speckle_pair_2d(160, 128)


@pytest.fixture(scope="module")
def pair3d():
    return synth.speckle_pair_3d(40, 36, 44)


def test_gradient2d_formula_and_zero_borders(pair2d):
    ref, tar = pair2d
    o = Oracle2D(ref, tar)
    o.prepare()
    gx, gy = o.gradients()
    f = ref.astype(np.float64)
    ex = np.zeros_like(f)
    ex[:, 2:-2] = -f[:, 4:] / 12 + f[:, 3:-1] * 2 / 3 - f[:, 1:-3] * 2 / 3 + f[:, :-4] / 12
    ey = np.zeros_like(f)
    ey[2:-2, :] = -f[4:, :] / 12 + f[3:-1, :] * 2 / 3 - f[1:-3, :] * 2 / 3 + f[:-4, :] / 12
    assert np.abs(gx - ex).max() < 1e-4 and np.abs(gy - ey).max() < 1e-4
    assert not gx[:, :2].any() and not gx[:, -2:].any() and not gy[:2].any() and not gy[-2:].any()


def test_bicubic_interpolates_pixels_and_rejects_outside(pair2d):
    ref, tar = pair2d
    h, w = tar.shape
    o = Oracle2D(ref, tar)
    o.prepare()
    ys, xs = np.mgrid[1:h - 2, 1:w - 2]
    v = o.bicubic(np.stack([xs.ravel(), ys.ravel()], 1).astype(np.float32))
    assert np.abs(v - tar[1:h - 2, 1:w - 2].ravel()).max() < 1e-3  # BC row 3 = {0,1,0,0}: interpolating
    out = o.bicubic(np.array([[0.99, 5], [5, 0.5], [w - 2, 5], [5, h - 2], [np.nan, 5]], np.float32))
    assert (out == -1).all()
    # closed form  value = (BC^T ty)^T q (BC^T tx)  
    BC = np.array([[-144, 384, -384, 144], [342, -702, 450, -90], [-198, -18, 270, -54], [0, 336, 0, 0]], np.float64) / 336
    rng = np.random.default_rng(1)
    pts = np.stack([rng.uniform(1, w - 2.001, 200), rng.uniform(1, h - 2.001, 200)], 1).astype(np.float32)
    got = o.bicubic(pts, exact=True)
    for (x, y), g in zip(pts.astype(np.float64), got):
        xi, yi = int(np.floor(x)), int(np.floor(y))
        tx, ty = x - xi, y - yi
        wx = BC.T @ np.array([tx ** 3, tx ** 2, tx, 1])
        wy = BC.T @ np.array([ty ** 3, ty ** 2, ty, 1])
        q = tar[yi - 1:yi + 3, xi - 1:xi + 3].astype(np.float64)
        assert abs(wy @ q @ wx - g) < 2e-4


def test_fftcc2d_is_the_circular_cross_correlation(pair2d):
    ref, tar = pair2d
    o = Oracle2D(ref, tar)
    for r, (x, y) in ((8, (60, 50)), (15, (70, 64)), (10, (100, 40))):
        q = make_poi2d([[x, y]])
        o.fftcc2d(q, r, r, exact=True)
        a = ref[y - r:y + r, x - r:x + r].astype(np.float64)
        b = tar[y - r:y + r, x - r:x + r].astype(np.float64)
        a -= a.mean()
        b -= b.mean()
        c = np.fft.ifft2(np.conj(np.fft.fft2(a)) * np.fft.fft2(b)).real  # c[d] = sum_n a[n] b[n+d]
        idx = int(np.argmax(c))
        dv, du = divmod(idx, 2 * r)
        du = du - 2 * r if du > r else du
        dv = dv - 2 * r if dv > r else dv
        assert (q[0, 2], q[0, 8]) == (du, dv)
        zn = c.max() / np.sqrt((a * a).sum() * (b * b).sum())
        assert abs(q[0, 16] - zn) < 1e-6
        q32 = make_poi2d([[x, y]])
        o.fftcc2d(q32, r, r, exact=False)
        assert (q32[0, 2], q32[0, 8]) == (du, dv) and abs(q32[0, 16] - zn) < 1e-5


def test_fftcc2d_border_guard_leaves_poi_untouched(pair2d):
    ref, tar = pair2d
    o = Oracle2D(ref, tar)
    q = make_poi2d([[7, 60], [60, 7], [ref.shape[1] - 8, 60]])
    q[:, 16] = 0.25
    before = q.copy()
    o.fftcc2d(q, 8, 8)
    assert np.array_equal(q, before)


def test_icgn2d1_recovers_known_translation():
    ref, tar = synth.speckle_pair_2d(200, 200)
    xy = synth.grid_2d(60, 60, 4, 4, 25, 25)
    q = make_poi2d(xy)
    o = Oracle2D(ref, tar)
    o.fftcc2d(q, 16, 16)
    o.icgn2d1(q, 16, 16, 0.001, 10)
    u, v = synth.displacement_2d(xy[:, 0], xy[:, 1], 200, 200)
    assert (q[:, 16] > 0.95).all()
    assert np.abs(q[:, 2] - u).max() < 0.03 and np.abs(q[:, 8] - v).max() < 0.03
    assert (q[:, 23] == 16).all() and (q[:, 24] == 16).all()


def test_icgn2d_sentinel_state_machine(pair2d):
    """-3 guard / skip on zncc<0 / -3 when samples leave the target / -4 at stop."""
    ref, tar = pair2d
    h, w = ref.shape
    o = Oracle2D(ref, tar)
    q = make_poi2d([[5, 60], [80, 64], [80, 64], [80, 64], [80, 64]])
    q[1, 16] = -1.0            # arrives failed: untouched
    q[2, 2] = float(w)         # |u| >= width
    q[3, 2] = 70.0             # pushes the target subset out of the image: -3 inside the loop
    q[4, 2], q[4, 8] = 2.0, -2.0
    before = q.copy()
    o.icgn2d1(q, 10, 10, 1e-6, 1)
    assert list(q[:4, 16]) == [-3, -1, -3, -3]
    assert np.array_equal(np.delete(q[:4], 16, axis=1), np.delete(before[:4], 16, axis=1))
    assert q[4, 16] == -4 and q[4, 17] == 1 and q[4, 18] >= 1e-6
    assert q[4, 14] == 2.0 and q[4, 15] == -2.0


def test_icgn2d2_drops_incoming_second_order_guess():
    ref, tar = synth.speckle_pair_2d(160, 160, second_order=True)
    o = Oracle2D(ref, tar)
    qa = make_poi2d([[80, 80]])
    o.fftcc2d(qa, 12, 12)
    qb = qa.copy()
    qb[0, 5:8] = 0.5  # uxx uxy uyy garbage that ICGN2D2 must ignore (src/oc_icgn.cpp:765-770)
    o.icgn2d2(qa, 12, 12, 0.001, 10)
    o.icgn2d2(qb, 12, 12, 0.001, 10)
    assert np.array_equal(qa, qb)


def test_prefilter_and_tricubic(pair3d):
    ref, tar = pair3d
    o = Oracle3D(ref, tar)
    o.prepare()
    coef = o.coefficients()
    b = np.array([1.732176555412860, -0.464135309171000, 0.124364681271139, -0.033323415913556, 0.008928982383084,
                  -0.002392513618779, 0.000641072092032, -0.000171774749350])
    cur = tar.astype(np.float64)
    for axis in (2, 1, 0):  # x, y, z
        n = cur.shape[axis]
        idx = np.arange(n)
        acc = b[0] * cur
        for t in range(1, 8):
            acc = acc + b[t] * (np.take(cur, np.clip(idx - t, 0, n - 1), axis) + np.take(cur, np.clip(idx + t, 0, n - 1), axis))
        cur = acc
    assert np.abs(coef - cur).max() < 2e-3
    # cubic B-spline evaluation at integer nodes reproduces (1/6, 4/6, 1/6) smoothing of the coefficients
    dz, dy, dx = tar.shape
    pts = np.array([[10, 12, 14], [20, 9, 30]], np.float32)
    got = o.tricubic(pts, exact=True)
    k = np.array([1 / 6, 4 / 6, 1 / 6])
    for (x, y, z), g in zip(pts.astype(int), got):
        blk = coef[z - 1:z + 2, y - 1:y + 2, x - 1:x + 2].astype(np.float64)
        assert abs(np.einsum("i,j,k,ijk->", k, k, k, blk) - g) < 1e-3
    # the prefilter inverts that smoothing: interpolant ~ image at the nodes
    assert abs(got[0] - tar[14, 12, 10]) < 0.5
    assert (o.tricubic(np.array([[0.5, 5, 5], [5, 5, dz - 2.0], [np.nan, 5, 5]], np.float32)) == -1).all()


def test_gradient3d_zero_borders(pair3d):
    ref, tar = pair3d
    o = Oracle3D(ref, tar)
    o.prepare()
    gx, gy, gz = o.gradients()
    f = ref.astype(np.float64)
    ez = np.zeros_like(f)
    ez[2:-2] = -f[4:] / 12 + f[3:-1] * 2 / 3 - f[1:-3] * 2 / 3 + f[:-4] / 12
    assert np.abs(gz - ez).max() < 1e-4
    assert not gx[:, :, :2].any() and not gy[:, -2:, :].any() and not gz[:2].any()


def test_fftcc3d_and_icgn3d1_recover_known_field(pair3d):
    ref, tar = pair3d
    dz, dy, dx = ref.shape
    xyz = np.array([[20, 18, 22], [19, 17, 20]], np.float32)
    o = Oracle3D(ref, tar)
    q = make_poi3d(xyz)
    o.fftcc3d(q, 8, 8, 8)
    u, v, w = synth.displacement_3d(xyz[:, 0], xyz[:, 1], xyz[:, 2], dx, dy, dz)
    assert np.abs(q[:, 3] - u).max() < 1.0 and np.abs(q[:, 7] - v).max() < 1.0 and np.abs(q[:, 11] - w).max() < 1.0  # integer-pixel guess
    assert np.array_equal(q[:, [3, 7, 11]], np.round(q[:, [3, 7, 11]]))
    o.icgn3d1(q, 8, 8, 8, 0.001, 20)
    assert (q[:, 18] > 0.9).all()
    assert np.abs(q[:, 3] - u).max() < 0.1 and np.abs(q[:, 7] - v).max() < 0.1 and np.abs(q[:, 11] - w).max() < 0.1
    # a POI whose window leaves the volume is left untouched by FFT-CC and gets -3 from IC-GN
    q2 = make_poi3d([[4, 18, 22]])
    before = q2.copy()
    o.fftcc3d(q2, 8, 8, 8)
    assert np.array_equal(q2, before)
    o.icgn3d1(q2, 8, 8, 8, 0.001, 20)
    assert q2[0, 18] == -3


def test_icgn2d_ex_reduces_to_plain_overloads():
    """Zero offsets == compute(queue); self-adaptive with a uniform radius == that radius."""
    ref, tar = synth.speckle_pair_2d(220, 200)
    xy = synth.grid_2d(60, 60, 4, 3, 28, 31)
    o = Oracle2D(ref, tar)
    q = make_poi2d(xy)
    o.fftcc2d(q, 12, 12)
    for order, plain in ((1, o.icgn2d1), (2, o.icgn2d2)):
        a, b, c = q.copy(), q.copy(), q.copy()
        plain(a, 12, 12, 0.001, 10)
        o.icgn2d_ex(order, b, 12, 12, 0.001, 10, center_offsets=np.zeros((len(q), 2), np.float32))
        assert np.array_equal(a, b)
        c[:, 23:25] = 12
        o.icgn2d_ex(order, c, 5, 7, 0.001, 10, self_adaptive=True)
        assert np.array_equal(a, c)
    # a centre offset reports the displacement of the offset point: u changes by ~ du/dx * off_x
    d = q.copy()
    off = np.tile(np.array([[3.0, 0.0]], np.float32), (len(q), 1))
    o.icgn2d_ex(1, d, 12, 12, 0.001, 10, center_offsets=off)
    e = q.copy()
    o.icgn2d1(e, 12, 12, 0.001, 10)
    assert abs(float(np.mean(d[:, 2] - e[:, 2])) - 3.0 * 1.5e-3) < 1.5e-3
