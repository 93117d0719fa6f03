"""CPU: unit checks of the oracle's restated third-party arithmetic (PCL / Eigen / Ceres pieces)."""
import numpy as np


def test_pcl_rng_is_mt19937_seed_12345_halved(oracle):
    """PCL SampleConsensusModel: boost::mt19937(12345) through uniform_int<>(0, INT_MAX) = mt() >> 1."""
    rs = np.random.RandomState(12345)          # legacy seeding = init_genrand(12345), the boost/std mt19937 seeding
    raw = rs._bit_generator.random_raw(700)    # crosses the 624-word regeneration boundary
    got = [oracle.pcl_rnd(k) for k in range(700)]
    assert got == [int(v) >> 1 for v in raw]


def test_eigen3f_matches_numpy_up_to_sign(oracle):
    rng = np.random.default_rng(1)
    for _ in range(300):
        A = rng.normal(size=(60, 3)) * np.array([0.01, 0.3, 0.4])
        R, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        cov = ((A @ R.T).T @ (A @ R.T) / 60).astype(np.float32)
        ev, V = oracle.eigen3f(cov)
        w, Vn = np.linalg.eigh(cov.astype(np.float64))
        assert np.allclose(ev, w, rtol=2e-3, atol=2e-6)
        for c in range(3):
            assert abs(abs(V[:, c] @ Vn[:, c]) - 1) < 2e-3


def _calhist_reference(inten, rate=2.5, nb=100):
    """Literal Python transcription of calHist + get_gray_zone (LidarCornersEst.cpp:224-328) incl. std::map::insert."""
    datas = sorted(float(v) for v in inten)
    mn, mx = datas[0], datas[-1]
    factor = nb / (mx - mn)
    hist = [0] * (nb + 1)
    for v in datas:
        hist[int(np.floor((v - mn) * factor + 0.5))] += 1      # round(): values are >= 0
    mean = sum(datas) / len(datas)
    width = (mx - mn) / nb
    m = {}
    for i in range(nb):
        m.setdefault(float(hist[i]), i)                        # insert keeps the first bin per count
    low = high = None
    for cnt in sorted(m, reverse=True):
        edge = width * m[cnt] + mn
        if edge > mean and high is None:
            high = edge
        if edge < mean and low is None:
            low = edge
        if low is not None and high is not None:
            break
    return mn, mx, mean, low, high, ((rate - 1) * low + high) / rate, (low + (rate - 1) * high) / rate


def test_gray_zone_rule(oracle):
    rng = np.random.default_rng(3)
    for trial in range(40):
        n = int(rng.integers(300, 3000))
        black = rng.normal(20, rng.uniform(2, 8), n // 2)
        white = rng.normal(70, rng.uniform(2, 8), n - n // 2)
        inten = np.rint(np.clip(np.concatenate([black, white]), 0, 255)).astype(np.float32)
        rng.shuffle(inten)
        st, g = oracle.gray_zone(inten)
        ref = _calhist_reference(inten)
        assert st == 0
        got = (g["min"], g["max"], g["mean"], g["r_low"], g["r_high"], g["gray"][0], g["gray"][1])
        assert np.allclose(got, ref, rtol=0, atol=1e-12)


def test_gray_zone_degenerate_cases(oracle):
    assert oracle.gray_zone(np.full(100, 7.0, np.float32))[0] == 4          # max == min: division by zero (:235)
    st, _ = oracle.gray_zone(np.array([0.0] * 99 + [1.0], np.float32))      # no populated bin edge above the mean
    assert st in (0, 4)


def _residual_reference(y, z, white, x, W, H, g, tlw, oob):
    """VirtualboardError::operator() (Optimization.h:31-107) transcribed literally, T = double."""
    th, ty, tz = x
    yy = y * np.cos(th) - z * np.sin(th) + ty
    zz = z * np.cos(th) + y * np.sin(th) + tz
    i = (yy + W * g / 2.0) / g
    j = (zz + H * g / 2.0) / g
    if 0 < i < W and 0 < j < H:
        ifl, jfl = np.floor(i), np.floor(j)
        ii, jj = np.floor(ifl / 2) * 2, np.floor(jfl / 2) * 2
        White = not tlw
        if ifl == ii and jfl == jj:
            White = tlw
        if ifl != ii and jfl != jj:
            White = tlw
        if bool(white) == White:
            return 0.0
        ie = np.ceil(i) - i if i - ifl > 0.5 else i - ifl
        je = np.ceil(j) - j if j - jfl > 0.5 else j - jfl
        return ie + je
    if oob:
        return min(abs(i), abs(i - W)) + min(abs(j), abs(j - H))
    return 0.0


def test_residual_and_jacobian(oracle):
    rng = np.random.default_rng(5)
    W, H, g = 6, 8, 0.15
    checked = 0
    for _ in range(4000):
        y, z = rng.uniform(-0.8, 0.8), rng.uniform(-0.9, 0.9)
        x = np.array([rng.uniform(-0.3, 0.3), rng.uniform(-0.1, 0.1), rng.uniform(-0.1, 0.1)])
        white, tlw, oob = int(rng.integers(2)), bool(rng.integers(2)), bool(rng.integers(2))
        r, jac = oracle.residual(y, z, white, x, W, H, g, tlw, oob)
        assert abs(r - _residual_reference(y, z, white, x, W, H, g, tlw, oob)) < 1e-12
        if r > 1e-3:
            # autodiff Jacobian == central differences away from the kinks of the piecewise-linear residual
            h = 1e-7
            num = np.zeros(3)
            smooth = True
            for k in range(3):
                xp, xm = x.copy(), x.copy()
                xp[k] += h
                xm[k] -= h
                rp = oracle.residual(y, z, white, xp, W, H, g, tlw, oob)[0]
                rm = oracle.residual(y, z, white, xm, W, H, g, tlw, oob)[0]
                if rp == 0 or rm == 0 or abs(rp - rm) > 1e-3:
                    smooth = False
                num[k] = (rp - rm) / (2 * h)
            if smooth:
                assert np.allclose(jac, num, atol=2e-5), (jac, num)
                checked += 1
    assert checked > 300


def test_tr_boundary_2d_is_the_boundary_minimiser(oracle):
    rng = np.random.default_rng(7)
    ang = np.linspace(0, 2 * np.pi, 200001)
    for _ in range(200):
        L = rng.normal(size=(2, 2))
        B = L @ L.T + 1e-3 * np.eye(2)
        g = rng.normal(size=2) * 3
        r = float(np.linalg.norm(np.linalg.solve(B, g)) * rng.uniform(0.05, 0.9))   # GN step lies outside
        ok, x = oracle.tr_boundary_2d([B[0, 0], B[0, 1], B[1, 1]], g, r)
        assert ok and abs(np.linalg.norm(x) - r) < 1e-9 * r
        f = 0.5 * x @ B @ x + g @ x
        pts = r * np.stack([np.cos(ang), np.sin(ang)])
        fscan = 0.5 * np.einsum("in,ij,jn->n", pts, B, pts) + g @ pts
        assert f <= fscan.min() + 1e-7 * (1 + abs(fscan.min()))


def test_solver_decreases_cost_and_warm_start_is_a_fixed_point(oracle):
    from lidar_camera_calibration_b200 import synth
    rng = np.random.default_rng(11)
    # a noiseless virtual board sampled on a grid, displaced by a known pose
    W, H, g = 6, 8, 0.15
    yy, zz = np.meshgrid(np.linspace(-0.44, 0.44, 45), np.linspace(-0.59, 0.59, 60), indexing="ij")
    y, z = yy.ravel(), zz.ravel()
    i, j = np.floor((y + W * g / 2) / g), np.floor((z + H * g / 2) / g)
    white = ((i + j) % 2 == 1).astype(np.uint8)          # topleftWhite = False colouring
    th, ty, tz = 0.03, 0.02, -0.015
    # points observed in a frame rotated/translated by the inverse pose
    c, s = np.cos(-th), np.sin(-th)
    yo, zo = (y - ty) * c - (z - tz) * s, (z - tz) * c + (y - ty) * s
    yz = np.stack([yo, zo], 1)
    term, x, cost, it, ne = oracle.solve(yz, white, [0, 0, 0], W, H, g, False, True)
    term2, x2, cost2, it2, ne2 = oracle.solve(yz, white, x, W, H, g, False, False)
    assert cost2 <= cost + 1e-12
    # the zero-cost set of a noiseless board is a small region, not a point: any pose in it is a minimiser
    _, _, cost0, _, _ = oracle.solve(yz, white, [0, 0, 0], W, H, g, False, False, max_iter=0)
    assert abs(x2[0] - th) < 1e-2 and abs(x2[1] - ty) < 6e-3 and abs(x2[2] - tz) < 6e-3 and cost2 < 0.2 * cost0
    assert it <= 50 and it2 <= 50


def test_reciprocal_correction_equals_true_division():
    """The kernels' cost-only pass takes x / g (Optimization.h:45-46 with T = double) as q = x * (1/g); r = fma(-q, g, x);
    q + r * (1/g) -- exact IEEE division for the correctly rounded reciprocal.  Checked here with exact fma emulation
    (math.fma needs 3.13; fractions are exact) on values around the cell borders the floor() of the residual sees."""
    from fractions import Fraction
    rng = np.random.default_rng(7)
    g = 0.15
    inv_g = 1.0 / g

    def fma(a, b, c):  # correctly rounded a*b+c
        return float(Fraction(a) * Fraction(b) + Fraction(c))

    xs = np.concatenate([rng.uniform(-0.2, 1.4, 4000), np.arange(0, 9) * 0.15, np.nextafter(np.arange(0, 9) * 0.15, 10),
                         np.nextafter(np.arange(0, 9) * 0.15, -10), rng.uniform(-1e-12, 1e-12, 200)])
    for x in xs:
        x = float(x)
        q = x * inv_g
        r = fma(-q, g, x)
        assert fma(r, inv_g, q) == x / g, x


def test_restatement_switches_move_the_fixture_result_by_microns(oracle, bag_clouds, fixtures):
    """profiles/deviation_table_r2.md in one assertion: the arithmetic choices of the restatement (reciprocal vs true
    division, rotation axis, PCL's sequential fp32 moments, refit solver) move the corners of bag 1 by < 5e-6 m."""
    fx = np.asarray(fixtures["lidar_corners"]["1"])
    L = oracle.lib()
    base = oracle.extract_frame(bag_clouds[1], fx.mean(0), flags=1)[1]
    try:
        for mask in (1, 2, 4, 8):
            L.ilcc2o_set_variant(mask, 0, 0)
            c = oracle.extract_frame(bag_clouds[1], fx.mean(0), flags=1)[1]
            assert np.abs(c - base).max() < 5e-6, mask
    finally:
        L.ilcc2o_set_variant(0, 0, 0)
