"""The oracle against the reference's own outputs (tests/golden, made by oracle/make_golden.py
from the unmodified reference) and against independent statements of the same maths."""
import numpy as np
import pytest

from oracle import oracle as O


def test_overlap_fft_restatement_matches_reference_goldens(overlap_golden, manifest):
    g = overlap_golden
    for case in manifest["overlap"]:
        i, alpha = case["idx"], case["alpha"]
        a, b, dr = g[f"a{i}"], g[f"b{i}"], g[f"dr{i}"]
        got = O.density_overlap_fft(a ** alpha, b ** alpha, dr) if alpha != 1.0 else O.density_overlap_fft(a, b, dr)
        ref = g[f"E{i}"]                                   # same numpy calls; pocketfft's SIMD path
        assert np.abs(got - ref).max() <= 1e-14 * np.abs(ref).max(), case   # may differ in the last bit


def test_direct_sum_equals_fft_expression(overlap_golden, manifest):
    g = overlap_golden
    for case in manifest["overlap"]:
        if np.prod(case["shape"]) > 4000:
            continue
        i, alpha = case["idx"], case["alpha"]
        a, b, dr = g[f"a{i}"] ** alpha, g[f"b{i}"] ** alpha, g[f"dr{i}"]
        ref = g[f"E{i}"]
        got = O.corr_direct(a, b, dr)
        assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()
        z_lo, z_hi = 3, 9
        assert np.array_equal(O.corr_direct(a, b, dr, z_lo, z_hi), got[:, :, z_lo:z_hi])


def test_tricubic_lekien_marsden_equals_catmull_rom_and_hermite_solve():
    rng = np.random.default_rng(7)
    d = rng.standard_normal((7, 6, 5))
    tc = O.Tricubic(list(d), list(d.shape))
    for _ in range(60):
        p = rng.uniform(-13, 16, 3)
        a = tc.ip(list(p))
        assert abs(a - O.tricubic_lm_python(d, p)) < 5e-13
        assert abs(a - tc.ip_catmull(p)) < 5e-13
    # interpolates the nodes, periodic in all three axes
    assert tc.ip([2, 3, 4]) == d[2, 3, 4]
    assert tc.ip([-1, 6, 5]) == d[6, 0, 0]
    assert abs(tc.ip([2.25, 3.5, -0.75]) - tc.ip([2.25 + 7, 3.5 - 12, -0.75 + 10])) < 1e-13


def test_c_powell_is_bit_identical_to_installed_scipy():
    from scipy.optimize import minimize

    def rosen(x):
        return (1 - x[0]) ** 2 + 100 * (x[1] - x[0] ** 2) ** 2

    def bowl(x):
        return np.cos(x[0]) + 0.1 * (x[0] - 3.0) ** 2 + 0.5 * (x[1] - 0.3) ** 2 + 0.05 * np.sin(3 * x[1]) * x[0]

    def flat_phi(x):
        return 0.1 * (np.pi - x[0]) ** 2 + 1e-3 * np.sin(x[0]) * np.cos(x[1])

    for f, x0 in [(rosen, [-1.2, 1.0]), (bowl, [0.0, 0.0]), (bowl, [np.pi, 0.0]), (flat_phi, [np.pi, 0.0])]:
        ref = minimize(f, x0, method="Powell")
        x, fun, nit, nfev = O.powell_minimize(f, x0)
        assert nfev == ref.nfev and nit == ref.nit
        assert np.array_equal(x, ref.x) and fun == ref.fun
    # maxfev exhaustion takes scipy's exception path
    ref = minimize(rosen, [-1.2, 1.0], method="Powell", options={"maxfev": 25, "maxiter": 2000})
    x, fun, nit, nfev = O.powell_minimize(rosen, [-1.2, 1.0], maxfun=25)
    assert nfev == ref.nfev and np.array_equal(x, ref.x) and fun == ref.fun


def test_relax_restatements_match_reference_goldens(relax_golden, manifest):
    g = relax_golden
    static, dr, cols = g["static"], g["dr"], g["cols"]
    m = manifest["relax"]
    for c, (i, j) in enumerate(cols):
        got = O.relax_tile(static, i, i + 1, j, j + 1, m["nz"], m["kappa"], m["rpivot"], dr)[0, 0]
        assert np.array_equal(got, g["ortho"][c]), (i, j)     # C restatement == reference + scipy, bit for bit
    for c, (i, j) in enumerate(cols[:3]):
        got = O.relax_tile(static, i, i + 1, j, j + 1, m["nz"], m["kappa"], m["rpivot"], dr, dR=np.diag(dr))[0, 0]
        # np.linalg.inv (LAPACK) vs closed-form 3x3 inverse: last-bit differences in the offset
        assert np.abs(got[:, 0] - g["noortho"][c][:, 0]).max() < 1e-12
        assert np.abs(got[:, 1] - g["noortho"][c][:, 1]).max() < 1e-6
        s = m["stiff"]
        got = O.relax_tile(static, i, i + 1, j, j + 1, s["nz"], s["kappa"], s["rpivot"], dr)[0, 0]
        assert np.array_equal(got, g["stiff"][c])
    # the Python restatement on the installed scipy (what the reference really executes)
    tc = O.Tricubic(static)
    i, j = cols[2]
    py = O.relax_tip_z_scipy(i, j, tc, m["kappa"], m["rpivot"] / dr[2], np.arange(m["nz"]), dr)
    assert np.array_equal(py, g["ortho"][2])


def test_pivot_height_truncation_quirk(relax_golden):
    """origin is int64 in the reference: pivot plane = trunc(k + npivot), not k + npivot."""
    g = relax_golden
    static, dr = g["static"], g["dr"]
    out = O.relax_tile(static, 4, 5, 4, 5, 3, 0.2, 3.02, dr)[0, 0]
    tc = O.Tricubic(static)
    th, az = out[0, 1], out[0, 2]
    r = (3.02 / dr[2]) * dr[2]
    z_trunc = int(0 + 3.02 / dr[2]) + r * np.cos(th) / dr[2]
    e = tc.ip([4 + r * np.sin(th) * np.cos(az) / dr[0], 4 + r * np.sin(th) * np.sin(az) / dr[1], z_trunc])
    assert abs(e + 0.5 * 0.2 * (np.pi - th) ** 2 - out[0, 0]) < 1e-14


def test_get_fs_restatement_matches_reference_golden():
    """tools.py:5-32 restated with numpy (the checker used by the GPU Delta-f test)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "nextrow_reference.npz"))
    F = g["F"]
    for key, (dz, k0, f0, n) in {"fs": (0.075, 1800, 30300, 10), "fs6": (0.1, 1500.0, 25000.0, 6)}.items():
        x = np.linspace(-1, 1, n + 1)
        y = np.sqrt(1 - x * x)
        dy = (y[1:] - y[:-1]) / (dz * n)
        pref = (1 + (n - 2) ** 2 * (2 / np.pi)) / ((n - 2) ** 2 + 1)
        got = -pref * np.apply_along_axis(lambda m: np.convolve(m, dy, mode="valid"), 2, F) * 16.0217656 * f0 / k0
        assert np.array_equal(got, g[key])
