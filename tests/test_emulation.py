"""The CUDA kernel bodies (dbspm_b200/csrc/*.h), compiled for the CPU by tests/emu, against
the oracle.  This checks the code that ships -- plan, butterflies, tiling, the packed-real
untangle / product / pruned inverse algebra, the Powell state machine -- without a GPU."""
import numpy as np
import pytest

from oracle import oracle as O

LENGTHS = [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 14, 15, 16, 20, 21, 27, 28, 30, 32, 49, 54, 64, 72, 77, 80, 84, 90, 96, 98,
           100, 108, 112, 120, 128, 160, 196, 240, 256, 320, 384, 512, 576, 1024, 11, 13, 22, 26, 33, 62, 169]


def _factors(n):
    out, p = [], 2
    while n > 1:
        while n % p == 0:
            out.append(p)
            n //= p
        p += 1
    return out


@pytest.mark.parametrize("n", LENGTHS)
def test_tile_fft_matches_numpy(emu, n):
    rng = np.random.default_rng(n)
    for tshift in (0, 2, 3):
        x = rng.standard_normal((n, 1 << tshift)) + 1j * rng.standard_normal((n, 1 << tshift))
        fwd = emu.fft_lines(x, -1, tshift)
        inv = emu.fft_lines(x, +1, tshift)
        scale = max(1.0, np.abs(np.fft.fft(x, axis=0)).max())
        assert np.abs(fwd - np.fft.fft(x, axis=0)).max() < 2e-14 * scale * max(1, np.log2(n + 1))
        assert np.abs(inv - np.fft.ifft(x, axis=0) * n).max() < 2e-14 * scale * max(1, np.log2(n + 1))
        if all(f in (2, 3, 5, 7) for f in _factors(n)) and n > 1:
            # transposed (decimation-in-time) stages: permuted input -> natural output
            fwd_t = emu.fft_lines(x, -1, tshift, transposed=True)
            inv_t = emu.fft_lines(x, +1, tshift, transposed=True)
            assert np.abs(fwd_t - np.fft.fft(x, axis=0)).max() < 2e-14 * scale * max(1, np.log2(n + 1))
            assert np.abs(inv_t - np.fft.ifft(x, axis=0) * n).max() < 2e-14 * scale * max(1, np.log2(n + 1))


def test_unsupported_length_is_reported(emu):
    with pytest.raises(RuntimeError):
        emu.fft_lines(np.zeros((37, 1), dtype=complex), -1, 0)       # prime > 31


def test_corr3d_matches_reference_goldens(emu, overlap_golden, manifest):
    g = overlap_golden
    for case in manifest["overlap"]:
        i, alpha = case["idx"], case["alpha"]
        a, b, dr, ref = g[f"a{i}"], g[f"b{i}"], g[f"dr{i}"], g[f"E{i}"]
        nz = a.shape[2]
        for z_lo, z_hi in [(0, nz), (nz // 4, nz // 4 + nz // 2), (3, 8), (nz - 1, nz)]:
            got = emu.corr3d(a, b, dr, z_lo, z_hi, alpha, alpha)
            err = np.abs(got - ref[:, :, z_lo:z_hi]).max() / np.abs(ref).max()
            assert err <= 1e-12, (case, z_lo, z_hi, err)


@pytest.mark.parametrize("shape,window", [
    ((4, 4, 4), (0, 4)), ((2, 2, 2), (0, 2)), ((1, 1, 2), (0, 2)), ((3, 1, 6), (1, 6)), ((1, 5, 4), (2, 3)),
    ((6, 5, 8), (3, 7)), ((7, 9, 10), (1, 2)), ((12, 10, 28), (0, 1)), ((14, 15, 16), (5, 12)),
    ((5, 14, 12), (11, 12)), ((16, 3, 6), (0, 5)),
])
def test_corr3d_edge_shapes_all_tile_widths(emu, shape, window):
    rng = np.random.default_rng(sum(shape))
    a, b = rng.random(shape), rng.random(shape) - 0.4
    dr = np.array([0.1, 0.2, 0.3])
    ref = O.density_overlap_fft(a, b, dr)[:, :, window[0]:window[1]]
    # -1/0/1/2: register-first/last body with auto/forced tile width; 100+: the all-in-smem body;
    # 200+: the TMA-fed persistent body (loads/stores emulated by copies); 300+: the persistent software-pipelined body
    for tshift in (-1, 0, 1, 2, 100, 101, 103, 200, 201, 203, 300, 301, 303):
        got = emu.corr3d(a, b, dr, *window, tshift=tshift)
        assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max(), (shape, window, tshift)


@pytest.mark.parametrize("shape,window", [((64, 128, 8), (2, 7)), ((128, 64, 6), (0, 6)), ((512, 3, 4), (1, 3)), ((1024, 2, 4), (0, 2)), ((20, 96, 10), (3, 4)),
                                          ((128, 5, 6), (1, 4)), ((64, 7, 10), (0, 9))])       # the last two: ragged last tiles (inner = 15, 35)
def test_pipelined_axis_pass_is_bit_identical_to_the_one_shot_pass(emu, shape, window):
    """axis_pass_body_v2s (persistent blocks, next tile's rows requested before the remaining stages of the current
    one) does the arithmetic of axis_pass_body_v2 in the same order: same bits, with and without the fused power."""
    rng = np.random.default_rng(sum(shape))
    a, b = rng.random(shape) + 0.01, rng.random(shape) + 0.02
    dr = np.array([0.1, 0.2, 0.3])
    for alpha in (1.0, 1.08):
        ref = O.density_overlap_fft(a ** alpha, b ** alpha, dr)[:, :, window[0]:window[1]]
        for one_shot, piped in ((-1, 300), (0, 301), (2, 303)):
            x = emu.corr3d(a, b, dr, *window, alpha0=alpha, alpha1=alpha, tshift=one_shot)
            y = emu.corr3d(a, b, dr, *window, alpha0=alpha, alpha1=alpha, tshift=piped)
            assert np.array_equal(x, y), (shape, alpha, piped)
            assert np.abs(y - ref).max() <= 1e-12 * np.abs(ref).max()


@pytest.mark.parametrize("shape,window,G", [
    ((4, 4, 4), (0, 4), 1), ((4, 4, 4), (1, 3), 2), ((6, 8, 8), (3, 7), 2), ((8, 16, 12), (2, 9), 4),
    ((12, 24, 10), (0, 5), 3), ((16, 16, 6), (5, 6), 8), ((10, 20, 28), (7, 21), 5), ((2, 2, 2), (0, 2), 1),
    ((14, 28, 16), (0, 16), 2), ((9, 12, 8), (1, 8), 3),
])
def test_corr3d_distributed_x_slabs(emu, shape, window, G):
    """x-slab decomposition over G ranks (y pass -> all-to-all in slot order -> x pass, z kernel, x inverse ->
    gather -> y inverse through the slot table) equals the reference's periodic correlation."""
    rng = np.random.default_rng(sum(shape) + G)
    a, b = rng.random(shape), rng.random(shape) - 0.4
    dr = np.array([0.1, 0.2, 0.3])
    ref = O.density_overlap_fft(a ** 1.08, b, dr)[:, :, window[0]:window[1]]
    # 1000+: fused exchange (the y pass stores every row straight into its owner's receive buffer)
    for tshift in (-1, 0, 2, 1000, 1001, 1003):
        got = emu.corr3d_dist(a, b, dr, window[0], window[1], G, alpha0=1.08, tshift=tshift)
        assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max(), (shape, window, G, tshift)


@pytest.mark.parametrize("shape,window,support,G", [
    ((8, 8, 64), (20, 30), (60, 9), 2), ((6, 12, 48), (10, 22), (44, 8), 3), ((8, 16, 40), (3, 9), (0, 5), 4),
    ((4, 8, 32), (0, 32), (30, 4), 2), ((12, 8, 96), (40, 61), (90, 13), 2), ((8, 8, 64), (20, 30), (60, 9), 1),
])
def test_corr3d_distributed_pruned_equals_full_correlation(emu, shape, window, support, G):
    """Compact tip on the distributed path: the forward y pass reads the z-ranges of the pruned problem in place, all
    later stages run on the shorter z length; same window as the full periodic correlation."""
    rng = np.random.default_rng(sum(shape) + window[0] + G)
    nz = shape[2]
    a = rng.random(shape)
    b = np.zeros(shape)
    s_lo, s_len = support
    planes = (s_lo + np.arange(s_len)) % nz
    b[:, :, planes] = rng.random(shape[:2] + (s_len,)) + 0.1
    dr = np.array([0.1, 0.2, 0.3])
    ref = O.density_overlap_fft(a ** 1.08, b ** 1.08, dr)[:, :, window[0]:window[1]]
    for tshift in (-1, 0, 1000):
        got, used = emu.corr3d_dist_pruned(a, b, dr, window[0], window[1], s_lo, s_len, G, alpha0=1.08, alpha1=1.08, tshift=tshift)
        assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max(), (shape, G, tshift)
        assert (used == nz) == (shape == (4, 8, 32)), used          # the full window of (4,8,32) cannot be pruned


def test_corr3d_distributed_random_shapes(emu):
    """Property test (hypothesis): for random 2-3-5-7-smooth shapes, rank counts, windows and tile widths the distributed
    step -- with all-to-alls or with both exchanges fused as peer stores -- equals the reference's periodic correlation."""
    from hypothesis import given, settings, strategies as st, HealthCheck

    smooth = [v for v in range(1, 33) if all(v % p for p in (11, 13, 17, 19, 23, 29, 31))]

    @st.composite
    def cases(draw):
        G = draw(st.sampled_from([1, 2, 3, 4, 5, 8]))
        nx = G * draw(st.sampled_from([v for v in smooth if v * G <= 40]))
        ny = 2 * G * draw(st.sampled_from([v for v in smooth if 2 * v * G <= 48]))
        nz = 2 * draw(st.sampled_from([v for v in smooth if v <= 16]))
        z_lo = draw(st.integers(0, nz - 1))
        z_hi = draw(st.integers(z_lo + 1, nz))
        ts = draw(st.sampled_from([-1, 0, 1, 2, 1000, 1001, 1003]))
        return G, (nx, ny, nz), (z_lo, z_hi), ts

    @settings(max_examples=40, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.too_slow])
    @given(cases())
    def run(case):
        G, shape, (z_lo, z_hi), ts = case
        rng = np.random.default_rng(sum(shape) * 7 + G)
        a, b = rng.random(shape), rng.random(shape) - 0.4
        dr = np.array([0.1, 0.2, 0.3])
        ref = O.density_overlap_fft(a ** 1.08, b, dr)[:, :, z_lo:z_hi]
        got = emu.corr3d_dist(a, b, dr, z_lo, z_hi, G, alpha0=1.08, tshift=ts)
        assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max(), case

    run()


def test_slot_order_host_and_device_agree(emu):
    """parallel.slot_order (host) and dist_slot_of / dist_mirror_slot / dist_row_of (kernels) describe the same deal of
    y-frequencies: k and -k are slot neighbours, equal even blocks are closed under k -> -k."""
    from dbspm_b200 import parallel
    L = emu.L
    for n in (2, 4, 6, 8, 20, 30, 196, 1024):
        order = parallel.slot_order(n)
        assert sorted(order) == list(range(n))
        for S, K in enumerate(order):
            assert L.emu_dist_slot_of(K, n) == S
            assert order[L.emu_dist_mirror_slot(S)] == (n - K) % n
        for G in (1, 2, 4, 5, 8):
            if n % (2 * G):
                continue
            nl = n // G
            for h in range(G):
                block = {order[S] for S in range(h * nl, (h + 1) * nl)}
                assert block == {(n - K) % n for K in block}            # closed under the mirror
            for K in range(n):
                S = L.emu_dist_slot_of(K, n)
                assert L.emu_dist_row_of(K, n, G, 1000) == (S // nl) * 1000 + S % nl


def test_pow_fast_accuracy(emu):
    """The table-driven v**alpha of the powered first pass against numpy's pow in extended precision."""
    rng = np.random.default_rng(11)
    for alpha in (1.08, 1.12, 0.5, 2.0, 3.7):
        for (lo, hi), tol in (((-12, 2), 3e-14), ((-300, 300), 3e-13)):
            x = np.concatenate([10.0 ** rng.uniform(lo, hi, 50000), [1.0, 2.0, 0.5, np.nextafter(1, 2), np.nextafter(1, 0),
                                                                     1 + 1 / 32, 1 + 1 / 64, 2.3e-308]])
            ref = np.power(x.astype(np.longdouble), np.longdouble(alpha))
            with np.errstate(over="ignore"):
                ok = np.isfinite(ref.astype(np.float64)) & (ref > 1e-300)
            got = emu.pow_fast(x, alpha)
            assert float(np.abs((got[ok] - ref[ok]) / ref[ok]).max()) <= tol, (alpha, lo, hi)
    # zero, negative, non-finite and subnormal inputs take the libm path: same values as numpy
    x = np.array([0.0, -1.0, np.inf, np.nan, 5e-324, 1e-310])
    with np.errstate(all="ignore"):
        want = np.power(x, 1.08)
    assert np.array_equal(emu.pow_fast(x, 1.08), want, equal_nan=True)
    # unfiltered densities with small negative noise: numpy's `**` stays finite for an integer-valued exponent
    x = np.array([-2.0, -1e-7, -0.5, 3.0, 0.0])
    for alpha in (2.0, 3.0):
        got = emu.pow_fast(x, alpha)
        assert np.allclose(got, np.power(x, alpha), rtol=1e-14, atol=0) and np.array_equal(np.sign(got), np.sign(np.power(x, alpha)))
    a = np.random.default_rng(5).standard_normal((6, 4, 8)) * 0.1
    b = np.random.default_rng(6).random((6, 4, 8))
    ref = O.density_overlap_fft(a ** 2.0, b ** 2.0, np.ones(3))
    assert np.abs(emu.corr3d(a, b, np.ones(3), alpha0=2.0, alpha1=2.0) - ref).max() <= 1e-12 * np.abs(ref).max()


def test_corr3d_tiny_config_window(emu, overlap_golden):
    from dbspm_b200 import synth
    cfg = synth.CONFIGS["tiny"]
    rhoS, potS, _ = synth.make_sample(cfg)
    rhoT, nrhoT = synth.make_tip(cfg, variant="noisy")
    dr = overlap_golden["tiny_dr"]
    sr = emu.corr3d(rhoS.numpy(), rhoT.numpy(), dr, 20, 44, cfg.alpha, cfg.alpha)
    es = emu.corr3d(potS.numpy(), nrhoT.numpy(), dr, 20, 44)
    for got, ref in ((sr, overlap_golden["tiny_sr_win"]), (es, overlap_golden["tiny_es_win"])):
        assert np.abs(got - ref).max() <= 1e-9 * np.abs(ref).max()
    # zero density stays zero under the fused power (0**alpha = 0), negative -> NaN like numpy
    z = np.zeros((4, 4, 4)); z[1, 2, 3] = 2.0
    out = emu.corr3d(z, z, dr, 0, 4, 1.08, 1.08)
    assert np.isfinite(out).all() and abs(out[0, 0, 0] - np.prod(dr) * (2.0 ** 1.08) ** 2) < 1e-12


@pytest.mark.parametrize("shape,window", [((4, 4, 5), (0, 5)), ((6, 5, 9), (2, 7)), ((3, 8, 15), (14, 15)),
                                          ((1, 1, 1), (0, 1)), ((7, 1, 3), (1, 2)), ((10, 12, 21), (0, 21))])
def test_odd_nz_takes_the_complex_path(emu, shape, window):
    """The packed real transform needs an even z length; odd lengths go through plain complex transforms
    (get_density_overlap of the reference accepts any shape)."""
    rng = np.random.default_rng(sum(shape))
    a, b = rng.random(shape), rng.random(shape) - 0.3
    dr = np.array([0.2, 0.1, 0.3])
    ref = O.density_overlap_fft(a ** 1.08, b, dr)[:, :, window[0]:window[1]]
    for ts in (-1, 101, 200):
        got = emu.corr3d(a, b, dr, window[0], window[1], alpha0=1.08, tshift=ts)
        assert np.abs(got - ref).max() <= 1e-12 * max(np.abs(ref).max(), 1e-300), (shape, ts)


@pytest.mark.parametrize("ni,nj", [(1, 1), (8, 4), (70, 45), (3, 45), (70, 2), (33, 31), (128, 1024)])
def test_relax_patch_order_covers_every_column_once(emu, ni, nj):
    """relax_patch_column: whatever patch a warp relaxes (foot_x x 32/foot_x), every column of the tile is taken by exactly
    one thread, idle lanes return ncol, a warp's columns lie in one patch, and foot_x = 1 is the identity."""
    ncol = ni * nj
    for fx in (1, 2, 4, 8, 16, 32):
        cols = emu.relax_patch_columns(ni, nj, fx)
        assert cols.size % 32 == 0 or fx == 1
        taken = cols[cols < ncol]
        assert taken.size == ncol and np.array_equal(np.sort(taken), np.arange(ncol)), (ni, nj, fx)
        assert np.all(cols[cols >= ncol] == ncol)
        if fx == 1:
            assert np.array_equal(cols, np.arange(ncol))
            continue
        fy = 32 // fx
        for w in range(0, cols.size, 32):
            c = cols[w:w + 32]; c = c[c < ncol]
            if c.size:
                ci, cj = c // nj, c % nj
                assert ci.max() - ci.min() < fx and cj.max() - cj.min() < fy and ci.min() % fx == 0 and cj.min() % fy == 0


def test_tricubic_stencil_equals_oracle(emu, relax_golden):
    s = relax_golden["static"]
    tc = O.Tricubic(s)
    rng = np.random.default_rng(3)
    for _ in range(200):
        p = rng.uniform(-60, 120, 3)
        assert abs(emu.tricubic(s, *p) - tc.ip(list(p))) <= 1e-13 * max(1.0, np.abs(s).max())
    assert emu.tricubic(s, 3.0, 4.0, 5.0) == s[3, 4, 5]


def test_relax_matches_reference_goldens(emu, relax_golden, manifest):
    g, m = relax_golden, manifest["relax"]
    static, dr, cols = g["static"], g["dr"], g["cols"]
    rp = m["rpivot"]
    dev = []
    for c, (i, j) in enumerate(cols):
        etp, cart, nfev = emu.relax(static, (i, i + 1, j, j + 1), m["nz"], m["kappa"], rp, dr)
        ref = g["ortho"][c]
        assert np.abs(etp[0, 0, :, 0] - ref[:, 0]).max() <= 1e-6 * np.abs(ref[:, 0]).max()
        want = np.stack(O.sph2cart(ref[:, 1], ref[:, 2], rp))
        dev.append(np.abs(cart[:, 0, 0] - want).max(axis=0) / rp)
    dev = np.concatenate(dev)
    # the reference stops at xtol = 1e-4 with a 0.01 line-search tolerance, so its answer depends
    # (chaotically, through the warm start) on last-bit differences of the interpolant: a few
    # positions leave the 1e-6*rpivot band; none leaves 1e-4*rpivot
    assert (dev > 1e-6).mean() <= 0.05 and dev.max() <= 1e-4, (dev.max(), (dev > 1e-6).mean())
    for c, (i, j) in enumerate(cols[:3]):
        etp, cart, _ = emu.relax(static, (i, i + 1, j, j + 1), m["nz"], m["kappa"], rp, dr, dR=np.diag(dr))
        assert np.abs(etp[0, 0, :, 0] - g["noortho"][c][:, 0]).max() <= 1e-6 * np.abs(g["noortho"][c][:, 0]).max()
        s = m["stiff"]
        etp, cart, _ = emu.relax(static, (i, i + 1, j, j + 1), s["nz"], s["kappa"], s["rpivot"], dr)
        assert np.abs(etp[0, 0, :, 0] - g["stiff"][c][:, 0]).max() <= 1e-6 * np.abs(g["stiff"][c][:, 0]).max()


def test_relax_tile_against_oracle_with_outlier_budget(emu, relax_golden):
    """Whole tile: E within 1e-6 relative; tip positions within 1e-6*rpivot except for a small
    fraction of Powell branch flips (ulp-level differences of the interpolant)."""
    static, dr = relax_golden["static"], relax_golden["dr"]
    ref, nf_ref = O.relax_tile(static, 0, 24, 0, 20, 24, 0.2, 3.02, dr, want_nfev=True)
    etp, cart, nfev = emu.relax(static, (0, 24, 0, 20), 24, 0.2, 3.02, dr)
    assert np.abs(etp[..., 0] - ref[..., 0]).max() <= 1e-6 * np.abs(ref[..., 0]).max()
    want = np.stack(O.sph2cart(ref[..., 1], ref[..., 2], 3.02))
    bad = (np.abs(cart - want).max(axis=0) > 1e-6 * 3.02).mean()
    assert bad < 2e-3, bad
    assert abs(nfev.mean() - nf_ref.mean()) < 1.0
    # maxfev exhaustion: same early-exit semantics as scipy's _MaxFuncCallError path
    r2, n2 = O.relax_tile(static, 2, 6, 2, 6, 24, 0.2, 3.02, dr, maxfev=17, want_nfev=True)
    e2, _, m2 = emu.relax(static, (2, 6, 2, 6), 24, 0.2, 3.02, dr, maxfev=17)
    assert np.array_equal(n2, m2) and np.abs(e2[..., 0] - r2[..., 0]).max() < 1e-9


def test_relax_transposed_layout_is_bit_identical(emu, relax_golden):
    """The y-fastest copy of the potential (coalesced loads on the GPU) must not change a bit."""
    static, dr = relax_golden["static"], relax_golden["dr"]
    a = emu.relax(static, (0, 24, 0, 20), 8, 0.2, 3.02, dr)
    b = emu.relax(static, (0, 24, 0, 20), 8, 0.2, 3.02, dr, transposed=1)
    for u, v in zip(a, b):
        assert np.array_equal(u, v)
    odd = np.ascontiguousarray(static[:23, :19, :37])            # tile edges of the 32x32 transpose
    a = emu.relax(odd, (3, 9, 2, 17), 5, 0.2, 3.02, dr)
    b = emu.relax(odd, (3, 9, 2, 17), 5, 0.2, 3.02, dr, transposed=1)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2])


def test_relax_quad_layout_is_bit_identical(emu, relax_golden):
    """LAYOUT 2 (the four y neighbours of a stencil row as one 32-byte sector, 256-bit loads on the GPU) through
    quad_yz_tile, including grids whose edges do not fill the 32 x 32 transpose tiles."""
    static, dr = relax_golden["static"], relax_golden["dr"]
    a = emu.relax(static, (0, 24, 0, 20), 8, 0.2, 3.02, dr)
    b = emu.relax(static, (0, 24, 0, 20), 8, 0.2, 3.02, dr, transposed=2)
    for u, v in zip(a, b):
        assert np.array_equal(u, v)
    rng = np.random.default_rng(4)
    big = np.ascontiguousarray(np.tile(static, (1, 2, 1))[:, :37, :41]) + 1e-3 * rng.random((24, 37, 41))   # ny = 37, nz = 41
    a = emu.relax(big, (3, 9, 0, 37), 5, 0.2, 3.02, dr)
    b = emu.relax(big, (3, 9, 0, 37), 5, 0.2, 3.02, dr, transposed=2)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2])


def test_relax_x_slab_mode_is_bit_identical(emu, relax_golden):
    """Multi-GPU relax: a rank holds its x-tile of the static window plus a halo of the lever length; the kernel maps
    the global (periodically wrapped) plane index into the slab.  Same bits as the whole window, across the wrap."""
    static, dr = relax_golden["static"], relax_golden["dr"]
    nx = static.shape[0]
    rp, halo = 0.3, 6                                               # ceil(0.3 / 0.0765) + 2
    full = emu.relax(static, (0, nx, 0, 20), 6, 0.2, rp, dr)
    for i_lo, i_hi in ((20, 24), (0, 5), (9, 12)):
        x_base = (i_lo - halo) % nx
        planes = (x_base + np.arange((i_hi - i_lo) + 2 * halo)) % nx
        for layout in (0, 1, 2):
            got = emu.relax_slab(static[planes], nx, x_base, (i_lo, i_hi, 0, 20), 6, 0.2, rp, dr, transposed=layout)
            for u, v in zip(got, full):
                assert np.array_equal(u, v[:, i_lo:i_hi] if v.shape[0] == 3 else v[i_lo:i_hi]), (i_lo, layout)


def test_kernel_body_equals_the_oracle_twin_bit_for_bit(emu, relax_golden):
    """The kernel body (state machine, stencil, fixed-sequence sin/cos, correctly rounded division) against the
    INDEPENDENT restatement in oracle/dbspm_oracle.c ("kernel twin": scipy's Powell in C + tc_ip_twin + det_sincos):
    every E, theta, phi, nfev and Cartesian tip position must have the same bits.  The GPU test of the same name
    (tests/test_gpu_parity.py) then pins the sm_100a kernel to the twin; what remains between the twin and the
    reference-faithful mode is measured below, on the CPU alone."""
    static, dr = relax_golden["static"], relax_golden["dr"]
    tw, nf = O.relax_tile(static, 0, 24, 0, 20, 24, 0.2, 3.02, dr, want_nfev=True, twin=True)
    for layout in (0, 1, 2):
        etp, cart, nfev = emu.relax(static, (0, 24, 0, 20), 24, 0.2, 3.02, dr, transposed=layout)
        assert np.array_equal(etp, tw) and np.array_equal(nfev, nf), layout
        assert np.array_equal(cart, O.sph2cart_twin(tw[..., 1], tw[..., 2], 3.02))
    # stiff spring, short lever, maxfev exhaustion
    tw2, nf2 = O.relax_tile(static, 2, 9, 1, 12, 10, 0.5, 2.0, dr, want_nfev=True, twin=True, maxfev=23)
    etp, _, nfev = emu.relax(static, (2, 9, 1, 12), 10, 0.5, 2.0, dr, maxfev=23)
    assert np.array_equal(etp, tw2) and np.array_equal(nfev, nf2)
    # twin vs reference-faithful mode (Lekien-Marsden + libm): E to 1e-6, positions to 1e-6 * rpivot except for a
    # few Powell branch flips -- this is the whole difference between the GPU result and the reference's arithmetic
    ref, nfr = O.relax_tile(static, 0, 24, 0, 20, 24, 0.2, 3.02, dr, want_nfev=True)
    assert np.abs(tw[..., 0] - ref[..., 0]).max() <= 1e-6 * np.abs(ref[..., 0]).max()
    d = np.abs(O.sph2cart_twin(tw[..., 1], tw[..., 2], 3.02) - np.stack(O.sph2cart(ref[..., 1], ref[..., 2], 3.02))).max(axis=0) / 3.02
    assert (d > 1e-6).mean() <= 1e-3 and d.max() <= 3e-5, ((d > 1e-6).mean(), d.max())


def test_noortho_skew_cell_matches_reference_golden_and_twin(emu, relax_golden, manifest):
    """relax_noortho_tip_z (interactions.py:88-105) with a genuinely non-orthogonal dR: the fixture was produced by
    the reference's own function (oracle/make_golden.py noortho); the kernel body must agree with it to the
    north_star tolerances and with the oracle twin bit for bit."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "relax_noortho_reference.npz"))
    m = manifest["relax_noortho"]
    static, dr, dR = relax_golden["static"], relax_golden["dr"], g["dR"]
    assert np.abs(dR - np.diag(np.diag(dR))).max() > 1e-3                  # really skew
    devs = []
    for c, (i, j) in enumerate(g["cols"]):
        i, j = int(i), int(j)
        ref = g["skew"][c]
        orc = O.relax_tile(static, i, i + 1, j, j + 1, m["nz"], m["kappa"], m["rpivot"], dr, dR=dR)[0, 0]
        # the reference inverts dR.T with LAPACK (np.linalg.inv) and multiplies with BLAS on every evaluation; a cofactor
        # inverse differs from that in the last bits, so for a skew cell the C restatement is pinned to the tolerance, not
        # to the bit (with dR = diag(dr) it IS bit-identical: tests/test_oracle.py)
        assert np.abs(orc[:, 0] - ref[:, 0]).max() <= 1e-6 * np.abs(ref[:, 0]).max(), (i, j)
        etp, cart, _ = emu.relax(static, (i, i + 1, j, j + 1), m["nz"], m["kappa"], m["rpivot"], dr, dR=dR)
        tw = O.relax_tile(static, i, i + 1, j, j + 1, m["nz"], m["kappa"], m["rpivot"], dr, dR=dR, twin=True)
        assert np.array_equal(etp, tw)
        assert np.abs(etp[0, 0, :, 0] - ref[:, 0]).max() <= 1e-6 * np.abs(ref[:, 0]).max()
        want = np.stack(O.sph2cart(ref[:, 1], ref[:, 2], m["rpivot"]))
        devs.append(np.abs(cart[:, 0, 0] - want).max(axis=0) / m["rpivot"])
    devs = np.concatenate(devs)
    assert devs.max() <= 1e-6, devs.max()                              # every position (measured: 4.7e-8)
    # the skew really matters: the orthogonal path gives a different answer
    etp_o, _, _ = emu.relax(static, (5, 6, 7, 8), m["nz"], m["kappa"], m["rpivot"], dr)
    assert np.abs(etp_o[0, 0, :, 0] - g["skew"][1][:, 0]).max() > 1e-5 * np.abs(g["skew"][1][:, 0]).max()


def test_fixed_sequence_sincos(emu):
    """dbspm_sincos: within one ulp of the host libm over the range a relaxation visits (and far beyond), exact at the
    special angles of the start point, and the kernel's statement equals the oracle's bit for bit."""
    import ctypes as C
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-10, 10, 400000), rng.uniform(-1e5, 1e5, 400000), np.pi + rng.uniform(-1e-6, 1e-6, 50000),
                        np.array([np.pi, 0.0, -np.pi, np.pi / 2, 1e-300, -0.0, 3 * np.pi / 4, 2e5, -3.7e8])])
    s, c = np.empty_like(x), np.empty_like(x)
    f = emu.L.emu_sincos
    f.argtypes = [C.POINTER(C.c_double)] * 3 + [C.c_longlong]
    f(emu._p(x), emu._p(s), emu._p(c), x.size)
    ulp = lambda a, b: np.abs(a - b) / np.spacing(np.abs(b))
    assert ulp(s, np.sin(x)).max() <= 1.0 and ulp(c, np.cos(x)).max() <= 1.0
    assert s[850000] == np.sin(np.pi) and c[850000] == -1.0 and s[850001] == 0.0 and c[850001] == 1.0
    so, co = O.det_sincos(x)
    assert np.array_equal(s, so) and np.array_equal(c, co)


def test_relax_empty_tile_and_periodic_wrap(emu, relax_golden):
    static, dr = relax_golden["static"], relax_golden["dr"]
    etp, cart, nfev = emu.relax(static, (3, 3, 0, 5), 4, 0.2, 3.02, dr)
    assert etp.size == 0
    # columns at the lateral edges use the periodic wrap of the interpolant: shift the grid
    # by one cell and the same physical column must give the same answer
    rolled = np.ascontiguousarray(np.roll(static, (1, 1), axis=(0, 1)))
    a, _, _ = emu.relax(static, (0, 1, 0, 1), 6, 0.2, 3.02, dr)
    b, _, _ = emu.relax(rolled, (1, 2, 1, 2), 6, 0.2, 3.02, dr)
    assert np.abs(a[..., 0] - b[..., 0]).max() < 1e-12


def test_division_by_known_spacing_is_correctly_rounded(emu):
    """relax_objective divides the tip offset by dr with `div_by` (multiply + two fused corrections);
    it must give the bits of the IEEE quotient the reference's numpy computes."""
    import ctypes as C
    f = emu.L.emu_check_div_by
    f.restype, f.argtypes = C.c_longlong, [C.c_longlong, C.c_ulonglong]
    assert f(10_000_000, 1234) == 0


@pytest.mark.parametrize("shape,window,support", [
    ((8, 6, 40), (10, 16), (36, 9)),      # support wraps through z = 0, even start
    ((8, 6, 40), (11, 18), (37, 8)),      # odd start of both cuts, odd window
    ((6, 10, 48), (0, 5), (45, 7)),       # window at the bottom: the A cut wraps
    ((6, 10, 48), (40, 48), (2, 5)),      # window at the top: the A cut wraps upwards
    ((12, 4, 64), (20, 21), (0, 1)),      # single-plane tip
    ((4, 4, 32), (3, 30), (30, 6)),       # window + support > 0.9 nz -> full-length path
    ((22, 4, 40), (10, 16), (36, 9)),     # nx = 2*11: remapped first pass unavailable -> full-length path
])
def test_corr3d_pruned_equals_full_correlation(emu, shape, window, support):
    """Tip exactly zero outside `support` planes: the z-pruned path (cut, zero-pad, shorter transform)
    gives the window of the full periodic correlation."""
    rng = np.random.default_rng(sum(shape) + window[0])
    nz = shape[2]
    a = rng.random(shape)
    b = np.zeros(shape)
    s_lo, s_len = support
    planes = (s_lo + np.arange(s_len)) % nz
    b[:, :, planes] = rng.random(shape[:2] + (s_len,)) - 0.3
    dr = np.array([0.1, 0.2, 0.3])
    ref = O.density_overlap_fft(a ** 1.08, b, dr)[:, :, window[0]:window[1]]
    got, used = emu.corr3d_pruned(a, b, dr, window[0], window[1], s_lo, s_len, alpha0=1.08)
    assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()
    expect_full = shape in [(4, 4, 32), (22, 4, 40)]
    assert (used == nz) == expect_full, used
    # a support claim wider than the real one is harmless
    got2, _ = emu.corr3d_pruned(a, b, dr, window[0], window[1], s_lo - 2, s_len + 3, alpha0=1.08)
    assert np.abs(got2 - ref).max() <= 1e-12 * np.abs(ref).max()


@pytest.mark.parametrize("shape,window,center,radius", [
    ((12, 10, 16), (3, 11), (0, 0, 0), 2.2),          # ball around the origin: wraps on all three axes
    ((9, 14, 20), (0, 20), (4, 7, 10), 1.5),          # off-centre support, whole z range
    ((16, 16, 24), (20, 24), (15, 1, 23), 2.9),       # window at the top, support straddling the corner
    ((7, 5, 8), (2, 3), (0, 0, 0), 0.0),              # a single point (delta tip)
    ((40, 36, 32), (5, 27), (0, 0, 0), 3.4),          # several tiles in x and y
])
def test_direct_real_space_correlation(emu, shape, window, center, radius):
    """Direct path (tip = a few hundred support points): plan (circular bounding box, chunks), pad kernel with periodic halo,
    tile kernel with the emulated tensor-map load -- against the reference's FFT expression and the explicit periodic sum."""
    rng = np.random.default_rng(sum(shape) + window[0])
    a = rng.random(shape) + 0.05
    b = np.zeros(shape)
    g = np.stack(np.meshgrid(*[np.arange(n) for n in shape], indexing="ij"), axis=-1)
    d = g - np.array(center)
    d = np.minimum(np.abs(d), np.array(shape) - np.abs(d))                       # periodic distance
    mask = (d ** 2).sum(axis=-1) <= radius ** 2 + 1e-9
    b[mask] = rng.random(int(mask.sum())) - 0.3
    b[tuple(center)] = 0.7
    if radius > 2:
        b[tuple((np.array(center) + [1, 0, 1]) % shape)] = 0.0                   # a hole inside a chunk
    pts = np.argwhere(b != 0.0)
    dr = np.array([0.1, 0.2, 0.3])
    for alpha in (1.0, 1.08):
        w = np.abs(b[b != 0.0]) ** alpha if alpha != 1.0 else b[b != 0.0]
        bb = np.zeros(shape); bb[b != 0.0] = w
        ref = O.density_overlap_fft(a ** alpha, bb, dr)[:, :, window[0]:window[1]]
        got, cost = emu.corr3d_direct(a, pts, w, dr, window[0], window[1], alpha0=alpha)
        assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max(), (shape, alpha)
        assert cost >= len(w)
    if np.prod(shape) <= 4000:
        direct = O.corr_direct(a, b, dr, window[0], window[1])
        got, _ = emu.corr3d_direct(a, pts, b[b != 0.0], dr, window[0], window[1])
        assert np.abs(got - direct).max() <= 1e-13 * np.abs(direct).max()
