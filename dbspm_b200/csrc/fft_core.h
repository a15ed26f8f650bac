// fft_core.h -- batched in-place mixed-radix FFT on a shared-memory tile.
//
// Tile layout: s[pos * ld + t], pos in [0,n) along the transform axis, t in [0,T) the
// batch ("line") index, T a power of two.  Decimation in frequency, in place: after the
// last stage the value of frequency K sits at position pos_of[K] (mixed-radix digit
// reversal, table built on the host in plan.cpp).  DIR = -1 forward (e^{-i..}), +1 inverse.
//
// Threads map to (t fastest, butterfly index slow): the 8 threads of a quarter warp
// read 8 consecutive 16-byte values -> conflict-free 128-bit shared-memory accesses
// at every stage, independent of the butterfly stride.
#pragma once
#include "hd.h"

#define DBSPM_MAX_STAGES 16
#define DBSPM_MAX_RADIX 32

struct FftPlanDev {
    int n;
    int nstages;
    int radix[DBSPM_MAX_STAGES];
};

// ---- radix butterflies (values in registers) --------------------------------------------
template <int D> HD void bfly2(cplx *v)
{
    cplx a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
}

template <int D> HD void bfly4(cplx *v)
{
    cplx s02 = cadd(v[0], v[2]), d02 = csub(v[0], v[2]);
    cplx s13 = cadd(v[1], v[3]), d13 = cmul_iD<D>(csub(v[1], v[3]));
    v[0] = cadd(s02, s13);
    v[1] = cadd(d02, d13);
    v[2] = csub(s02, s13);
    v[3] = csub(d02, d13);
}

template <int D> HD void bfly8(cplx *v)
{
    const double h = 0.70710678118654752440;
    cplx e[4] = {v[0], v[2], v[4], v[6]};
    cplx o[4] = {v[1], v[3], v[5], v[7]};
    bfly4<D>(e);
    bfly4<D>(o);
    // W8^k = e^{i D pi k / 4}
    cplx o1 = cmake(h * (o[1].x - D * o[1].y), h * (o[1].y + D * o[1].x));   // * (1 + iD)/sqrt2
    cplx o2 = cmul_iD<D>(o[2]);                                              // * iD
    cplx o3 = cmake(h * (-o[3].x - D * o[3].y), h * (-o[3].y + D * o[3].x)); // * (-1 + iD)/sqrt2
    v[0] = cadd(e[0], o[0]); v[4] = csub(e[0], o[0]);
    v[1] = cadd(e[1], o1);   v[5] = csub(e[1], o1);
    v[2] = cadd(e[2], o2);   v[6] = csub(e[2], o2);
    v[3] = cadd(e[3], o3);   v[7] = csub(e[3], o3);
}

// v * (cr + i*D*ci)
template <int D> HD cplx cmul_cD(cplx v, double cr, double ci)
{
    return cmake(v.x * cr - D * (v.y * ci), v.y * cr + D * (v.x * ci));
}

// radix 16 = 4 x 4: X[k1 + 4 k2] = sum_b W16^{b k1} W4^{b k2} sum_a x[4a + b] W4^{a k1}
template <int D> HD void bfly16(cplx *v)
{
    const double c1 = 0.92387953251128674, s1 = 0.38268343236508977, h = 0.70710678118654752440;
    cplx t[4][4];                                     // t[b][k1]
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        cplx col[4] = {v[b], v[4 + b], v[8 + b], v[12 + b]};
        bfly4<D>(col);
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1) t[b][k1] = col[k1];
    }
    // twiddles W16^{b k1} = exp(i D 2 pi b k1 / 16)
    t[1][1] = cmul_cD<D>(t[1][1], c1, s1);            // W^1
    t[1][2] = cmul_cD<D>(t[1][2], h, h);              // W^2
    t[1][3] = cmul_cD<D>(t[1][3], s1, c1);            // W^3
    t[2][1] = cmul_cD<D>(t[2][1], h, h);              // W^2
    t[2][2] = cmul_iD<D>(t[2][2]);                    // W^4
    t[2][3] = cmul_cD<D>(t[2][3], -h, h);             // W^6
    t[3][1] = cmul_cD<D>(t[3][1], s1, c1);            // W^3
    t[3][2] = cmul_cD<D>(t[3][2], -h, h);             // W^6
    t[3][3] = cmul_cD<D>(t[3][3], -c1, -s1);          // W^9
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) {
        cplx row[4] = {t[0][k1], t[1][k1], t[2][k1], t[3][k1]};
        bfly4<D>(row);
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) v[k1 + 4 * k2] = row[k2];
    }
}

template <int P> struct OddTab;
template <> struct OddTab<3> {
    static HD double c(int m) { const double t[3] = {1.0, -0.5, -0.5}; return t[m]; }
    static HD double s(int m) { const double t[3] = {0.0, 0.8660254037844386, -0.8660254037844386}; return t[m]; }
};
template <> struct OddTab<5> {
    static HD double c(int m) { const double t[5] = {1.0, 0.30901699437494745, -0.8090169943749475, -0.8090169943749475, 0.30901699437494745}; return t[m]; }
    static HD double s(int m) { const double t[5] = {0.0, 0.9510565162951535, 0.5877852522924731, -0.5877852522924731, -0.9510565162951535}; return t[m]; }
};
template <> struct OddTab<7> {
    static HD double c(int m) { const double t[7] = {1.0, 0.6234898018587335, -0.2225209339563144, -0.9009688679024191, -0.9009688679024191, -0.2225209339563144, 0.6234898018587335}; return t[m]; }
    static HD double s(int m) { const double t[7] = {0.0, 0.7818314824680298, 0.9749279121818236, 0.4338837391175581, -0.4338837391175581, -0.9749279121818236, -0.7818314824680298}; return t[m]; }
};

// odd prime P: X_k, X_{P-k} from the symmetric / antisymmetric input pairs
template <int P, int D> HD void bfly_odd(cplx *v)
{
    constexpr int H = (P - 1) / 2;
    cplx sp[H], dm[H];
#pragma unroll
    for (int j = 1; j <= H; ++j) {
        sp[j - 1] = cadd(v[j], v[P - j]);
        dm[j - 1] = csub(v[j], v[P - j]);
    }
    cplx x0 = v[0];
    cplx sum = x0;
#pragma unroll
    for (int j = 0; j < H; ++j) sum = cadd(sum, sp[j]);
    v[0] = sum;
#pragma unroll
    for (int k = 1; k <= H; ++k) {
        cplx a = x0, b = cmake(0.0, 0.0);
#pragma unroll
        for (int j = 1; j <= H; ++j) {
            const double cc = OddTab<P>::c((j * k) % P), ss = OddTab<P>::s((j * k) % P);
            a.x += cc * sp[j - 1].x; a.y += cc * sp[j - 1].y;
            b.x += ss * dm[j - 1].x; b.y += ss * dm[j - 1].y;
        }
        cplx ib = cmul_iD<D>(b);
        v[k] = cadd(a, ib);
        v[P - k] = csub(a, ib);
    }
}

template <int R, int D> HD void bfly(cplx *v)
{
    if (R == 2) bfly2<D>(v);
    else if (R == 3) bfly_odd<3, D>(v);
    else if (R == 4) bfly4<D>(v);
    else if (R == 5) bfly_odd<5, D>(v);
    else if (R == 7) bfly_odd<7, D>(v);
    else if (R == 8) bfly8<D>(v);
    else if (R == 16) bfly16<D>(v);
}

// ---- one DIF stage over the whole tile ----------------------------------------------------
// m = current sub-transform length, R | m.  tw[e] = exp(-2 pi i e / n), e in [0,n).
template <int R, int D>
HD void fft_stage(cplx *s, int ld, int tshift, int n, int m, const cplx *tw, int tid, int nthr)
{
    const int q = m / R;
    const int nb = n / R;
    const int twstride = n / m;
    const int tmask = (1 << tshift) - 1;
    const int total = nb << tshift;
    for (int idx = tid; idx < total; idx += nthr) {
        const int t = idx & tmask;
        const int b = idx >> tshift;
        const int blk = b / q;
        const int j = b - blk * q;
        cplx *p = s + ((blk * m + j) * ld + t);          // 32-bit index: tiles are far below 2^31 elements
        const int step = q * ld;
        cplx v[R];
#pragma unroll
        for (int k = 0; k < R; ++k) v[k] = p[k * step];
        bfly<R, D>(v);
        if (q > 1) {
            const int e1 = j * twstride;
#pragma unroll
            for (int k = 1; k < R; ++k) {
                cplx w = tw[e1 * k];
                v[k] = D < 0 ? cmul(v[k], w) : cmulc(v[k], w);
            }
        }
#pragma unroll
        for (int k = 0; k < R; ++k) p[k * step] = v[k];
    }
}

// generic radix (any r <= DBSPM_MAX_RADIX dividing m): O(r^2) with table twiddles
template <int D>
HD void fft_stage_generic(cplx *s, int ld, int tshift, int n, int m, int r, const cplx *tw, int tid, int nthr)
{
    const int q = m / r;
    const int nb = n / r;
    const int twstride = n / m;
    const int rstride = n / r;          // tw[rstride * a] = exp(-2 pi i a / r)
    const int tmask = (1 << tshift) - 1;
    const int total = nb << tshift;
    for (int idx = tid; idx < total; idx += nthr) {
        const int t = idx & tmask;
        const int b = idx >> tshift;
        const int blk = b / q;
        const int j = b - blk * q;
        cplx *p = s + ((blk * m + j) * ld + t);          // 32-bit index: tiles are far below 2^31 elements
        const int step = q * ld;
        cplx v[DBSPM_MAX_RADIX], o[DBSPM_MAX_RADIX];
        for (int k = 0; k < r; ++k) v[k] = p[k * step];
        for (int k = 0; k < r; ++k) {
            cplx acc = v[0];
            for (int a = 1; a < r; ++a) {
                cplx w = tw[((a * k) % r) * rstride];
                acc = cadd(acc, D < 0 ? cmul(v[a], w) : cmulc(v[a], w));
            }
            if (k > 0 && q > 1) {
                cplx w = tw[j * twstride * k];
                acc = D < 0 ? cmul(acc, w) : cmulc(acc, w);
            }
            o[k] = acc;
        }
        for (int k = 0; k < r; ++k) p[k * step] = o[k];
    }
}

// ---- transposed stage (decimation in time): twiddle first, then butterfly ----------------------
// The DFT matrix is symmetric, so running the transposes of the DIF stages in reverse order maps
// an input stored in the DIF *output* order (value of frequency K at position pos_of[K]) to the
// transform in natural order.  Used for the inverse z transform: the spectrum never has to be
// un-permuted between the forward and the inverse FFT.
template <int R, int D>
HD void fft_stage_t(cplx *s, int ld, int tshift, int n, int m, const cplx *tw, int tid, int nthr)
{
    const int q = m / R;
    const int nb = n / R;
    const int twstride = n / m;
    const int tmask = (1 << tshift) - 1;
    const int total = nb << tshift;
    for (int idx = tid; idx < total; idx += nthr) {
        const int t = idx & tmask;
        const int b = idx >> tshift;
        const int blk = b / q;
        const int j = b - blk * q;
        cplx *p = s + ((blk * m + j) * ld + t);          // 32-bit index: tiles are far below 2^31 elements
        const int step = q * ld;
        cplx v[R];
#pragma unroll
        for (int k = 0; k < R; ++k) v[k] = p[k * step];
        if (q > 1) {
            const int e1 = j * twstride;
#pragma unroll
            for (int k = 1; k < R; ++k) {
                cplx w = tw[e1 * k];
                v[k] = D < 0 ? cmul(v[k], w) : cmulc(v[k], w);
            }
        }
        bfly<R, D>(v);
#pragma unroll
        for (int k = 0; k < R; ++k) p[k * step] = v[k];
    }
}

// all stages, last to first; fast radices only (callers check the plan)
template <int D>
HD void fft_tile_t(cplx *s, int ld, int tshift, const FftPlanDev &pl, const cplx *tw, int tid, int nthr)
{
    int m = 1;
    for (int st = pl.nstages - 1; st >= 0; --st) {
        const int r = pl.radix[st];
        m *= r;
        switch (r) {
        case 2: fft_stage_t<2, D>(s, ld, tshift, pl.n, m, tw, tid, nthr); break;
        case 3: fft_stage_t<3, D>(s, ld, tshift, pl.n, m, tw, tid, nthr); break;
        case 4: fft_stage_t<4, D>(s, ld, tshift, pl.n, m, tw, tid, nthr); break;
        case 5: fft_stage_t<5, D>(s, ld, tshift, pl.n, m, tw, tid, nthr); break;
        case 7: fft_stage_t<7, D>(s, ld, tshift, pl.n, m, tw, tid, nthr); break;
#if !defined(ZF2_MAX_RADIX) || ZF2_MAX_RADIX >= 16
        case 8: fft_stage_t<8, D>(s, ld, tshift, pl.n, m, tw, tid, nthr); break;
        default: fft_stage_t<16, D>(s, ld, tshift, pl.n, m, tw, tid, nthr); break;
#else
        default: fft_stage_t<8, D>(s, ld, tshift, pl.n, m, tw, tid, nthr); break;
#endif
        }
        BLOCK_SYNC();
    }
}

// ---- whole transform; caller guarantees a BLOCK_SYNC() before entry, one is issued after
// every stage (so the tile is consistent on return) ----------------------------------------
template <int D>
HD void fft_tile(cplx *s, int ld, int tshift, const FftPlanDev &pl, const cplx *tw, int tid, int nthr)
{
    int m = pl.n;
    for (int st = 0; st < pl.nstages; ++st) {
        const int r = pl.radix[st];
        switch (r) {
        case 2: fft_stage<2, D>(s, ld, tshift, pl.n, m, tw, tid, nthr); break;
        case 3: fft_stage<3, D>(s, ld, tshift, pl.n, m, tw, tid, nthr); break;
        case 4: fft_stage<4, D>(s, ld, tshift, pl.n, m, tw, tid, nthr); break;
        case 5: fft_stage<5, D>(s, ld, tshift, pl.n, m, tw, tid, nthr); break;
        case 7: fft_stage<7, D>(s, ld, tshift, pl.n, m, tw, tid, nthr); break;
        case 8: fft_stage<8, D>(s, ld, tshift, pl.n, m, tw, tid, nthr); break;
        case 16: fft_stage<16, D>(s, ld, tshift, pl.n, m, tw, tid, nthr); break;
        default: fft_stage_generic<D>(s, ld, tshift, pl.n, m, r, tw, tid, nthr); break;
        }
        m /= r;
        BLOCK_SYNC();
    }
}
