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

template <int R, int D> HD void bfly(cplx *v);

// ---- composite radices: fewer shared-memory stages for lengths like 120 = 12*10, 196 = 14*14 ----
HD constexpr int mod_inverse(int a, int m)
{
    int r = 0;
    for (int x = 1; x < m; ++x) if ((a * x) % m == 1) r = x;
    return m == 1 ? 0 : r;
}

// R = R1*R2 with gcd(R1,R2) = 1, Good-Thomas: no twiddles between the two passes, only index maps
//   input  n = (R2*n1 + R1*n2) mod R,   output k = k1 (mod R1), k2 (mod R2)  (Chinese remainder)
template <int R1, int R2, int D> HD void bfly_pfa(cplx *v)
{
    constexpr int R = R1 * R2;
    constexpr int A = R2 * mod_inverse(R2 % R1, R1), B = R1 * mod_inverse(R1 % R2, R2);
    cplx t[R2][R1];
#pragma unroll
    for (int n2 = 0; n2 < R2; ++n2) {
#pragma unroll
        for (int n1 = 0; n1 < R1; ++n1) t[n2][n1] = v[(R2 * n1 + R1 * n2) % R];
        bfly<R1, D>(t[n2]);
    }
#pragma unroll
    for (int k1 = 0; k1 < R1; ++k1) {
        cplx c[R2];
#pragma unroll
        for (int n2 = 0; n2 < R2; ++n2) c[n2] = t[n2][k1];
        bfly<R2, D>(c);
#pragma unroll
        for (int k2 = 0; k2 < R2; ++k2) v[(k1 * A + k2 * B) % R] = c[k2];
    }
}

// radix 9 = 3 x 3: X[k1 + 3 k2] = sum_b W9^{b k1} W3^{b k2} sum_a x[3a + b] W3^{a k1}
template <int D> HD void bfly9(cplx *v)
{
    const double c1 = 0.766044443118978, s1 = 0.6427876096865393;      // cos, sin of 2 pi / 9
    const double c2 = 0.17364817766693036, s2 = 0.984807753012208;     // 4 pi / 9
    const double c4 = -0.9396926207859084, s4 = 0.3420201433256687;    // 8 pi / 9
    cplx t[3][3];                                     // t[b][k1]
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        cplx col[3] = {v[b], v[3 + b], v[6 + b]};
        bfly_odd<3, D>(col);
#pragma unroll
        for (int k1 = 0; k1 < 3; ++k1) t[b][k1] = col[k1];
    }
    t[1][1] = cmul_cD<D>(t[1][1], c1, s1);            // W9^1
    t[1][2] = cmul_cD<D>(t[1][2], c2, s2);            // W9^2
    t[2][1] = cmul_cD<D>(t[2][1], c2, s2);            // W9^2
    t[2][2] = cmul_cD<D>(t[2][2], c4, s4);            // W9^4
#pragma unroll
    for (int k1 = 0; k1 < 3; ++k1) {
        cplx row[3] = {t[0][k1], t[1][k1], t[2][k1]};
        bfly_odd<3, D>(row);
#pragma unroll
        for (int k2 = 0; k2 < 3; ++k2) v[k1 + 3 * k2] = row[k2];
    }
}

template <int R, int D> HD void bfly(cplx *v)
{
    if (R == 2) bfly2<D>(v);
    else if (R == 3) bfly_odd<3, D>(v);
    else if (R == 4) bfly4<D>(v);
    else if (R == 5) bfly_odd<5, D>(v);
    else if (R == 6) bfly_pfa<2, 3, D>(v);
    else if (R == 7) bfly_odd<7, D>(v);
    else if (R == 8) bfly8<D>(v);
    else if (R == 9) bfly9<D>(v);
    else if (R == 10) bfly_pfa<2, 5, D>(v);
    else if (R == 12) bfly_pfa<4, 3, D>(v);
    else if (R == 14) bfly_pfa<2, 7, D>(v);
    else if (R == 15) bfly_pfa<3, 5, D>(v);
    else if (R == 16) bfly16<D>(v);
}

// the radices with a register butterfly above ("fast" radices); every other radix (primes 11..31)
// goes through fft_stage_generic
HD bool fast_radix(int r)
{
    return (r >= 2 && r <= 10) || r == 12 || r == 14 || r == 15 || r == 16;
}

// runtime radix -> template instantiation; CALL(R) is a statement that uses the constant R
#define DBSPM_FAST_RADIX_SWITCH(r, CALL)                                                            \
    switch (r) {                                                                                    \
    case 2: CALL(2); break;   case 3: CALL(3); break;   case 4: CALL(4); break;   case 5: CALL(5); break;   \
    case 6: CALL(6); break;   case 7: CALL(7); break;   case 8: CALL(8); break;   case 9: CALL(9); break;   \
    case 10: CALL(10); break; case 12: CALL(12); break; case 14: CALL(14); break; case 15: CALL(15); break; \
    default: CALL(16); break;                                                                       \
    }

// the same with the instantiations above a compile-time cap left out: a kernel built for plans of radix <= MAXR does not
// carry the registers of the radix-16 butterfly (ptxas allocates for the largest path, taken or not)
#define DBSPM_RADIX_SWITCH_MAX(r, CALL, MAXR)                                                                   \
    switch (r) {                                                                                                \
    case 2: CALL(2); break;   case 3: CALL(3); break;   case 4: CALL(4); break;   case 5: CALL(5); break;       \
    case 6: CALL(6); break;   case 7: CALL(7); break;   case 8: CALL(8); break;                                 \
    case 9: if (MAXR >= 9) CALL(9); break;    case 10: if (MAXR >= 10) CALL(10); break;                         \
    case 12: if (MAXR >= 12) CALL(12); break; case 14: if (MAXR >= 14) CALL(14); break;                         \
    case 15: if (MAXR >= 15) CALL(15); break;                                                                   \
    default: if (MAXR >= 16) CALL(16); break;                                                                   \
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
template <int D, int MAXR = 16>
HD void fft_tile_t(cplx *s, int ld, int tshift, const FftPlanDev &pl, const cplx *tw, int tid, int nthr)
{
    int m = 1;
    for (int st = pl.nstages - 1; st >= 0; --st) {
        const int r = pl.radix[st];
        m *= r;
#define CALL_(R) fft_stage_t<R, D>(s, ld, tshift, pl.n, m, tw, tid, nthr)
        DBSPM_RADIX_SWITCH_MAX(r, CALL_, MAXR)
#undef CALL_
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
        if (fast_radix(r)) {
#define CALL_(R) fft_stage<R, D>(s, ld, tshift, pl.n, m, tw, tid, nthr)
            DBSPM_FAST_RADIX_SWITCH(r, CALL_)
#undef CALL_
        } else {
            fft_stage_generic<D>(s, ld, tshift, pl.n, m, r, tw, tid, nthr);
        }
        m /= r;
        BLOCK_SYNC();
    }
}
