// corr_body.h -- block bodies of the sr/es correlation kernels.
//
// Replaces pydbspm/interactions.py:9-22 (get_density_overlap) together with the `**ALPHA`
// and the window slice at dbspm:329-330,348-349 of the reference.
//
// Data model.  A real (Nx,Ny,Nz) array, z fastest, Nz even, is viewed
// as a complex (Nx,Ny,Nz/2) array c[x,y,zp] = a[x,y,2zp] + i a[x,y,2zp+1] -- the bytes are
// not moved.  Forward: complex FFTs along x, then y (axis_pass_body), then along zp fused
// with the even/odd untangle, the spectral product A.conj(B), the window phase, and the
// inverse z transform restricted to the output window (zfused_body), which emits the
// window already packed as complex pairs of planes.  Inverse: complex FFTs along y, then x
// of the packed window; its bytes are then the real (Nx,Ny,Nzw) result.
#pragma once
#include <string.h>
#include "fft_core.h"

#ifndef AXIS2_SWITCH_MAXR
#define AXIS2_SWITCH_MAXR 16   /* largest radix instantiated in the one-shot axis pass (experiment builds: 8) */
#endif
#ifndef AXIS2_TILE_BUDGET
#define AXIS2_TILE_BUDGET (220 * 1024 / 3)   /* shared-memory bytes per CTA the axis-pass tile may use */
#endif
#ifndef ZF2_MAX_RADIX
#define ZF2_MAX_RADIX 16   /* power-of-two radix cap of the z-kernel plan (registers vs stage count) */
#endif

// ------------------------------------------------------------------------------------------
// Strided-axis pass: FFT along an axis of length n whose elements are `inner` complex apart;
// `outer` independent slabs.  One block = T adjacent lines of one slab, staged in shared
// memory as s[n][T]; global rows are T*16-byte contiguous segments (128 B for T = 8).
// blockIdx.y selects one of two arrays (A and B travel in one launch).
// ------------------------------------------------------------------------------------------
#define DBSPM_MAX_PEERS 8
#define DIST_PEER_SHIFT 20
struct AxisParams {
    const cplx *in[2];
    cplx *out[2];
    double alpha[2];          // applied on load when use_pow (x pass): v -> v^alpha componentwise
    int use_pow;
    int narrays;
    int n;
    int tshift;               // T = 1 << tshift
    long long inner;
    long long outer;
    long long tiles;          // ceil(inner / T)
    FftPlanDev plan;
    const cplx *tw;           // device tables for length n
    const int *pos_of;        // frequency -> in-place position
    const int *freq_of;       // in-place position -> frequency
    int v2;                   // 1: register-first/last body (axis_pass_body_v2)
    int tw_smem;              // v2: stage the twiddle table in shared memory (else read it through L1)
    // z-pruned correlation (v2 body, first pass only): the pass reads a z-range of the caller's arrays
    // instead of a packed copy.  Destination line (y, zp) with zp < g_valid[w] comes from source column
    // y*g_src_nzh + (g_zp0[w] + zp) mod g_src_nzh, row stride g_src_inner; zp >= g_valid[w] reads as zero.
    // Slab o of the source starts at o * n * g_src_inner (the distributed y pass: n = ny, lines are zp only).
    int gather;
    long long g_src_inner;
    int g_dst_nzh, g_src_nzh;
    int g_zp0[2], g_valid[2];
    // distributed correlation (v2 body, MAP variant): row K of the transform axis lives at row in_map[K] /
    // goes to row out_map[K] (units of `inner` elements) and slab o starts at o * in_ostride / o * out_ostride
    // elements; null maps = natural rows.  Without maps the strides are n * inner.
    const int *in_map, *out_map;
    long long in_ostride, out_ostride;
    // MAP == 2 (fused exchange): out_map[K] = (h << DIST_PEER_SHIFT) | row, and the row is stored straight into rank h's
    // receive buffer out_peer[which][h] (peer memory over NVLink) at element out_base_off + o * out_ostride + row * inner
    cplx *out_peer[2][DBSPM_MAX_PEERS];
    long long out_base_off;
    // > 0: every block first asks the L2 for the rows of the tile that block `bid + pf_ahead` will read (the block that
    // takes this SM slot next), so that its first-stage loads hit the L2 instead of waiting on DRAM
    int pf_ahead;
    // >= 0: byte offset in dynamic shared memory of the staged pow_fast table (powered first pass); < 0: pow_nonneg
    int pow_tab_off;
    double pow_lo[2], pow_hi[2];   // pow_fast_bounds(alpha[w])
    // pipelined body (axis_pass_body_v2s): byte offset of the per-thread staging slots (8 x nthr x 16 bytes) of the rows that
    // travel to the next first stage through cp.async instead of registers
    int stage_off;
};

// ---- distributed correlation: slot order of a transform axis --------------------------------------
// The x-slab decomposition hands each of G ranks n/G frequencies of the y axis.  The fused
// z kernel needs the frequencies ky and -ky on the same rank, so frequencies are dealt in "slot" order
// 0, n/2, 1, n-1, 2, n-2, ...: slot(0) = 0, slot(n/2) = 1, slot(k) = 2k, slot(n-k) = 2k+1 (0 < k < n/2); the
// mirror of slot S is S itself for S < 2 and S^1 otherwise.  Rank h owns slots [h*n/G, (h+1)*n/G), n/G even.
HD int dist_slot_of(int K, int n)
{
    if (K == 0) return 0;
    if (2 * K == n) return 1;
    return 2 * K < n ? 2 * K : 2 * (n - K) + 1;
}

HD int dist_mirror_slot(int S) { return S < 2 ? S : (S ^ 1); }

// row (units of `inner`) of frequency K inside a buffer laid out (G, rows_per_rank_block / (n/G), n/G, inner)
HD int dist_row_of(int K, int n, int G, long long blockrows)
{
    const int nl = n / G, S = dist_slot_of(K, n);
    return (int)((long long)(S / nl) * blockrows + (S % nl));
}

// fused-exchange table entry of frequency K: (owner rank << DIST_PEER_SHIFT) | row inside the owner's block
HD int dist_peer_entry(int K, int n, int G)
{
    const int nl = n / G, S = dist_slot_of(K, n);
    return ((S / nl) << DIST_PEER_SHIFT) | (S % nl);
}

// fused second exchange: position X of the x axis belongs to the rank that owns its x-slab
HD int dist_block_entry(int X, int n, int G)
{
    const int nl = n / G;
    return ((X / nl) << DIST_PEER_SHIFT) | (X % nl);
}

HD bool dist_shape_ok(int nx, int ny, int nz, int G)
{
    return G >= 1 && nx >= 1 && ny >= 2 && nz >= 2 && (nz % 2) == 0 && (ny % (2 * G)) == 0 && (nx % G) == 0;
}

// ---- z-pruned correlation: geometry ---------------------------------------------------------
// When the tip array B is exactly zero outside the planes (s_lo + m) mod nz, m in [0, s_len), the window
// planes [z_lo, z_hi) of the periodic correlation only see the planes z_lo + s_lo ... z_hi - 1 + s_lo +
// s_len - 1 of A.  Both ranges are cut out (starting on even planes, the packed transform pairs z = 2k,
// 2k+1), zero-padded to a common FFT-friendly length Mp >= their linear-correlation length, and
// correlated periodically over Mp: no wrap-around term reaches the window, which sits at [w_lo, w_lo+nzw).
struct PrunedGeom {
    int use;                  // 0: no gain (support too wide) -> full-length path
    int Mp;                   // pruned z length (even, Mp/2 is 2-3-5-7 smooth)
    int zA0, zB0;             // first source plane of the A / B cut (even)
    int validA, validB;       // planes taken from the source (even); the rest of Mp is zero
    int w_lo;                 // window start inside the pruned problem (0 or 1)
};

HD bool pruned_smooth(int m)
{
    const int pr[4] = {2, 3, 5, 7};
    for (int i = 0; i < 4; ++i) while (m % pr[i] == 0) m /= pr[i];
    return m == 1;
}

HD PrunedGeom pruned_geometry(int nz, int z_lo, int z_hi, int s_lo, int s_len)
{
    PrunedGeom g;
    g.use = 0; g.Mp = nz; g.zA0 = g.zB0 = 0; g.validA = g.validB = nz; g.w_lo = z_lo;
    if (s_len < 1 || s_len >= nz || (nz & 1)) return g;
    s_lo = ((s_lo % nz) + nz) % nz;
    const int nzw = z_hi - z_lo;
    g.zB0 = s_lo & ~1;
    const int Lb = s_lo - g.zB0 + s_len;
    const int a0 = (z_lo + g.zB0) % nz;
    g.zA0 = a0 & ~1;
    g.w_lo = a0 - g.zA0;
    const int Ma = g.w_lo + nzw + Lb - 1;
    g.validB = (Lb + 1) & ~1;
    g.validA = (Ma + 1) & ~1;
    int Mp = g.validA > g.validB ? g.validA : g.validB;
    while (!pruned_smooth(Mp / 2)) Mp += 2;
    g.Mp = Mp;
    g.use = (Mp < nz && 10 * (long long)Mp <= 9 * (long long)nz) ? 1 : 0;     // worth it only below 90 % of the full length
    return g;
}

HD size_t axis_smem_bytes(int n, int tshift)
{
    return (size_t)n * 16 + (size_t)n * 4 + ((size_t)n << tshift) * 16;
}

// v^alpha as numpy's `**` gives it for the cases a density array can hold: rho >= 0 upstream (grid.py:450-464
// filter_neg) -> exp(alpha log v), 0 -> exp(-inf) = 0; a negative value (unfiltered noise) -> NaN for a fractional
// alpha, like numpy, and the signed finite power for an integer-valued alpha (numpy: (-2.0)**2.0 = 4.0)
HD double pow_signed(double v, double alpha)
{
    if (v < 0.0 && alpha == trunc(alpha) && fabs(alpha) < 9007199254740992.0) {
        const double r = exp(alpha * log(-v));
        return (fmod(alpha, 2.0) != 0.0) ? -r : r;
    }
    return exp(alpha * log(v));
}

HD double pow_nonneg(double v, double alpha) { return pow_signed(v, alpha); }

// ---- table-driven v^alpha: 21 FP64 operations instead of the 46 of exp(alpha * log(v)) ----------------------
// The powered first pass is FP64-issue bound, so the operation count is what matters.
//   log2 v = e + t_i + log2(1 + r),  r = m * invc_i - 1  (|r| <= 2^-6; i = top 5 mantissa bits), degree-7 Taylor
//   2^y    = 2^n * E_j * 2^t,        y = n + j/32 + t    (|t| <= 2^-6), degree-6 Taylor
// Relative error <= 1e-14 for 1e-12 <= v <= 100 and <= 2e-13 over the whole double range (it grows with |log2 v|;
// tests/test_emulation.py::test_pow_fast_accuracy): four orders below the 1e-9 the sr grid is judged at.  Zero, negative, subnormal, non-finite inputs and results outside 2^+-1000 take pow_nonneg.
#include "pow_tables.h"
static const double dbspm_pow_table_h[DBSPM_POW_TABLE_LEN] = DBSPM_POW_TABLE;
#if defined(__CUDACC__)
static __device__ const double dbspm_pow_table_d[DBSPM_POW_TABLE_LEN] = DBSPM_POW_TABLE;
#endif

HD long long dbspm_d2ll(double v)
{
#if defined(__CUDA_ARCH__)
    return __double_as_longlong(v);
#else
    long long b; memcpy(&b, &v, sizeof b); return b;
#endif
}

HD double dbspm_ll2d(long long b)
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double(b);
#else
    double v; memcpy(&v, &b, sizeof v); return v;
#endif
}

// out-of-line cold path (zero, negative, subnormal, huge): keeps the unrolled first stage small
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#else
__attribute__((noinline))
#endif
static double pow_cold(double v, double alpha) { return pow_signed(v, alpha); }

// inputs for which pow_fast_core is valid: normal numbers whose result stays inside 2^+-1000
HD void pow_fast_bounds(double alpha, double *lo, double *hi)
{
    const double a = fabs(alpha) > 1.0 ? fabs(alpha) : 1.0;
    *lo = exp2(-1000.0 / a);
    *hi = exp2(1000.0 / a);
}

// Table addressing: entry e of copy c sits at tab[e * pstride + c].  On the device the block stages POW_TAB_COPIES = 16
// copies (pstride = 16) and a lane reads copy (lane & 15): a 64-bit shared-memory load is served per half-warp, and the 16
// lanes of a half-warp then hit 16 different bank pairs whatever entries they ask for.  With one copy the 32 random
// entries of a warp collided two- to four-fold (ncu, round 1: 43 % of the pass's shared-memory wavefronts were conflicts).
#ifndef POW_TAB_COPIES
#define POW_TAB_COPIES 16
#endif
HD double pow_fast_core(double v, double alpha, const double *tab, int pstride, int pcopy);
HD double pow_fast(double v, double alpha, const double *tab)
{
    double lo, hi;
    pow_fast_bounds(alpha, &lo, &hi);
    return (v >= lo && v <= hi) ? pow_fast_core(v, alpha, tab, 1, 0) : pow_cold(v, alpha);
}

HD double pow_fast_core(double v, double alpha, const double *tab, int pstride, int pcopy)
{
    const long long b = dbspm_d2ll(v);
    const int hi = (int)(b >> 32);
    const int i = (hi >> 15) & 31;
    const double m = dbspm_ll2d((b & 0x000FFFFFFFFFFFFFLL) | 0x3FF0000000000000LL);
    // exponent as a double without a conversion instruction: (1.5 * 2^52 + e) - 1.5 * 2^52
    const double ed = dbspm_ll2d(0x4338000000000000LL + (long long)((hi >> 20) - 1023)) - 6755399441055744.0;
    const double invc = tab[(2 * i) * pstride + pcopy], ti = tab[(2 * i + 1) * pstride + pcopy];
    const double r = fma(m, invc, -1.0);
    double p = 0.2060992915555662;
    p = fma(p, r, -0.2404491734814939);
    p = fma(p, r, 0.28853900817779266);
    p = fma(p, r, -0.36067376022224085);
    p = fma(p, r, 0.4808983469629878);
    p = fma(p, r, -0.7213475204444817);
    p = fma(p, r, 1.4426950408889634);
    const double y = alpha * fma(p, r, ti + ed);
    const double shifter = 211106232532992.0;                     // 1.5 * 2^47: one ulp = 2^-5
    const double z = y + shifter;
    const int ki = (int)(unsigned int)(dbspm_d2ll(z) & 0xFFFFFFFFLL);   // round(32 y), two's complement
    const double t = y - (z - shifter);
    double q = 0.0001540353039338161;
    q = fma(q, t, 0.0013333558146428443);
    q = fma(q, t, 0.009618129107628477);
    q = fma(q, t, 0.05550410866482158);
    q = fma(q, t, 0.24022650695910072);
    q = fma(q, t, 0.6931471805599453);
    q = q * t;
    const double E = tab[(64 + (ki & 31)) * pstride + pcopy];
    const double res = fma(E, q, E);
    return dbspm_ll2d(dbspm_d2ll(res) + ((long long)(ki >> 5) << 52));
}

template <int D>
HD void axis_pass_body(const AxisParams &P, unsigned char *smem_raw, long long bid, int which, int tid, int nthr)
{
    const int n = P.n, T = 1 << P.tshift, tmask = T - 1;
    cplx *tw = (cplx *)smem_raw;
    cplx *s = tw + n;
    int *pos_of = (int *)(s + ((size_t)n << P.tshift));

    const long long o = bid / P.tiles;
    const long long tile = bid - o * P.tiles;
    const long long i0 = tile << P.tshift;
    const int tcount = (int)((P.inner - i0) < T ? (P.inner - i0) : T);
    const cplx *gin = P.in[which] + (size_t)o * n * P.inner + i0;
    cplx *gout = P.out[which] + (size_t)o * n * P.inner + i0;
    const double alpha = P.alpha[which];
    const bool do_pow = P.use_pow && alpha != 1.0;

    for (int e = tid; e < n; e += nthr) {
        tw[e] = ldg_c(P.tw + e);
        pos_of[e] = LDG(P.pos_of + e);
    }
    const int total = n << P.tshift;
    for (int idx = tid; idx < total; idx += nthr) {
        const int t = idx & tmask, row = idx >> P.tshift;
        cplx v = cmake(0.0, 0.0);
        if (t < tcount) {
            v = gin[(size_t)row * P.inner + t];
            if (do_pow) { v.x = pow_nonneg(v.x, alpha); v.y = pow_nonneg(v.y, alpha); }
        }
        s[idx] = v;
    }
    BLOCK_SYNC();
    fft_tile<D>(s, T, P.tshift, P.plan, tw, tid, nthr);
    for (int idx = tid; idx < total; idx += nthr) {
        const int t = idx & tmask, K = idx >> P.tshift;
        if (t < tcount) gout[(size_t)K * P.inner + t] = s[((size_t)pos_of[K] << P.tshift) + t];
    }
}

// ------------------------------------------------------------------------------------------
// v2 of the strided-axis pass ("register first, register last").  The first radix stage reads
// its r rows straight from global memory into registers (r independent 128-bit streaming loads
// in flight per thread), applies the optional power, the butterfly and the twiddles, and only
// then touches shared memory; the last stage reads shared memory, does its butterfly and
// streams the rows to their frequency positions in global memory.  Shared-memory traffic per
// element drops from 2*(stages+1) to 2*(stages-1) 16-byte accesses and two block barriers
// disappear.  Used when every radix of the plan is one of {2,3,4,5,7,8,16}.
// ------------------------------------------------------------------------------------------
HD cplx ld_stream(const cplx *p)
{
#if defined(__CUDA_ARCH__)
    const double2 v = __ldcs(reinterpret_cast<const double2 *>(p));
    return cmake(v.x, v.y);
#else
    return *p;
#endif
}

HD void st_stream(cplx *p, cplx v)
{
#if defined(__CUDA_ARCH__)
    __stcs(reinterpret_cast<double2 *>(p), make_double2(v.x, v.y));
#else
    *p = v;
#endif
}

HD bool axis_v2_ok(const FftPlanDev &pl)
{
    if (pl.nstages < 1) return false;
    for (int i = 0; i < pl.nstages; ++i) {
        const int r = pl.radix[i];
        if (!fast_radix(r)) return false;
    }
    return true;
}

// shared memory of the v2 body: [twiddles n*16 if tw_smem] [tile n*T*16 if nstages > 1]
HD size_t axis2_smem_bytes(int n, int tshift, int nstages, int tw_smem)
{
    return (tw_smem ? (size_t)n * 16 : 0) + (nstages > 1 ? ((size_t)n << tshift) * 16 : 0);
}

// launch geometry of the v2 body: widest tile (8,4,2,1 lines) that still leaves room for three
// resident CTAs per SM; the twiddle table is staged in shared memory only if that still holds
HD void axis2_config(int n, long long inner, int nstages, int *tshift, int *tw_smem)
{
    const size_t budget = AXIS2_TILE_BUDGET;
    int ts = 3;
    while (ts > 0 && ((1LL << ts) > 2 * inner || axis2_smem_bytes(n, ts, nstages, 0) > budget)) --ts;
    *tshift = ts;
    *tw_smem = axis2_smem_bytes(n, ts, nstages, 1) <= budget ? 1 : 0;
}

template <int D, int R, int MAP>
HD void axis2_first(const AxisParams &P, cplx *s, const cplx *tw, const cplx *gin, cplx *gout, int tcount,
                    bool do_pow, double alpha, const double *ptab, int which, long long i0, size_t gbase, int tid, int nthr)
{
    const int n = P.n, q = n / R, tmask = (1 << P.tshift) - 1;
    const bool single = P.plan.nstages == 1;
    const int total = q << P.tshift;
    for (int idx = tid; idx < total; idx += nthr) {
        const int t = idx & tmask, j = idx >> P.tshift;
        cplx v[R];
        bool have = t < tcount;
        const cplx *src = gin + (size_t)j * P.inner + t;
        size_t rstride = (size_t)P.inner;
        if (P.gather && have) {
            const long long i = i0 + t, y = i / P.g_dst_nzh;
            const int zp = (int)(i - y * P.g_dst_nzh);
            int zs = P.g_zp0[which] + zp;
            zs = zs >= P.g_src_nzh ? zs - P.g_src_nzh : zs;
            have = zp < P.g_valid[which];
            rstride = (size_t)P.g_src_inner;
            src = P.in[which] + gbase + ((size_t)y * P.g_src_nzh + zs) + (size_t)j * rstride;
        }
        if (have) {
            if (MAP && P.in_map) {
#pragma unroll
                for (int k = 0; k < R; ++k) v[k] = ld_stream(gin + (size_t)LDG(P.in_map + j + k * q) * rstride + t);
            } else {
#pragma unroll
                for (int k = 0; k < R; ++k) v[k] = ld_stream(src + (size_t)k * q * rstride);
            }
            if (do_pow) {
                if (ptab) {
                    const double plo = P.pow_lo[which], phi = P.pow_hi[which];
#if defined(__CUDA_ARCH__)
                    const int pstride = POW_TAB_COPIES, pcopy = tid & (POW_TAB_COPIES - 1);
#else
                    const int pstride = 1, pcopy = 0;
#endif
#pragma unroll
                    for (int k = 0; k < R; ++k) {
                        v[k].x = (v[k].x >= plo && v[k].x <= phi) ? pow_fast_core(v[k].x, alpha, ptab, pstride, pcopy) : pow_cold(v[k].x, alpha);
                        v[k].y = (v[k].y >= plo && v[k].y <= phi) ? pow_fast_core(v[k].y, alpha, ptab, pstride, pcopy) : pow_cold(v[k].y, alpha);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < R; ++k) { v[k].x = pow_cold(v[k].x, alpha); v[k].y = pow_cold(v[k].y, alpha); }
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < R; ++k) v[k] = cmake(0.0, 0.0);
        }
        bfly<R, D>(v);
        if (single) {
            if (t < tcount) {
#pragma unroll
                for (int k = 0; k < R; ++k) {
                    if (MAP == 2) {
                        const int m = LDG(P.out_map + k);
                        st_stream(P.out_peer[which][m >> DIST_PEER_SHIFT] + (gout - P.out[which]) +
                                  (size_t)(m & ((1 << DIST_PEER_SHIFT) - 1)) * P.inner + t, v[k]);
                    } else {
                        st_stream(gout + (size_t)((MAP && P.out_map) ? LDG(P.out_map + k) : k) * P.inner + t, v[k]);
                    }
                }
            }
        } else {
#pragma unroll
            for (int k = 1; k < R; ++k) {
                const cplx w = tw[j * k];
                v[k] = D < 0 ? cmul(v[k], w) : cmulc(v[k], w);
            }
            cplx *dst = s + ((j << P.tshift) + t);
#pragma unroll
            for (int k = 0; k < R; ++k) dst[(k * q) << P.tshift] = v[k];
        }
    }
}

template <int D, int R, int MAP>
HD void axis2_last(const AxisParams &P, const cplx *s, cplx *gout, int tcount, int which, int tid, int nthr)
{
    const int n = P.n, nb = n / R, tmask = (1 << P.tshift) - 1;
    const int total = nb << P.tshift;
    for (int idx = tid; idx < total; idx += nthr) {
        const int t = idx & tmask, blk = idx >> P.tshift;
        const cplx *src = s + (((blk * R) << P.tshift) + t);
        cplx v[R];
#pragma unroll
        for (int k = 0; k < R; ++k) v[k] = src[k << P.tshift];
        bfly<R, D>(v);
        if (t < tcount) {
            const int K0 = LDG(P.freq_of + blk * R);           // frequency held by position blk*R
            if (MAP == 2) {
                const size_t goff = (size_t)(gout - P.out[which]) + t;
#pragma unroll
                for (int k = 0; k < R; ++k) {
                    const int m = LDG(P.out_map + K0 + k * nb);
                    st_stream(P.out_peer[which][m >> DIST_PEER_SHIFT] + goff + (size_t)(m & ((1 << DIST_PEER_SHIFT) - 1)) * P.inner, v[k]);
                }
            } else if (MAP && P.out_map) {
#pragma unroll
                for (int k = 0; k < R; ++k) st_stream(gout + (size_t)LDG(P.out_map + K0 + k * nb) * P.inner + t, v[k]);
            } else {
                cplx *dst = gout + (size_t)K0 * P.inner + t;
#pragma unroll
                for (int k = 0; k < R; ++k) st_stream(dst + (size_t)k * nb * P.inner, v[k]);
            }
        }
    }
}

template <int D, int MAP = 0>
HD void axis_pass_body_v2(const AxisParams &P, unsigned char *smem_raw, long long bid, int which, int tid, int nthr)
{
    const int n = P.n, T = 1 << P.tshift;
    const FftPlanDev &pl = P.plan;
    cplx *twbuf = (cplx *)smem_raw;
    cplx *s = P.tw_smem ? twbuf + n : twbuf;
    const cplx *tw = P.tw_smem ? (const cplx *)twbuf : P.tw;

    const long long o = bid / P.tiles;
    const long long tile = bid - o * P.tiles;
    const long long i0 = tile << P.tshift;
    const int tcount = (int)((P.inner - i0) < T ? (P.inner - i0) : T);
    const cplx *gin = P.in[which] + (MAP ? (size_t)o * P.in_ostride : (size_t)o * n * P.inner) + i0;
    cplx *gout = P.out[which] + (MAP == 2 ? (size_t)P.out_base_off : 0) + (MAP ? (size_t)o * P.out_ostride : (size_t)o * n * P.inner) + i0;
    const double alpha = P.alpha[which];
    const bool do_pow = P.use_pow && alpha != 1.0;

#if defined(__CUDA_ARCH__)
    if (P.pf_ahead > 0 && !P.gather && !(MAP && P.in_map)) {
        const long long nb = bid + P.pf_ahead;
        if (nb < P.outer * P.tiles) {
            const long long o2 = nb / P.tiles, t2 = nb - o2 * P.tiles;
            const cplx *g2 = P.in[which] + (MAP ? (size_t)o2 * P.in_ostride : (size_t)o2 * n * P.inner) + (t2 << P.tshift);
            for (int r = tid; r < n; r += nthr) asm volatile("prefetch.global.L2 [%0];" ::"l"(g2 + (size_t)r * P.inner));
        }
    }
#endif
    const double *ptab = nullptr;
    if (do_pow && P.pow_tab_off >= 0) {
#if defined(__CUDA_ARCH__)
        double *pt = (double *)(smem_raw + P.pow_tab_off);
        for (int e = tid; e < DBSPM_POW_TABLE_LEN * POW_TAB_COPIES; e += nthr) pt[e] = dbspm_pow_table_d[e / POW_TAB_COPIES];
        ptab = pt;
#else
        ptab = dbspm_pow_table_h;
#endif
    }
    if ((P.tw_smem && pl.nstages > 1) || ptab) {
        if (P.tw_smem && pl.nstages > 1)
            for (int e = tid; e < n; e += nthr) twbuf[e] = ldg_c(P.tw + e);
        BLOCK_SYNC();
    }
    // z-range cut (gather): slab o of the source starts o * n source rows further (x pass: one slab; distributed y pass:
    // one slab per x of the rank's x-slab)
    const size_t gbase = P.gather ? (size_t)o * n * (size_t)P.g_src_inner : 0;
#define CALL_(R) axis2_first<D, R, MAP>(P, s, tw, gin, gout, tcount, do_pow, alpha, ptab, which, i0, gbase, tid, nthr)
    DBSPM_RADIX_SWITCH_MAX(pl.radix[0], CALL_, AXIS2_SWITCH_MAXR)
#undef CALL_
    if (pl.nstages == 1) return;
    BLOCK_SYNC();
    int m = n / pl.radix[0];
    for (int st = 1; st + 1 < pl.nstages; ++st) {
        const int r = pl.radix[st];
#define CALL_(R) fft_stage<R, D>(s, T, P.tshift, n, m, tw, tid, nthr)
        DBSPM_RADIX_SWITCH_MAX(r, CALL_, AXIS2_SWITCH_MAXR)
#undef CALL_
        m /= r;
        BLOCK_SYNC();
    }
#define CALL_(R) axis2_last<D, R, MAP>(P, s, gout, tcount, which, tid, nthr)
    DBSPM_RADIX_SWITCH_MAX(pl.radix[pl.nstages - 1], CALL_, AXIS2_SWITCH_MAXR)
#undef CALL_
}

// ------------------------------------------------------------------------------------------
// v2s: the register-first / register-last pass as a PERSISTENT, software-pipelined block.  A block walks over the tiles
// b0, b0 + bstride, ...; as soon as the first stage of tile i has left the registers for shared memory, the rows of
// tile i + 1 are requested (half of a thread's R rows as cp.async copies into its own shared-memory slots, half as 128-bit
// streaming loads into registers) and stay in flight while the middle stages and the storing last stage of tile i run.  The one-shot body above exposes the whole HBM latency of a
// tile at the start of every block (two resident CTAs per SM cannot hide it: ncu showed 45-55 % of DRAM peak with the
// issue slots 30 % busy); here a CTA always has its next 64 KB on the way.  Same arithmetic in the same order as
// axis_pass_body_v2 (bit-identical results).  Plain passes only (no row map, no z-range gather).
// The register array has the largest radix; every index is a compile-time constant after unrolling.
// ------------------------------------------------------------------------------------------
HD void axis2s_tile(const AxisParams &P, long long b, int which, const cplx **gin, cplx **gout, int *tcount)
{
    const int T = 1 << P.tshift;
    const long long o = b / P.tiles, tile = b - o * P.tiles, i0 = tile << P.tshift;
    *tcount = (int)((P.inner - i0) < T ? (P.inner - i0) : T);
    *gin = P.in[which] + (size_t)o * P.n * P.inner + i0;
    *gout = P.out[which] + (size_t)o * P.n * P.inner + i0;
}

HD void axis2s_load(const cplx *src, size_t rstride, int R, bool have, cplx *v)
{
#pragma unroll
    for (int k = 0; k < 16; ++k)
        if (k < R) v[k] = have ? ld_stream(src + (size_t)k * rstride) : cmake(0.0, 0.0);
}

// Request of the carried item of the NEXT tile.  Rows 0..7 of the butterfly go through cp.async (16 bytes each, L2 ->
// shared memory, no register) into the thread's own staging slots stg[k * nthr + tid]; rows 8..15 into the registers hi[].
// 64 registers of rows in flight did not survive the later stages (ptxas spilled them, i.e. waited for them at once);
// 32 do.  A thread only ever reads the slots it filled itself, so cp.async.wait_group is all the synchronisation needed.
HD void axis2s_request(const cplx *src, size_t rstride, int R, bool have, cplx *stg, int tid, int nthr, cplx *hi)
{
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (k < R) {
            cplx *dst = stg + (k * nthr + tid);
#if defined(__CUDA_ARCH__)
            if (have) {
                const unsigned sa = (unsigned)__cvta_generic_to_shared(dst);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(src + (size_t)k * rstride) : "memory");
            } else {
                *dst = cmake(0.0, 0.0);
            }
#else
            *dst = have ? src[(size_t)k * rstride] : cmake(0.0, 0.0);
#endif
        }
    }
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (8 + k < R) hi[k] = have ? ld_stream(src + (size_t)(8 + k) * rstride) : cmake(0.0, 0.0);
}

HD void axis2s_collect(const cplx *stg, int R, int tid, int nthr, const cplx *hi, cplx *v)
{
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (k < R) v[k] = stg[k * nthr + tid];
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (8 + k < R) v[8 + k] = hi[k];
}

// first stage of one item (t, j) whose R rows are already in v: power, butterfly, twiddles, store to the tile
template <int D, int R>
HD void axis2s_stage1(const AxisParams &P, cplx *s, const cplx *tw, cplx *v, bool have, bool do_pow, double alpha,
                      const double *ptab, int which, int q, int j, int t, int tid)
{
    if (do_pow && have) {
        if (ptab) {
            const double plo = P.pow_lo[which], phi = P.pow_hi[which];
#if defined(__CUDA_ARCH__)
            const int pstride = POW_TAB_COPIES, pcopy = tid & (POW_TAB_COPIES - 1);
#else
            const int pstride = 1, pcopy = 0;
#endif
#pragma unroll
            for (int k = 0; k < R; ++k) {
                v[k].x = (v[k].x >= plo && v[k].x <= phi) ? pow_fast_core(v[k].x, alpha, ptab, pstride, pcopy) : pow_cold(v[k].x, alpha);
                v[k].y = (v[k].y >= plo && v[k].y <= phi) ? pow_fast_core(v[k].y, alpha, ptab, pstride, pcopy) : pow_cold(v[k].y, alpha);
            }
        } else {
#pragma unroll
            for (int k = 0; k < R; ++k) { v[k].x = pow_cold(v[k].x, alpha); v[k].y = pow_cold(v[k].y, alpha); }
        }
    }
    bfly<R, D>(v);
#pragma unroll
    for (int k = 1; k < R; ++k) {
        const cplx w = tw[j * k];
        v[k] = D < 0 ? cmul(v[k], w) : cmulc(v[k], w);
    }
    cplx *dst = s + ((j << P.tshift) + t);
#pragma unroll
    for (int k = 0; k < R; ++k) dst[(k * q) << P.tshift] = v[k];
}

// The body is compiled for ONE plan shape at a time -- first radix R0, middle stages of radix RM, last radix RL, power or
// not -- because only a lean instruction stream keeps the rows in flight in registers: with the run-time radix
// switches of the one-shot body around them ptxas spilled them right behind the loads.
//   <16, 8, 8>: n = 128 (16*8), 1024 (16*8*8);  <16, 8, 4>: n = 64 (16*4), 512 (16*8*4)
template <int R0, int RM, int RL>
HD bool axis2s_plan_is(const FftPlanDev &pl)
{
    if (pl.nstages < 2 || pl.radix[0] != R0 || pl.radix[pl.nstages - 1] != RL) return false;
    for (int i = 1; i + 1 < pl.nstages; ++i)
        if (pl.radix[i] != RM) return false;
    return true;
}
HD bool axis2s_ok(const FftPlanDev &pl) { return axis2s_plan_is<16, 8, 8>(pl) || axis2s_plan_is<16, 8, 4>(pl); }

template <int D, int R0, int RM, int RL, bool POW>
HD void axis_pass_body_v2s(const AxisParams &P, unsigned char *smem_raw, long long b0, long long bstride, int which, int tid, int nthr)
{
    const int n = P.n, T = 1 << P.tshift, tmask = T - 1;
    const FftPlanDev &pl = P.plan;
    cplx *twbuf = (cplx *)smem_raw;
    cplx *s = P.tw_smem ? twbuf + n : twbuf;
    const cplx *tw = P.tw_smem ? (const cplx *)twbuf : P.tw;
    const double alpha = P.alpha[which];
    const bool do_pow = POW && P.use_pow && alpha != 1.0;
    const long long nblk = P.outer * P.tiles;
    const int q = n / R0, total = q << P.tshift;
    const size_t rstride = (size_t)q * (size_t)P.inner;        // rows k*q of the first butterfly

    const double *ptab = nullptr;
    if (do_pow && P.pow_tab_off >= 0) {
#if defined(__CUDA_ARCH__)
        double *pt = (double *)(smem_raw + P.pow_tab_off);
        for (int e = tid; e < DBSPM_POW_TABLE_LEN * POW_TAB_COPIES; e += nthr) pt[e] = dbspm_pow_table_d[e / POW_TAB_COPIES];
        ptab = pt;
#else
        ptab = dbspm_pow_table_h;
#endif
    }
    if (P.tw_smem)
        for (int e = tid; e < n; e += nthr) twbuf[e] = ldg_c(P.tw + e);
    BLOCK_SYNC();

    // the item of the first stage this thread keeps in flight across tiles
    const bool mine = tid < total;
    const int t0 = tid & tmask, j0 = tid >> P.tshift;
    cplx *stg = (cplx *)(smem_raw + P.stage_off);
    cplx hi[8];
    const cplx *gin; cplx *gout; int tcount;
    // (hi is assigned unconditionally -- zeros when there is nothing to load -- so that it is dead between the first stage
    // and the next request; a conditional assignment keeps its registers live through the whole loop)
    {
        const long long bb = b0 < nblk ? b0 : 0;
        axis2s_tile(P, bb, which, &gin, &gout, &tcount);
        axis2s_request(gin + (size_t)j0 * P.inner + t0, rstride, R0, b0 < nblk && mine && t0 < tcount, stg, tid, nthr, hi);
    }
    for (long long b = b0; b < nblk; b += bstride) {
        axis2s_tile(P, b, which, &gin, &gout, &tcount);
        if (mine) {
            cplx v[16];
            axis2s_collect(stg, R0, tid, nthr, hi, v);
            axis2s_stage1<D, R0>(P, s, tw, v, t0 < tcount, do_pow, alpha, ptab, which, q, j0, t0, tid);
        }
#if !defined(__CUDA_ARCH__)
        // CPU emulation only (one "thread" per block): the remaining items of the tile, loaded directly.  On the device the
        // launcher guarantees one item per thread (total <= nthr)
        for (int idx = tid + nthr; idx < total; idx += nthr) {
            const int t = idx & tmask, j = idx >> P.tshift;
            cplx u[16];
            axis2s_load(gin + (size_t)j * P.inner + t, rstride, R0, t < tcount, u);
            axis2s_stage1<D, R0>(P, s, tw, u, t < tcount, do_pow, alpha, ptab, which, q, j, t, tid);
        }
#endif
        BLOCK_SYNC();
        {   // next tile's rows: in flight during the remaining stages of this one
            const long long bn = b + bstride;
            const cplx *gin2; cplx *gout2; int tc2;
            axis2s_tile(P, bn < nblk ? bn : b, which, &gin2, &gout2, &tc2);
            axis2s_request(gin2 + (size_t)j0 * P.inner + t0, rstride, R0, bn < nblk && mine && t0 < tc2, stg, tid, nthr, hi);
        }
        int m = n / R0;
        for (int st = 1; st + 1 < pl.nstages; ++st) {
            fft_stage<RM, D>(s, T, P.tshift, n, m, tw, tid, nthr);
            m /= RM;
            BLOCK_SYNC();
        }
        axis2_last<D, RL, 0>(P, s, gout, tcount, which, tid, nthr);
        BLOCK_SYNC();                       // the tile is overwritten by the next first stage
    }
}

// ------------------------------------------------------------------------------------------
// v3 of the strided-axis pass: TMA-fed persistent blocks.  Each block walks over its share of the
// tiles with a ring of NBUF shared-memory buffers.  One thread issues
//   - a tensor-map LOAD (cp.async.bulk.tensor, completion on an mbarrier) of the whole n x T tile
//     for the tile NBUF-1 steps ahead, and
//   - after the in-place FFT, ONE tensor-map STORE whose box dimensions are the radix digits in
//     reversed order with their natural global strides: the TMA unit performs the mixed-radix
//     digit reversal (tile row k1*(n/r1)+...+ks  ->  global row k1 + r1*k2 + ...) on the fly.
// No global load/store instruction, no register staging and no L1 traffic is left in the block;
// the FFT stages run on shared memory while the loads of the next two tiles and the store of
// the previous one are in flight.  On the host (tests/emu) the same body runs with the two TMA
// operations emulated by plain copies, so tile bookkeeping and stage order are still checked.
// ------------------------------------------------------------------------------------------
#ifndef AXIS3_NBUF
#define AXIS3_NBUF 3
#endif
#ifndef AXIS3_MAX_SMEM
#define AXIS3_MAX_SMEM (225 * 1024)
#endif

template <int D, int MAXR = 16>
HD void fft_stages_from(cplx *s, int ld, int tshift, const FftPlanDev &pl, int first, const cplx *tw, int tid, int nthr);

struct AxisTmaParams {
    const cplx *in[2];        // used by the host emulation only (the device goes through tensor maps)
    cplx *out[2];
    double alpha[2];
    int use_pow;
    int narrays;
    int n;
    int tshift;               // T = 1 << tshift lines per tile
    long long inner;
    long long outer;
    long long tiles;          // tiles per (array, slab) = ceil(inner / T)
    FftPlanDev plan;
    const cplx *tw;
    const int *freq_of;       // host emulation of the permuting store
};

HD size_t axis3_tile_bytes(int n, int tshift) { return ((size_t)n << tshift) * 16; }
HD size_t axis3_smem_bytes(int n, int tshift)
{
    return AXIS3_NBUF * axis3_tile_bytes(n, tshift) + (size_t)n * 16 + 2 * AXIS3_NBUF * 8 + 128;
}
// widest tile (8,4,2,1 lines) whose ring fits; -1 if even one line does not
HD int axis3_config(int n, long long inner)
{
    for (int ts = 3; ts >= 0; --ts) {
        if ((1LL << ts) > 2 * inner && ts > 0) continue;
        if (axis3_smem_bytes(n, ts) <= AXIS3_MAX_SMEM) return ts;
    }
    return -1;
}

#if defined(__CUDA_ARCH__)
HD unsigned axis3_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
#endif

// Warp-specialised layout of the block: AXIS3_GROUPS consumer groups of AXIS3_GROUP_THREADS threads
// (each group runs the FFT of one tile, so two tiles are in compute at any time) followed by one
// producer warp whose lane 0 issues the tensor-map loads.  full[b] / empty[b] mbarriers per ring slot.
#ifndef AXIS3_GROUPS
#define AXIS3_GROUPS 2
#endif
#ifndef AXIS3_GROUP_THREADS
#define AXIS3_GROUP_THREADS 192
#endif
#define AXIS3_BLOCK_THREADS (AXIS3_GROUPS * AXIS3_GROUP_THREADS + 32)

#if defined(__CUDA_ARCH__)
#define GROUP_SYNC(id, n) asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory")
#else
#define GROUP_SYNC(id, n) ((void)0)
#endif

// all DIF stages on a tile, synchronising only the `nthr` threads of one consumer group (named barrier)
template <int D>
HD void fft_stages_group(cplx *s, int ld, int tshift, const FftPlanDev &pl, const cplx *tw, int tid, int nthr, int bar_id)
{
    int m = pl.n;
    for (int st = 0; st < pl.nstages; ++st) {
        const int r = pl.radix[st];
#define CALL_(R) fft_stage<R, D>(s, ld, tshift, pl.n, m, tw, tid, nthr)
        DBSPM_FAST_RADIX_SWITCH(r, CALL_)
#undef CALL_
        m /= r;
        GROUP_SYNC(bar_id, nthr);
    }
}

#if defined(__CUDA_ARCH__)
HD void mbar_wait(unsigned bar, unsigned parity)
{
    unsigned done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
#endif

// tensor-map arguments: lmaps/smaps point at CUtensorMap objects in kernel parameter space.
// Device: called by all AXIS3_BLOCK_THREADS threads (tid = threadIdx.x).  Host emulation: tid = 0, nthr = 1.
template <int D>
HD void axis_tma_body(const AxisTmaParams &P, const void *const *lmaps, const void *const *smaps,
                      unsigned char *smem_raw, long long bid, long long nblocks, int tid, int nthr)
{
    const int n = P.n, T = 1 << P.tshift;
    const size_t tile_bytes = axis3_tile_bytes(n, P.tshift);
    // pointer arithmetic, not a round trip through an integer: the compiler must keep knowing that this is shared memory
    // (round 1 aligned through size_t and every access of the tile became a generic LD.E / ST.E instead of LDS / STS)
    unsigned char *base = smem_raw + ((128 - ((size_t)smem_raw & 127)) & 127);
    cplx *tw = (cplx *)(base + AXIS3_NBUF * tile_bytes);
    unsigned long long *full = (unsigned long long *)(tw + n);
    unsigned long long *empty = full + AXIS3_NBUF;
    const long long per_array = P.outer * P.tiles;
    const long long total = per_array * P.narrays;
    const long long mine = bid < total ? (total - bid + nblocks - 1) / nblocks : 0;

    for (int e = tid; e < n; e += nthr) tw[e] = ldg_c(P.tw + e);
#if defined(__CUDA_ARCH__)
    if (tid == 0) {
        for (int b = 0; b < AXIS3_NBUF; ++b) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(axis3_smem_u32(full + b)));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(axis3_smem_u32(empty + b)));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
#endif
    BLOCK_SYNC();

#if defined(__CUDA_ARCH__)
    const int group = tid / AXIS3_GROUP_THREADS;               // AXIS3_GROUPS = producer warp
    if (group >= AXIS3_GROUPS) {
        // ---------------- producer: one lane streams the tiles into the ring ----------------
        if ((tid & 31) == 0) {
            for (long long i = 0; i < mine; ++i) {
                const int slot = (int)(i % AXIS3_NBUF);
                if (i >= AXIS3_NBUF) mbar_wait(axis3_smem_u32(empty + slot), (unsigned)(((i / AXIS3_NBUF) - 1) & 1));
                const long long t = bid + i * nblocks;
                const int which = (int)(t / per_array);
                const long long r = t - which * per_array;
                const long long o = r / P.tiles, tile = r - o * P.tiles;
                const unsigned bar = axis3_smem_u32(full + slot);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((unsigned)tile_bytes) : "memory");
                asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                             ::"r"(axis3_smem_u32(base + (size_t)slot * tile_bytes)), "l"(lmaps[which]),
                               "r"((int)(tile << (P.tshift + 1))), "r"(0), "r"(0), "r"((int)o), "r"(bar)
                             : "memory");
            }
        }
        return;
    }
    // ---------------- consumers: group g takes my tiles g, g + GROUPS, ... ----------------------
    const int gtid = tid - group * AXIS3_GROUP_THREADS;
    const int gn = AXIS3_GROUP_THREADS;
    const int bar_id = 1 + group;
    for (long long i = group; i < mine; i += AXIS3_GROUPS) {
        const int slot = (int)(i % AXIS3_NBUF);
        const long long t = bid + i * nblocks;
        const int which = (int)(t / per_array);
        const long long r = t - which * per_array;
        const long long o = r / P.tiles, tile = r - o * P.tiles;
        cplx *s = (cplx *)(base + (size_t)slot * tile_bytes);
        mbar_wait(axis3_smem_u32(full + slot), (unsigned)((i / AXIS3_NBUF) & 1));
        const double alpha = P.alpha[which];
        if (P.use_pow && alpha != 1.0) {
            const int total_e = n << P.tshift;
            for (int idx = gtid; idx < total_e; idx += gn) {
                cplx v = s[idx];
                v.x = pow_nonneg(v.x, alpha);
                v.y = pow_nonneg(v.y, alpha);
                s[idx] = v;
            }
            GROUP_SYNC(bar_id, gn);
        }
        fft_stages_group<D>(s, T, P.tshift, P.plan, tw, gtid, gn, bar_id);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the TMA unit
        GROUP_SYNC(bar_id, gn);
        if (gtid == 0) {
            const int c0 = (int)(tile << (P.tshift + 1));
            const int ns = P.plan.nstages;
            const void *sm = smaps[which];
            const unsigned src = axis3_smem_u32(s);
            if (ns == 1)
                asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
                             ::"l"(sm), "r"(c0), "r"(0), "r"((int)o), "r"(src) : "memory");
            else if (ns == 2)
                asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4}], [%5];"
                             ::"l"(sm), "r"(c0), "r"(0), "r"(0), "r"((int)o), "r"(src) : "memory");
            else
                asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4, %5}], [%6];"
                             ::"l"(sm), "r"(c0), "r"(0), "r"(0), "r"(0), "r"((int)o), "r"(src) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // the TMA unit has read the slot
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(axis3_smem_u32(empty + slot)) : "memory");
        }
    }
    if (gtid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all stores landed before the block exits
#else
    // ---------------- host emulation: the two TMA operations as plain copies --------------------
    (void)lmaps; (void)smaps; (void)full; (void)empty;
    for (long long i = 0; i < mine; ++i) {
        const long long t = bid + i * nblocks;
        const int which = (int)(t / per_array);
        const long long r = t - which * per_array;
        const long long o = r / P.tiles, tile = r - o * P.tiles;
        cplx *s = (cplx *)(base + (size_t)(i % AXIS3_NBUF) * tile_bytes);
        const cplx *src = P.in[which] + (size_t)o * n * P.inner + (tile << P.tshift);
        const int tcount = (int)((P.inner - (tile << P.tshift)) < T ? (P.inner - (tile << P.tshift)) : T);
        for (int row = 0; row < n; ++row)
            for (int c = 0; c < T; ++c) s[(size_t)row * T + c] = c < tcount ? src[(size_t)row * P.inner + c] : cmake(0.0, 0.0);
        const double alpha = P.alpha[which];
        if (P.use_pow && alpha != 1.0)
            for (int idx = 0; idx < (n << P.tshift); ++idx) { s[idx].x = pow_nonneg(s[idx].x, alpha); s[idx].y = pow_nonneg(s[idx].y, alpha); }
        fft_stages_group<D>(s, T, P.tshift, P.plan, tw, tid, nthr, 0);
        cplx *dst = P.out[which] + (size_t)o * n * P.inner + (tile << P.tshift);
        for (int p = 0; p < n; ++p)
            for (int c = 0; c < tcount; ++c) dst[(size_t)P.freq_of[p] * P.inner + c] = s[(size_t)p * T + c];
    }
#endif
}

// ------------------------------------------------------------------------------------------
// Fused z kernel.  One block = ZL line pairs {p=(kx,ky), m=(-kx,-ky)}; per pair 4 lines
// (A_p, A_m, B_p, B_m) of n = Nz/2 complex go through a forward FFT; the untangle + product
// + window phase + first inverse radix-2 split produce 2 lines (spectra of the packed window
// pairs for p and for m); an inverse FFT of length n and the first nwh = ceil(Nzw/2) outputs
// are written to W[(kx,ky)] and W[(-kx,-ky)].
// ------------------------------------------------------------------------------------------
#define ZL 2                  // line pairs per block
#define ZT_FWD 8              // 4*ZL lines in the forward tile
#define ZT_INV 4              // 2*ZL lines in the inverse tile
#define ZLD_FWD (ZT_FWD + 1)  // padded leading dimensions: the transposing load/store is
#define ZLD_INV (ZT_INV + 1)  // bank-conflict free, butterfly accesses stay contiguous

struct ZFusedParams {
    const cplx *ZA, *ZB;      // (nx,ny,n) complex, x/y already transformed
    cplx *W;                  // (nx,ny,nwh) complex
    int nx, ny, n;            // n = nz/2
    int nwh;                  // ceil(nzw/2)
    int z_lo;
    double scale;             // dV / (nx*ny*nz) / 4
    FftPlanDev plan;          // length n
    const cplx *tw;           // length-n table
    const int *pos_of;        // length n: frequency -> tile row
    const int *freq_of;       // length n: tile row -> frequency
    const cplx *twN;          // length-2n table: twN[e] = exp(-2 pi i e / nz)
    long long npairs;         // (nx/2+1) * ny candidate pairs
    // distributed correlation: this rank holds the ny rows of the y axis that are slots [slot_base, slot_base + ny)
    // of the global axis (dist_slot_of); row r pairs with the row of its mirror slot.  slot_mode = 0: natural rows
    int slot_mode, slot_base;
    int prefetch;             // 1: a persistent block asks the L2 for the 4*ZLV lines of its next pair group (bulk prefetch)
};

HD size_t zfused_smem_bytes(int n)
{
    return (size_t)n * 16 * 2 + (size_t)n * 4 + (size_t)n * ZLD_FWD * 16 + (size_t)n * ZLD_INV * 16;
}

struct ZPair { int kx, ky, mx, my, valid, self; };

HD ZPair zpair_of(const ZFusedParams &P, long long pid)
{
    ZPair z;
    z.valid = 0; z.self = 0; z.kx = z.ky = z.mx = z.my = 0;
    if (pid >= P.npairs) return z;
    z.kx = (int)(pid / P.ny);
    z.ky = (int)(pid - (long long)z.kx * P.ny);
    z.mx = z.kx == 0 ? 0 : P.nx - z.kx;
    z.my = P.slot_mode ? dist_mirror_slot(P.slot_base + z.ky) - P.slot_base : (z.ky == 0 ? 0 : P.ny - z.ky);
    if (z.mx == z.kx) {                       // kx is its own mirror: keep one of (ky, -ky)
        if (z.ky > z.my) return z;
        z.self = (z.ky == z.my);
    }
    z.valid = 1;
    return z;
}

HD void zfused_body(const ZFusedParams &P, unsigned char *smem_raw, long long bid, int tid, int nthr)
{
    const int n = P.n;
    cplx *tw = (cplx *)smem_raw;              // exp(-2 pi i e / n)
    cplx *twh = tw + n;                       // exp(-2 pi i e / (2n)), e < n
    cplx *U = twh + n;                        // forward tile  [n][ZLD_FWD]
    cplx *S = U + (size_t)n * ZLD_FWD;        // inverse tile  [n][ZLD_INV]
    int *pos_of = (int *)(S + (size_t)n * ZLD_INV);

    ZPair pr[ZL];
#pragma unroll
    for (int l = 0; l < ZL; ++l) pr[l] = zpair_of(P, bid * ZL + l);

    for (int e = tid; e < n; e += nthr) {
        tw[e] = ldg_c(P.tw + e);
        twh[e] = ldg_c(P.twN + e);
        pos_of[e] = LDG(P.pos_of + e);
    }
    // load 4*ZL lines (contiguous in global memory) into the transposed tile
    for (int idx = tid; idx < n * ZT_FWD; idx += nthr) {
        const int line = idx / n, i = idx - line * n;
        const int l = line >> 2, role = line & 3;
        cplx v = cmake(0.0, 0.0);
        if (pr[l].valid) {
            const cplx *src = (role & 2) ? P.ZB : P.ZA;
            const int x = (role & 1) ? pr[l].mx : pr[l].kx, y = (role & 1) ? pr[l].my : pr[l].ky;
            v = src[((size_t)x * P.ny + y) * n + i];
        }
        U[(size_t)i * ZLD_FWD + line] = v;
    }
    BLOCK_SYNC();
    fft_tile<-1>(U, ZLD_FWD, 3, P.plan, tw, tid, nthr);

    // pointwise stage: one work item = (pair l, frequency k <= n/2) handles k and k* = (n-k)%n
    const int nk = n / 2 + 1;
    const double sgn = (P.z_lo & 1) ? -1.0 : 1.0;
    const long long N2 = 2LL * n;
    for (int idx = tid; idx < nk * ZL; idx += nthr) {
        const int l = idx / nk, k = idx - l * nk;
        const int ks = k == 0 ? 0 : n - k;
        if (k > ks) continue;
        const cplx *Uk = U + (size_t)pos_of[k] * ZLD_FWD + 4 * l;
        const cplx *Us = U + (size_t)pos_of[ks] * ZLD_FWD + 4 * l;
        cplx G1[2], G2[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            // h = 0: frequency k (needs U_p[k], U_m[k*]); h = 1: frequency k* (needs U_p[k*], U_m[k])
            const int kk = h == 0 ? k : ks;
            const cplx *Ua = h == 0 ? Uk : Us, *Ub = h == 0 ? Us : Uk;
            const cplx w = twh[kk];                                   // exp(-2 pi i kk / nz)
            const cplx miw = cmake(w.y, -w.x);                        // -i * w
            cplx a1 = Ua[0], a2 = cconj(Ub[1]);
            cplx b1 = Ua[2], b2 = cconj(Ub[3]);
            cplx ea = cadd(a1, a2), ta = cmul(miw, csub(a1, a2));
            cplx eb = cadd(b1, b2), tb = cmul(miw, csub(b1, b2));
            cplx alo = cadd(ea, ta), ahi = csub(ea, ta);              // 2*A^[kk], 2*A^[kk+n]
            cplx blo = cadd(eb, tb), bhi = csub(eb, tb);
            cplx plo = cscale(cmulc(alo, blo), P.scale);
            cplx phi = cscale(cmulc(ahi, bhi), P.scale * sgn);
            const cplx ph = cconj(ldg_c(P.twN + (int)(((long long)kk * P.z_lo) % N2)));   // e^{+2 pi i kk z_lo / nz}
            plo = cmul(plo, ph);
            phi = cmul(phi, ph);
            G1[h] = cadd(plo, phi);
            G2[h] = cmulc(csub(plo, phi), w);                         // * e^{+2 pi i kk / nz}
        }
        cplx *Sk = S + (size_t)k * ZLD_INV + 2 * l;
        cplx *Ss = S + (size_t)ks * ZLD_INV + 2 * l;
        // S_p[kk] = G1 + i G2 ; S_m[kk] = conj(G1[kk*]) + i conj(G2[kk*])
        Sk[0] = cmake(G1[0].x - G2[0].y, G1[0].y + G2[0].x);
        Sk[1] = cmake(G1[1].x + G2[1].y, -G1[1].y + G2[1].x);
        if (ks != k) {
            Ss[0] = cmake(G1[1].x - G2[1].y, G1[1].y + G2[1].x);
            Ss[1] = cmake(G1[0].x + G2[0].y, -G1[0].y + G2[0].x);
        }
    }
    BLOCK_SYNC();
    fft_tile<1>(S, ZLD_INV, 2, P.plan, tw, tid, nthr);

    for (int idx = tid; idx < P.nwh * ZT_INV; idx += nthr) {
        const int line = idx / P.nwh, j = idx - line * P.nwh;
        const int l = line >> 1, role = line & 1;
        if (!pr[l].valid || (role == 1 && pr[l].self)) continue;
        const int x = role ? pr[l].mx : pr[l].kx, y = role ? pr[l].my : pr[l].ky;
        P.W[((size_t)x * P.ny + y) * P.nwh + j] = S[(size_t)pos_of[j] * ZLD_INV + line];
    }
}

// ------------------------------------------------------------------------------------------
// v2 of the fused z kernel: ZLV line pairs per block (4*ZLV forward lines keep every butterfly
// stage at >= one work item per thread with radix-16 stages), and the first forward stage reads
// the contiguous z-lines straight from global memory into registers (coalesced: consecutive
// threads take consecutive positions of one line) and writes the transposed, padded tile.
// Pair bookkeeping lives in shared memory (no per-thread arrays).  Same algebra as zfused_body.
// ------------------------------------------------------------------------------------------
template <int ZLV> HD size_t zfused2_smem_bytes(int n)
{
    return (size_t)n * 16 * 2 + (size_t)n * (4 * ZLV + 1) * 16 + (size_t)n * (2 * ZLV + 1) * 16 + (size_t)n * 8 + ZLV * 6 * 4 + 16;
}

template <int D, int R, int ZLV>
HD void zfused2_first(const ZFusedParams &P, cplx *U, const cplx *tw, const int *pinfo, int tid, int nthr)
{
    constexpr int LDF = 4 * ZLV + 1;
    const int n = P.n, q = n / R;
    const int total = 4 * ZLV * q;
    for (int idx = tid; idx < total; idx += nthr) {
        const int line = idx / q, j = idx - line * q;
        const int l = line >> 2, role = line & 3;
        const int *pi = pinfo + 6 * l;
        cplx v[R];
        if (pi[4]) {
            const cplx *src = ((role & 2) ? P.ZB : P.ZA)
                              + ((size_t)((role & 1) ? pi[2] : pi[0]) * P.ny + ((role & 1) ? pi[3] : pi[1])) * n + j;
#pragma unroll
            for (int k = 0; k < R; ++k) v[k] = ld_stream(src + k * q);
        } else {
#pragma unroll
            for (int k = 0; k < R; ++k) v[k] = cmake(0.0, 0.0);
        }
        bfly<R, D>(v);
        if (q > 1) {
#pragma unroll
            for (int k = 1; k < R; ++k) {
                const cplx w = tw[j * k];
                v[k] = D < 0 ? cmul(v[k], w) : cmulc(v[k], w);
            }
        }
        cplx *dst = U + (j * LDF + line);
#pragma unroll
        for (int k = 0; k < R; ++k) dst[k * q * LDF] = v[k];
    }
}

template <int D, int MAXR>
HD void fft_stages_from(cplx *s, int ld, int tshift, const FftPlanDev &pl, int first, const cplx *tw, int tid, int nthr)
{
    int m = pl.n;
    for (int st = 0; st < first; ++st) m /= pl.radix[st];
    for (int st = first; st < pl.nstages; ++st) {
        const int r = pl.radix[st];
#define CALL_(R) fft_stage<R, D>(s, ld, tshift, pl.n, m, tw, tid, nthr)
        DBSPM_RADIX_SWITCH_MAX(r, CALL_, MAXR)
#undef CALL_
        m /= r;
        BLOCK_SYNC();
    }
}

template <int ZLV, int MAXR = 16>
HD void zfused_body_v2(const ZFusedParams &P, unsigned char *smem_raw, long long bid0, long long bstride, int tid, int nthr)
{
    constexpr int TF = 4 * ZLV, TI = 2 * ZLV, LDF = TF + 1, LDI = TI + 1;
    constexpr int TSF = ZLV == 1 ? 2 : (ZLV == 2 ? 3 : (ZLV == 4 ? 4 : 5));
    constexpr int TSI = TSF - 1;
    const int n = P.n;
    cplx *tw = (cplx *)smem_raw;
    cplx *twh = tw + n;
    cplx *U = twh + n;
    cplx *S = U + (size_t)n * LDF;
    int *pos_of = (int *)(S + (size_t)n * LDI);
    int *freq_of = pos_of + n;
    int *pinfo = freq_of + n;                 // per pair: kx, ky, mx, my, valid, self

    for (int e = tid; e < n; e += nthr) {
        tw[e] = ldg_c(P.tw + e);
        twh[e] = ldg_c(P.twN + e);
        pos_of[e] = LDG(P.pos_of + e);
        freq_of[e] = LDG(P.freq_of + e);
    }
    // persistent block: the tables above are staged once, then the block walks over its share of
    // the pair groups
    const long long ngroups = (P.npairs + ZLV - 1) / ZLV;
    for (long long bid = bid0; bid < ngroups; bid += bstride) {
    for (int l = tid; l < ZLV; l += nthr) {
        const ZPair z = zpair_of(P, bid * ZLV + l);
        int *pi = pinfo + 6 * l;
        pi[0] = z.kx; pi[1] = z.ky; pi[2] = z.mx; pi[3] = z.my; pi[4] = z.valid; pi[5] = z.self;
    }
    BLOCK_SYNC();
#if defined(__CUDA_ARCH__)
    if (P.prefetch && bid + bstride < ngroups) {
        for (int line = tid; line < 4 * ZLV; line += nthr) {
            const ZPair z = zpair_of(P, (bid + bstride) * ZLV + (line >> 2));
            if (z.valid) {
                const int role = line & 3;
                const cplx *src = ((role & 2) ? P.ZB : P.ZA) + ((size_t)((role & 1) ? z.mx : z.kx) * P.ny + ((role & 1) ? z.my : z.ky)) * n;
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(n * 16) : "memory");
            }
        }
    }
#endif
#define CALL_(R) zfused2_first<-1, R, ZLV>(P, U, tw, pinfo, tid, nthr)
    DBSPM_RADIX_SWITCH_MAX(P.plan.radix[0], CALL_, MAXR)
#undef CALL_
    BLOCK_SYNC();
    fft_stages_from<-1, MAXR>(U, LDF, TSF, P.plan, 1, tw, tid, nthr);

    // pointwise stage in POSITION order: one work item = (pair l, tile row p) owns frequency
    // k = freq_of[p].  Consecutive threads read consecutive rows (bank-conflict free); only the
    // mirror frequency k* = (n-k)%n lives in a scattered row.  It writes S_p at its own row and
    // S_m at the mirror row, i.e. the inverse tile holds the spectra in the same permuted order
    // and the inverse transform runs the transposed (decimation-in-time) stages.
    const double sgn = (P.z_lo & 1) ? -1.0 : 1.0;
    const long long N2 = 2LL * n;
    for (int idx = tid; idx < n * ZLV; idx += nthr) {
        const int l = idx / n, p = idx - l * n;
        const int k = freq_of[p];
        const int ks = k == 0 ? 0 : n - k;
        const int ps = pos_of[ks];
        const cplx *Uk = U + (p * LDF + 4 * l);
        const cplx *Us = U + (ps * LDF + 4 * l);
        const cplx w = twh[k];                                    // exp(-2 pi i k / nz)
        const cplx miw = cmake(w.y, -w.x);                        // -i * w
        const cplx a1 = Uk[0], a2 = cconj(Us[1]);
        const cplx b1 = Uk[2], b2 = cconj(Us[3]);
        const cplx ea = cadd(a1, a2), ta = cmul(miw, csub(a1, a2));
        const cplx eb = cadd(b1, b2), tb = cmul(miw, csub(b1, b2));
        const cplx alo = cadd(ea, ta), ahi = csub(ea, ta);        // 2*A^[k], 2*A^[k+n]
        const cplx blo = cadd(eb, tb), bhi = csub(eb, tb);
        cplx plo = cscale(cmulc(alo, blo), P.scale);
        cplx phi = cscale(cmulc(ahi, bhi), P.scale * sgn);
        const cplx ph = cconj(ldg_c(P.twN + (int)(((long long)k * P.z_lo) % N2)));   // e^{+2 pi i k z_lo / nz}
        plo = cmul(plo, ph);
        phi = cmul(phi, ph);
        const cplx G1 = cadd(plo, phi);
        const cplx G2 = cmulc(csub(plo, phi), w);                 // * e^{+2 pi i k / nz}
        // S_p[k] = G1 + i G2 ; S_m[k*] = conj(G1[k]) + i conj(G2[k])
        S[p * LDI + 2 * l] = cmake(G1.x - G2.y, G1.y + G2.x);
        S[ps * LDI + 2 * l + 1] = cmake(G1.x + G2.y, -G1.y + G2.x);
    }
    BLOCK_SYNC();
    fft_tile_t<1, MAXR>(S, LDI, TSI, P.plan, tw, tid, nthr);

    for (int idx = tid; idx < P.nwh * TI; idx += nthr) {
        const int line = idx / P.nwh, j = idx - line * P.nwh;
        const int l = line >> 1, role = line & 1;
        const int *pi = pinfo + 6 * l;
        if (!pi[4] || (role == 1 && pi[5])) continue;
        const int x = role ? pi[2] : pi[0], y = role ? pi[3] : pi[1];
        st_stream(P.W + ((size_t)x * P.ny + y) * P.nwh + j, S[j * LDI + line]);
    }
    BLOCK_SYNC();                              // S and pinfo are rewritten by the next group
    }
}

// ------------------------------------------------------------------------------------------
// compaction when the window has an odd number of planes: W (nx*ny, 2*nwh) -> E (nx*ny, nzw)
// ------------------------------------------------------------------------------------------
// ---- odd nz: plain complex transforms (the packed real transform needs an even z length) --------
// Elementwise bodies around six ordinary axis passes: real -> complex (power fused), spectrum product
// A * conj(B) * scale, real part of the window planes.  Twice the traffic of the packed path; VASP FFT
// grids are even, so this only keeps get_density_overlap() total over shapes like the reference's.
HD void r2c_body(const double *in, cplx *out, size_t n, double alpha, size_t gid, size_t gstride)
{
    const bool do_pow = alpha != 1.0;
    for (size_t i = gid; i < n; i += gstride) {
        const double v = in[i];
        out[i] = cmake(do_pow ? pow_nonneg(v, alpha) : v, 0.0);
    }
}

HD void cmulconj_body(cplx *a, const cplx *b, size_t n, double scale, size_t gid, size_t gstride)
{
    for (size_t i = gid; i < n; i += gstride) a[i] = cscale(cmulc(a[i], b[i]), scale);
}

HD void c2r_window_body(const cplx *W, double *E, long long ncol, int nz, int z_lo, int nzw, long long gid, long long gstride)
{
    const long long total = ncol * nzw;
    for (long long i = gid; i < total; i += gstride) {
        const long long c = i / nzw;
        const int z = (int)(i - c * nzw);
        E[i] = W[c * nz + z_lo + z].x;
    }
}

HD void compact_body(const double *W, double *E, long long ncol, int nzw, int ldw, long long gid, long long gstride)
{
    const long long total = ncol * nzw;
    for (long long i = gid; i < total; i += gstride) {
        const long long c = i / nzw;
        const int z = (int)(i - c * nzw);
        E[i] = W[c * ldw + z];
    }
}
