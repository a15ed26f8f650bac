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
#include "fft_core.h"

// ------------------------------------------------------------------------------------------
// Strided-axis pass: FFT along an axis of length n whose elements are `inner` complex apart;
// `outer` independent slabs.  One block = T adjacent lines of one slab, staged in shared
// memory as s[n][T]; global rows are T*16-byte contiguous segments (128 B for T = 8).
// blockIdx.y selects one of two arrays (A and B travel in one launch).
// ------------------------------------------------------------------------------------------
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
    const int *pos_of;
};

HD size_t axis_smem_bytes(int n, int tshift)
{
    return (size_t)n * 16 + (size_t)n * 4 + ((size_t)n << tshift) * 16;
}

HD double pow_nonneg(double v, double alpha)
{
    // rho >= 0 upstream (grid.py:450-464 filter_neg); 0 -> exp(-inf) = 0; v < 0 -> NaN like numpy
    return exp(alpha * log(v));
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
    const int *pos_of;        // length n
    const cplx *twN;          // length-2n table: twN[e] = exp(-2 pi i e / nz)
    long long npairs;         // (nx/2+1) * ny candidate pairs
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
    z.my = z.ky == 0 ? 0 : P.ny - z.ky;
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
// compaction when the window has an odd number of planes: W (nx*ny, 2*nwh) -> E (nx*ny, nzw)
// ------------------------------------------------------------------------------------------
HD void compact_body(const double *W, double *E, long long ncol, int nzw, int ldw, long long gid, long long gstride)
{
    const long long total = ncol * nzw;
    for (long long i = gid; i < total; i += gstride) {
        const long long c = i / nzw;
        const int z = (int)(i - c * nzw);
        E[i] = W[c * ldw + z];
    }
}
