// relax_body.h -- per-thread body of the relax kernel.
//
// Replaces, for one lateral position (i,j), the reference's
//   relax_tip_z / relax_noortho_tip_z   pydbspm/interactions.py:70-105
//   Emin / Emin_norm / spring_energy    pydbspm/interactions.py:49-67
//   sph2grid / sph2cart                 pydbspm/grid.py:364-386
//   tricubic(...).ip                    third-party `tricubic` (Lekien-Marsden, periodic,
//                                       index units) == separable Catmull-Rom stencil
//   scipy.optimize.minimize("Powell")   scipy/optimize/_optimize.py:3413-3643 with
//                                       _linesearch_powell :3214-3247, bracket :2954-3111,
//                                       Brent.optimize :2497-2611, bracket-error recovery
//                                       :3117-3148, maxfun wrapper :549-572
//
// The optimiser is restated as an explicit state machine with ONE objective-evaluation site
// per loop trip: every thread of a warp runs cheap, divergent scalar bookkeeping until it
// needs a function value, then all threads evaluate the 64-point stencil together.  Control
// flow, constants and comparison directions follow scipy line by line; Powell bookkeeping is
// compiled without FMA contraction (plain * and +, like numpy).
#pragma once
#include <string.h>
#include "hd.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

struct RelaxParams {
    const double *grid;       // static potential: (nx,ny,nzc) z fastest [LAYOUT 0] or (nx,nzc,nyp) [LAYOUT 1]
    int nx, ny, nzc;
    int x_base, nx_local;     // x-slab mode: `grid` holds the nx_local planes x_base, x_base + 1, ... (mod nx) of the window (a rank's
                              // x-tile plus a halo of the lever length on both sides); x_base = 0, nx_local = nx: the whole window
    int nyp;                  // LAYOUT 1: row length ny + 3 (periodic halo: stored index t holds y = t - 1)
    int i_lo, i_hi, j_lo, j_hi;
    int foot_x;               // lanes of a warp along x (1, 2, 4, 8, 16 or 32; the other 32/foot_x along y): which 32 columns a warp relaxes.
                              // x-neighbouring lanes share 12 of their 16 stencil sectors in the y-quad layout (L1 hits / hit-on-miss)
    int nz_relax;
    int k_hi, k_lo;           // this launch sweeps the heights k_hi, k_hi - 1, ..., k_lo (whole sweep: nz_relax - 1 ... 0); the start
                              // point of height k_hi < nz_relax - 1 is the relaxed (theta, phi) of height k_hi + 1 in out_etp
    double kappa;
    double rpivot;            // params.RPIVOT (used by sph2cart, dbspm:605)
    double npivot;            // rpivot / dr[2]                     (dbspm:558)
    double rpivot_rt;         // npivot * dr[2]                     (interactions.py:73)
    double dr[3];
    double inv_dr[3];         // RN(1/dr): division by dr as multiply + two fused corrections (div_by)
    double inv_n[3];          // 1/nx, 1/ny, 1/nzc
    double nd[3];             // nx, ny, nzc as doubles
    int noortho;
    double Minv[9];           // inv(dR^T), row-major
    double xtol, ftol;
    int maxfev, maxiter;
    double *out_etp;          // (ni,nj,nz,3)
    double *out_cart;         // (3,ni,nj,nz) or null
    int *out_nfev;            // (ni,nj,nz) or null
};

// ---- sin and cos, same bits on the device and on the host -----------------------------------
// A fixed sequence of IEEE operations (explicit fma, no library call): Cody-Waite reduction by pi/2 in three
// pieces, then the classical degree-13 / degree-14 minimax kernels on [-pi/4, pi/4] (coefficients of fdlibm's
// __kernel_sin / __kernel_cos; the cosine with fdlibm's compensated 1 - z/2).  Within one ulp of glibc for
// |x| <= 1e5 (tests/test_emulation.py), which is what CUDA's own sincos() delivers; unlike a library call the
// result does not depend on who compiled the caller, so the CPU twin of this kernel (oracle/dbspm_oracle.c,
// "kernel twin" mode) reproduces it bit for bit.  Larger or non-finite arguments (never produced by a
// relaxation) go to the library.  On the device the 16 constants come from constant memory (two per 128-bit
// uniform load) instead of two 32-bit immediate moves each: a third of the instructions of the first version.
#define DBSPM_SC_TABLE {0.63661977236758138, 1.5707963267948966e+00, 6.1232339957367574e-17, 8.4784276603688985e-32,            \
                        1.58969099521155010221e-10, -2.50507602534068634195e-08, 2.75573137070700676789e-06,                   \
                        -1.98412698298579493134e-04, 8.33333333332248946124e-03, -1.66666666666666324348e-01,                  \
                        -1.13596475577881948265e-11, 2.08757232129817482790e-09, -2.75573143513906633035e-07,                  \
                        2.48015872894767294178e-05, -1.38888888888741095749e-03, 4.16666666666666019037e-02}
static const double dbspm_sc_h[16] = DBSPM_SC_TABLE;
#if defined(__CUDACC__)
static __constant__ double dbspm_sc_d[16] = DBSPM_SC_TABLE;
#endif
#if defined(__CUDA_ARCH__)
#define SCC(i) dbspm_sc_d[i]
#else
#define SCC(i) dbspm_sc_h[i]
#endif

// reduced argument r in [-pi/4, pi/4] and quadrant q of x (|x| <= 1e5)
HD double dbspm_sc_reduce(double x, int *q)
{
    const double magic = 6755399441055744.0;                       // 1.5 * 2^52: (v + magic) - magic = rint(v)
    const double t = fma(x, SCC(0), magic);
    const double k = t - magic;
#if defined(__CUDA_ARCH__)
    *q = __double2loint(t);
#else
    long long tb; memcpy(&tb, &t, sizeof tb);
    *q = (int)(tb & 0xffffffffLL);
#endif
    double r = fma(-k, SCC(1), x);
    r = fma(-k, SCC(2), r);
    return fma(-k, SCC(3), r);
}

HD void dbspm_sc_kernel(double r, int q, double *sn, double *cs)
{
    const double z = r * r;
    double ps = SCC(4);
    ps = fma(ps, z, SCC(5));
    ps = fma(ps, z, SCC(6));
    ps = fma(ps, z, SCC(7));
    ps = fma(ps, z, SCC(8));
    ps = fma(ps, z, SCC(9));
    const double s0 = fma(r * z, ps, r);
    double pc = SCC(10);
    pc = fma(pc, z, SCC(11));
    pc = fma(pc, z, SCC(12));
    pc = fma(pc, z, SCC(13));
    pc = fma(pc, z, SCC(14));
    pc = fma(pc, z, SCC(15));
    const double hz = 0.5 * z, w = 1.0 - hz;                       // fdlibm's compensated 1 - z/2
    const double c0 = w + fma(z * z, pc, (1.0 - w) - hz);
    const double sa = (q & 1) ? c0 : s0, ca = (q & 1) ? s0 : c0;
    *sn = (q & 2) ? -sa : sa;
    *cs = ((q + 1) & 2) ? -ca : ca;
}

HD void dbspm_sincos(double x, double *sn, double *cs)
{
    if (!(fabs(x) <= 1.0e5)) {
#if defined(__CUDA_ARCH__)
        sincos(x, sn, cs);
#else
        *sn = sin(x); *cs = cos(x);
#endif
        return;
    }
    int q;
    const double r = dbspm_sc_reduce(x, &q);
    dbspm_sc_kernel(r, q, sn, cs);
}

// both angles of an evaluation in one basic block (one range test, the constants loaded once, four independent
// polynomial chains for the scheduler); same bits as two dbspm_sincos calls
HD void dbspm_sincos2(double a, double b, double *sa, double *ca, double *sb, double *cb)
{
    if (!(fabs(a) <= 1.0e5 && fabs(b) <= 1.0e5)) {
        dbspm_sincos(a, sa, ca);
        dbspm_sincos(b, sb, cb);
        return;
    }
    int qa, qb;
    const double ra = dbspm_sc_reduce(a, &qa), rb = dbspm_sc_reduce(b, &qb);
    dbspm_sc_kernel(ra, qa, sa, ca);
    dbspm_sc_kernel(rb, qb, sb, cb);
}

// ---- periodic Catmull-Rom (a = -1/2) tricubic stencil in index units ---------------------
// d = fmod(p, n), then "+= n if negative": the reference's periodic wrap, bit for bit.  fmod is
// exact by definition; q = trunc(p/n) computed with a reciprocal can be off by one, the
// fused p - q*n is still exact (Sterbenz) and one +-n correction restores fmod's range.
HD double wrap_coord(double p, double n, double inv_n)
{
    // inside the cell (every probe except those of columns within a lever length of the border, and the bottom
    // height, whose pivot plane index is -0.27): fmod(p, n) is p itself.  One almost always uniform branch instead
    // of a truncation, a fused multiply-add and nine selects on doubles
    if (p >= 0.0 && p < n) return p;
    const double q = trunc(p * inv_n);
    const double r = fma(-q, n, p);
    const double pos = (r < 0.0) ? r + n : ((r >= n) ? r - n : r);          // p >= 0: fmod in [0,n)
    double neg = (r > 0.0) ? r - n : ((r <= -n) ? r + n : r);               // p <  0: fmod in (-n,0]
    neg = (neg < 0.0) ? neg + n : neg;                 // the reference's `if (dx < 0) dx += n` (rounds)
    return (p >= 0.0) ? pos : neg;
}

// x / d for a divisor known in advance, correctly rounded (Markstein): with y = RN(1/d), q = RN(x*y) is
// within one ulp of x/d, r = x - q*d is exact in one fma, and RN(q + r*y) is the correctly rounded quotient
// (finite, non-extreme operands -- offsets of a few grid spacings here).  3 instructions instead of
// the ~20 of the general division; tests/test_emulation.py compares it with `/` on 10^7 random pairs.
HD double div_by(double x, double d, double inv_d)
{
    const double q = x * inv_d;
    const double r = fma(-q, d, x);
    return fma(r, inv_d, q);
}

// indices of the four stencil planes around floor(coordinate) = i0 in [0, n] (n itself when the
// reference's `+= n` rounds up), periodic; chained selects, valid for every n >= 1
HD void stencil_idx(int i0, int n, int *idx)
{
    const int i1 = (i0 >= n) ? i0 - n : i0;
    idx[1] = i1;
    idx[0] = (i1 == 0) ? n - 1 : i1 - 1;
    idx[2] = (i1 + 1 == n) ? 0 : i1 + 1;
    idx[3] = (idx[2] + 1 == n) ? 0 : idx[2] + 1;
}

// cell index and offset inside the cell
HD double cr_cell(double p, double nd, double inv_n, int *i0)
{
    const double d = wrap_coord(p, nd, inv_n);
    const double fl = floor(d);
    *i0 = (int)fl;
    return d - fl;
}

HD void cr_weights(double u, double *w)
{
    const double u2 = u * u, u3 = u2 * u;
    w[0] = -0.5 * u3 + u2 - 0.5 * u;
    w[1] = 1.5 * u3 - 2.5 * u2 + 1.0;
    w[2] = -1.5 * u3 + 2.0 * u2 + 0.5 * u;
    w[3] = 0.5 * u3 - 0.5 * u2;
}

HD void cr_axis(double p, int n, double inv_n, int *i0, double *w)
{
    cr_weights(cr_cell(p, (double)n, inv_n, i0), w);
}

#ifndef RELAX_PF
#define RELAX_PF 0     /* 1 / 2: prefetch.global.L1 / .L2 of the 16 stencil sectors as soon as their addresses are known (layout 2).
                          Measured on B200: 105 ms against 72.6 ms without -- the CCTL prefetches cost more than the waits they hide */
#endif

// LAYOUT 0: g[(x*ny + y)*nz + z] (the caller's array, z fastest)
// LAYOUT 1: g[(x*nz + z)*nyp + (y + 1)], nyp = ny + 3 (y fastest with a periodic halo: the 32 lanes of a
//           warp, which own adjacent y columns, read adjacent addresses -> 2-3 L1 wavefronts per load
//           instead of ~8, and the four y neighbours of a stencil row are four consecutive doubles at
//           immediate offsets from one base address, wrap or no wrap)
// LAYOUT 2: g[((x*nz + z)*ny + y)*4 + c] = value at (x, (y - 1 + c) mod ny, z): every stencil row's four y
//           neighbours are ONE aligned 32-byte sector, fetched by one 256-bit load (16 loads per evaluation instead
//           of 64).  Lanes of a warp mostly sit in different (x, z) cells (their tips are deflected differently), so a
//           load instruction of layout 1 touches about one 128-byte line per lane for 8 useful bytes; here the same
//           line delivers 32 (traced on the CPU emulation: 3.6x fewer L1 wavefronts per evaluation).  Costs four
//           times the memory of the window (1.8 GB for 1024 x 1024 x 54) -- HBM has it.
HD void ld_quad(const double *p, double *q)
{
#if defined(__CUDA_ARCH__)
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(q[0]), "=d"(q[1]), "=d"(q[2]), "=d"(q[3]) : "l"(p));
#else
    q[0] = p[0]; q[1] = p[1]; q[2] = p[2]; q[3] = p[3];
#endif
}

template <int LAYOUT>
HD double tricubic_eval(const double *g, int nx, int ny, int nz, int nyp, const double *nd, const double *inv_n, double x, double y, double z,
                        int x_base = 0, int nx_local = 0)
{
    int ix, iy, iz;
    double wx[4], wy[4], wz[4];
    const double ux = cr_cell(x, nd[0], inv_n[0], &ix);
    const double uy = cr_cell(y, nd[1], inv_n[1], &iy);
    const double uz = cr_cell(z, nd[2], inv_n[2], &iz);
    // 32-bit element offsets (the C ABI guarantees the element count < 2^31)
    int xo[4], yo[4], zo[4];
    stencil_idx(ix, nx, xo);
    stencil_idx(iz, nz, zo);
    if (nx_local) {
        // x-slab mode: global (periodically wrapped) plane index -> plane of the local slab.  The cell index and the
        // weights above come from the global coordinate, so the value is the one the whole window gives, bit for bit.
        // A plane outside the slab cannot occur when the halo covers the lever length; it is clamped, never read out of bounds.
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            int l = xo[c] - x_base;
            l = l < 0 ? l + nx : l;
            xo[c] = l < nx_local ? l : nx_local - 1;
        }
    }
    if (LAYOUT == 0) {
        stencil_idx(iy, ny, yo);
#pragma unroll
        for (int c = 0; c < 4; ++c) { xo[c] *= ny * nz; yo[c] *= nz; }
    } else if (LAYOUT == 2) {
        const int y1 = (iy >= ny) ? iy - ny : iy;
#pragma unroll
        for (int c = 0; c < 4; ++c) { xo[c] *= 4 * ny * nz; zo[c] *= 4 * ny; yo[c] = 4 * y1; }
#if defined(__CUDA_ARCH__) && RELAX_PF
        // The 16 loads below do not fit in the register file at once (128 registers of results): the compiler issues
        // them two or three at a time and every batch costs a trip to the L2 (ncu: five serialised waits per
        // evaluation).  Prefetches need no registers -- but measured slower (see RELAX_PF), so this is off.
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 4; ++c)
#if RELAX_PF == 2
                asm volatile("prefetch.global.L2 [%0];" ::"l"(g + (xo[a] + zo[c] + yo[0])));
#else
                asm volatile("prefetch.global.L1 [%0];" ::"l"(g + (xo[a] + zo[c] + yo[0])));
#endif
#endif
    } else {
        const int y1 = (iy >= ny) ? iy - ny : iy;          // stored index of plane y1 - 1
#pragma unroll
        for (int c = 0; c < 4; ++c) { xo[c] *= nyp * nz; zo[c] *= nyp; yo[c] = y1 + c; }
    }
    cr_weights(ux, wx);
    cr_weights(uy, wy);
    cr_weights(uz, wz);
    double r = 0.0;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        double rx = 0.0;
        // same summation order for both layouts (y innermost), so the result does not depend on
        // whether the caller supplied the re-layout workspace
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            double ry;
            if (LAYOUT == 0) {
                const int o = xo[a] + zo[c];
                ry = wy[0] * LDG(g + (o + yo[0]));
                ry = fma(wy[1], LDG(g + (o + yo[1])), ry);
                ry = fma(wy[2], LDG(g + (o + yo[2])), ry);
                ry = fma(wy[3], LDG(g + (o + yo[3])), ry);
            } else if (LAYOUT == 2) {
                double q[4];
                ld_quad(g + (xo[a] + zo[c] + yo[0]), q);
                ry = wy[0] * q[0];
                ry = fma(wy[1], q[1], ry);
                ry = fma(wy[2], q[2], ry);
                ry = fma(wy[3], q[3], ry);
            } else {
                const double *row = g + (xo[a] + zo[c] + yo[0]);
                ry = wy[0] * LDG(row);
                ry = fma(wy[1], LDG(row + 1), ry);
                ry = fma(wy[2], LDG(row + 2), ry);
                ry = fma(wy[3], LDG(row + 3), ry);
            }
            rx = fma(wz[c], ry, rx);
        }
        r = fma(wx[a], rx, r);
    }
    return r;
}

#if defined(DBSPM_EMU_INTERP_HOOK) && !defined(__CUDA_ARCH__)
extern double (*dbspm_emu_interp_hook)(double, double, double);
// test/analysis only: called before every objective evaluation with (column, height, entry state, theta, phi)
extern void (*dbspm_emu_trace_hook)(long long, int, int, double, double);
#endif

// ---- objective: Emin / Emin_norm ----------------------------------------------------------
template <int LAYOUT>
HD double relax_objective(const RelaxParams &P, int oi, int oj, int ok, double th, double az)
{
    const double d = M_PI - th;
    const double spr = 0.5 * P.kappa * (d * d);
    double sth, cth, saz, caz;
    dbspm_sincos2(th, az, &sth, &cth, &saz, &caz);
    double ox, oy, oz;
    if (!P.noortho) {
        ox = div_by(P.rpivot_rt * sth * caz, P.dr[0], P.inv_dr[0]);
        oy = div_by(P.rpivot_rt * sth * saz, P.dr[1], P.inv_dr[1]);
        oz = div_by(P.rpivot_rt * cth, P.dr[2], P.inv_dr[2]);
    } else {
        const double X = P.rpivot_rt * sth * caz, Y = P.rpivot_rt * sth * saz, Z = P.rpivot_rt * cth;
        ox = P.Minv[0] * X + P.Minv[1] * Y + P.Minv[2] * Z;
        oy = P.Minv[3] * X + P.Minv[4] * Y + P.Minv[5] * Z;
        oz = P.Minv[6] * X + P.Minv[7] * Y + P.Minv[8] * Z;
    }
#if defined(DBSPM_EMU_INTERP_HOOK) && !defined(__CUDA_ARCH__)
    // test-only: lets the CPU emulation swap in the oracle's interpolant so the optimiser
    // state machine can be compared bit for bit with the scipy restatement
    if (dbspm_emu_interp_hook)
        return dbspm_emu_interp_hook((double)oi + ox, (double)oj + oy, (double)ok + oz) + spr;
#endif
    return tricubic_eval<LAYOUT>(P.grid, P.nx, P.ny, P.nzc, P.nyp, P.nd, P.inv_n, (double)oi + ox, (double)oj + oy, (double)ok + oz,
                                 P.x_base, P.nx_local) + spr;
}

// ---- Powell / Brent / bracket state machine -----------------------------------------------
// A lane is always in one of the ENTRY states (a function value it asked for has arrived).  One
// trip of the column loop runs, for the whole warp together, a short dispatch on the entry state
// followed by a fixed sequence of blocks `if (pc == X) { ... }` in the order in which scipy's control
// flow visits them (bracket loop -> bracket check -> Brent iteration -> line-search finish ->
// direction bookkeeping -> iteration end -> store/next height -> iteration begin -> line-search
// begin -> request).  Every block is entered once per trip by all the lanes that need it, whatever
// state they came from, and the warp re-converges after each block; the rare backward edges (maxfev
// exhausted, a zero search direction) repeat the sequence for the lanes concerned.
enum RelaxPc {
    PC_IDLE = 0,                   // column finished
    // entry states
    PC_INIT,                       // f(x0)
    PC_BR_FB, PC_BR_FC,            // bracket: f(xb), f(xc)   (f(xa) is never requested, see PC_LS_BEGIN)
    PC_BR_A1, PC_BR_C1, PC_BR_GEN, // bracket loop: f(w) for the three kinds of trial point
    PC_BT_U,                       // Brent: f(u)
    PC_EXTRAP,                     // Powell: f(2x - x1)
    // blocks
    PC_BR_LOOP, PC_BR_DONE, PC_BT_LOOP, PC_LS_FINISH, PC_AFTER_LS, PC_ITER_END, PC_FINISHED,
    PC_ITER_BEGIN, PC_LS_BEGIN, PC_REQUEST
};

// Powell-level state, touched only at line-search boundaries: one shared-memory column per thread
// (field-major, so a warp's accesses are conflict-free).  Keeps ~28 registers free for occupancy.
enum RelaxShared {
    PS_X_0, PS_X_1,                // current point (theta, phi)
    PS_X1_0, PS_X1_1,              // point at the start of the iteration
    PS_D0_0, PS_D0_1, PS_D1_0, PS_D1_1,   // direction set
    PS_FVAL, PS_FX, PS_FX2, PS_DELTA,
    PS_COUNT,
    // optional residents (RELAX_SMEM_STATE): the bracket / Brent variables, then the line-search origin and direction
    PS_S0 = PS_COUNT, PS_P_0 = PS_S0 + 11, PS_P_1, PS_XI_0, PS_XI_1,
    PS_TOTAL
};

// Which optimiser state lives in the thread's shared-memory column instead of registers (occupancy against LDS/STS
// traffic): 0 = Powell-level state only; 1 = + the 11 bracket / Brent variables; 2 = + line-search origin and direction
#ifndef RELAX_SMEM_STATE
#define RELAX_SMEM_STATE 0
#endif

struct PowellRegs {
#if RELAX_SMEM_STATE < 2
    double p[2], xi[2];            // line search: origin and direction
#endif
#if RELAX_SMEM_STATE < 1
    double s[11];                  // bracket (xa xb xc fa fb fc w) and Brent (x w v fx fw fv a b deltax rat u) share storage
#endif
    int pc, which, bigind, iter, nfev, it;
};
#if RELAX_SMEM_STATE < 1
#define RS(i) R.s[i]
#else
#define RS(i) sm[(PS_S0 + (i)) * NTHR + tid]
#endif
#if RELAX_SMEM_STATE < 2
#define RP(i) R.p[i]
#define RXI(i) R.xi[i]
#else
#define RP(i) sm[(PS_P_0 + (i)) * NTHR + tid]
#define RXI(i) sm[(PS_XI_0 + (i)) * NTHR + tid]
#endif
// bracket view
#define BR_XA RS(0)
#define BR_XB RS(1)
#define BR_XC RS(2)
#define BR_FA RS(3)
#define BR_FB RS(4)
#define BR_FC RS(5)
#define BR_W  RS(6)
// Brent view
#define BT_X   RS(1)
#define BT_FX  RS(4)
#define BT_W   RS(0)
#define BT_V   RS(2)
#define BT_FW  RS(3)
#define BT_FV  RS(5)
#define BT_A   RS(6)
#define BT_B   RS(7)
#define BT_DX  RS(8)
#define BT_RAT RS(9)
#define BT_U   RS(10)

#ifndef RELAX_DRIFT
#define RELAX_DRIFT 0
#endif

// ---- which column a thread relaxes ------------------------------------------------------------
// foot_x = 1: thread g takes column g (a warp = 32 consecutive y).  foot_x > 1: warp w relaxes the foot_x x (32 / foot_x)
// patch of columns number w (patches in row-major order over the tile), lane = dx * fy + dy; lanes beyond a ragged edge
// of the tile idle (return ncol).  Column index = (i - i_lo) * nj + (j - j_lo), the row order of the outputs.
HD long long relax_patch_column(const RelaxParams &P, long long g, long long ncol)
{
    if (P.foot_x <= 1) return g;
    const int fy = 32 / P.foot_x, nj = P.j_hi - P.j_lo, ni = P.i_hi - P.i_lo;
    const long long w = g >> 5;
    const int lane = (int)(g & 31), pj = (nj + fy - 1) / fy;
    const long long s = w / pj;
    const long long ci = s * P.foot_x + lane / fy;
    const int cj = (int)(w - s * pj) * fy + lane % fy;
    return (ci < ni && cj < nj) ? ci * nj + cj : ncol;
}

// number of threads the patch order needs for a tile (a multiple of 32)
HD long long relax_patch_threads(int ni, int nj, int foot_x)
{
    if (foot_x <= 1) return (long long)ni * nj;
    const long long fy = 32 / foot_x;
    return (((long long)ni + foot_x - 1) / foot_x) * (((long long)nj + fy - 1) / fy) * 32;
}

// ---- one lateral column: sweep z from the top with warm start ------------------------------
// NTHR = threads per block (the stride of the shared-memory columns); sm = PS_TOTAL*NTHR doubles.
// k_hi, k_lo: the heights this call sweeps (k_hi, k_hi - 1, ..., k_lo); the start point of a height k_hi < nz_relax - 1 is the
// relaxed (theta, phi) of height k_hi + 1 in out_etp -- written by an earlier launch or, under the dynamic schedule of
// relax_dyn_kernel, by another warp (hence the L2-coherent load)
template <int LAYOUT, int NTHR>
HDN inline void relax_column_body(const RelaxParams &P, long long col, long long ncol, double *sm, int tid, int k_hi, int k_lo)
{
#define PS(f) sm[(f) * NTHR + tid]
    const double _gold = 1.618034, _verysmall_num = 1e-21, grow_limit = 110.0;
    const double _mintol = 1.0e-11, _cg = 0.3819660;
    const double tol = P.xtol * 100;
    // Every lane of the warp stays in the loop until all of them are finished (lanes past the
    // end of the tile start out finished), so the full-mask warp barrier below is legal.
    bool done = col >= ncol;
    const int nj = P.j_hi - P.j_lo;
    const long long c = done ? 0 : col;
    const int ci = (int)(c / nj), cj = (int)(c - (long long)ci * nj);
    const int oi = P.i_lo + ci, oj = P.j_lo + cj;
    int k = k_hi;
    int ok = (int)((double)k + P.npivot);                    // int64 store truncates
    PowellRegs R;
    // start of a minimisation at the current point (scipy: x = asarray(x0); fval = func(x)); the
    // request "p + 0*0" is x itself.  The top height starts from (pi, 0); a launch that continues a sweep starts
    // from the relaxed angles of the height above (warm start, interactions.py:83), written by the previous launch
    double th0 = M_PI, az0 = 0.0;
    if (!done && k < P.nz_relax - 1) {
        const size_t o = (size_t)col * P.nz_relax + (k + 1);
        th0 = LD_CG(P.out_etp + 3 * o + 1); az0 = LD_CG(P.out_etp + 3 * o + 2);
    }
    PS(PS_X_0) = th0; PS(PS_X_1) = az0;
    PS(PS_D0_0) = 1.0; PS(PS_D0_1) = 0.0; PS(PS_D1_0) = 0.0; PS(PS_D1_1) = 1.0;
    RP(0) = th0; RP(1) = az0; RXI(0) = 0.0; RXI(1) = 0.0;
    R.nfev = 1; R.iter = 0; R.which = 0; R.bigind = 0; R.it = 0;
    R.pc = done ? PC_IDLE : PC_INIT;
#pragma unroll
    for (int q = 0; q < 11; ++q) RS(q) = 0.0;
    double alpha = 0.0;
    for (;;) {
        // expensive part: the warp is converged here and all unfinished lanes evaluate the 64-point
        // stencil together
        WARP_SYNC();
        if (WARP_ALL(done)) break;
        double fe = 0.0;
#if defined(__CUDA_ARCH__)
        // A lane that has finished a height waits for the slowest lane of its warp before it starts the
        // next one (RELAX_DRIFT = heights it may run ahead).  The kernel is bound by L1 wavefronts: lanes on
        // the same height read the same z rows of the potential, which more than pays for the idle lanes
        // (B200, 1024x1024x24: no limit 118 ms, 2 heights 115 ms, 1 height 107 ms, lockstep 89.5 ms).
        const int kmax = __reduce_max_sync(0xffffffffu, done ? -1 : k);
        const bool hold = !done && R.pc == PC_INIT && k < kmax - RELAX_DRIFT;
#else
        const bool hold = false;
#endif
#if defined(DBSPM_EMU_INTERP_HOOK) && !defined(__CUDA_ARCH__)
        if (!done && !hold && dbspm_emu_trace_hook) dbspm_emu_trace_hook(col, k, R.pc, RP(0) + alpha * RXI(0), RP(1) + alpha * RXI(1));
#endif
        if (!done && !hold) fe = relax_objective<LAYOUT>(P, oi, oj, ok, RP(0) + alpha * RXI(0), RP(1) + alpha * RXI(1));
        WARP_SYNC();

        // ---- dispatch on the entry state (cheap, per state) ----
        int pc = hold ? PC_IDLE : R.pc, next = PC_IDLE;
        bool fresh = false;                  // request that starts a minimisation: no maxfev check
        double alpha_min = 0.0, fret = 0.0;  // line-search result on its way to PC_LS_FINISH
        if (pc == PC_BT_U) {
            const double fu = fe, u = BT_U;
            if (fu > BT_FX) {
                if (u < BT_X) BT_A = u; else BT_B = u;
                if ((fu <= BT_FW) || (BT_W == BT_X)) { BT_V = BT_W; BT_W = u; BT_FV = BT_FW; BT_FW = fu; }
                else if ((fu <= BT_FV) || (BT_V == BT_X) || (BT_V == BT_W)) { BT_V = u; BT_FV = fu; }
            } else {
                if (u >= BT_X) BT_A = BT_X; else BT_B = BT_X;
                BT_V = BT_W; BT_W = BT_X; BT_X = u;
                BT_FV = BT_FW; BT_FW = BT_FX; BT_FX = fu;
            }
            R.it += 1;
            pc = PC_BT_LOOP;
        } else if (pc == PC_BR_FB) {
            BR_FB = fe;
            if (BR_FA < BR_FB) {
                double t = BR_XA; BR_XA = BR_XB; BR_XB = t;
                t = BR_FA; BR_FA = BR_FB; BR_FB = t;
            }
            BR_XC = BR_XB + _gold * (BR_XB - BR_XA);
            alpha = BR_XC; next = PC_BR_FC; pc = PC_REQUEST;
        } else if (pc == PC_BR_FC) {
            BR_FC = fe;
            R.it = 0;
            pc = PC_BR_LOOP;
        } else if (pc == PC_BR_GEN) {
            BR_XA = BR_XB; BR_XB = BR_XC; BR_XC = BR_W;
            BR_FA = BR_FB; BR_FB = BR_FC; BR_FC = fe;
            pc = PC_BR_LOOP;
        } else if (pc == PC_BR_A1) {
            const double fw = fe;
            if (fw < BR_FC) {
                BR_XA = BR_XB; BR_XB = BR_W; BR_FA = BR_FB; BR_FB = fw;
                pc = PC_BR_DONE;
            } else if (fw > BR_FB) {
                BR_XC = BR_W; BR_FC = fw;
                pc = PC_BR_DONE;
            } else {
                BR_W = BR_XC + _gold * (BR_XC - BR_XB);
                alpha = BR_W; next = PC_BR_GEN; pc = PC_REQUEST;
            }
        } else if (pc == PC_BR_C1) {
            const double fw = fe;
            if (fw < BR_FC) {
                BR_XB = BR_XC; BR_XC = BR_W; BR_W = BR_XC + _gold * (BR_XC - BR_XB);
                BR_FB = BR_FC; BR_FC = fw;
                alpha = BR_W; next = PC_BR_GEN; pc = PC_REQUEST;
            } else {
                BR_XA = BR_XB; BR_XB = BR_XC; BR_XC = BR_W;
                BR_FA = BR_FB; BR_FB = BR_FC; BR_FC = fw;
                pc = PC_BR_LOOP;
            }
        } else if (pc == PC_EXTRAP) {
            const double fx2 = fe, fx = PS(PS_FX);
            PS(PS_FX2) = fx2;
            bool again = false;
            if (fx > fx2) {
                const double fval = PS(PS_FVAL), delta = PS(PS_DELTA);
                double t = 2.0 * (fx + fx2 - 2.0 * fval);
                double temp = (fx - fval - delta);
                t *= temp * temp;
                temp = fx - fx2;
                t -= delta * temp * temp;
                if (t < 0.0) again = true;
            }
            if (again) { R.which = 2; pc = PC_LS_BEGIN; }
            else pc = PC_ITER_BEGIN;
        } else if (pc == PC_INIT) {
            PS(PS_FVAL) = fe;
            PS(PS_X1_0) = PS(PS_X_0); PS(PS_X1_1) = PS(PS_X_1);
            pc = PC_ITER_BEGIN;
        }

        // ---- block sequence ----
        bool repeat;
        do {
            repeat = false;
            if (pc == PC_BR_LOOP) {
                if (!(BR_FC < BR_FB)) {
                    pc = PC_BR_DONE;
                } else {
                    const double tmp1 = (BR_XB - BR_XA) * (BR_FB - BR_FC);
                    const double tmp2 = (BR_XB - BR_XC) * (BR_FB - BR_FA);
                    const double val = tmp2 - tmp1;
                    double denom;
                    if (fabs(val) < _verysmall_num) denom = 2.0 * _verysmall_num;
                    else denom = 2.0 * val;
                    BR_W = BR_XB - ((BR_XB - BR_XC) * tmp2 - (BR_XB - BR_XA) * tmp1) / denom;
                    const double wlim = BR_XB + grow_limit * (BR_XC - BR_XB);
                    if (R.it > 1000) {
                        pc = PC_BR_DONE;
                    } else {
                        R.it += 1;
                        if ((BR_W - BR_XC) * (BR_XB - BR_W) > 0.0) {
                            next = PC_BR_A1;
                        } else if ((BR_W - wlim) * (wlim - BR_XC) >= 0.0) {
                            BR_W = wlim;
                            next = PC_BR_GEN;
                        } else if ((BR_W - wlim) * (BR_XC - BR_W) > 0.0) {
                            next = PC_BR_C1;
                        } else {
                            BR_W = BR_XC + _gold * (BR_XC - BR_XB);
                            next = PC_BR_GEN;
                        }
                        alpha = BR_W; pc = PC_REQUEST;
                    }
                }
            }
            if (pc == PC_BR_DONE) {
                const double xa = BR_XA, xb = BR_XB, xc = BR_XC, fa = BR_FA, fb = BR_FB, fc = BR_FC;
                const bool cond1 = (fb < fc && fb <= fa) || (fb < fa && fb <= fc);
                const bool cond2 = (xa < xb && xb < xc) || (xc < xb && xb < xa);
                const bool cond3 = isfinite(xa) && isfinite(xb) && isfinite(xc);
                if (!(cond1 && cond2 && cond3)) {
                    if (isnan(xa) || isnan(xb) || isnan(xc) || isnan(fa) || isnan(fb) || isnan(fc)) {
                        alpha_min = NAN; fret = NAN;
                    } else {
                        alpha_min = xa; fret = fa;                    /* np.argmin: first minimum */
                        if (fb < fret) { alpha_min = xb; fret = fb; }
                        if (fc < fret) { alpha_min = xc; fret = fc; }
                    }
                    pc = PC_LS_FINISH;
                } else {
                    BT_X = xb; BT_W = xb; BT_V = xb;
                    BT_FX = fb; BT_FW = fb; BT_FV = fb;
                    if (xa < xc) { BT_A = xa; BT_B = xc; } else { BT_A = xc; BT_B = xa; }
                    BT_DX = 0.0; BT_RAT = 0.0;
                    R.it = 0;
                    pc = PC_BT_LOOP;
                }
            }
            if (pc == PC_BT_LOOP) {
                const double x = BT_X;
                const double tol1 = tol * fabs(x) + _mintol;
                const double tol2 = 2.0 * tol1;
                const double xmid = 0.5 * (BT_A + BT_B);
                if (!(R.it < 500) || fabs(x - xmid) < (tol2 - 0.5 * (BT_B - BT_A))) {
                    alpha_min = x; fret = BT_FX;
                    pc = PC_LS_FINISH;
                } else {
                    if (fabs(BT_DX) <= tol1) {
                        if (x >= xmid) BT_DX = BT_A - x; else BT_DX = BT_B - x;
                        BT_RAT = _cg * BT_DX;
                    } else {
                        double tmp1 = (x - BT_W) * (BT_FX - BT_FV);
                        double tmp2 = (x - BT_V) * (BT_FX - BT_FW);
                        double pp = (x - BT_V) * tmp2 - (x - BT_W) * tmp1;
                        tmp2 = 2.0 * (tmp2 - tmp1);
                        if (tmp2 > 0.0) pp = -pp;
                        tmp2 = fabs(tmp2);
                        const double dx_temp = BT_DX;
                        BT_DX = BT_RAT;
                        if ((pp > tmp2 * (BT_A - x)) && (pp < tmp2 * (BT_B - x)) && (fabs(pp) < fabs(0.5 * tmp2 * dx_temp))) {
                            BT_RAT = pp * 1.0 / tmp2;
                            const double uu = x + BT_RAT;
                            if ((uu - BT_A) < tol2 || (BT_B - uu) < tol2) {
                                if (xmid - x >= 0) BT_RAT = tol1; else BT_RAT = -tol1;
                            }
                        } else {
                            if (x >= xmid) BT_DX = BT_A - x; else BT_DX = BT_B - x;
                            BT_RAT = _cg * BT_DX;
                        }
                    }
                    if (fabs(BT_RAT) < tol1) { if (BT_RAT >= 0) BT_U = x + tol1; else BT_U = x - tol1; }
                    else BT_U = x + BT_RAT;
                    alpha = BT_U; next = PC_BT_U; pc = PC_REQUEST;
                }
            }
            if (pc == PC_LS_FINISH) {
                RXI(0) = alpha_min * RXI(0); RXI(1) = alpha_min * RXI(1);
                PS(PS_X_0) = RP(0) + RXI(0); PS(PS_X_1) = RP(1) + RXI(1);
                PS(PS_FVAL) = fret;
                pc = PC_AFTER_LS;
            }
            if (pc == PC_AFTER_LS) {
                if (R.which < 2) {
                    const double gain = PS(PS_FX2) - PS(PS_FVAL);
                    if (gain > PS(PS_DELTA)) { PS(PS_DELTA) = gain; R.bigind = R.which; }
                    if (R.which == 0) { R.which = 1; pc = PC_LS_BEGIN; }
                    else pc = PC_ITER_END;
                } else {
                    if (RXI(0) != 0.0 || RXI(1) != 0.0) {
                        if (R.bigind == 0) { PS(PS_D0_0) = PS(PS_D1_0); PS(PS_D0_1) = PS(PS_D1_1); }
                        PS(PS_D1_0) = RXI(0); PS(PS_D1_1) = RXI(1);
                    }
                    pc = PC_ITER_BEGIN;
                }
            }
            if (pc == PC_ITER_END) {
                R.iter += 1;
                const double fx = PS(PS_FX), fval = PS(PS_FVAL);
                const double bnd = P.ftol * (fabs(fx) + fabs(fval)) + 1e-20;
                if (2.0 * (fx - fval) <= bnd || R.nfev >= P.maxfev || R.iter >= P.maxiter || (isnan(fx) && isnan(fval))) {
                    pc = PC_FINISHED;
                } else {
                    const double x0 = PS(PS_X_0), x1 = PS(PS_X_1);
                    RXI(0) = x0 - PS(PS_X1_0); RXI(1) = x1 - PS(PS_X1_1);         /* direc1 */
                    PS(PS_X1_0) = x0; PS(PS_X1_1) = x1;
                    RP(0) = x0; RP(1) = x1;
                    alpha = 1.0; next = PC_EXTRAP; pc = PC_REQUEST;                   /* x + 1*direc1 */
                }
            }
            if (pc == PC_FINISHED) {
                // store this height's row and start the next one from the relaxed angles (warm
                // start), so lanes do not wait for the slowest lane at every height
                const size_t o = (size_t)col * P.nz_relax + k;
                const double x0 = PS(PS_X_0), x1 = PS(PS_X_1);
                P.out_etp[3 * o + 0] = PS(PS_FVAL);
                P.out_etp[3 * o + 1] = x0;
                P.out_etp[3 * o + 2] = x1;
                if (P.out_nfev) P.out_nfev[o] = R.nfev;
                if (--k < k_lo) {
                    done = true; pc = PC_IDLE;
                } else {
                    ok = (int)((double)k + P.npivot);
                    R.nfev = 0; R.iter = 0;
                    PS(PS_D0_0) = 1.0; PS(PS_D0_1) = 0.0; PS(PS_D1_0) = 0.0; PS(PS_D1_1) = 1.0;
                    RP(0) = x0; RP(1) = x1; RXI(0) = 0.0; RXI(1) = 0.0;
                    alpha = 0.0; next = PC_INIT; fresh = true; pc = PC_REQUEST;
                }
            }
            if (pc == PC_ITER_BEGIN) {
                PS(PS_FX) = PS(PS_FVAL); R.bigind = 0; PS(PS_DELTA) = 0.0;
                R.which = 0;
                pc = PC_LS_BEGIN;
            }
            if (pc == PC_LS_BEGIN) {
                if (R.which == 0) { RXI(0) = PS(PS_D0_0); RXI(1) = PS(PS_D0_1); }
                else if (R.which == 1) { RXI(0) = PS(PS_D1_0); RXI(1) = PS(PS_D1_1); }
                /* which == 2: xi already holds the extrapolation direction */
                if (R.which < 2) PS(PS_FX2) = PS(PS_FVAL);
                if (RXI(0) == 0.0 && RXI(1) == 0.0) {
                    pc = PC_AFTER_LS; repeat = true;
                } else {
                    RP(0) = PS(PS_X_0); RP(1) = PS(PS_X_1);
                    BR_XA = 0.0; BR_XB = 1.0;
                    // bracket's first call is f(p + 0*xi) = f(x), the point whose value is fval: the
                    // objective is a pure function, so the call is counted (the maxfun wrapper may
                    // raise on it) but not recomputed -- one evaluation in ten saved, same bits
                    if (R.nfev >= P.maxfev) { pc = PC_FINISHED; repeat = true; }
                    else {
                        R.nfev += 1;
                        BR_FA = PS(PS_FVAL);
                        alpha = 1.0; next = PC_BR_FB; pc = PC_REQUEST;
                    }
                }
            }
            if (pc == PC_REQUEST) {
                /* maxfun wrapper: the call that would exceed maxfev raises instead (-> exit with
                   the x, fval of the last completed line search) */
                if (!fresh && R.nfev >= P.maxfev) { pc = PC_FINISHED; repeat = true; }
                else { R.nfev += 1; pc = next; }
            }
        } while (repeat);
        if (!hold) R.pc = pc;
    }
#undef PS
}

// sph2cart of every relaxed position (dbspm:603-605), one converged pass instead of divergent
// tail work inside the column loop
HD void relax_cart_body(const double *etp, double *cart, long long npos, double rpivot, long long gid, long long gstride)
{
    for (long long o = gid; o < npos; o += gstride) {
        const double th = etp[3 * o + 1], az = etp[3 * o + 2];
        double sth, cth, saz, caz;
        dbspm_sincos2(th, az, &sth, &cth, &saz, &caz);
        cart[o] = rpivot * sth * caz;
        cart[npos + o] = rpivot * sth * saz;
        cart[2 * npos + o] = rpivot * cth;
    }
}

// (nx, ny, nz) -> (nx, nz, ny, 4) with quad y holding the values at y - 1, y, y + 1, y + 2 (periodic): a tile of
// 32 y-quads x 32 z through shared memory (tile[35][33]: 32 + 3 halo rows), one x-slab per blockIdx.z
HD void quad_yz_tile(const double *in, double *out, int ny, int nz, double *tile /* 35*33 */,
                     int bx /* z tile */, int by /* y tile */, long long slab, int tx, int ty, int nty)
{
    const double *src = in + (size_t)slab * ny * nz;
    double *dst = out + (size_t)slab * ny * nz * 4;
    for (int r = ty; r < 35; r += nty) {
        const int y = by * 32 + r - 1, z = bx * 32 + tx;               // rows y0 - 1 ... y0 + 33
        if (z < nz) tile[r * 33 + tx] = src[(size_t)(((y % ny) + ny) % ny) * nz + z];
    }
    BLOCK_SYNC();
    for (int r = ty; r < 32; r += nty) {
        const int z = bx * 32 + r, y = by * 32 + tx;
        if (y < ny && z < nz) {
            double *d = dst + ((size_t)z * ny + y) * 4;
#if defined(__CUDA_ARCH__)
            reinterpret_cast<double2 *>(d)[0] = make_double2(tile[tx * 33 + r], tile[(tx + 1) * 33 + r]);
            reinterpret_cast<double2 *>(d)[1] = make_double2(tile[(tx + 2) * 33 + r], tile[(tx + 3) * 33 + r]);
#else
            for (int c = 0; c < 4; ++c) d[c] = tile[(tx + c) * 33 + r];
#endif
        }
    }
}

// (nx, ny, nz) -> (nx, nz, nyp = ny + 3) with stored index t holding y = (t - 1) mod ny: 32x32 tiles
// through shared memory (tile[32][33]), one x-slab per blockIdx.z; `by` tiles the stored index t
HD void transpose_yz_tile(const double *in, double *out, int ny, int nz, double *tile /* 32*33 */,
                          int bx /* z tile */, int by /* t tile */, long long slab, int tx, int ty, int nty)
{
    const int nyp = ny + 3;
    const double *src = in + (size_t)slab * ny * nz;
    double *dst = out + (size_t)slab * nyp * nz;
    for (int r = ty; r < 32; r += nty) {
        const int t = by * 32 + r, z = bx * 32 + tx;
        if (t < nyp && z < nz) tile[r * 33 + tx] = src[(size_t)((t - 1 + ny) % ny) * nz + z];
    }
    BLOCK_SYNC();
    for (int r = ty; r < 32; r += nty) {
        const int z = bx * 32 + r, t = by * 32 + tx;
        if (t < nyp && z < nz) dst[(size_t)z * nyp + t] = tile[tx * 33 + r];
    }
}
