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
#include "hd.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

struct RelaxParams {
    const double *grid;       // static potential: (nx,ny,nzc) z fastest [LAYOUT 0] or (nx,nzc,ny) [LAYOUT 1]
    int nx, ny, nzc;
    int i_lo, i_hi, j_lo, j_hi;
    int nz_relax;
    double kappa;
    double rpivot;            // params.RPIVOT (used by sph2cart, dbspm:605)
    double npivot;            // rpivot / dr[2]                     (dbspm:558)
    double rpivot_rt;         // npivot * dr[2]                     (interactions.py:73)
    double dr[3];
    double inv_n[3];          // 1/nx, 1/ny, 1/nzc
    int noortho;
    double Minv[9];           // inv(dR^T), row-major
    double xtol, ftol;
    int maxfev, maxiter;
    double *out_etp;          // (ni,nj,nz,3)
    double *out_cart;         // (3,ni,nj,nz) or null
    int *out_nfev;            // (ni,nj,nz) or null
};

// ---- periodic Catmull-Rom (a = -1/2) tricubic stencil in index units ---------------------
// d = fmod(p, n), then "+= n if negative": the reference's periodic wrap, bit for bit.  fmod is
// exact by definition; q = trunc(p/n) computed with a reciprocal can be off by one, the
// fused p - q*n is still exact (Sterbenz) and one +-n correction restores fmod's range.
HD double wrap_coord(double p, double n, double inv_n)
{   // selects only, no branches: the 64-point stencil that follows must stay warp-converged
    const double q = trunc(p * inv_n);
    const double r = fma(-q, n, p);
    const double pos = (r < 0.0) ? r + n : ((r >= n) ? r - n : r);          // p >= 0: fmod in [0,n)
    double neg = (r > 0.0) ? r - n : ((r <= -n) ? r + n : r);               // p <  0: fmod in (-n,0]
    neg = (neg < 0.0) ? neg + n : neg;                 // the reference's `if (dx < 0) dx += n` (rounds)
    return (p >= 0.0) ? pos : neg;
}

HD void cr_axis(double p, int n, double inv_n, int *i0, double *w)
{
#if defined(RELAX_OLD_WRAP)
    double d = fmod(p, (double)n);
    if (d < 0) d += (double)n;
#else
    const double d = wrap_coord(p, (double)n, inv_n);
#endif
    const double fl = floor(d);
    const double u = d - fl, u2 = u * u, u3 = u2 * u;
    *i0 = (int)fl;
    w[0] = -0.5 * u3 + u2 - 0.5 * u;
    w[1] = 1.5 * u3 - 2.5 * u2 + 1.0;
    w[2] = -1.5 * u3 + 2.0 * u2 + 0.5 * u;
    w[3] = 0.5 * u3 - 0.5 * u2;
}

HD int wrap_idx(int i, int n)
{   // i in [-1, n+2] (stencil around a wrapped coordinate): selects instead of integer division;
    // tiny grids (n < 4) can need more than one correction and take the modulo (uniform branch)
#if defined(RELAX_OLD_WRAP)
    i %= n;
    return i < 0 ? i + n : i;
#else
    if (n < 4) {
        i %= n;
        return i < 0 ? i + n : i;
    }
    i = (i < 0) ? i + n : i;
    return (i >= n) ? i - n : i;
#endif
}

// LAYOUT 0: g[(x*ny + y)*nz + z] (the caller's array, z fastest)
// LAYOUT 1: g[(x*nz + z)*ny + y] (y fastest: the 32 lanes of a warp, which own adjacent y columns,
//           read adjacent addresses -> 2-3 L1 wavefronts per load instead of ~8 and a 5x smaller
//           per-warp cache footprint)
template <int LAYOUT>
HD double tricubic_eval(const double *g, int nx, int ny, int nz, const double *inv_n, double x, double y, double z)
{
    int ix, iy, iz;
    double wx[4], wy[4], wz[4];
    cr_axis(x, nx, inv_n[0], &ix, wx);
    cr_axis(y, ny, inv_n[1], &iy, wy);
    cr_axis(z, nz, inv_n[2], &iz, wz);
    // 32-bit element offsets (the C ABI guarantees nx*ny*nz < 2^31): one 3-input add and one
    // widening multiply-add per load instead of 64-bit pointer arithmetic
    int xo[4], yo[4], zo[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        xo[c] = wrap_idx(ix - 1 + c, nx) * (ny * nz);
        yo[c] = wrap_idx(iy - 1 + c, ny) * (LAYOUT == 0 ? nz : 1);
        zo[c] = wrap_idx(iz - 1 + c, nz) * (LAYOUT == 0 ? 1 : ny);
    }
    double r = 0.0;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        double rx = 0.0;
        // same summation order for both layouts (y innermost), so the result does not depend on
        // whether the caller supplied the transposition workspace
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int o = xo[a] + zo[c];
            double ry = wy[0] * LDG(g + (o + yo[0]));
            ry = fma(wy[1], LDG(g + (o + yo[1])), ry);
            ry = fma(wy[2], LDG(g + (o + yo[2])), ry);
            ry = fma(wy[3], LDG(g + (o + yo[3])), ry);
            rx = fma(wz[c], ry, rx);
        }
        r = fma(wx[a], rx, r);
    }
    return r;
}

#if defined(DBSPM_EMU_INTERP_HOOK) && !defined(__CUDA_ARCH__)
extern double (*dbspm_emu_interp_hook)(double, double, double);
#endif

// ---- objective: Emin / Emin_norm ----------------------------------------------------------
template <int LAYOUT>
HD double relax_objective(const RelaxParams &P, int oi, int oj, int ok, double th, double az)
{
    const double d = M_PI - th;
    const double spr = 0.5 * P.kappa * (d * d);
    double sth, cth, saz, caz;
#if defined(__CUDA_ARCH__)
    sincos(th, &sth, &cth);
    sincos(az, &saz, &caz);
#else
    sth = sin(th); cth = cos(th); saz = sin(az); caz = cos(az);
#endif
    double ox, oy, oz;
    if (!P.noortho) {
        ox = P.rpivot_rt * sth * caz / P.dr[0];
        oy = P.rpivot_rt * sth * saz / P.dr[1];
        oz = P.rpivot_rt * cth / P.dr[2];
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
    return tricubic_eval<LAYOUT>(P.grid, P.nx, P.ny, P.nzc, P.inv_n, (double)oi + ox, (double)oj + oy, (double)ok + oz) + spr;
}

// ---- Powell / Brent / bracket state machine -----------------------------------------------
enum RelaxState {
    RS_START = 0,      // begin a minimisation at the current guess
    RS_INIT,           // f(x0) arrived
    RS_ITER_BEGIN,
    RS_LS_BEGIN,       // start line search `which`
    RS_BR_FA, RS_BR_FB, RS_BR_FC,
    RS_BR_LOOP,
    RS_BR_A1, RS_BR_C1, RS_BR_GEN,
    RS_BR_DONE,
    RS_BT_LOOP,        // Brent loop head
    RS_BT_U,           // f(u) arrived
    RS_LS_FINISH,
    RS_AFTER_LS,
    RS_ITER_END,
    RS_EXTRAP,         // f(x2) arrived
    RS_FINISHED
};

struct PowellMachine {
    // Powell
    double x[2], x1[2], d0[2], d1[2], fval, fx, fx2, delta;
    int bigind, iter, nfev, which, state;
    // line search
    double p[2], xi[2], alpha_min, fret;
    // bracket
    double xa, xb, xc, fa, fb, fc, w, fw;
    int br_iter;
    // Brent
    double bx, bw, bv, bfx, bfw, bfv, a, b, deltax, rat, u;
    int bt_iter;
    // evaluation request
    double ex[2];
};

HD void pm_request_alpha(PowellMachine &M, double alpha, int next)
{
    M.ex[0] = M.p[0] + alpha * M.xi[0];
    M.ex[1] = M.p[1] + alpha * M.xi[1];
    M.state = next;
}

// Advance until an objective value is needed (returns true, point in M.ex) or the
// minimisation has finished (returns false; result in M.x, M.fval).  `fe` is the value of
// the previously requested point.
HDN inline bool powell_advance(PowellMachine &M, double fe, double xtol, double ftol, int maxiter, int maxfev)
{
    const double _gold = 1.618034, _verysmall_num = 1e-21, grow_limit = 110.0;
    const double _mintol = 1.0e-11, _cg = 0.3819660;
    const double tol = xtol * 100;
    for (;;) {
        switch (M.state) {
        case RS_START:
            M.nfev = 0; M.iter = 0;
            M.d0[0] = 1.0; M.d0[1] = 0.0; M.d1[0] = 0.0; M.d1[1] = 1.0;
            M.ex[0] = M.x[0]; M.ex[1] = M.x[1];
            M.state = RS_INIT;
            M.nfev += 1;
            return true;
        case RS_INIT:
            M.fval = fe;
            M.x1[0] = M.x[0]; M.x1[1] = M.x[1];
            goto L_ITER_BEGIN;
        case RS_ITER_BEGIN:
        L_ITER_BEGIN:
            M.fx = M.fval; M.bigind = 0; M.delta = 0.0;
            M.which = 0;
            goto L_LS_BEGIN;
        case RS_LS_BEGIN:
        L_LS_BEGIN:
            if (M.which == 0) { M.xi[0] = M.d0[0]; M.xi[1] = M.d0[1]; }
            else if (M.which == 1) { M.xi[0] = M.d1[0]; M.xi[1] = M.d1[1]; }
            /* which == 2: xi already holds the extrapolation direction */
            if (M.which < 2) M.fx2 = M.fval;
            if (M.xi[0] == 0.0 && M.xi[1] == 0.0) { goto L_AFTER_LS; }
            M.p[0] = M.x[0]; M.p[1] = M.x[1];
            M.xa = 0.0; M.xb = 1.0;
            pm_request_alpha(M, M.xa, RS_BR_FA);
            goto request;
        case RS_BR_FA:
            M.fa = fe;
            pm_request_alpha(M, M.xb, RS_BR_FB);
            goto request;
        case RS_BR_FB:
            M.fb = fe;
            if (M.fa < M.fb) {
                double t = M.xa; M.xa = M.xb; M.xb = t;
                t = M.fa; M.fa = M.fb; M.fb = t;
            }
            M.xc = M.xb + _gold * (M.xb - M.xa);
            pm_request_alpha(M, M.xc, RS_BR_FC);
            goto request;
        case RS_BR_FC:
            M.fc = fe;
            M.br_iter = 0;
            goto L_BR_LOOP;
        case RS_BR_LOOP:
        L_BR_LOOP: {
            if (!(M.fc < M.fb)) { goto L_BR_DONE; }
            const double tmp1 = (M.xb - M.xa) * (M.fb - M.fc);
            const double tmp2 = (M.xb - M.xc) * (M.fb - M.fa);
            const double val = tmp2 - tmp1;
            double denom;
            if (fabs(val) < _verysmall_num) denom = 2.0 * _verysmall_num;
            else denom = 2.0 * val;
            M.w = M.xb - ((M.xb - M.xc) * tmp2 - (M.xb - M.xa) * tmp1) / denom;
            const double wlim = M.xb + grow_limit * (M.xc - M.xb);
            if (M.br_iter > 1000) { goto L_BR_DONE; }
            M.br_iter += 1;
            if ((M.w - M.xc) * (M.xb - M.w) > 0.0) {
                pm_request_alpha(M, M.w, RS_BR_A1);
            } else if ((M.w - wlim) * (wlim - M.xc) >= 0.0) {
                M.w = wlim;
                pm_request_alpha(M, M.w, RS_BR_GEN);
            } else if ((M.w - wlim) * (M.xc - M.w) > 0.0) {
                pm_request_alpha(M, M.w, RS_BR_C1);
            } else {
                M.w = M.xc + _gold * (M.xc - M.xb);
                pm_request_alpha(M, M.w, RS_BR_GEN);
            }
            goto request;
        }
        case RS_BR_A1:
            M.fw = fe;
            if (M.fw < M.fc) {
                M.xa = M.xb; M.xb = M.w; M.fa = M.fb; M.fb = M.fw;
                goto L_BR_DONE;
            } else if (M.fw > M.fb) {
                M.xc = M.w; M.fc = M.fw;
                goto L_BR_DONE;
            }
            M.w = M.xc + _gold * (M.xc - M.xb);
            pm_request_alpha(M, M.w, RS_BR_GEN);
            goto request;
        case RS_BR_C1:
            M.fw = fe;
            if (M.fw < M.fc) {
                M.xb = M.xc; M.xc = M.w; M.w = M.xc + _gold * (M.xc - M.xb);
                M.fb = M.fc; M.fc = M.fw;
                pm_request_alpha(M, M.w, RS_BR_GEN);
                goto request;
            }
            M.xa = M.xb; M.xb = M.xc; M.xc = M.w;
            M.fa = M.fb; M.fb = M.fc; M.fc = M.fw;
            goto L_BR_LOOP;
        case RS_BR_GEN:
            M.fw = fe;
            M.xa = M.xb; M.xb = M.xc; M.xc = M.w;
            M.fa = M.fb; M.fb = M.fc; M.fc = M.fw;
            goto L_BR_LOOP;
        case RS_BR_DONE:
        L_BR_DONE: {
            const bool cond1 = (M.fb < M.fc && M.fb <= M.fa) || (M.fb < M.fa && M.fb <= M.fc);
            const bool cond2 = (M.xa < M.xb && M.xb < M.xc) || (M.xc < M.xb && M.xb < M.xa);
            const bool cond3 = isfinite(M.xa) && isfinite(M.xb) && isfinite(M.xc);
            if (!(cond1 && cond2 && cond3)) {
                if (isnan(M.xa) || isnan(M.xb) || isnan(M.xc) || isnan(M.fa) || isnan(M.fb) || isnan(M.fc)) {
                    M.alpha_min = NAN; M.fret = NAN;
                } else {
                    M.alpha_min = M.xa; M.fret = M.fa;                /* np.argmin: first minimum */
                    if (M.fb < M.fret) { M.alpha_min = M.xb; M.fret = M.fb; }
                    if (M.fc < M.fret) { M.alpha_min = M.xc; M.fret = M.fc; }
                }
                goto L_LS_FINISH;
            }
            M.bx = M.bw = M.bv = M.xb;
            M.bfw = M.bfv = M.bfx = M.fb;
            if (M.xa < M.xc) { M.a = M.xa; M.b = M.xc; } else { M.a = M.xc; M.b = M.xa; }
            M.deltax = 0.0; M.rat = 0.0;
            M.bt_iter = 0;
            goto L_BT_LOOP;
        }
        case RS_BT_LOOP:
        L_BT_LOOP: {
            if (!(M.bt_iter < 500)) { M.alpha_min = M.bx; M.fret = M.bfx; goto L_LS_FINISH; }
            const double x = M.bx;
            const double tol1 = tol * fabs(x) + _mintol;
            const double tol2 = 2.0 * tol1;
            const double xmid = 0.5 * (M.a + M.b);
            if (fabs(x - xmid) < (tol2 - 0.5 * (M.b - M.a))) {
                M.alpha_min = M.bx; M.fret = M.bfx; goto L_LS_FINISH;
            }
            if (fabs(M.deltax) <= tol1) {
                if (x >= xmid) M.deltax = M.a - x; else M.deltax = M.b - x;
                M.rat = _cg * M.deltax;
            } else {
                double tmp1 = (x - M.bw) * (M.bfx - M.bfv);
                double tmp2 = (x - M.bv) * (M.bfx - M.bfw);
                double pp = (x - M.bv) * tmp2 - (x - M.bw) * tmp1;
                tmp2 = 2.0 * (tmp2 - tmp1);
                if (tmp2 > 0.0) pp = -pp;
                tmp2 = fabs(tmp2);
                const double dx_temp = M.deltax;
                M.deltax = M.rat;
                if ((pp > tmp2 * (M.a - x)) && (pp < tmp2 * (M.b - x)) && (fabs(pp) < fabs(0.5 * tmp2 * dx_temp))) {
                    M.rat = pp * 1.0 / tmp2;
                    const double uu = x + M.rat;
                    if ((uu - M.a) < tol2 || (M.b - uu) < tol2) {
                        if (xmid - x >= 0) M.rat = tol1; else M.rat = -tol1;
                    }
                } else {
                    if (x >= xmid) M.deltax = M.a - x; else M.deltax = M.b - x;
                    M.rat = _cg * M.deltax;
                }
            }
            if (fabs(M.rat) < tol1) { if (M.rat >= 0) M.u = x + tol1; else M.u = x - tol1; }
            else M.u = x + M.rat;
            pm_request_alpha(M, M.u, RS_BT_U);
            goto request;
        }
        case RS_BT_U: {
            const double fu = fe, u = M.u;
            if (fu > M.bfx) {
                if (u < M.bx) M.a = u; else M.b = u;
                if ((fu <= M.bfw) || (M.bw == M.bx)) { M.bv = M.bw; M.bw = u; M.bfv = M.bfw; M.bfw = fu; }
                else if ((fu <= M.bfv) || (M.bv == M.bx) || (M.bv == M.bw)) { M.bv = u; M.bfv = fu; }
            } else {
                if (u >= M.bx) M.a = M.bx; else M.b = M.bx;
                M.bv = M.bw; M.bw = M.bx; M.bx = u;
                M.bfv = M.bfw; M.bfw = M.bfx; M.bfx = fu;
            }
            M.bt_iter += 1;
            goto L_BT_LOOP;
        }
        case RS_LS_FINISH:
        L_LS_FINISH:
            M.xi[0] = M.alpha_min * M.xi[0]; M.xi[1] = M.alpha_min * M.xi[1];
            M.x[0] = M.p[0] + M.xi[0]; M.x[1] = M.p[1] + M.xi[1];
            M.fval = M.fret;
            goto L_AFTER_LS;
        case RS_AFTER_LS:
        L_AFTER_LS:
            if (M.which < 2) {
                if ((M.fx2 - M.fval) > M.delta) { M.delta = M.fx2 - M.fval; M.bigind = M.which; }
                if (M.which == 0) { M.which = 1; goto L_LS_BEGIN; }
                goto L_ITER_END;
            } else {
                if (M.xi[0] != 0.0 || M.xi[1] != 0.0) {
                    if (M.bigind == 0) { M.d0[0] = M.d1[0]; M.d0[1] = M.d1[1]; }
                    M.d1[0] = M.xi[0]; M.d1[1] = M.xi[1];
                }
                goto L_ITER_BEGIN;
            }
        case RS_ITER_END:
        L_ITER_END: {
            M.iter += 1;
            const double bnd = ftol * (fabs(M.fx) + fabs(M.fval)) + 1e-20;
            if (2.0 * (M.fx - M.fval) <= bnd) { M.state = RS_FINISHED; return false; }
            if (M.nfev >= maxfev) { M.state = RS_FINISHED; return false; }
            if (M.iter >= maxiter) { M.state = RS_FINISHED; return false; }
            if (isnan(M.fx) && isnan(M.fval)) { M.state = RS_FINISHED; return false; }
            M.xi[0] = M.x[0] - M.x1[0]; M.xi[1] = M.x[1] - M.x1[1];     /* direc1 */
            M.x1[0] = M.x[0]; M.x1[1] = M.x[1];
            M.ex[0] = M.x[0] + 1 * M.xi[0]; M.ex[1] = M.x[1] + 1 * M.xi[1];
            M.state = RS_EXTRAP;
            goto request;
        }
        case RS_EXTRAP: {
            M.fx2 = fe;
            bool again = false;
            if (M.fx > M.fx2) {
                double t = 2.0 * (M.fx + M.fx2 - 2.0 * M.fval);
                double temp = (M.fx - M.fval - M.delta);
                t *= temp * temp;
                temp = M.fx - M.fx2;
                t -= M.delta * temp * temp;
                if (t < 0.0) again = true;
            }
            if (again) { M.which = 2; goto L_LS_BEGIN; }
            goto L_ITER_BEGIN;
        }
        default:
            return false;
        }
        continue;
    request:
        /* maxfun wrapper: the call that would exceed maxfev raises instead (-> loop exit
           with the x, fval of the last completed line search) */
        if (M.nfev >= maxfev) { M.state = RS_FINISHED; return false; }
        M.nfev += 1;
        return true;
    }
}

// ---- one lateral column: sweep z from the top with warm start ------------------------------
template <int LAYOUT>
HDN inline void relax_column_body(const RelaxParams &P, long long col, long long ncol)
{
    // Every lane of the warp stays in the loop until all of them are finished (lanes past the
    // end of the tile start out finished), so the full-mask warp barrier below is legal.
    bool done = col >= ncol;
    const int nj = P.j_hi - P.j_lo;
    const long long c = done ? 0 : col;
    const int ci = (int)(c / nj), cj = (int)(c - (long long)ci * nj);
    const int oi = P.i_lo + ci, oj = P.j_lo + cj;
    PowellMachine M;
    M.x[0] = M_PI; M.x[1] = 0.0;
    int k = P.nz_relax - 1;
    int ok = (int)((double)k + P.npivot);                    // int64 store truncates
    M.state = RS_START;
    double fe = 0.0;
    for (;;) {
        // cheap, divergent part: advance this lane's optimiser until it needs a function value.
        // Finishing a height stores its row and immediately starts the next one, so lanes do not
        // wait for the slowest lane at every height.
        while (!done && !powell_advance(M, fe, P.xtol, P.ftol, P.maxiter, P.maxfev)) {
            const size_t o = (size_t)col * P.nz_relax + k;
            P.out_etp[3 * o + 0] = M.fval;
            P.out_etp[3 * o + 1] = M.x[0];
            P.out_etp[3 * o + 2] = M.x[1];
            if (P.out_nfev) P.out_nfev[o] = M.nfev;
            if (--k < 0) { done = true; break; }
            ok = (int)((double)k + P.npivot);
            M.state = RS_START;
        }
        // expensive part: re-converge the warp, then all unfinished lanes evaluate the 64-point
        // stencil together (without the barrier the compiler threads the jump from each state
        // straight into a private copy of the evaluation and the lanes run it one group at a time)
        WARP_SYNC();
        if (WARP_ALL(done)) break;
        if (!done) fe = relax_objective<LAYOUT>(P, oi, oj, ok, M.ex[0], M.ex[1]);
    }
}

// sph2cart of every relaxed position (dbspm:603-605), one converged pass instead of divergent
// tail work inside the column loop
HD void relax_cart_body(const double *etp, double *cart, long long npos, double rpivot, long long gid, long long gstride)
{
    for (long long o = gid; o < npos; o += gstride) {
        const double th = etp[3 * o + 1], az = etp[3 * o + 2];
        cart[o] = rpivot * sin(th) * cos(az);
        cart[npos + o] = rpivot * sin(th) * sin(az);
        cart[2 * npos + o] = rpivot * cos(th);
    }
}

// (nx, ny, nz) -> (nx, nz, ny): 32x32 tiles through shared memory (tile[32][33]), one x-slab per blockIdx.z
HD void transpose_yz_tile(const double *in, double *out, int ny, int nz, double *tile /* 32*33 */,
                          int bx /* z tile */, int by /* y tile */, long long slab, int tx, int ty, int nty)
{
    const double *src = in + (size_t)slab * ny * nz;
    double *dst = out + (size_t)slab * ny * nz;
    for (int r = ty; r < 32; r += nty) {
        const int y = by * 32 + r, z = bx * 32 + tx;
        if (y < ny && z < nz) tile[r * 33 + tx] = src[(size_t)y * nz + z];
    }
    BLOCK_SYNC();
    for (int r = ty; r < 32; r += nty) {
        const int z = bx * 32 + r, y = by * 32 + tx;
        if (y < ny && z < nz) dst[(size_t)z * ny + y] = tile[tx * 33 + r];
    }
}
