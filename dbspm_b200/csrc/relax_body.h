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
    const double *grid;       // (nx,ny,nzc) static potential, z fastest
    int nx, ny, nzc;
    int i_lo, i_hi, j_lo, j_hi;
    int nz_relax;
    double kappa;
    double rpivot;            // params.RPIVOT (used by sph2cart, dbspm:605)
    double npivot;            // rpivot / dr[2]                     (dbspm:558)
    double rpivot_rt;         // npivot * dr[2]                     (interactions.py:73)
    double dr[3];
    int noortho;
    double Minv[9];           // inv(dR^T), row-major
    double xtol, ftol;
    int maxfev, maxiter;
    double *out_etp;          // (ni,nj,nz,3)
    double *out_cart;         // (3,ni,nj,nz) or null
    int *out_nfev;            // (ni,nj,nz) or null
};

// ---- periodic Catmull-Rom (a = -1/2) tricubic stencil in index units ---------------------
HD void cr_axis(double p, int n, int *i0, double *w)
{
    double d = fmod(p, (double)n);
    if (d < 0) d += (double)n;
    const double fl = floor(d);
    const double u = d - fl, u2 = u * u, u3 = u2 * u;
    *i0 = (int)fl;
    w[0] = -0.5 * u3 + u2 - 0.5 * u;
    w[1] = 1.5 * u3 - 2.5 * u2 + 1.0;
    w[2] = -1.5 * u3 + 2.0 * u2 + 0.5 * u;
    w[3] = 0.5 * u3 - 0.5 * u2;
}

HD int wrap_idx(int i, int n)
{
    i %= n;
    return i < 0 ? i + n : i;
}

HD double tricubic_eval(const double *g, int nx, int ny, int nz, double x, double y, double z)
{
    int ix, iy, iz;
    double wx[4], wy[4], wz[4];
    cr_axis(x, nx, &ix, wx);
    cr_axis(y, ny, &iy, wy);
    cr_axis(z, nz, &iz, wz);
    int zi[4], yo[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        zi[c] = wrap_idx(iz - 1 + c, nz);
        yo[c] = wrap_idx(iy - 1 + c, ny) * nz;
    }
    double r = 0.0;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const double *gx = g + (size_t)wrap_idx(ix - 1 + a, nx) * ny * nz;
        double ry = 0.0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const double *gy = gx + yo[b];
            double rz = wz[0] * LDG(gy + zi[0]);
            rz = fma(wz[1], LDG(gy + zi[1]), rz);
            rz = fma(wz[2], LDG(gy + zi[2]), rz);
            rz = fma(wz[3], LDG(gy + zi[3]), rz);
            ry = fma(wy[b], rz, ry);
        }
        r = fma(wx[a], ry, r);
    }
    return r;
}

#if defined(DBSPM_EMU_INTERP_HOOK) && !defined(__CUDA_ARCH__)
extern double (*dbspm_emu_interp_hook)(double, double, double);
#endif

// ---- objective: Emin / Emin_norm ----------------------------------------------------------
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
    return tricubic_eval(P.grid, P.nx, P.ny, P.nzc, (double)oi + ox, (double)oj + oy, (double)ok + oz) + spr;
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
            M.state = RS_ITER_BEGIN;
            break;
        case RS_ITER_BEGIN:
            M.fx = M.fval; M.bigind = 0; M.delta = 0.0;
            M.which = 0;
            M.state = RS_LS_BEGIN;
            break;
        case RS_LS_BEGIN:
            if (M.which == 0) { M.xi[0] = M.d0[0]; M.xi[1] = M.d0[1]; }
            else if (M.which == 1) { M.xi[0] = M.d1[0]; M.xi[1] = M.d1[1]; }
            /* which == 2: xi already holds the extrapolation direction */
            if (M.which < 2) M.fx2 = M.fval;
            if (M.xi[0] == 0.0 && M.xi[1] == 0.0) { M.state = RS_AFTER_LS; break; }
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
            M.state = RS_BR_LOOP;
            break;
        case RS_BR_LOOP: {
            if (!(M.fc < M.fb)) { M.state = RS_BR_DONE; break; }
            const double tmp1 = (M.xb - M.xa) * (M.fb - M.fc);
            const double tmp2 = (M.xb - M.xc) * (M.fb - M.fa);
            const double val = tmp2 - tmp1;
            double denom;
            if (fabs(val) < _verysmall_num) denom = 2.0 * _verysmall_num;
            else denom = 2.0 * val;
            M.w = M.xb - ((M.xb - M.xc) * tmp2 - (M.xb - M.xa) * tmp1) / denom;
            const double wlim = M.xb + grow_limit * (M.xc - M.xb);
            if (M.br_iter > 1000) { M.state = RS_BR_DONE; break; }
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
                M.state = RS_BR_DONE; break;
            } else if (M.fw > M.fb) {
                M.xc = M.w; M.fc = M.fw;
                M.state = RS_BR_DONE; break;
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
            M.state = RS_BR_LOOP;
            break;
        case RS_BR_GEN:
            M.fw = fe;
            M.xa = M.xb; M.xb = M.xc; M.xc = M.w;
            M.fa = M.fb; M.fb = M.fc; M.fc = M.fw;
            M.state = RS_BR_LOOP;
            break;
        case RS_BR_DONE: {
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
                M.state = RS_LS_FINISH;
                break;
            }
            M.bx = M.bw = M.bv = M.xb;
            M.bfw = M.bfv = M.bfx = M.fb;
            if (M.xa < M.xc) { M.a = M.xa; M.b = M.xc; } else { M.a = M.xc; M.b = M.xa; }
            M.deltax = 0.0; M.rat = 0.0;
            M.bt_iter = 0;
            M.state = RS_BT_LOOP;
            break;
        }
        case RS_BT_LOOP: {
            if (!(M.bt_iter < 500)) { M.alpha_min = M.bx; M.fret = M.bfx; M.state = RS_LS_FINISH; break; }
            const double x = M.bx;
            const double tol1 = tol * fabs(x) + _mintol;
            const double tol2 = 2.0 * tol1;
            const double xmid = 0.5 * (M.a + M.b);
            if (fabs(x - xmid) < (tol2 - 0.5 * (M.b - M.a))) {
                M.alpha_min = M.bx; M.fret = M.bfx; M.state = RS_LS_FINISH; break;
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
            M.state = RS_BT_LOOP;
            break;
        }
        case RS_LS_FINISH:
            M.xi[0] = M.alpha_min * M.xi[0]; M.xi[1] = M.alpha_min * M.xi[1];
            M.x[0] = M.p[0] + M.xi[0]; M.x[1] = M.p[1] + M.xi[1];
            M.fval = M.fret;
            M.state = RS_AFTER_LS;
            break;
        case RS_AFTER_LS:
            if (M.which < 2) {
                if ((M.fx2 - M.fval) > M.delta) { M.delta = M.fx2 - M.fval; M.bigind = M.which; }
                if (M.which == 0) { M.which = 1; M.state = RS_LS_BEGIN; }
                else M.state = RS_ITER_END;
            } else {
                if (M.xi[0] != 0.0 || M.xi[1] != 0.0) {
                    if (M.bigind == 0) { M.d0[0] = M.d1[0]; M.d0[1] = M.d1[1]; }
                    M.d1[0] = M.xi[0]; M.d1[1] = M.xi[1];
                }
                M.state = RS_ITER_BEGIN;
            }
            break;
        case RS_ITER_END: {
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
            if (again) { M.which = 2; M.state = RS_LS_BEGIN; }
            else M.state = RS_ITER_BEGIN;
            break;
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
HDN inline void relax_column_body(const RelaxParams &P, long long col)
{
    const int nj = P.j_hi - P.j_lo;
    const int ci = (int)(col / nj), cj = (int)(col - (long long)ci * nj);
    const int oi = P.i_lo + ci, oj = P.j_lo + cj;
    const long long ncol = (long long)(P.i_hi - P.i_lo) * nj;
    PowellMachine M;
    M.x[0] = M_PI; M.x[1] = 0.0;
    int k = P.nz_relax - 1;
    int ok = (int)((double)k + P.npivot);                    // int64 store truncates
    M.state = RS_START;
    double fe = 0.0;
    // One flat loop: a thread that finishes height k stores its row and immediately asks for
    // the first value of height k-1, so the lanes of a warp stay together at the (expensive)
    // objective evaluation instead of waiting for the slowest lane at every height.
    for (;;) {
        if (powell_advance(M, fe, P.xtol, P.ftol, P.maxiter, P.maxfev)) {
            fe = relax_objective(P, oi, oj, ok, M.ex[0], M.ex[1]);
            continue;
        }
        const size_t o = (size_t)col * P.nz_relax + k;
        P.out_etp[3 * o + 0] = M.fval;
        P.out_etp[3 * o + 1] = M.x[0];
        P.out_etp[3 * o + 2] = M.x[1];
        if (P.out_cart) {
            const double th = M.x[0], az = M.x[1];
            const size_t plane = (size_t)ncol * P.nz_relax;
            P.out_cart[o] = P.rpivot * sin(th) * cos(az);
            P.out_cart[plane + o] = P.rpivot * sin(th) * sin(az);
            P.out_cart[2 * plane + o] = P.rpivot * cos(th);
        }
        if (P.out_nfev) P.out_nfev[o] = M.nfev;
        if (--k < 0) break;
        ok = (int)((double)k + P.npivot);
        M.state = RS_START;
    }
}
