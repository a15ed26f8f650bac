/*
 * dbspm_oracle.c -- CPU restatement of the reference's sr/es/relax algorithm.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under dbspm_b200/ may import, link or
 * call this file.  It is used by tests/, __graft_entry__.smoke() and by
 * bench.py's cpu_baseline / --impl reference legs as the checker and the
 * reported CPU baseline; it is never the thing shipped or measured as product.
 *
 * What it restates (paths relative to /root/reference unless absolute):
 *   corr_direct()        pydbspm/interactions.py:9-22  get_density_overlap, written as the
 *                        direct periodic sum the FFT expression evaluates
 *   tricubic (tc_*)      third-party PyPI package `tricubic` (unpinned, setup.py:14; source
 *                        NOT in the container).  Restated from its published algorithm
 *                        (Lekien & Marsden, IJNME 63 (2005) 455: 64 coefficients per voxel
 *                        from values + central-difference derivatives at the 8 corners)
 *                        and anchored on the reference's call sites dbspm:562,
 *                        interactions.py:59,66, grid.py:125-138 (index-unit coordinates,
 *                        periodic on all three axes, nested-list constructor, .ip(list)).
 *   powell_minimize()    scipy 1.18.1 scipy/optimize/_optimize.py:3413-3643 (_minimize_powell),
 *                        :3214-3247 (_linesearch_powell, unbounded branch), :3117-3148
 *                        (_recover_from_bracket_error), :2497-2611 (Brent.optimize),
 *                        :2954-3111 (bracket), :549-572 (maxfun wrapper)
 *   relax_column()       pydbspm/interactions.py:70-86 relax_tip_z with Emin :55-60,
 *                        spring_energy :49-52, sph2grid grid.py:364-369; and the
 *                        non-orthogonal variant :88-105 with Emin_norm :62-67
 *
 *   "kernel twin" mode   (oracle_relax_tile_twin) the same relax loop and the same scipy restatement, with
 *                        the two ingredients in which the CUDA kernel differs from the reference at the
 *                        1e-14 level swapped for CPU statements of the kernel's own arithmetic: the
 *                        separable Catmull-Rom stencil in the kernel's summation order with fused
 *                        multiply-adds (tc_ip_twin) and the library-independent sin/cos (det_sincos).
 *                        A GPU run must equal this mode BIT FOR BIT (tests/test_gpu_parity.py); the
 *                        difference between the two modes is then exactly "Lekien-Marsden vs
 *                        Catmull-Rom form + libm vs fixed-sequence sin/cos", measured on the CPU
 *                        (tests/test_oracle.py), with no GPU effect mixed in.
 *
 * Parity status: the sr/es part is pinned against the reference's own
 * get_density_overlap run in the build container (oracle/make_golden.py ->
 * tests/golden/).  The Powell part is pinned against the installed scipy
 * (bit-identical on the golden cases).  The tricubic boundary is "parity
 * unpinned": the package is absent and the reference holds no fixture for it.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off: no FMA contraction so
 * the arithmetic matches numpy/scipy's separate multiply and add).
 */
#include <math.h>
#include <setjmp.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------------- */
/* 1. direct periodic cross-correlation (interactions.py:9-22)               */
/*    E[R] = dV * sum_r rho0[r] * rho1[(r - R) mod N]                         */
/*    (fftn(rho0)*fftn(rho1[::-1,::-1,::-1]) -> ifftn -> roll(+1) evaluates   */
/*     exactly this sum; checked in tests/test_oracle.py against numpy.)      */
/* ------------------------------------------------------------------------- */
void oracle_corr_direct(const double *rho0, const double *rho1, int nx, int ny, int nz,
                        double dV, int z_lo, int z_hi, double *E /* (nx,ny,z_hi-z_lo) */)
{
    int nzw = z_hi - z_lo;
    for (int Rx = 0; Rx < nx; ++Rx)
        for (int Ry = 0; Ry < ny; ++Ry)
            for (int Rz = z_lo; Rz < z_hi; ++Rz) {
                long double acc = 0.0L;
                for (int x = 0; x < nx; ++x) {
                    int bx = x - Rx; if (bx < 0) bx += nx;
                    for (int y = 0; y < ny; ++y) {
                        int by = y - Ry; if (by < 0) by += ny;
                        const double *a = rho0 + ((size_t)x * ny + y) * nz;
                        const double *b = rho1 + ((size_t)bx * ny + by) * nz;
                        for (int z = 0; z < nz; ++z) {
                            int bz = z - Rz; bz %= nz; if (bz < 0) bz += nz;
                            acc += (long double)a[z] * (long double)b[bz];
                        }
                    }
                }
                E[((size_t)Rx * ny + Ry) * nzw + (Rz - z_lo)] = (double)(acc * (long double)dV);
            }
}

/* ------------------------------------------------------------------------- */
/* 2. tricubic interpolator (Lekien-Marsden), periodic, index units          */
/* ------------------------------------------------------------------------- */
typedef struct {
    int n1, n2, n3;
    const double *data;     /* borrowed, C order (z fastest) */
    double C[64][64];       /* coefficient matrix = inverse of the Hermite constraint matrix */
    int init, i1, i2, i3;   /* cached voxel (the package re-uses coefficients per voxel) */
    double coefs[64];
} tc_t;

static double pw_at(int power, int deriv, int x)
{   /* d^deriv/dx^deriv of x^power at x in {0,1}, deriv in {0,1} */
    if (deriv == 0) return x == 0 ? (power == 0) : 1.0;
    return x == 0 ? (power == 1) : (double)power;
}

static int tc_build_matrix(double C[64][64])
{
    /* rows: constraint r = 8*t + p, t in {f,fx,fy,fz,fxy,fxz,fyz,fxyz}, corner p = x+2y+4z
       cols: coefficient c = i + 4j + 16k of x^i y^j z^k                                     */
    static const int DX[8] = {0, 1, 0, 0, 1, 1, 0, 1};
    static const int DY[8] = {0, 0, 1, 0, 1, 0, 1, 1};
    static const int DZ[8] = {0, 0, 0, 1, 0, 1, 1, 1};
    static double B[64][128];
    for (int t = 0; t < 8; ++t)
        for (int p = 0; p < 8; ++p) {
            int r = 8 * t + p, x = p & 1, y = (p >> 1) & 1, z = (p >> 2) & 1;
            for (int k = 0; k < 4; ++k)
                for (int j = 0; j < 4; ++j)
                    for (int i = 0; i < 4; ++i)
                        B[r][i + 4 * j + 16 * k] = pw_at(i, DX[t], x) * pw_at(j, DY[t], y) * pw_at(k, DZ[t], z);
            for (int c = 0; c < 64; ++c) B[r][64 + c] = (r == c);
        }
    for (int col = 0; col < 64; ++col) {           /* Gauss-Jordan, partial pivoting */
        int piv = col;
        for (int r = col + 1; r < 64; ++r)
            if (fabs(B[r][col]) > fabs(B[piv][col])) piv = r;
        if (fabs(B[piv][col]) < 1e-12) return -1;
        if (piv != col)
            for (int c = 0; c < 128; ++c) { double t = B[piv][c]; B[piv][c] = B[col][c]; B[col][c] = t; }
        double inv = 1.0 / B[col][col];
        for (int c = 0; c < 128; ++c) B[col][c] *= inv;
        for (int r = 0; r < 64; ++r)
            if (r != col && B[r][col] != 0.0) {
                double f = B[r][col];
                for (int c = 0; c < 128; ++c) B[r][c] -= f * B[col][c];
            }
    }
    double worst = 0.0;
    for (int r = 0; r < 64; ++r)
        for (int c = 0; c < 64; ++c) {
            double v = B[r][64 + c], rv = nearbyint(v);   /* the inverse is an integer matrix */
            if (fabs(v - rv) > worst) worst = fabs(v - rv);
            C[r][c] = rv;
        }
    return worst < 1e-8 ? 0 : -2;
}

tc_t *tc_create(const double *data, int n1, int n2, int n3)
{
    tc_t *t = (tc_t *)calloc(1, sizeof(tc_t));
    if (!t) return NULL;
    t->n1 = n1; t->n2 = n2; t->n3 = n3; t->data = data; t->init = 0;
    if (tc_build_matrix(t->C) != 0) { free(t); return NULL; }
    return t;
}

void tc_destroy(tc_t *t) { free(t); }

static inline int wrapi(int i, int n) { i %= n; return i < 0 ? i + n : i; }

static inline double tc_at(const tc_t *t, int i1, int i2, int i3)
{
    return t->data[(size_t)wrapi(i3, t->n3) + (size_t)t->n3 * ((size_t)wrapi(i2, t->n2) + (size_t)t->n2 * wrapi(i1, t->n1))];
}

double tc_ip(tc_t *t, double x, double y, double z)
{
    double dx = fmod(x, (double)t->n1), dy = fmod(y, (double)t->n2), dz = fmod(z, (double)t->n3);
    if (dx < 0) dx += t->n1;            /* periodicity is built in */
    if (dy < 0) dy += t->n2;
    if (dz < 0) dz += t->n3;
    int xi = (int)floor(dx), yi = (int)floor(dy), zi = (int)floor(dz);
    if (!t->init || xi != t->i1 || yi != t->i2 || zi != t->i3) {
        double b[64];
        for (int p = 0; p < 8; ++p) {
            int a = xi + (p & 1), c = yi + ((p >> 1) & 1), e = zi + ((p >> 2) & 1);
            b[p]      = tc_at(t, a, c, e);
            b[8 + p]  = 0.5 * (tc_at(t, a + 1, c, e) - tc_at(t, a - 1, c, e));
            b[16 + p] = 0.5 * (tc_at(t, a, c + 1, e) - tc_at(t, a, c - 1, e));
            b[24 + p] = 0.5 * (tc_at(t, a, c, e + 1) - tc_at(t, a, c, e - 1));
            b[32 + p] = 0.25 * (tc_at(t, a + 1, c + 1, e) - tc_at(t, a - 1, c + 1, e)
                              - tc_at(t, a + 1, c - 1, e) + tc_at(t, a - 1, c - 1, e));
            b[40 + p] = 0.25 * (tc_at(t, a + 1, c, e + 1) - tc_at(t, a - 1, c, e + 1)
                              - tc_at(t, a + 1, c, e - 1) + tc_at(t, a - 1, c, e - 1));
            b[48 + p] = 0.25 * (tc_at(t, a, c + 1, e + 1) - tc_at(t, a, c - 1, e + 1)
                              - tc_at(t, a, c + 1, e - 1) + tc_at(t, a, c - 1, e - 1));
            b[56 + p] = 0.125 * (tc_at(t, a + 1, c + 1, e + 1) - tc_at(t, a - 1, c + 1, e + 1)
                               - tc_at(t, a + 1, c - 1, e + 1) + tc_at(t, a - 1, c - 1, e + 1)
                               - tc_at(t, a + 1, c + 1, e - 1) + tc_at(t, a - 1, c + 1, e - 1)
                               + tc_at(t, a + 1, c - 1, e - 1) - tc_at(t, a - 1, c - 1, e - 1));
        }
        for (int r = 0; r < 64; ++r) {
            double s = 0.0;
            for (int c = 0; c < 64; ++c) s += t->C[r][c] * b[c];
            t->coefs[r] = s;
        }
        t->init = 1; t->i1 = xi; t->i2 = yi; t->i3 = zi;
    }
    dx -= xi; dy -= yi; dz -= zi;
    int ijkn = 0;
    double dzpow = 1.0, result = 0.0;
    for (int k = 0; k < 4; ++k) {
        double dypow = 1.0;
        for (int j = 0; j < 4; ++j) {
            const double *c = t->coefs + ijkn;
            result += dypow * dzpow * (c[0] + dx * (c[1] + dx * (c[2] + dx * c[3])));
            ijkn += 4;
            dypow *= dy;
        }
        dzpow *= dz;
    }
    return result;
}

/* separable Catmull-Rom (a = -1/2) form of the same interpolant; the CUDA
   kernel uses this form.  Kept here so the equivalence is tested on the CPU. */
double tc_ip_catmull(const tc_t *t, double x, double y, double z)
{
    double p[3] = {x, y, z}, w[3][4];
    int n[3] = {t->n1, t->n2, t->n3}, i0[3];
    for (int a = 0; a < 3; ++a) {
        double d = fmod(p[a], (double)n[a]);
        if (d < 0) d += n[a];
        int fi = (int)floor(d);
        double u = d - fi, u2 = u * u, u3 = u2 * u;
        i0[a] = fi;
        w[a][0] = -0.5 * u3 + u2 - 0.5 * u;
        w[a][1] = 1.5 * u3 - 2.5 * u2 + 1.0;
        w[a][2] = -1.5 * u3 + 2.0 * u2 + 0.5 * u;
        w[a][3] = 0.5 * u3 - 0.5 * u2;
    }
    double r = 0.0;
    for (int a = 0; a < 4; ++a) {
        double ry = 0.0;
        for (int b = 0; b < 4; ++b) {
            double rz = 0.0;
            for (int c = 0; c < 4; ++c) rz += w[2][c] * tc_at(t, i0[0] - 1 + a, i0[1] - 1 + b, i0[2] - 1 + c);
            ry += w[1][b] * rz;
        }
        r += w[0][a] * ry;
    }
    return r;
}

/* "kernel twin" of the interpolant: the CUDA kernel's arithmetic (dbspm_b200/csrc/relax_body.h tricubic_eval) stated
   independently -- same weights, y innermost / z middle / x outermost, the first product of every row a plain multiply
   and every accumulation a fused multiply-add */
double tc_ip_twin(const tc_t *t, double x, double y, double z)
{
    double p[3] = {x, y, z}, w[3][4];
    int n[3] = {t->n1, t->n2, t->n3}, i0[3];
    for (int a = 0; a < 3; ++a) {
        double d = fmod(p[a], (double)n[a]);
        if (d < 0) d += n[a];
        double fl = floor(d);
        double u = d - fl, u2 = u * u, u3 = u2 * u;
        i0[a] = (int)fl;
        w[a][0] = -0.5 * u3 + u2 - 0.5 * u;
        w[a][1] = 1.5 * u3 - 2.5 * u2 + 1.0;
        w[a][2] = -1.5 * u3 + 2.0 * u2 + 0.5 * u;
        w[a][3] = 0.5 * u3 - 0.5 * u2;
    }
    double r = 0.0;
    for (int a = 0; a < 4; ++a) {
        double rx = 0.0;
        for (int c = 0; c < 4; ++c) {
            double ry = w[1][0] * tc_at(t, i0[0] - 1 + a, i0[1] - 1, i0[2] - 1 + c);
            for (int b = 1; b < 4; ++b) ry = fma(w[1][b], tc_at(t, i0[0] - 1 + a, i0[1] - 1 + b, i0[2] - 1 + c), ry);
            rx = fma(w[2][c], ry, rx);
        }
        r = fma(w[0][a], rx, r);
    }
    return r;
}

/* library-independent sin and cos (the kernel's dbspm_sincos stated independently): rint(x * 2/pi) by the
   1.5 * 2^52 shift, three-piece Cody-Waite reduction with fused multiply-adds, fdlibm's __kernel_sin / __kernel_cos
   polynomials in Horner form with fused multiply-adds, quadrant from the two low bits of the rounded multiple */
static void det_sincos(double x, double *sn, double *cs)
{
    if (!(fabs(x) <= 1.0e5)) { *sn = sin(x); *cs = cos(x); return; }
    const double shift = 6755399441055744.0;
    double t = fma(x, 0.63661977236758138, shift);
    double k = t - shift;
    long long bits; memcpy(&bits, &t, sizeof bits);
    int q = (int)(bits & 3);
    double r = fma(-k, 1.5707963267948966e+00, x);
    r = fma(-k, 6.1232339957367574e-17, r);
    r = fma(-k, 8.4784276603688985e-32, r);
    double z = r * r;
    static const double S[6] = {-1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,
                                2.75573137070700676789e-06, -2.50507602534068634195e-08, 1.58969099521155010221e-10};
    static const double Cc[6] = {4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05,
                                 -2.75573143513906633035e-07, 2.08757232129817482790e-09, -1.13596475577881948265e-11};
    double ps = S[5], pc = Cc[5];
    for (int i = 4; i >= 0; --i) { ps = fma(ps, z, S[i]); pc = fma(pc, z, Cc[i]); }
    double s0 = fma(r * z, ps, r);
    double hz = 0.5 * z, w = 1.0 - hz;
    double c0 = w + fma(z * z, pc, (1.0 - w) - hz);
    double sa = (q & 1) ? c0 : s0, ca = (q & 1) ? s0 : c0;
    *sn = (q & 2) ? -sa : sa;
    *cs = ((q + 1) & 2) ? -ca : ca;
}

void oracle_det_sincos(const double *x, double *s, double *c, long n)
{
    for (long i = 0; i < n; ++i) det_sincos(x[i], s + i, c + i);
}

/* ------------------------------------------------------------------------- */
/* 3. scipy Powell (N = 2), Brent line search, bracket                       */
/* ------------------------------------------------------------------------- */
typedef double (*obj2_t)(const double x[2], void *ctx);

typedef struct {
    obj2_t f; void *ctx;
    int ncalls, maxfun;
    jmp_buf too_many;                  /* _MaxFuncCallError (_optimize.py:557-559) */
} fwrap_t;

static double wrapped(fwrap_t *w, const double x[2])
{
    if (w->ncalls >= w->maxfun) longjmp(w->too_many, 1);
    w->ncalls += 1;
    return w->f(x, w->ctx);
}

typedef struct { fwrap_t *w; double p[2], xi[2]; } line_t;

static double myfunc(line_t *L, double alpha)        /* _optimize.py:3236-3237 */
{
    double x[2] = {L->p[0] + alpha * L->xi[0], L->p[1] + alpha * L->xi[1]};
    return wrapped(L->w, x);
}

/* bracket(), _optimize.py:2954-3111; returns 0 for a valid bracket, 1 for BracketError */
static int bracket(line_t *L, double *pxa, double *pxb, double *pxc,
                   double *pfa, double *pfb, double *pfc, int *pfuncalls)
{
    const double _gold = 1.618034, _verysmall_num = 1e-21, grow_limit = 110.0;
    const int maxiter = 1000;
    double xa = 0.0, xb = 1.0, xc, fa, fb, fc, w, fw, tmp1, tmp2, val, denom, wlim;
    int funcalls, iter = 0;
    fa = myfunc(L, xa);
    fb = myfunc(L, xb);
    if (fa < fb) { double t = xa; xa = xb; xb = t; t = fa; fa = fb; fb = t; }
    xc = xb + _gold * (xb - xa);
    fc = myfunc(L, xc);
    funcalls = 3;
    while (fc < fb) {
        tmp1 = (xb - xa) * (fb - fc);
        tmp2 = (xb - xc) * (fb - fa);
        val = tmp2 - tmp1;
        if (fabs(val) < _verysmall_num) denom = 2.0 * _verysmall_num;
        else denom = 2.0 * val;
        w = xb - ((xb - xc) * tmp2 - (xb - xa) * tmp1) / denom;
        wlim = xb + grow_limit * (xc - xb);
        if (iter > maxiter) break;      /* RuntimeError in scipy; unreachable for finite f */
        iter += 1;
        if ((w - xc) * (xb - w) > 0.0) {
            fw = myfunc(L, w); funcalls += 1;
            if (fw < fc) { xa = xb; xb = w; fa = fb; fb = fw; break; }
            else if (fw > fb) { xc = w; fc = fw; break; }
            w = xc + _gold * (xc - xb);
            fw = myfunc(L, w); funcalls += 1;
        } else if ((w - wlim) * (wlim - xc) >= 0.0) {
            w = wlim;
            fw = myfunc(L, w); funcalls += 1;
        } else if ((w - wlim) * (xc - w) > 0.0) {
            fw = myfunc(L, w); funcalls += 1;
            if (fw < fc) {
                xb = xc; xc = w; w = xc + _gold * (xc - xb);
                fb = fc; fc = fw;
                fw = myfunc(L, w); funcalls += 1;
            }
        } else {
            w = xc + _gold * (xc - xb);
            fw = myfunc(L, w); funcalls += 1;
        }
        xa = xb; xb = xc; xc = w;
        fa = fb; fb = fc; fc = fw;
    }
    int cond1 = (fb < fc && fb <= fa) || (fb < fa && fb <= fc);
    int cond2 = (xa < xb && xb < xc) || (xc < xb && xb < xa);
    int cond3 = isfinite(xa) && isfinite(xb) && isfinite(xc);
    *pxa = xa; *pxb = xb; *pxc = xc; *pfa = fa; *pfb = fb; *pfc = fc; *pfuncalls = funcalls;
    return (cond1 && cond2 && cond3) ? 0 : 1;
}

/* _recover_from_bracket_error(_minimize_scalar_brent, myfunc, None, (), xtol=tol) */
static void brent_linemin(line_t *L, double tol, double *xmin, double *fmin)
{
    const double _mintol = 1.0e-11, _cg = 0.3819660;
    const int maxiter = 500;
    double xa, xb, xc, fa, fb, fc;
    int funcalls;
    if (bracket(L, &xa, &xb, &xc, &fa, &fb, &fc, &funcalls)) {
        /* _optimize.py:3137-3147: best of the three points (np.argmin: first minimum) */
        double xs[3] = {xa, xb, xc}, fs[3] = {fa, fb, fc};
        if (isnan(xa) || isnan(xb) || isnan(xc) || isnan(fa) || isnan(fb) || isnan(fc)) {
            *xmin = NAN; *fmin = NAN; return;
        }
        int im = 0;
        for (int q = 1; q < 3; ++q) if (fs[q] < fs[im]) im = q;
        *xmin = xs[im]; *fmin = fs[im];
        return;
    }
    /* Brent.optimize, _optimize.py:2497-2611 */
    double x, w, v, fx, fw, fv, a, b, deltax = 0.0, rat = 0.0, u, fu;
    double tol1, tol2, xmid, tmp1, tmp2, p, dx_temp;
    int iter = 0;
    x = w = v = xb;
    fw = fv = fx = fb;
    if (xa < xc) { a = xa; b = xc; } else { a = xc; b = xa; }
    while (iter < maxiter) {
        tol1 = tol * fabs(x) + _mintol;
        tol2 = 2.0 * tol1;
        xmid = 0.5 * (a + b);
        if (fabs(x - xmid) < (tol2 - 0.5 * (b - a))) break;
        if (fabs(deltax) <= tol1) {
            if (x >= xmid) deltax = a - x; else deltax = b - x;
            rat = _cg * deltax;
        } else {
            tmp1 = (x - w) * (fx - fv);
            tmp2 = (x - v) * (fx - fw);
            p = (x - v) * tmp2 - (x - w) * tmp1;
            tmp2 = 2.0 * (tmp2 - tmp1);
            if (tmp2 > 0.0) p = -p;
            tmp2 = fabs(tmp2);
            dx_temp = deltax;
            deltax = rat;
            if ((p > tmp2 * (a - x)) && (p < tmp2 * (b - x)) && (fabs(p) < fabs(0.5 * tmp2 * dx_temp))) {
                rat = p * 1.0 / tmp2;
                u = x + rat;
                if ((u - a) < tol2 || (b - u) < tol2) {
                    if (xmid - x >= 0) rat = tol1; else rat = -tol1;
                }
            } else {
                if (x >= xmid) deltax = a - x; else deltax = b - x;
                rat = _cg * deltax;
            }
        }
        if (fabs(rat) < tol1) { if (rat >= 0) u = x + tol1; else u = x - tol1; }
        else u = x + rat;
        fu = myfunc(L, u);
        funcalls += 1;
        if (fu > fx) {
            if (u < x) a = u; else b = u;
            if ((fu <= fw) || (w == x)) { v = w; w = u; fv = fw; fw = fu; }
            else if ((fu <= fv) || (v == x) || (v == w)) { v = u; fv = fu; }
        } else {
            if (u >= x) a = x; else b = x;
            v = w; w = x; x = u;
            fv = fw; fw = fx; fx = fu;
        }
        iter += 1;
    }
    *xmin = x; *fmin = fx;
}

/* _linesearch_powell, unbounded branch (_optimize.py:3239-3247) */
static void linesearch_powell(fwrap_t *W, double p[2], double xi[2], double tol, double *fval)
{
    if (xi[0] == 0.0 && xi[1] == 0.0) return;     /* "if not np.any(xi)": fval given -> unchanged */
    line_t L; L.w = W; L.p[0] = p[0]; L.p[1] = p[1]; L.xi[0] = xi[0]; L.xi[1] = xi[1];
    double alpha_min, fret;
    brent_linemin(&L, tol, &alpha_min, &fret);
    xi[0] = alpha_min * xi[0]; xi[1] = alpha_min * xi[1];
    p[0] = p[0] + xi[0]; p[1] = p[1] + xi[1];
    *fval = fret;
}

/* _minimize_powell with N = 2, bounds=None, direc=None (_optimize.py:3413-3643).
   Returns nfev; writes x (2), fun, nit. */
int oracle_powell_minimize(obj2_t f, void *ctx, const double x0[2], double xtol, double ftol,
                           int maxiter, int maxfun, double xout[2], double *fun, int *nit)
{
    enum { N = 2 };
    fwrap_t W; W.f = f; W.ctx = ctx; W.ncalls = 0; W.maxfun = maxfun;
    /* volatile: values must survive the longjmp that models _MaxFuncCallError */
    volatile double vx[2], vfval;
    volatile int viter = 0;
    double x[2] = {x0[0], x0[1]}, x1[2], x2[2], direc[N][N] = {{1.0, 0.0}, {0.0, 1.0}}, direc1[2];
    double fval, fx, fx2, delta, bnd, t, temp;
    int bigind;
    vx[0] = x[0]; vx[1] = x[1]; vfval = NAN;
    if (setjmp(W.too_many) == 0) {
        fval = wrapped(&W, x);                       /* outside scipy's try: maxfun >= 1 assumed */
        vfval = fval;
        x1[0] = x[0]; x1[1] = x[1];
        for (;;) {
            fx = fval; bigind = 0; delta = 0.0;
            for (int i = 0; i < N; ++i) {
                direc1[0] = direc[i][0]; direc1[1] = direc[i][1];
                fx2 = fval;
                linesearch_powell(&W, x, direc1, xtol * 100, &fval);
                vx[0] = x[0]; vx[1] = x[1]; vfval = fval;
                if ((fx2 - fval) > delta) { delta = fx2 - fval; bigind = i; }
            }
            viter = viter + 1;
            bnd = ftol * (fabs(fx) + fabs(fval)) + 1e-20;
            if (2.0 * (fx - fval) <= bnd) break;
            if (W.ncalls >= maxfun) break;
            if (viter >= maxiter) break;
            if (isnan(fx) && isnan(fval)) break;
            direc1[0] = x[0] - x1[0]; direc1[1] = x[1] - x1[1];
            x1[0] = x[0]; x1[1] = x[1];
            x2[0] = x[0] + 1 * direc1[0]; x2[1] = x[1] + 1 * direc1[1];
            fx2 = wrapped(&W, x2);
            if (fx > fx2) {
                t = 2.0 * (fx + fx2 - 2.0 * fval);
                temp = (fx - fval - delta);
                t *= temp * temp;
                temp = fx - fx2;
                t -= delta * temp * temp;
                if (t < 0.0) {
                    linesearch_powell(&W, x, direc1, xtol * 100, &fval);
                    vx[0] = x[0]; vx[1] = x[1]; vfval = fval;
                    if (direc1[0] != 0.0 || direc1[1] != 0.0) {
                        direc[bigind][0] = direc[N - 1][0]; direc[bigind][1] = direc[N - 1][1];
                        direc[N - 1][0] = direc1[0]; direc[N - 1][1] = direc1[1];
                    }
                }
            }
        }
    }
    xout[0] = vx[0]; xout[1] = vx[1]; *fun = vfval; *nit = viter;
    return W.ncalls;
}

/* ------------------------------------------------------------------------- */
/* 4. relax_tip_z / relax_noortho_tip_z (interactions.py:70-105)             */
/* ------------------------------------------------------------------------- */
typedef struct {
    tc_t *tc;
    long origin[3];            /* np.array([i, j, 0]) is int64: origin[2] = k + npivot truncates */
    double kappa, rpivot, dr[3];
    int noortho;
    double Minv[9];            /* inv(dR.T) for Emin_norm */
    int twin;                  /* 1: "kernel twin" mode (tc_ip_twin + det_sincos), see the file header */
} emin_ctx_t;

static double emin_obj(const double pol[2], void *vctx)
{
    emin_ctx_t *c = (emin_ctx_t *)vctx;
    double th = pol[0], az = pol[1];
    double d = M_PI - th;
    double spr = 0.5 * c->kappa * (d * d);                       /* interactions.py:49-52 */
    double ox, oy, oz, sth, cth, saz, caz;
    if (c->twin) { det_sincos(th, &sth, &cth); det_sincos(az, &saz, &caz); }
    else { sth = sin(th); cth = cos(th); saz = sin(az); caz = cos(az); }
    if (!c->noortho) {                                           /* grid.py:364-369 */
        ox = c->rpivot * sth * caz / c->dr[0];
        oy = c->rpivot * sth * saz / c->dr[1];
        oz = c->rpivot * cth / c->dr[2];
    } else {                                                     /* grid.py:381-386 + :66 */
        double X = c->rpivot * sth * caz, Y = c->rpivot * sth * saz, Z = c->rpivot * cth;
        ox = c->Minv[0] * X + c->Minv[1] * Y + c->Minv[2] * Z;
        oy = c->Minv[3] * X + c->Minv[4] * Y + c->Minv[5] * Z;
        oz = c->Minv[6] * X + c->Minv[7] * Y + c->Minv[8] * Z;
    }
    double px = (double)c->origin[0] + ox, py = (double)c->origin[1] + oy, pz = (double)c->origin[2] + oz;
    return (c->twin ? tc_ip_twin(c->tc, px, py, pz) : tc_ip(c->tc, px, py, pz)) + spr;
}

static void inv3(const double m[9], double out[9])
{
    double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
    double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
    out[0] = (e * i - f * h) / det; out[1] = (c * h - b * i) / det; out[2] = (b * f - c * e) / det;
    out[3] = (f * g - d * i) / det; out[4] = (a * i - c * g) / det; out[5] = (c * d - a * f) / det;
    out[6] = (d * h - e * g) / det; out[7] = (b * g - a * h) / det; out[8] = (a * e - b * d) / det;
}

/* One lateral column.  zi: nz z-indices (ascending); sweep is zi[::-1] with warm start.
   out: (nz,3) rows [fun, theta, phi] in ascending-z order; nfev: (nz) or NULL.
   dR: 9 doubles (row-major 3x3) or NULL for the orthogonal path. */
static void relax_column_mode(tc_t *tc, long i, long j, double kappa, double npivot,
                              const long *zi, int nz, const double dr[3], const double *dR,
                              double xtol, double ftol, int maxfev, double *out, int *nfev, int twin)
{
    emin_ctx_t c;
    c.twin = twin;
    c.tc = tc; c.origin[0] = i; c.origin[1] = j; c.origin[2] = 0;
    c.kappa = kappa; c.rpivot = npivot * dr[2];                  /* interactions.py:73 */
    c.dr[0] = dr[0]; c.dr[1] = dr[1]; c.dr[2] = dr[2];
    c.noortho = dR != NULL;
    if (dR) {
        double dRT[9] = {dR[0], dR[3], dR[6], dR[1], dR[4], dR[7], dR[2], dR[5], dR[8]};
        inv3(dRT, c.Minv);
    }
    double guess[2] = {M_PI, 0.0};
    for (int q = nz - 1; q >= 0; --q) {
        c.origin[2] = (long)((double)zi[q] + npivot);            /* int64 store truncates */
        double x[2], fun; int nit;
        int n = oracle_powell_minimize(emin_obj, &c, guess, xtol, ftol, 2000, maxfev, x, &fun, &nit);
        guess[0] = x[0]; guess[1] = x[1];
        out[3 * q + 0] = fun; out[3 * q + 1] = x[0]; out[3 * q + 2] = x[1];
        if (nfev) nfev[q] = n;
    }
}

void oracle_relax_column(tc_t *tc, long i, long j, double kappa, double npivot,
                         const long *zi, int nz, const double dr[3], const double *dR,
                         double xtol, double ftol, int maxfev, double *out, int *nfev)
{
    relax_column_mode(tc, i, j, kappa, npivot, zi, nz, dr, dR, xtol, ftol, maxfev, out, nfev, 0);
}

/* whole lateral tile, plain loop (threads are added by the caller with processes) */
static void relax_tile_mode(const double *static_win, int nx, int ny, int nzc,
                            int i_lo, int i_hi, int j_lo, int j_hi, int nz_relax,
                            double kappa, double rpivot_param, const double dr[3], const double *dR,
                            double xtol, double ftol, int maxfev,
                            double *out /* (ni,nj,nz,3) */, int *nfev /* (ni,nj,nz) or NULL */, int twin)
{
    tc_t *tc = tc_create(static_win, nx, ny, nzc);
    long *zi = (long *)malloc(sizeof(long) * (size_t)nz_relax);
    for (int k = 0; k < nz_relax; ++k) zi[k] = k;
    double npivot = rpivot_param / dr[2];                        /* dbspm:558 */
    int nj = j_hi - j_lo;
    for (int i = i_lo; i < i_hi; ++i)
        for (int j = j_lo; j < j_hi; ++j) {
            size_t col = (size_t)(i - i_lo) * nj + (j - j_lo);
            relax_column_mode(tc, i, j, kappa, npivot, zi, nz_relax, dr, dR, xtol, ftol, maxfev,
                              out + col * nz_relax * 3, nfev ? nfev + col * nz_relax : NULL, twin);
        }
    free(zi);
    tc_destroy(tc);
}

void oracle_relax_tile(const double *static_win, int nx, int ny, int nzc,
                       int i_lo, int i_hi, int j_lo, int j_hi, int nz_relax,
                       double kappa, double rpivot_param, const double dr[3], const double *dR,
                       double xtol, double ftol, int maxfev, double *out, int *nfev)
{
    relax_tile_mode(static_win, nx, ny, nzc, i_lo, i_hi, j_lo, j_hi, nz_relax, kappa, rpivot_param, dr, dR,
                    xtol, ftol, maxfev, out, nfev, 0);
}

/* "kernel twin" mode: what the CUDA kernel must reproduce bit for bit */
void oracle_relax_tile_twin(const double *static_win, int nx, int ny, int nzc,
                            int i_lo, int i_hi, int j_lo, int j_hi, int nz_relax,
                            double kappa, double rpivot_param, const double dr[3], const double *dR,
                            double xtol, double ftol, int maxfev, double *out, int *nfev)
{
    relax_tile_mode(static_win, nx, ny, nzc, i_lo, i_hi, j_lo, j_hi, nz_relax, kappa, rpivot_param, dr, dR,
                    xtol, ftol, maxfev, out, nfev, 1);
}

/* sph2cart of relaxed angles in twin mode (the kernel's relax_cart_body): (3, n) x, y, z */
void oracle_sph2cart_twin(const double *th, const double *az, double r, long n, double *out)
{
    for (long i = 0; i < n; ++i) {
        double sth, cth, saz, caz;
        det_sincos(th[i], &sth, &cth);
        det_sincos(az[i], &saz, &caz);
        out[i] = r * sth * caz; out[n + i] = r * sth * saz; out[2 * n + i] = r * cth;
    }
}

/* generic objective hook so tests can drive oracle_powell_minimize from Python */
typedef double (*py_obj_t)(double, double);
static double py_tramp(const double x[2], void *ctx) { return ((py_obj_t)ctx)(x[0], x[1]); }
int oracle_powell_minimize_cb(py_obj_t f, const double x0[2], double xtol, double ftol,
                              int maxiter, int maxfun, double xout[2], double *fun, int *nit)
{
    return oracle_powell_minimize(py_tramp, (void *)f, x0, xtol, ftol, maxiter, maxfun, xout, fun, nit);
}
