// relax.cu -- sm_100a kernels + C ABI for the relax step, the static-potential assembly and
// the two "next" rows (tricubic gather, dE/dz).  Compiled with --fmad=false: the Powell
// bookkeeping must use separate multiply and add like numpy/scipy; the interpolation stencil
// asks for fused multiply-adds explicitly (fma()).
// Replaces dbspm:479-486, dbspm:558-605, pydbspm/interactions.py:49-105, grid.py:325-386.
#include <stdlib.h>
#include "capi_common.h"
#include "relax_body.h"

// One warp per block: the kernel has no block barrier, so a slot of the SM frees when its warp finishes instead of when
// the slowest of four does (B200, 1024x1024x24, y-quad layout: 4 x 128 threads 72.5 ms, 16 x 32 threads 70.3 ms).
// 16 blocks = 128 registers, no spills; 96 registers (21 warps) and 80 (25 warps) measured slower (75.6 / 89.2 ms) even with
// the bracket / Brent state moved to shared memory (RELAX_SMEM_STATE): fewer stencil loads in flight per warp cost more
// than the extra warps hide.
#ifndef RELAX_THREADS
#define RELAX_THREADS 32
#endif
#ifndef RELAX_MIN_BLOCKS
#define RELAX_MIN_BLOCKS 16
#endif

__global__ void __launch_bounds__(256) relax_cart_kernel(const double *etp, double *cart, long long npos, double rpivot)
{
    relax_cart_body(etp, cart, npos, rpivot, (long long)blockIdx.x * blockDim.x + threadIdx.x,
                    (long long)gridDim.x * blockDim.x);
}

#ifndef RELAX_FOOT_X_DEFAULT
#define RELAX_FOOT_X_DEFAULT 8
#endif

template <int LAYOUT>
__global__ void __launch_bounds__(RELAX_THREADS, RELAX_MIN_BLOCKS) relax_kernel(const __grid_constant__ RelaxParams P, long long ncol)
{
    const long long col = relax_patch_column(P, (long long)blockIdx.x * blockDim.x + threadIdx.x, ncol);
    __shared__ double powell_sm[(RELAX_SMEM_STATE == 0 ? PS_COUNT : (RELAX_SMEM_STATE == 1 ? PS_P_0 : PS_TOTAL)) * RELAX_THREADS];       // Powell-level state, one column per thread
    relax_column_body<LAYOUT, RELAX_THREADS>(P, col, ncol, powell_sm, threadIdx.x, P.k_hi, P.k_lo);      // lanes past the end idle inside (warp barrier)
}

// Dynamic schedule: a persistent grid (one wave: SMs x resident warps) whose warps pull tasks from a global counter.  A task =
// (patch of 32 columns, chunk of `hpc` heights), issued chunk-major: all patches' top chunk first.  A static launch gives every
// warp one patch with all its heights -- with 4096 patches for 2368 resident warps (one rank of 8 on the 1024 x 1024 grid) the
// second wave runs at 73 % occupancy and the slowest patches (over the molecule) finish alone; tasks of a quarter of the length
// fill the tail.  The chunk c of a patch needs chunk c - 1 (warm start): `progress[patch]` counts finished chunks; the task
// that needs it was issued a whole round of patches later, so the wait is almost never taken, and it cannot deadlock: a task
// only waits for tasks issued BEFORE it, which run on resident warps.  Each height starts a fresh Powell minimisation from the
// previous height's angles (as the reference), so chunking does not change a single bit.
template <int LAYOUT>
__global__ void __launch_bounds__(RELAX_THREADS, RELAX_MIN_BLOCKS) relax_dyn_kernel(const __grid_constant__ RelaxParams P, long long ncol,
                                                                                    int npatch, int nchunks, int hpc, int *ctrl)
{
    __shared__ double powell_sm[(RELAX_SMEM_STATE == 0 ? PS_COUNT : (RELAX_SMEM_STATE == 1 ? PS_P_0 : PS_TOTAL)) * RELAX_THREADS];
    const int lane = threadIdx.x & 31;
    int *progress = ctrl + 1;
    for (;;) {
        int t = 0;
        if (lane == 0) t = atomicAdd(ctrl, 1);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= npatch * nchunks) break;
        const int c = t / npatch, p = t - c * npatch;
        if (c > 0) {
            if (lane == 0) {
                int seen;
                do {
                    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(progress + p) : "memory");
                    if (seen < c) __nanosleep(256);
                } while (seen < c);
            }
            __syncwarp();
        }
        const int k_hi = P.nz_relax - 1 - c * hpc, k_lo = (k_hi - hpc + 1) > 0 ? (k_hi - hpc + 1) : 0;
        const long long col = relax_patch_column(P, (long long)p * 32 + lane, ncol);
        relax_column_body<LAYOUT, RELAX_THREADS>(P, col, ncol, powell_sm, threadIdx.x, k_hi, k_lo);
        __threadfence();                      // every lane's rows of out_etp before the flag
        __syncwarp();
        if (lane == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(progress + p), "r"(c + 1) : "memory");
    }
}

// (nx,ny,nz) -> (nx,nz,ny+3) so that the lanes of a warp (adjacent y columns) read adjacent addresses
// and the y neighbours of a stencil row are contiguous across the periodic boundary
__global__ void __launch_bounds__(256) transpose_yz_kernel(const double *in, double *out, int ny, int nz)
{
    __shared__ double tile[32 * 33];
    transpose_yz_tile(in, out, ny, nz, tile, (int)blockIdx.x, (int)blockIdx.y, (long long)blockIdx.z,
                      (int)threadIdx.x, (int)threadIdx.y, (int)blockDim.y);
}

// (nx,ny,nz) -> (nx,nz,ny,4): the four y neighbours of every stencil row as one aligned 32-byte sector (LAYOUT 2)
__global__ void __launch_bounds__(256) quad_yz_kernel(const double *in, double *out, int ny, int nz)
{
    __shared__ double tile[35 * 33];
    quad_yz_tile(in, out, ny, nz, tile, (int)blockIdx.x, (int)blockIdx.y, (long long)blockIdx.z,
                 (int)threadIdx.x, (int)threadIdx.y, (int)blockDim.y);
}

__global__ void __launch_bounds__(256) accumulate_kernel(double *acc, const double *comp, double scale, size_t n, int init)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double term = comp[i] * scale;             // numpy: temp = comp * scale
        acc[i] = init ? (0.0 + term) : (acc[i] + term);  //        static += temp
    }
}

__global__ void __launch_bounds__(256) gather_kernel(const double *data, int nx, int ny, int nz,
                                                     const double *pts, size_t npts, double *out)
{
    const double inv_n[3] = {1.0 / nx, 1.0 / ny, 1.0 / nz}, nd[3] = {(double)nx, (double)ny, (double)nz};
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npts; i += stride)
        out[i] = tricubic_eval<0>(data, nx, ny, nz, 0, nd, inv_n, pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
}

// np.gradient(E, -dz, axis=2): central differences inside, one-sided first order at the ends
__global__ void __launch_bounds__(256) gradz_kernel(const double *E, long long ncol, int nz, double h, double *out)
{
    const long long total = ncol * nz;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int k = (int)(i % nz);
        double g;
        if (nz == 1) g = 0.0;
        else if (k == 0) g = (E[i + 1] - E[i]) / h;
        else if (k == nz - 1) g = (E[i] - E[i - 1]) / h;
        else g = (E[i + 1] - E[i - 1]) / (2.0 * h);
        out[i] = g;
    }
}

// out[c, i] = scale * sum_j w[j] * F[c, i + nw - 1 - j]: np.convolve(F[c,:], w, mode="valid") along z
// (pydbspm/tools.py:30, the Giessibl weights of get_fs); nw <= 64 weights live in shared memory
__global__ void __launch_bounds__(256) convz_valid_kernel(const double *F, long long ncol, int nz, const double *w, int nw,
                                                          double scale, double *out)
{
    __shared__ double sw[64];
    if (threadIdx.x < nw) sw[threadIdx.x] = w[threadIdx.x];
    __syncthreads();
    const int nout = nz - nw + 1;
    const long long total = ncol * nout;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long c = i / nout;
        const int z = (int)(i - c * nout);
        const double *f = F + c * nz + z + nw - 1;
        double acc = 0.0;
        for (int j = 0; j < nw; ++j) acc = fma(sw[j], f[-j], acc);
        out[i] = acc * scale;
    }
}

static unsigned grid_for(size_t n, int threads)
{
    size_t b = (n + threads - 1) / threads;
    const size_t cap = 148 * 16;                        // a few waves of the 148 SMs
    return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

static bool inv3(const double m[9], double out[9])
{
    const double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
    const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
    if (det == 0.0 || !isfinite(det)) return false;
    out[0] = (e * i - f * h) / det; out[1] = (c * h - b * i) / det; out[2] = (b * f - c * e) / det;
    out[3] = (f * g - d * i) / det; out[4] = (a * i - c * g) / det; out[5] = (c * d - a * f) / det;
    out[6] = (d * h - e * g) / det; out[7] = (b * g - a * h) / det; out[8] = (a * e - b * d) / det;
    return true;
}

extern "C" int dbspm_static_accumulate_f64(double *acc, const double *comp, double scale, size_t n,
                                           int init, void *stream)
{
    DBSPM_REQUIRE(acc && comp, DBSPM_EINVAL, "null pointer argument");
    if (n == 0) return DBSPM_OK;
    accumulate_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(acc, comp, scale, n, init);
    DBSPM_CUDA_TRY(cudaGetLastError());
    return DBSPM_OK;
}

// layout 1: y-fastest copy with a 3-plane periodic halo; layout 2: y-quads (four times the window)
// control block of the dynamic schedule (task counter + one progress word per patch of 32 columns), behind the re-laid potential
static size_t relax_ctrl_bytes(int nx, int ny)
{
    const size_t words = (size_t)nx * ny / 32 + (size_t)nx + ny + 8;          // >= patches of any patch shape, + counter
    return (words * sizeof(int) + 255) / 256 * 256;
}

// bytes of the re-laid potential alone (a multiple of 256, so the control block behind it is aligned)
static size_t relax_layout_bytes(int nx, int ny, int nzc, int layout)
{
    if (nx < 1 || ny < 1 || nzc < 1) return 0;
    size_t b = 0;
    if (layout == 1) b = (size_t)nx * (ny + 3) * nzc * sizeof(double);
    if (layout == 2) b = (long long)nx * ny * nzc * 4 < 2147483647LL ? (size_t)nx * ny * nzc * 4 * sizeof(double) : 0;
    return (b + 255) / 256 * 256;
}

extern "C" size_t dbspm_relax_workspace_bytes_layout(int nx, int ny, int nzc, int layout)
{
    const size_t b = relax_layout_bytes(nx, ny, nzc, layout);
    return b ? b + relax_ctrl_bytes(nx, ny) : 0;
}

// the preferred workspace: the quad layout where its 32-bit element offsets allow, else the y-fastest copy
extern "C" size_t dbspm_relax_workspace_bytes(int nx, int ny, int nzc)
{
    const size_t q = dbspm_relax_workspace_bytes_layout(nx, ny, nzc, 2);
    return q ? q : dbspm_relax_workspace_bytes_layout(nx, ny, nzc, 1);
}

static int relax_launch(const double *static_win, int nx, int ny, int nzc, int x_base, int nx_local,
                        int i_lo, int i_hi, int j_lo, int j_hi, int nz_relax,
                        double kappa, double rpivot, const double dr[3], const double *dR,
                        double xtol, double ftol, int maxfev,
                        double *out_E_theta_phi, double *out_cart, int *out_nfev,
                        void *workspace, size_t workspace_bytes, void *stream);

extern "C" int dbspm_relax_powell_f64(const double *static_win, int nx, int ny, int nzc,
                                      int i_lo, int i_hi, int j_lo, int j_hi, int nz_relax,
                                      double kappa, double rpivot, const double dr[3], const double *dR,
                                      double xtol, double ftol, int maxfev,
                                      double *out_E_theta_phi, double *out_cart, int *out_nfev,
                                      void *workspace, size_t workspace_bytes, void *stream)
{
    return relax_launch(static_win, nx, ny, nzc, 0, 0, i_lo, i_hi, j_lo, j_hi, nz_relax, kappa, rpivot, dr, dR, xtol, ftol, maxfev,
                        out_E_theta_phi, out_cart, out_nfev, workspace, workspace_bytes, stream);
}

// lateral halo (planes) a slab needs on each side of its tile: the lever reaches rpivot/dx planes (|sin| <= 1; row norms
// of inv(dR^T) for a skew cell) and the stencil two more
extern "C" int dbspm_relax_slab_halo(double rpivot, const double dr[3], const double *dR)
{
    if (!dr || !(dr[0] > 0) || !(dr[2] > 0) || !(rpivot >= 0)) return -1;
    double reach = (rpivot / dr[2]) * dr[2] / dr[0];
    if (dR) {
        const double dRT[9] = {dR[0], dR[3], dR[6], dR[1], dR[4], dR[7], dR[2], dR[5], dR[8]};
        double M[9];
        if (!inv3(dRT, M)) return -1;
        reach = (rpivot / dr[2]) * dr[2] * sqrt(M[0] * M[0] + M[1] * M[1] + M[2] * M[2]);
    }
    return (int)ceil(reach) + 2;
}

extern "C" int dbspm_relax_powell_slab_f64(const double *static_slab, int nx, int ny, int nzc, int x_base, int nx_local,
                                           int i_lo, int i_hi, int j_lo, int j_hi, int nz_relax,
                                           double kappa, double rpivot, const double dr[3], const double *dR,
                                           double xtol, double ftol, int maxfev,
                                           double *out_E_theta_phi, double *out_cart, int *out_nfev,
                                           void *workspace, size_t workspace_bytes, void *stream)
{
    DBSPM_REQUIRE(nx >= 1 && nx_local >= 1 && nx_local <= nx && x_base >= 0 && x_base < nx, DBSPM_EINVAL,
                  "bad slab: planes [%d, %d + %d) of %d", x_base, x_base, nx_local, nx);
    DBSPM_REQUIRE(dr, DBSPM_EINVAL, "null pointer argument");
    const int halo = dbspm_relax_slab_halo(rpivot, dr, dR);
    DBSPM_REQUIRE(halo >= 0, DBSPM_EINVAL, "bad rpivot / dr / dR");
    if (nx_local < nx) {
        // every plane a probe of the tile [i_lo, i_hi) can touch must be in the slab
        const int lo = ((i_lo - halo - x_base) % nx + nx) % nx, span = (i_hi - i_lo) + 2 * halo;
        DBSPM_REQUIRE(i_hi > i_lo ? lo + span <= nx_local : true, DBSPM_EINVAL,
                      "slab [%d, +%d) does not cover the tile [%d, %d) with its halo of %d planes", x_base, nx_local, i_lo, i_hi, halo);
    }
    return relax_launch(static_slab, nx, ny, nzc, x_base, nx_local, i_lo, i_hi, j_lo, j_hi, nz_relax, kappa, rpivot, dr, dR, xtol, ftol,
                        maxfev, out_E_theta_phi, out_cart, out_nfev, workspace, workspace_bytes, stream);
}

static int relax_launch(const double *static_win, int nx, int ny, int nzc, int x_base, int nx_local,
                        int i_lo, int i_hi, int j_lo, int j_hi, int nz_relax,
                        double kappa, double rpivot, const double dr[3], const double *dR,
                        double xtol, double ftol, int maxfev,
                        double *out_E_theta_phi, double *out_cart, int *out_nfev,
                        void *workspace, size_t workspace_bytes, void *stream)
{
    DBSPM_REQUIRE(static_win && dr && out_E_theta_phi, DBSPM_EINVAL, "null pointer argument");
    DBSPM_REQUIRE(nx >= 1 && ny >= 1 && nzc >= 1, DBSPM_EINVAL, "bad grid shape (%d,%d,%d)", nx, ny, nzc);
    DBSPM_REQUIRE(i_lo <= i_hi && j_lo <= j_hi && nz_relax >= 0, DBSPM_EINVAL, "bad tile");
    DBSPM_REQUIRE(dr[0] > 0 && dr[1] > 0 && dr[2] > 0, DBSPM_EINVAL, "grid spacings must be positive");
    DBSPM_REQUIRE(maxfev >= 1, DBSPM_EINVAL, "maxfev must be >= 1");
    const int nxs = nx_local ? nx_local : nx;              // planes actually stored
    DBSPM_REQUIRE((long long)nxs * (ny + 3) * nzc < 2147483647LL, DBSPM_EUNSUPPORTED,
                  "potential grid too large for 32-bit element offsets (%d x %d x %d)", nxs, ny, nzc);
    RelaxParams P;
    P.grid = static_win; P.nx = nx; P.ny = ny; P.nzc = nzc; P.nyp = ny + 3;
    P.x_base = x_base; P.nx_local = nx_local;
    P.i_lo = i_lo; P.i_hi = i_hi; P.j_lo = j_lo; P.j_hi = j_hi; P.nz_relax = nz_relax;
    P.kappa = kappa; P.rpivot = rpivot;
    P.npivot = rpivot / dr[2];
    P.rpivot_rt = P.npivot * dr[2];
    P.dr[0] = dr[0]; P.dr[1] = dr[1]; P.dr[2] = dr[2];
    for (int q = 0; q < 3; ++q) P.inv_dr[q] = 1.0 / dr[q];
    P.inv_n[0] = 1.0 / nx; P.inv_n[1] = 1.0 / ny; P.inv_n[2] = 1.0 / nzc;
    P.nd[0] = nx; P.nd[1] = ny; P.nd[2] = nzc;
    P.noortho = dR != nullptr;
    for (int q = 0; q < 9; ++q) P.Minv[q] = 0.0;
    if (dR) {
        const double dRT[9] = {dR[0], dR[3], dR[6], dR[1], dR[4], dR[7], dR[2], dR[5], dR[8]};
        DBSPM_REQUIRE(inv3(dRT, P.Minv), DBSPM_EINVAL, "dR is singular");
    }
    P.xtol = xtol; P.ftol = ftol; P.maxfev = maxfev; P.maxiter = 2000;   // N*1000 (_optimize.py:3514-3516)
    P.out_etp = out_E_theta_phi; P.out_cart = out_cart; P.out_nfev = out_nfev;
    const long long ncol = (long long)(i_hi - i_lo) * (j_hi - j_lo);
    if (ncol == 0 || nz_relax == 0) return DBSPM_OK;
    // columns of a warp: 32 consecutive y (foot_x = 1) or a foot_x x (32 / foot_x) patch (relax_kernel)
    // Measured on B200 with the y-quad layout (1024 x 1024 x 24 positions): 1 x 32 columns per warp 71.9 ms, 2 x 16 68.9, 4 x 8 66.9,
    // 8 x 4 65.7, 16 x 2 70.0, 32 x 1 81.1 -- x-neighbours reuse each other's sectors in L1, but every x plane of a load is
    // another 128-byte line, so the patch must stay wide enough in y to fill its lines.  The other layouts keep 1 x 32.
    P.foot_x = 1;
    if (workspace && workspace_bytes >= dbspm_relax_workspace_bytes_layout(nxs, ny, nzc, 2) && dbspm_relax_workspace_bytes_layout(nxs, ny, nzc, 2)) {
        P.foot_x = RELAX_FOOT_X_DEFAULT;
        while (P.foot_x > 1 && (P.foot_x > i_hi - i_lo || 32 / P.foot_x > j_hi - j_lo)) P.foot_x >>= 1;   // narrow tiles
    }
    if (const char *e = getenv("DBSPM_RELAX_FOOT_X")) { const int v = atoi(e); if (v == 1 || v == 2 || v == 4 || v == 8 || v == 16 || v == 32) P.foot_x = v; }
    const long long blocks = (relax_patch_threads(i_hi - i_lo, j_hi - j_lo, P.foot_x) + RELAX_THREADS - 1) / RELAX_THREADS;
    DBSPM_REQUIRE(blocks < 2147483647LL, DBSPM_EUNSUPPORTED, "tile too large");
    // Heights per launch (default: the whole sweep in one launch).  A launch can also continue a sweep: it finds its warm
    // start in out_etp.  One launch per height keeps every resident warp on the same few z rows (L2-resident working set
    // with the y-quad layout, which otherwise reads 21 GB from HBM), but measured slower on B200 (77.9 against 72.5 ms:
    // 24 launch tails cost more than the L2 misses); DBSPM_RELAX_HEIGHTS_PER_LAUNCH keeps the experiment reproducible.
    int hpl = nz_relax;
    if (const char *e = getenv("DBSPM_RELAX_HEIGHTS_PER_LAUNCH")) { const int v = atoi(e); if (v >= 1) hpl = v; }
    if (workspace) {
        // re-lay the potential once per call (~2-5 passes over the window): the largest layout the workspace holds
        const size_t need1 = dbspm_relax_workspace_bytes_layout(nxs, ny, nzc, 1), need2 = dbspm_relax_workspace_bytes_layout(nxs, ny, nzc, 2);
        DBSPM_REQUIRE(workspace_bytes >= need1 && ((uintptr_t)workspace & 31) == 0,
                      DBSPM_EWORKSPACE, "relax workspace needs %zu bytes (y-fastest copy) or %zu (y-quads), 32-byte aligned (got %zu)",
                      need1, need2, workspace_bytes);
        DBSPM_REQUIRE(nxs <= 65535, DBSPM_EUNSUPPORTED, "nx too large for the transpose grid");
        P.grid = (const double *)workspace;
        const bool quads = need2 && workspace_bytes >= need2;
        // dynamic schedule (relax_dyn_kernel): chunks of `hpc` heights pulled by one wave of persistent warps; the control
        // block sits behind the re-laid potential.  DBSPM_RELAX_CHUNK=0 keeps the static launches (A/B; with
        // DBSPM_RELAX_HEIGHTS_PER_LAUNCH), any other value sets the chunk length
        int hpc = 4;     // B200, 1024 x 1024 x 24: chunks of 2 ... 8 heights 63.9-64.1 ms, 12: 64.6, static launch 65.7
        if (const char *e = getenv("DBSPM_RELAX_CHUNK")) hpc = atoi(e);
        int *ctrl = (int *)((char *)workspace + relax_layout_bytes(nxs, ny, nzc, quads ? 2 : 1));
        const long long npatch = relax_patch_threads(i_hi - i_lo, j_hi - j_lo, P.foot_x) / 32 + (P.foot_x <= 1 && ncol % 32 ? 1 : 0);
        const int nchunks = hpc > 0 ? (nz_relax + hpc - 1) / hpc : 0;
        bool dynamic = hpc > 0 && getenv("DBSPM_RELAX_HEIGHTS_PER_LAUNCH") == nullptr && npatch * nchunks < 2147483647LL &&
                             (size_t)(npatch + 1) * sizeof(int) <= relax_ctrl_bytes(nxs, ny);
        int resident = 0;
        if (dynamic) {
            int dev = 0, sms = 0;
            DBSPM_CUDA_TRY(cudaGetDevice(&dev));
            DBSPM_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
            const long long warps = (long long)sms * RELAX_MIN_BLOCKS * (RELAX_THREADS / 32);
            // Only with more patches than resident warps.  Otherwise every patch has its own warp from the start, a static
            // launch IS the best schedule, and the chunk-major order would make early finishers wait for the patch whose
            // next chunk they drew (16-set pyridine sweep on four streams, 1225 patches per launch: 41.8 -> 44.4 ms).
            dynamic = npatch > warps || getenv("DBSPM_RELAX_CHUNK") != nullptr;
            resident = (int)((npatch < warps ? npatch : warps) + (RELAX_THREADS / 32) - 1) / (RELAX_THREADS / 32);
            if (dynamic) DBSPM_CUDA_TRY(cudaMemsetAsync(ctrl, 0, (size_t)(npatch + 1) * sizeof(int), (cudaStream_t)stream));
        }
        if (quads) {
            dim3 tg((unsigned)((nzc + 31) / 32), (unsigned)((ny + 31) / 32), (unsigned)nxs);
            quad_yz_kernel<<<tg, dim3(32, 8), 0, (cudaStream_t)stream>>>(static_win, (double *)workspace, ny, nzc);
            DBSPM_CUDA_TRY(cudaGetLastError());
            if (dynamic) relax_dyn_kernel<2><<<(unsigned)resident, RELAX_THREADS, 0, (cudaStream_t)stream>>>(P, ncol, (int)npatch, nchunks, hpc, ctrl);
            else for (int kh = nz_relax - 1; kh >= 0; kh -= hpl) {
                P.k_hi = kh; P.k_lo = kh - hpl + 1 > 0 ? kh - hpl + 1 : 0;
                relax_kernel<2><<<(unsigned)blocks, RELAX_THREADS, 0, (cudaStream_t)stream>>>(P, ncol);
            }
        } else {
            dim3 tg((unsigned)((nzc + 31) / 32), (unsigned)((ny + 3 + 31) / 32), (unsigned)nxs);
            transpose_yz_kernel<<<tg, dim3(32, 8), 0, (cudaStream_t)stream>>>(static_win, (double *)workspace, ny, nzc);
            DBSPM_CUDA_TRY(cudaGetLastError());
            if (dynamic) relax_dyn_kernel<1><<<(unsigned)resident, RELAX_THREADS, 0, (cudaStream_t)stream>>>(P, ncol, (int)npatch, nchunks, hpc, ctrl);
            else for (int kh = nz_relax - 1; kh >= 0; kh -= hpl) {
                P.k_hi = kh; P.k_lo = kh - hpl + 1 > 0 ? kh - hpl + 1 : 0;
                relax_kernel<1><<<(unsigned)blocks, RELAX_THREADS, 0, (cudaStream_t)stream>>>(P, ncol);
            }
        }
    } else {
        P.k_hi = nz_relax - 1; P.k_lo = 0;
        relax_kernel<0><<<(unsigned)blocks, RELAX_THREADS, 0, (cudaStream_t)stream>>>(P, ncol);
    }
    DBSPM_CUDA_TRY(cudaGetLastError());
    if (out_cart) {
        const long long npos = ncol * nz_relax;
        relax_cart_kernel<<<grid_for((size_t)npos, 256), 256, 0, (cudaStream_t)stream>>>(out_E_theta_phi, out_cart, npos, rpivot);
        DBSPM_CUDA_TRY(cudaGetLastError());
    }
    return DBSPM_OK;
}

extern "C" int dbspm_tricubic_gather_f64(const double *data, int nx, int ny, int nz,
                                         const double *pts, size_t npts, double *out, void *stream)
{
    DBSPM_REQUIRE(data && pts && out, DBSPM_EINVAL, "null pointer argument");
    DBSPM_REQUIRE(nx >= 1 && ny >= 1 && nz >= 1, DBSPM_EINVAL, "bad grid shape");
    DBSPM_REQUIRE((long long)nx * ny * nz < 2147483647LL, DBSPM_EUNSUPPORTED, "grid too large for 32-bit element offsets");
    if (npts == 0) return DBSPM_OK;
    gather_kernel<<<grid_for(npts, 256), 256, 0, (cudaStream_t)stream>>>(data, nx, ny, nz, pts, npts, out);
    DBSPM_CUDA_TRY(cudaGetLastError());
    return DBSPM_OK;
}

extern "C" int dbspm_dfz_gradient_f64(const double *E, int nx, int ny, int nz, double dz, double *out, void *stream)
{
    DBSPM_REQUIRE(E && out, DBSPM_EINVAL, "null pointer argument");
    DBSPM_REQUIRE(nx >= 1 && ny >= 1 && nz >= 1 && dz != 0.0, DBSPM_EINVAL, "bad arguments");
    const long long ncol = (long long)nx * ny;
    gradz_kernel<<<grid_for((size_t)ncol * nz, 256), 256, 0, (cudaStream_t)stream>>>(E, ncol, nz, -dz, out);
    DBSPM_CUDA_TRY(cudaGetLastError());
    return DBSPM_OK;
}

extern "C" int dbspm_convz_valid_f64(const double *F, int nx, int ny, int nz, const double *w, int nw, double scale,
                                     double *out, void *stream)
{
    DBSPM_REQUIRE(F && w && out, DBSPM_EINVAL, "null pointer argument");
    DBSPM_REQUIRE(nx >= 1 && ny >= 1 && nz >= 1 && nw >= 1 && nw <= 64 && nw <= nz, DBSPM_EINVAL,
                  "need 1 <= nw <= min(64, nz) (nw=%d, nz=%d)", nw, nz);
    const long long ncol = (long long)nx * ny;
    convz_valid_kernel<<<grid_for((size_t)ncol * (nz - nw + 1), 256), 256, 0, (cudaStream_t)stream>>>(F, ncol, nz, w, nw, scale, out);
    DBSPM_CUDA_TRY(cudaGetLastError());
    return DBSPM_OK;
}
