// corr3d.cu -- sm_100a kernels + C ABI for the sr / es interaction grids.
// Replaces pydbspm/interactions.py:9-22 and the call sites dbspm:329-330, dbspm:348-349.
// Kernel bodies live in corr_body.h (shared with the CPU emulation used by the tests).
#include <map>
#include <mutex>
#include "capi_common.h"
#include "corr_body.h"
#include "plan.h"

#define DBSPM_THREADS 256
#ifndef AXIS2_THREADS
#define AXIS2_THREADS 256
#endif
#ifndef AXIS2_MIN_BLOCKS
#define AXIS2_MIN_BLOCKS 2
#endif

template <int D>
__global__ void __launch_bounds__(DBSPM_THREADS) axis_pass_kernel(const __grid_constant__ AxisParams P)
{
    extern __shared__ __align__(16) unsigned char smem[];
    axis_pass_body<D>(P, smem, (long long)blockIdx.x, (int)blockIdx.y, (int)threadIdx.x, (int)blockDim.x);
}

template <int D>
__global__ void __launch_bounds__(AXIS2_THREADS, AXIS2_MIN_BLOCKS) axis_pass2_kernel(const __grid_constant__ AxisParams P)
{
    extern __shared__ __align__(16) unsigned char smem[];
    axis_pass_body_v2<D>(P, smem, (long long)blockIdx.x, (int)blockIdx.y, (int)threadIdx.x, (int)blockDim.x);
}

__global__ void __launch_bounds__(DBSPM_THREADS) zfused_kernel(const __grid_constant__ ZFusedParams P)
{
    extern __shared__ __align__(16) unsigned char smem[];
    zfused_body(P, smem, (long long)blockIdx.x, (int)threadIdx.x, (int)blockDim.x);
}

#ifndef ZF2_PAIRS
#define ZF2_PAIRS 4
#endif
#ifndef ZF2_THREADS
#define ZF2_THREADS 128
#endif
#ifndef ZF2_MIN_BLOCKS
#define ZF2_MIN_BLOCKS 1
#endif
__global__ void __launch_bounds__(ZF2_THREADS, ZF2_MIN_BLOCKS) zfused2_kernel(const __grid_constant__ ZFusedParams P)
{
    extern __shared__ __align__(16) unsigned char smem[];
    zfused_body_v2<ZF2_PAIRS>(P, smem, (long long)blockIdx.x, (long long)gridDim.x, (int)threadIdx.x, (int)blockDim.x);
}

__global__ void __launch_bounds__(DBSPM_THREADS) compact_kernel(const double *W, double *E, long long ncol, int nzw, int ldw)
{
    compact_body(W, E, ncol, nzw, ldw, (long long)blockIdx.x * blockDim.x + threadIdx.x,
                 (long long)gridDim.x * blockDim.x);
}

// ---- device-resident plan cache (twiddle + position tables; a few KB per distinct length) --
struct DevPlan {
    HostPlan host;
    cplx *tw = nullptr;
    int *pos_of = nullptr;
    int *freq_of = nullptr;
};

static std::mutex g_plan_mu;
static std::map<std::pair<int, int>, DevPlan *> g_plans;   // (device, n + 65536 * max radix)

static int get_plan(int n, cudaStream_t st, DevPlan **out, int max_radix = 16)
{
    int dev = 0;
    DBSPM_CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_plan_mu);
    const int key = n + 65536 * max_radix;
    auto it = g_plans.find({dev, key});
    if (it != g_plans.end()) { *out = it->second; return DBSPM_OK; }
    DevPlan *p = new DevPlan();
    if (!dbspm_make_plan(n, p->host, max_radix)) {
        delete p;
        dbspm_set_error("FFT length %d has a prime factor > %d (unsupported)", n, DBSPM_MAX_RADIX);
        return DBSPM_EUNSUPPORTED;
    }
    DBSPM_CUDA_TRY(cudaMalloc((void **)&p->tw, sizeof(cplx) * (size_t)n));
    DBSPM_CUDA_TRY(cudaMalloc((void **)&p->pos_of, sizeof(int) * (size_t)n));
    // the host vectors live as long as the cache entry, so the async copies stay valid
    DBSPM_CUDA_TRY(cudaMemcpyAsync(p->tw, p->host.tw.data(), sizeof(cplx) * (size_t)n, cudaMemcpyHostToDevice, st));
    DBSPM_CUDA_TRY(cudaMemcpyAsync(p->pos_of, p->host.pos_of.data(), sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, st));
    DBSPM_CUDA_TRY(cudaMalloc((void **)&p->freq_of, sizeof(int) * (size_t)n));
    DBSPM_CUDA_TRY(cudaMemcpyAsync(p->freq_of, p->host.freq_of.data(), sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, st));
    // tables may be used from other streams later: make them visible before returning
    DBSPM_CUDA_TRY(cudaStreamSynchronize(st));
    g_plans[{dev, key}] = p;
    *out = p;
    return DBSPM_OK;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// tile width for an axis pass: widest of 8,4,2,1 lines whose tile leaves room for 2 CTAs/SM
static int pick_tshift(int n, long long inner)
{
    for (int ts = 3; ts >= 0; --ts) {
        if ((1LL << ts) > inner && ts > 0) continue;
        if (axis_smem_bytes(n, ts) <= 100 * 1024) return ts;
    }
    return axis_smem_bytes(n, 0) <= 227 * 1024 ? 0 : -1;
}

template <int D>
static int launch_axis(const cplx *in0, cplx *out0, const cplx *in1, cplx *out1, int narrays,
                       double alpha0, double alpha1, int use_pow, int n, long long inner, long long outer,
                       cudaStream_t st, int flags)
{
    DevPlan *pl = nullptr;
    int rc = get_plan(n, st, &pl);
    if (rc) return rc;
    AxisParams P;
    P.in[0] = in0; P.in[1] = in1; P.out[0] = out0; P.out[1] = out1;
    P.alpha[0] = alpha0; P.alpha[1] = alpha1;
    P.use_pow = use_pow; P.narrays = narrays;
    P.n = n; P.inner = inner; P.outer = outer;
    P.plan = pl->host.dev; P.tw = pl->tw; P.pos_of = pl->pos_of; P.freq_of = pl->freq_of;
    P.v2 = axis_v2_ok(P.plan) && !(flags & DBSPM_CORR_AXIS_V1);
    P.tw_smem = 0;
    size_t smem;
    if (P.v2) {
        axis2_config(n, inner, P.plan.nstages, &P.tshift, &P.tw_smem);
        smem = axis2_smem_bytes(n, P.tshift, P.plan.nstages, P.tw_smem);
        if (smem > 227 * 1024) P.v2 = 0;
    }
    if (!P.v2) {
        P.tshift = pick_tshift(n, inner);
        DBSPM_REQUIRE(P.tshift >= 0, DBSPM_EUNSUPPORTED, "axis length %d does not fit in shared memory", n);
        smem = axis_smem_bytes(n, P.tshift);
    }
    P.tiles = (inner + (1LL << P.tshift) - 1) >> P.tshift;
    const long long nblocks = outer * P.tiles;
    DBSPM_REQUIRE(nblocks > 0 && nblocks < 2147483647LL, DBSPM_EUNSUPPORTED, "grid too large (%lld blocks)", nblocks);
    if (P.v2) {
        DBSPM_CUDA_TRY(cudaFuncSetAttribute(axis_pass2_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        axis_pass2_kernel<D><<<dim3((unsigned)nblocks, (unsigned)narrays), AXIS2_THREADS, smem, st>>>(P);
    } else {
        DBSPM_CUDA_TRY(cudaFuncSetAttribute(axis_pass_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        axis_pass_kernel<D><<<dim3((unsigned)nblocks, (unsigned)narrays), DBSPM_THREADS, smem, st>>>(P);
    }
    DBSPM_CUDA_TRY(cudaGetLastError());
    return DBSPM_OK;
}

extern "C" size_t dbspm_corr3d_workspace_bytes(int nx, int ny, int nz, int z_lo, int z_hi, int flags)
{
    (void)flags;
    if (nx < 1 || ny < 1 || nz < 2 || z_lo < 0 || z_hi <= z_lo || z_hi > nz) return 0;
    const size_t N = (size_t)nx * ny * nz;
    const int nzw = z_hi - z_lo, nwh = (nzw + 1) / 2;
    size_t bytes = 2 * align_up(N * 8, 256);
    if (nzw & 1) bytes += align_up((size_t)nx * ny * nwh * 16, 256);
    return bytes;
}

extern "C" int dbspm_corr3d_f64(const double *A, const double *B, double alphaA, double alphaB, double dV,
                                int nx, int ny, int nz, int z_lo, int z_hi, double *E_win,
                                void *workspace, size_t workspace_bytes, int flags, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    DBSPM_REQUIRE(A && B && E_win && workspace, DBSPM_EINVAL, "null pointer argument");
    DBSPM_REQUIRE(nx >= 1 && ny >= 1 && nz >= 2, DBSPM_EINVAL, "bad shape (%d,%d,%d)", nx, ny, nz);
    DBSPM_REQUIRE(z_lo >= 0 && z_hi > z_lo && z_hi <= nz, DBSPM_EINVAL, "bad window [%d,%d) for nz=%d", z_lo, z_hi, nz);
    DBSPM_REQUIRE((nz & 1) == 0, DBSPM_EUNSUPPORTED, "nz=%d is odd: the packed real transform needs an even z length", nz);
    DBSPM_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0 && ((uintptr_t)E_win & 15) == 0,
                  DBSPM_EINVAL, "A, B and E_win must be 16-byte aligned");
    const size_t need = dbspm_corr3d_workspace_bytes(nx, ny, nz, z_lo, z_hi, flags);
    DBSPM_REQUIRE(workspace_bytes >= need && ((uintptr_t)workspace & 255) == 0, DBSPM_EWORKSPACE,
                  "workspace needs %zu bytes, 256-byte aligned (got %zu)", need, workspace_bytes);

    const int n = nz / 2, nzw = z_hi - z_lo, nwh = (nzw + 1) / 2;
    const size_t N = (size_t)nx * ny * nz;
    unsigned char *ws = (unsigned char *)workspace;
    cplx *WA = (cplx *)ws;
    cplx *WB = (cplx *)(ws + align_up(N * 8, 256));
    cplx *W = (nzw & 1) ? (cplx *)(ws + 2 * align_up(N * 8, 256)) : (cplx *)E_win;

    int rc;
    // forward x (out of place, power fused into the load), then y in place
    if (!(flags & DBSPM_CORR_SKIP_XFWD)) {
        rc = launch_axis<-1>((const cplx *)A, WA, (const cplx *)B, WB, 2, alphaA, alphaB, 1, nx, (long long)ny * n, 1, st, flags);
        if (rc) return rc;
    }
    if (!(flags & DBSPM_CORR_SKIP_YFWD)) {
        rc = launch_axis<-1>(WA, WA, WB, WB, 2, 1.0, 1.0, 0, ny, n, nx, st, flags);
        if (rc) return rc;
    }

    // fused z: forward, untangle, product, window phase, pruned inverse
    if (!(flags & DBSPM_CORR_SKIP_Z)) {
        DevPlan *pl = nullptr, *plN = nullptr;
        rc = get_plan(n, st, &pl, ZF2_MAX_RADIX);
        if (rc) return rc;
        rc = get_plan(nz, st, &plN);
        if (rc) return rc;
        ZFusedParams P;
        P.ZA = WA; P.ZB = WB; P.W = W;
        P.nx = nx; P.ny = ny; P.n = n; P.nwh = nwh; P.z_lo = z_lo;
        P.scale = dV / ((double)nx * (double)ny * (double)nz) * 0.25;
        P.plan = pl->host.dev; P.tw = pl->tw; P.pos_of = pl->pos_of; P.freq_of = pl->freq_of; P.twN = plN->tw;
        P.npairs = (long long)(nx / 2 + 1) * ny;
        const bool v2 = axis_v2_ok(P.plan) && zfused2_smem_bytes<ZF2_PAIRS>(n) <= 113 * 1024 &&
                        !(flags & DBSPM_CORR_AXIS_V1);
        if (v2) {
            const long long ngroups = (P.npairs + ZF2_PAIRS - 1) / ZF2_PAIRS;
            const size_t smem = zfused2_smem_bytes<ZF2_PAIRS>(n);
            // persistent blocks: a few per SM (148 SMs), each walking over its share of the pair groups
            int per_sm = (int)((220 * 1024) / (smem + 1024));
            per_sm = per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm);
            long long nblocks = 148LL * per_sm * 2;
            if (nblocks > ngroups) nblocks = ngroups;
            DBSPM_CUDA_TRY(cudaFuncSetAttribute(zfused2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            DBSPM_CUDA_TRY(cudaFuncSetAttribute(zfused2_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            zfused2_kernel<<<(unsigned)nblocks, ZF2_THREADS, smem, st>>>(P);
        } else {
            const long long nblocks = (P.npairs + ZL - 1) / ZL;
            const size_t smem = zfused_smem_bytes(n);
            DBSPM_REQUIRE(smem <= 227 * 1024, DBSPM_EUNSUPPORTED, "nz=%d too long for the fused z kernel", nz);
            DBSPM_REQUIRE(nblocks < 2147483647LL, DBSPM_EUNSUPPORTED, "grid too large");
            DBSPM_CUDA_TRY(cudaFuncSetAttribute(zfused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            DBSPM_CUDA_TRY(cudaFuncSetAttribute(zfused_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            zfused_kernel<<<(unsigned)nblocks, DBSPM_THREADS, smem, st>>>(P);
        }
        DBSPM_CUDA_TRY(cudaGetLastError());
    }

    // inverse y, then x on the packed window (in place)
    if (!(flags & DBSPM_CORR_SKIP_YINV)) {
        rc = launch_axis<1>(W, W, nullptr, nullptr, 1, 1.0, 1.0, 0, ny, nwh, nx, st, flags);
        if (rc) return rc;
    }
    if (!(flags & DBSPM_CORR_SKIP_XINV)) {
        rc = launch_axis<1>(W, W, nullptr, nullptr, 1, 1.0, 1.0, 0, nx, (long long)ny * nwh, 1, st, flags);
        if (rc) return rc;
    }

    if (nzw & 1) {
        const long long ncol = (long long)nx * ny;
        long long blocks = (ncol * nzw + DBSPM_THREADS - 1) / DBSPM_THREADS;
        if (blocks > 148LL * 32) blocks = 148LL * 32;
        compact_kernel<<<(unsigned)blocks, DBSPM_THREADS, 0, st>>>((const double *)W, E_win, ncol, nzw, 2 * nwh);
        DBSPM_CUDA_TRY(cudaGetLastError());
    }
    return DBSPM_OK;
}
