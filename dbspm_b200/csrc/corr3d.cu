// corr3d.cu -- sm_100a kernels + C ABI for the sr / es interaction grids.
// Replaces pydbspm/interactions.py:9-22 and the call sites dbspm:329-330, dbspm:348-349.
// Kernel bodies live in corr_body.h (shared with the CPU emulation used by the tests).
#include <cuda.h>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>
#include "capi_common.h"
#include "corr_body.h"
#include "direct_plan.h"
#include "plan.h"

#define DBSPM_THREADS 256
#ifndef AXIS2_THREADS
#define AXIS2_THREADS 256
#endif
#ifndef AXIS2_MIN_BLOCKS
#define AXIS2_MIN_BLOCKS 2
#endif
#ifndef AXIS_PLAN_MAXR
#define AXIS_PLAN_MAXR 16   /* power-of-two radix cap of the axis-pass plans (experiment: 8 with three resident CTAs) */
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

// wide variant for the slowest axis at long lengths (n ~ 1024): one CTA of 512 threads per SM with an 8-line
// tile, so that every row segment is a full 128-byte line (rows of the x pass are megabytes apart)
template <int D>
__global__ void __launch_bounds__(512, 1) axis_pass2w_kernel(const __grid_constant__ AxisParams P)
{
    extern __shared__ __align__(16) unsigned char smem[];
    axis_pass_body_v2<D>(P, smem, (long long)blockIdx.x, (int)blockIdx.y, (int)threadIdx.x, (int)blockDim.x);
}

// persistent, software-pipelined variants (axis_pass_body_v2s, one instantiation per plan shape): blockIdx.x walks the tiles
// with stride gridDim.x
template <int D, int R0, int RM, int RL, bool POW>
__global__ void __launch_bounds__(AXIS2_THREADS, AXIS2_MIN_BLOCKS) axis_pass2s_kernel(const __grid_constant__ AxisParams P)
{
    extern __shared__ __align__(16) unsigned char smem[];
    axis_pass_body_v2s<D, R0, RM, RL, POW>(P, smem, (long long)blockIdx.x, (long long)gridDim.x, (int)blockIdx.y, (int)threadIdx.x, (int)blockDim.x);
}

template <int D, int R0, int RM, int RL, bool POW>
__global__ void __launch_bounds__(512, 1) axis_pass2sw_kernel(const __grid_constant__ AxisParams P)
{
    extern __shared__ __align__(16) unsigned char smem[];
    axis_pass_body_v2s<D, R0, RM, RL, POW>(P, smem, (long long)blockIdx.x, (long long)gridDim.x, (int)blockIdx.y, (int)threadIdx.x, (int)blockDim.x);
}

template <int D, int R0, int RM, int RL, bool POW>
static int launch_axis_pipelined(const AxisParams &P, unsigned gx, int narrays, bool wide, size_t smem, cudaStream_t st)
{
    if (wide) {
        DBSPM_CUDA_TRY(cudaFuncSetAttribute(axis_pass2sw_kernel<D, R0, RM, RL, POW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        axis_pass2sw_kernel<D, R0, RM, RL, POW><<<dim3(gx, (unsigned)narrays), 512, smem, st>>>(P);
    } else {
        DBSPM_CUDA_TRY(cudaFuncSetAttribute(axis_pass2s_kernel<D, R0, RM, RL, POW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        axis_pass2s_kernel<D, R0, RM, RL, POW><<<dim3(gx, (unsigned)narrays), AXIS2_THREADS, smem, st>>>(P);
    }
    return DBSPM_OK;
}

// row-mapped variant (distributed correlation): rows of the transform axis are read / written through a table
template <int D>
__global__ void __launch_bounds__(AXIS2_THREADS, AXIS2_MIN_BLOCKS) axis_pass2m_kernel(const __grid_constant__ AxisParams P)
{
    extern __shared__ __align__(16) unsigned char smem[];
    axis_pass_body_v2<D, 1>(P, smem, (long long)blockIdx.x, (int)blockIdx.y, (int)threadIdx.x, (int)blockDim.x);
}

// fused-exchange variant: every output row is stored straight into its owner's receive buffer (peer memory, NVLink)
template <int D>
__global__ void __launch_bounds__(AXIS2_THREADS, AXIS2_MIN_BLOCKS) axis_pass2p_kernel(const __grid_constant__ AxisParams P)
{
    extern __shared__ __align__(16) unsigned char smem[];
    axis_pass_body_v2<D, 2>(P, smem, (long long)blockIdx.x, (int)blockIdx.y, (int)threadIdx.x, (int)blockDim.x);
}

// same with 8-line tiles (one CTA of 512 threads per SM): every remote store is a full 128-byte segment
template <int D>
__global__ void __launch_bounds__(512, 1) axis_pass2pw_kernel(const __grid_constant__ AxisParams P)
{
    extern __shared__ __align__(16) unsigned char smem[];
    axis_pass_body_v2<D, 2>(P, smem, (long long)blockIdx.x, (int)blockIdx.y, (int)threadIdx.x, (int)blockDim.x);
}

__global__ void __launch_bounds__(DBSPM_THREADS) zfused_kernel(const __grid_constant__ ZFusedParams P)
{
    extern __shared__ __align__(16) unsigned char smem[];
    zfused_body(P, smem, (long long)blockIdx.x, (int)threadIdx.x, (int)blockDim.x);
}

// TMA-fed persistent axis pass (axis_tma_body): no global load/store instructions in the block
#ifndef AXIS3_L2_PROMOTION
#define AXIS3_L2_PROMOTION CU_TENSOR_MAP_L2_PROMOTION_NONE
#endif
#define AXIS3_THREADS AXIS3_BLOCK_THREADS   /* consumer groups + producer warp (corr_body.h) */
#ifndef AXIS3_MIN_BLOCKS
#define AXIS3_MIN_BLOCKS 1
#endif
template <int D>
__global__ void __launch_bounds__(AXIS3_THREADS, AXIS3_MIN_BLOCKS)
axis_tma_kernel(const __grid_constant__ AxisTmaParams P, const __grid_constant__ CUtensorMap l0, const __grid_constant__ CUtensorMap s0,
                const __grid_constant__ CUtensorMap l1, const __grid_constant__ CUtensorMap s1)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const void *lm[2] = {&l0, &l1};
    const void *sm[2] = {&s0, &s1};
    axis_tma_body<D>(P, lm, sm, smem, (long long)blockIdx.x, (long long)gridDim.x, (int)threadIdx.x, (int)blockDim.x);
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
#ifndef ZF2_MAX_PER_SM
#define ZF2_MAX_PER_SM 4      /* cap on resident persistent blocks per SM */
#endif
#ifndef ZF2_WAVES
#define ZF2_WAVES 2
#endif
// MAXR = largest radix the kernel is built for (8 or 16): the radix-16 butterfly costs 166 registers, taken or not
template <int MAXR>
__global__ void __launch_bounds__(ZF2_THREADS, ZF2_MIN_BLOCKS) zfused2_kernel(const __grid_constant__ ZFusedParams P)
{
    extern __shared__ __align__(16) unsigned char smem[];
    zfused_body_v2<ZF2_PAIRS, MAXR>(P, smem, (long long)blockIdx.x, (long long)gridDim.x, (int)threadIdx.x, (int)blockDim.x);
}

__global__ void __launch_bounds__(DBSPM_THREADS) compact_kernel(const double *W, double *E, long long ncol, int nzw, int ldw)
{
    compact_body(W, E, ncol, nzw, ldw, (long long)blockIdx.x * blockDim.x + threadIdx.x,
                 (long long)gridDim.x * blockDim.x);
}

__global__ void __launch_bounds__(256) r2c_kernel(const double *in, cplx *out, size_t n, double alpha)
{
    r2c_body(in, out, n, alpha, (size_t)blockIdx.x * blockDim.x + threadIdx.x, (size_t)gridDim.x * blockDim.x);
}

__global__ void __launch_bounds__(256) cmulconj_kernel(cplx *a, const cplx *b, size_t n, double scale)
{
    cmulconj_body(a, b, n, scale, (size_t)blockIdx.x * blockDim.x + threadIdx.x, (size_t)gridDim.x * blockDim.x);
}

__global__ void __launch_bounds__(256) c2r_window_kernel(const cplx *W, double *E, long long ncol, int nz, int z_lo, int nzw)
{
    c2r_window_body(W, E, ncol, nz, z_lo, nzw, (long long)blockIdx.x * blockDim.x + threadIdx.x,
                    (long long)gridDim.x * blockDim.x);
}

// which z-planes of a (nlines, nz) array hold a nonzero value: flags[z] = 1 (flags zeroed by the caller).
// Streaming 128-bit reads, per-block flags in shared memory, one global store per set plane and block.
__global__ void __launch_bounds__(256) zsupport_kernel(const double2 *B2, size_t npairs, int nzh, int *flags)
{
    extern __shared__ unsigned char seen[];               // 2*nzh
    for (int i = threadIdx.x; i < 2 * nzh; i += blockDim.x) seen[i] = 0;
    __syncthreads();
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npairs; i += stride) {
        const double2 v = __ldcs(B2 + i);
        const int zp = (int)(i % (size_t)nzh);
        if (v.x != 0.0) seen[2 * zp] = 1;
        if (v.y != 0.0) seen[2 * zp + 1] = 1;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * nzh; i += blockDim.x)
        if (seen[i]) flags[i] = 1;
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

// device-resident row tables of the distributed correlation: map[K] = dist_row_of(K, n, G, blockrows)
static std::map<std::tuple<int, int, int, long long>, int *> g_rowmaps;

static int get_rowmap(int n, int G, long long blockrows, cudaStream_t st, const int **out)
{
    int dev = 0;
    DBSPM_CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_plan_mu);
    const auto key = std::make_tuple(dev, n, G, blockrows);
    auto it = g_rowmaps.find(key);
    if (it != g_rowmaps.end()) { *out = it->second; return DBSPM_OK; }
    DBSPM_REQUIRE((long long)G * blockrows < 2147483647LL, DBSPM_EUNSUPPORTED, "row table overflows 32 bits");
    std::vector<int> h((size_t)n);
    for (int K = 0; K < n; ++K)
        h[K] = blockrows == -2 ? dist_block_entry(K, n, G) : (blockrows < 0 ? dist_peer_entry(K, n, G) : dist_row_of(K, n, G, blockrows));
    int *d = nullptr;
    DBSPM_CUDA_TRY(cudaMalloc((void **)&d, sizeof(int) * (size_t)n));
    DBSPM_CUDA_TRY(cudaMemcpyAsync(d, h.data(), sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, st));
    DBSPM_CUDA_TRY(cudaStreamSynchronize(st));
    g_rowmaps[key] = d;
    *out = d;
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

// ---- tensor maps for the TMA axis pass -----------------------------------------------------
typedef CUresult (*TmapEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static TmapEncodeFn tmap_encoder()
{
    static TmapEncodeFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        cudaDriverEntryPointQueryResult q;
        void *p = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (TmapEncodeFn)p;
    }
    return fn;
}

// load map: (2*inner doubles, n_lo rows, n_hi row groups, outer slabs), box = (2T, n_lo, n_hi, 1): natural row order
// store map: (2*inner, r_s, ..., r_1, outer) with the radix digits' natural global strides, box = (2T, r_s, ..., r_1, 1):
//            tile row k_1*(n/r_1) + ... + k_s lands on global row k_1 + r_1*k_2 + ... (mixed-radix digit reversal)
static bool make_axis_maps(TmapEncodeFn enc, const cplx *in, cplx *out, int n, long long inner, long long outer, int T,
                           const FftPlanDev &pl, CUtensorMap *lmap, CUtensorMap *smap)
{
    int n_lo = 1;
    for (int d = 256; d >= 1; --d)
        if (n % d == 0) { n_lo = d; break; }
    const int n_hi = n / n_lo;
    if (n_hi > 256 || pl.nstages < 1 || pl.nstages > 3) return false;
    const cuuint64_t rowb = (cuuint64_t)inner * 16;
    cuuint32_t ones[5] = {1, 1, 1, 1, 1};
    {
        cuuint64_t dims[4] = {(cuuint64_t)(2 * inner), (cuuint64_t)n_lo, (cuuint64_t)n_hi, (cuuint64_t)outer};
        cuuint64_t str[3] = {rowb, rowb * n_lo, rowb * n};
        cuuint32_t box[4] = {(cuuint32_t)(2 * T), (cuuint32_t)n_lo, (cuuint32_t)n_hi, 1};
        if (enc(lmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, (void *)in, dims, str, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, AXIS3_L2_PROMOTION, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return false;
    }
    {
        const int ns = pl.nstages;
        cuuint64_t dims[5], str[4];
        cuuint32_t box[5];
        dims[0] = (cuuint64_t)(2 * inner); box[0] = (cuuint32_t)(2 * T);
        // digit i (1-based) has global stride r_1*...*r_{i-1} rows; listed innermost-first as s, s-1, ..., 1
        long long mult[4] = {1, 1, 1, 1};
        for (int i = 1; i < ns; ++i) mult[i] = mult[i - 1] * pl.radix[i - 1];
        for (int d = 1; d <= ns; ++d) {
            const int i = ns - d;                       // digit index 0-based
            dims[d] = (cuuint64_t)pl.radix[i]; box[d] = (cuuint32_t)pl.radix[i];
            str[d - 1] = rowb * (cuuint64_t)mult[i];
        }
        dims[ns + 1] = (cuuint64_t)outer; box[ns + 1] = 1;
        str[ns] = rowb * n;
        if (enc(smap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, (cuuint32_t)(ns + 2), (void *)out, dims, str, box, ones,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return false;
    }
    return true;
}

template <int D>
static int launch_axis_tma(const cplx *in0, cplx *out0, const cplx *in1, cplx *out1, int narrays, double alpha0, double alpha1,
                           int use_pow, int n, long long inner, long long outer, DevPlan *pl, cudaStream_t st, bool *done)
{
    *done = false;
    TmapEncodeFn enc = tmap_encoder();
    if (!enc || !axis_v2_ok(pl->host.dev) || pl->host.dev.nstages > 3) return DBSPM_OK;
    if (2 * inner >= (1LL << 31) || outer >= (1LL << 31)) return DBSPM_OK;
    const int ts = axis3_config(n, inner);
    if (ts < 0) return DBSPM_OK;
    AxisTmaParams P;
    P.in[0] = in0; P.in[1] = in1; P.out[0] = out0; P.out[1] = out1;
    P.alpha[0] = alpha0; P.alpha[1] = alpha1; P.use_pow = use_pow; P.narrays = narrays;
    P.n = n; P.tshift = ts; P.inner = inner; P.outer = outer;
    P.tiles = (inner + (1LL << ts) - 1) >> ts;
    P.plan = pl->host.dev; P.tw = pl->tw; P.freq_of = pl->freq_of;
    CUtensorMap lm[2], sm[2];
    for (int a = 0; a < 2; ++a) {
        const int w = a < narrays ? a : 0;
        if (!make_axis_maps(enc, w == 0 ? in0 : in1, w == 0 ? out0 : out1, n, inner, outer, 1 << ts, P.plan, &lm[a], &sm[a]))
            return DBSPM_OK;                 // fall back to the register-staged kernel
    }
    const size_t smem = axis3_smem_bytes(n, ts);
    const long long total = (long long)narrays * outer * P.tiles;
    long long nblocks = 148LL * (smem <= 113 * 1024 ? 2 : 1);
    if (nblocks > total) nblocks = total;
    DBSPM_CUDA_TRY(cudaFuncSetAttribute(axis_tma_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    axis_tma_kernel<D><<<(unsigned)nblocks, AXIS3_THREADS, smem, st>>>(P, lm[0], sm[0], lm[1], sm[1]);
    DBSPM_CUDA_TRY(cudaGetLastError());
    *done = true;
    return DBSPM_OK;
}

// z-range cut of the caller's arrays read by the first pass of the pruned correlation (planes)
struct GatherSpec {
    int src_nz, dst_nz;
    int z0[2], valid[2];
};

// row tables of a mapped pass (distributed correlation); null table = natural rows
struct AxisMap {
    const int *in_map, *out_map;
    long long in_ostride, out_ostride;      // elements between consecutive slabs (0 = n * inner)
    int npeers;                             // > 0: fused exchange, out_map holds dist_peer_entry values
    cplx *peer[2][DBSPM_MAX_PEERS];
    long long out_base_off;
};

template <int D>
static int launch_axis(const cplx *in0, cplx *out0, const cplx *in1, cplx *out1, int narrays,
                       double alpha0, double alpha1, int use_pow, int n, long long inner, long long outer,
                       cudaStream_t st, int flags, const GatherSpec *g = nullptr, const AxisMap *mp = nullptr)
{
    DevPlan *pl = nullptr;
    int rc = get_plan(n, st, &pl, AXIS_PLAN_MAXR);
    if (rc) return rc;
    AxisParams P;
    P.in[0] = in0; P.in[1] = in1; P.out[0] = out0; P.out[1] = out1;
    P.alpha[0] = alpha0; P.alpha[1] = alpha1;
    P.use_pow = use_pow; P.narrays = narrays;
    P.n = n; P.inner = inner; P.outer = outer;
    // path picked per shape: the TMA-fed persistent kernel wins where a ring of three 8-line tiles fits
    // (192 <= n <= 576; measured on B200 after its tile accesses became real LDS/STS: es 0.346 against 0.374 ms at
    // 196x196x240, x pass 0.155 against 0.189 ms at n = 240, 0.186 against 0.228 ms at n = 384); the register-staged
    // kernel wins at n = 1024 (4-line tiles: 2.50 against 3.31 ms)
    const bool tma_auto = n >= 192 && axis3_config(n, inner) == 3;
    P.gather = 0;
    P.in_map = P.out_map = nullptr;
    P.in_ostride = P.out_ostride = (long long)n * inner;
    if (mp) {
        DBSPM_REQUIRE(axis_v2_ok(pl->host.dev), DBSPM_EUNSUPPORTED, "row-mapped pass needs an all-fast-radix plan (n=%d)", n);
        P.in_map = mp->in_map; P.out_map = mp->out_map;
        if (mp->in_ostride) P.in_ostride = mp->in_ostride;
        if (mp->out_ostride) P.out_ostride = mp->out_ostride;
        P.out_base_off = 0;
        if (mp->npeers > 0) {
            for (int w = 0; w < 2; ++w)
                for (int q = 0; q < DBSPM_MAX_PEERS; ++q) P.out_peer[w][q] = q < mp->npeers ? mp->peer[w][q] : nullptr;
            P.out_base_off = mp->out_base_off;
        }
    }
    if (g) {
        // only the register-staged body can remap its loads
        DBSPM_REQUIRE(axis_v2_ok(pl->host.dev), DBSPM_EUNSUPPORTED, "pruned first pass needs an all-fast-radix plan (n=%d)", n);
        P.gather = 1;
        P.g_src_nzh = g->src_nz / 2; P.g_dst_nzh = g->dst_nz / 2;
        P.g_src_inner = inner / P.g_dst_nzh * P.g_src_nzh;
        for (int w = 0; w < 2; ++w) { P.g_zp0[w] = g->z0[w] / 2; P.g_valid[w] = g->valid[w] / 2; }
    }
    if (!g && !mp && ((flags & DBSPM_CORR_USE_TMA) || (tma_auto && !(flags & DBSPM_CORR_NO_TMA))) && !(flags & DBSPM_CORR_AXIS_V1)) {
        bool done = false;
        rc = launch_axis_tma<D>(in0, out0, in1, out1, narrays, alpha0, alpha1, use_pow, n, inner, outer, pl, st, &done);
        if (rc || done) return rc;
    }
    P.plan = pl->host.dev; P.tw = pl->tw; P.pos_of = pl->pos_of; P.freq_of = pl->freq_of;
    P.v2 = axis_v2_ok(P.plan) && (g || mp || !(flags & DBSPM_CORR_AXIS_V1));
    P.tw_smem = 0;
    size_t smem;
    bool wide = false;
    // persistent software-pipelined blocks (axis_pass_body_v2s) for the plain passes; not with the power fused into the
    // load (the pass is FP64-issue bound and measured slower pipelined), see DBSPM_CORR_NO_PIPE in the header
    const bool powered = use_pow && (alpha0 != 1.0 || alpha1 != 1.0);
    const bool pipe_off = (flags & DBSPM_CORR_NO_PIPE) || g || mp || !axis2s_ok(pl->host.dev);
    // powered pass: pipelined on 8-line tiles where those apply (slowest axis of a large grid: 3.91 -> 3.61 ms at n = 1024);
    // pipelined on 4-line tiles it measured slower than one-shot (4.28 against 3.90 ms), so there it stays one-shot
    const bool pow_narrow = powered && (flags & DBSPM_CORR_POW_NARROW);
    if (P.v2) {
        axis2_config(n, inner, P.plan.nstages, &P.tshift, &P.tw_smem);
        smem = axis2_smem_bytes(n, P.tshift, P.plan.nstages, P.tw_smem);
        if (smem > 227 * 1024) P.v2 = 0;
#if !defined(AXIS2_NO_WIDE)
        // slowest axis (outer == 1: rows inner*16 bytes apart) and a tile budget that only allows 4-line tiles:
        // measured on B200 at n = 1024, 2 CTAs x 256 threads x T=4 -> 1 CTA x 512 threads x T=8: 2.60 -> 2.32 ms
        // (the same change slows the y pass, whose rows are 2 KB apart, 2.20 -> 2.36 ms)
        // With the power fused into the load the one-shot pass is FP64-issue bound and two CTAs hide it better
        // (4.21 ms for 2 x 256 against 4.87 ms for 1 x 512); pipelined, the 8-line tiles win (3.61 against 3.91 ms).
        else if (!mp && outer == 1 && P.tshift == 2 && inner >= 65536 && P.plan.nstages > 1 &&
                 (!powered || (!pipe_off && !pow_narrow) || (flags & DBSPM_CORR_POW_WIDE)) &&
                 axis2_smem_bytes(n, 3, P.plan.nstages, 0) <= 200 * 1024) {
            wide = true;
            P.tshift = 3; P.tw_smem = 0;
            smem = axis2_smem_bytes(n, 3, P.plan.nstages, 0);
        }
#endif
        // The pipelined pass takes 8-line tiles for the other axes as well: full 128-byte row segments halve the L1
        // wavefronts per byte (the one-shot y pass sat at 72 % L1/LSU), and the pipelining covers the latency that one
        // resident CTA exposes (one-shot, the same tiles were slower: 2.36 against 2.20 ms).  B200, n = 1024, inner = 128:
        // y pass 2.29 -> 2.08 ms.  Not for ragged or short rows (inverse y pass on the window, inner = 27: 0.301 -> 0.313 ms).
        if (P.v2 && !wide && !pipe_off && !(flags & DBSPM_CORR_PIPE_NARROW_Y) && outer > 1 && P.tshift == 2 && inner >= 64 &&
            inner % 8 == 0 && P.plan.nstages > 1 && axis2_smem_bytes(n, 3, P.plan.nstages, 0) <= 160 * 1024) {
            wide = true;
            P.tshift = 3; P.tw_smem = 0;
            smem = axis2_smem_bytes(n, 3, P.plan.nstages, 0);
        }
        // fused exchange: NVLink moves 128-byte segments much better than 64-byte ones, so the peer-store pass takes
        // 8-line tiles whenever they fit (measured in profiles/README.md)
        if (P.v2 && mp && mp->npeers > 0 && !(flags & DBSPM_CORR_PEER_NARROW) && P.tshift < 3 && inner >= 8 && P.plan.nstages > 1 &&
            axis2_smem_bytes(n, 3, P.plan.nstages, 0) <= 200 * 1024) {
            wide = true;
            P.tshift = 3; P.tw_smem = 0;
            smem = axis2_smem_bytes(n, 3, P.plan.nstages, 0);
        }
    }
    DBSPM_REQUIRE(P.v2 || !mp, DBSPM_EUNSUPPORTED, "row-mapped pass: axis length %d does not fit in shared memory", n);
    if (!P.v2) {
        P.tshift = pick_tshift(n, inner);
        DBSPM_REQUIRE(P.tshift >= 0, DBSPM_EUNSUPPORTED, "axis length %d does not fit in shared memory", n);
        smem = axis_smem_bytes(n, P.tshift);
    }
    // powered first pass: the table of pow_fast is staged behind the tile (DBSPM_CORR_SLOW_POW: libdevice exp(alpha log v))
    P.pow_tab_off = -1;
    if (P.v2 && use_pow && (alpha0 != 1.0 || alpha1 != 1.0) && !(flags & DBSPM_CORR_SLOW_POW)) {
        pow_fast_bounds(alpha0, &P.pow_lo[0], &P.pow_hi[0]);
        pow_fast_bounds(alpha1, &P.pow_lo[1], &P.pow_hi[1]);
        P.pow_tab_off = (int)align_up(smem, 16);
        smem = (size_t)P.pow_tab_off + DBSPM_POW_TABLE_LEN * POW_TAB_COPIES * sizeof(double);   // 12 KB: 16 bank-disjoint copies
    }
    P.tiles = (inner + (1LL << P.tshift) - 1) >> P.tshift;
    const long long nblocks = outer * P.tiles;
    // L2 prefetch distance = the blocks resident at once (148 SMs x CTAs per SM): block b + distance takes the slot of b
    P.pf_ahead = 0;
    // (not for the slowest axis: its rows are megabytes apart and the prefetch measured slower, 2.49 -> 2.75 ms at n = 1024)
    if (P.v2 && outer > 1 && !(flags & DBSPM_CORR_NO_PREFETCH)) {
        const int per_sm = wide ? 1 : (int)((227 * 1024) / (smem + 1024) < AXIS2_MIN_BLOCKS ? (227 * 1024) / (smem + 1024) : AXIS2_MIN_BLOCKS);
        P.pf_ahead = 148 * (per_sm < 1 ? 1 : per_sm);
    }
    DBSPM_REQUIRE(nblocks > 0 && nblocks < 2147483647LL, DBSPM_EUNSUPPORTED, "grid too large (%lld blocks)", nblocks);
    if (P.v2 && mp && mp->npeers > 0 && wide) {
        DBSPM_CUDA_TRY(cudaFuncSetAttribute(axis_pass2pw_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        axis_pass2pw_kernel<D><<<dim3((unsigned)nblocks, (unsigned)narrays), 512, smem, st>>>(P);
    } else if (P.v2 && mp && mp->npeers > 0) {
        DBSPM_CUDA_TRY(cudaFuncSetAttribute(axis_pass2p_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        axis_pass2p_kernel<D><<<dim3((unsigned)nblocks, (unsigned)narrays), AXIS2_THREADS, smem, st>>>(P);
    } else if (P.v2 && mp) {
        DBSPM_CUDA_TRY(cudaFuncSetAttribute(axis_pass2m_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        axis_pass2m_kernel<D><<<dim3((unsigned)nblocks, (unsigned)narrays), AXIS2_THREADS, smem, st>>>(P);
    } else if (P.v2 && !pipe_off && (outer > 1 || wide) && (!powered || wide) && nblocks * narrays >= 4LL * 148 * (wide ? 1 : AXIS2_MIN_BLOCKS) &&
               ((long long)(n / P.plan.radix[0]) << P.tshift) == (wide ? 512 : AXIS2_THREADS)) {
        // exactly one first-stage item per thread: its rows travel to the next tile's first stage through cp.async slots + registers
        const int per_sm = wide ? 1 : AXIS2_MIN_BLOCKS;
        P.stage_off = (int)align_up(smem, 16);
        smem = (size_t)P.stage_off + 8 * (size_t)(wide ? 512 : AXIS2_THREADS) * sizeof(cplx);
        const unsigned gx = (unsigned)((148 * per_sm + narrays - 1) / narrays);
        // (with the power: the forward instantiation only -- inverse passes have no power)
        if (powered && D < 0) {
            if (axis2s_plan_is<16, 8, 8>(P.plan)) rc = launch_axis_pipelined<-1, 16, 8, 8, true>(P, gx, narrays, wide, smem, st);
            else rc = launch_axis_pipelined<-1, 16, 8, 4, true>(P, gx, narrays, wide, smem, st);
        } else if (axis2s_plan_is<16, 8, 8>(P.plan)) rc = launch_axis_pipelined<D, 16, 8, 8, false>(P, gx, narrays, wide, smem, st);
        else rc = launch_axis_pipelined<D, 16, 8, 4, false>(P, gx, narrays, wide, smem, st);
        if (rc) return rc;
    } else if (P.v2 && wide) {
        DBSPM_CUDA_TRY(cudaFuncSetAttribute(axis_pass2w_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        axis_pass2w_kernel<D><<<dim3((unsigned)nblocks, (unsigned)narrays), 512, smem, st>>>(P);
    } else if (P.v2) {
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
    if (nx < 1 || ny < 1 || nz < 1 || z_lo < 0 || z_hi <= z_lo || z_hi > nz) return 0;
    const size_t N = (size_t)nx * ny * nz;
    const int nzw = z_hi - z_lo, nwh = (nzw + 1) / 2;
    if (nz & 1) return 2 * align_up(N * 16, 256);         // odd nz: two full complex arrays
    size_t bytes = 2 * align_up(N * 8, 256);
    if (nzw & 1) bytes += align_up((size_t)nx * ny * nwh * 16, 256);
    return bytes;
}

// odd nz: real -> complex, three forward passes on both arrays, product, three inverse passes, real window
static int corr3d_complex(const double *A, const double *B, double alphaA, double alphaB, double dV,
                          int nx, int ny, int nz, int z_lo, int z_hi, double *E_win, void *workspace, int flags, cudaStream_t st)
{
    const size_t N = (size_t)nx * ny * nz;
    cplx *WA = (cplx *)workspace;
    cplx *WB = (cplx *)((unsigned char *)workspace + align_up(N * 16, 256));
    size_t blocks = (N + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    r2c_kernel<<<(unsigned)blocks, 256, 0, st>>>(A, WA, N, alphaA);
    r2c_kernel<<<(unsigned)blocks, 256, 0, st>>>(B, WB, N, alphaB);
    DBSPM_CUDA_TRY(cudaGetLastError());
    const int f = (flags | DBSPM_CORR_NO_TMA) & ~DBSPM_CORR_USE_TMA;
    int rc;
    if ((rc = launch_axis<-1>(WA, WA, WB, WB, 2, 1.0, 1.0, 0, nx, (long long)ny * nz, 1, st, f))) return rc;
    if ((rc = launch_axis<-1>(WA, WA, WB, WB, 2, 1.0, 1.0, 0, ny, nz, nx, st, f))) return rc;
    if ((rc = launch_axis<-1>(WA, WA, WB, WB, 2, 1.0, 1.0, 0, nz, 1, (long long)nx * ny, st, f))) return rc;
    cmulconj_kernel<<<(unsigned)blocks, 256, 0, st>>>(WA, WB, N, dV / ((double)nx * (double)ny * (double)nz));
    DBSPM_CUDA_TRY(cudaGetLastError());
    if ((rc = launch_axis<1>(WA, WA, nullptr, nullptr, 1, 1.0, 1.0, 0, nz, 1, (long long)nx * ny, st, f))) return rc;
    if ((rc = launch_axis<1>(WA, WA, nullptr, nullptr, 1, 1.0, 1.0, 0, ny, nz, nx, st, f))) return rc;
    if ((rc = launch_axis<1>(WA, WA, nullptr, nullptr, 1, 1.0, 1.0, 0, nx, (long long)ny * nz, 1, st, f))) return rc;
    const long long ncol = (long long)nx * ny;
    c2r_window_kernel<<<(unsigned)blocks, 256, 0, st>>>(WA, E_win, ncol, nz, z_lo, z_hi - z_lo);
    DBSPM_CUDA_TRY(cudaGetLastError());
    return DBSPM_OK;
}

// fused z kernel on (nx, ny_rows, nz/2) spectra -> packed window (nx, ny_rows, nwh).  slot_mode: the rows are
// slots [slot_base, slot_base + ny_rows) of a distributed y axis (corr_body.h, dist_slot_of)
static int launch_zfused(const cplx *WA, const cplx *WB, cplx *W, int nx, int ny_rows, int nz, int z_lo, int nwh,
                         double scale, int slot_mode, int slot_base, int flags, cudaStream_t st)
{
    const int n = nz / 2;
    DevPlan *pl = nullptr, *plN = nullptr;
    int rc = get_plan(n, st, &pl, (flags & DBSPM_CORR_Z_RADIX8) ? 8 : ZF2_MAX_RADIX);
    if (rc) return rc;
    rc = get_plan(nz, st, &plN);
    if (rc) return rc;
    ZFusedParams P;
    P.ZA = WA; P.ZB = WB; P.W = W;
    P.nx = nx; P.ny = ny_rows; P.n = n; P.nwh = nwh; P.z_lo = z_lo;
    P.scale = scale;
    P.plan = pl->host.dev; P.tw = pl->tw; P.pos_of = pl->pos_of; P.freq_of = pl->freq_of; P.twN = plN->tw;
    P.npairs = (long long)(nx / 2 + 1) * ny_rows;
    P.slot_mode = slot_mode; P.slot_base = slot_base;
    P.prefetch = (flags & DBSPM_CORR_NO_PREFETCH) ? 0 : 1;
    const bool v2 = axis_v2_ok(P.plan) && zfused2_smem_bytes<ZF2_PAIRS>(n) <= 113 * 1024 &&
                    !(flags & DBSPM_CORR_AXIS_V1);
    if (v2) {
        const long long ngroups = (P.npairs + ZF2_PAIRS - 1) / ZF2_PAIRS;
        const size_t smem = zfused2_smem_bytes<ZF2_PAIRS>(n);
        // persistent blocks: a few per SM (148 SMs), each walking over its share of the pair groups
        int per_sm = (int)((228 * 1024) / (smem + 1024));          // 228 KB per SM, 1 KB reserved per resident block
        per_sm = per_sm < 1 ? 1 : (per_sm > ZF2_MAX_PER_SM ? ZF2_MAX_PER_SM : per_sm);
        long long nblocks = 148LL * per_sm * ZF2_WAVES;
        if (nblocks > ngroups) nblocks = ngroups;
        int maxr = 0;
        for (int i = 0; i < P.plan.nstages; ++i) maxr = P.plan.radix[i] > maxr ? P.plan.radix[i] : maxr;
        auto kern = maxr <= 8 ? zfused2_kernel<8> : zfused2_kernel<16>;
        DBSPM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        DBSPM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        kern<<<(unsigned)nblocks, ZF2_THREADS, smem, st>>>(P);
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
    return DBSPM_OK;
}

// the five launches on a problem of z length nz (the caller's, or the pruned length when g != null)
static int corr3d_core(const double *A, const double *B, double alphaA, double alphaB, double dV,
                       int nx, int ny, int nz, int z_lo, int z_hi, double *E_win,
                       void *workspace, int flags, cudaStream_t st, const GatherSpec *g)
{
    const int n = nz / 2, nzw = z_hi - z_lo, nwh = (nzw + 1) / 2;
    const size_t N = (size_t)nx * ny * nz;
    unsigned char *ws = (unsigned char *)workspace;
    cplx *WA = (cplx *)ws;
    cplx *WB = (cplx *)(ws + align_up(N * 8, 256));
    cplx *W = (nzw & 1) ? (cplx *)(ws + 2 * align_up(N * 8, 256)) : (cplx *)E_win;

    int rc;
    // forward x (out of place, power fused into the load), then y in place
    if (!(flags & DBSPM_CORR_SKIP_XFWD)) {
        rc = launch_axis<-1>((const cplx *)A, WA, (const cplx *)B, WB, 2, alphaA, alphaB, 1, nx, (long long)ny * n, 1, st, flags, g);
        if (rc) return rc;
    }
    if (!(flags & DBSPM_CORR_SKIP_YFWD)) {
        rc = launch_axis<-1>(WA, WA, WB, WB, 2, 1.0, 1.0, 0, ny, n, nx, st, flags);
        if (rc) return rc;
    }

    // fused z: forward, untangle, product, window phase, pruned inverse
    if (!(flags & DBSPM_CORR_SKIP_Z)) {
        rc = launch_zfused(WA, WB, W, nx, ny, nz, z_lo, nwh, dV / ((double)nx * (double)ny * (double)nz) * 0.25, 0, 0, flags, st);
        if (rc) return rc;
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

extern "C" int dbspm_corr3d_f64(const double *A, const double *B, double alphaA, double alphaB, double dV,
                                int nx, int ny, int nz, int z_lo, int z_hi, double *E_win,
                                void *workspace, size_t workspace_bytes, int flags, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    DBSPM_REQUIRE(A && B && E_win && workspace, DBSPM_EINVAL, "null pointer argument");
    DBSPM_REQUIRE(nx >= 1 && ny >= 1 && nz >= 1, DBSPM_EINVAL, "bad shape (%d,%d,%d)", nx, ny, nz);
    DBSPM_REQUIRE(z_lo >= 0 && z_hi > z_lo && z_hi <= nz, DBSPM_EINVAL, "bad window [%d,%d) for nz=%d", z_lo, z_hi, nz);
    DBSPM_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0 && ((uintptr_t)E_win & 15) == 0,
                  DBSPM_EINVAL, "A, B and E_win must be 16-byte aligned");
    const size_t need = dbspm_corr3d_workspace_bytes(nx, ny, nz, z_lo, z_hi, flags);
    DBSPM_REQUIRE(workspace_bytes >= need && ((uintptr_t)workspace & 255) == 0, DBSPM_EWORKSPACE,
                  "workspace needs %zu bytes, 256-byte aligned (got %zu)", need, workspace_bytes);
    if (nz & 1)      // the packed real transform needs an even z length: plain complex transforms instead
        return corr3d_complex(A, B, alphaA, alphaB, dV, nx, ny, nz, z_lo, z_hi, E_win, workspace, flags, st);

    return corr3d_core(A, B, alphaA, alphaB, dV, nx, ny, nz, z_lo, z_hi, E_win, workspace, flags, st, nullptr);
}

// ---- z-pruned correlation for a tip that is exactly zero outside s_len planes -----------------
extern "C" size_t dbspm_corr3d_pruned_workspace_bytes(int nx, int ny, int nz, int z_lo, int z_hi, int s_lo, int s_len, int flags)
{
    if (nx < 1 || ny < 1 || nz < 1 || z_lo < 0 || z_hi <= z_lo || z_hi > nz) return 0;
    const PrunedGeom G = pruned_geometry(nz, z_lo, z_hi, s_lo, s_len);
    if (!G.use) return dbspm_corr3d_workspace_bytes(nx, ny, nz, z_lo, z_hi, flags);
    return dbspm_corr3d_workspace_bytes(nx, ny, G.Mp, G.w_lo, G.w_lo + (z_hi - z_lo), flags);
}

extern "C" int dbspm_corr3d_pruned_f64(const double *A, const double *B, double alphaA, double alphaB, double dV,
                                       int nx, int ny, int nz, int z_lo, int z_hi, int s_lo, int s_len,
                                       double *E_win, void *workspace, size_t workspace_bytes, int flags, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    DBSPM_REQUIRE(A && B && E_win && workspace, DBSPM_EINVAL, "null pointer argument");
    DBSPM_REQUIRE(nx >= 1 && ny >= 1 && nz >= 1, DBSPM_EINVAL, "bad shape (%d,%d,%d)", nx, ny, nz);
    DBSPM_REQUIRE(z_lo >= 0 && z_hi > z_lo && z_hi <= nz, DBSPM_EINVAL, "bad window [%d,%d) for nz=%d", z_lo, z_hi, nz);
    DBSPM_REQUIRE(s_len >= 1 && s_len <= nz, DBSPM_EINVAL, "bad tip support length %d for nz=%d", s_len, nz);
    DBSPM_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0 && ((uintptr_t)E_win & 15) == 0,
                  DBSPM_EINVAL, "A, B and E_win must be 16-byte aligned");
    const size_t need = dbspm_corr3d_pruned_workspace_bytes(nx, ny, nz, z_lo, z_hi, s_lo, s_len, flags);
    DBSPM_REQUIRE(workspace_bytes >= need && ((uintptr_t)workspace & 255) == 0, DBSPM_EWORKSPACE,
                  "workspace needs %zu bytes, 256-byte aligned (got %zu)", need, workspace_bytes);
    PrunedGeom G = pruned_geometry(nz, z_lo, z_hi, s_lo, s_len);
    if (G.use) {
        // the remapped first pass exists in the register-staged body only (all radices of nx fast)
        DevPlan *pl = nullptr;
        int rc = get_plan(nx, st, &pl);
        if (rc) return rc;
        if (!axis_v2_ok(pl->host.dev)) G.use = 0;
    }
    if (!G.use) {
        DBSPM_REQUIRE(workspace_bytes >= dbspm_corr3d_workspace_bytes(nx, ny, nz, z_lo, z_hi, flags), DBSPM_EWORKSPACE,
                      "workspace too small for the full-length path");
        if (nz & 1)
            return corr3d_complex(A, B, alphaA, alphaB, dV, nx, ny, nz, z_lo, z_hi, E_win, workspace, flags, st);
        return corr3d_core(A, B, alphaA, alphaB, dV, nx, ny, nz, z_lo, z_hi, E_win, workspace, flags, st, nullptr);
    }
    GatherSpec g;
    g.src_nz = nz; g.dst_nz = G.Mp;
    g.z0[0] = G.zA0; g.z0[1] = G.zB0; g.valid[0] = G.validA; g.valid[1] = G.validB;
    return corr3d_core(A, B, alphaA, alphaB, dV, nx, ny, G.Mp, G.w_lo, G.w_lo + (z_hi - z_lo), E_win, workspace, flags, st, &g);
}

// ---- distributed correlation: x-slab decomposition over the G ranks of a group -----------------------------
// Rank g of G holds the x-slab [g*nx/G, (g+1)*nx/G) of A and B.  (1) fwd: y transform of the slab (power fused),
// written destination-major: block h of the send buffer = the ny/G frequencies (slot order) rank h owns.
// (2) the caller exchanges the blocks (all-to-all: the one data-path collective of the sr/es step).  (3) mid: x
// transform of the received (nx, ny/G, nz/2) spectra in place, fused z kernel, inverse x transform of this rank's
// packed window rows.  (4) the caller gathers the G packed windows.  (5) fin: inverse y transform reading its rows
// through the slot table, into the real window.  Every output plane needs every input plane, so nothing else of
// the transform can be split without an exchange; this order needs one all-to-all of 16N/G bytes per rank.
extern "C" int dbspm_corr3d_dist_supported(int nx, int ny, int nz, int G)
{
    if (!dist_shape_ok(nx, ny, nz, G)) return 0;
    HostPlan hx, hy, hz;
    if (!dbspm_make_plan(nx, hx) || !dbspm_make_plan(ny, hy) || !dbspm_make_plan(nz / 2, hz, ZF2_MAX_RADIX)) return 0;
    if (!axis_v2_ok(hx.dev) || !axis_v2_ok(hy.dev)) return 0;
    int ts, tws;
    axis2_config(ny, nz / 2, hy.dev.nstages, &ts, &tws);
    if (axis2_smem_bytes(ny, ts, hy.dev.nstages, tws) > 227 * 1024) return 0;
    axis2_config(nx, (long long)(ny / G) * (nz / 2), hx.dev.nstages, &ts, &tws);      // row-mapped inverse x pass of mid_peer
    if (axis2_smem_bytes(nx, ts, hx.dev.nstages, tws) > 227 * 1024) return 0;
    if (zfused_smem_bytes(nz / 2) > 227 * 1024) return 0;
    return 1;
}

// geometry of the z-pruned problem (compact tip), for the callers of the *_cut_f64 entries:
// out = {use, Mp, zA0, zB0, validA, validB, w_lo} (corr_body.h, pruned_geometry)
extern "C" int dbspm_corr3d_pruned_geometry(int nz, int z_lo, int z_hi, int s_lo, int s_len, int *out)
{
    DBSPM_REQUIRE(out, DBSPM_EINVAL, "null pointer argument");
    DBSPM_REQUIRE(nz >= 1 && z_lo >= 0 && z_hi > z_lo && z_hi <= nz && s_len >= 1 && s_len <= nz, DBSPM_EINVAL,
                  "bad window [%d,%d) / support length %d for nz=%d", z_lo, z_hi, s_len, nz);
    const PrunedGeom G = pruned_geometry(nz, z_lo, z_hi, s_lo, s_len);
    out[0] = G.use; out[1] = G.Mp; out[2] = G.zA0; out[3] = G.zB0; out[4] = G.validA; out[5] = G.validB; out[6] = G.w_lo;
    return DBSPM_OK;
}

// cut (nullable): the forward y pass reads the z-ranges of the pruned problem in place (as the pruned x pass of
// dbspm_corr3d_pruned_f64 does) and writes spectra of z length Mp = cut[1]; everything downstream runs with nz = Mp
static int dist_cut_spec(const int *cut, int nz_src, GatherSpec *g, int *nz_problem)
{
    *nz_problem = nz_src;
    if (!cut) return 0;
    DBSPM_REQUIRE(cut[0] == 1 && cut[1] >= 2 && (cut[1] & 1) == 0 && cut[1] <= nz_src && (nz_src & 1) == 0 &&
                  cut[2] >= 0 && cut[3] >= 0 && cut[4] >= 0 && cut[5] >= 0 && cut[4] <= cut[1] && cut[5] <= cut[1] &&
                  !((cut[2] | cut[3] | cut[4] | cut[5]) & 1), DBSPM_EINVAL, "bad z-range cut");
    g->src_nz = nz_src; g->dst_nz = cut[1];
    g->z0[0] = cut[2]; g->z0[1] = cut[3]; g->valid[0] = cut[4]; g->valid[1] = cut[5];
    *nz_problem = cut[1];
    return 1;
}

extern "C" int dbspm_corr3d_dist_fwd_cut_f64(const double *A_slab, const double *B_slab, double alphaA, double alphaB,
                                             int nxl, int ny, int nz_src, int G, const int *cut, void *sendA, void *sendB,
                                             int flags, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    DBSPM_REQUIRE(A_slab && B_slab && sendA && sendB, DBSPM_EINVAL, "null pointer argument");
    GatherSpec gs;
    int nz = nz_src;
    const int has_cut = dist_cut_spec(cut, nz_src, &gs, &nz);
    if (has_cut < 0) return has_cut;
    DBSPM_REQUIRE(dist_shape_ok(nxl * G, ny, nz, G), DBSPM_EUNSUPPORTED,
                  "distributed correlation needs even nz, ny %% (2G) == 0 (got slab %d x %d x %d, G=%d)", nxl, ny, nz, G);
    DBSPM_REQUIRE(((uintptr_t)A_slab & 15) == 0 && ((uintptr_t)B_slab & 15) == 0 && ((uintptr_t)sendA & 15) == 0 &&
                  ((uintptr_t)sendB & 15) == 0, DBSPM_EINVAL, "buffers must be 16-byte aligned");
    const int n = nz / 2, nyl = ny / G;
    const int *omap = nullptr;
    int rc = get_rowmap(ny, G, (long long)nxl * nyl, st, &omap);
    if (rc) return rc;
    AxisMap mp;
    mp.in_map = nullptr; mp.out_map = omap; mp.in_ostride = 0; mp.out_ostride = (long long)nyl * n;
    mp.npeers = 0; mp.out_base_off = 0;
    return launch_axis<-1>((const cplx *)A_slab, (cplx *)sendA, (const cplx *)B_slab, (cplx *)sendB, 2, alphaA, alphaB, 1,
                           ny, n, nxl, st, flags, has_cut ? &gs : nullptr, &mp);
}

extern "C" int dbspm_corr3d_dist_fwd_f64(const double *A_slab, const double *B_slab, double alphaA, double alphaB,
                                         int nxl, int ny, int nz, int G, void *sendA, void *sendB, int flags, void *stream)
{
    return dbspm_corr3d_dist_fwd_cut_f64(A_slab, B_slab, alphaA, alphaB, nxl, ny, nz, G, nullptr, sendA, sendB, flags, stream);
}

// fused exchange: the y pass of rank `me` stores every row straight into its owner's receive buffer
// peer_recv[h] = rank h's (2, nx, ny/G, nz/2) complex buffer (A then B), mapped into this process (symmetric memory)
extern "C" int dbspm_corr3d_dist_fwd_peer_cut_f64(const double *A_slab, const double *B_slab, double alphaA, double alphaB,
                                                  int nxl, int ny, int nz_src, int G, int me, const int *cut,
                                                  void *const *peer_recv, int flags, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    DBSPM_REQUIRE(A_slab && B_slab && peer_recv, DBSPM_EINVAL, "null pointer argument");
    GatherSpec gs;
    int nz = nz_src;
    const int has_cut = dist_cut_spec(cut, nz_src, &gs, &nz);
    if (has_cut < 0) return has_cut;
    DBSPM_REQUIRE(dist_shape_ok(nxl * G, ny, nz, G) && G <= DBSPM_MAX_PEERS && me >= 0 && me < G, DBSPM_EUNSUPPORTED,
                  "fused exchange needs even nz, ny %% (2G) == 0, G <= %d (got slab %d x %d x %d, rank %d of %d)",
                  DBSPM_MAX_PEERS, nxl, ny, nz, me, G);
    const int n = nz / 2, nyl = ny / G, nx = nxl * G;
    DBSPM_REQUIRE(nyl < (1 << DIST_PEER_SHIFT), DBSPM_EUNSUPPORTED, "ny/G too large for the peer table");
    const int *omap = nullptr;
    int rc = get_rowmap(ny, G, -1, st, &omap);
    if (rc) return rc;
    AxisMap mp;
    mp.in_map = nullptr; mp.out_map = omap; mp.in_ostride = 0; mp.out_ostride = (long long)nyl * n;
    mp.npeers = G; mp.out_base_off = (long long)me * nxl * nyl * n;
    for (int h = 0; h < G; ++h) {
        DBSPM_REQUIRE(peer_recv[h] && ((uintptr_t)peer_recv[h] & 15) == 0, DBSPM_EINVAL, "peer buffer %d null or misaligned", h);
        mp.peer[0][h] = (cplx *)peer_recv[h];
        mp.peer[1][h] = (cplx *)peer_recv[h] + (size_t)nx * nyl * n;
    }
    return launch_axis<-1>((const cplx *)A_slab, mp.peer[0][me], (const cplx *)B_slab, mp.peer[1][me], 2, alphaA, alphaB, 1,
                           ny, n, nxl, st, flags, has_cut ? &gs : nullptr, &mp);
}

extern "C" int dbspm_corr3d_dist_fwd_peer_f64(const double *A_slab, const double *B_slab, double alphaA, double alphaB,
                                              int nxl, int ny, int nz, int G, int me, void *const *peer_recv, int flags, void *stream)
{
    return dbspm_corr3d_dist_fwd_peer_cut_f64(A_slab, B_slab, alphaA, alphaB, nxl, ny, nz, G, me, nullptr, peer_recv, flags, stream);
}

extern "C" int dbspm_corr3d_dist_mid_f64(void *recvA, void *recvB, double dV, int nx, int ny, int nz, int G, int h,
                                         int z_lo, int z_hi, void *W_local, int flags, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    DBSPM_REQUIRE(recvA && recvB && W_local, DBSPM_EINVAL, "null pointer argument");
    DBSPM_REQUIRE(dist_shape_ok(nx, ny, nz, G) && h >= 0 && h < G, DBSPM_EUNSUPPORTED,
                  "distributed correlation: bad shape (%d,%d,%d) / rank %d of %d", nx, ny, nz, h, G);
    DBSPM_REQUIRE(z_lo >= 0 && z_hi > z_lo && z_hi <= nz, DBSPM_EINVAL, "bad window [%d,%d) for nz=%d", z_lo, z_hi, nz);
    const int n = nz / 2, nyl = ny / G, nwh = (z_hi - z_lo + 1) / 2;
    cplx *WA = (cplx *)recvA, *WB = (cplx *)recvB, *W = (cplx *)W_local;
    int rc = launch_axis<-1>(WA, WA, WB, WB, 2, 1.0, 1.0, 0, nx, (long long)nyl * n, 1, st, flags);
    if (rc) return rc;
    rc = launch_zfused(WA, WB, W, nx, nyl, nz, z_lo, nwh, dV / ((double)nx * (double)ny * (double)nz) * 0.25, 1, h * nyl, flags, st);
    if (rc) return rc;
    return launch_axis<1>(W, W, nullptr, nullptr, 1, 1.0, 1.0, 0, nx, (long long)nyl * nwh, 1, st, flags);
}

// mid with the second exchange fused into the inverse x pass: row x of this rank's packed window goes straight into
// the buffer of the rank that owns x-slab x / (nx/G): peer_w2[q] = rank q's (G, nx/G, ny/G, ceil(nzw/2)) complex buffer
extern "C" int dbspm_corr3d_dist_mid_peer_f64(void *recvA, void *recvB, double dV, int nx, int ny, int nz, int G, int h,
                                              int z_lo, int z_hi, void *W_local, void *const *peer_w2, int flags, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    DBSPM_REQUIRE(recvA && recvB && W_local && peer_w2, DBSPM_EINVAL, "null pointer argument");
    DBSPM_REQUIRE(dist_shape_ok(nx, ny, nz, G) && G <= DBSPM_MAX_PEERS && h >= 0 && h < G, DBSPM_EUNSUPPORTED,
                  "distributed correlation: bad shape (%d,%d,%d) / rank %d of %d", nx, ny, nz, h, G);
    DBSPM_REQUIRE(z_lo >= 0 && z_hi > z_lo && z_hi <= nz, DBSPM_EINVAL, "bad window [%d,%d) for nz=%d", z_lo, z_hi, nz);
    const int n = nz / 2, nyl = ny / G, nxl = nx / G, nwh = (z_hi - z_lo + 1) / 2;
    DBSPM_REQUIRE(nxl < (1 << DIST_PEER_SHIFT), DBSPM_EUNSUPPORTED, "nx/G too large for the peer table");
    cplx *WA = (cplx *)recvA, *WB = (cplx *)recvB, *W = (cplx *)W_local;
    int rc = launch_axis<-1>(WA, WA, WB, WB, 2, 1.0, 1.0, 0, nx, (long long)nyl * n, 1, st, flags);
    if (rc) return rc;
    rc = launch_zfused(WA, WB, W, nx, nyl, nz, z_lo, nwh, dV / ((double)nx * (double)ny * (double)nz) * 0.25, 1, h * nyl, flags, st);
    if (rc) return rc;
    const int *omap = nullptr;
    rc = get_rowmap(nx, G, -2, st, &omap);
    if (rc) return rc;
    const long long inner = (long long)nyl * nwh;
    AxisMap mp;
    mp.in_map = nullptr; mp.out_map = omap; mp.in_ostride = 0; mp.out_ostride = 0;
    mp.npeers = G; mp.out_base_off = (long long)h * nxl * inner;
    for (int q = 0; q < G; ++q) {
        DBSPM_REQUIRE(peer_w2[q] && ((uintptr_t)peer_w2[q] & 15) == 0, DBSPM_EINVAL, "peer buffer %d null or misaligned", q);
        mp.peer[0][q] = (cplx *)peer_w2[q];
        mp.peer[1][q] = nullptr;
    }
    return launch_axis<1>(W, mp.peer[0][h], nullptr, nullptr, 1, 1.0, 1.0, 0, nx, inner, 1, st, flags, nullptr, &mp);
}

extern "C" size_t dbspm_corr3d_dist_fin_workspace_bytes(int nx, int ny, int z_lo, int z_hi)
{
    const int nzw = z_hi - z_lo;
    if (nx < 1 || ny < 1 || nzw < 1) return 0;
    return (nzw & 1) ? align_up((size_t)nx * ny * ((nzw + 1) / 2) * 16, 256) : 0;
}

extern "C" int dbspm_corr3d_dist_fin_f64(const void *W_all, int nx, int ny, int G, int z_lo, int z_hi, double *E_win,
                                         void *workspace, size_t workspace_bytes, int flags, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    DBSPM_REQUIRE(W_all && E_win, DBSPM_EINVAL, "null pointer argument");
    DBSPM_REQUIRE(G >= 1 && ny % (2 * G) == 0 && z_hi > z_lo, DBSPM_EUNSUPPORTED, "distributed correlation: ny %% (2G) != 0");
    const int nzw = z_hi - z_lo, nwh = (nzw + 1) / 2, nyl = ny / G;
    const size_t need = dbspm_corr3d_dist_fin_workspace_bytes(nx, ny, z_lo, z_hi);
    DBSPM_REQUIRE(need == 0 || (workspace && workspace_bytes >= need && ((uintptr_t)workspace & 255) == 0), DBSPM_EWORKSPACE,
                  "workspace needs %zu bytes, 256-byte aligned (got %zu)", need, workspace_bytes);
    cplx *W = (nzw & 1) ? (cplx *)workspace : (cplx *)E_win;
    const int *imap = nullptr;
    int rc = get_rowmap(ny, G, (long long)nx * nyl, st, &imap);
    if (rc) return rc;
    AxisMap mp;
    mp.in_map = imap; mp.out_map = nullptr; mp.in_ostride = (long long)nyl * nwh; mp.out_ostride = 0;
    mp.npeers = 0; mp.out_base_off = 0;
    rc = launch_axis<1>((const cplx *)W_all, W, nullptr, nullptr, 1, 1.0, 1.0, 0, ny, nwh, nx, st, flags, nullptr, &mp);
    if (rc) return rc;
    if (nzw & 1) {
        const long long ncol = (long long)nx * ny;
        long long blocks = (ncol * nzw + DBSPM_THREADS - 1) / DBSPM_THREADS;
        if (blocks > 148LL * 32) blocks = 148LL * 32;
        compact_kernel<<<(unsigned)blocks, DBSPM_THREADS, 0, st>>>((const double *)W, E_win, ncol, nzw, 2 * nwh);
        DBSPM_CUDA_TRY(cudaGetLastError());
    }
    return DBSPM_OK;
}

extern "C" int dbspm_zsupport_f64(const double *B, int nx, int ny, int nz, int *plane_nonzero, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    DBSPM_REQUIRE(B && plane_nonzero, DBSPM_EINVAL, "null pointer argument");
    DBSPM_REQUIRE(nx >= 1 && ny >= 1 && nz >= 2, DBSPM_EINVAL, "bad shape (%d,%d,%d)", nx, ny, nz);
    DBSPM_REQUIRE((nz & 1) == 0 && nz <= 65536, DBSPM_EUNSUPPORTED, "nz=%d: need an even z length <= 65536", nz);
    DBSPM_REQUIRE(((uintptr_t)B & 15) == 0, DBSPM_EINVAL, "B must be 16-byte aligned");
    DBSPM_CUDA_TRY(cudaMemsetAsync(plane_nonzero, 0, sizeof(int) * (size_t)nz, st));
    const size_t npairs = (size_t)nx * ny * (nz / 2);
    size_t blocks = (npairs + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if ((size_t)nz > 48 * 1024)
        DBSPM_CUDA_TRY(cudaFuncSetAttribute(zsupport_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, nz));
    zsupport_kernel<<<(unsigned)blocks, 256, (size_t)nz, st>>>((const double2 *)B, npairs, nz / 2, plane_nonzero);
    DBSPM_CUDA_TRY(cudaGetLastError());
    return DBSPM_OK;
}


// ---- direct real-space correlation for a tip with a few hundred support points (direct_body.h) ------------------------
__global__ void __launch_bounds__(256) support_points_kernel(const double *B, long long n, int max_points, int *count,
                                                              long long *idx, double *val)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double v = __ldcs(B + i);
        if (v != 0.0) {
            const int pos = atomicAdd(count, 1);
            if (pos < max_points) { idx[pos] = i; val[pos] = v; }
        }
    }
}

__global__ void __launch_bounds__(256) direct_pad_kernel(const __grid_constant__ DirectPadParams Q)
{
    direct_pad_body(Q, (long long)blockIdx.x * blockDim.x + threadIdx.x, (long long)gridDim.x * blockDim.x);
}

// one block = one output tile: a single tensor-map load stages the sample box, then every thread runs its register tile
__global__ void __launch_bounds__(DIRECT_THREADS) direct_kernel(const __grid_constant__ DirectParams Q, const __grid_constant__ CUtensorMap map)
{
#if defined(__CUDA_ARCH__)
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *base = smem + ((128 - ((size_t)smem & 127)) & 127);      // keeps the shared address space (LDS, not generic LD)
    double *box = (double *)base;
    unsigned long long *bar = (unsigned long long *)(base + ((direct_box_bytes(Q.BX, Q.BY, Q.BZ) + 127) & ~(size_t)127));
    const long long tile = blockIdx.x;
    const unsigned bar_s = axis3_smem_u32(bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int tz = (int)(tile % Q.tiles_z);
        const long long r = tile / Q.tiles_z;
        const int ty = (int)(r % Q.tiles_y), tx = (int)(r / Q.tiles_y);
        const unsigned bytes = (unsigned)direct_box_bytes(Q.BX, Q.BY, Q.BZ);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(axis3_smem_u32(box)), "l"(&map), "r"(tz * Q.Tz), "r"(ty * Q.Ty), "r"(tx * Q.Tx), "r"(bar_s) : "memory");
    }
    mbar_wait(bar_s, 0);
    direct_tile_compute(Q, box, tile, (int)threadIdx.x);
#endif
}

extern "C" int dbspm_support_points_f64(const double *B, int nx, int ny, int nz, int max_points, int *count_dev,
                                        long long *idx_dev, double *val_dev, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    DBSPM_REQUIRE(B && count_dev && idx_dev && val_dev, DBSPM_EINVAL, "null pointer argument");
    DBSPM_REQUIRE(nx >= 1 && ny >= 1 && nz >= 1 && max_points >= 1, DBSPM_EINVAL, "bad arguments");
    DBSPM_CUDA_TRY(cudaMemsetAsync(count_dev, 0, sizeof(int), st));
    const long long n = (long long)nx * ny * nz;
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    support_points_kernel<<<(unsigned)blocks, 256, 0, st>>>(B, n, max_points, count_dev, idx_dev, val_dev);
    DBSPM_CUDA_TRY(cudaGetLastError());
    return DBSPM_OK;
}

extern "C" size_t dbspm_corr3d_direct_workspace_bytes(int nx, int ny, int nz, int z_lo, int z_hi, int npts,
                                                      const int *sx, const int *sy, const int *sz)
{
    if (!sx || !sy || !sz) return 0;
    std::vector<double> w((size_t)(npts > 0 ? npts : 0), 1.0);
    const DirectPlan D = direct_make_plan(nx, ny, nz, z_lo, z_hi, npts, sx, sy, sz, w.data(), 1.0);
    return D.ok ? direct_workspace_bytes(D) : 0;
}

// fused multiply-adds per output point the direct kernel would execute (zero padding of the chunks included); 0 = the
// support does not fit a shared-memory box.  The host compares N_w * this with the cost of the spectral path.
extern "C" long long dbspm_corr3d_direct_cost(int nx, int ny, int nz, int z_lo, int z_hi, int npts,
                                              const int *sx, const int *sy, const int *sz)
{
    if (!sx || !sy || !sz) return 0;
    std::vector<double> w((size_t)(npts > 0 ? npts : 0), 1.0);
    const DirectPlan D = direct_make_plan(nx, ny, nz, z_lo, z_hi, npts, sx, sy, sz, w.data(), 1.0);
    return D.ok ? D.fma_per_output : 0;
}

extern "C" int dbspm_corr3d_direct_f64(const double *A, double alphaA, double dV, int nx, int ny, int nz, int z_lo, int z_hi,
                                       int npts, const int *sx, const int *sy, const int *sz, const double *w,
                                       double *E_win, void *workspace, size_t workspace_bytes, int flags, void *stream)
{
    (void)flags;
    cudaStream_t st = (cudaStream_t)stream;
    DBSPM_REQUIRE(A && E_win && workspace && sx && sy && sz && w, DBSPM_EINVAL, "null pointer argument");
    DBSPM_REQUIRE(nx >= 1 && ny >= 1 && nz >= 1, DBSPM_EINVAL, "bad shape (%d,%d,%d)", nx, ny, nz);
    DBSPM_REQUIRE(z_lo >= 0 && z_hi > z_lo && z_hi <= nz, DBSPM_EINVAL, "bad window [%d,%d) for nz=%d", z_lo, z_hi, nz);
    DBSPM_REQUIRE(npts >= 1, DBSPM_EINVAL, "need at least one support point");
    const DirectPlan D = direct_make_plan(nx, ny, nz, z_lo, z_hi, npts, sx, sy, sz, w, dV);
    DBSPM_REQUIRE(D.ok, DBSPM_EUNSUPPORTED, "tip support too wide for the direct path (use dbspm_corr3d_f64)");
    DBSPM_REQUIRE(workspace_bytes >= direct_workspace_bytes(D) && ((uintptr_t)workspace & 255) == 0, DBSPM_EWORKSPACE,
                  "workspace needs %zu bytes, 256-byte aligned (got %zu)", direct_workspace_bytes(D), workspace_bytes);
    TmapEncodeFn enc = tmap_encoder();
    DBSPM_REQUIRE(enc != nullptr, DBSPM_EUNSUPPORTED, "cuTensorMapEncodeTiled unavailable");
    unsigned char *ws = (unsigned char *)workspace;
    double *P = (double *)ws;
    int *off_d = (int *)(ws + direct_align((size_t)D.PX * D.PY * D.PZ * sizeof(double), 256));
    double *w_d = (double *)((unsigned char *)off_d + direct_align(D.chunk_off.size() * sizeof(int), 256));
    // pageable host -> device: the runtime stages the source before returning, the vectors may die with this frame
    DBSPM_CUDA_TRY(cudaMemcpyAsync(off_d, D.chunk_off.data(), D.chunk_off.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    DBSPM_CUDA_TRY(cudaMemcpyAsync(w_d, D.chunk_w.data(), D.chunk_w.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    DirectPadParams Qp;
    Qp.A = A; Qp.P = P; Qp.nx = nx; Qp.ny = ny; Qp.nz = nz; Qp.PX = D.PX; Qp.PY = D.PY; Qp.PZ = D.PZ;
    Qp.x0 = D.start[0]; Qp.y0 = D.start[1]; Qp.z0 = (z_lo + D.start[2]) % nz; Qp.alpha = alphaA;
    const long long np = (long long)D.PX * D.PY * D.PZ;
    long long pb = (np + 255) / 256;
    if (pb > 148LL * 32) pb = 148LL * 32;
    direct_pad_kernel<<<(unsigned)pb, 256, 0, st>>>(Qp);
    DBSPM_CUDA_TRY(cudaGetLastError());
    CUtensorMap map;
    {
        cuuint64_t dims[3] = {(cuuint64_t)D.PZ, (cuuint64_t)D.PY, (cuuint64_t)D.PX};
        cuuint64_t str[2] = {(cuuint64_t)D.PZ * 8, (cuuint64_t)D.PZ * D.PY * 8};
        cuuint32_t box[3] = {(cuuint32_t)D.BZ, (cuuint32_t)D.BY, (cuuint32_t)D.BX};
        cuuint32_t ones[3] = {1, 1, 1};
        DBSPM_REQUIRE(enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void *)P, dims, str, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS,
                      DBSPM_ECUDA, "cuTensorMapEncodeTiled failed for the direct correlation (P %d x %d x %d, box %d x %d x %d)",
                      D.PX, D.PY, D.PZ, D.BX, D.BY, D.BZ);
    }
    DirectParams Q;
    Q.P = P; Q.PX = D.PX; Q.PY = D.PY; Q.PZ = D.PZ; Q.E = E_win; Q.nx = nx; Q.ny = ny; Q.nzw = z_hi - z_lo;
    Q.Tx = D.Tx; Q.Ty = D.Ty; Q.Tz = D.Tz; Q.BX = D.BX; Q.BY = D.BY; Q.BZ = D.BZ;
    Q.tiles_y = D.tiles_y; Q.tiles_z = D.tiles_z; Q.nchunks = (int)D.chunk_off.size();
    Q.chunk_off = off_d; Q.chunk_w = w_d;
    const long long ntiles = (long long)D.tiles_x * D.tiles_y * D.tiles_z;
    DBSPM_REQUIRE(ntiles < 2147483647LL, DBSPM_EUNSUPPORTED, "grid too large");
    const size_t smem = ((direct_box_bytes(D.BX, D.BY, D.BZ) + 127) & ~(size_t)127) + 128 + 16;
    DBSPM_CUDA_TRY(cudaFuncSetAttribute(direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    direct_kernel<<<(unsigned)ntiles, DIRECT_THREADS, smem, st>>>(Q, map);
    DBSPM_CUDA_TRY(cudaGetLastError());
    return DBSPM_OK;
}
