// hd.h -- host/device portability shims.
//
// Every kernel body in this directory is written as a plain function of
// (params, shared-memory pointer, block index, thread index, thread count).
// nvcc compiles those bodies into the sm_100a kernels that ship; tests/emu/
// compiles the *same* bodies with g++ and runs each block with one "thread"
// (tid = 0, nthr = 1) so the index arithmetic can be checked on a machine
// without a GPU.  The emulation is test infrastructure: the Python package
// never loads it.
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define HD __host__ __device__ __forceinline__
#define HDN __host__ __device__
#else
#define HD inline
#define HDN
#endif

#if defined(__CUDA_ARCH__)
#define BLOCK_SYNC() __syncthreads()
#define WARP_SYNC() __syncwarp()
#define WARP_ALL(p) __all_sync(0xffffffffu, (p))
#define LDG(p) __ldg(p)
#define LD_CG(p) __ldcg(p)          /* L2-coherent load: data another SM may have written during this kernel */
#else
#define BLOCK_SYNC() ((void)0)
#define WARP_SYNC() ((void)0)
#define WARP_ALL(p) (p)
#define LDG(p) (*(p))
#define LD_CG(p) (*(p))
#endif

struct __attribute__((aligned(16))) cplx {
    double x, y;
};

HD cplx cmake(double x, double y) { cplx r; r.x = x; r.y = y; return r; }
HD cplx cadd(cplx a, cplx b) { return cmake(a.x + b.x, a.y + b.y); }
HD cplx csub(cplx a, cplx b) { return cmake(a.x - b.x, a.y - b.y); }
HD cplx cmul(cplx a, cplx b) { return cmake(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
HD cplx cmulc(cplx a, cplx b) { /* a * conj(b) */ return cmake(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }
HD cplx cconj(cplx a) { return cmake(a.x, -a.y); }
HD cplx cscale(cplx a, double s) { return cmake(a.x * s, a.y * s); }
// a * (i*D), D = +1 or -1
HD cplx ldg_c(const cplx *p)
{
#if defined(__CUDA_ARCH__)
    const double2 v = __ldg(reinterpret_cast<const double2 *>(p));
    return cmake(v.x, v.y);
#else
    return *p;
#endif
}
template <int D> HD cplx cmul_iD(cplx a) { return D > 0 ? cmake(-a.y, a.x) : cmake(a.y, -a.x); }
