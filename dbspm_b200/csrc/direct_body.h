// direct_body.h -- direct real-space correlation for a tip whose support is a few hundred points.
//
// Replaces pydbspm/interactions.py:18-22 (with the `**ALPHA` and the window slice of dbspm:329-330 / :348-349) when the
// second array B is nonzero at only K points s_k (a truncated or coarse tip):
//
//     E[R] = dV * sum_r A[r]^a * B[(r - R) mod N]^b = sum_k w_k * P[(R + s_k) mod N],   P = A^a,  w_k = dV * B[s_k]^b
//
// N_w * K fused multiply-adds instead of five passes over both grids: cheaper than the spectral path up to K of a few
// hundred on B200 (FP64 issue is the bound; the crossover is measured in profiles/, the host picks per shape).
//
// Two kernels.  (1) direct_pad_body: P on the z-planes the window can see, with a periodic halo of the support's
// extents on every axis, so that the main kernel needs no wrap-around logic.  (2) direct_tile_body: one block = one tile
// of (Tx, Ty, Tz) outputs; ONE tensor-map load (cp.async.bulk.tensor.3d, completion on an mbarrier) stages the sample box
// the tile can see in shared memory; every thread owns DIRECT_RZ consecutive z outputs of one (x, y) column and walks
// over the support in chunks of DIRECT_LCH z-neighbours of one (s_x, s_y) column: a sliding window of RZ + LCH - 1
// values in registers feeds RZ * LCH fused multiply-adds (weights zero-padded to whole chunks: 0 * P adds nothing).
#pragma once
#include "hd.h"

#ifndef DIRECT_RZ
#define DIRECT_RZ 16          /* z outputs per thread */
#endif
#ifndef DIRECT_LCH
#define DIRECT_LCH 4          /* support points per chunk (consecutive in z) */
#endif
#define DIRECT_THREADS 128

struct DirectPadParams {
    const double *A;          // (nx, ny, nz) z fastest
    double *P;                // (PX, PY, PZ) z fastest
    int nx, ny, nz;
    int PX, PY, PZ;
    int x0, y0, z0;           // P[x', y', z'] = A[(x0 + x') mod nx, (y0 + y') mod ny, (z0 + z') mod nz]^alpha  (x0, y0, z0 in [0, n))
    double alpha;
};

HD double direct_pow(double v, double alpha)
{
    // numpy's `**` for the values a density array holds (corr_body.h pow_signed, restated here to keep this header standalone)
    if (alpha == 1.0) return v;
    if (v < 0.0 && alpha == trunc(alpha) && fabs(alpha) < 9007199254740992.0) {
        const double r = exp(alpha * log(-v));
        return (fmod(alpha, 2.0) != 0.0) ? -r : r;
    }
    return exp(alpha * log(v));
}

// one thread per element of P (grid-stride); z fastest so reads and writes coalesce
HD void direct_pad_body(const DirectPadParams &Q, long long gid, long long gstride)
{
    const long long total = (long long)Q.PX * Q.PY * Q.PZ;
    for (long long i = gid; i < total; i += gstride) {
        const int zp = (int)(i % Q.PZ);
        const long long r = i / Q.PZ;
        const int yp = (int)(r % Q.PY), xp = (int)(r / Q.PY);
        int x = Q.x0 + xp; x = x >= Q.nx ? x % Q.nx : x;
        int y = Q.y0 + yp; y = y >= Q.ny ? y % Q.ny : y;
        int z = Q.z0 + zp; z = z >= Q.nz ? z % Q.nz : z;
        Q.P[i] = direct_pow(Q.A[((size_t)x * Q.ny + y) * Q.nz + z], Q.alpha);
    }
}

struct DirectParams {
    const double *P;          // host emulation only (the device goes through the tensor map)
    int PX, PY, PZ;
    double *E;                // (nx, ny, nzw)
    int nx, ny, nzw;
    int Tx, Ty, Tz;           // outputs per tile; Tx * Ty * (Tz / DIRECT_RZ) == DIRECT_THREADS
    int BX, BY, BZ;           // box the tile sees: Tx + hx, Ty + hy, Tz + hz (+ padding; BZ even)
    int tiles_y, tiles_z;     // tiles along y and z (x is the slowest tile index)
    int nchunks;
    const int *chunk_off;     // per chunk: (ox * BY + oy) * BZ + oz -- offset of its first support point inside the box
    const double *chunk_w;    // nchunks * DIRECT_LCH weights (dV * B^b, zero-padded)
};

HD size_t direct_box_bytes(int BX, int BY, int BZ) { return (size_t)BX * BY * BZ * sizeof(double); }

// box: the staged sample box (BX, BY, BZ), z fastest, already complete (device: the mbarrier has flipped)
HD void direct_tile_compute(const DirectParams &Q, const double *box, long long tile, int tid)
{
    const int ZQ = Q.Tz / DIRECT_RZ;
    const int tz = (int)(tile % Q.tiles_z);
    const long long r = tile / Q.tiles_z;
    const int ty = (int)(r % Q.tiles_y), tx = (int)(r / Q.tiles_y);
    const int zq = tid % ZQ, y = (tid / ZQ) % Q.Ty, x = tid / (ZQ * Q.Ty);
    const int gx = tx * Q.Tx + x, gy = ty * Q.Ty + y, gz = tz * Q.Tz + zq * DIRECT_RZ;
    double acc[DIRECT_RZ];
#pragma unroll
    for (int i = 0; i < DIRECT_RZ; ++i) acc[i] = 0.0;
    const double *mine = box + ((size_t)x * Q.BY + y) * Q.BZ + zq * DIRECT_RZ;
    for (int c = 0; c < Q.nchunks; ++c) {
        const double *src = mine + LDG(Q.chunk_off + c);
        double win[DIRECT_RZ + DIRECT_LCH - 1], w[DIRECT_LCH];
#pragma unroll
        for (int i = 0; i < DIRECT_RZ + DIRECT_LCH - 1; ++i) win[i] = src[i];
#pragma unroll
        for (int j = 0; j < DIRECT_LCH; ++j) w[j] = LDG(Q.chunk_w + (size_t)c * DIRECT_LCH + j);
#pragma unroll
        for (int j = 0; j < DIRECT_LCH; ++j)
#pragma unroll
            for (int i = 0; i < DIRECT_RZ; ++i) acc[i] = fma(w[j], win[i + j], acc[i]);
    }
    if (gx < Q.nx && gy < Q.ny) {
        double *dst = Q.E + ((size_t)gx * Q.ny + gy) * Q.nzw + gz;
#pragma unroll
        for (int i = 0; i < DIRECT_RZ; ++i)
            if (gz + i < Q.nzw) dst[i] = acc[i];
    }
}

// host emulation of the tensor-map load: the box of tile (tx, ty, tz) out of P, zero beyond P's extents (TMA's OOB fill)
HD void direct_tile_stage_host(const DirectParams &Q, double *box, long long tile)
{
    const int tz = (int)(tile % Q.tiles_z);
    const long long r = tile / Q.tiles_z;
    const int ty = (int)(r % Q.tiles_y), tx = (int)(r / Q.tiles_y);
    for (int bx = 0; bx < Q.BX; ++bx)
        for (int by = 0; by < Q.BY; ++by)
            for (int bz = 0; bz < Q.BZ; ++bz) {
                const int px = tx * Q.Tx + bx, py = ty * Q.Ty + by, pz = tz * Q.Tz + bz;
                box[((size_t)bx * Q.BY + by) * Q.BZ + bz] =
                    (px < Q.PX && py < Q.PY && pz < Q.PZ) ? Q.P[((size_t)px * Q.PY + py) * Q.PZ + pz] : 0.0;
            }
}
