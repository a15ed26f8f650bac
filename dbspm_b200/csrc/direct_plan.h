// direct_plan.h -- host-side geometry of the direct real-space correlation (direct_body.h): circular bounding box of the
// tip's support, chunk tables, tile shape.  Shared by the C ABI (corr3d.cu) and the CPU emulation (tests/emu).
#pragma once
#include <algorithm>
#include <vector>
#include "direct_body.h"

struct DirectPlan {
    int ok = 0;
    int start[3] = {0, 0, 0};     // first plane of the support's circular bounding interval per axis
    int h[3] = {0, 0, 0};         // extent - 1 per axis (offsets run 0 ... h)
    int PX = 0, PY = 0, PZ = 0;   // padded sample
    int Tx = 0, Ty = 0, Tz = 0, BX = 0, BY = 0, BZ = 0;
    int tiles_x = 0, tiles_y = 0, tiles_z = 0;
    std::vector<int> chunk_off;
    std::vector<double> chunk_w;
    long long fma_per_output = 0; // nchunks * LCH (work the kernel does per output, zero padding included)
};

// smallest circular interval of Z_n covering `vals`: (start, extent)
static inline void direct_circular_extent(std::vector<int> vals, int n, int *start, int *extent)
{
    std::sort(vals.begin(), vals.end());
    vals.erase(std::unique(vals.begin(), vals.end()), vals.end());
    int best_gap = -1, best_i = 0;
    for (size_t i = 0; i < vals.size(); ++i) {
        const int nxt = i + 1 < vals.size() ? vals[i + 1] : vals[0] + n;
        const int gap = nxt - vals[i] - 1;
        if (gap > best_gap) { best_gap = gap; best_i = (int)i; }
    }
    *start = vals[(best_i + 1) % vals.size()];
    *extent = n - best_gap;
}

// sx, sy, sz: support points as indices into B (0 <= s < n); w: B[s]^alphaB (dV is applied here)
static inline DirectPlan direct_make_plan(int nx, int ny, int nz, int z_lo, int z_hi, int npts, const int *sx, const int *sy,
                                          const int *sz, const double *w, double dV, size_t max_box_bytes = 220 * 1024)
{
    DirectPlan D;
    if (npts < 1 || nx < 1 || ny < 1 || nz < 1 || z_lo < 0 || z_hi <= z_lo || z_hi > nz) return D;
    const int n[3] = {nx, ny, nz};
    const int *s[3] = {sx, sy, sz};
    int ext[3];
    for (int a = 0; a < 3; ++a) {
        for (int k = 0; k < npts; ++k) if (s[a][k] < 0 || s[a][k] >= n[a]) return D;
        direct_circular_extent(std::vector<int>(s[a], s[a] + npts), n[a], &D.start[a], &ext[a]);
        D.h[a] = ext[a] - 1;
    }
    const int nzw = z_hi - z_lo;
    // support points as offsets inside the bounding box, grouped by (ox, oy) column, sorted along z
    struct Pt { int ox, oy, oz; double w; };
    std::vector<Pt> pts((size_t)npts);
    for (int k = 0; k < npts; ++k)
        pts[k] = {((sx[k] - D.start[0]) % nx + nx) % nx, ((sy[k] - D.start[1]) % ny + ny) % ny, ((sz[k] - D.start[2]) % nz + nz) % nz, w[k] * dV};
    std::sort(pts.begin(), pts.end(), [](const Pt &a, const Pt &b) {
        return a.ox != b.ox ? a.ox < b.ox : (a.oy != b.oy ? a.oy < b.oy : a.oz < b.oz); });
    // tile shape: DIRECT_THREADS threads, each DIRECT_RZ outputs along z; the box with the least staging per output that fits
    double best = 1e300;
    for (int ZQ = 1; ZQ <= 8; ZQ *= 2) {
        for (int Ty = 1; Ty <= DIRECT_THREADS / ZQ; Ty *= 2) {
            const int Tx = DIRECT_THREADS / (ZQ * Ty), Tz = ZQ * DIRECT_RZ;
            const int BX = Tx + D.h[0], BY = Ty + D.h[1];
            int BZ = Tz + D.h[2] + DIRECT_LCH - 1;
            BZ += BZ & 1;                                   // rows of 16-byte multiples for the tensor map
            // lanes of a warp own neighbouring y: their rows start BZ doubles apart, and the tensor-map load writes the box
            // densely, so BZ itself decides the bank pattern.  BZ = 2 (mod 4) spreads eight rows over eight different bank
            // pairs; BZ = 0 (mod 16) would put them all on one (measured: 87 ms instead of 9 for the same work)
            if ((BZ & 3) == 0) BZ += 2;
            if (BX > 256 || BY > 256 || BZ > 256) continue;
            const size_t bytes = direct_box_bytes(BX, BY, BZ);
            if (bytes > max_box_bytes) continue;
            // staged bytes per output, with a penalty for boxes that leave room for only one CTA per SM
            const double cost = (double)bytes / ((double)Tx * Ty * Tz) * (bytes > 110 * 1024 ? 1.5 : 1.0)
                                * (Tz > nzw + DIRECT_RZ ? 2.0 : 1.0);
            if (cost < best) { best = cost; D.Tx = Tx; D.Ty = Ty; D.Tz = Tz; D.BX = BX; D.BY = BY; D.BZ = BZ; }
        }
    }
    if (D.Tx == 0) return D;                                // support too wide for a shared-memory box: spectral path
    // chunks: DIRECT_LCH consecutive z offsets of one (ox, oy) column, zero weights where the support has a hole
    for (size_t i = 0; i < pts.size();) {
        const int ox = pts[i].ox, oy = pts[i].oy, oz0 = pts[i].oz;
        double wc[DIRECT_LCH];
        for (int j = 0; j < DIRECT_LCH; ++j) wc[j] = 0.0;
        while (i < pts.size() && pts[i].ox == ox && pts[i].oy == oy && pts[i].oz < oz0 + DIRECT_LCH) {
            wc[pts[i].oz - oz0] += pts[i].w;               // += : a point listed twice counts twice, as in the sum
            ++i;
        }
        D.chunk_off.push_back((ox * D.BY + oy) * D.BZ + oz0);
        for (int j = 0; j < DIRECT_LCH; ++j) D.chunk_w.push_back(wc[j]);
    }
    D.fma_per_output = (long long)D.chunk_off.size() * DIRECT_LCH;
    D.PX = nx + D.h[0]; D.PY = ny + D.h[1];
    D.PZ = nzw + D.h[2]; D.PZ += D.PZ & 1;
    D.tiles_x = (nx + D.Tx - 1) / D.Tx; D.tiles_y = (ny + D.Ty - 1) / D.Ty; D.tiles_z = (nzw + D.Tz - 1) / D.Tz;
    D.ok = 1;
    return D;
}

static inline size_t direct_align(size_t v, size_t a) { return (v + a - 1) / a * a; }

// workspace: [P padded sample][chunk offsets][chunk weights]
static inline size_t direct_workspace_bytes(const DirectPlan &D)
{
    return direct_align((size_t)D.PX * D.PY * D.PZ * sizeof(double), 256) + direct_align(D.chunk_off.size() * sizeof(int), 256)
           + direct_align(D.chunk_w.size() * sizeof(double), 256);
}
