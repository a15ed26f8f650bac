// emu.cpp -- CPU emulation of the CUDA kernel bodies (TEST INFRASTRUCTURE ONLY).
//
// Compiles dbspm_b200/csrc/{corr_body.h,relax_body.h,plan.cpp} with g++ and runs each
// CUDA block serially as a single "thread" (tid = 0, nthr = 1; BLOCK_SYNC is a no-op), so
// the index arithmetic, FFT schedule, untangle/product algebra and the Powell state machine
// can be checked against the oracle on a machine without a GPU.  It validates the code that
// ships; it is not a fallback -- dbspm_b200/_lib.py only ever loads libdbspm_b200.so.
#define DBSPM_EMU_INTERP_HOOK 1
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../dbspm_b200/csrc/corr_body.h"
#include "../../dbspm_b200/csrc/relax_body.h"
#include "../../dbspm_b200/csrc/direct_plan.h"
#include "../../dbspm_b200/csrc/plan.h"

double (*dbspm_emu_interp_hook)(double, double, double) = nullptr;
extern "C" void emu_set_interp_hook(double (*f)(double, double, double)) { dbspm_emu_interp_hook = f; }
void (*dbspm_emu_trace_hook)(long long, int, int, double, double) = nullptr;
extern "C" void emu_set_trace_hook(void (*f)(long long, int, int, double, double)) { dbspm_emu_trace_hook = f; }

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static int pick_tshift(int n, long long inner)
{
    for (int ts = 3; ts >= 0; --ts) {
        if ((1LL << ts) > inner && ts > 0) continue;
        if (axis_smem_bytes(n, ts) <= 100 * 1024) return ts;
    }
    return 0;
}

struct EmuGather { int src_nz, dst_nz, z0[2], valid[2]; };
struct EmuMap { const int *in_map, *out_map; long long in_ostride, out_ostride; int npeers; cplx *peer[2][DBSPM_MAX_PEERS]; long long out_base_off; };

template <int D>
static int run_axis(const cplx *in0, cplx *out0, const cplx *in1, cplx *out1, int narrays, double a0, double a1,
                    int use_pow, int n, long long inner, long long outer, int force_tshift, const EmuGather *g = nullptr,
                    const EmuMap *mp = nullptr)
{
    HostPlan hp;
    if (!dbspm_make_plan(n, hp)) return -2;
    AxisParams P;
    P.in[0] = in0; P.in[1] = in1; P.out[0] = out0; P.out[1] = out1;
    P.alpha[0] = a0; P.alpha[1] = a1; P.use_pow = use_pow; P.narrays = narrays;
    P.gather = 0;
    P.in_map = P.out_map = nullptr;
    P.pf_ahead = 0;
    P.pow_tab_off = (force_tshift == 1 || force_tshift >= 1000) ? -1 : 0;     // both power implementations are exercised
    pow_fast_bounds(a0, &P.pow_lo[0], &P.pow_hi[0]);
    pow_fast_bounds(a1, &P.pow_lo[1], &P.pow_hi[1]);
    P.in_ostride = P.out_ostride = (long long)n * inner;
    if (mp) {   // as launch_axis(): row-mapped passes exist in the register-staged body only
        if (!axis_v2_ok(hp.dev)) return -2;
        if (force_tshift >= 100) force_tshift = -1;
        P.in_map = mp->in_map; P.out_map = mp->out_map;
        if (mp->in_ostride) P.in_ostride = mp->in_ostride;
        if (mp->out_ostride) P.out_ostride = mp->out_ostride;
        P.out_base_off = mp->out_base_off;
        for (int w = 0; w < 2; ++w)
            for (int q = 0; q < DBSPM_MAX_PEERS; ++q) P.out_peer[w][q] = q < mp->npeers ? mp->peer[w][q] : nullptr;
    }
    if (g) {   // as launch_axis(): the remapped first pass exists in the register-staged body only
        if (!axis_v2_ok(hp.dev)) return -2;
        if (force_tshift >= 100) force_tshift = -1;
        P.gather = 1;
        P.g_src_nzh = g->src_nz / 2; P.g_dst_nzh = g->dst_nz / 2;
        P.g_src_inner = inner / P.g_dst_nzh * P.g_src_nzh;
        for (int w = 0; w < 2; ++w) { P.g_zp0[w] = g->z0[w] / 2; P.g_valid[w] = g->valid[w] / 2; }
    }
    // 300 = persistent software-pipelined body (axis_pass_body_v2s), auto tile width; 301.. = forced tshift.  Falls back to the
    // one-shot body where the launcher does (plan not eligible, row maps, z-range gather)
    bool pipelined = false;
    if (force_tshift >= 300 && force_tshift < 1000) {
        pipelined = axis2s_ok(hp.dev) && !g && !mp;
        force_tshift = force_tshift > 300 ? force_tshift - 301 : -1;
    }
    if (force_tshift >= 200 && (!axis_v2_ok(hp.dev) || hp.dev.nstages > 3)) force_tshift = -1;   // launcher falls back too
    if (force_tshift >= 200) {
        // TMA-fed persistent body (axis_tma_body) with the two TMA operations emulated by copies;
        // 200 = auto tile width, 201.. = forced tshift; three "persistent blocks" walk the tiles
        AxisTmaParams Q;
        Q.in[0] = in0; Q.in[1] = in1; Q.out[0] = out0; Q.out[1] = out1;
        Q.alpha[0] = a0; Q.alpha[1] = a1; Q.use_pow = use_pow; Q.narrays = narrays;
        Q.n = n; Q.inner = inner; Q.outer = outer;
        Q.tshift = force_tshift > 200 ? force_tshift - 201 : axis3_config(n, inner);
        Q.tiles = (inner + (1LL << Q.tshift) - 1) >> Q.tshift;
        Q.plan = hp.dev; Q.tw = hp.tw.data(); Q.freq_of = hp.freq_of.data();
        // in-place passes (in == out): a real block reads its tile before storing it and tiles are disjoint,
        // so the serial emulation is faithful
        std::vector<unsigned char> sm3(axis3_smem_bytes(n, Q.tshift) + 256);
        for (long long b = 0; b < 3; ++b) axis_tma_body<D>(Q, nullptr, nullptr, sm3.data(), b, 3, 0, 1);
        return 0;
    }
    P.n = n; P.inner = inner; P.outer = outer;
    P.plan = hp.dev; P.tw = hp.tw.data(); P.pos_of = hp.pos_of.data(); P.freq_of = hp.freq_of.data();
    // same selection as launch_axis() in corr3d.cu; force_tshift >= 100 selects the v1 body
    P.v2 = axis_v2_ok(P.plan) && force_tshift < 100;
    P.tw_smem = 0;
    if (force_tshift >= 100) force_tshift -= 101;
    size_t bytes;
    if (P.v2) {
        axis2_config(n, inner, P.plan.nstages, &P.tshift, &P.tw_smem);
        if (force_tshift >= 0) { P.tshift = force_tshift; P.tw_smem = force_tshift & 1; }
        bytes = axis2_smem_bytes(n, P.tshift, P.plan.nstages, P.tw_smem);
    } else {
        P.tshift = force_tshift >= 0 ? force_tshift : pick_tshift(n, inner);
        bytes = axis_smem_bytes(n, P.tshift);
    }
    P.tiles = (inner + (1LL << P.tshift) - 1) >> P.tshift;
    std::vector<unsigned char> smem(bytes + 64);
    unsigned char *sm = (unsigned char *)align_up((size_t)smem.data(), 16);
    if (pipelined && P.v2) {
        P.stage_off = (int)align_up(bytes, 16);
        std::vector<unsigned char> smem2(P.stage_off + 8 * sizeof(cplx) + 64);
        sm = (unsigned char *)align_up((size_t)smem2.data(), 16);
        // three "persistent blocks" walk the tiles with stride 3; one thread each, so a tile has more items than
        // threads and both the carried first item and the directly loaded ones are exercised.  In-place passes: a real
        // block requests tile i + 1 before it stores tile i, and tiles are disjoint, so the serial order is faithful
        for (int w = 0; w < narrays; ++w)
            for (long long b = 0; b < 3; ++b) {
                if (axis2s_plan_is<16, 8, 8>(P.plan)) axis_pass_body_v2s<D, 16, 8, 8, true>(P, sm, b, 3, w, 0, 1);
                else axis_pass_body_v2s<D, 16, 8, 4, true>(P, sm, b, 3, w, 0, 1);
            }
        return 0;
    }
    for (int w = 0; w < narrays; ++w)
        for (long long b = 0; b < outer * P.tiles; ++b) {
            if (P.v2 && mp && mp->npeers > 0) axis_pass_body_v2<D, 2>(P, sm, b, w, 0, 1);
            else if (P.v2 && mp) axis_pass_body_v2<D, 1>(P, sm, b, w, 0, 1);
            else if (P.v2) axis_pass_body_v2<D>(P, sm, b, w, 0, 1);
            else axis_pass_body<D>(P, sm, b, w, 0, 1);
        }
    return 0;
}

extern "C" size_t emu_corr3d_workspace_bytes(int nx, int ny, int nz, int z_lo, int z_hi)
{
    const size_t N = (size_t)nx * ny * nz;
    const int nzw = z_hi - z_lo, nwh = (nzw + 1) / 2;
    size_t bytes = 2 * align_up(N * 8, 256);
    if (nzw & 1) bytes += align_up((size_t)nx * ny * nwh * 16, 256);
    return bytes;
}

static int emu_corr3d_core(const double *A, const double *B, double alphaA, double alphaB, double dV,
                           int nx, int ny, int nz, int z_lo, int z_hi, double *E_win, int force_tshift, const EmuGather *g)
{
    if (nz & 1) {
        // odd nz: corr3d_complex() of corr3d.cu
        const size_t N = (size_t)nx * ny * nz;
        std::vector<cplx> WA(N), WB(N);
        int rc;
        r2c_body(A, WA.data(), N, alphaA, 0, 1);
        r2c_body(B, WB.data(), N, alphaB, 0, 1);
        if (force_tshift >= 200) force_tshift = -1;
        if ((rc = run_axis<-1>(WA.data(), WA.data(), WB.data(), WB.data(), 2, 1.0, 1.0, 0, nx, (long long)ny * nz, 1, force_tshift))) return rc;
        if ((rc = run_axis<-1>(WA.data(), WA.data(), WB.data(), WB.data(), 2, 1.0, 1.0, 0, ny, nz, nx, force_tshift))) return rc;
        if ((rc = run_axis<-1>(WA.data(), WA.data(), WB.data(), WB.data(), 2, 1.0, 1.0, 0, nz, 1, (long long)nx * ny, force_tshift))) return rc;
        cmulconj_body(WA.data(), WB.data(), N, dV / ((double)nx * (double)ny * (double)nz), 0, 1);
        if ((rc = run_axis<1>(WA.data(), WA.data(), nullptr, nullptr, 1, 1.0, 1.0, 0, nz, 1, (long long)nx * ny, force_tshift))) return rc;
        if ((rc = run_axis<1>(WA.data(), WA.data(), nullptr, nullptr, 1, 1.0, 1.0, 0, ny, nz, nx, force_tshift))) return rc;
        if ((rc = run_axis<1>(WA.data(), WA.data(), nullptr, nullptr, 1, 1.0, 1.0, 0, nx, (long long)ny * nz, 1, force_tshift))) return rc;
        c2r_window_body(WA.data(), E_win, (long long)nx * ny, nz, z_lo, z_hi - z_lo, 0, 1);
        return 0;
    }
    const int n = nz / 2, nzw = z_hi - z_lo, nwh = (nzw + 1) / 2;
    const size_t N = (size_t)nx * ny * nz;
    std::vector<cplx> WA(N / 2), WB(N / 2), Wodd;
    cplx *W = (cplx *)E_win;
    if (nzw & 1) { Wodd.resize((size_t)nx * ny * nwh); W = Wodd.data(); }
    int rc;
    if ((rc = run_axis<-1>((const cplx *)A, WA.data(), (const cplx *)B, WB.data(), 2, alphaA, alphaB, 1, nx, (long long)ny * n, 1, force_tshift, g))) return rc;
    if ((rc = run_axis<-1>(WA.data(), WA.data(), WB.data(), WB.data(), 2, 1.0, 1.0, 0, ny, n, nx, force_tshift))) return rc;
    {
        HostPlan hp, hpN;
        if (!dbspm_make_plan(n, hp) || !dbspm_make_plan(nz, hpN)) return -2;
        ZFusedParams P;
        P.ZA = WA.data(); P.ZB = WB.data(); P.W = W;
        P.nx = nx; P.ny = ny; P.n = n; P.nwh = nwh; P.z_lo = z_lo;
        P.scale = dV / ((double)nx * (double)ny * (double)nz) * 0.25;
        P.plan = hp.dev; P.tw = hp.tw.data(); P.pos_of = hp.pos_of.data(); P.freq_of = hp.freq_of.data(); P.twN = hpN.tw.data();
        P.npairs = (long long)(nx / 2 + 1) * ny;
        P.slot_mode = 0; P.slot_base = 0; P.prefetch = 0;
        // same selection as dbspm_corr3d_f64: v2 (4 pairs / block) when the plan is all fast radices
        const bool v2 = axis_v2_ok(P.plan) && (force_tshift < 100 || force_tshift >= 200);
        std::vector<unsigned char> smem((v2 ? zfused2_smem_bytes<4>(n) : zfused_smem_bytes(n)) + 64);
        unsigned char *sm = (unsigned char *)align_up((size_t)smem.data(), 16);
        if (v2) for (long long b = 0; b < 3; ++b) zfused_body_v2<4>(P, sm, b, 3, 0, 1);   // 3 persistent blocks
        else for (long long b = 0; b < (P.npairs + ZL - 1) / ZL; ++b) zfused_body(P, sm, b, 0, 1);
    }
    if ((rc = run_axis<1>(W, W, nullptr, nullptr, 1, 1.0, 1.0, 0, ny, nwh, nx, force_tshift))) return rc;
    if ((rc = run_axis<1>(W, W, nullptr, nullptr, 1, 1.0, 1.0, 0, nx, (long long)ny * nwh, 1, force_tshift))) return rc;
    if (nzw & 1) compact_body((const double *)W, E_win, (long long)nx * ny, nzw, 2 * nwh, 0, 1);
    return 0;
}

extern "C" int emu_corr3d_f64(const double *A, const double *B, double alphaA, double alphaB, double dV,
                              int nx, int ny, int nz, int z_lo, int z_hi, double *E_win, int force_tshift)
{
    return emu_corr3d_core(A, B, alphaA, alphaB, dV, nx, ny, nz, z_lo, z_hi, E_win, force_tshift, nullptr);
}

// what dbspm_corr3d_pruned_f64 does; returns the pruned z length (or nz when the full path was taken), < 0 on error
extern "C" int emu_corr3d_pruned_f64(const double *A, const double *B, double alphaA, double alphaB, double dV,
                                     int nx, int ny, int nz, int z_lo, int z_hi, int s_lo, int s_len, double *E_win, int force_tshift)
{
    PrunedGeom G = pruned_geometry(nz, z_lo, z_hi, s_lo, s_len);
    HostPlan hp;
    if (!dbspm_make_plan(nx, hp)) return -2;
    if (G.use && !axis_v2_ok(hp.dev)) G.use = 0;
    if (!G.use) {
        const int rc = emu_corr3d_core(A, B, alphaA, alphaB, dV, nx, ny, nz, z_lo, z_hi, E_win, force_tshift, nullptr);
        return rc ? rc : nz;
    }
    EmuGather g;
    g.src_nz = nz; g.dst_nz = G.Mp; g.z0[0] = G.zA0; g.z0[1] = G.zB0; g.valid[0] = G.validA; g.valid[1] = G.validB;
    const int rc = emu_corr3d_core(A, B, alphaA, alphaB, dV, nx, ny, G.Mp, G.w_lo, G.w_lo + (z_hi - z_lo), E_win, force_tshift, &g);
    return rc ? rc : G.Mp;
}

// what parallel.corr3d_distributed does with G ranks, run serially: dbspm_corr3d_dist_fwd_f64 per rank, the
// all-to-all as plain block copies, dbspm_corr3d_dist_mid_f64 per rank, the gather, dbspm_corr3d_dist_fin_f64
// cut (nullable): the z-range cut of the pruned problem (compact tip); nz_src = z length of the caller's arrays
static int emu_dist_core(const double *A, const double *B, double alphaA, double alphaB, double dV,
                         int nx, int ny, int nz_src, int nz, int z_lo, int z_hi, double *E_win, int G, int force_tshift,
                         const EmuGather *cut)
{
    const bool peer = force_tshift >= 1000;          // fused exchange: the y pass stores into the owners' buffers
    if (peer) force_tshift -= 1001;
    if (!dist_shape_ok(nx, ny, nz, G) || (peer && G > DBSPM_MAX_PEERS)) return -3;
    const int n = nz / 2, nsrc = nz_src / 2, nxl = nx / G, nyl = ny / G, nzw = z_hi - z_lo, nwh = (nzw + 1) / 2;
    const size_t blk = (size_t)nxl * nyl * n;                     // one (source, destination) block of the exchange
    std::vector<std::vector<cplx>> sendA(G), sendB(G), recvA(G), recvB(G);
    std::vector<int> omap(ny), imap(ny);
    for (int K = 0; K < ny; ++K) { omap[K] = dist_row_of(K, ny, G, (long long)nxl * nyl); imap[K] = dist_row_of(K, ny, G, (long long)nx * nyl); }
    int rc;
    std::vector<int> pmap(ny);
    for (int K = 0; K < ny; ++K) pmap[K] = dist_peer_entry(K, ny, G);
    std::vector<std::vector<cplx>> recvAB(G);                      // peer mode: rank h's (2, nx, nyl, n) receive buffer
    if (peer) for (int h = 0; h < G; ++h) recvAB[h].assign(2 * blk * G, cmake(0, 0));
    for (int g = 0; g < G; ++g) {
        const cplx *As = (const cplx *)A + (size_t)g * nxl * ny * nsrc, *Bs = (const cplx *)B + (size_t)g * nxl * ny * nsrc;
        if (peer) {
            EmuMap mp = {nullptr, pmap.data(), 0, (long long)nyl * n, G, {}, (long long)g * nxl * nyl * n};
            for (int h = 0; h < G; ++h) { mp.peer[0][h] = recvAB[h].data(); mp.peer[1][h] = recvAB[h].data() + blk * G; }
            if ((rc = run_axis<-1>(As, mp.peer[0][g], Bs, mp.peer[1][g], 2, alphaA, alphaB, 1, ny, n, nxl, force_tshift, cut, &mp))) return rc;
            continue;
        }
        sendA[g].assign(blk * G, cmake(0, 0)); sendB[g].assign(blk * G, cmake(0, 0));
        EmuMap mp = {nullptr, omap.data(), 0, (long long)nyl * n, 0, {}, 0};
        if ((rc = run_axis<-1>(As, sendA[g].data(), Bs, sendB[g].data(), 2, alphaA, alphaB, 1, ny, n, nxl, force_tshift, cut, &mp))) return rc;
    }
    for (int h = 0; h < G; ++h) {                                  // all-to-all: block h of rank g -> block g of rank h
        recvA[h].resize(blk * G); recvB[h].resize(blk * G);
        if (peer) {
            memcpy(recvA[h].data(), recvAB[h].data(), blk * G * sizeof(cplx));
            memcpy(recvB[h].data(), recvAB[h].data() + blk * G, blk * G * sizeof(cplx));
            continue;
        }
        for (int g = 0; g < G; ++g) {
            memcpy(recvA[h].data() + blk * g, sendA[g].data() + blk * h, blk * sizeof(cplx));
            memcpy(recvB[h].data() + blk * g, sendB[g].data() + blk * h, blk * sizeof(cplx));
        }
    }
    // packed window rows: W2[q] = rank q's (G, nxl, nyl, nwh) buffer = the rows of x-slab q from every rank
    std::vector<std::vector<cplx>> W2(G);
    for (int q = 0; q < G; ++q) W2[q].assign((size_t)G * nxl * nyl * nwh, cmake(0, 0));
    std::vector<int> bmap(nx);
    for (int X = 0; X < nx; ++X) bmap[X] = dist_block_entry(X, nx, G);
    HostPlan hp, hpN;
    if (!dbspm_make_plan(n, hp) || !dbspm_make_plan(nz, hpN)) return -2;
    const long long winner = (long long)nyl * nwh;
    for (int h = 0; h < G; ++h) {
        cplx *WA = recvA[h].data(), *WB = recvB[h].data();
        std::vector<cplx> Wl((size_t)nx * nyl * nwh);
        cplx *W = Wl.data();
        if ((rc = run_axis<-1>(WA, WA, WB, WB, 2, 1.0, 1.0, 0, nx, (long long)nyl * n, 1, force_tshift))) return rc;
        ZFusedParams P;
        P.ZA = WA; P.ZB = WB; P.W = W;
        P.nx = nx; P.ny = nyl; P.n = n; P.nwh = nwh; P.z_lo = z_lo;
        P.scale = dV / ((double)nx * (double)ny * (double)nz) * 0.25;
        P.plan = hp.dev; P.tw = hp.tw.data(); P.pos_of = hp.pos_of.data(); P.freq_of = hp.freq_of.data(); P.twN = hpN.tw.data();
        P.npairs = (long long)(nx / 2 + 1) * nyl;
        P.slot_mode = 1; P.slot_base = h * nyl; P.prefetch = 0;
        const bool v2 = axis_v2_ok(P.plan) && (force_tshift < 100 || force_tshift >= 200);
        std::vector<unsigned char> smem((v2 ? zfused2_smem_bytes<4>(n) : zfused_smem_bytes(n)) + 64);
        unsigned char *sm = (unsigned char *)align_up((size_t)smem.data(), 16);
        if (v2) for (long long b = 0; b < 3; ++b) zfused_body_v2<4>(P, sm, b, 3, 0, 1);
        else for (long long b = 0; b < (P.npairs + ZL - 1) / ZL; ++b) zfused_body(P, sm, b, 0, 1);
        if (peer) {
            // dbspm_corr3d_dist_mid_peer_f64: the inverse x pass stores row x into the owner of x-slab x / nxl
            EmuMap mp = {nullptr, bmap.data(), 0, 0, G, {}, (long long)h * nxl * winner};
            for (int q = 0; q < G; ++q) { mp.peer[0][q] = W2[q].data(); mp.peer[1][q] = nullptr; }
            if ((rc = run_axis<1>(W, mp.peer[0][h], nullptr, nullptr, 1, 1.0, 1.0, 0, nx, winner, 1, force_tshift, nullptr, &mp))) return rc;
        } else {
            if ((rc = run_axis<1>(W, W, nullptr, nullptr, 1, 1.0, 1.0, 0, nx, winner, 1, force_tshift))) return rc;
            for (int q = 0; q < G; ++q)                            // second all-to-all: x-slab q of rank h -> block h of rank q
                memcpy(W2[q].data() + (size_t)h * nxl * winner, W + (size_t)q * nxl * winner, (size_t)nxl * winner * sizeof(cplx));
        }
    }
    // every rank: inverse y pass of its x-slab through the slot table -> real slab, placed at its x range
    std::vector<int> imap2(ny);
    for (int K = 0; K < ny; ++K) imap2[K] = dist_row_of(K, ny, G, (long long)nxl * nyl);
    for (int q = 0; q < G; ++q) {
        std::vector<cplx> Wodd;
        double *Eq = E_win + (size_t)q * nxl * ny * nzw;
        cplx *Wf = (cplx *)Eq;
        if (nzw & 1) { Wodd.resize((size_t)nxl * ny * nwh); Wf = Wodd.data(); }
        EmuMap mp = {imap2.data(), nullptr, (long long)nyl * nwh, 0, 0, {}, 0};
        if ((rc = run_axis<1>(W2[q].data(), Wf, nullptr, nullptr, 1, 1.0, 1.0, 0, ny, nwh, nxl, force_tshift, nullptr, &mp))) return rc;
        if (nzw & 1) compact_body((const double *)Wf, Eq, (long long)nxl * ny, nzw, 2 * nwh, 0, 1);
    }
    (void)imap;
    return 0;
}

extern "C" int emu_corr3d_dist_f64(const double *A, const double *B, double alphaA, double alphaB, double dV,
                                   int nx, int ny, int nz, int z_lo, int z_hi, double *E_win, int G, int force_tshift)
{
    return emu_dist_core(A, B, alphaA, alphaB, dV, nx, ny, nz, nz, z_lo, z_hi, E_win, G, force_tshift, nullptr);
}

// the distributed step for a tip that is zero outside (s_lo, s_len): returns the z length transformed, < 0 on error
extern "C" int emu_corr3d_dist_pruned_f64(const double *A, const double *B, double alphaA, double alphaB, double dV,
                                          int nx, int ny, int nz, int z_lo, int z_hi, int s_lo, int s_len, double *E_win,
                                          int G, int force_tshift)
{
    const PrunedGeom P = pruned_geometry(nz, z_lo, z_hi, s_lo, s_len);
    if (!P.use) {
        const int rc = emu_dist_core(A, B, alphaA, alphaB, dV, nx, ny, nz, nz, z_lo, z_hi, E_win, G, force_tshift, nullptr);
        return rc ? rc : nz;
    }
    EmuGather g;
    g.src_nz = nz; g.dst_nz = P.Mp; g.z0[0] = P.zA0; g.z0[1] = P.zB0; g.valid[0] = P.validA; g.valid[1] = P.validB;
    const int rc = emu_dist_core(A, B, alphaA, alphaB, dV, nx, ny, nz, P.Mp, P.w_lo, P.w_lo + (z_hi - z_lo), E_win, G, force_tshift, &g);
    return rc ? rc : P.Mp;
}

extern "C" int emu_dist_slot_of(int K, int n) { return dist_slot_of(K, n); }
extern "C" int emu_dist_mirror_slot(int S) { return dist_mirror_slot(S); }
extern "C" int emu_dist_row_of(int K, int n, int G, long long blockrows) { return dist_row_of(K, n, G, blockrows); }

extern "C" void emu_pow_fast(const double *in, double *out, long long n, double alpha)
{
    for (long long i = 0; i < n; ++i) out[i] = pow_fast(in[i], alpha, dbspm_pow_table_h);
}

// plain 1-D FFT of T interleaved lines through the tile routine (plan / butterfly unit test)
extern "C" int emu_fft_lines(const double *in, double *out, int n, int tshift, int dir)
{
    HostPlan hp;
    if (!dbspm_make_plan(n, hp)) return -2;
    const int T = 1 << tshift;
    std::vector<cplx> s((size_t)n * T);
    memcpy(s.data(), in, sizeof(cplx) * s.size());
    if (dir < 0) fft_tile<-1>(s.data(), T, tshift, hp.dev, hp.tw.data(), 0, 1);
    else fft_tile<1>(s.data(), T, tshift, hp.dev, hp.tw.data(), 0, 1);
    cplx *o = (cplx *)out;
    for (int K = 0; K < n; ++K)
        for (int t = 0; t < T; ++t) o[(size_t)K * T + t] = s[(size_t)hp.pos_of[K] * T + t];
    return hp.dev.nstages;
}

// transposed (decimation-in-time) stages: natural input is placed in DIF output order first
extern "C" int emu_fft_lines_t(const double *in, double *out, int n, int tshift, int dir)
{
    HostPlan hp;
    if (!dbspm_make_plan(n, hp) || !axis_v2_ok(hp.dev)) return -2;
    const int T = 1 << tshift;
    std::vector<cplx> s((size_t)n * T);
    const cplx *u = (const cplx *)in;
    for (int p = 0; p < n; ++p)
        for (int t = 0; t < T; ++t) s[(size_t)p * T + t] = u[(size_t)hp.freq_of[p] * T + t];
    if (dir < 0) fft_tile_t<-1>(s.data(), T, tshift, hp.dev, hp.tw.data(), 0, 1);
    else fft_tile_t<1>(s.data(), T, tshift, hp.dev, hp.tw.data(), 0, 1);
    memcpy(out, s.data(), sizeof(cplx) * s.size());
    return hp.dev.nstages;
}

static bool inv3(const double m[9], double out[9])
{
    const double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
    const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
    if (det == 0.0) return false;
    out[0] = (e * i - f * h) / det; out[1] = (c * h - b * i) / det; out[2] = (b * f - c * e) / det;
    out[3] = (f * g - d * i) / det; out[4] = (a * i - c * g) / det; out[5] = (c * d - a * f) / det;
    out[6] = (d * h - e * g) / det; out[7] = (b * g - a * h) / det; out[8] = (a * e - b * d) / det;
    return true;
}

static int emu_relax_impl(const double *static_win, int nx_global, int ny, int nzc, int x_base, int nx_local,
                          int i_lo, int i_hi, int j_lo, int j_hi, int nz_relax,
                          double kappa, double rpivot, const double dr[3], const double *dR,
                          double xtol, double ftol, int maxfev,
                          double *out_etp, double *out_cart, int *out_nfev, int transposed);

extern "C" int emu_relax_powell_f64(const double *static_win, int nx, int ny, int nzc,
                                    int i_lo, int i_hi, int j_lo, int j_hi, int nz_relax,
                                    double kappa, double rpivot, const double dr[3], const double *dR,
                                    double xtol, double ftol, int maxfev,
                                    double *out_etp, double *out_cart, int *out_nfev, int transposed)
{
    return emu_relax_impl(static_win, nx, ny, nzc, 0, 0, i_lo, i_hi, j_lo, j_hi, nz_relax, kappa, rpivot, dr, dR, xtol, ftol, maxfev,
                          out_etp, out_cart, out_nfev, transposed);
}

// x-slab mode: static_slab holds the planes x_base ... x_base + nx_local - 1 (mod nx) of the window
extern "C" int emu_relax_powell_slab_f64(const double *static_slab, int nx, int ny, int nzc, int x_base, int nx_local,
                                         int i_lo, int i_hi, int j_lo, int j_hi, int nz_relax,
                                         double kappa, double rpivot, const double dr[3], const double *dR,
                                         double xtol, double ftol, int maxfev,
                                         double *out_etp, double *out_cart, int *out_nfev, int transposed)
{
    return emu_relax_impl(static_slab, nx, ny, nzc, x_base, nx_local, i_lo, i_hi, j_lo, j_hi, nz_relax, kappa, rpivot, dr, dR, xtol, ftol,
                          maxfev, out_etp, out_cart, out_nfev, transposed);
}

static int emu_relax_impl(const double *static_win, int nx_global, int ny, int nzc, int x_base, int nx_local,
                          int i_lo, int i_hi, int j_lo, int j_hi, int nz_relax,
                          double kappa, double rpivot, const double dr[3], const double *dR,
                          double xtol, double ftol, int maxfev,
                          double *out_etp, double *out_cart, int *out_nfev, int transposed)
{
    const int nx = nx_local ? nx_local : nx_global;          // planes stored (loops of the re-layout below)
    double powell_sm[PS_TOTAL];
    RelaxParams P;
    P.grid = static_win; P.nx = nx_global; P.ny = ny; P.nzc = nzc; P.nyp = ny + 3;
    P.i_lo = i_lo; P.i_hi = i_hi; P.j_lo = j_lo; P.j_hi = j_hi; P.nz_relax = nz_relax;
    P.kappa = kappa; P.rpivot = rpivot; P.npivot = rpivot / dr[2]; P.rpivot_rt = P.npivot * dr[2];
    P.dr[0] = dr[0]; P.dr[1] = dr[1]; P.dr[2] = dr[2];
    for (int q = 0; q < 3; ++q) P.inv_dr[q] = 1.0 / dr[q];
    P.inv_n[0] = 1.0 / nx_global; P.inv_n[1] = 1.0 / ny; P.inv_n[2] = 1.0 / nzc;
    P.nd[0] = nx_global; P.nd[1] = ny; P.nd[2] = nzc;
    P.noortho = dR != nullptr;
    for (int q = 0; q < 9; ++q) P.Minv[q] = 0.0;
    if (dR) {
        const double dRT[9] = {dR[0], dR[3], dR[6], dR[1], dR[4], dR[7], dR[2], dR[5], dR[8]};
        if (!inv3(dRT, P.Minv)) return -1;
    }
    P.xtol = xtol; P.ftol = ftol; P.maxfev = maxfev; P.maxiter = 2000;
    P.k_hi = nz_relax - 1; P.k_lo = 0;
    P.x_base = x_base; P.nx_local = nx_local;
    P.out_etp = out_etp; P.out_cart = out_cart; P.out_nfev = out_nfev;
    const long long ncol = (long long)(i_hi - i_lo) * (j_hi - j_lo);
    std::vector<double> tr;
    if (transposed == 2) {
        // quad_yz_kernel block by block.  The body has one block barrier; the 32 x 1 "threads" of a block are run one
        // after the other, twice: the first round fills the shared tile, the second one stores from the complete tile
        tr.resize((size_t)nx * ny * nzc * 4);
        std::vector<double> tile(35 * 33);
        for (int x = 0; x < nx; ++x)
            for (int by = 0; by < (ny + 31) / 32; ++by)
                for (int bx = 0; bx < (nzc + 31) / 32; ++bx)
                    for (int round = 0; round < 2; ++round)
                        for (int tx = 0; tx < 32; ++tx)
                            quad_yz_tile(static_win, tr.data(), ny, nzc, tile.data(), bx, by, x, tx, 0, 1);
        P.grid = tr.data();
        // one "launch" per height, as the library does for large tiles: the warm start travels through out_etp
        for (int kh = nz_relax - 1; kh >= 0; --kh) {
            P.k_hi = P.k_lo = kh;
            for (long long c = 0; c < ncol; ++c) relax_column_body<2, 1>(P, c, ncol, powell_sm, 0, P.k_hi, P.k_lo);
        }
    } else if (transposed) {
        // what transpose_yz_kernel does, block by block: the body has one block barrier, so it is run
        // as a 32 x 1 thread block with each phase completed for all tx before the next starts --
        // done here by calling the body per tx with a private tile copy of phase 1 (tile is filled
        // identically by every call's phase 1 for its own tx column only, so fill first, then store)
        const int nyp = ny + 3;
        tr.resize((size_t)nx * nyp * nzc);
        std::vector<double> tile(32 * 33);
        for (int x = 0; x < nx; ++x)
            for (int by = 0; by < (nyp + 31) / 32; ++by)
                for (int bx = 0; bx < (nzc + 31) / 32; ++bx) {
                    const double *src = static_win + (size_t)x * ny * nzc;
                    for (int tx = 0; tx < 32; ++tx)
                        for (int r = 0; r < 32; ++r) {
                            const int t = by * 32 + r, z = bx * 32 + tx;
                            if (t < nyp && z < nzc) tile[r * 33 + tx] = src[(size_t)((t - 1 + ny) % ny) * nzc + z];
                        }
                    double *dst = tr.data() + (size_t)x * nyp * nzc;
                    for (int tx = 0; tx < 32; ++tx)
                        for (int r = 0; r < 32; ++r) {
                            const int z = bx * 32 + r, t = by * 32 + tx;
                            if (t < nyp && z < nzc) dst[(size_t)z * nyp + t] = tile[tx * 33 + r];
                        }
                }
        P.grid = tr.data();
        // launches of five heights
        for (int kh = nz_relax - 1; kh >= 0; kh -= 5) {
            P.k_hi = kh; P.k_lo = kh - 4 > 0 ? kh - 4 : 0;
            for (long long c = 0; c < ncol; ++c) relax_column_body<1, 1>(P, c, ncol, powell_sm, 0, P.k_hi, P.k_lo);
        }
    } else {
        for (long long c = 0; c < ncol; ++c) relax_column_body<0, 1>(P, c, ncol, powell_sm, 0, P.k_hi, P.k_lo);
    }
    if (out_cart) relax_cart_body(out_etp, out_cart, ncol * nz_relax, rpivot, 0, 1);
    return 0;
}

extern "C" void emu_sincos(const double *x, double *s, double *c, long long n)
{
    for (long long i = 0; i < n; ++i) dbspm_sincos(x[i], s + i, c + i);
}

extern "C" double emu_tricubic(const double *g, int nx, int ny, int nz, double x, double y, double z)
{
    const double inv_n[3] = {1.0 / nx, 1.0 / ny, 1.0 / nz}, nd[3] = {(double)nx, (double)ny, (double)nz};
    return tricubic_eval<0>(g, nx, ny, nz, 0, nd, inv_n, x, y, z);
}

// the column each thread of the relax launch takes (relax_patch_column); returns the number of threads launched
extern "C" long long emu_relax_patch_columns(int ni, int nj, int foot_x, long long *cols, long long cap)
{
    RelaxParams P;
    P.i_lo = 0; P.i_hi = ni; P.j_lo = 0; P.j_hi = nj; P.foot_x = foot_x;
    const long long ncol = (long long)ni * nj, nthr = relax_patch_threads(ni, nj, foot_x);
    for (long long g = 0; g < nthr && g < cap; ++g) cols[g] = relax_patch_column(P, g, ncol);
    return nthr;
}

// div_by(x, d, RN(1/d)) against x / d on random operands: returns the number of mismatching bit patterns
extern "C" long long emu_check_div_by(long long n, unsigned long long seed)
{
    unsigned long long st = seed * 6364136223846793005ULL + 1442695040888963407ULL;
    auto rnd = [&]() { st = st * 6364136223846793005ULL + 1442695040888963407ULL; return (double)(st >> 11) / 9007199254740992.0; };
    long long bad = 0;
    for (long long i = 0; i < n; ++i) {
        const double d = 0.01 + rnd() * 2.0;                       // grid spacings (Angstrom)
        const double x = (rnd() - 0.5) * 20.0 * pow(10.0, (int)(rnd() * 12) - 8);
        const double got = div_by(x, d, 1.0 / d), want = x / d;
        if (memcmp(&got, &want, sizeof got) != 0 && !(got == 0.0 && want == 0.0)) ++bad;
    }
    return bad;
}


// what dbspm_corr3d_direct_f64 does: plan on the host, the pad kernel body, then every tile (tensor-map load emulated by a
// plain copy with zero fill, all DIRECT_THREADS "threads" one after the other).  Returns the multiply-adds per output (> 0), < 0 on error
extern "C" long long emu_corr3d_direct_f64(const double *A, double alphaA, double dV, int nx, int ny, int nz, int z_lo, int z_hi,
                                           int npts, const int *sx, const int *sy, const int *sz, const double *w, double *E_win)
{
    const DirectPlan D = direct_make_plan(nx, ny, nz, z_lo, z_hi, npts, sx, sy, sz, w, dV);
    if (!D.ok) return -2;
    std::vector<double> P((size_t)D.PX * D.PY * D.PZ);
    DirectPadParams Qp;
    Qp.A = A; Qp.P = P.data(); Qp.nx = nx; Qp.ny = ny; Qp.nz = nz; Qp.PX = D.PX; Qp.PY = D.PY; Qp.PZ = D.PZ;
    Qp.x0 = D.start[0]; Qp.y0 = D.start[1]; Qp.z0 = (z_lo + D.start[2]) % nz; Qp.alpha = alphaA;
    direct_pad_body(Qp, 0, 1);
    DirectParams Q;
    Q.P = P.data(); Q.PX = D.PX; Q.PY = D.PY; Q.PZ = D.PZ; Q.E = E_win; Q.nx = nx; Q.ny = ny; Q.nzw = z_hi - z_lo;
    Q.Tx = D.Tx; Q.Ty = D.Ty; Q.Tz = D.Tz; Q.BX = D.BX; Q.BY = D.BY; Q.BZ = D.BZ;
    Q.tiles_y = D.tiles_y; Q.tiles_z = D.tiles_z; Q.nchunks = (int)D.chunk_off.size();
    Q.chunk_off = D.chunk_off.data(); Q.chunk_w = D.chunk_w.data();
    std::vector<double> box((size_t)D.BX * D.BY * D.BZ);
    const long long ntiles = (long long)D.tiles_x * D.tiles_y * D.tiles_z;
    for (long long t = 0; t < ntiles; ++t) {
        direct_tile_stage_host(Q, box.data(), t);
        for (int tid = 0; tid < DIRECT_THREADS; ++tid) direct_tile_compute(Q, box.data(), t, tid);
    }
    return D.fma_per_output;
}
