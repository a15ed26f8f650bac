/*
 * dbspm_b200.h -- C ABI of libdbspm_b200.so: the B200 (sm_100a) replacement for the
 * data-parallel hot path of SPMTH/DBSPM (sr / es interaction grids and the relax step).
 *
 * The reference is pure Python and has no FFI layer; the "interface each entry point
 * replaces" is therefore a Python function or a block of the `dbspm` driver script
 * (paths relative to the reference repository):
 *
 *   dbspm_corr3d_f64            pydbspm/interactions.py:9-22  get_density_overlap()
 *                               + the `**ALPHA` on both densities and the window slice
 *                               `sr[denscalc.nidx]` at dbspm:329-330 / dbspm:348-349
 *   dbspm_static_accumulate_f64 dbspm:479-486  static += sr*V ; static += es ; static += vdw
 *   dbspm_relax_powell_f64      dbspm:558-581 (tricubic(...) + loop over lateral points)
 *                               -> pydbspm/interactions.py:70-105 relax_tip_z /
 *                               relax_noortho_tip_z -> scipy.optimize.minimize("Powell")
 *                               -> tricubic .ip ; plus sph2cart at dbspm:603-605
 *   dbspm_tricubic_gather_f64   pydbspm/grid.py:325-361 interpolate_data (tricubic .ip at
 *                               explicit index-space points; also grid.py:122-139)
 *   dbspm_dfz_gradient_f64      pydbspm/SPMGrid.py:375-379 np.gradient(E, -dz, axis=2)
 *   dbspm_convz_valid_f64       pydbspm/tools.py:5-32 get_fs (np.convolve(..., 'valid') along z)
 *
 * Conventions
 *   - every array pointer is a DEVICE pointer to C-contiguous float64 memory (z fastest),
 *     owned by the caller for the whole call (e.g. a torch.cuda tensor's data_ptr());
 *   - no persistent allocation except cached FFT tables (a few KB per distinct length);
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = default);
 *     calls return after enqueueing, they do not synchronise;
 *   - return value 0 = success, negative = DBSPM_E* below; dbspm_last_error() gives a
 *     thread-local message; no C++ exception crosses this boundary;
 *   - one host thread per GPU; the library never changes the current device.
 */
#ifndef DBSPM_B200_H
#define DBSPM_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DBSPM_OK              0
#define DBSPM_EINVAL         -1   /* bad argument (NULL pointer, non-positive size, bad window) */
#define DBSPM_EUNSUPPORTED   -2   /* shape not supported (prime factor > 31 in a length, grid too large) */
#define DBSPM_EWORKSPACE     -3   /* workspace too small or misaligned */
#define DBSPM_ECUDA          -4   /* a CUDA runtime call / kernel launch failed */

/* flags for dbspm_corr3d_f64 */
#define DBSPM_CORR_DEFAULT    0
#define DBSPM_CORR_AXIS_V1    (1 << 1)   /* force the all-in-shared-memory axis pass (A/B timing) */
#define DBSPM_CORR_USE_TMA    (1 << 2)   /* force the TMA-fed persistent axis pass (tensor-map load, digit-reversing
                                            tensor-map store, two consumer groups + producer warp) */
#define DBSPM_CORR_NO_TMA     (1 << 3)   /* force the register-staged axis pass.  Default: picked per axis length --
                                            TMA for 192 <= n <= 576 (8-20 % faster on B200), register-staged
                                            otherwise (n = 1024 is FP64/shared-memory bound with 4-line tiles) */
#define DBSPM_CORR_PEER_NARROW (1 << 4)  /* fused exchange: keep the 4-line tiles of the local pass (64-byte remote stores; A/B timing) */
#define DBSPM_CORR_NO_PREFETCH (1 << 5)  /* no L2 prefetch of the next tile / pair group (A/B timing) */
#define DBSPM_CORR_SLOW_POW   (1 << 6)   /* v^alpha as libdevice exp(alpha*log(v)) instead of the table-driven pow_fast (A/B) */
#define DBSPM_CORR_POW_WIDE   (1 << 7)   /* with DBSPM_CORR_NO_PIPE: the one-shot powered x pass on 8-line tiles / one CTA of 512 threads (A/B timing;
                                            pipelined, the powered pass takes 8-line tiles by default, see DBSPM_CORR_POW_NARROW) */
/* profiling aid: skip stages (results are then meaningless); bench.py times each launch alone */
#define DBSPM_CORR_SKIP_XFWD  (1 << 8)
#define DBSPM_CORR_SKIP_YFWD  (1 << 9)
#define DBSPM_CORR_SKIP_Z     (1 << 10)
#define DBSPM_CORR_SKIP_YINV  (1 << 11)
#define DBSPM_CORR_SKIP_XINV  (1 << 12)
#define DBSPM_CORR_Z_RADIX8   (1 << 13)  /* fused z kernel: plan with radices <= 8 and the kernel instantiated without the radix-16
                                            butterfly (120 instead of 166 registers, one more shared-memory stage; A/B timing:
                                            2.67 against 2.75 ms at nz = 256 on B200 -- the kernel is not occupancy-bound) */

#define DBSPM_CORR_NO_PIPE    (1 << 14)  /* plain axis passes as one-shot blocks only.  Default: where the plan is 16*8*8 or 16*8*4 (n = 1024, 512,
                                            128, 64) and the pass has many more tiles than resident CTAs, persistent blocks request the next
                                            tile's rows (8 per thread through cp.async into the thread's own shared-memory slots, 8 into
                                            registers) before the remaining stages of the current tile.  B200, n = 1024: y pass 2.39 -> 2.30 ms (2.08 on 8-line tiles),
                                            x pass on 8-line tiles 2.68 -> 2.42 ms (3.91 -> 3.61 ms with the power fused); slower with 4-line
                                            tiles on the slowest axis (3.23 ms; 4.28 against 3.90 ms with the power), where the one-shot blocks stay */

#define DBSPM_CORR_PIPE_NARROW_Y (1 << 15) /* A/B: keep 4-line tiles (two CTAs of 256 threads) in the pipelined y-type passes.  Default: 8-line
                                            tiles / one CTA of 512 threads when the rows are at least 64 elements and a multiple of 8
                                            (B200, n = 1024: 2.29 -> 2.08 ms) */

#define DBSPM_CORR_POW_NARROW  (1 << 16)  /* A/B: powered x pass as one-shot blocks on 4-line tiles (the default of round 1) instead of the
                                            pipelined pass on 8-line tiles (B200, n = 1024: 3.91 against 3.61 ms) */

/* library / build identification: "dbspm_b200 <version> sm_100a" */
const char *dbspm_version(void);
const char *dbspm_last_error(void);

/* ---- sr / es: periodic 3-D cross-correlation, output restricted to a z-window ------------
 * E_win[x,y,z-z_lo] = dV * sum_r A[r]^alphaA * B[(r - R) mod N]^alphaB ,  R = (x,y,z), z in [z_lo,z_hi)
 * A, B: (nx,ny,nz); E_win: (nx,ny,z_hi-z_lo).  alpha == 1.0 skips the power.
 * Requires 0 <= z_lo < z_hi <= nz and nx, ny, nz (nz/2 for even nz) with prime factors <= 31.  An odd nz
 * takes plain complex transforms (twice the traffic of the packed real path; VASP grids are even).
 * workspace: at least dbspm_corr3d_workspace_bytes(...) bytes, 256-byte aligned.           */
size_t dbspm_corr3d_workspace_bytes(int nx, int ny, int nz, int z_lo, int z_hi, int flags);
int dbspm_corr3d_f64(const double *A, const double *B, double alphaA, double alphaB, double dV,
                     int nx, int ny, int nz, int z_lo, int z_hi, double *E_win,
                     void *workspace, size_t workspace_bytes, int flags, void *stream);

/* ---- sr / es with a tip that is exactly zero outside a z-range ("compact tip") ----
 * Same result as dbspm_corr3d_f64 when B[x,y,z] == 0 for every z outside the s_len planes
 * (s_lo + m) mod nz, m in [0, s_len): only the planes of A the window can see are transformed, on a
 * z length Mp ~ (z_hi - z_lo) + s_len instead of nz (all five passes shrink by Mp/nz; the first pass
 * reads the z-ranges of A and B in place, no packed copy).  Falls back to the full-length path when
 * Mp > 0.9 nz or nx has a prime factor > 7.  The support is the caller's claim: find it with
 * dbspm_zsupport_f64 (plane_nonzero: device int[nz], 1 where a plane of B holds a nonzero value;
 * the smallest circular interval covering the ones is (s_lo, s_len)).  Replaces the same reference
 * lines as dbspm_corr3d_f64; the reference always pays the full transform.                      */
size_t dbspm_corr3d_pruned_workspace_bytes(int nx, int ny, int nz, int z_lo, int z_hi, int s_lo, int s_len, int flags);
int dbspm_corr3d_pruned_f64(const double *A, const double *B, double alphaA, double alphaB, double dV,
                            int nx, int ny, int nz, int z_lo, int z_hi, int s_lo, int s_len, double *E_win,
                            void *workspace, size_t workspace_bytes, int flags, void *stream);
int dbspm_zsupport_f64(const double *B, int nx, int ny, int nz, int *plane_nonzero, void *stream);

/* ---- sr / es as a direct real-space sum for a tip with a few hundred support points ---------------------------------
 * E_win[x,y,z-z_lo] = dV * sum_k w_k * A[(R + s_k) mod N]^alphaA: the same quantity as dbspm_corr3d_f64 when B is nonzero
 * only at the npts points s_k = (sx,sy,sz)[k] (indices into B) with w_k = B[s_k]^alphaB.  N_w * K fused multiply-adds
 * instead of five passes over both grids: cheaper than the spectral path for small K (the crossover is measured in
 * profiles/; dbspm_corr3d_direct_cost returns the multiply-adds per output point for the caller's choice, 0 when the
 * support's bounding box does not fit a shared-memory tile).  One kernel builds A^alphaA on the planes the window can see
 * with a periodic halo; the main kernel stages one sample box per output tile with a single tensor-map load
 * (cp.async.bulk.tensor.3d) and keeps a sliding window of the box in registers.  sx, sy, sz, w are HOST arrays;
 * dbspm_support_points_f64 compacts the nonzero entries of a device array (count_dev: device int, entries beyond
 * max_points are counted but not stored).  Replaces pydbspm/interactions.py:18-22 + dbspm:329-330 / :348-349.            */
int dbspm_support_points_f64(const double *B, int nx, int ny, int nz, int max_points, int *count_dev,
                             long long *idx_dev, double *val_dev, void *stream);
size_t dbspm_corr3d_direct_workspace_bytes(int nx, int ny, int nz, int z_lo, int z_hi, int npts,
                                           const int *sx, const int *sy, const int *sz);
long long dbspm_corr3d_direct_cost(int nx, int ny, int nz, int z_lo, int z_hi, int npts,
                                   const int *sx, const int *sy, const int *sz);
int dbspm_corr3d_direct_f64(const double *A, double alphaA, double dV, int nx, int ny, int nz, int z_lo, int z_hi,
                            int npts, const int *sx, const int *sy, const int *sz, const double *w,
                            double *E_win, void *workspace, size_t workspace_bytes, int flags, void *stream);

/* ---- sr / es distributed over the G ranks of a group (x-slab decomposition) ------------------
 * Replaces the same reference lines as dbspm_corr3d_f64 (the reference computes sr / es on rank 0 only,
 * dbspm:320,339).  Rank g holds the x-slab [g*nx/G, (g+1)*nx/G) of A and B -- contiguous in the .npz member, so a
 * rank reads and uploads 1/G of each input.  The caller owns the two exchanges (NCCL through torch.distributed):
 *   1. dbspm_corr3d_dist_fwd_f64: y transform of the slab (power fused on load) written destination-major:
 *      sendA/sendB = (G, nxl, ny/G, nz/2) complex128; block h holds the ny/G y-frequencies rank h owns.
 *      Frequencies are dealt in slot order 0, ny/2, 1, ny-1, 2, ny-2, ... so that ky and -ky sit on one rank.
 *   2. all-to-all of the G blocks  ->  recvA/recvB = (nx, ny/G, nz/2) on every rank.
 *   3. dbspm_corr3d_dist_mid_f64 (rank h of G): x transform in place, fused z kernel (forward z, spectral product,
 *      inverse z of the window planes only), inverse x transform -> W_local = (nx, ny/G, ceil(nzw/2)) complex128.
 *   4. gather of the G W_local blocks -> W_all = (G, nx, ny/G, ceil(nzw/2)).
 *   5. dbspm_corr3d_dist_fin_f64: inverse y transform reading its rows through the slot table -> E_win (nx,ny,nzw).
 * Requires even nz, nx % G == 0, ny % (2G) == 0 and lengths nx, ny whose prime factors are <= 7
 * (dbspm_corr3d_dist_supported returns 0 otherwise; callers then use dbspm_corr3d_f64 on replicated inputs).
 * G = 1 needs no exchange and equals dbspm_corr3d_f64.  fin needs a workspace only for an odd window.          */
int dbspm_corr3d_dist_supported(int nx, int ny, int nz, int G);
int dbspm_corr3d_dist_fwd_f64(const double *A_slab, const double *B_slab, double alphaA, double alphaB,
                              int nxl, int ny, int nz, int G, void *sendA, void *sendB, int flags, void *stream);
/* Steps 1 + 2 fused (compute + exchange in one kernel): the y pass of rank `me` stores every output row straight into
 * its owner's receive buffer through peer memory (NVLink / NVSwitch); no send buffer and no all-to-all.
 * peer_recv: HOST array of G device pointers, peer_recv[h] = rank h's receive buffer (2, nx, ny/G, nz/2) complex128
 * (A's spectra, then B's) mapped into this process (e.g. torch symmetric memory buffer_ptrs).  The caller brackets the
 * call with two group barriers: every rank has finished reading its buffer / every rank's stores have landed.      */
int dbspm_corr3d_dist_fwd_peer_f64(const double *A_slab, const double *B_slab, double alphaA, double alphaB,
                                   int nxl, int ny, int nz, int G, int me, void *const *peer_recv, int flags, void *stream);
/* Compact tip on the distributed path: the forward y pass reads the z-ranges of the pruned problem in place (as
 * dbspm_corr3d_pruned_f64 does in its x pass) and writes spectra of z length Mp.  dbspm_corr3d_pruned_geometry fills
 * cut = {use, Mp, zA0, zB0, validA, validB, w_lo}; with use == 1 pass `cut` to a *_cut_f64 entry (buffers sized with Mp
 * instead of nz) and run mid / fin with nz = Mp and the window [w_lo, w_lo + z_hi - z_lo).  cut == NULL: full length. */
int dbspm_corr3d_pruned_geometry(int nz, int z_lo, int z_hi, int s_lo, int s_len, int *cut /* 7 ints */);
int dbspm_corr3d_dist_fwd_cut_f64(const double *A_slab, const double *B_slab, double alphaA, double alphaB,
                                  int nxl, int ny, int nz_src, int G, const int *cut, void *sendA, void *sendB,
                                  int flags, void *stream);
int dbspm_corr3d_dist_fwd_peer_cut_f64(const double *A_slab, const double *B_slab, double alphaA, double alphaB,
                                       int nxl, int ny, int nz_src, int G, int me, const int *cut,
                                       void *const *peer_recv, int flags, void *stream);
int dbspm_corr3d_dist_mid_f64(void *recvA, void *recvB, double dV, int nx, int ny, int nz, int G, int h,
                              int z_lo, int z_hi, void *W_local, int flags, void *stream);
/* Step 3 with a second, small exchange fused into its inverse x pass: row x of the packed window is stored straight into
 * the buffer of the rank that owns x-slab x / (nx/G), so that step 5 runs distributed (every rank: its own x-slab) and
 * rank 0 only gathers finished real slabs.  peer_w2[q] = rank q's (G, nx/G, ny/G, ceil(nzw/2)) complex128 buffer;
 * afterwards (group barrier) rank q calls dbspm_corr3d_dist_fin_f64(its buffer, nx/G, ny, G, ...) -> (nx/G, ny, nzw).  */
int dbspm_corr3d_dist_mid_peer_f64(void *recvA, void *recvB, double dV, int nx, int ny, int nz, int G, int h,
                                   int z_lo, int z_hi, void *W_local, void *const *peer_w2, int flags, void *stream);
size_t dbspm_corr3d_dist_fin_workspace_bytes(int nx, int ny, int z_lo, int z_hi);
int dbspm_corr3d_dist_fin_f64(const void *W_all, int nx, int ny, int G, int z_lo, int z_hi, double *E_win,
                              void *workspace, size_t workspace_bytes, int flags, void *stream);

/* ---- static potential assembly: acc = (init ? 0 : acc) + comp*scale  (separate mul, add) -- */
int dbspm_static_accumulate_f64(double *acc, const double *comp, double scale, size_t n,
                                int init, void *stream);

/* ---- relax: Powell minimisation of E(theta,phi) at every tip position of a lateral tile ---
 * static_win: (nx,ny,nzc) window potential, periodic in all three axes for interpolation.
 * Lateral tile i in [i_lo,i_hi), j in [j_lo,j_hi); z indices k = 0..nz_relax-1 swept from the
 * top down with warm start.  rpivot in Angstrom, dr = grid spacings; npivot = rpivot/dr[2] and
 * the pivot plane index is trunc(k + npivot) exactly as the reference's int64 origin does.
 * dR: 9 doubles (row-major 3x3 lattice step matrix) selects the non-orthogonal offset
 * inv(dR^T).(x,y,z); NULL selects the orthogonal path.
 * out_E_theta_phi: (ni,nj,nz_relax,3) rows [E_min, theta, phi]
 * out_cart:        (3,ni,nj,nz_relax) tip_x, tip_y, tip_z = sph2cart(theta, phi, rpivot); nullable
 * out_nfev:        (ni,nj,nz_relax) objective evaluations per position; nullable           */
int dbspm_relax_powell_f64(const double *static_win, int nx, int ny, int nzc,
                           int i_lo, int i_hi, int j_lo, int j_hi, int nz_relax,
                           double kappa, double rpivot, const double dr[3], const double *dR,
                           double xtol, double ftol, int maxfev,
                           double *out_E_theta_phi, double *out_cart, int *out_nfev,
                           void *workspace, size_t workspace_bytes, void *stream);
/* x-slab form (multi-GPU: a rank holds only its tile of the static window plus a halo): static_slab holds the nx_local
 * planes x_base, x_base + 1, ... (mod nx) of the (nx,ny,nzc) window.  The slab must cover the tile [i_lo, i_hi) and
 * dbspm_relax_slab_halo(rpivot, dr, dR) planes on both sides (the lever length in planes + the stencil).  Cell indices
 * and interpolation weights are computed from the global coordinate, so the results are bit-identical to
 * dbspm_relax_powell_f64 on the whole window.  Replaces the replicated `static` of dbspm:540-556.              */
int dbspm_relax_slab_halo(double rpivot, const double dr[3], const double *dR);
int dbspm_relax_powell_slab_f64(const double *static_slab, int nx, int ny, int nzc, int x_base, int nx_local,
                                int i_lo, int i_hi, int j_lo, int j_hi, int nz_relax,
                                double kappa, double rpivot, const double dr[3], const double *dR,
                                double xtol, double ftol, int maxfev,
                                double *out_E_theta_phi, double *out_cart, int *out_nfev,
                                void *workspace, size_t workspace_bytes, void *stream);
/* workspace (nullable), 32-byte aligned: the potential is re-laid once per call so that the stencil loads of a warp
 * waste less of every cache line.  Layout 1 (dbspm_relax_workspace_bytes_layout(.., 1) bytes): y fastest with a periodic
 * halo.  Layout 2 (4x the window; returns 0 when 32-bit element offsets would overflow): the four y neighbours of every
 * stencil row as one 32-byte sector, one 256-bit load per row.  The kernel takes the largest layout the workspace
 * holds; dbspm_relax_workspace_bytes returns the preferred size (layout 2, else 1).  Without a workspace the kernel
 * reads static_win in place.  Results are bit-identical in all three cases.
 * The sizes include a small control block behind the re-laid potential (a task counter and one progress word per patch
 * of 32 columns): with more patches than resident warps the heights are relaxed in chunks pulled by one wave of
 * persistent warps (dynamic schedule, same bits; environment: DBSPM_RELAX_CHUNK = heights per chunk, 0 = static
 * launches; DBSPM_RELAX_FOOT_X = columns of a warp's patch along x, 1 = 32 consecutive y).                          */
size_t dbspm_relax_workspace_bytes(int nx, int ny, int nzc);
size_t dbspm_relax_workspace_bytes_layout(int nx, int ny, int nzc, int layout);

/* ---- tricubic gather: out[p] = interp(data)(pts[p,0], pts[p,1], pts[p,2]), index units ---- */
int dbspm_tricubic_gather_f64(const double *data, int nx, int ny, int nz,
                              const double *pts /* (npts,3) */, size_t npts, double *out, void *stream);

/* ---- dE/dz image: out = np.gradient(E, -dz, axis=2) for E of shape (nx,ny,nz) ------------- */
int dbspm_dfz_gradient_f64(const double *E, int nx, int ny, int nz, double dz, double *out, void *stream);

/* ---- valid 1-D convolution along z: out (nx,ny,nz-nw+1) = scale * np.convolve(F[x,y,:], w, "valid") ----
 * (the force -> frequency-shift weighting of pydbspm/tools.py:5-32 get_fs; w, nw <= 64, on the device) */
int dbspm_convz_valid_f64(const double *F, int nx, int ny, int nz, const double *w, int nw, double scale,
                          double *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DBSPM_B200_H */
