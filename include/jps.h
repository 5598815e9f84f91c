/*
 * jps.h -- C ABI of libjps.so: the B200 (sm_100a) implementation of jax-powspec's
 * mesh-painting + Fourier-space clustering hot path.
 *
 * This is the drop-in boundary.  Each entry point replaces one reference function
 * (paths relative to the reference repository root):
 *
 *   jps_paint                 src/mas.py:88-153   cic_mas_vec   (variant JPS_VARIANT_VEC)
 *                             src/mas.py:5-87     cic_mas       (variant JPS_VARIANT_SCAN)
 *                             + TSC / PCS (order 3 / 4), absent from the reference
 *   jps_powspec               src/correlations.py:7-56     powspec_vec
 *   jps_powspec_fundamental   src/correlations.py:60-117   powspec_vec_fundamental
 *   jps_xi                    src/correlations.py:120-187  xi_vec  (and the xi blocks :522-543, :689-710)
 *   jps_xi_fundamental        src/correlations.py:191-261  xi_vec_fundamental
 *   jps_bispec                src/correlations.py:334-462  bispec
 *   jps_paint_powspec         tests/correlations.py:41-78  paint -> delta=rho/mean-1 -> powspec_vec
 *   jps_text_parse            tests/correlations.py:29-31  np.loadtxt(usecols, float32) + box mask
 *
 * The reference has no FFI of its own (it is pure Python on jax.numpy); these are the
 * symbols a jax.ffi handler, a ctypes stub or any other host binds.
 *
 * Conventions
 *   - every function returns 0 on success, a negative JPS_ERR_* otherwise;
 *     jps_last_error() returns a thread-local message for the last failure.
 *   - all pointers documented "device" are CUDA device pointers BORROWED from the caller
 *     (they must stay alive until `stream` reaches the operation); "host" pointers are
 *     read before the call returns.
 *   - all work is enqueued on the caller's `stream`; no call synchronises the device or
 *     touches the default stream (exception: jps_plan_create, which builds cuFFT plans).
 *   - no hidden device allocation: the caller passes the workspace whose size the
 *     matching *_workspace_bytes() function reports.
 *   - a plan is bound to the CUDA device current at creation and is not thread-safe;
 *     distinct plans may be used concurrently from different threads / streams.
 *   - `stream` is a cudaStream_t passed as void* so that this header needs no CUDA include.
 */
#ifndef JPS_H_
#define JPS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JPS_VERSION 100

#if defined(__GNUC__)
#define JPS_API __attribute__((visibility("default")))
#else
#define JPS_API
#endif

/* error codes */
#define JPS_OK                 0
#define JPS_ERR_INVALID       -1   /* bad argument */
#define JPS_ERR_CUDA          -2   /* CUDA runtime error */
#define JPS_ERR_CUFFT         -3   /* cuFFT error */
#define JPS_ERR_WORKSPACE     -4   /* workspace too small */
#define JPS_ERR_UNSUPPORTED   -5   /* valid request this build cannot serve */

/* mass-assignment order */
#define JPS_ORDER_CIC 2
#define JPS_ORDER_TSC 3
#define JPS_ORDER_PCS 4

/* compat: JPS_COMPAT_REFERENCE reproduces the reference bit-for-bit in semantics (SURVEY.md
 * section 8 quirks Q1-Q3, Q18); JPS_COMPAT_FIXED uses the textbook weights, floor() cell
 * choice and periodic wrap of every index. */
#define JPS_COMPAT_REFERENCE 0
#define JPS_COMPAT_FIXED     1

/* which of the two reference painters' wrap handling to follow (only differs for wrap=0 and
 * for particles outside the box) */
#define JPS_VARIANT_VEC  0
#define JPS_VARIANT_SCAN 1

/* painter algorithm */
#define JPS_PAINT_AUTO    0
#define JPS_PAINT_ATOMIC  1   /* one thread per particle, global red.add (any order of particles) */
#define JPS_PAINT_SORTED  2   /* bucket by mesh tile, deposit in shared memory, float4 flush */

/* plan flags */
#define JPS_PLAN_DEFAULT      0
#define JPS_PLAN_TABLES_ONLY  1   /* bin tables + accumulators only: no 3-D FFT plans, no delta_k buffer */
#define JPS_PLAN_FFT_PENCIL   2   /* forward transform as three contiguous batched 1-D cuFFT passes with two
                                     transposing kernels in between (spectrum left as [kz][ky][kx]); for big
                                     meshes (2048^3: the monolithic 3-D plan needs 93 ms, this ~60).  Such a
                                     plan serves jps_powspec / jps_paint_powspec only; it holds a second
                                     delta_k-sized buffer (n_shell_fields must be 0). */

typedef struct jps_plan jps_plan_t;
typedef struct jps_slab_plan jps_slab_plan_t;

JPS_API int         jps_version(void);
JPS_API const char* jps_last_error(void);

/* ------------------------------------------------------------------ launch accounting */
/* The library counts every kernel it launches, per kernel kind.  With profiling enabled each
 * launch is also bracketed by CUDA events on the launching stream; jps_profile_get()
 * synchronises on those events and returns the accumulated device time.  Used by bench.py
 * for `gpu_launches` and the per-kernel roofline; off by default (zero overhead). */
JPS_API int jps_profile_enable(int on);
JPS_API int jps_profile_reset(void);
JPS_API int jps_profile_num_kernels(void);
JPS_API int jps_profile_get(int id, const char** name, unsigned long long* launches, double* ms);

/* ------------------------------------------------------------------ plan ------------- */
/* Bytes of device workspace a plan for an n_mesh^3 grid needs (delta_k buffer, cuFFT work
 * area, lookup tables, accumulators).  n_shell_fields > 0 additionally reserves room for
 * that many real-space shell fields (bispectrum); 0 for P(k)-only plans. */
JPS_API int jps_plan_workspace_bytes(int n_mesh, int n_shell_fields, int flags, size_t* bytes);

/* Build a plan on the current device.  `workspace` is device memory of at least
 * jps_plan_workspace_bytes(...) bytes, 256-byte aligned, owned by the caller and kept
 * alive until jps_plan_destroy. */
JPS_API int jps_plan_create(int n_mesh, int n_shell_fields, int flags,
                    void* workspace, size_t workspace_bytes, jps_plan_t** plan);
JPS_API int jps_plan_destroy(jps_plan_t* plan);

/* ------------------------------------------------------------------ painting --------- */
/* Workspace for jps_paint (bucketed copy of the particles + tile offsets). */
JPS_API int jps_paint_workspace_bytes(int n_mesh, int64_t n_part, int order, int method, size_t* bytes);

/* mesh[n,n,n] (device, float32, C order) += deposit of n_part particles.
 *   x,y,z,w : device float32; element i is at p[i*stride] (stride 1 = SoA arrays, 3 = the
 *             columns of an (Np,3) row-major array); w may be NULL (unit weights, stride 1).
 *   xmin..  : per-axis origin; box_size: cubic box; pos = (x-xmin) * (1/(box_size/n_mesh)) in
 *             float32 exactly as src/mas.py:100-105.
 *   wrap    : the reference's `wrap` argument. */
JPS_API int jps_paint(int n_mesh,
              const float* x, const float* y, const float* z, const float* w,
              int64_t stride, int64_t n_part,
              float xmin, float ymin, float zmin, float box_size,
              int order, int wrap, int compat, int variant, int method,
              float* mesh, void* workspace, size_t workspace_bytes, void* stream);

/* Same deposit into an x-SLAB of the mesh: `mesh` is [nx_alloc, n, n] and its plane 0 holds global
 * plane x0 (x0 may be negative: planes are taken mod n_mesh).  Stencil nodes whose plane is not in
 * [x0, x0+nx_alloc) are dropped, so the caller sizes the slab with ghost planes and adds them to
 * the neighbours (jax_powspec_b200/slab.py).  jps_paint is jps_paint_slab(x0=0, nx_alloc=n_mesh). */
JPS_API int jps_paint_slab(int n_mesh, int x0, int nx_alloc,
                   const float* x, const float* y, const float* z, const float* w,
                   int64_t stride, int64_t n_part,
                   float xmin, float ymin, float zmin, float box_size,
                   int order, int wrap, int compat, int variant, int method,
                   float* mesh, void* workspace, size_t workspace_bytes, void* stream);

/* The two halves of jps_paint_slab(method = JPS_PAINT_SORTED) as separate calls, for pipelines that overlap the
 * deposit with the transform of the planes already finished (jax_powspec_b200/slab.py): phase
 * JPS_PAINT_PHASE_BUCKET partitions the catalogue by 16^3-cell tile into `workspace`; JPS_PAINT_PHASE_DEPOSIT
 * deposits the tiles of tile rows [tx_begin, tx_end) along x (16 planes each, jps_paint_tile_rows(nx_alloc) rows
 * in all; a tile row also touches the next order-1 planes) from a workspace bucketed by an earlier BUCKET call
 * with the SAME arguments.  JPS_PAINT_PHASE_ALL is jps_paint_slab.  (No reference counterpart.) */
#define JPS_PAINT_PHASE_ALL     0
#define JPS_PAINT_PHASE_BUCKET  1
#define JPS_PAINT_PHASE_DEPOSIT 2
JPS_API int jps_paint_tile_rows(int nx_alloc);
JPS_API int jps_paint_slab_phase(int n_mesh, int x0, int nx_alloc,
                         const float* x, const float* y, const float* z, const float* w,
                         int64_t stride, int64_t n_part,
                         float xmin, float ymin, float zmin, float box_size,
                         int order, int wrap, int compat, int variant, int method,
                         float* mesh, void* workspace, size_t workspace_bytes,
                         int phase, int tx_begin, int tx_end, void* stream);

/* ------------------------------------------------------------------ P(k) ------------- */
/* Power-spectrum multipoles of a mesh in user bins.
 *   mesh       : device float32 [n,n,n]; not modified.
 *   normalise  : 0 = mesh already holds delta (what powspec_vec receives);
 *                1 = mesh holds rho: delta = rho/mean(rho) - 1 is applied in Fourier space
 *                    through the DC mode (tests/correlations.py:49-50 folded in).
 *   k_edges    : HOST float32 [nb+1], ascending, in h/Mpc; converted to grid units in float32
 *                as src/correlations.py:42.
 *   mas_order  : exponent p of the window deconvolution (1/sinc)^p per axis; the reference
 *                hard-codes 2.
 *   shot_noise : subtracted from P0 (the reference never does: pass 0).
 * Outputs (device float32, as the reference returns): k3d[nb], pk3d[nb*3] row-major
 * (P0,P2,P4), nmodes[nb].  Optional (may be NULL): sums[nb*3] float64 raw sums of
 * |delta_k|^2 L_ell, counts[nb] int64 exact mode counts. */
JPS_API int jps_powspec(jps_plan_t* plan, const float* mesh, int normalise, float box_size,
                const float* k_edges, int nb, int mas_order, float shot_noise,
                float* k3d, float* pk3d, float* nmodes,
                double* sums, int64_t* counts, void* stream);

/* jps_powspec with the two estimator options the reference lacks (parity unpinned,
 * defined by oracle/correlations.py:powspec and the analytic tests):
 *   mesh2 : NULL, or the SAME particles painted on the grid displaced by +half a cell (jps_paint with
 *           xmin + cell/2 on every axis) -> interlaced spectrum (dk1 + dk2 e^{-i pi (kx+ky+kz)/N}) / 2,
 *           the aliased images with odd m_x+m_y+m_z cancel; needs n_shell_fields >= 1
 *   flags : JPS_PK_HERMITIAN counts every stored mode with 0 < kz < N/2 twice (its mirror image is not
 *           stored), i.e. the full-space shell average; 0 = the reference's half-space counting (Q7). */
#define JPS_PK_HERMITIAN 1
JPS_API int jps_powspec_ex(jps_plan_t* plan, const float* mesh, const float* mesh2, int normalise, float box_size,
                   const float* k_edges, int nb, int mas_order, float shot_noise, int flags,
                   float* k3d, float* pk3d, float* nmodes, double* sums, int64_t* counts, void* stream);

/* Number of rows powspec_vec_fundamental returns for an n_mesh grid: floor(sqrt(3)*(n/2)). */
JPS_API int jps_fundamental_nbins(int n_mesh);

/* kF-wide integer bins, bin 0 dropped.  compat selects how k3d is formed (Q18). */
JPS_API int jps_powspec_fundamental(jps_plan_t* plan, const float* mesh, int normalise, float box_size,
                            int mas_order, int compat,
                            float* k3d, float* pk3d, float* nmodes,
                            double* sums, int64_t* counts, void* stream);

/* ------------------------------------------------------------------ xi(s) ------------ */
/* Configuration-space multipoles (needs a plan with n_shell_fields >= 1).
 *   s_edges  : HOST float32 [nb+1] separations in Mpc/h; converted to grid units as
 *              src/correlations.py:126,170.
 *   guard_mu : 0 = xi_vec (mu = rz/|r| is NaN at r = 0, which poisons the first bin of xi2/xi4
 *              when s_edges[0] == 0, quirk Q22); 1 = the composites' guarded mu (:527).
 * Outputs (device float32): r3d[nb], xi3d[nb*3], nmodes[nb] (empty bins: inf, as the reference). */
JPS_API int jps_xi(jps_plan_t* plan, const float* mesh, int normalise, float box_size,
           const float* s_edges, int nb, int mas_order, int guard_mu,
           float* r3d, float* xi3d, float* nmodes, double* sums, int64_t* counts, void* stream);

/* Integer-lag bins, bin 0 dropped; jps_fundamental_nbins(n_mesh) rows. */
JPS_API int jps_xi_fundamental(jps_plan_t* plan, const float* mesh, int normalise, float box_size,
                       int mas_order, float* r3d, float* xi3d, float* nmodes,
                       double* sums, int64_t* counts, void* stream);

/* ------------------------------------------------------------------ bispectrum ------- */
/* FFT bispectrum for fixed (k1, k2) and nbins opening angles (needs n_shell_fields >= 6).
 *   theta : HOST float32 [nbins] radians.
 * Outputs (device float32): k_all[nbins+2], pk[nbins+2] (P at every shell), B[nbins], Q[nbins]. */
JPS_API int jps_bispec(jps_plan_t* plan, const float* mesh, int normalise, float box_size,
               float k1, float k2, const float* theta, int nbins, int mas_order,
               float* k_all, float* pk, float* B, float* Q, void* stream);

/* BASELINE.json configs[2] ("all triangle bins up to k_max"): the reference's bispec() evaluated for
 * npairs (k1, k2) pairs -- what a caller's double loop over shell centres does
 * (tests/bispec.py:53-56 is one such call) -- sharing ONE forward FFT and the cached indicator sums.
 * Pair p gives exactly what jps_bispec(k1[p], k2[p]) gives.
 *   k1, k2 : HOST float32 [npairs];  theta : HOST float32 [nbins]
 * Outputs (device float32, row p = pair p): k_all[npairs][nbins+2], pk[npairs][nbins+2],
 * B[npairs][nbins], Q[npairs][nbins]. */
JPS_API int jps_bispec_pairs(jps_plan_t* plan, const float* mesh, int normalise, float box_size,
                     const float* k1, const float* k2, int npairs, const float* theta, int nbins,
                     int mas_order, float* k_all, float* pk, float* B, float* Q, void* stream);

/* ------------------------------------------------------------------ composites ------- */
/* src/correlations.py:640-712: P(k) + xi(s) sharing ONE forward FFT (n_shell_fields >= 1). */
JPS_API int jps_compute_2pt_correlations(jps_plan_t* plan, const float* mesh, int normalise, float box_size,
                                 const float* s_edges, int ns, const float* k_edges, int nk, int mas_order,
                                 float* k3d, float* pk3d, float* nmodes_pk,
                                 float* r3d, float* xi3d, float* nmodes_xi, void* stream);

/* src/correlations.py:464-637: P(k) + xi(s) + bispectrum sharing ONE forward FFT
 * (n_shell_fields >= 6). */
JPS_API int jps_compute_all_correlations(jps_plan_t* plan, const float* mesh, int normalise, float box_size,
                                 const float* s_edges, int ns, const float* k_edges, int nk,
                                 float k1, float k2, const float* theta, int nbins, int mas_order,
                                 float* k3d, float* pk3d, float* nmodes_pk,
                                 float* r3d, float* xi3d, float* nmodes_xi,
                                 float* k_all, float* pk_shell, float* B, float* Q, void* stream);

/* ------------------------------------------------------------------ slab-sharded mesh - */
/* Per-rank compute stages of the distributed path (one process per GPU; rank r of nranks owns
 * x-planes [r*n/nranks, (r+1)*n/nranks); n_mesh % nranks == 0).  The reference has no
 * multi-device code: these have no reference counterpart; the host side that strings them
 * together with the halo exchange, the all-to-all transpose and the allreduce of the bin sums is
 * jax_powspec_b200/slab.py.  All buffers are caller-owned device memory:
 *   slab      float32  [n/nranks][n][n]            owned planes of the painted mesh
 *   yz        complex64 [n/nranks][n][n/2+1]       after jps_slab_fft_yz
 *   packed    complex64 [nranks][n/nranks][n/nranks][n/2+1]   after jps_slab_pack (all-to-all send buffer)
 *   dk        complex64 [n][n/nranks][n/2+1]       all-to-all receive buffer; jps_slab_fft_x in place;
 *                                                  element (ix, yl, kz) is mode (kx(ix), ky(rank*n/nranks+yl), kz)
 */
JPS_API int jps_slab_plan_workspace_bytes(int n_mesh, int nranks, size_t* bytes);
JPS_API int jps_slab_plan_create(int n_mesh, int nranks, int rank, void* workspace, size_t workspace_bytes,
                         jps_slab_plan_t** plan);
JPS_API int jps_slab_plan_destroy(jps_slab_plan_t* plan);
JPS_API int jps_slab_fft_yz(jps_slab_plan_t* plan, const float* slab, void* yz, void* stream);
JPS_API int jps_slab_pack(jps_slab_plan_t* plan, const void* yz, void* packed, void* stream);
/* Fused pack + all-to-all through NVLink peer memory: writes block q of `yz` straight into
 * peer_recv[q] (HOST array of nranks DEVICE pointers to every rank's `dk` buffer, peer-mapped with
 * CUDA IPC by the caller; peer_recv[rank] is the local buffer), at this rank's slot.  Replaces
 * jps_slab_pack + the NCCL all-to-all.  The caller orders it against the peers' use of `dk`
 * (stream-ordered collectives before and after; see jax_powspec_b200/slab.py). */
/* CUDA IPC: map a peer process's allocation into this process on the CURRENT device (handle = the
 * 64-byte cudaIpcMemHandle_t of the allocation's base); returns the base pointer. */
JPS_API int jps_ipc_open(const void* handle, void** out_ptr);
JPS_API int jps_ipc_close(void* ptr);
/* Let kernels on the current device dereference memory of `peer_device` (NVLink peer access). */
JPS_API int jps_enable_peer_access(int peer_device);
JPS_API int jps_slab_pack_p2p(jps_slab_plan_t* plan, const void* yz, void* const* peer_recv, void* stream);
/* Plane-range variants: transform / send the owned planes [x_begin, x_begin + x_count) only, so
 * that the 2-D FFT of one chunk overlaps the NVLink transfer of the previous one (two streams on
 * the host side).  x_count must be the whole slab or jps_slab_chunk_planes() (0 = no chunking). */
JPS_API int jps_slab_chunk_planes(jps_slab_plan_t* plan);
/* Layout of the transposed shard and form of the 2-D transform of the owned planes (default 0).
 * layout & JPS_SLAB_LAYOUT_XFAST == 0: shard [n][n/nranks][n/2+1], x slowest -- what jps_slab_pack + an all-to-all
 * produce; the 1-D FFT along x is strided.  JPS_SLAB_LAYOUT_XFAST: shard [n/nranks][n/2+1][n], x fastest -- produced
 * by jps_slab_pack_p2p[_planes], which then transposes 32x32 tiles on the way to the peers; the 1-D FFT is contiguous
 * and jps_slab_powspec_partial runs its lanes along kx.
 * With the x-fast layout the owned planes are transformed either by cuFFT's batched 2-D R2C plan (the planes leave
 * jps_slab_fft_yz* as [x_local][n][n/2+1]) or in "pencil form" -- C2C of length n/2 along z on the real lines read as
 * complex pairs, real-to-complex untangle fused with a transpose, contiguous C2C along y (the planes leave as
 * [x_local][n/2+1][n], and the peer-store kernel transposes x <-> y).  Default: pencil form when a rank's planes hold
 * 6 GB or more (measured crossover at 2048^3); JPS_SLAB_FFT_PENCIL / JPS_SLAB_FFT_CUFFT2D force one (even n only for
 * the pencil form).  Set before the first step. */
#define JPS_SLAB_LAYOUT_XFAST  1
#define JPS_SLAB_FFT_PENCIL    2
#define JPS_SLAB_FFT_CUFFT2D   4
JPS_API int jps_slab_set_layout(jps_slab_plan_t* plan, int layout);
JPS_API int jps_slab_fft_yz_planes(jps_slab_plan_t* plan, const float* slab, void* yz, int x_begin, int x_count,
                           void* stream);
JPS_API int jps_slab_pack_p2p_planes(jps_slab_plan_t* plan, const void* yz, void* const* peer_recv, int x_begin,
                             int x_count, void* stream);
JPS_API int jps_slab_fft_x(jps_slab_plan_t* plan, void* dk, void* stream);
/* This rank's partial sums of |delta_k|^2 L_l per user bin (sums[nb*3], float64, overwritten) and the
 * GLOBAL exact mode counts (counts[nb], identical on every rank; may be NULL).  dc: device pointer to
 * Re rho_hat(k=0) (lives on rank 0; broadcast it), used when normalise != 0. */
JPS_API int jps_slab_powspec_partial(jps_slab_plan_t* plan, const void* dk, const float* dc, int normalise,
                             float box_size, const float* k_edges, int nb, int mas_order,
                             double* sums, int64_t* counts, void* stream);
/* After the allreduce of sums: the reference's output arrays (src/correlations.py:49-54). */
JPS_API int jps_slab_powspec_finalize(jps_slab_plan_t* plan, float box_size, const float* k_edges, int nb,
                              const double* sums, const int64_t* counts, float shot_noise,
                              float* k3d, float* pk3d, float* nmodes, void* stream);

/* ------------------------------------------------------------------ gradients -------- */
/* Reverse-mode derivatives (the reference is differentiated with jax.value_and_grad, e.g.
 * tests/lognormal.py:99-107).  grad_pk: device float32 [nb*3] cotangent of Pk3D (NaN-free; empty bins
 * are ignored).  grad_mesh: device float32 [n,n,n] cotangent of the mesh passed to jps_powspec with
 * the same arguments. */
JPS_API int jps_powspec_grad(jps_plan_t* plan, const float* mesh, int normalise, float box_size,
                     const float* k_edges, int nb, int mas_order, const float* grad_pk,
                     float* grad_mesh, void* stream);
/* Cotangents of the particles of jps_paint: gx, gy, gz, gw device float32 [n_part] (any may be NULL).
 * Cell choice (int / floor of the grid coordinate) has zero derivative, as under JAX autodiff. */
JPS_API int jps_paint_grad(int n_mesh, const float* x, const float* y, const float* z, const float* w,
                   int64_t stride, int64_t n_part, float xmin, float ymin, float zmin, float box_size,
                   int order, int wrap, int compat, int variant, const float* grad_mesh,
                   float* gx, float* gy, float* gz, float* gw, void* stream);

/* ------------------------------------------------------------------ fused ------------ */
/* paint (into the plan-owned mesh, zeroed first) -> R2C FFT -> multipoles; the call the
 * benchmark times.  Arguments as jps_paint + jps_powspec(normalise=1, mas_order=order). */
JPS_API int jps_paint_powspec(jps_plan_t* plan,
                      const float* x, const float* y, const float* z, const float* w,
                      int64_t stride, int64_t n_part,
                      float xmin, float ymin, float zmin, float box_size,
                      int order, int wrap, int compat, int method,
                      const float* k_edges, int nb, float shot_noise,
                      float* mesh, void* paint_workspace, size_t paint_workspace_bytes,
                      float* k3d, float* pk3d, float* nmodes,
                      double* sums, int64_t* counts, void* stream);

/* Gradients of xi(s) and of the bispectrum with respect to the mesh (normalise = 0 semantics: the
 * caller differentiates delta = rho/mean - 1 itself, as the reference's scripts do under JAX,
 * tests/lognormal_bispec.py:71-106).  NaN cotangents (empty bins, Q11/Q22) count as zero.
 *   jps_xi_grad     : grad_xi [nb][3] (cotangent of xi3D of jps_xi / the composites) -> grad_mesh [N^3];
 *                     plan needs n_shell_fields >= 1
 *   jps_bispec_grad : grad_pk [nbins+2], grad_B [nbins] (cotangents of Pk and B of jps_bispec; fold
 *                     the cotangent of Q = B/(P0 P1 + P0 P3 + P1 P3) into them first) -> grad_mesh;
 *                     plan needs n_shell_fields >= 6.  7 FFTs whatever nbins. */
JPS_API int jps_xi_grad(jps_plan_t* plan, const float* mesh, float box_size, const float* s_edges /* host */,
                int nb, int mas_order, const float* grad_xi, float* grad_mesh, void* stream);
JPS_API int jps_bispec_grad(jps_plan_t* plan, const float* mesh, float box_size, float k1, float k2,
                    const float* theta /* host */, int nbins, int mas_order,
                    const float* grad_pk, const float* grad_B, float* grad_mesh, void* stream);

/* ------------------------------------------------------------------ catalogue text --- */
/* Whitespace-separated ASCII catalogue -> float32 rows on the device.
 * Replaces  np.loadtxt(path, usecols=(0,1,2), dtype=np.float32)  (tests/correlations.py:29,
 * tests/all_corr.py:26, tests/bispec.py:28) and the pd.read_csv(..., delim_whitespace=True)
 * variant (tests/positions.py:25), plus the box mask ((p < box) & (p > 0)).all(axis=1) the scripts
 * apply next (tests/correlations.py:30).  `text` is the raw file content in device memory
 * (16-byte aligned).  Fields are converted exactly as NumPy does: decimal -> nearest double ->
 * nearest float.  Lines that are blank or hold only a comment are skipped; `skiprows` leading
 * lines are ignored (pandas' header line: skiprows=1).
 *
 * Two calls, because the output size is data dependent:
 *   jps_text_count_lines : *n_lines (device int64) = number of text lines
 *   jps_text_parse       : out[n_rows][ncols] (capacity n_lines rows), rows in file order;
 *       counters (device int64[4]) = { n_rows, n_slow, n_bad, first_bad_line or -1 }
 *       filter != 0 keeps only rows with lo < value < hi in every requested column
 *       slow_rows (device int64[slow_capacity][2]) = (output row, byte offset of the line) of
 *       the rows holding a field outside the exact envelope of the device converter (more than
 *       19 significant digits, |decimal exponent| > 27, nan / inf): the host re-parses those
 *       fields (n_slow may exceed slow_capacity: then call again with a larger list)
 *       n_bad counts rows with a non-numeric or missing requested column (np.loadtxt raises). */
JPS_API size_t jps_text_workspace_bytes(int64_t nbytes, int64_t n_lines, int ncols);
JPS_API int jps_text_count_lines(const char* text, int64_t nbytes, int64_t* n_lines,
                         void* workspace, size_t workspace_bytes, void* stream);
JPS_API int jps_text_parse(const char* text, int64_t nbytes, int64_t n_lines, int skiprows, int comment,
                   const int* usecols /* host */, int ncols, int filter, float lo, float hi,
                   float* out, int64_t* counters, int64_t* slow_rows, int64_t slow_capacity,
                   void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ mock generator --- */
/* the reference's input pipeline (tests/create_lognormal.py:44-55) on the device.
 *
 * jps_mock_gaussian_field replaces gaussian_field(grid, kf, Pkf, Rayleigh_sampling, seed, BoxSize)
 * (src/gauss_field.py:5-80): delta_k (device complex64 [n][n][n/2+1]) with amplitude
 * sqrt(P(|k|) (n^2/box)^3) (times sqrt(-log u) when rayleigh != 0), uniform random phase, the
 * reference's Hermitian pairing on the kz = 0 and kz = n/2 planes and delta_k[0,0,0] = 0.  P(|k|) is
 * interpolated linearly in the HOST table (kf, pkf)[nk] exactly as the reference's bisection does
 * (extrapolating along the end segments).  The random stream is Philox4x32-10 with one counter per
 * mode -- NOT the reference's sequential Mersenne-Twister draws, which no parallel generator can
 * follow -- so fields agree with the reference in distribution, not draw by draw.
 *
 * jps_mock_populate_count / _fill replace populate_field(rho, n_bins, box_size, density, key)
 * (src/populate_field.py:11-29, NumPy twin src/gauss_field.py:90-110): Poisson counts with mean
 * rho * (box/n)^3 * density / mean(rho) per cell, every particle at its cell centre plus the
 * triangular offset sign(r)(1 - sqrt|r|) * bin_size per axis, wrapped into [0, box).  lognormal != 0
 * samples exp(bias * rho) instead (rho = the Gaussian field in real space; tests/create_lognormal.py:49).
 * Two calls because the particle number is data dependent: _count leaves the per-cell counts and
 * offsets in the workspace and the total in *total (device int64); after reading it the caller
 * allocates pos[total][3] (device float32) and calls _fill with the SAME workspace and seed.
 * Particles come out grouped by cell in C order. */
JPS_API size_t jps_mock_field_workspace_bytes(int nk);
JPS_API int jps_mock_gaussian_field(int n_mesh, const double* kf /* host */, const double* pkf /* host */, int nk,
                            int rayleigh, unsigned long long seed, float box_size, void* delta_k,
                            void* workspace, size_t workspace_bytes, void* stream);
JPS_API size_t jps_mock_populate_workspace_bytes(int n_mesh);
/* byte offset, inside the populate workspace, of the per-cell counts (uint32 [n][n][n]) _count leaves there */
JPS_API size_t jps_mock_populate_counts_offset(int n_mesh);
JPS_API int jps_mock_populate_count(const float* rho, int n_mesh, float box_size, float density, int lognormal,
                            float bias, unsigned long long seed, void* workspace, size_t workspace_bytes,
                            int64_t* total, void* stream);
JPS_API int jps_mock_populate_fill(int n_mesh, float box_size, unsigned long long seed, const void* workspace,
                           size_t workspace_bytes, int64_t n_out, float* pos, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* JPS_H_ */
