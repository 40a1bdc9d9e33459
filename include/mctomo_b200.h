/*
 * mctomo_b200.h -- C ABI of the B200-native surface-wave forward-modelling hot path of MCTomo.
 *
 * This is the drop-in boundary: plain C, raw pointers and sizes, arrays in the Fortran
 * (column-major) layout the reference already uses across its C++ boundary
 * (src/cgal_delaunay.cpp:470-478; src/fastMarching_wrapper.f90:18-29).  The Fortran host
 * reaches it through ISO_C_BINDING interfaces (fortran/mctomo_b200_shim.f90);
 * the bodies it replaces are cited per function (paths relative to the reference tree).
 *
 * Conventions
 *   - 3-D model arrays are (nz,ny,nx) column-major: element (k,j,i), 1-based, lives at
 *     [((i-1)*ny + (j-1))*nz + (k-1)].
 *   - Column windows ix0..ix1, iy0..iy1 are 1-based and inclusive, as in the Fortran.
 *   - Every function returns an int status: 0 = OK, > 0 = data condition (see MCT_E_*),
 *     < 0 = CUDA/runtime failure (text via mct_last_error()).  The library never exits.
 *   - Host-pointer entry points are synchronous: on return all outputs are in host memory.
 *     *_dev entry points take device pointers and enqueue on `stream` (a cudaStream_t cast to
 *     void*; NULL = the library's own stream) without synchronising.
 *   - The caller owns every array; the library keeps no host pointer after returning.
 *   - There is no CPU fallback: without a usable CUDA device mct_init fails and every other
 *     call returns MCT_E_NOINIT.
 */
#ifndef MCTOMO_B200_H
#define MCTOMO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCT_OK 0
#define MCT_E_INVALID_ARG 1    /* bad sizes / NULL pointers */
#define MCT_E_GRT_NEEDED 2     /* >= 1 column has a low-velocity layer (surfmodes.f90:84-87,96-99): ierr=2 */
#define MCT_E_TOO_MANY_LAYERS 3 /* a column needs more than MCT_MAX_LAYERS layers: ierr=3 */
#define MCT_E_DEGENERATE_NUCLEI 4 /* > 13 nuclei coincident in ALL coordinates: the reference kd-tree build never terminates
                                     (nuclei that merely share one or two coordinates are fine: kdtree2.f90:818-826) */
#define MCT_E_FLUID_BELOW_TOP 5 /* vs ~ 0 below the first layer: the reference `stop`s (surfmodes.f90:342-345): ierr=4 */
#define MCT_E_ZERO_NOISE 6 /* misfit: a ray that carries data has sigma < 1e-10 (likelihood_surf.F90:387-390 raises an error) */
#define MCT_E_FM2D_STALE 7 /* fm2d: a source lies in the model's last cell row/column.  The reference's refined march stops at once
                             there (fm2d_ttime.f90:76-87 compares a refined extent with a coarse index) and returns the PREVIOUS
                             source's field; outputs are delivered, that source's times are those of the dead march (unreached
                             nodes read 0) and the per-problem status (d_err) is 6 */
#define MCT_E_NOINIT (-1)
#define MCT_E_CUDA (-2)

#define MCT_MAX_LAYERS 200 /* NL of surfdisp96.f:57 */
#define MCT_MAX_PERIODS 60 /* NP of surfdisp96.f:59 */

/* Mirrors the fields of T_GRID (src/settings.f90:20-28) the hot path reads. */
typedef struct {
  int32_t nx, ny, nz;
  double xmin, ymin, zmin;
  double dx, dy, dz;
  double waterDepth; /* grid%waterDepth */
  double scaling;    /* grid%scaling (thick = thick/scaling, likelihood_surf.F90:615) */
} mct_grid;

/* Options of the dispersion call; mirrors T_MODES_PARA (surfmodes.f90:23-30) plus the two
 * constants in which the reference's two layering routines differ. */
typedef struct {
  int32_t raylov;     /* paras%modetype: 1 Rayleigh, 0 Love */
  int32_t phaseGroup; /* 0 phase only, 1 phase + group (surfdisp96 igr) */
  int32_t nmodes;     /* <= 0: surfmodes -> surfdisp96 (fundamental, outputs preset to 100.0)
                         >= 1: surfmmodes -> surfdisp_mmodes with that many modes (outputs preset to 0) */
  int32_t check_scope; /* columns check_model scans: 0 = the whole grid (the reference, likelihood_surf.F90:631-646);
                          1 = only the window's columns.  In the sampler the model outside the perturbed box is
                          the previously accepted -- hence valid -- model, so both give the same answer there,
                          and scope 1 spares the host-pointer call the upload of the whole vs grid. */
  double dphase;      /* paras%dc = settings%dPhaseVel */
  double layer_eps;   /* new layer when |vs(k)-vs_run| > layer_eps: (double)1e-10f in
                         likelihood_surf.F90:37,560; (double)1e-5f in forward_modelling.f90:27 */
  double water_thresh; /* water layer iff waterDepth > water_thresh: (double)1e-10f
                          (likelihood_surf.F90:546) or 0 (forward_modelling.f90:91) */
  double preset;      /* value left in pvel/gvel for columns the solver never ran on (ierr >= 2):
                         100.0 (likelihood_surf.F90:188-189) or 1000.0 (forward_modelling.f90:410-411) */
} mct_disp_opts;

/* Counters accumulated since mct_init / mct_reset_stats. */
typedef struct {
  /* REPRESENTED work: what the reference's own loops count for the same input (every column solved on its own).
   * These equal the oracle's counters -- an identical search path is itself a parity check. */
  int64_t n_dltar;       /* secular-function evaluations (surfdisp96.f:1036) */
  int64_t n_layer_steps; /* layer steps inside them (loops surfdisp96.f:1078,1159) */
  int64_t n_columns;     /* columns solved */
  int64_t n_nodes;       /* grid nodes assigned by the nearest-nucleus kernel */
  int64_t n_launches;    /* kernels launched by this library */
  /* EXECUTED work: what the dispersion kernel really ran.  Smaller than the represented figures when bit-identical
   * layer stacks were folded (mct_set_dedup); the roofline is computed from these. */
  int64_t n_dltar_executed;
  int64_t n_layer_steps_executed;
  int64_t n_columns_solved;
} mct_stats;

/* What the last dispersion launch did (the bench's roofline block reports the kernel from here). */
typedef struct {
  char kernel[48];          /* __global__ function launched, e.g. "k2_dispersion_fast_r128" */
  int32_t columns;          /* columns of the call */
  int32_t columns_solved;   /* distinct layer stacks actually solved; -1 = not read back (calls below 8192 columns
                               stay asynchronous) */
  int32_t lanes_per_column; /* 1 = one thread per column, else the cooperative group size */
  int32_t sm_count;
} mct_launch_info;

/* ---- lifetime ------------------------------------------------------------------------ */
int mct_init(int device);       /* selects the device, creates the stream; idempotent */
int mct_shutdown(void);         /* frees every device/pinned buffer */
const char* mct_last_error(void);
int mct_get_stats(mct_stats* out);
int mct_reset_stats(void);
/* Enables (1) / disables (0) the per-column work counters inside the dispersion kernel.
 * Default on; they cost two integer adds per layer step. */
int mct_set_counters(int on);

/* ---- stage 1: Voronoi -> grid ---------------------------------------------------------
 * Replaces the body of kdtree_to_grid (src/mcmc_loc2.f90:2029-2078): kdtree2_create over
 * points(:,1:ncells), the OpenMP loop of kdtree2_n_nearest(nn=1) over the nodes inside
 * bnd_box, and the writes of sites_id / vp / vs / rho.  Results are identical to kdtree2's,
 * including which nucleus wins an exact distance tie.
 *   points  (3,ncells)  RTI%points     params (3,ncells) RTI%parameters = (vp,vs,rho)
 *   box     {x0,y0,z0,x1,y1,z1}        bnd_box(1:2)
 *   pm      NULL, or {vp,vs,rho}: only nodes with |vs-pm.vs|<1e-8 and |vp-pm.vp|<1e-8 are
 *           reassigned (mcmc_loc2.f90:2055-2062)
 *   vp,vs,rho,sites_id  (nz,ny,nx) host arrays, updated in place inside the window only.
 */
int mct_voronoi_to_grid(const double* points, const double* params, int ncells, const mct_grid* g,
                        const double box[6], const double* pm, double* vp, double* vs, double* rho,
                        int32_t* sites_id);

/* Device-pointer form: nuclei are still host arrays (48 B each; the tree is built on the
 * host and uploaded), the model arrays are device pointers. */
int mct_voronoi_to_grid_dev(const double* points, const double* params, int ncells, const mct_grid* g,
                            const double box[6], const double* pm, double* d_vp, double* d_vs,
                            double* d_rho, int32_t* d_sites_id, void* stream);

/* The *_dev entry points that run the nearest-nucleus kernel (mct_voronoi_to_grid_dev, mct_forward_eval_dev,
 * mct_forward_batch_dev) cannot report a device-side condition without synchronising.  mct_k1_status synchronises
 * `stream` and returns MCT_E_CUDA if the last such call on it overflowed the traversal stack of the tie replay
 * (a tree deeper than 64 levels: never seen with kdtree2's mean splits), MCT_OK otherwise. */
int mct_k1_status(void* stream);

/* Index window of a box, exactly as mcmc_loc2.f90:2034-2045 computes it (1-based, clamped):
 * w = {ix0,ix1,iy0,iy1,iz0,iz1}. */
int mct_box_window(const mct_grid* g, const double box[6], int32_t w[6]);

/* ---- datatype-2 property maps -----------------------------------------------------------
 * vs2vp_3d + vp2rho_3d over n values (src/utils.f90:102-112,125-134; called from
 * src/likelihood.f90:75-76): vp = vs*1.730, rho = 1.74*vp**0.25. */
int mct_vs2vp_rho(const double* vs, double* vp, double* rho, int64_t n);
int mct_vs2vp_rho_dev(const double* d_vs, double* d_vp, double* d_rho, int64_t n, void* stream);
/* Same maps restricted to the nodes of an index window w = {ix0,ix1,iy0,iy1,iz0,iz1} (1-based inclusive, e.g. from
 * mct_box_window) of (nz,ny,nx) host arrays.  After a kdtree_to_grid call on a box only those nodes of vs have
 * changed, so this leaves vp and rho exactly as the reference's whole-grid recomputation does. */
int mct_vs2vp_rho_window(const double* vs, double* vp, double* rho, const mct_grid* g, const int32_t w[6]);

/* ---- stage 2: per-column modal dispersion ------------------------------------------------
 * Replaces surf_likelihood's block likelihood_surf.F90:161-206 (check_model,
 * convert_to_layer, the OpenMP loop over surfmodes) and the twin in
 * forward_modelling.f90:393-429.
 *   vp,vs,rho   (nz,ny,nx) host arrays (whole grid: check_model scans all of vs)
 *   ix0..iy1    clamped 1-based inclusive column window (likelihood_surf.F90:155-169)
 *   freqs       np frequencies in Hz (dat%freqs; periods = 1/freqs must ascend)
 *   pvel, gvel  (np*nm, iy0:iy1, ix0:ix1), nm = max(nmodes,1); mode-major inside a column:
 *               index ifreq + (imode-1)*np (surfmodes.f90:179-180)
 *   ierr        (iy0:iy1, ix0:ix1): 0/1 as surfdisp96 sets it; 2/3/4 see MCT_E_*
 *   model_invalid  out: 1 if check_model(vs) is .true. (likelihood_surf.F90:161-164,631-646);
 *               the reference then returns before solving anything, and so does this call
 *               (outputs untouched).  Pass NULL to skip the check (forward_modelling.f90 has none).
 * Returns MCT_OK, or the largest per-column condition code >= 2 that occurred.
 */
int mct_surf_dispersion(const double* vp, const double* vs, const double* rho, const mct_grid* g, int ix0,
                        int ix1, int iy0, int iy1, const double* freqs, int np, const mct_disp_opts* opt,
                        double* pvel, double* gvel, int32_t* ierr, int32_t* model_invalid);

int mct_surf_dispersion_dev(const double* d_vp, const double* d_vs, const double* d_rho, const mct_grid* g,
                            int ix0, int ix1, int iy0, int iy1, const double* freqs, int np,
                            const mct_disp_opts* opt, double* d_pvel, double* d_gvel, int32_t* d_ierr,
                            int32_t* d_flags /* [0]=model_invalid, [1]=max condition code; may be NULL */,
                            void* stream);

/* surfmodes / surfmmodes for columns that are already layered (surfmodes.f90:39-49,110-120):
 * column c owns layers offsets[c] .. offsets[c+1]-1 of thick/vp/vs/rho (top to bottom, last
 * one the half-space).  phase, group: (np*nm, ncol); ierr: (ncol). */
int mct_surfmodes_batch(const double* thick, const double* vp, const double* vs, const double* rho,
                        const int64_t* offsets, int ncol, const double* freqs, int np,
                        const mct_disp_opts* opt, double* phase, double* group, int32_t* ierr);

/* ---- fused forward evaluation --------------------------------------------------------------
 * Stage 1 over the whole grid, the datatype-2 property maps, check_model and stage 2 over all
 * columns, with the gridded model staying in HBM (the sequence kdtree_to_grid -> likelihood ->
 * surf_likelihood of mcmc_loc2.f90:139-143 for a full bounding box).  Only the nuclei go in and
 * only the dispersion maps (+ optionally the gridded model) come out.
 *   derive_vp_rho  1: datatype 2 (vp, rho recomputed from vs over the grid); 0: keep params' vp, rho
 *   vp,vs,rho,sites_id  optional host outputs (NULL = leave on device)
 */
int mct_forward_eval(const double* points, const double* params, int ncells, const mct_grid* g,
                     int derive_vp_rho, const double* freqs, int np, const mct_disp_opts* opt, double* pvel,
                     double* gvel, int32_t* ierr, int32_t* model_invalid, double* vp, double* vs,
                     double* rho, int32_t* sites_id);

/* Same with device outputs (d_pvel/d_gvel (np*nm,ny,nx), d_ierr (ny,nx), d_flags[2]) and an
 * x-slab restriction ixs0..ixs1 (1-based inclusive; whole grid = 1..nx) used to shard the
 * columns of one chain across GPUs: only that slab is gridded and solved, outputs are indexed
 * relative to the slab.  d_model = {d_vp,d_vs,d_rho,d_sites} whole-grid device arrays. */
int mct_forward_eval_dev(const double* points, const double* params, int ncells, const mct_grid* g,
                         int derive_vp_rho, int ixs0, int ixs1, const double* freqs, int np,
                         const mct_disp_opts* opt, double* d_vp, double* d_vs, double* d_rho,
                         int32_t* d_sites_id, double* d_pvel, double* d_gvel, int32_t* d_ierr,
                         int32_t* d_flags, void* stream);

/* ---- batches of independent models (chains) ---------------------------------------------------
 * MCTomo runs one independent chain per MPI rank (src/MCTomo.F90:82-86,131-133).  Several chains that
 * share a GPU are evaluated as ONE batch: model b owns nuclei offsets[b]..offsets[b+1]-1; every model
 * array gains a slowest axis of length nb ((nz,ny,nx,nb), outputs (np*nm,ny,nx,nb), ierr (ny,nx,nb),
 * flags int32[2*nb] = {model_invalid, max condition code} per model).  mct_set_nuclei_batch builds the
 * nb search trees on the host and makes them resident; mct_forward_batch_dev then runs every kernel of
 * the forward evaluation on the resident set without touching host memory. */
int mct_set_nuclei_batch(const double* points, const double* params, const int64_t* offsets, int nb);
int mct_forward_batch_dev(const mct_grid* g, int nb, int derive_vp_rho, int ixs0, int ixs1, const double* freqs,
                          int np, const mct_disp_opts* opt, double* d_vp, double* d_vs, double* d_rho,
                          int32_t* d_sites_id, double* d_pvel, double* d_gvel, int32_t* d_ierr, int32_t* d_flags,
                          void* stream);
/* Host-pointer form: nuclei in, maps out;
 * model_invalid int32[nb]; vp/vs/rho/sites_id optional (NULL = stay on the device). */
int mct_forward_eval_batch(const double* points, const double* params, const int64_t* offsets, int nb,
                           const mct_grid* g, int derive_vp_rho, const double* freqs, int np,
                           const mct_disp_opts* opt, double* pvel, double* gvel, int32_t* ierr,
                           int32_t* model_invalid, double* vp, double* vs, double* rho, int32_t* sites_id);

/* ---- multi-GPU: one chain's columns sharded over several GPUs (BASELINE config 5) ------------------
 * MCTomo runs one MPI rank per chain (src/MCTomo.F90:82-86,131-133) and chains never talk on the data path: chains ->
 * GPUs needs nothing from this section.  For ONE chain on a large grid the x axis is cut into contiguous slabs, one
 * per rank (process per GPU); nuclei are replicated (48 B each), every rank grids and solves its slab, and the
 * dispersion maps are all-gathered IN PLACE with NCCL so that every rank holds the full (np,ny,nx) field for fm2d
 * (src/likelihood_surf.F90:295-336); check_model's whole-grid `any` (:631-646) is a MAX all-reduce of one flag.
 * NCCL is bound at run time (dlopen libnccl.so.2; MCT_NCCL_LIB overrides): no link-time dependency.
 *   rank 0:  mct_comm_unique_id(id)  ->  host broadcasts the 128 bytes (MPI_Bcast)  ->  all: mct_comm_init(id, rank, n)
 */
#define MCT_COMM_ID_BYTES 128
int mct_comm_unique_id(void* id128);
int mct_comm_init(const void* id128, int rank, int nranks); /* after mct_init(device); collective over the ranks */
int mct_comm_destroy(void);
int mct_comm_info(int* rank, int* nranks, int* nccl_version); /* MCT_E_INVALID_ARG when no communicator exists */
/* How mct_forward_sharded_dev divides one chain's columns.  0 (default): contiguous x-slabs, as SURVEY.md 8(e) specifies.
 * 1: balanced -- every rank grids, layers and de-duplicates the WHOLE model (milliseconds), solves every nranks-th entry
 * of the sorted list of distinct columns, and the compact results are all-gathered in place; robust against models whose
 * structure -- hence cost -- varies along x.  In mode 1 the outputs are the plain (nout, ny, nx) maps / (ny, nx) ierr. */
int mct_comm_set_mode(int mode);
/* x-slab of `rank`: 1-based inclusive ix0..ix1, equal width per = ceil(nx/nranks); trailing slabs may be short or
 * empty (ix1 < ix0). */
int mct_slab_bounds(int nx, int nranks, int rank, int* ix0, int* ix1, int* per);
/* In-place all-gather of a device buffer of nranks chunks of bytes_per_rank bytes (this rank owns the rank-th). */
int mct_allgather_inplace(void* d_buf, int64_t bytes_per_rank, void* stream);
/* MAX all-reduce of n int32 flags in place ({model_invalid, max condition code}). */
int mct_allreduce_flags(int32_t* d_flags, int n, void* stream);
/* The sharded forward evaluation: K1 + property maps + check_model + dispersion on this rank's slab of the resident
 * nuclei set (mct_set_nuclei_batch, nb = 1, same nuclei on every rank), outputs written into this rank's chunk of the
 * FULL maps d_pvel/d_gvel (nout, ny, per*nranks), d_ierr (ny, per*nranks); then in-place all-gather of pvel, ierr
 * (and gvel when opt->phaseGroup == 1) and MAX all-reduce of d_flags[2], all enqueued on `stream`.  With one rank (no
 * communicator) it is mct_forward_batch_dev over the whole grid. */
int mct_forward_sharded_dev(const mct_grid* g, int derive_vp_rho, const double* freqs, int np, const mct_disp_opts* opt,
                            double* d_vp, double* d_vs, double* d_rho, int32_t* d_sites_id, double* d_pvel,
                            double* d_gvel, int32_t* d_ierr, int32_t* d_flags, void* stream);
/* Device milliseconds the collectives of the last mct_forward_sharded_dev call took (synchronises). */
int mct_comm_last_ms(double* ms);

/* ---- the 2-D product and point location ---------------------------------------------------------------
 * kdtree_to_grid of the 2-D variant (mcmc2d/mcmc.f90:1469-1526): nuclei points2 (2,ncells), every node of the (ny,nx)
 * grid (element (j,i) at [(i-1)*ny + (j-1)]), query point (xmin+(i-1)dx, ymin+(j-1)dy).  Same cells as kdtree2 run with
 * two dimensions, ties included. */
int mct_voronoi_to_grid_2d(const double* points2, const double* params, int ncells, int nx, int ny, double xmin,
                           double ymin, double dx, double dy, double* vp, double* vs, double* rho, int32_t* sites_id);
/* kdtree_locate (mcmc2d/mcmc.f90:1528-1551) for a batch: the nucleus (1-based) nearest to each of nq query points;
 * points (dim,ncells), queries (dim,nq), dim = 2 or 3. */
int mct_nearest_nucleus(const double* points, int dim, int ncells, const double* queries, int64_t nq, int32_t* idx);
/* sites_locate (src/likelihood_body.F90:799-831) for a batch of 3-D points: the cell index of the node below the point
 * (point2idx, :1038-1054) when its eight neighbours agree, else kdtree2's nearest nucleus.  sites_id (nz,ny,nx): host
 * array, or device array for the _dev form (e.g. a session's resident cell map). */
int mct_sites_locate(const double* points, int ncells, const int32_t* sites_id, const mct_grid* g, const double* queries,
                     int64_t nq, int32_t* idx);
int mct_sites_locate_dev(const double* points, int ncells, const int32_t* d_sites_id, const mct_grid* g,
                         const double* queries, int64_t nq, int32_t* idx);

/* ---- measurement helpers ----------------------------------------------------------------------
 * mct_set_profiling(1) brackets every kernel launch with CUDA events on the launching stream;
 * mct_kernel_times returns accumulated milliseconds {K1 nearest-nucleus, K2 dispersion, other kernels,
 * number of timed launches} (and zeroes them when reset != 0).  mct_fp64_peak_probe measures this
 * GPU's FP64 pipe with 8 independent chains per thread: DFMA (2 flop/instr) and DMUL+DADD. */
int mct_set_profiling(int on);
int mct_kernel_times(double ms[4], int reset);
int mct_fp64_peak_probe(double* tflops_fma, double* tflops_mul_add);
/* Posterior accumulation of `program sample` (src/sample.f90:481-487), on the device: after a full-grid
 * mct_voronoi_to_grid_dev of a kept sample, aveS += vs, stdS += vs**2, aveP += vp, stdP += vp**2 elementwise
 * over n = nx*ny*nz values (same operations, same order: bit-identical sums).  Keeps the thousands of regrids
 * the post-processor issues resident in HBM; only the four accumulators are read back at the end. */
int mct_accumulate_stats_dev(const double* d_vs, const double* d_vp, double* d_aveS, double* d_stdS, double* d_aveP,
                             double* d_stdP, int64_t n, void* stream);

/* Shape of the nearest-nucleus kernel.  mode 0 (default): a block per tile of 4 x 16 columns; conservative culls leave,
 * per column and per z-segment of 8 nodes, a list of at most 7 candidate nuclei in shared memory; nodes are resolved
 * two at a time and written with 16-byte stores; kdtree2's traversal is replayed only for (near-)tied nodes.  mode 2:
 * the round-1 shape (a warp per column, every node scans its column's survivors).  mode 1: kdtree2's traversal for
 * every node.  Results are identical in all three. */
/* ---- 2-D fast-marching travel times ----------------------------------------------------------------
 * Replaces `modrays` as surf_likelihood calls it for phase-velocity data (src/likelihood_surf.F90:295-336 with uar = 1:
 * travel times at the receivers, no ray geometry): gridder, bsplrefine, the source loop with source-grid refinement,
 * srtimes (fm2d/fm2dray_cartesian.f90:67-478,490-668,676-770) and travel / fouds1 / fouds2 / the narrow-band heap
 * (fm2d/fm2d_ttime.f90).  All np periods x nsrc sources are independent problems and go out in ONE launch.
 * Options mirror settings%gridx, gridy, sgref, sgdic, sgext, order, band (likelihood_surf.F90:247-253). */
typedef struct {
  int32_t gridx, gridy; /* dicing of the propagation grid */
  int32_t sgref;        /* source-grid refinement on (1) / off (0) */
  int32_t sgdic, sgext; /* its dicing level and extent */
  int32_t order;        /* 0 first-order, 1 mixed-order stencils */
  double band;          /* narrow band size as a fraction of nx*ny */
} mct_fm2d_opts;
/* Host form.  src/rcv coordinates: x along the FIRST grid axis (grid%x), z along the second (grid%y).
 *   srs  (nrc, nsrc, nmaps) int32: 1 where the pair carries data (dat%raystat(:,1,period));
 *   vel  (nvz+2, nvx+2, nmaps): like%vel(period,:,:) with its replicated edge, one period after the other;
 *   ttime (nrc, nsrc, nmaps) in/out: entries without data are left untouched (like%phaseTime);
 *   field optional (nnz, nnx, nsrc, nmaps): the travel-time field of every problem (like%field4d), or NULL.
 * nvx = grid%nx, nvz = grid%ny, gox/goz = xmin/ymin, dvx/dvz = dx/dy. */
int mct_fm2d_times(const double* src_x, const double* src_z, int nsrc, const double* rcv_x, const double* rcv_z, int nrc,
                   const int32_t* srs, const double* vel, int nmaps, int nvx, int nvz, double gox, double goz, double dvx, double dvz,
                   const mct_fm2d_opts* o, double* ttime, double* field);
/* The same with the ray geometry (uar = 0, group-velocity data): rpaths (fm2d/fm2dray_cartesian.f90:773-1456, cfd = 0).
 *   srsv (nrc, nsrc, nmaps) int32: dat%raystat(:,2,period), the 1-based slot in which the Fortran stores the pair's ray;
 *   ray_npts (nrc*nsrc, nmaps): rays(slot)%npoints (0 for a pair without data);
 *   ray_pts (2, ray_cap, nrc*nsrc, nmaps): rays(slot)%points, receiver first, source last;
 *   ray_len (nrc*nsrc, nmaps): rays(slot)%length() (like%srdist for group-velocity data);
 *   crazy (nmaps): crazyray of every period (> 0: surf_likelihood returns like = huge).
 * A slot holds ray_cap points (8*(nnx+nnz) is ample); a ray still wandering after that many steps is counted as crazy
 * (the Fortran only gives up after nnx*nnz points). */
int mct_fm2d_rays(const double* src_x, const double* src_z, int nsrc, const double* rcv_x, const double* rcv_z, int nrc,
                  const int32_t* srs, const int32_t* srsv, const double* vel, int nmaps, int nvx, int nvz, double gox, double goz, double dvx,
                  double dvz, const mct_fm2d_opts* o, double* ttime, int ray_cap, int32_t* ray_npts, double* ray_pts, double* ray_len,
                  int32_t* crazy);
/* Device form, asynchronous on `stream`.  d_src_xz = [x(nsrc) | z(nsrc)], d_rcv_xz = [x(nrc) | z(nrc)].  The velocity
 * maps are addressed as d_vel[m*vel_map_stride + (b*(nvz+2) + a)*vel_elem_stride], so the padded map that
 * mct_assemble_vel_dev builds -- the Fortran's like%vel(np, ny+2, nx+2) -- is consumed in place with
 * (vel_elem_stride, vel_map_stride) = (np, 1); srs_map_stride = nrc*nsrc, or 2*nrc*nsrc for dat%raystat(nrev*nsrc, 2, np).
 * d_err int32[nmaps*nsrc]: 0, 1 source outside the model, 2 narrow band overflow, 3 receiver outside the model, 6 see
 * MCT_E_FM2D_STALE. */
int mct_fm2d_times_dev(const double* d_src_xz, int nsrc, const double* d_rcv_xz, int nrc, const int32_t* d_srs, long long srs_map_stride,
                       const double* d_vel, long long vel_elem_stride, long long vel_map_stride, int nmaps, int nvx, int nvz, double gox,
                       double goz, double dvx, double dvz, const mct_fm2d_opts* o, double* d_ttime, int32_t* d_err, void* stream);
/* out2 = {nodes accepted, stencil updates} since mct_reset_stats, as the reference's own loops count them */
int mct_fm2d_stats(int64_t out2[2]);

int mct_set_k1_mode(int mode);

/* ---- low-velocity columns: the generalized reflection/transmission branch of surfmodes ------------------------------
 * Replaces RayleighModes / LoveModes (surfmodes/surfmodes.f90:84-87,96-99,185-306) with SearchRayleigh / SearchLove
 * (allmodes = 0), C_Interval[_L], the secular functions of Rayleigh.f90 / Love.f90 and bisecim (util.f90).
 * enable = 1: columns the nlvls1 predicate sends to that branch are solved on the device after the surfdisp96 kernel
 * (ierr 0/1 as RayleighModes / LoveModes set it; entries the Fortran never assigns keep opt->preset) instead of being
 * reported as ierr = 2 / MCT_E_GRT_NEEDED.  Applies to nmodes <= 0 only (surfmmodes prints "not supported yet").
 * par6 = {tolmin, tolmax, smin_min, smin_max, dcm, dc2} of T_MODES_PARA (surfmodes.f90:23-30); NULL keeps the current
 * values (default: likelihood_surf.F90:175-182 with settings%tol = 1e-6).  paras%dc is opt->dphase.  Default: off. */
int mct_set_grt(int enable, const double* par6);
/* out3 = {columns the last dispersion call solved on that branch, secular-function evaluations and interface steps spent
 * there since mct_reset_stats (as the reference's own loops would count them)} */
int mct_grt_stats(int64_t out3[3]);
/* Shape of the dispersion kernel.  mode 0 (default): batches of fewer than coop_max_columns columns (default 0 =
 * 1.7x the resident lanes of the GPU, 128 819 columns on a B200; pass -1 to keep) give every column a GROUP OF G LANES that
 * split getsol's bracketing scan: G = 256, 128 or 64 (a block of 8, 4 or 2 warps per column, while that keeps the
 * launch within the GPU's resident warps: proposal-sized calls), 32 (one warp), down to 2 -- the largest power of two
 * that keeps the launch within about two waves of resident lanes.  Larger batches run one THREAD per column.  mode 1 / 2 force
 * one or the other; mct_set_k2_lanes fixes G (0 = automatic, else a power of two from 2 to 256).  Results are
 * identical in every shape. */
int mct_set_k2_mode(int mode, int coop_max_columns);
/* Exact de-duplication of columns in front of the dispersion kernel (default on): columns whose float32 layer stacks
 * are bit-identical (same model of a batch) are solved once and the outputs copied.  Results are identical either way;
 * calls of 8192 columns or more read the distinct count back (one stream synchronisation) to size the launch. */
int mct_set_dedup(int on);
int mct_last_launch(mct_launch_info* out);
int mct_set_k2_lanes(int lanes_per_column);
/* Device self-test: the shared-reciprocal division the dispersion kernel uses is compared, bit for
 * bit, with the compiler's IEEE division on *tested random operand pairs whose exponents are drawn
 * from [-emax, emax]; *mismatches must come back 0. */
int mct_selftest_division(int emax, int64_t* tested, int64_t* mismatches);

/* Map assembly of likelihood_surf.F90:259-264 on the device: scatter a window of pvel into the
 * padded (np, ny+2, nx+2) field and replicate the edges the window touches. */
int mct_assemble_vel_dev(const double* d_pvel, int np, int nx, int ny, int ix0, int ix1, int iy0, int iy1,
                         double* d_vel, void* stream);

/* ---- resident session: one chain's model kept in HBM between proposals -------------------------------------
 * Replaces, for a sampler that adopts it, the per-iteration sequence of src/mcmc_loc2.f90:199-228,556-566 +
 * src/likelihood.f90:75-83 + src/likelihood_surf.F90:155-231 -- backup, kdtree_to_grid(box), vs2vp_3d/vp2rho_3d,
 * check_model (whole grid), dispersion of the box's columns + one-column halo, keep or restore -- with calls that
 * move only the nuclei (48 B each) in and the window's dispersion maps out.  Same results as the host-pointer
 * entry points (tests/test_gpu_session.py).  Not thread-safe per session; one pending proposal at a time. */
typedef struct mct_session mct_session;
int mct_session_create(const mct_grid* g, const double* freqs, int np, const mct_disp_opts* opt, int derive_vp_rho,
                       mct_session** out);
int mct_session_destroy(mct_session* s);
/* Full evaluation of a nuclei set (every node, every column); it becomes the current model.  Host outputs
 * pvel/gvel (nout,ny,nx), ierr (ny,nx), model_invalid are optional (NULL: results stay on the device). */
int mct_session_set_model(mct_session* s, const double* points, const double* params, int ncells, double* pvel,
                          double* gvel, int32_t* ierr, int32_t* model_invalid);
/* One proposal: nuclei AFTER the move, the box handed to kdtree_to_grid, pm = NULL or the moved cell's
 * (vp,vs,rho) for a value-only move.  Out: win = {ix0,ix1,iy0,iy1} (1-based; box columns + halo) and that
 * window's maps packed (nout,wy,wx) / (wy,wx); the caller provides room for nx*ny columns.  When check_model
 * rejects the model (*model_invalid = 1) nothing is solved and the map buffers are not meaningful.  The call
 * commits nothing: follow it with mct_session_accept or mct_session_reject. */
int mct_session_propose(mct_session* s, const double* points, const double* params, int ncells, const double box[6],
                        const double* pm, int32_t win[4], double* pvel_win, double* gvel_win, int32_t* ierr_win,
                        int32_t* model_invalid);
int mct_session_accept(mct_session* s);
int mct_session_reject(mct_session* s);
/* Read the resident arrays back (checkpoint/restart, tests); any pointer may be NULL. */
int mct_session_get_model(mct_session* s, double* vp, double* vs, double* rho, int32_t* sites_id);
int mct_session_get_maps(mct_session* s, double* pvel, double* gvel, int32_t* ierr);

/* CalGroupTime / GetVelocity (src/likelihood_surf.F90:454-521) on a DEVICE (np,ny,nx) velocity map: travel time of
 * every ray through the map of its period (bilinear interpolation at the ray points, dist*2/(vhead+vtail)
 * accumulated point by point, in the reference's order: same bits).  Rays are HOST data, packed: ray r of period ip
 * owns points ray_offsets[ip*nrays + r] .. ray_offsets[ip*nrays + r + 1]-1 of ray_points (x,y pairs); a ray with
 * fewer than two points gets time 0.  time: HOST out, (nrays, np), ray index fastest (= time(k,j,i) of the
 * reference flattened over (k,j)).  mct_session_group_times runs it on the session's resident group-velocity map,
 * so that map never leaves the device. */
int mct_group_times_dev(const double* d_vel, int np, const mct_grid* g, const double* ray_points,
                        const int64_t* ray_offsets, int nrays, double* time, void* stream);
int mct_session_group_times(mct_session* s, const double* ray_points, const int64_t* ray_offsets, int nrays,
                            double* time);
/* mct_session_group_times integrates through like%gvel of the session's CURRENT (accepted) model: the group map
 * when opt.phaseGroup == 1, the phase map otherwise (likelihood_surf.F90:226-230).  The _pending form does the same for
 * the pending proposal -- its window maps overlaid on the resident ones -- which is where the sampler calls
 * CalGroupTime (mcmc_loc2.f90:228 -> likelihood_surf.F90:350).  ray_points == NULL: the rays made resident by
 * mct_session_set_rays (the reference's straight-ray mode sets its rays up once, likelihood_surf.F90:233-243). */
int mct_session_group_times_pending(mct_session* s, const double* ray_points, const int64_t* ray_offsets, int nrays,
                                    double* time);
int mct_session_set_rays(mct_session* s, const double* ray_points, const int64_t* ray_offsets, int nrays);

/* ---- the Gaussian misfit of surf_likelihood (src/likelihood_surf.F90:356-404) ----------------------------------
 *   time     (nrr, np)     like%phaseTime(k,j,i) flattened over (k,j): nrr = nsrc*nrev, receiver index fastest
 *   ttime    (nrr, 3, np)  dat%ttime: [.,1,.] observed time, [.,2,.] noise level (used when sigdep == 0)
 *   raystat  (nrr, 2, np)  dat%raystat: [.,1,.] == 1 for a ray that carries data
 *   snoise0/1 (np)         RTI%snoise0, RTI%snoise1 (sigdep /= 0: sigma = snoise0*srdist + snoise1); else NULL
 *   srdist   (nrr, np)     like%srdist (sigdep /= 0); else NULL
 *   nrays_total            dat%nrays
 *   out[3]                 like%like, like%misfit, like%unweighted_misfit; sigma (nrr,np): like%sigma, optional
 * Sums are accumulated in the reference's order (period, source, receiver): same bits.  Returns MCT_E_ZERO_NOISE
 * where the reference raises 'The noise level is 0!'. */
int mct_surf_misfit(const double* time, int nrr, int np, int sigdep, int nrays_total, const double* ttime,
                    const int32_t* raystat, const double* snoise0, const double* snoise1, const double* srdist,
                    double out[3], double* sigma);
/* The same chained after CalGroupTime on a session's resident maps: observed data made resident once
 * (mct_session_set_data), then per likelihood only the noise parameters go in and three doubles come out.
 * pending = 0: the current model; 1: the pending proposal.  Rays as for mct_session_group_times.
 * phase_time, sigma: optional host outputs (nrr, np). */
int mct_session_set_data(mct_session* s, int nrr, int sigdep, int nrays_total, const double* ttime,
                         const int32_t* raystat, const double* srdist);
int mct_session_likelihood(mct_session* s, int pending, const double* ray_points, const int64_t* ray_offsets,
                           int nrays, const double* snoise0, const double* snoise1, double out[3],
                           double* phase_time, double* sigma);
/* Curved rays (settings%isStraight == 0) for phase-velocity data: sources, receivers and the fast-marching settings made
 * resident once (mct_session_set_fm2d), then mct_session_likelihood_fm2d assembles like%vel from the resident phase map
 * (+ the pending window), marches every (period, source) (mct_fm2d_times_dev), sets like%srdist = like%phaseTime
 * (likelihood_surf.F90:327-333) and evaluates the misfit: nuclei in (mct_session_propose), three doubles out.
 * Needs mct_session_set_data (raystat decides which sources are marched and, for group-velocity data, in which slot a
 * pair's ray is kept).  A session with phaseGroup == 1 traces the rays as well (uar = 0): a crazy ray gives
 * out = {huge, 0, 0} as likelihood_surf.F90:338-343, else CalGroupTime integrates the resident group map along the rays
 * and like%srdist is their length. */
int mct_session_set_fm2d(mct_session* s, const double* src_x, const double* src_z, int nsrc, const double* rcv_x, const double* rcv_z,
                         int nrc, const mct_fm2d_opts* o);
int mct_session_likelihood_fm2d(mct_session* s, int pending, const double* snoise0, const double* snoise1, double out[3],
                                double* phase_time, double* sigma);
/* stat_rti (src/mcmc_loc2.f90:1966-1978): aveS += vs, stdS += vs**2, aveP += vp, stdP += vp**2 over the session's
 * resident current model; the accumulators stay on the device until mct_session_stat_get. */
int mct_session_stat_accumulate(mct_session* s);
int mct_session_stat_get(mct_session* s, double* aveS, double* stdS, double* aveP, double* stdP, int64_t* nsamples);
int mct_session_stat_reset(mct_session* s);

#ifdef __cplusplus
}
#endif
#endif /* MCTOMO_B200_H */
