/*
 * wabbit_gpu.h -- C ABI of the B200-resident WABBIT block hot path.
 *
 * This is the drop-in boundary: plain pointers, int32 / double scalars, no C++ or torch types.
 * A Fortran host binds these with ISO_C_BINDING (`bind(C)`, scalars `value`); the binding a WABBIT
 * maintainer would add is shown in INTEGRATION.md (fortran/module_gpu_bridge.f90).
 *
 * Each entry point names the tree-level reference routine it replaces (paths relative to the
 * reference checkout).  The per-block plugin calls RHS_meta / GET_DT_BLOCK_meta
 * (LIB/EQUATION/module_physics_metamodule.f90:268,558) are useless granularity for a GPU, so the
 * boundary sits one level up, at the routines that own the block loops.
 *
 * Conventions
 *   - Heavy arrays on the HOST side keep the reference layout: Fortran column-major
 *     hvy(nx,ny,nz,ncomp,number_blocks), nx = Bs+2g, x fastest, float64 (LIB/MESH/allocate_forest.f90:228-267).
 *   - hvy ids and lgt ids are 1-based as in Fortran; -1 means "none".
 *   - Every function returns 0 on success, else a non-zero WABBIT-style integer code; the message is
 *     available from wgpu_last_error().  The library never calls exit()/abort().
 *   - One host thread drives a context (the reference is single-threaded per rank as well).
 *   - Device arrays are owned by the library and indexed by the SAME hvy_id as the host arrays.
 */
#ifndef WABBIT_GPU_H
#define WABBIT_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WGPU_MAX_STAGES 8   /* rows of the Butcher tableau minus one */
#define WGPU_NCOLORS 16

/* order_discretization (LIB/EQUATION/ACMnew/rhs_ACM.f90:1014,1137,1344,1461) */
enum { WGPU_FD_2ND_CENTRAL = 2, WGPU_FD_4TH_CENTRAL = 4, WGPU_FD_6TH_CENTRAL = 6, WGPU_FD_4TH_CENTRAL_OPTIMIZED = 40 };

/* device-resident heavy arrays (LIB/MESH/allocate_forest.f90:228-267) */
enum { WGPU_HVY_BLOCK = 0, WGPU_HVY_WORK = 1, WGPU_HVY_MASK = 2, WGPU_HVY_TMP = 3 };

/* error codes (0 = ok).  WGPU_ERR_DIVERGED mirrors abort(0409201933) in rhs_ACM.f90:145. */
enum {
    WGPU_OK = 0,
    WGPU_ERR_ARG = 1001,
    WGPU_ERR_CUDA = 1002,
    WGPU_ERR_UNSUPPORTED = 1003,
    WGPU_ERR_NO_DEVICE = 1004,
    WGPU_ERR_DIVERGED = 409201933
};

/*
 * Configuration: the subset of type_params (LIB/PARAMS/module_params.f90:18-233) and type_params_acm
 * (LIB/EQUATION/ACMnew/module_ACM.f90:56-154) the hot path reads.  Filled by the host from the same
 * .ini file (LIB/MESH/ini_file_to_params.f90, module_ACM.f90:172-609).
 */
typedef struct wgpu_config {
    int32_t dim;                 /* [Domain] dim: 2 or 3 */
    int32_t Bs[3];               /* [Blocks] number_block_nodes (even) */
    int32_t g;                   /* [Blocks] number_ghost_nodes (host layout ghost width) */
    int32_t g_rhs;               /* [Blocks] number_ghost_nodes_rhs */
    int32_t n_eqn;               /* [Blocks] number_equations: 3 (2-D) or 4 (3-D) for ACM */
    int32_t n_mask;              /* components of hvy_mask held on the device: 0, 5 or 6 */
    int32_t max_blocks;          /* params%number_blocks (per rank) */
    int32_t Jmax;                /* [Blocks] max_treelevel */
    int32_t periodic[3];         /* [Domain] periodic_BC */
    int32_t fd;                  /* WGPU_FD_* */
    int32_t skew_symmetry;       /* [ACM-new] skew_symmetry */
    int32_t penalization;        /* [VPM] penalization */
    int32_t use_sponge;          /* [Sponge] use_sponge */
    int32_t n_stages;            /* s: the tableau below is (s+1) x (s+1) */
    int32_t write_method_fixed_time; /* [Time] write_method == "fixed_time" */
    int32_t device;              /* CUDA device ordinal */
    double domain[3];            /* [Domain] domain_size */
    double c0, nu, gamma_p;      /* [ACM-new] */
    double C_eta;                /* [VPM] C_eta */
    double C_sponge;             /* [Sponge] C_sponge */
    double u_mean_set[3];        /* [ACM-new] u_mean_set */
    double CFL, CFL_eta, CFL_nu; /* [Time] */
    double dt_fixed, dt_max;     /* [Time] */
    double time_max;             /* [Time] */
    double write_time, write_time_first;   /* [Time] */
    double tsave_stats;          /* [Statistics] tsave_stats (9999999.9 = off) */
    double butcher[(WGPU_MAX_STAGES + 1) * (WGPU_MAX_STAGES + 1)]; /* row-major (s+1)x(s+1), ini_file_to_params.f90:640-646 */
} wgpu_config;

typedef struct wgpu_ctx wgpu_ctx;

/* ---- lifecycle: replaces allocate_forest (LIB/MESH/allocate_forest.f90:1) for the device copies ---- */
int32_t wgpu_create(const wgpu_config *cfg, wgpu_ctx **out);
int32_t wgpu_destroy(wgpu_ctx *ctx);
/* wgpu_last_error: copies the last error message of `ctx` (or of a failed wgpu_create if ctx == NULL) into buf -- the text the reference passes
 * to abort(code, msg) (LIB/MODULE/module_globals.f90:131-150); the library returns the code instead of stopping the program */
int32_t wgpu_last_error(const wgpu_ctx *ctx, char *buf, int32_t len);
/* wgpu_set_stream / wgpu_synchronize (no reference counterpart): all work of this context is issued on `cuda_stream` (a cudaStream_t);
 * NULL = default stream */
int32_t wgpu_set_stream(wgpu_ctx *ctx, void *cuda_stream);
int32_t wgpu_synchronize(wgpu_ctx *ctx);

/*
 * ---- topology upload: consumes what updateMetadata_tree leaves in the forest metadata
 *      (LIB/MESH/module_forestMetaData.f90:33-56; neighbour slots LIB/TREE/neighborhood.f90:10-22).
 * hvy_active[n_active]        1-based hvy ids in the order of hvy_active(:,tree_ID)
 * level[n_active]             mesh level of each active block (lgt_block(lgt_id, IDX_MESH_LVL))
 * hvy_neighbor                Fortran array hvy_neighbor(ld, 168) of LGT ids (lgt = rank*max_blocks + hvy), -1 = none
 * rank                        this process' rank (to translate lgt <-> hvy ids)
 */
int32_t wgpu_set_topology(wgpu_ctx *ctx, int32_t n_active, const int32_t *hvy_active, const int32_t *level,
                          const int32_t *hvy_neighbor, int32_t ld, int32_t rank);

/*
 * wgpu_set_treecodes: the numerical binary treecodes of the active blocks (get_tc(lgt_block(lgt_id, IDX_TC_1:IDX_TC_2)),
 * LIB/TREE/module_treelib.f90:227-239, encoding_b :837-871 with max_level = Jmax), same order as hvy_active / level.
 * Blocks listed here but not in the following wgpu_set_topology are known as data sources only (they can be looked up by position).
 * Required BEFORE wgpu_set_topology whenever the grid has coarser / finer neighbour relations: the level-jump ghost
 * patches (restriction, prediction; LIB/MPI/restrict_predict_data.f90:45-202) locate their sources by block position.
 * Grids with level jumps also need wgpu_set_wavelet first (the predictor order is the wavelet's, params%order_predictor).
 */
int32_t wgpu_set_treecodes(wgpu_ctx *ctx, int32_t n_active, const int32_t *hvy_active, const int32_t *level, const int64_t *treecode);

/*
 * wgpu_set_grid / wgpu_set_active: the device-native alternative to wgpu_set_treecodes + wgpu_set_topology ("topology on
 * device"): the neighbour search of find_neighbors / updateNeighbors_tree (LIB/MESH/find_neighbors.f90:18-180, updateNeighbors_tree.f90)
 * runs on the GPU.  The host names the RESIDENT blocks -- hvy id, mesh level, numerical treecode (get_tc(lgt_block(:, IDX_TC_1:IDX_TC_2))) --
 * and the active list; no hvy_neighbor table is needed.  A position hash is built on the device and one thread per (active block,
 * direction) establishes the relation exactly as find_neighbor does (same level / finer / coarser, incl. the last-digit rule for coarser edge
 * and corner neighbours); the gather tables of the stencil and wavelet kernels, the level-jump patch lists, the coarse-extension list, the
 * senders of restricted data and the interior / partition-boundary split follow by stable compaction in the order of the active list.
 *   Resident blocks that are not active are data sources only: halo copies of other ranks' blocks (the slots wgpu_set_halo declared),
 *   and, for the passes of adapt_tree's full wavelet transformation, all blocks of the full tree (leaves + mothers, init_full_tree):
 *   register the tree once with wgpu_set_grid, then name each pass's blocks with wgpu_set_active (wavelet_decompose_full_tree's level
 *   loop, adapt_tree.f90:268-545).  A leaf next to a refined region then finds the region's mother on its own level, as
 *   sync_TMP_from_MF provides it.  An entry with hvy id <= 0 names a block that exists but whose data are not resident on this rank: it
 *   takes part in the relations (a leaf next to it is "next to a coarser block": the coarse extension applies) but is never read.
 * wgpu_topology_tables / wgpu_topology_list: read the derived tables back (tests, debugging): nbr27 / wnbr27 [max_blocks][27] gather codes
 *   (>= 0 block index, -1 none, <= -2 pool patch), counts[8] = n_active, n_jump, n_wjump, n_ce, n_rst, n_interior, n_boundary, has_jumps;
 *   list `which`: 0 stage-kernel face patches (block, dir), 1 wavelet ghost patches (block, dir), 2 coarse extension (block, dir),
 *   3 restriction senders (block, direction mask), 4 interior blocks, 5 partition-boundary blocks; 0-based block indices.
 */
int32_t wgpu_set_grid(wgpu_ctx *ctx, int32_t n_resident, const int32_t *hvy_ids, const int32_t *level, const int64_t *treecode, int32_t n_active,
                      const int32_t *hvy_active);
int32_t wgpu_set_active(wgpu_ctx *ctx, int32_t n_active, const int32_t *hvy_active);
int32_t wgpu_topology_tables(wgpu_ctx *ctx, int32_t *nbr27, int32_t *wnbr27, int32_t *counts);
int32_t wgpu_topology_list(wgpu_ctx *ctx, int32_t which, int32_t n, int32_t *out_a, int32_t *out_b);

/*
 * ---- data movement between the host's Fortran arrays and the resident device arrays.
 * host points at element (1,1,1,1,1) of hvy(nx,ny,nz,ncomp_host,number_blocks); blocks listed in hvy_ids
 * (1-based, n of them) are moved.  `slot` selects hvy_work(:,:,:,:,:,slot) (ignored for other arrays).
 * Download fills ghost layers of width g_sync (0..g) by running the ghost synchronisation on the fly (what
 * sync_ghosts_tree with ignore_Filter would have left there: same-level copy, and on grids with level jumps decimation from finer
 * and prediction from coarser neighbours for all 26 relations); the rest of the ghost region is untouched.
 */
int32_t wgpu_upload(wgpu_ctx *ctx, int32_t array_id, int32_t slot, const int32_t *hvy_ids, int32_t n,
                    const double *host, int32_t ncomp_host);
int32_t wgpu_download(wgpu_ctx *ctx, int32_t array_id, int32_t slot, const int32_t *hvy_ids, int32_t n,
                      double *host, int32_t ncomp_host, int32_t g_sync);
/* wgpu_set_transfer_mode: how a PAGE-LOCKED 3-D host array crosses PCIe, per direction (no reference counterpart: the reference's
 *   hvy arrays never leave host memory; this is the cost of the drop-in boundary of RungeKuttaGeneric, runge_kutta_generic.f90:136-154).
 *   1 (default): copy engines -- for every interior xy plane the contiguous span first..last interior node is moved by DMA
 *      (cudaMemcpy3DAsync, rows of (Bs-1)*nx+Bs doubles) and a layout kernel converts between that and the resident arrays; uploads and
 *      downloads of different contexts run concurrently at the full rate of each direction.  A download with g_sync = 0 then also writes
 *      the x ghost nodes that lie between the interior rows: with the same-level x neighbour's values (what sync_ghosts would put
 *      there), or 0 where no such neighbour is resident on this GPU.  Ghost nodes are invalid after a time step in the reference too.
 *   0: the layout kernels read / write the host array directly (zero-copy); exactly the interiors (+ the g_sync shell) are touched.
 *   Pageable host arrays, 2-D, g_sync > 0 and ncomp_host != the array's component count always take the other paths. */
int32_t wgpu_set_transfer_mode(wgpu_ctx *ctx, int32_t upload_mode, int32_t download_mode);

/*
 * ---- compute ----
 * wgpu_sync_ghosts: replaces sync_ghosts_RHS_tree / sync_ghosts_tree
 *   (LIB/MPI/synchronize_ghosts_generic.f90:155-174, 181-343).  Same-level relations are not materialised in HBM
 *   (the stencil kernels gather them from the neighbour's interior); this call refreshes the level-jump / remote
 *   patch pool and is a no-op on a uniform single-GPU grid.
 */
int32_t wgpu_sync_ghosts(wgpu_ctx *ctx, int32_t array_id, int32_t slot, int32_t g_minus, int32_t g_plus);

/*
 * ---- mask function and statistics on the device (the physics module's CREATE_MASK / STATISTICS entry points for the ACM module)
 * wgpu_create_mask: createMask_tree -> create_mask_2D_ACM / create_mask_3D_ACM (LIB/MESH/createMask_tree.f90, LIB/EQUATION/ACMnew/
 *   create_mask.f90:6-320) for the closed-form geometries: WGPU_GEOM_CYLINDER (2-D "cylinder" / "circle", draw_circle) with the p-norm sponge
 *   of sponge_2D (LIB/EQUATION/ACMnew/sponge.f90) and WGPU_GEOM_SPHERE (3-D, draw_sphere; centre(t) = center0 + velocity * time, u_s = velocity),
 *   cosine smoothing of width `smoothing_width` (= C_smooth * dx_min, module_ACM.f90:459-471).  Writes the six components [chi, u_s(3),
 *   colour = 1, sponge] of every active block's interior into the resident hvy_mask; nothing crosses PCIe.  chi is drawn only with
 *   penalization = 1, the sponge only with use_sponge = 1 (as the reference).  Other geometries (insects, STL): fill hvy_mask with wgpu_upload.
 * wgpu_statistics: the integral_stage + post_stage reductions of STATISTICS_ACM (LIB/EQUATION/ACMnew/statistics_ACM.f90:138-430) over the
 *   active blocks (and over the ranks of the communicator): out[0..2] mean flow * volume, [3] e_kin, [4] ACM energy, [5] mask volume (colour 1),
 *   [6] sponge volume, [7..9] penalization power (solid input, solid dissipation, sponge), [10..12] force on colour 1, [13] max |u|^2,
 *   [14] / [15] max / min of div(u) outside the solid (with_divergence = 1: one RHS evaluation into hvy_work slot 2, whose pressure row is
 *   -c0^2 div(u) - gamma_p p; else 0), [16..18] residual velocity in the solid (sum over blocks of max * dV, as the reference).  out: 19 doubles.
 *   flags: WGPU_STAT_DIVERGENCE (1) as above; WGPU_STAT_VORTICITY (2) adds, on one rank, [19] enstrophy, [20] max |vorticity|, [21] helicity
 *   (3-D; 0 in 2-D), [22] dissipation = -nu * integral of u . laplace(u) (0 for nu = 0) -- compute_vorticity and compute_dissipation
 *   (LIB/OPERATORS/compute_vorticity.f90:3-67, compute_dissipation.f90:5-78; statistics_ACM.f90:371-387) with the discretization's first- and
 *   second-derivative stencils on ghost-synchronised copies of the blocks (level jumps: the wavelet's predictor, wgpu_set_wavelet).
 *   out: 23 doubles with WGPU_STAT_VORTICITY.
 */
enum { WGPU_STAT_DIVERGENCE = 1, WGPU_STAT_VORTICITY = 2 };
enum { WGPU_GEOM_CYLINDER = 1, WGPU_GEOM_SPHERE = 2 };
int32_t wgpu_create_mask(wgpu_ctx *ctx, double time, int32_t geometry, const double *center0, const double *velocity, double radius,
                         double smoothing_width, double L_sponge, double p_sponge);
int32_t wgpu_statistics(wgpu_ctx *ctx, double time, int32_t flags, double *out);

/* wgpu_set_ghost_filter: the ignore_Filter switch of sync_ghosts_tree (LIB/MPI/synchronize_ghosts_generic.f90:125-153).  With a lifted
 *   wavelet (CDFXY, Y > 0) the reference's default synchronisation restricts through the HD filter: a ghost node owned by a finer
 *   neighbour receives the filtered value, except next to that neighbour's own coarser / finer neighbours, where the plain value is
 *   copied (restrict_copy_at_CE, LIB/MPI/restrict_predict_data.f90:121-172; blockFilterXYZ_vct, LIB/WAVELETS/module_wavelets.f90:307-401).
 *   This is the default here too (ignore_filter = 0) for every wavelet-side synchronisation: wgpu_download with g_sync > 0, wgpu_fwt /
 *   wgpu_iwt and wgpu_refine.  ignore_filter = 1 makes them restrict by plain decimation.  The right-hand side always ignores the
 *   filter (sync_ghosts_RHS_tree). */
int32_t wgpu_set_ghost_filter(wgpu_ctx *ctx, int32_t ignore_filter);

/* wgpu_set_mask_sphere: replaces, for a (translating) sphere, the per-stage createMask_tree -> CREATE_MASK_meta -> create_mask_3D_ACM ->
 *   draw_sphere chain (LIB/TIME/RHS_wrapper.f90:51, LIB/MESH/createMask_tree.f90, LIB/EQUATION/ACMnew/create_mask.f90:6-174,
 *   LIB/EQUATION/insects/module_geometry.f90 draw_sphere) and the hvy_mask reads of the penalization term (rhs_ACM.f90:1192-1195): the
 *   stage kernel evaluates chi = step_cosine(|x - (center0 + velocity*t)| - radius, smoothing_width) / C_eta and u_s = velocity at the stage
 *   time t = time + dt*rk_coeffs(j,1) itself, so no mask array is generated, uploaded or read.  enable = 0 returns to hvy_mask.
 *   3-D, penalization = 1, FD_4th_central. */
int32_t wgpu_set_mask_sphere(wgpu_ctx *ctx, int32_t enable, const double *center0, const double *velocity, double radius, double smoothing_width);

/* wgpu_rhs: replaces RHS_wrapper (LIB/TIME/RHS_wrapper.f90:16-260) for physics "ACM-new":
 *   hvy_work(:,:,:,:,:,dst_slot) = RHS(src), src = hvy_block (src_slot = 0) or hvy_work(..., src_slot).
 *   Includes the integral_stage divergence guard (rhs_ACM.f90:133-146) -> WGPU_ERR_DIVERGED. */
int32_t wgpu_rhs(wgpu_ctx *ctx, double time, int32_t src_slot, int32_t dst_slot);

/* wgpu_calculate_time_step: replaces calculate_time_step (LIB/TIME/calculate_time_step.f90:2-125) incl. the
 *   per-block GET_DT_BLOCK_ACM (module_ACM.f90:617-691) and the global MIN. */
int32_t wgpu_calculate_time_step(wgpu_ctx *ctx, double time, double *dt);

/* wgpu_rk_step: replaces RungeKuttaGeneric (LIB/TIME/runge_kutta_generic.f90:1-156): ghost sync, dt, all stages,
 *   final combination; hvy_block is advanced in place on the device.  *dt receives the step taken. */
int32_t wgpu_rk_step(wgpu_ctx *ctx, double time, int32_t iteration, double *dt);

/*
 * ---- wavelet side (grid adaptation): decomposition, thresholding, reconstruction ----
 * wgpu_set_wavelet: replaces setup_wavelet (LIB/WAVELETS/module_wavelets.f90:1031-1417) for "CDFXY", X in {2,4,6},
 *   Y in {0,2,4,6}, Y <= X; returns the default ghost widths g = X-1+max(Y-1,0) and g_rhs = X/2 (ini_file_to_params.f90:467-468).
 * wgpu_fwt / wgpu_iwt: replace the block loops over waveletDecomposition_optimized_block /
 *   waveletReconstruction_optimized_block (LIB/WAVELETS/wavelet_decomposition_reconstruction.f90:23,426) together with
 *   the sync_ghosts_tree that precedes them (LIB/MESH/adapt_tree.f90:403-446, 813-843): dst = transform(src), result in
 *   spaghetti order (scaling coefficients at interior offsets 0,2,4,..).  src and dst must be different arrays.
 * wgpu_norm: componentWiseNorm_tree (LIB/OPERATORS/componentWiseNorm_tree.f90:1) over the leaf interiors, every component on its own:
 *   norm_id 0 Linfty (bit-exact), 1 L1, 2 L2, 3 H1 (= L2).  The sums are formed per block in a fixed order on the device and
 *   added over the blocks in hvy_active order on the host: deterministic, equal to the reference's sequential sum to round-off.
 * wgpu_threshold: wavelet_renorm_block + threshold_block on a decomposed array (LIB/INDICATORS/threshold_block.f90:1-130,
 *   module_wavelets.f90:1848-1960): refinement_status[n_active] = -1 iff all(detail <= eps*norm) else 0, in the order of
 *   hvy_active; eps_norm_id 0 Linfty / 1 L1 / 2 L2 / 3 H1; thresh_comp = params%threshold_state_vector_component;
 *   norm may be NULL (eps not normalised); detail_out[n_active*n_eqn] may be NULL.
 */
/* wgpu_coarse_extension: coarse_extension_modify(CE_case="tree") on a leaf grid (LIB/MPI/reconstruction_step.f90:3-100 ->
 *   coarseExtensionManipulateWC_block / ...SC_block, LIB/WAVELETS/module_wavelets.f90:877-1027): on every block, in every direction
 *   whose neighbour is coarser, the wavelet coefficients of the decomposed array (wd) are zeroed in a strip Nwcl/Nwcr deep
 *   (clear_wc) and its scaling coefficients are copied from the array of original values (orig) in a strip Nscl/Nscr deep
 *   (copy_sc); the sizes are setup_wavelet's (module_wavelets.f90:1368-1417, incl. the widening to 2*FD_max_size).  Interiors
 *   only: ghost nodes are not stored on the device. */
int32_t wgpu_coarse_extension(wgpu_ctx *ctx, int32_t wd_id, int32_t wd_slot, int32_t orig_id, int32_t orig_slot, int32_t clear_wc, int32_t copy_sc);
/* wgpu_patch_details: the block-level half of addSecurityZone_CE_tree (LIB/MESH/securityZone_tree.f90:140-298): for every pair
 *   (hvy_ids[k], dirs[k] = (dz+1)*9+(dy+1)*3+(dx+1)) the Linfty detail per component of the decomposed block inside the Nwcl / Nwcr deep
 *   strip that faces the neighbour in that direction (threshold_block with `indices`), detail_out[k*n_eqn + c].  The host compares with
 *   eps*norm and keeps the insignificant neighbour alive if the strip is significant. */
int32_t wgpu_patch_details(wgpu_ctx *ctx, int32_t array_id, int32_t slot, int32_t n, const int32_t *hvy_ids, const int32_t *dirs, double *detail_out);
/* wgpu_patch_details_norm: the same with the coefficients renormalised for eps_norm = L1 / L2 / H1 first (wavelet_renorm_block,
 *   LIB/WAVELETS/module_wavelets.f90:1900-1945; eps_norm_id as in wgpu_threshold: 0 Linfty, 1 L1, 2 L2, 3 H1; level_ref = Jmax) */
int32_t wgpu_patch_details_norm(wgpu_ctx *ctx, int32_t array_id, int32_t slot, int32_t eps_norm_id, int32_t level_ref, int32_t n,
                                const int32_t *hvy_ids, const int32_t *dirs, double *detail_out);
int32_t wgpu_set_wavelet(wgpu_ctx *ctx, const char *name, int32_t *g_default, int32_t *g_rhs_default);
int32_t wgpu_fwt(wgpu_ctx *ctx, int32_t src_id, int32_t src_slot, int32_t dst_id, int32_t dst_slot);
int32_t wgpu_iwt(wgpu_ctx *ctx, int32_t src_id, int32_t src_slot, int32_t dst_id, int32_t dst_slot);
/* wgpu_iwt_ce: the reconstruction step of wavelet_reconstruct_full_tree_CEoptimized (LIB/MESH/adapt_tree.f90:686-987) for the active blocks:
 *   sync_SCWC_from_MC + coarse_extension_modify + waveletReconstruction_optimized_block.  Ghost nodes that face a same-level block take
 *   that block's coefficients from array wd; ghost nodes that face a coarser leaf take the leaf's value in array `coarse` (hvy_tmp of
 *   the reference) at the scaling positions and zero elsewhere.  dst may be the `coarse` array (the reference writes hvy_tmp too). */
int32_t wgpu_iwt_ce(wgpu_ctx *ctx, int32_t wd_id, int32_t wd_slot, int32_t coarse_id, int32_t coarse_slot, int32_t dst_id, int32_t dst_slot);
int32_t wgpu_norm(wgpu_ctx *ctx, int32_t array_id, int32_t slot, int32_t norm_id, double *out);
int32_t wgpu_threshold(wgpu_ctx *ctx, int32_t array_id, int32_t slot, int32_t eps_norm_id, int32_t level_ref, const int32_t *thresh_comp,
                       const double *eps, const double *norm, int32_t *refinement_status, double *detail_out);

/*
 * ---- refinement / coarsening of heavy data (the light-data side -- which blocks, which free ids -- stays with the host) ----
 * wgpu_refine: replaces the block loop of refinement_execute_tree -> refineBlock (LIB/MESH/refinementExecute.f90:1-120) on hvy_block:
 *   every mother (ghosts synchronised on the fly, as the sync_ghosts_tree before refine_tree, LIB/MAIN/main.f90:314-322, with
 *   ignore_Filter) is interpolated with the wavelet's predictor to 2^dim daughters.  daughter_hvy[i*2^dim + digit], digit bit0 -> y,
 *   bit1 -> x, bit2 -> z (treecode digit).  Blocks that are not refined either stay in place (n_keep < 0) or move from hvy id
 *   keep_src[i] to keep_dst[i] (the block_xfer of balanceLoad_tree("refine_post"), LIB/MESH/balanceLoad_tree.f90, on one rank);
 *   every active block must then be a mother or listed, and no destination slot may be used twice.  hvy_tmp is used as the second buffer and holds the old grid's data afterwards.  The topology on the device
 *   is the OLD one during the call; upload the new one (wgpu_set_treecodes + wgpu_set_topology) before any other compute call.
 * wgpu_coarsen: replaces sync_D2M + the mother assembly of executeCoarsening (LIB/MESH/executeCoarsening_tree.f90:125-230): octant
 *   `digit` of hvy_block(mother) = the scaling coefficients (even spaghetti positions, conversion_routines.f90:149) of the decomposed
 *   daughter in array (src_id, src_slot) -- the result of wgpu_fwt; the source must not be hvy_block.  A mother may reuse the id of one
 *   of its daughters.
 */
int32_t wgpu_refine(wgpu_ctx *ctx, int32_t n, const int32_t *mother_hvy, const int32_t *daughter_hvy, int32_t n_keep, const int32_t *keep_src,
                    const int32_t *keep_dst);
int32_t wgpu_coarsen(wgpu_ctx *ctx, int32_t n, const int32_t *mother_hvy, const int32_t *daughter_hvy, int32_t src_id, int32_t src_slot);
/* wgpu_move_blocks: the same-rank part of block_xfer (LIB/MPI/block_xfer_nonblocking.f90:16) as used by balanceLoad_tree: hvy_block(dst[i]) =
 *   hvy_block(src[i]) for all i at once (a permutation is fine).  Every block that must survive has to be listed (identity pairs
 *   allowed): the move goes through hvy_tmp, which becomes hvy_block.  Upload the new topology afterwards. */
int32_t wgpu_move_blocks(wgpu_ctx *ctx, int32_t n, const int32_t *src_hvy, const int32_t *dst_hvy);

/*
 * ---- multi-GPU (one process per GPU): the Runge-Kutta step split at the points where ranks must talk.
 * The reference exchanges ghost patches with MPI_Isend/Irecv once per sync (LIB/MPI/xfer_block_data.f90:10-99) and
 * all-reduces dt with MPI_MIN (LIB/TIME/calculate_time_step.f90:48).  Here the host moves bytes with NCCL
 * (torch.distributed) between the library's pack kernel and the stage kernels, which read remote halos straight
 * from the receive buffer ("patch pool") -- there is no unpack pass.
 *
 *   wgpu_rk_begin(time)        local CFL candidate -> device scalar (wgpu_dtmin_pointer); host all-reduces it (MIN)
 *   wgpu_rk_dt(time)           calculate_time_step's clipping, dt stays on the device
 *   for j = 1..s:  wgpu_pack_halo(j); <host: all-to-all send_buf -> pool>; wgpu_rk_stage(j, WGPU_BLOCKS_ALL)
 *                  (or stage(j, INTERIOR) overlapped with the exchange, then stage(j, BOUNDARY))
 *   wgpu_rk_end(&dt)
 * wgpu_rk_step() is exactly this sequence without the host steps.
 *
 * wgpu_set_exchange: face patches this rank receives (block hvy id, direction index (dz+1)*9+(dy+1)*3+(dx+1)) in
 * the order they arrive in `pool`, and the patches it sends in the order they are laid out in `send_buf`.
 * Patch = nc x (g_rhs deep strip), laid out as the receiver's ghost strip, x fastest.  pool / send_buf are device
 * buffers owned by the caller (n_recv resp. n_send patches of wgpu_patch_doubles() doubles).
 * Must be called after wgpu_set_topology (which accepts neighbours on other ranks only if this call follows).
 */
/*
 * Halo blocks: the multi-GPU mode for grids with level jumps and for the wavelet side.  Every block of another rank that appears in
 * the hvy_neighbor rows of this rank's blocks (any of the 168 relations: same level, coarser, finer) is mirrored in a local slot
 * behind the rank's own blocks.  The copies are part of the block lookup and of the neighbour tables, so every kernel (stage, level-jump
 * patches, wavelet transforms, refinement, download with ghosts) runs unchanged; they are never advanced.  The reference moves only
 * ghost patches (xfer_block_data.f90); whole blocks cost about Bs/(2 g) times the bytes but need no second index space, and the
 * transfer is hidden behind the blocks that have no halo neighbour.
 *   wgpu_set_halo            halo_lgt[k] (lgt id = owner_rank*max_blocks + owner_hvy) is mirrored in slot halo_hvy[k] (consecutive slots,
 *                            in the order the owners' data arrive); send_hvy: own blocks other ranks mirror, in the order they are laid
 *                            out in send_buf (device memory of the caller, n_send blocks of n_eqn*Bs^dim doubles).  Call it before
 *                            wgpu_set_treecodes (which then also lists the halo slots) and wgpu_set_topology.
 *   wgpu_pack_halo(stage)    in halo mode: copies the send blocks of the stage input into send_buf
 *   wgpu_rk_stage_halo_pointer  where the halo slots of the stage input start: the host receives straight into them (no unpack pass)
 *   wgpu_pack_blocks / wgpu_halo_pointer   the same for a named array (before wgpu_fwt, wgpu_refine, wgpu_download with ghosts)
 */
int32_t wgpu_set_halo(wgpu_ctx *ctx, int32_t n_halo, const int32_t *halo_lgt, const int32_t *halo_hvy, const int32_t *halo_level, int32_t n_send,
                      const int32_t *send_hvy, double *send_buf);
int32_t wgpu_pack_blocks(wgpu_ctx *ctx, int32_t array_id, int32_t slot);
int32_t wgpu_halo_pointer(wgpu_ctx *ctx, int32_t array_id, int32_t slot, void **ptr, int64_t *n_doubles);
int32_t wgpu_rk_stage_halo_pointer(wgpu_ctx *ctx, int32_t stage, void **ptr, int64_t *n_doubles);

/* Filtered restriction across ranks (lifted wavelets, wgpu_set_ghost_filter(0)): a block whose FINER neighbour lives on another rank
 *   needs that neighbour's HD-filtered, decimated copy (restrict_copy_at_CE), which only the owner can form (it reads the fine block's own
 *   same-level neighbours).  wgpu_set_halo_restrict (after wgpu_set_topology): recv_halo_hvy = halo slots of the finer neighbours, in
 *   arrival order; send_hvy = own blocks whose copies other ranks need, in the order they are laid out in send_buf (n_eqn*(Bs/2)^dim
 *   doubles each).  Per synchronisation of an array: wgpu_pack_blocks + all-to-all (the blocks), then wgpu_restrict_pack(array) +
 *   all-to-all of send_buf into wgpu_restrict_halo_pointer, then the consumer (wgpu_fwt, wgpu_refine, wgpu_download with ghosts). */
int32_t wgpu_set_halo_restrict(wgpu_ctx *ctx, int32_t n_recv, const int32_t *recv_halo_hvy, int32_t n_send, const int32_t *send_hvy, double *send_buf);
int32_t wgpu_restrict_pack(wgpu_ctx *ctx, int32_t array_id, int32_t slot);
int32_t wgpu_restrict_halo_pointer(wgpu_ctx *ctx, void **ptr, int64_t *n_doubles);

/* wgpu_gather_blocks / wgpu_scatter_blocks: the two local halves of block_xfer (LIB/MPI/block_xfer_nonblocking.f90:16) between ranks, as
 *   balanceLoad_tree and the gathering of sister blocks before a coarsening need it: whole blocks (interiors) of a resident array are
 *   packed into / unpacked from a contiguous device buffer of the caller, block k of the list at k*n_eqn*Bs^dim doubles; the host moves
 *   the buffer between ranks (NCCL).  A block received into a free slot is ordinary local data afterwards (wgpu_move_blocks,
 *   wgpu_coarsen and wgpu_refine address it by its slot). */
int32_t wgpu_gather_blocks(wgpu_ctx *ctx, int32_t array_id, int32_t slot, int32_t n, const int32_t *hvy_ids, double *device_buf);
int32_t wgpu_scatter_blocks(wgpu_ctx *ctx, int32_t array_id, int32_t slot, int32_t n, const int32_t *hvy_ids, const double *device_buf);

enum { WGPU_BLOCKS_ALL = 0, WGPU_BLOCKS_INTERIOR = 1, WGPU_BLOCKS_BOUNDARY = 2 };
int64_t wgpu_patch_doubles(const wgpu_ctx *ctx);
int32_t wgpu_set_exchange(wgpu_ctx *ctx, int32_t n_recv, const int32_t *recv_hvy, const int32_t *recv_dir, double *pool,
                          int32_t n_send, const int32_t *send_hvy, const int32_t *send_dir, double *send_buf);
int32_t wgpu_pack_halo(wgpu_ctx *ctx, int32_t stage);
int32_t wgpu_rk_begin(wgpu_ctx *ctx, double time);
int32_t wgpu_dtmin_pointer(wgpu_ctx *ctx, void **ptr);
int32_t wgpu_rk_dt(wgpu_ctx *ctx, double time);
int32_t wgpu_rk_stage(wgpu_ctx *ctx, int32_t stage, int32_t which_blocks);
int32_t wgpu_rk_end(wgpu_ctx *ctx, double *dt);
/* number of blocks in each set (ALL / INTERIOR / BOUNDARY) */
int32_t wgpu_block_count(const wgpu_ctx *ctx, int32_t which_blocks);

/*
 * ---- multi-GPU inside the library: NCCL over NVLink on a communicator the library owns (multigpu.cu).  One process per GPU.
 * The reference's counterparts: MPI_Isend / Irecv of ghost patches (LIB/MPI/xfer_block_data.f90:10-99), MPI_Allreduce(MIN) of dt
 * (LIB/TIME/calculate_time_step.f90:48), block_xfer (LIB/MPI/block_xfer_nonblocking.f90:16), synchronize_lgt_data (LIB/MESH/).
 *   wgpu_comm_unique_id      rank 0: a 128-byte id (ncclGetUniqueId); the host broadcasts it (MPI_Bcast) ...
 *   wgpu_comm_init           ... and every rank joins (ncclCommInitRank).  NCCL is loaded at run time (dlopen).
 *   wgpu_comm_set_counts     per-peer counts [world] of the exchange declared last: face patches (wgpu_set_exchange) or halo blocks
 *                            (wgpu_set_halo), and of the filtered copies of wgpu_set_halo_restrict (may be NULL)
 *   wgpu_rk_steps            n_steps of RungeKuttaGeneric back to back (the N_dt_per_grid loop of LIB/POSTPROCESSING/performance_test.f90:
 *                            194-199): time, dt and the divergence flag stay on the device, dt's MIN over ranks is an ncclAllReduce on the
 *                            device scalar, every stage = pack -> grouped ncclSend / ncclRecv on a second stream, straight into the patch pool
 *                            resp. the halo slots of the stage input (face patches by default: the pack kernel stores them into the peers'
 *                            pools over NVLink, see wgpu_comm_set_transport) || stage kernel on interior blocks -> stage kernel on
 *                            partition-boundary blocks.  One host synchronisation at the end: *time_out = time after the last step, *dt_last its dt.
 *                            Works without a communicator (single GPU) as well.
 *   wgpu_exchange_array      refresh the halo copies (and, filtered != 0 with a lifted wavelet, the filtered copies of finer neighbours)
 *                            of a named array: before wgpu_fwt, wgpu_refine, wgpu_download with ghosts
 *   wgpu_ship_blocks         block_xfer between ranks.  Item k: a block in slot src_slot[k] (1-based) of the array on rank src_rank[k] is
 *                            needed on dst_rank[k]; all ranks pass the same lists.  Remote blocks are received straight into the free slots
 *                            first_free, first_free + 1, ... (peer-major, item order); local_slot[] receives the local slot of every item with
 *                            dst_rank == me (item order), *next_free the first slot still free.
 *   wgpu_comm_allreduce      small host arrays (norms, flags): op 0 MAX, 1 MIN, 2 SUM
 *   wgpu_comm_allgatherv_i32 concatenation over ranks of int32 lists whose lengths every rank knows (refinement flags)
 */
int32_t wgpu_comm_unique_id(char *id128);
int32_t wgpu_comm_init(wgpu_ctx *ctx, const char *id128, int32_t rank, int32_t world);
int32_t wgpu_comm_destroy(wgpu_ctx *ctx);
int32_t wgpu_comm_info(const wgpu_ctx *ctx, int32_t *rank, int32_t *world);
int32_t wgpu_comm_set_counts(wgpu_ctx *ctx, const int32_t *send_counts, const int32_t *recv_counts, const int32_t *restrict_send_counts,
                             const int32_t *restrict_recv_counts);
/* wgpu_comm_set_transport / wgpu_comm_transport: how the face patches of wgpu_set_exchange travel inside wgpu_rk_steps (the reference's
 *   MPI_Isend / Irecv of xfer_block_data.f90:10-99).  1 (default): PEER STORES -- wgpu_comm_set_counts exports the library's receive pools
 *   by CUDA IPC, the pack kernel writes every patch straight into the receiver's pool over NVLink and releases a flag per peer, the
 *   receiver acquires the flags in front of its partition-boundary blocks (no send buffer, no NCCL kernel).  0: grouped ncclSend / ncclRecv.
 *   If peer access or IPC is unavailable on any rank, every rank uses NCCL (agreed by an all-reduce); wgpu_comm_transport reports which
 *   transport is active (1 peer stores, 0 NCCL).  Set before wgpu_comm_set_counts. */
int32_t wgpu_comm_set_transport(wgpu_ctx *ctx, int32_t peer_stores);
int32_t wgpu_comm_transport(const wgpu_ctx *ctx);
int32_t wgpu_rk_steps(wgpu_ctx *ctx, double time, int32_t n_steps, double *time_out, double *dt_last);
int32_t wgpu_exchange_array(wgpu_ctx *ctx, int32_t array_id, int32_t slot, int32_t filtered);
int32_t wgpu_ship_blocks(wgpu_ctx *ctx, int32_t array_id, int32_t slot, int32_t n_items, const int32_t *src_rank, const int32_t *src_slot,
                         const int32_t *dst_rank, int32_t first_free, int32_t *local_slot, int32_t *next_free);
int32_t wgpu_comm_allreduce(wgpu_ctx *ctx, double *inout, int32_t n, int32_t op);
int32_t wgpu_comm_allgatherv_i32(wgpu_ctx *ctx, const int32_t *mine, const int32_t *counts, int32_t *out);

/*
 * wgpu_rkc_step: RungeKuttaChebychev (LIB/TIME/runge_kutta_chebychev.f90:6-146; time_step_method = RungeKuttaChebychev in timeStep_tree.f90:38):
 *   one step of the s-stage scheme (s >= 4, else code 1715929 as the reference) with the coefficient rows mu(s,1:s), mu_tilde, nu, gamma_tilde,
 *   c of the host's tables (setup_RKC_coefficients, or the RKC_custom_scheme of the parameter file): calculate_time_step, F0 = rhs(y00),
 *   y1 = y0 + mu~_1 dt F0, then for i = 2..s  F1 = rhs(y1) at tau = time + c(i-1) dt  and
 *   y2 = (1 - mu_i - nu_i) y00 + mu_i y1 + nu_i y0 + mu~_i dt F1 + gamma~_i dt F0  (evaluated left to right, uncontracted).  Six registers as the
 *   reference: hvy_block, two stage inputs and hvy_work slots 2..4.  Ghost synchronisation is fused into every right-hand side.  One rank.
 */
int32_t wgpu_rkc_step(wgpu_ctx *ctx, double time, int32_t iteration, int32_t s, const double *mu, const double *mu_tilde, const double *nu,
                      const double *gamma_tilde, const double *c, double *dt);

/* wgpu_krylov_step: timeStep_tree -> krylov_time_stepper (time_step_method = "Krylov", LIB/TIME/krylov.f90:1-190): the exponential integrator
 *   u(t + dt) = u + dt phi_1(dt J) F(u) on the Krylov space of the Jacobian (finite differences of the right-hand side with
 *   eps = |u| sqrt(epsilon)).  calculate_time_step, M right-hand sides (ghost synchronisation fused), Arnoldi with modified Gram-Schmidt on
 *   the block interiors (wabbit_norm / scalarproduct, :500-595; every scalar product is read by the host, as the reference's MPI_Allreduce),
 *   the matrix exponential of the augmented Hessenberg matrix on the host (wgpu_expm_pade), error estimate |beta h(M+1,M) phi(M,M+2)|.
 *   M_max = params%M_krylov (M_max + 3 registers of the size of hvy_block are allocated on first use); dynamic = 1
 *   (krylov_subspace_dimension = "dynamic"): stop at the first M with err <= err_threshold, and at M_max shrink dt by 0.9 until it is.
 *   Outputs: dt (possibly shrunk), M_used, err (what the reference appends to krylov_err.t).  One rank.
 * wgpu_expm_pade: expM_pade -> DGPADM (krylov.f90:193-396; Expokit): exp(H) of an m x m matrix (row- or column-major alike), degree-6 Pade
 *   fraction with scaling and squaring.  Host code, no device needed. */
int32_t wgpu_krylov_step(wgpu_ctx *ctx, double time, int32_t iteration, int32_t M_max, int32_t dynamic, double err_threshold, double *dt,
                         int32_t *M_used, double *err);
int32_t wgpu_expm_pade(const double *H, int32_t m, double *E);

/*
 * wgpu_filter: filter_wrapper (LIB/TIME/filter_wrapper.f90:1-78) on the resident hvy_block: filter_type = "explicit_3pt" ... "explicit_21pt" or
 *   "superviscosity_2nd" ... "_20th" (the binomial stencils of generate_superviscosity_stencil, + identity), applied with blockFilterXYZ_vct
 *   (LIB/WAVELETS/module_wavelets.f90:307-401: x, then y, then z; every sum starts from 0 and adds the shifts in increasing order) to the components
 *   with filter_component[c] != 0 (NULL: all) of every active block, or only of the blocks on Jmax / all but those.  Ghost values are gathered
 *   on the fly as sync_ghosts_tree leaves them.  Error codes 251106 / 251107 / 251108 as the reference.  3-D.
 */
int32_t wgpu_filter(wgpu_ctx *ctx, const char *filter_type, const int32_t *filter_component, int32_t only_maxlevel, int32_t all_except_maxlevel);

/* Stage-kernel timing with CUDA events on the context's stream (for the roofline line of bench.py; the reference times the same region on the
 * host with toc("timestep (RHS wrapper)"), LIB/TIMING/module_timing.f90:76, runge_kutta_generic.f90:71-73, 125-127):
 * wgpu_profile(ctx, 1) starts recording an event pair around every stage-kernel launch (at most 4096 pairs),
 * wgpu_profile_read synchronises, returns their number and summed duration in milliseconds, and resets. */
int32_t wgpu_profile(wgpu_ctx *ctx, int32_t enable);
int32_t wgpu_profile_read(wgpu_ctx *ctx, int32_t *n_launches, double *total_ms);

/* wgpu_launch_count (no reference counterpart): number of kernels launched by this context since creation (bench.py's gpu_launches) */
int64_t wgpu_launch_count(const wgpu_ctx *ctx);
/* wgpu_device_bytes: bytes of device memory held by the context (the reference prints its heavy-data footprint in allocate_forest.f90:228-267) */
int64_t wgpu_device_bytes(const wgpu_ctx *ctx);
/* wgpu_device_pointer (no reference counterpart): raw device pointer + element count of a resident array (for zero-copy interop, e.g. NCCL
 * via torch) */
int32_t wgpu_device_pointer(wgpu_ctx *ctx, int32_t array_id, int32_t slot, void **ptr, int64_t *n_doubles);

#ifdef __cplusplus
}
#endif
#endif /* WABBIT_GPU_H */
