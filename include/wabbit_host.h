/*
 * wabbit_host.h -- C ABI of the host-side forest metadata used by the drivers, tests and benchmark
 * of this repository (libwabbit_host.so, plain C++17, no CUDA).
 *
 * In a WABBIT build the octree / light data stay in host Fortran (LIB/TREE, LIB/MESH) and only their
 * RESULT -- hvy_active, block levels and the 168-slot hvy_neighbor table -- crosses the boundary
 * through wgpu_set_topology().  No Fortran compiler exists in this image, so this library produces
 * the same tables for the grids the drivers need (equidistant grids, and grids given as an explicit
 * list of leaf blocks), with WABBIT's conventions:
 *   - binary treecode, coarsest digit in the highest bits, digit bit0 -> y, bit1 -> x, bit2 -> z
 *     (LIB/TREE/module_treelib.f90:793-871),
 *   - neighbour slots 1-56 same level, 57-112 coarser, 113-168 finer; faces 1-24, edges 25-48,
 *     corners 49-56 (LIB/TREE/neighborhood.f90:10-22, LIB/MESH/find_neighbors.f90:18-180),
 *   - lgt_id = rank*number_blocks + hvy_id, 1-based (LIB/MESH/hvy2lgt.f90, lgt2proc.f90),
 *   - blocks sorted along a space-filling curve (Z or Hilbert) and cut into contiguous chunks,
 *     one per rank (LIB/MESH/balanceLoad_tree.f90:203-285, 600-715).
 */
#ifndef WABBIT_HOST_H
#define WABBIT_HOST_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct whost_forest whost_forest;

enum { WHOST_SFC_Z = 0, WHOST_SFC_HILBERT = 1 };

/* Equidistant grid on level J (createEquidistantGrid_tree), distributed over n_ranks by the SFC. */
int32_t whost_create_uniform(int32_t dim, int32_t J, int32_t Jmax, int32_t sfc, int32_t n_ranks, int32_t max_blocks_per_rank,
                             const int32_t periodic[3], whost_forest **out);
/* Grid from an explicit list of leaf blocks: level[n], ixyz[3*n] zero-based block coordinates on their level. */
int32_t whost_create_from_blocks(int32_t dim, int32_t Jmax, int32_t sfc, int32_t n_ranks, int32_t max_blocks_per_rank,
                                 const int32_t periodic[3], int32_t n, const int32_t *level, const int32_t *ixyz, whost_forest **out);
int32_t whost_destroy(whost_forest *f);

int32_t whost_n_blocks(const whost_forest *f);                 /* lgt_n */
int32_t whost_n_active(const whost_forest *f, int32_t rank);   /* hvy_n of a rank */
/* per rank, in SFC order: hvy ids (1-based), level, zero-based block coordinates, treecode */
int32_t whost_get_active(const whost_forest *f, int32_t rank, int32_t *hvy_active, int32_t *level, int32_t *ixyz, int64_t *treecode);
/* hvy_neighbor(ld, 168) of a rank, ld = whost_n_active(f, rank) (hvy ids are 1..ld), Fortran column-major, lgt ids
 * (rank*max_blocks_per_rank + hvy), -1 = none */
int32_t whost_get_neighbors(const whost_forest *f, int32_t rank, int32_t *hvy_neighbor);
/* the same table without a copy: pointer to the forest's own storage (valid until whost_destroy), or NULL */
const int32_t *whost_neighbors_ptr(const whost_forest *f, int32_t rank);
/* 1 if every neighbour relation of every block is same-level */
int32_t whost_is_uniform(const whost_forest *f);

/* Halo plan of a rank from the block positions (no hvy_neighbor table needed): the blocks of other ranks that appear in a neighbour relation
 * (same level / finer / coarser, find_neighbor's cases) of the rank's blocks, ascending lgt id, with level, treecode and a flag for finer
 * neighbours (whose HD-filtered copies travel as well, restrict_copy_at_CE); the own blocks (hvy ids) the peers mirror, peer-major and hvy
 * ascending, and those among them that are finer neighbours of a peer's block.  recv_counts / send_counts / fine_send_counts: n_ranks entries.
 * The reference derives the same per synchronisation in prepare_ghost_synch_metadata (LIB/MPI/synchronize_ghosts_generic.f90:352-694). */
int32_t whost_halo_plan(const whost_forest *f, int32_t rank, int32_t *n_halo, int32_t *halo_lgt, int32_t *halo_level, int64_t *halo_tc,
                        int32_t *halo_fine, int32_t *recv_counts, int32_t *n_send, int32_t *send_hvy, int32_t *send_counts, int32_t *n_fine_send,
                        int32_t *fine_send_hvy, int32_t *fine_send_counts);

/*
 * Grid adaptation, light data only, single rank (n_ranks == 1).  Stand-ins for the id bookkeeping of refinement_execute_tree +
 * balanceLoad_tree and for respectJmaxJmin_tree / completeness / ensureGradedness_tree; they produce the id lists that
 * wgpu_refine / wgpu_move_blocks / wgpu_coarsen consume (all hvy ids 1-based).
 *   whost_refine : flags[n] > 0 (NULL = everywhere) marks blocks to refine (blocks on Jmax are skipped).  mothers[nm] are old ids,
 *                  daughters[nm*2^dim] new ids in digit order, keep_src/keep_dst[nk] old -> new ids of the other blocks.
 *   whost_coarsen: status[n] in: -1 = wants to coarsen; out: final status (-1 only for complete sister groups above Jmin whose
 *                  finer neighbours coarsen too).  mothers[nm] are new ids, daughters[nm*2^dim] OLD ids in digit order.
 * Output arrays must hold n entries (daughters: n); returns 2 if the new grid needs more than max_blocks blocks.
 */
int32_t whost_refine(const whost_forest *f, const int32_t *flags, int32_t max_blocks, whost_forest **out, int32_t *n_mothers, int32_t *mothers,
                     int32_t *daughters, int32_t *n_keep, int32_t *keep_src, int32_t *keep_dst);
int32_t whost_coarsen(const whost_forest *f, int32_t *status, int32_t Jmin, int32_t max_blocks, whost_forest **out, int32_t *n_mothers,
                      int32_t *mothers, int32_t *daughters, int32_t *n_keep, int32_t *keep_src, int32_t *keep_dst);

/* The same for a grid partitioned over any number of ranks: flags / status and every id are 1-based positions in the GLOBAL
 * space-filling-curve order (= rank-major order of the active lists) of the old resp. new grid; the new grid is partitioned over the
 * same ranks, at most max_blocks_per_rank blocks each (else 2).  The caller turns positions into (rank, hvy id) with the ranks' counts. */
int32_t whost_refine_global(const whost_forest *f, const int32_t *flags, int32_t max_blocks_per_rank, whost_forest **out, int32_t *n_mothers,
                            int32_t *mothers, int32_t *daughters, int32_t *n_keep, int32_t *keep_src, int32_t *keep_dst);
int32_t whost_coarsen_global(const whost_forest *f, int32_t *status, int32_t Jmin, int32_t max_blocks_per_rank, whost_forest **out,
                             int32_t *n_mothers, int32_t *mothers, int32_t *daughters, int32_t *n_keep, int32_t *keep_src, int32_t *keep_dst);

/* Full tree (leaves + all ancestors) of adapt_tree's full wavelet transformation (LIB/MESH/adapt_tree.f90:268-545, init_full_tree): blocks as
 * (level[n], pos[3n]) in any order.  whost_ft_tables: same-level neighbour per direction nb[n][3^dim-1] (dz, dy, dx ascending, periodic), mother
 * par[n], daughters child[n][2^dim] (column = x + 2y + 4z offset); indices into the list or -1.  whost_ft_decide: respectJmaxJmin_tree +
 * ensureGradedness_tree(check_daughters) (LIB/MESH/ensureGradedness_tree.f90): status -1 survives only for blocks that are deleted. */
/* whost_ft_build: init_full_tree's light-data half -- the leaves plus all their ancestors down to Jmin, sorted by position code
 * (level << 57 | z << 38 | y << 19 | x); leaf_of[i] = index into the input list, -1 for a mother.  Returns 2 if cap is too small.
 * whost_encode_many: numerical treecodes (encoding_b) of a list of blocks. */
int32_t whost_ft_build(int32_t dim, int32_t Jmin, int32_t n_leaf, const int32_t *level, const int32_t *pos, int32_t cap, int32_t *n_out,
                       int32_t *level_out, int32_t *pos_out, int32_t *leaf_of);
int32_t whost_encode_many(int32_t dim, int32_t Jmax, int32_t n, const int32_t *level, const int32_t *pos, int64_t *tc);
int32_t whost_ft_tables(int32_t dim, int32_t n, const int32_t *level, const int32_t *pos, int32_t *nb, int32_t *par, int32_t *child);
/* hvy_neighbor rows [168][ld] (column = slot - 1) of every block of the full tree for the passes of the full-tree adapt_tree: same-level
 * relations from nb (row dir_code[q] - 1, dir_code = the slot code of find_neighbor for direction q); for blocks flagged in `leaf`,
 * directions without a same-level block get the coarser relation (row + 56) to the covering block one level up.  Unset entries keep
 * their value (preset -1). */
int32_t whost_ft_rows(int32_t dim, int32_t n, const int32_t *level, const int32_t *pos, const int32_t *nb, const int32_t *slots,
                      const int32_t *leaf, const int32_t *dir_code, int64_t ld, int32_t *rows);
int32_t whost_ft_decide(int32_t dim, int32_t n, const int32_t *level, const int32_t *nb, const int32_t *par, const int32_t *child, int32_t Jmin,
                        int32_t *status);
/* the (significant block, direction) pairs addSecurityZone_CE_tree examines (LIB/MESH/securityZone_tree.f90:140-298): sig[i] != 0 and the
 * same-level neighbour nb[i][q] exists with status -1.  Direction-major, blocks ascending; returns the number of pairs (-2: cap too small). */
int32_t whost_ft_security_pairs(int32_t dim, int32_t n, const int32_t *nb, const int32_t *status, const uint8_t *sig, int32_t cap, int32_t *blk,
                                int32_t *dir);

/* treecode helpers (module_treelib.f90:793-871) */
int64_t whost_encode(int32_t dim, int32_t level, int32_t Jmax, const int32_t ixyz[3]);
int32_t whost_decode(int32_t dim, int32_t level, int32_t Jmax, int64_t treecode, int32_t ixyz[3]);
/* position along the space-filling curve of a block (sort key) */
uint64_t whost_sfc_key(int32_t dim, int32_t sfc, int32_t level, int32_t Jmax, const int32_t ixyz[3]);

#ifdef __cplusplus
}
#endif
#endif
